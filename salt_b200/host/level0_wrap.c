/*
 * level0_wrap.c -- linked INTO a program that uses libsalt_level0.so, together with -Wl,--wrap=mixRef_restore:
 * the reference loads its SNP-aware reference exactly as always (mixRef_restore, metaref.c:61-93), and the copy it
 * loaded is uploaded to the GPU once, before the first ed_mismatch / ed_diff call needs it.  (The wrapper has to be part
 * of the program's own link step: --wrap rewrites references in the objects being linked, not in a shared library.)
 */
#include <stdint.h>
#include <stdlib.h>

typedef struct { uint32_t *seq; uint32_t l; } mixref_view_t;          /* mixRef_t, metaref.h:2-5 */
mixref_view_t *__real_mixRef_restore(const char *fn);
int salt_level0_attach_words(const uint32_t *words, uint32_t l);

mixref_view_t *__wrap_mixRef_restore(const char *fn)
{
    mixref_view_t *m = __real_mixRef_restore(fn);
    if (m && salt_level0_attach_words(m->seq, m->l) != 0) exit(1);
    return m;
}
