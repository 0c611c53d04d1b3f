// engine.cu -- the C ABI of libsalt_b200.so (include/salt_b200.h): device residency of the
// reference and the current read chunk, staging of batches, and the verification pipeline.
// There is deliberately no CPU implementation behind any entry point.
#if !defined(SALT_EMUL)
#include <cuda_runtime.h>
#endif

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"

using namespace salt;

namespace {

thread_local std::string g_err;

int fail(int code, const char *what, cudaError_t e = cudaSuccess)
{
    char buf[512];
    if (e != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    else snprintf(buf, sizeof buf, "%s", what);
    g_err = buf;
    return code;
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(SALT_ERR_CUDA, #call, e_); } while (0)

// growable device buffer
struct DBuf {
    void *p = nullptr; size_t cap = 0;
    cudaError_t need(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaError_t e = cudaFree(p); p = nullptr; cap = 0; if (e != cudaSuccess) return e; }
        size_t want = bytes + bytes / 8 + 4096;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) return e;
        cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

}  // namespace

// One in-flight read chunk: its own stream, its own copy of the chunk's reads and candidate
// lists in HBM, and the scratch of the verification stage.  Slot 0 also serves the synchronous
// entry points; the other slots exist so that chunk c+1 can upload while chunk c computes and
// chunk c-1 downloads (salt_b200_verify_submit / _wait).
struct Slot {
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    DBuf codes, offs3, rd4, rd_len;                           // offs3: read offsets, then the two candidate-offset arrays
    uint32_t n_reads = 0, W64 = 0, l_max = 0;
    bool offs_merged = false;                                 // all three offset arrays arrived in one copy
    uint32_t *d_roffs() const { return offs3.as<uint32_t>(); }
    uint32_t *d_coffs(int strand) const { return offs3.as<uint32_t>() + (size_t)(strand + 1) * ((size_t)n_reads + 1); }
    DBuf c_loci0, c_loci1, vpairs, acc, rec, lvlist, ciglist, counters, cig, fpairs, fslots, lvreads, fpairs2, fslots2;
    DBuf pk_bases, pk_lens, pk_cnt, pk_npos, pk_scan;          // compact transport (salt_packed_chunk_t): raw uploads + scan scratch
    DBuf tl_md, tl_out, tl_len, tl_offs, tl_packed, tl_cigrow, tl_xv, tl_scan;   // SAM tail of the chunk's primaries
    bool tail_pending = false; size_t tail_eager = 0, tail_cap = 0; uint32_t *tail_offs = nullptr; char *tail_md = nullptr;
    bool have_rec = false; int rec_cig_stride = 0;                // the slot's rec / cig buffers hold a finished verify
    DBuf cigslim;                                               // 32-byte copies of the CIGAR rows that go back eagerly
    DBuf sd_status;                                             // per strand: flags of the paired-end flavour of locate
    DBuf sd_long;                                               // (read, strand) ids whose lists need the long sort, [0] = count
    DBuf sd_sai, sd_counts, sd_lists;                          // seeding scratch: intervals, per-strand counts, fixed-stride lists
    uint32_t *h_tot = nullptr;                                  // pinned: the two list totals of a seeded chunk
    size_t seeded_n0 = 0, seeded_n1 = 0; int seeded = 0;        // 1: totals in flight, 2: lists gathered into c_loci0/1
    int seed_max_locate = 0;
    // asynchronous verify in flight: where the compact CIGAR list has to be scattered to
    bool pending = false;
    char *u_cigars = nullptr; int u_stride = 0;
    uint8_t *h_stage = nullptr; size_t h_stage_cap = 0;      // pinned: count, read ids, strings
    uint32_t eager = 0;                                       // list entries already copied by submit
    uint32_t *h_offs = nullptr; size_t h_offs_cap = 0;        // pinned; salt_b200_verify_batch: the chunk's rebased offsets
    // optional per-stage timing (salt_b200_profile): events at the boundaries of the stages
    cudaEvent_t ev_verify[7] = {nullptr};
    bool have_verify_prof = false;

    void release()
    {
        DBuf *all[] = {&codes, &offs3, &rd4, &rd_len, &c_loci0, &c_loci1, &vpairs, &acc, &rec,
                       &lvlist, &ciglist, &counters, &cig, &fpairs, &fslots, &lvreads, &fpairs2, &fslots2,
                       &pk_bases, &pk_lens, &pk_cnt, &pk_npos, &pk_scan, &sd_sai, &sd_counts, &sd_lists, &sd_long, &sd_status, &cigslim,
                       &tl_md, &tl_out, &tl_len, &tl_offs, &tl_packed, &tl_cigrow, &tl_xv, &tl_scan};
        for (DBuf *b : all) b->release();
        if (h_stage) cudaFreeHost(h_stage);
        h_stage = nullptr; h_stage_cap = 0;
        if (h_offs) cudaFreeHost(h_offs);
        h_offs = nullptr; h_offs_cap = 0;
        if (h_tot) cudaFreeHost(h_tot);
        h_tot = nullptr;
        for (int i = 0; i < 7; ++i) if (ev_verify[i]) { cudaEventDestroy(ev_verify[i]); ev_verify[i] = nullptr; }
        if (stream && own_stream) cudaStreamDestroy(stream);
        stream = nullptr;
    }
};

constexpr int SALT_SLOTS = 4;
constexpr uint32_t CIG_EAGER = 8192;       // CIGARs of a chunk copied back without waiting for their count

struct salt_b200 {
    int device = 0;
    int sm_count = 148;
    uint8_t *d_mixref_alloc = nullptr;      // allocation: REF_FRONT zero bytes, the words, REF_PAD zero bytes
    uint32_t *d_mixref = nullptr; uint32_t l = 0;
    uint8_t *d_pac = nullptr; int64_t l_pac = 0;
    size_t ref_cap = 0, pac_cap = 0;        // bytes allocated for the words / the pac (salt_b200_reload_ref)
    bool borrowed = false;                  // reference (and index) belong to the handle this one was attached to
    Slot slot[SALT_SLOTS];
    DBuf fpairs, fslots, fcount;            // LV filter survivors of the per-pair entry point
    int lv_filter = 1;                      // pigeonhole filter in front of Landau-Vishkin (salt_b200_set_lv_filter)
    int lv_two_pass = 1;                    // Landau-Vishkin in two passes behind the filter (salt_b200_set_lv_mapping bit 4 clears it)
    DBuf fpairs2, fslots2;                  // second-pass list of the per-pair entry point
    // staging / scratch of the synchronous per-pair and SSW entry points (slot 0's stream)
    DBuf pairs, out8, kbuf, cig, wins, sswout, sswcig, sswscratch, sswovf, md_in, md_cig, md_str, md_xv, md_out;
    DBuf ix_cbwt, ix_csa, ix_lkt, ix_rbwt, ix_rocc, ix_rmaj, ix_rsa;      // FM-indexes (salt_b200_set_index)
    DBuf ix_c32, ix_r64;                                                  // ... and the device's own dense layouts built from them
    FmIndexDev fm{}; bool have_index = false;
    uint64_t launches = 0;
    int cur = 0;                // slot whose reads the per-pair / SSW entry points work on (salt_b200_use_slot)
    int lv_mapping = 0;         // 0 = auto, 1 = warp per pair, 2 = thread per pair (salt_b200_set_lv_mapping)
    int max_window = 1024;      // widest rescue window the SSW scratch is sized for
    bool profiling = false;
    cudaEvent_t ev_ssw[7] = {nullptr};
    bool have_ssw_prof = false;

    DevCtx ctx(int si = 0) const
    {
        const Slot &s = slot[si];
        DevCtx c;
        c.mixref = d_mixref; c.l = l; c.pac = d_pac; c.l_pac = l_pac;
        c.rd4 = s.rd4.as<uint64_t>(); c.rd_len = s.rd_len.as<uint16_t>();
        c.n_reads = s.n_reads; c.W64 = s.W64; c.l_max = s.l_max;
        return c;
    }
};

namespace {

int use_device(salt_b200_t *h)
{
    if (!h) return fail(SALT_ERR_ARG, "null handle");
    CU(cudaSetDevice(h->device));
    return SALT_OK;
}

salt_b200_t *new_handle(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) { fail(SALT_ERR_NODEVICE, "no CUDA device available (libsalt_b200 has no CPU path)", e); return nullptr; }
    if (device < 0 || device >= n) { fail(SALT_ERR_ARG, "device index out of range"); return nullptr; }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { fail(SALT_ERR_CUDA, "cudaSetDevice", e); return nullptr; }
    salt_b200_t *h = new (std::nothrow) salt_b200();
    if (!h) { fail(SALT_ERR_NOMEM, "out of host memory"); return nullptr; }
    h->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
    for (int i = 0; i < SALT_SLOTS; ++i)
        if ((e = cudaStreamCreateWithFlags(&h->slot[i].stream, cudaStreamNonBlocking)) != cudaSuccess) {
            fail(SALT_ERR_CUDA, "cudaStreamCreate", e); salt_b200_destroy(h); return nullptr;
        }
    return h;
}

const size_t REF_PAD = 1024;  // zero bytes after the reference so vector loads may run past the end
const size_t REF_FRONT = 256; // zero bytes before it: the LV filter looks 32 bases to the left of a window

// device allocation of the reference words with zeroed padding on both sides
cudaError_t alloc_mixref(salt_b200_t *h, size_t nb, cudaStream_t st)
{
    cudaError_t e = cudaMalloc(&h->d_mixref_alloc, REF_FRONT + nb + REF_PAD);
    if (e != cudaSuccess) return e;
    h->d_mixref = reinterpret_cast<uint32_t *>(h->d_mixref_alloc + REF_FRONT);
    h->ref_cap = nb;
    if ((e = cudaMemsetAsync(h->d_mixref_alloc, 0, REF_FRONT, st)) != cudaSuccess) return e;
    return cudaMemsetAsync(h->d_mixref_alloc + REF_FRONT + nb, 0, REF_PAD, st);
}

// Upload + pack one chunk of reads into a slot.  Asynchronous on the slot's stream.
// `with` non-null: if the caller keeps the read offsets and both candidate-offset arrays back to back in host
// memory (salt_b200_verify_batch and the host layer's chunk queues do), all three travel in one copy.
int load_reads(salt_b200_t *h, Slot &s, const salt_reads_t *reads, const salt_cands_t *with = nullptr)
{
    if (!reads || !reads->offs || (reads->n_reads && !reads->codes)) return fail(SALT_ERR_ARG, "reads is null");
    if (reads->n_reads >= (1u << 31)) return fail(SALT_ERR_ARG, "too many reads in one chunk");
    const uint32_t n = reads->n_reads;
    uint32_t l_max = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t a = reads->offs[i], b = reads->offs[i + 1];
        if (b < a) return fail(SALT_ERR_ARG, "read offsets not monotone");
        if (b - a > l_max) l_max = b - a;
    }
    if (l_max > 1024) return fail(SALT_ERR_UNSUPPORTED, "reads longer than 1024 bases are not supported");
    const size_t total = n ? (size_t)reads->offs[n] - reads->offs[0] : 0;
    if (n && reads->offs[0] != 0) return fail(SALT_ERR_ARG, "read offsets must start at 0");
    s.n_reads = n; s.l_max = l_max; s.W64 = (l_max + 15) / 16 + 1; s.have_rec = false; s.seeded = 0;      // lists seeded from the reads that were here are gone with them
    CU(s.codes.need(total + 16));
    CU(s.offs3.need(3 * ((size_t)n + 1) * 4));
    CU(s.rd4.need((size_t)n * 2 * s.W64 * 8 + 64));
    CU(s.rd_len.need((size_t)n * 2 + 64));
    if (n) {
        CU(cudaMemcpyAsync(s.codes.p, reads->codes, total, cudaMemcpyHostToDevice, s.stream));
        s.offs_merged = with && with->offs[0] == reads->offs + ((size_t)n + 1) && with->offs[1] == with->offs[0] + ((size_t)n + 1);
        CU(cudaMemcpyAsync(s.offs3.p, reads->offs, (s.offs_merged ? 3 : 1) * ((size_t)n + 1) * 4, cudaMemcpyHostToDevice, s.stream));
        CU(launch_pack_reads(s.codes.as<uint8_t>(), s.d_roffs(), n, s.W64, s.rd4.as<uint64_t>(),
                             s.rd_len.as<uint16_t>(), s.stream));
        h->launches += 1;
    }
    return SALT_OK;
}

// The verification stage of one slot with every input already in HBM (asynchronous).
int verify_on_device(salt_b200_t *h, int si, const uint32_t *d_offs0, const uint32_t *d_loci0, size_t n0,
                     const uint32_t *d_offs1, const uint32_t *d_loci1, size_t n1,
                     int nogap_T0, int lv_T0, salt_verify_out_t *d_rec, int8_t *d_acc0, int8_t *d_acc1,
                     char *d_cigars, int cigar_stride, uint32_t *d_cig_reads, uint32_t *d_cig_count)
{
    Slot &s = h->slot[si];
    if (!s.n_reads) return fail(SALT_ERR_ARG, "no reads set");
    if (!d_offs0 || !d_offs1 || !d_rec) return fail(SALT_ERR_ARG, "null buffer");
    if (nogap_T0 < 0 || nogap_T0 > 127) return fail(SALT_ERR_ARG, "nogap_T0 must be in 0..127");
    if (n0 + n1 >= (size_t)1 << 32) return fail(SALT_ERR_ARG, "too many candidates in one chunk");
    if (d_acc0 && d_acc1 && d_acc1 != d_acc0 + n0) return fail(SALT_ERR_ARG, "acc1 must follow acc0 contiguously");
    if (d_cigars && (!d_cig_reads || !d_cig_count)) return fail(SALT_ERR_ARG, "cigars need d_cig_reads and d_cig_count");
    if (d_cigars && cigar_stride < 2) return fail(SALT_ERR_ARG, "cigar stride too small");
    if (reinterpret_cast<uintptr_t>(d_rec) & 15u) return fail(SALT_ERR_ARG, "d_rec must be 16-byte aligned");
    const size_t n = n0 + n1;
    const DevCtx c = h->ctx(si);
    CU(s.vpairs.need((n + 1) * sizeof(salt_pair_t)));      // LV worklist: pairs ...
    CU(s.lvlist.need((n + 1) * 4));                         // ... and the acc slot each one reports to
    CU(s.counters.need(256));
    CU(s.lvreads.need(((size_t)s.n_reads + 1) * 4));        // reads that reach the gapped stage
    LvFilterScratch fs{nullptr, nullptr, nullptr};
    if (h->lv_filter) {
        CU(s.fpairs.need((n + 1) * sizeof(salt_pair_t)));
        CU(s.fslots.need((n + 1) * 4));
        fs.pairs = s.fpairs.as<salt_pair_t>(); fs.slots = s.fslots.as<uint32_t>(); fs.count = s.counters.as<uint32_t>() + 8;
        if (h->lv_two_pass) {
            CU(s.fpairs2.need((n + 1) * sizeof(salt_pair_t)));
            CU(s.fslots2.need((n + 1) * 4));
            fs.pairs2 = s.fpairs2.as<salt_pair_t>(); fs.slots2 = s.fslots2.as<uint32_t>(); fs.count2 = s.counters.as<uint32_t>() + 12;
        }
    }
    int8_t *acc = d_acc0;
    if (!acc) { CU(s.acc.need(n + 1)); acc = s.acc.as<int8_t>(); }
    salt_pair_t *lvp = s.vpairs.as<salt_pair_t>();
    uint32_t *lvs = s.lvlist.as<uint32_t>();
    uint32_t *cnt = s.counters.as<uint32_t>();        // [0] LV worklist length
    if (h->profiling)
        for (int i = 0; i < 7; ++i) if (!s.ev_verify[i]) CU(cudaEventCreate(&s.ev_verify[i]));
    cudaEvent_t *ev = h->profiling ? s.ev_verify : nullptr;
#define SALT_EV(i) do { if (ev) CU(cudaEventRecord(ev[i], s.stream)); } while (0)
    CU(cudaMemsetAsync(cnt, 0, 256, s.stream));
    if (d_cig_count) CU(cudaMemsetAsync(d_cig_count, 0, 4, s.stream));
    SALT_EV(0);
    SALT_EV(1);
    CU(launch_nogap_fused(c, d_offs0, d_loci0, d_offs1, d_loci1, n0, nogap_T0, acc, d_rec, lvp, lvs, cnt,
                          s.lvreads.as<uint32_t>(), s.stream));
    SALT_EV(2);
    SALT_EV(3);
    CU(launch_lv(c, lvp, n, lv_T0, lvs, cnt, n, acc, h->sm_count, s.stream, h->lv_mapping, h->lv_filter ? &fs : nullptr));
    SALT_EV(4);
    CU(launch_scan_gap(c, d_offs0, d_loci0, d_offs1, d_loci1, n0, lv_T0, acc, d_rec, s.lvreads.as<uint32_t>(), cnt + 1,
                       d_cigars ? d_cig_reads : nullptr, d_cig_count, h->sm_count, s.stream));
    SALT_EV(5);
    h->launches += h->lv_filter ? 4 : 3;
    if (d_cigars) {
        const int kmax = lv_T0 >= 0 ? lv_T0 : (int)s.l_max / 10;     // a gapped primary's n_diff never exceeds the stage threshold
        CU(launch_lv_cigar(c, nullptr, nullptr, 0, d_cig_reads, d_cig_count, s.n_reads, d_rec, kmax,
                           d_cigars, cigar_stride, nullptr, h->sm_count, s.stream, h->lv_mapping));
        h->launches += 1;
    }
    SALT_EV(6);
#undef SALT_EV
    s.have_verify_prof = h->profiling;
    return SALT_OK;
}

// Second half of a host-buffer verify: the slot's candidate offsets and loci are in HBM (n0 / n1 of them);
// run the stage and queue the downloads.  Returns without waiting; finish_verify completes it.
int run_and_download(salt_b200_t *h, int si, size_t n0, size_t n1, int nogap_T0, int lv_T0,
                     salt_verify_out_t *rec, int8_t *acc0, int8_t *acc1, char *cigars, int cigar_stride)
{
    Slot &s = h->slot[si];
    const uint32_t nr = s.n_reads;
    CU(s.acc.need(n0 + n1 + 1));
    CU(s.rec.need((size_t)nr * sizeof(salt_verify_out_t)));
    s.eager = 0;
    if (cigars) {
        if (cigar_stride < 2) return fail(SALT_ERR_ARG, "cigar stride too small");
        CU(s.cig.need((size_t)nr * cigar_stride));
        CU(s.ciglist.need(((size_t)nr + 2) * 4));      // [0] count, [1..] read ids
        s.eager = nr < CIG_EAGER ? nr : CIG_EAGER;
        const size_t want = 4 + (size_t)s.eager * 4 + (size_t)s.eager * 32;
        CU(s.cigslim.need((size_t)s.eager * 32 + 32));
        if (want > s.h_stage_cap) {
            if (s.h_stage) { CU(cudaFreeHost(s.h_stage)); s.h_stage = nullptr; s.h_stage_cap = 0; }
            CU(cudaMallocHost(&s.h_stage, want));
            s.h_stage_cap = want;
        }
    }
    int8_t *acc = s.acc.as<int8_t>();
    uint32_t *cl = cigars ? s.ciglist.as<uint32_t>() : nullptr;
    if (int rc = verify_on_device(h, si, s.d_coffs(0), s.c_loci0.as<uint32_t>(), n0,
                                  s.d_coffs(1), s.c_loci1.as<uint32_t>(), n1, nogap_T0, lv_T0,
                                  s.rec.as<salt_verify_out_t>(), acc, acc + n0,
                                  cigars ? s.cig.as<char>() : nullptr, cigar_stride, cl ? cl + 1 : nullptr, cl))
        return rc;
    CU(cudaMemcpyAsync(rec, s.rec.p, (size_t)nr * sizeof(salt_verify_out_t), cudaMemcpyDeviceToHost, s.stream));
    if (acc0 && n0) CU(cudaMemcpyAsync(acc0, acc, n0, cudaMemcpyDeviceToHost, s.stream));
    if (acc1 && n1) CU(cudaMemcpyAsync(acc1, acc + n0, n1, cudaMemcpyDeviceToHost, s.stream));
    if (cigars) {
        // only gapped primaries have a CIGAR (query_gen_cigar, query.c:282-295): the device keeps a
        // compact list; its head comes back with the rest of the results, the tail (rare) on demand
        CU(cudaMemcpyAsync(s.h_stage, cl, 4 + (size_t)s.eager * 4, cudaMemcpyDeviceToHost, s.stream));
        if (s.eager) {
            CU(launch_cig_slim(s.cig.as<char>(), cigar_stride, cl, s.eager, s.cigslim.as<char>(), s.stream));
            h->launches += 1;
            CU(cudaMemcpyAsync(s.h_stage + 4 + (size_t)s.eager * 4, s.cigslim.p, (size_t)s.eager * 32, cudaMemcpyDeviceToHost, s.stream));
        }
    }
    s.u_cigars = cigars; s.u_stride = cigar_stride;
    s.pending = true;
    s.have_rec = true; s.rec_cig_stride = cigars ? cigar_stride : 0;
    return SALT_OK;
}

// Host-buffer verify of one slot: uploads the candidate lists, runs the stage, queues the
// downloads.  Returns without waiting; finish_verify completes it.
int enqueue_verify(salt_b200_t *h, int si, const salt_cands_t *cands, int nogap_T0, int lv_T0,
                   salt_verify_out_t *rec, int8_t *acc0, int8_t *acc1, char *cigars, int cigar_stride)
{
    Slot &s = h->slot[si];
    if (!cands || !cands->offs[0] || !cands->offs[1] || !rec) return fail(SALT_ERR_ARG, "null buffer");
    if (!s.n_reads) return fail(SALT_ERR_ARG, "no reads set");
    if (s.pending) return fail(SALT_ERR_ARG, "slot has a verify in flight: call salt_b200_verify_wait first");
    const uint32_t nr = s.n_reads;
    const size_t n0 = cands->offs[0][nr], n1 = cands->offs[1][nr];
    if ((n0 && !cands->loci[0]) || (n1 && !cands->loci[1])) return fail(SALT_ERR_ARG, "null loci");
    CU(s.c_loci0.need(n0 * 4 + 4)); CU(s.c_loci1.need(n1 * 4 + 4));
    if (!s.offs_merged) {
        CU(cudaMemcpyAsync(s.d_coffs(0), cands->offs[0], ((size_t)nr + 1) * 4, cudaMemcpyHostToDevice, s.stream));
        CU(cudaMemcpyAsync(s.d_coffs(1), cands->offs[1], ((size_t)nr + 1) * 4, cudaMemcpyHostToDevice, s.stream));
    }
    s.offs_merged = false;                               // one use: a later verify on the same reads uploads its own lists
    if (n0) CU(cudaMemcpyAsync(s.c_loci0.p, cands->loci[0], n0 * 4, cudaMemcpyHostToDevice, s.stream));
    if (n1) CU(cudaMemcpyAsync(s.c_loci1.p, cands->loci[1], n1 * 4, cudaMemcpyHostToDevice, s.stream));
    return run_and_download(h, si, n0, n1, nogap_T0, lv_T0, rec, acc0, acc1, cigars, cigar_stride);
}

// One chunk in the compact transport format: what is known on the host about it.
struct PackedView {
    const salt_packed_chunk_t *pc;
    uint32_t first = 0, n = 0;            // reads [first, first + n) of pc
    uint64_t base_pos = 0;                // stream position of the view's first base
    uint64_t n_bases = 0;
    size_t np_lo = 0, np_hi = 0;          // range of pc->n_pos inside the view
    size_t c0 = 0, c1 = 0, n0 = 0, n1 = 0;   // where the view's loci start in pc->loci[] and how many there are
    uint32_t l_max = 0;
};

inline uint32_t packed_count(const salt_packed_chunk_t *pc, int strand, uint32_t i)
{
    return pc->count_bits == 16 ? (uint32_t)static_cast<const uint16_t *>(pc->n_cand[strand])[i]
                                : static_cast<const uint32_t *>(pc->n_cand[strand])[i];
}

int check_packed(const salt_packed_chunk_t *pc, bool need_cands)
{
    if (!pc) return fail(SALT_ERR_ARG, "packed chunk is null");
    if (pc->base_bits != 2 && pc->base_bits != 4) return fail(SALT_ERR_ARG, "base_bits must be 2 or 4");
    if (pc->n_reads && !pc->bases) return fail(SALT_ERR_ARG, "bases is null");
    if (pc->n_reads >= (1u << 31)) return fail(SALT_ERR_ARG, "too many reads");
    if (!pc->lens && (pc->l_seq == 0 || pc->l_seq > 1024)) return fail(SALT_ERR_ARG, "l_seq must be 1..1024 when lens is null");
    if (pc->n_n && (!pc->n_pos || pc->base_bits != 2)) return fail(SALT_ERR_ARG, "n_pos belongs to 2-bit streams");
    if (need_cands) {
        if (pc->count_bits != 16 && pc->count_bits != 32) return fail(SALT_ERR_ARG, "count_bits must be 16 or 32");
        if (pc->n_reads && (!pc->n_cand[0] || !pc->n_cand[1])) return fail(SALT_ERR_ARG, "n_cand is null");
    }
    return SALT_OK;
}

// Extend `v` (which ends at read v.first + v.n) by the next m reads of its chunk: host-side sums only.
int grow_view(PackedView &v, uint32_t m, bool with_cands)
{
    const salt_packed_chunk_t *pc = v.pc;
    for (uint32_t i = v.first + v.n; i < v.first + v.n + m; ++i) {
        const uint32_t L = pc->lens ? pc->lens[i] : pc->l_seq;
        if (L > v.l_max) v.l_max = L;
        v.n_bases += L;
        if (with_cands) { v.n0 += packed_count(pc, 0, i); v.n1 += packed_count(pc, 1, i); }
    }
    v.n += m;
    if (v.l_max > 1024) return fail(SALT_ERR_UNSUPPORTED, "reads longer than 1024 bases are not supported");
    if (v.n_bases >= ((uint64_t)1 << 32) || v.n0 + v.n1 >= ((size_t)1 << 32)) return fail(SALT_ERR_ARG, "chunk too large: bases and candidates are 32-bit offsets");
    const uint64_t end = v.base_pos + v.n_bases;
    v.np_hi = v.np_lo;
    while (v.np_hi < pc->n_n && pc->n_pos[v.np_hi] < end) ++v.np_hi;
    return SALT_OK;
}

// Upload one view and rebuild codes / offsets / packed reads in the slot.  Asynchronous on the slot's stream.
int load_packed(salt_b200_t *h, Slot &s, const PackedView &v, bool with_cands)
{
    const salt_packed_chunk_t *pc = v.pc;
    const uint32_t n = v.n;
    s.n_reads = n; s.l_max = v.l_max; s.W64 = (v.l_max + 15) / 16 + 1; s.have_rec = false; s.seeded = 0;
    s.offs_merged = false;
    if (!n) return SALT_OK;
    const uint32_t per = 8u / (uint32_t)pc->base_bits;                 // bases per byte
    const uint64_t byte_lo = v.base_pos / per, byte_hi = (v.base_pos + v.n_bases + per - 1) / per;
    const uint32_t phase = (uint32_t)(v.base_pos % per);
    CU(s.codes.need((v.n_bases + 3) / 4 * 4 + 16));
    CU(s.offs3.need(3 * ((size_t)n + 1) * 4));
    CU(s.rd4.need((size_t)n * 2 * s.W64 * 8 + 64));
    CU(s.rd_len.need((size_t)n * 2 + 64));
    CU(s.pk_bases.need(byte_hi - byte_lo + 8));
    CU(s.pk_scan.need(3 * (size_t)scan3_blocks(n) * 4 + 16));
    CU(cudaMemcpyAsync(s.pk_bases.p, pc->bases + byte_lo, byte_hi - byte_lo, cudaMemcpyHostToDevice, s.stream));
    Scan3 sc{};
    sc.n = n; sc.partial = s.pk_scan.as<uint32_t>();
    sc.out[0] = s.d_roffs(); sc.out[1] = s.d_coffs(0); sc.out[2] = s.d_coffs(1);
    if (pc->lens) {
        CU(s.pk_lens.need((size_t)n * 2 + 16));
        CU(cudaMemcpyAsync(s.pk_lens.p, pc->lens + v.first, (size_t)n * 2, cudaMemcpyHostToDevice, s.stream));
        sc.in[0] = s.pk_lens.p; sc.width[0] = 16;
    } else { sc.in[0] = nullptr; sc.uniform[0] = pc->l_seq; sc.width[0] = 32; }
    if (with_cands) {
        const size_t cb = (size_t)pc->count_bits / 8;
        const size_t stride = ((size_t)n * cb + 15) / 16 * 16;
        CU(s.pk_cnt.need(2 * stride + 16));
        for (int k = 0; k < 2; ++k) {
            uint8_t *dst = s.pk_cnt.as<uint8_t>() + k * stride;
            CU(cudaMemcpyAsync(dst, static_cast<const uint8_t *>(pc->n_cand[k]) + (size_t)v.first * cb, (size_t)n * cb,
                               cudaMemcpyHostToDevice, s.stream));
            sc.in[1 + k] = dst; sc.width[1 + k] = pc->count_bits;
        }
    }
    CU(launch_scan3(sc, with_cands ? 3 : 1, s.stream));
    const size_t nn = v.np_hi - v.np_lo;
    if (nn) {
        CU(s.pk_npos.need(nn * 4));
        CU(cudaMemcpyAsync(s.pk_npos.p, pc->n_pos + v.np_lo, nn * 4, cudaMemcpyHostToDevice, s.stream));
    }
    CU(launch_unpack_bases(s.pk_bases.as<uint8_t>(), phase, pc->base_bits, (size_t)v.n_bases, s.codes.as<uint8_t>(),
                           s.pk_npos.as<uint32_t>(), nn, (uint32_t)v.base_pos, s.stream));
    CU(launch_pack_reads(s.codes.as<uint8_t>(), s.d_roffs(), n, s.W64, s.rd4.as<uint64_t>(), s.rd_len.as<uint16_t>(), s.stream));
    h->launches += 5 + (nn ? 1 : 0);
    if (with_cands) {
        if ((v.n0 && !pc->loci[0]) || (v.n1 && !pc->loci[1])) return fail(SALT_ERR_ARG, "null loci");
        CU(s.c_loci0.need(v.n0 * 4 + 4)); CU(s.c_loci1.need(v.n1 * 4 + 4));
        if (v.n0) CU(cudaMemcpyAsync(s.c_loci0.p, pc->loci[0] + v.c0, v.n0 * 4, cudaMemcpyHostToDevice, s.stream));
        if (v.n1) CU(cudaMemcpyAsync(s.c_loci1.p, pc->loci[1] + v.c1, v.n1 * 4, cudaMemcpyHostToDevice, s.stream));
    }
    return SALT_OK;
}

// ---- seeding + locate (row f1) -----------------------------------------------------------------------
int check_seed_opt(const salt_b200_t *h, const Slot &s, const salt_seed_opt_t *o, int *max_seeds)
{
    if (!h->have_index) return fail(SALT_ERR_ARG, "no FM-index uploaded: call salt_b200_set_index first");
    if (!o) return fail(SALT_ERR_ARG, "seed options are null");
    if (o->l_seed < h->fm.l_lkt || o->l_seed > 1024) return fail(SALT_ERR_ARG, "l_seed must be at least the lookup length");
    if (o->l_overlap < 1) return fail(SALT_ERR_UNSUPPORTED, "l_overlap < 1: the reference's non-overlap seeding is not served");
    if (o->max_seed < 0) return fail(SALT_ERR_ARG, "max_seed must be >= 0");
    if (o->max_locate < 1 || o->max_locate > 16384) return fail(SALT_ERR_UNSUPPORTED, "max_locate must be in 1..16384");
    if (o->locate_mode != 0 && o->locate_mode != 1) return fail(SALT_ERR_ARG, "locate_mode must be 0 (alnse_locate_alt) or 1 (alnse_locate)");
    if (o->locate_mode == 1 && (o->list_cap < 64 || o->list_cap > 16384)) return fail(SALT_ERR_ARG, "list_cap must be in 64..16384 for locate_mode 1");
    const int ms = s.l_max >= (uint32_t)o->l_seed ? (int)((s.l_max - (uint32_t)o->l_seed) / (uint32_t)o->l_overlap) + 1 : 1;
    if (ms > 1024) return fail(SALT_ERR_UNSUPPORTED, "more than 1024 seed starts per strand");
    *max_seeds = ms;
    return SALT_OK;
}

// Phase A of a seeded chunk: seed, locate, scan the counts into the slot's candidate offsets, start the download of
// the two totals.  Asynchronous.
int seed_enqueue(salt_b200_t *h, int si, const salt_seed_opt_t *o)
{
    Slot &s = h->slot[si];
    int max_seeds = 1;
    if (int rc = check_seed_opt(h, s, o, &max_seeds)) return rc;
    const uint32_t n = s.n_reads;
    const int stride = o->locate_mode == 1 ? o->list_cap : o->max_locate;
    s.seeded = 0; s.seeded_n0 = s.seeded_n1 = 0; s.seed_max_locate = stride;
    if (!s.h_tot) CU(cudaMallocHost(&s.h_tot, 16));
    s.h_tot[0] = s.h_tot[1] = 0;
    if (!n) { s.seeded = 1; return SALT_OK; }
    SeedOpt so{o->l_seed, o->l_overlap, o->max_seed, o->max_locate, o->seed_only_ref, o->locate_mode, stride};
    CU(s.sd_sai.need(seed_sai_bytes(n, max_seeds)));
    CU(s.sd_counts.need((size_t)n * 2 * 4 + 16));
    CU(s.sd_lists.need((size_t)n * 2 * (size_t)stride * 4 + 16));
    CU(s.sd_status.need((size_t)n * 2 + 16));
    CU(s.pk_scan.need(3 * (size_t)scan3_blocks(n) * 4 + 16));
    CU(launch_seed(h->fm, so, s.codes.as<uint8_t>(), s.d_roffs(), n, max_seeds, s.sd_sai.as<SeedSai>(), s.stream));
    CU(s.sd_long.need(((size_t)n * 2 + 4) * 4));
    CU(launch_locate(h->fm, so, s.d_roffs(), n, max_seeds, h->l, s.sd_sai.as<SeedSai>(), s.sd_counts.as<uint32_t>(),
                     s.sd_lists.as<uint32_t>(), s.sd_long.as<uint32_t>() + 4, s.sd_long.as<uint32_t>(), s.sd_status.as<uint8_t>(), h->sm_count, s.stream));
    Scan3 sc{};
    sc.n = n; sc.partial = s.pk_scan.as<uint32_t>();
    sc.in[0] = s.sd_counts.as<uint32_t>(); sc.in[1] = s.sd_counts.as<uint32_t>() + n; sc.width[0] = sc.width[1] = 32;
    sc.out[0] = s.d_coffs(0); sc.out[1] = s.d_coffs(1);
    CU(launch_scan3(sc, 2, s.stream));
    CU(cudaMemcpyAsync(&s.h_tot[0], s.d_coffs(0) + n, 4, cudaMemcpyDeviceToHost, s.stream));
    CU(cudaMemcpyAsync(&s.h_tot[1], s.d_coffs(1) + n, 4, cudaMemcpyDeviceToHost, s.stream));
    h->launches += 6;
    s.offs_merged = false;
    s.seeded = 1;
    return SALT_OK;
}

// Phase B: wait for the totals, gather the fixed-stride lists into the slot's CSR candidate arrays.
int seed_finish(salt_b200_t *h, int si)
{
    Slot &s = h->slot[si];
    if (s.seeded != 1) return s.seeded == 2 ? SALT_OK : fail(SALT_ERR_ARG, "slot was not seeded");
    CU(cudaStreamSynchronize(s.stream));
    s.seeded_n0 = s.h_tot[0]; s.seeded_n1 = s.h_tot[1];
    CU(s.c_loci0.need(s.seeded_n0 * 4 + 4)); CU(s.c_loci1.need(s.seeded_n1 * 4 + 4));
    if (s.n_reads) {
        CU(launch_seed_gather(s.sd_lists.as<uint32_t>(), s.seed_max_locate, s.d_coffs(0), s.d_coffs(1), s.n_reads,
                              s.c_loci0.as<uint32_t>(), s.c_loci1.as<uint32_t>(), s.stream));
        h->launches += 1;
    }
    s.seeded = 2;
    return SALT_OK;
}

int finish_verify(salt_b200_t *h, int si)
{
    Slot &s = h->slot[si];
    if (!s.pending) return SALT_OK;
    s.pending = false;
    CU(cudaStreamSynchronize(s.stream));
    if (!s.u_cigars) return SALT_OK;
    uint32_t n_cig = 0;
    memcpy(&n_cig, s.h_stage, 4);
    if (n_cig > s.n_reads) return fail(SALT_ERR_CUDA, "corrupt CIGAR list length");
    const int stride = s.u_stride;
    const uint32_t head = n_cig < s.eager ? n_cig : s.eager;
    const uint32_t *ids = reinterpret_cast<const uint32_t *>(s.h_stage + 4);
    const char *strs = reinterpret_cast<const char *>(s.h_stage + 4 + (size_t)s.eager * 4);
    auto scatter = [&](const uint32_t *id, const char *sp, uint32_t m, int row) {
        // like the reference, touch only the string and its terminator in the caller's buffers
        for (uint32_t i = 0; i < m; ++i, sp += row) {
            if ((unsigned char)sp[0] == 0xFFu) continue;               // a slimmed row that did not fit: fetched below
            const size_t lim = (size_t)(row < stride ? row : stride) - 1;
            const size_t len = strnlen(sp, lim);
            char *dst = s.u_cigars + (size_t)id[i] * stride;
            memcpy(dst, sp, len);
            dst[len] = '\0';
        }
    };
    scatter(ids, strs, head, 32);
    for (uint32_t i = 0; i < head; ++i)
        if ((unsigned char)strs[(size_t)i * 32] == 0xFFu) {           // rare: a CIGAR of 32+ characters
            std::vector<char> one((size_t)stride);
            CU(cudaMemcpyAsync(one.data(), s.cig.as<char>() + (size_t)i * stride, (size_t)stride, cudaMemcpyDeviceToHost, s.stream));
            CU(cudaStreamSynchronize(s.stream));
            scatter(ids + i, one.data(), 1, stride);
        }
    if (n_cig > head) {
        const uint32_t m = n_cig - head;
        std::vector<uint32_t> id2(m);
        std::vector<char> tmp((size_t)m * stride);
        const uint32_t *cl = s.ciglist.as<uint32_t>();
        CU(cudaMemcpyAsync(id2.data(), cl + 1 + head, (size_t)m * 4, cudaMemcpyDeviceToHost, s.stream));
        CU(cudaMemcpyAsync(tmp.data(), s.cig.as<char>() + (size_t)head * stride, tmp.size(), cudaMemcpyDeviceToHost, s.stream));
        CU(cudaStreamSynchronize(s.stream));
        scatter(id2.data(), tmp.data(), m, stride);
    }
    return SALT_OK;
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

const char *salt_b200_last_error(void) { return g_err.c_str(); }
int salt_b200_abi_version(void) { return SALT_B200_ABI_VERSION; }

int salt_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

salt_b200_t *salt_b200_init(const uint32_t *mixref, uint32_t l, const uint8_t *pac, int64_t l_pac, int device)
{
    if (!mixref || l == 0) { fail(SALT_ERR_ARG, "mixref is null or empty"); return nullptr; }
    salt_b200_t *h = new_handle(device);
    if (!h) return nullptr;
    cudaStream_t st = h->slot[0].stream;
    const size_t nb = ((size_t)l + 7) / 8 * 4;
    cudaError_t e;
    if ((e = alloc_mixref(h, nb, st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(h->d_mixref, mixref, nb, cudaMemcpyHostToDevice, st)) != cudaSuccess) {
        fail(SALT_ERR_CUDA, "upload mixref", e); salt_b200_destroy(h); return nullptr;
    }
    h->l = l;
    if (pac && l_pac > 0) {
        const size_t pb = ((size_t)l_pac + 3) / 4;
        if ((e = cudaMalloc(&h->d_pac, pb + REF_PAD)) != cudaSuccess ||
            (e = cudaMemsetAsync(h->d_pac + pb, 0, REF_PAD, st)) != cudaSuccess ||
            (e = cudaMemcpyAsync(h->d_pac, pac, pb, cudaMemcpyHostToDevice, st)) != cudaSuccess) {
            fail(SALT_ERR_CUDA, "upload pac", e); salt_b200_destroy(h); return nullptr;
        }
        h->l_pac = l_pac; h->pac_cap = pb;
    }
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) { fail(SALT_ERR_CUDA, "sync", e); salt_b200_destroy(h); return nullptr; }
    return h;
}

salt_b200_t *salt_b200_init_from_bases(const char *bases, uint32_t l, const uint32_t *snp_pos,
                                       const uint8_t *snp_mask, size_t n_snp, int device)
{
    if (!bases || l == 0) { fail(SALT_ERR_ARG, "bases is null or empty"); return nullptr; }
    salt_b200_t *h = new_handle(device);
    if (!h) return nullptr;
    cudaStream_t st = h->slot[0].stream;
    const size_t nb = ((size_t)l + 7) / 8 * 4;
    char *d_bases = nullptr; uint32_t *d_pos = nullptr; uint8_t *d_mask = nullptr;
    cudaError_t e = cudaSuccess;
    do {
        if ((e = alloc_mixref(h, nb, st)) != cudaSuccess) break;
        if ((e = cudaMalloc(&d_bases, l)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(d_bases, bases, l, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        if (n_snp) {
            if ((e = cudaMalloc(&d_pos, n_snp * 4)) != cudaSuccess) break;
            if ((e = cudaMalloc(&d_mask, n_snp)) != cudaSuccess) break;
            if ((e = cudaMemcpyAsync(d_pos, snp_pos, n_snp * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
            if ((e = cudaMemcpyAsync(d_mask, snp_mask, n_snp, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        }
        if ((e = launch_build_mixref(d_bases, l, d_pos, d_mask, n_snp, h->d_mixref, st)) != cudaSuccess) break;
        h->launches += n_snp ? 2 : 1;
        e = cudaStreamSynchronize(st);
    } while (0);
    if (d_bases) cudaFree(d_bases);
    if (d_pos) cudaFree(d_pos);
    if (d_mask) cudaFree(d_mask);
    if (e != cudaSuccess) { fail(SALT_ERR_CUDA, "build mixref", e); salt_b200_destroy(h); return nullptr; }
    h->l = l;
    return h;
}

salt_b200_t *salt_b200_attach(salt_b200_t *parent)
{
    if (!parent) { fail(SALT_ERR_ARG, "null handle"); return nullptr; }
    salt_b200_t *h = new_handle(parent->device);
    if (!h) return nullptr;
    h->borrowed = true;
    h->d_mixref = parent->d_mixref; h->l = parent->l; h->d_pac = parent->d_pac; h->l_pac = parent->l_pac;
    h->fm = parent->fm; h->have_index = parent->have_index;
    h->lv_filter = parent->lv_filter; h->lv_mapping = parent->lv_mapping;
    return h;
}

int salt_b200_reload_ref(salt_b200_t *h, const uint32_t *mixref, uint32_t l, const uint8_t *pac, int64_t l_pac)
{
    if (int rc = use_device(h)) return rc;
    if (h->borrowed) return fail(SALT_ERR_ARG, "an attached handle does not own its reference");
    if (!mixref || l == 0) return fail(SALT_ERR_ARG, "mixref is null or empty");
    for (int i = 0; i < SALT_SLOTS; ++i) if (h->slot[i].pending) return fail(SALT_ERR_ARG, "a slot has a verify in flight");
    cudaStream_t st = h->slot[0].stream;
    const size_t nb = ((size_t)l + 7) / 8 * 4;
    if (nb > h->ref_cap) {
        CU(cudaStreamSynchronize(st));
        if (h->d_mixref_alloc) { CU(cudaFree(h->d_mixref_alloc)); h->d_mixref_alloc = nullptr; h->d_mixref = nullptr; h->ref_cap = 0; }
        CU(alloc_mixref(h, nb + nb / 2 + 256, st));
    }
    CU(cudaMemcpyAsync(h->d_mixref, mixref, nb, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(reinterpret_cast<uint8_t *>(h->d_mixref) + nb, 0, REF_PAD, st));      // zero pad right behind the new end
    h->l = l;
    if (pac && l_pac > 0) {
        const size_t pb = ((size_t)l_pac + 3) / 4;
        if (pb > h->pac_cap) {
            CU(cudaStreamSynchronize(st));
            if (h->d_pac) { CU(cudaFree(h->d_pac)); h->d_pac = nullptr; h->pac_cap = 0; }
            CU(cudaMalloc(&h->d_pac, pb + pb / 2 + 256 + REF_PAD));
            h->pac_cap = pb + pb / 2 + 256;
        }
        CU(cudaMemcpyAsync(h->d_pac, pac, pb, cudaMemcpyHostToDevice, st));
        CU(cudaMemsetAsync(h->d_pac + pb, 0, REF_PAD, st));
        h->l_pac = l_pac;
    } else h->l_pac = 0;
    CU(cudaStreamSynchronize(st));
    return SALT_OK;
}

int salt_b200_get_mixref(salt_b200_t *h, uint32_t *words_out, size_t n_words)
{
    if (int rc = use_device(h)) return rc;
    const size_t have = ((size_t)h->l + 7) / 8;
    if (!words_out || n_words < have) return fail(SALT_ERR_ARG, "output too small");
    CU(cudaMemcpyAsync(words_out, h->d_mixref, have * 4, cudaMemcpyDeviceToHost, h->slot[0].stream));
    CU(cudaStreamSynchronize(h->slot[0].stream));
    return SALT_OK;
}

void salt_b200_destroy(salt_b200_t *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    for (int i = 0; i < SALT_SLOTS; ++i) {
        if (h->slot[i].stream) cudaStreamSynchronize(h->slot[i].stream);
        h->slot[i].release();
    }
    DBuf *all[] = {&h->pairs, &h->out8, &h->kbuf, &h->cig, &h->wins, &h->sswout, &h->sswcig, &h->sswscratch, &h->sswovf, &h->ix_cbwt, &h->ix_csa, &h->ix_lkt, &h->ix_rbwt, &h->ix_rocc, &h->ix_rmaj, &h->ix_rsa, &h->ix_c32, &h->ix_r64,
                   &h->fpairs, &h->fslots, &h->fcount, &h->fpairs2, &h->fslots2, &h->md_in, &h->md_cig, &h->md_str, &h->md_xv, &h->md_out};
    for (DBuf *b : all) b->release();
    for (int i = 0; i < 7; ++i) if (h->ev_ssw[i]) cudaEventDestroy(h->ev_ssw[i]);
    if (!h->borrowed) {                     // an attached handle only points at its parent's reference and indexes
        if (h->d_mixref_alloc) cudaFree(h->d_mixref_alloc);
        if (h->d_pac) cudaFree(h->d_pac);
    }
    delete h;
}

int salt_b200_set_stream(salt_b200_t *h, void *cuda_stream)
{
    if (int rc = use_device(h)) return rc;
    Slot &s = h->slot[0];
    if (s.pending) return fail(SALT_ERR_ARG, "slot 0 has a verify in flight");
    if (s.own_stream && s.stream) { CU(cudaStreamSynchronize(s.stream)); CU(cudaStreamDestroy(s.stream)); s.stream = nullptr; }
    if (cuda_stream) { s.stream = static_cast<cudaStream_t>(cuda_stream); s.own_stream = false; }
    else { CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking)); s.own_stream = true; }
    return SALT_OK;
}

int salt_b200_sync(salt_b200_t *h)
{
    if (int rc = use_device(h)) return rc;
    for (int i = 0; i < SALT_SLOTS; ++i) {
        if (h->slot[i].pending) { if (int rc = finish_verify(h, i)) return rc; }
        else CU(cudaStreamSynchronize(h->slot[i].stream));
    }
    return SALT_OK;
}

void *salt_b200_host_alloc(size_t bytes)
{
    void *p = nullptr;
    cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) { fail(SALT_ERR_CUDA, "cudaMallocHost", e); return nullptr; }
    return p;
}

void salt_b200_host_free(void *p) { if (p) cudaFreeHost(p); }

uint64_t salt_b200_launch_count(salt_b200_t *h, int reset)
{
    if (!h) return 0;
    const uint64_t v = h->launches;
    if (reset) h->launches = 0;
    return v;
}

int salt_b200_set_reads(salt_b200_t *h, const salt_reads_t *reads)
{
    if (int rc = use_device(h)) return rc;
    Slot &s = h->slot[0];
    if (s.pending) return fail(SALT_ERR_ARG, "slot 0 has a verify in flight");
    if (int rc = load_reads(h, s, reads)) return rc;
    CU(cudaStreamSynchronize(s.stream));      // the caller's buffers are free to change after return
    return SALT_OK;
}

// ------------------------------------------------------------------ mismatch / LV
int salt_b200_mismatch_dev(salt_b200_t *h, const salt_pair_t *d_pairs, size_t n, int max_err, int8_t *d_out)
{
    if (int rc = use_device(h)) return rc;
    if (n && (!d_pairs || !d_out)) return fail(SALT_ERR_ARG, "null buffer");
    if (max_err < 0 || max_err > 127) return fail(SALT_ERR_ARG, "max_err must be in 0..127");
    if (!h->slot[h->cur].n_reads) return fail(SALT_ERR_ARG, "no reads set");
    CU(launch_mismatch(h->ctx(h->cur), d_pairs, n, max_err, d_out, h->slot[h->cur].stream));
    if (n) h->launches += 1;
    return SALT_OK;
}

int salt_b200_mismatch(salt_b200_t *h, const salt_pair_t *pairs, size_t n, int max_err, int8_t *out)
{
    if (int rc = use_device(h)) return rc;
    if (n && (!pairs || !out)) return fail(SALT_ERR_ARG, "null buffer");
    if (!n) return SALT_OK;
    cudaStream_t st = h->slot[h->cur].stream;
    CU(h->pairs.need(n * sizeof(salt_pair_t)));
    CU(h->out8.need(n));
    CU(cudaMemcpyAsync(h->pairs.p, pairs, n * sizeof(salt_pair_t), cudaMemcpyHostToDevice, st));
    if (int rc = salt_b200_mismatch_dev(h, h->pairs.as<salt_pair_t>(), n, max_err, h->out8.as<int8_t>())) return rc;
    CU(cudaMemcpyAsync(out, h->out8.p, n, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SALT_OK;
}

int salt_b200_lv_dev(salt_b200_t *h, const salt_pair_t *d_pairs, size_t n, int k, int8_t *d_out)
{
    if (int rc = use_device(h)) return rc;
    if (n && (!d_pairs || !d_out)) return fail(SALT_ERR_ARG, "null buffer");
    if (!h->slot[h->cur].n_reads) return fail(SALT_ERR_ARG, "no reads set");
    LvFilterScratch fs{nullptr, nullptr, nullptr};
    if (h->lv_filter && n) {
        CU(h->fpairs.need((n + 1) * sizeof(salt_pair_t)));
        CU(h->fslots.need((n + 1) * 4));
        CU(h->fcount.need(256));
        fs.pairs = h->fpairs.as<salt_pair_t>(); fs.slots = h->fslots.as<uint32_t>(); fs.count = h->fcount.as<uint32_t>();
        if (h->lv_two_pass) {
            CU(h->fpairs2.need((n + 1) * sizeof(salt_pair_t)));
            CU(h->fslots2.need((n + 1) * 4));
            fs.pairs2 = h->fpairs2.as<salt_pair_t>(); fs.slots2 = h->fslots2.as<uint32_t>(); fs.count2 = h->fcount.as<uint32_t>() + 4;
        }
    }
    CU(launch_lv(h->ctx(h->cur), d_pairs, n, k, nullptr, nullptr, 0, d_out, h->sm_count, h->slot[h->cur].stream, h->lv_mapping,
                 h->lv_filter ? &fs : nullptr));
    if (n) h->launches += h->lv_filter ? 2 : 1;
    return SALT_OK;
}

int salt_b200_lv(salt_b200_t *h, const salt_pair_t *pairs, size_t n, int k, int8_t *out)
{
    if (int rc = use_device(h)) return rc;
    if (n && (!pairs || !out)) return fail(SALT_ERR_ARG, "null buffer");
    if (!n) return SALT_OK;
    cudaStream_t st = h->slot[h->cur].stream;
    CU(h->pairs.need(n * sizeof(salt_pair_t)));
    CU(h->out8.need(n));
    CU(cudaMemcpyAsync(h->pairs.p, pairs, n * sizeof(salt_pair_t), cudaMemcpyHostToDevice, st));
    if (int rc = salt_b200_lv_dev(h, h->pairs.as<salt_pair_t>(), n, k, h->out8.as<int8_t>())) return rc;
    CU(cudaMemcpyAsync(out, h->out8.p, n, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SALT_OK;
}

int salt_b200_lv_cigar(salt_b200_t *h, const salt_pair_t *pairs, const uint8_t *k_each, size_t n,
                       char *cigars, int stride, int8_t *out)
{
    if (int rc = use_device(h)) return rc;
    if (n && (!pairs || !k_each || !cigars || !out)) return fail(SALT_ERR_ARG, "null buffer");
    if (stride < 2) return fail(SALT_ERR_ARG, "cigar stride too small");
    if (!h->slot[h->cur].n_reads) return fail(SALT_ERR_ARG, "no reads set");
    if (!n) return SALT_OK;
    int kmax = 0;
    for (size_t i = 0; i < n; ++i) {
        if (k_each[i] >= 31) return fail(SALT_ERR_ARG, "k must be < 31 (LandauVishkin.c:183 asserts)");
        if (k_each[i] > kmax) kmax = k_each[i];
    }
    cudaStream_t st = h->slot[h->cur].stream;
    CU(h->pairs.need(n * sizeof(salt_pair_t)));
    CU(h->out8.need(n));
    CU(h->kbuf.need(n));
    CU(h->cig.need(n * (size_t)stride));
    CU(cudaMemcpyAsync(h->pairs.p, pairs, n * sizeof(salt_pair_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->kbuf.p, k_each, n, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(h->cig.p, 0, n * (size_t)stride, st));
    CU(launch_lv_cigar(h->ctx(h->cur), h->pairs.as<salt_pair_t>(), h->kbuf.as<uint8_t>(), n, nullptr, nullptr, 0, nullptr, kmax,
                       h->cig.as<char>(), stride, h->out8.as<int8_t>(), h->sm_count, st, h->lv_mapping));
    h->launches += 1;
    std::vector<char> tmp(n * (size_t)stride);
    CU(cudaMemcpyAsync(tmp.data(), h->cig.p, tmp.size(), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out, h->out8.p, n, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    // like the reference, touch only the string and its terminator in the caller's buffers
    for (size_t i = 0; i < n; ++i) {
        const char *sp = tmp.data() + i * (size_t)stride;
        size_t len = strnlen(sp, (size_t)stride - 1);
        memcpy(cigars + i * (size_t)stride, sp, len);
        cigars[i * (size_t)stride + len] = '\0';
    }
    return SALT_OK;
}

// ------------------------------------------------------------------ SAM tail
int salt_b200_md_nm(salt_b200_t *h, int slot, const salt_mdnm_in_t *items, size_t n, const char *cigars, int cigar_stride,
                    char *md, int md_stride, uint16_t *xv, int xv_stride, salt_mdnm_out_t *out)
{
    if (int rc = use_device(h)) return rc;
    if (n && (!items || !cigars || !md || !out)) return fail(SALT_ERR_ARG, "null buffer");
    if (cigar_stride < 2 || md_stride < 2) return fail(SALT_ERR_ARG, "string stride too small");
    if (xv_stride < 0 || xv_stride > 64 || (xv_stride > 0 && !xv)) return fail(SALT_ERR_ARG, "xv_stride must be 0..64 with a buffer");
    if (!h->d_pac) return fail(SALT_ERR_ARG, "MD/NM need the 2-bit pac (salt_b200_init was given none)");
    if (slot < 0 || slot >= SALT_SLOTS) return fail(SALT_ERR_ARG, "slot out of range");
    Slot &s = h->slot[slot];
    if (!s.n_reads) return fail(SALT_ERR_ARG, "no reads set");
    if (s.pending) return fail(SALT_ERR_ARG, "slot has a verify in flight: call salt_b200_verify_wait first");
    if (!n) return SALT_OK;
    cudaStream_t st = s.stream;
    const size_t xs = (size_t)(xv_stride > 0 ? xv_stride : 1);
    CU(h->md_in.need(n * sizeof(salt_mdnm_in_t)));
    CU(h->md_cig.need(n * (size_t)cigar_stride));
    CU(h->md_str.need(n * (size_t)md_stride));
    CU(h->md_xv.need(n * xs * 2));
    CU(h->md_out.need(n * sizeof(salt_mdnm_out_t)));
    CU(cudaMemcpyAsync(h->md_in.p, items, n * sizeof(salt_mdnm_in_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->md_cig.p, cigars, n * (size_t)cigar_stride, cudaMemcpyHostToDevice, st));
    if (xv_stride > 0) CU(cudaMemsetAsync(h->md_xv.p, 0, n * xs * 2, st));
    CU(launch_md_nm(h->ctx(slot), s.codes.as<uint8_t>(), s.d_roffs(), h->md_in.as<salt_mdnm_in_t>(), n, h->md_cig.as<char>(),
                    cigar_stride, h->md_str.as<char>(), md_stride, h->md_xv.as<uint16_t>(), xv_stride,
                    h->md_out.as<salt_mdnm_out_t>(), st));
    h->launches += 1;
    CU(cudaMemcpyAsync(md, h->md_str.p, n * (size_t)md_stride, cudaMemcpyDeviceToHost, st));
    if (xv_stride > 0) CU(cudaMemcpyAsync(xv, h->md_xv.p, n * xs * 2, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out, h->md_out.p, n * sizeof(salt_mdnm_out_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SALT_OK;
}

// Queue the tags of the slot's primaries behind whatever the slot's stream still has to do (a verify just submitted
// included): kernels, scan, packing and the downloads, with the first `eager` bytes of the packed MD stream copied
// without waiting for its length.  tail_finish completes it.
static int tail_enqueue(salt_b200_t *h, int slot, salt_mdnm_out_t *out, uint32_t *md_offs, char *md_packed, size_t md_cap,
                        uint16_t *xv, int xv_stride)
{
    if (slot < 0 || slot >= SALT_SLOTS) return fail(SALT_ERR_ARG, "slot out of range");
    Slot &s = h->slot[slot];
    if (!s.have_rec || !s.n_reads) return fail(SALT_ERR_ARG, "slot holds no verified chunk");
    if (!s.rec_cig_stride) return fail(SALT_ERR_ARG, "the chunk was verified without CIGARs: gapped primaries would have no tags");
    if (!h->d_pac) return fail(SALT_ERR_ARG, "MD/NM need the 2-bit pac (salt_b200_init was given none)");
    if (!out || !md_offs || !md_packed) return fail(SALT_ERR_ARG, "null buffer");
    if (xv_stride < 0 || xv_stride > 64 || (xv_stride > 0 && !xv)) return fail(SALT_ERR_ARG, "xv_stride must be 0..64 with a buffer");
    if (s.tail_pending) return fail(SALT_ERR_ARG, "slot has a tail in flight: call salt_b200_tail_wait first");
    const uint32_t n = s.n_reads;
    const int mstride = 2 * (int)s.l_max + 16 < 64 ? 64 : 2 * (int)s.l_max + 16;     // an MD string never needs more than ~2 characters per base
    const size_t xs = (size_t)(xv_stride > 0 ? xv_stride : 1);
    cudaStream_t st = s.stream;
    CU(s.tl_md.need((size_t)n * mstride)); CU(s.tl_out.need((size_t)n * sizeof(salt_mdnm_out_t)));
    CU(s.tl_len.need((size_t)n * 4 + 16)); CU(s.tl_offs.need(((size_t)n + 1) * 4)); CU(s.tl_cigrow.need((size_t)n * 4));
    CU(s.tl_xv.need((size_t)n * xs * 2)); CU(s.tl_scan.need(3 * (size_t)scan3_blocks(n) * 4 + 16));
    CU(s.tl_packed.need((size_t)n * mstride + 16));                   // worst case: known before the lengths are
    if (xv_stride > 0) CU(cudaMemsetAsync(s.tl_xv.p, 0, (size_t)n * xs * 2, st));
    const uint32_t *cl = s.ciglist.as<uint32_t>();
    CU(launch_tail_primaries(h->ctx(slot), s.codes.as<uint8_t>(), s.d_roffs(), s.rec.as<salt_verify_out_t>(), cl + 1, cl,
                             s.cig.as<char>(), s.rec_cig_stride, s.tl_cigrow.as<int32_t>(), s.tl_md.as<char>(), mstride,
                             s.tl_xv.as<uint16_t>(), xv_stride, s.tl_out.as<salt_mdnm_out_t>(), s.tl_len.as<uint32_t>(), st));
    Scan3 sc{};
    sc.n = n; sc.partial = s.tl_scan.as<uint32_t>(); sc.in[0] = s.tl_len.p; sc.width[0] = 32; sc.out[0] = s.tl_offs.as<uint32_t>();
    CU(launch_scan3(sc, 1, st));
    CU(launch_tail_pack(s.tl_md.as<char>(), mstride, s.tl_offs.as<uint32_t>(), n, s.tl_packed.as<char>(), st));
    CU(cudaMemcpyAsync(md_offs, s.tl_offs.p, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out, s.tl_out.p, (size_t)n * sizeof(salt_mdnm_out_t), cudaMemcpyDeviceToHost, st));
    if (xv_stride > 0) CU(cudaMemcpyAsync(xv, s.tl_xv.p, (size_t)n * xs * 2, cudaMemcpyDeviceToHost, st));
    size_t eager = (size_t)n * 12;                                    // typical MD strings are a handful of characters
    if (eager > md_cap) eager = md_cap;
    if (eager > (size_t)n * mstride) eager = (size_t)n * mstride;
    if (eager) CU(cudaMemcpyAsync(md_packed, s.tl_packed.p, eager, cudaMemcpyDeviceToHost, st));
    s.tail_pending = true; s.tail_eager = eager; s.tail_offs = md_offs; s.tail_md = md_packed; s.tail_cap = md_cap;
    h->launches += 7;
    return SALT_OK;
}

static int tail_finish(salt_b200_t *h, int slot, size_t *md_bytes)
{
    Slot &s = h->slot[slot];
    if (!s.tail_pending) return fail(SALT_ERR_ARG, "slot has no tail in flight");
    s.tail_pending = false;
    if (s.pending) { if (int rc = finish_verify(h, slot)) return rc; }      // same stream: the verify is complete too
    else CU(cudaStreamSynchronize(s.stream));
    const size_t total = s.tail_offs[s.n_reads];
    if (md_bytes) *md_bytes = total;
    if (total > s.tail_cap) return fail(SALT_ERR_NOMEM, "packed MD buffer too small");
    if (total > s.tail_eager) {
        CU(cudaMemcpyAsync(s.tail_md + s.tail_eager, s.tl_packed.as<char>() + s.tail_eager, total - s.tail_eager, cudaMemcpyDeviceToHost, s.stream));
        CU(cudaStreamSynchronize(s.stream));
    }
    return SALT_OK;
}

int salt_b200_tail_submit(salt_b200_t *h, int slot, salt_mdnm_out_t *out, uint32_t *md_offs, char *md_packed, size_t md_cap,
                          uint16_t *xv, int xv_stride)
{
    if (int rc = use_device(h)) return rc;
    return tail_enqueue(h, slot, out, md_offs, md_packed, md_cap, xv, xv_stride);
}

int salt_b200_tail_wait(salt_b200_t *h, int slot, size_t *md_bytes)
{
    if (int rc = use_device(h)) return rc;
    if (slot < 0 || slot >= SALT_SLOTS) return fail(SALT_ERR_ARG, "slot out of range");
    return tail_finish(h, slot, md_bytes);
}

int salt_b200_tail_primaries(salt_b200_t *h, int slot, salt_mdnm_out_t *out, uint32_t *md_offs, char *md_packed, size_t md_cap,
                             size_t *md_bytes, uint16_t *xv, int xv_stride)
{
    if (int rc = use_device(h)) return rc;
    if (slot < 0 || slot >= SALT_SLOTS) return fail(SALT_ERR_ARG, "slot out of range");
    if (h->slot[slot].pending) return fail(SALT_ERR_ARG, "slot has a verify in flight: call salt_b200_verify_wait first");
    if (int rc = tail_enqueue(h, slot, out, md_offs, md_packed, md_cap, xv, xv_stride)) return rc;
    return tail_finish(h, slot, md_bytes);
}

// ------------------------------------------------------------------ SSW
static int fill_ssw_params(SswParams &p, int use_pac, const int8_t *mat, int n_sym, int gapO, int gapE, int flag,
                           int filters, int filterd, int mask_len)
{
    if (!mat || n_sym < 1 || n_sym > 16) return fail(SALT_ERR_ARG, "mat is null or n_sym outside 1..16");
    if (use_pac && n_sym < 5) return fail(SALT_ERR_ARG, "pac scoring needs n_sym >= 5 (codes 0..4)");
    if (!use_pac && n_sym != 16) return fail(SALT_ERR_ARG, "mixRef scoring needs n_sym == 16 (4-bit masks)");
    if (gapO < 0 || gapO > 255 || gapE < 0 || gapE > 255) return fail(SALT_ERR_ARG, "gap penalties are uint8 in ssw_align");
    if (gapO <= gapE)
        return fail(SALT_ERR_UNSUPPORTED, "gapO <= gapE: the SSE2 lazy-F loop of the reference under-propagates there and is not reproduced");
    p.use_pac = use_pac; p.n_sym = n_sym; p.gapO = gapO; p.gapE = gapE; p.flag = flag;
    p.filters = filters; p.filterd = filterd; p.mask_len = mask_len;
    const int nn = n_sym * n_sym;
    for (int sym = 0; sym < 17; ++sym)
        for (int code = 0; code < 8; ++code) {
            int v = 0;
            if (sym == 16 || code == 5) v = -128;             // pad column / row above the read
            else if (code == 6 || code == 7) v = 0;           // zero-score padded rows (ssw.c:361)
            else if (sym >= n_sym) v = -128;
            else {
                const int read_sym = use_pac ? code : (1 << code);      // alnpe.c:283 / :344
                int idx = sym * n_sym + read_sym;                        // ssw.c:361
                if (idx >= nn) idx = nn - 1;
                v = mat[idx];
            }
            p.table[sym * 8 + code] = (int8_t)v;
        }
    return SALT_OK;
}

int salt_b200_ssw_dev(salt_b200_t *h, const salt_win_t *d_wins, size_t n, int use_pac,
                      const int8_t *mat, int n_sym, int gapO, int gapE, int flag,
                      int filters, int filterd, int mask_len,
                      salt_ssw_out_t *d_out, uint32_t *d_cigars, int cigar_stride)
{
    if (int rc = use_device(h)) return rc;
    if (n && (!d_wins || !d_out || !d_cigars)) return fail(SALT_ERR_ARG, "null buffer");
    if (cigar_stride < 1) return fail(SALT_ERR_ARG, "cigar stride too small");
    if (!h->slot[h->cur].n_reads) return fail(SALT_ERR_ARG, "no reads set");
    if (use_pac && !h->d_pac) return fail(SALT_ERR_ARG, "no pac uploaded");
    SswParams prm;
    if (int rc = fill_ssw_params(prm, use_pac, mat, n_sym, gapO, gapE, flag, filters, filterd, mask_len)) return rc;
    int maxpos = 0;
    for (int i = 0; i < n_sym * n_sym; ++i) if (mat[i] > maxpos) maxpos = mat[i];
    if ((int64_t)maxpos * (int64_t)h->slot[h->cur].l_max >= 24000) return fail(SALT_ERR_UNSUPPORTED, "scores would overflow int16");
    if (!n) return SALT_OK;
    const int max_cols = h->max_window;
    size_t lay[13];
    const size_t need = ssw_scratch_bytes(n, max_cols, (int)h->slot[h->cur].l_max, lay);
    CU(h->sswscratch.need(need));
    CU(h->sswovf.need(ssw_overflow_bytes((int)h->slot[h->cur].l_max)));     // per handle: two handles never share direction bytes
    if (h->profiling)
        for (int i = 0; i < 7; ++i) if (!h->ev_ssw[i]) CU(cudaEventCreate(&h->ev_ssw[i]));
    CU(launch_ssw(h->ctx(h->cur), d_wins, n, prm, h->sswscratch.p, h->sswscratch.cap, max_cols, d_out, d_cigars,
                  cigar_stride, h->sm_count, h->slot[h->cur].stream, &h->launches, h->profiling ? h->ev_ssw : nullptr,
                  h->sswovf.as<uint8_t>()));
    h->have_ssw_prof = h->profiling;
    return SALT_OK;
}

int salt_b200_use_slot(salt_b200_t *h, int slot)
{
    if (!h) return fail(SALT_ERR_ARG, "null handle");
    if (slot < 0 || slot >= SALT_SLOTS) return fail(SALT_ERR_ARG, "slot out of range");
    if (h->slot[slot].pending) return fail(SALT_ERR_ARG, "slot has a verify in flight: call salt_b200_verify_wait first");
    h->cur = slot;
    return SALT_OK;
}

int salt_b200_set_lv_mapping(salt_b200_t *h, int mapping)
{
    if (!h) return fail(SALT_ERR_ARG, "null handle");
    if (mapping < 0 || (mapping & 15) > 2 || mapping > 18) return fail(SALT_ERR_ARG, "mapping must be 0 (auto), 1 (warp per pair) or 2 (thread per pair), + 16 for a single pass");
    h->lv_mapping = mapping & 15;
    h->lv_two_pass = (mapping & 16) ? 0 : 1;
    return SALT_OK;
}

int salt_b200_set_lv_filter(salt_b200_t *h, int enable)
{
    if (!h) return fail(SALT_ERR_ARG, "null handle");
    h->lv_filter = enable ? 1 : 0;
    return SALT_OK;
}

int salt_b200_set_max_window(salt_b200_t *h, int cols)
{
    if (!h) return fail(SALT_ERR_ARG, "null handle");
    if (cols < 1 || cols > 65536) return fail(SALT_ERR_ARG, "max window must be in 1..65536");
    h->max_window = cols;
    return SALT_OK;
}

int salt_b200_ssw(salt_b200_t *h, const salt_win_t *wins, size_t n, int use_pac,
                  const int8_t *mat, int n_sym, int gapO, int gapE, int flag,
                  int filters, int filterd, int mask_len,
                  salt_ssw_out_t *out, uint32_t *cigars, int cigar_stride)
{
    if (int rc = use_device(h)) return rc;
    if (n && (!wins || !out || !cigars)) return fail(SALT_ERR_ARG, "null buffer");
    if (cigar_stride < 1) return fail(SALT_ERR_ARG, "cigar stride too small");
    if (!n) return SALT_OK;
    cudaStream_t st = h->slot[h->cur].stream;
    uint32_t widest = 1;
    for (size_t i = 0; i < n; ++i)
        if (wins[i].end >= wins[i].start && wins[i].end - wins[i].start + 1 > widest) widest = wins[i].end - wins[i].start + 1;
    if (widest > 65536) return fail(SALT_ERR_UNSUPPORTED, "rescue window wider than 65536 bases");
    h->max_window = (int)((widest + 7) / 8 * 8);
    CU(h->wins.need(n * sizeof(salt_win_t)));
    CU(h->sswout.need(n * sizeof(salt_ssw_out_t)));
    CU(h->sswcig.need(n * (size_t)cigar_stride * 4));
    CU(cudaMemcpyAsync(h->wins.p, wins, n * sizeof(salt_win_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(h->sswcig.p, 0, n * (size_t)cigar_stride * 4, st));
    if (int rc = salt_b200_ssw_dev(h, h->wins.as<salt_win_t>(), n, use_pac, mat, n_sym, gapO, gapE, flag, filters,
                                   filterd, mask_len, h->sswout.as<salt_ssw_out_t>(), h->sswcig.as<uint32_t>(), cigar_stride))
        return rc;
    CU(cudaMemcpyAsync(out, h->sswout.p, n * sizeof(salt_ssw_out_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(cigars, h->sswcig.p, n * (size_t)cigar_stride * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SALT_OK;                                   // declined windows are reported per item: out[i].cigarLen < 0
}

// ------------------------------------------------------------------ verification stage
int salt_b200_verify_dev(salt_b200_t *h, const uint32_t *d_offs0, const uint32_t *d_loci0, size_t n0,
                         const uint32_t *d_offs1, const uint32_t *d_loci1, size_t n1,
                         int nogap_T0, int lv_T0, salt_verify_out_t *d_rec, int8_t *d_acc0, int8_t *d_acc1,
                         char *d_cigars, int cigar_stride, uint32_t *d_cig_reads, uint32_t *d_cig_count)
{
    if (int rc = use_device(h)) return rc;
    if (h->slot[0].pending) return fail(SALT_ERR_ARG, "slot 0 has a verify in flight");
    return verify_on_device(h, 0, d_offs0, d_loci0, n0, d_offs1, d_loci1, n1, nogap_T0, lv_T0, d_rec, d_acc0, d_acc1,
                            d_cigars, cigar_stride, d_cig_reads, d_cig_count);
}

int salt_b200_verify(salt_b200_t *h, const salt_cands_t *cands, int nogap_T0, int lv_T0,
                     salt_verify_out_t *rec, int8_t *acc0, int8_t *acc1, char *cigars, int cigar_stride)
{
    if (int rc = use_device(h)) return rc;
    if (int rc = enqueue_verify(h, 0, cands, nogap_T0, lv_T0, rec, acc0, acc1, cigars, cigar_stride)) return rc;
    return finish_verify(h, 0);
}

int salt_b200_n_slots(void) { return SALT_SLOTS; }

int salt_b200_verify_submit(salt_b200_t *h, int slot, const salt_reads_t *reads, const salt_cands_t *cands,
                            int nogap_T0, int lv_T0, salt_verify_out_t *rec, int8_t *acc0, int8_t *acc1,
                            char *cigars, int cigar_stride)
{
    if (int rc = use_device(h)) return rc;
    if (slot < 0 || slot >= SALT_SLOTS) return fail(SALT_ERR_ARG, "slot out of range");
    Slot &s = h->slot[slot];
    if (s.pending) return fail(SALT_ERR_ARG, "slot has a verify in flight: call salt_b200_verify_wait first");
    if (int rc = load_reads(h, s, reads, cands)) return rc;
    if (!s.n_reads) return SALT_OK;
    return enqueue_verify(h, slot, cands, nogap_T0, lv_T0, rec, acc0, acc1, cigars, cigar_stride);
}

int salt_b200_verify_wait(salt_b200_t *h, int slot)
{
    if (int rc = use_device(h)) return rc;
    if (slot < 0 || slot >= SALT_SLOTS) return fail(SALT_ERR_ARG, "slot out of range");
    return finish_verify(h, slot);
}

int salt_b200_verify_batch(salt_b200_t *h, const salt_reads_t *reads, const salt_cands_t *cands, uint32_t chunk_reads,
                           int nogap_T0, int lv_T0, salt_verify_out_t *rec, int8_t *acc0, int8_t *acc1,
                           char *cigars, int cigar_stride)
{
    if (int rc = use_device(h)) return rc;
    if (!reads || !reads->offs || !cands || !cands->offs[0] || !cands->offs[1] || !rec) return fail(SALT_ERR_ARG, "null buffer");
    if (chunk_reads == 0) chunk_reads = 100000;          // N_SEQS, aln.h:27
    const uint32_t n = reads->n_reads;
    int rc = SALT_OK;
    uint32_t k = 0;
    for (uint32_t b = 0; b < n && rc == SALT_OK; b += chunk_reads, ++k) {
        const int si = (int)(k % SALT_SLOTS);
        if ((rc = finish_verify(h, si)) != SALT_OK) break;
        const uint32_t m = n - b < chunk_reads ? n - b : chunk_reads;
        // per-chunk views need offsets that start at 0: rebased copies in pinned memory, owned by the slot
        Slot &sl = h->slot[si];
        if (sl.h_offs_cap < 3 * ((size_t)m + 1)) {
            if (sl.h_offs) { cudaFreeHost(sl.h_offs); sl.h_offs = nullptr; sl.h_offs_cap = 0; }
            const size_t want = 3 * ((size_t)(m > chunk_reads ? m : chunk_reads) + 1);
            if (cudaMallocHost(&sl.h_offs, want * 4) != cudaSuccess) { rc = fail(SALT_ERR_NOMEM, "pinned offsets"); break; }
            sl.h_offs_cap = want;
        }
        uint32_t *r_ = sl.h_offs, *a_ = r_ + (m + 1), *b_ = a_ + (m + 1);
        const uint32_t rb = reads->offs[b], c0 = cands->offs[0][b], c1 = cands->offs[1][b];
        for (uint32_t i = 0; i <= m; ++i) {
            r_[i] = reads->offs[b + i] - rb; a_[i] = cands->offs[0][b + i] - c0; b_[i] = cands->offs[1][b + i] - c1;
        }
        salt_reads_t rv; rv.codes = reads->codes + rb; rv.offs = r_; rv.n_reads = m;
        salt_cands_t cv; cv.offs[0] = a_; cv.offs[1] = b_;
        cv.loci[0] = cands->loci[0] ? cands->loci[0] + c0 : nullptr; cv.loci[1] = cands->loci[1] ? cands->loci[1] + c1 : nullptr;
        rc = salt_b200_verify_submit(h, si, &rv, &cv, nogap_T0, lv_T0, rec + b, acc0 ? acc0 + c0 : nullptr,
                                     acc1 ? acc1 + c1 : nullptr, cigars ? cigars + (size_t)b * cigar_stride : nullptr, cigar_stride);
    }
    for (int si = 0; si < SALT_SLOTS; ++si) { const int r2 = finish_verify(h, si); if (rc == SALT_OK) rc = r2; }
    return rc;
}

// ------------------------------------------------------------------ compact transport
int salt_b200_set_reads_packed(salt_b200_t *h, const salt_packed_chunk_t *pc)
{
    if (int rc = use_device(h)) return rc;
    if (int rc = check_packed(pc, false)) return rc;
    Slot &s = h->slot[0];
    if (s.pending) return fail(SALT_ERR_ARG, "slot 0 has a verify in flight");
    PackedView v; v.pc = pc; v.base_pos = pc->base_start;
    while (v.np_lo < pc->n_n && pc->n_pos[v.np_lo] < v.base_pos) ++v.np_lo;
    if (int rc = grow_view(v, pc->n_reads, false)) return rc;
    if (int rc = load_packed(h, s, v, false)) return rc;
    CU(cudaStreamSynchronize(s.stream));
    return SALT_OK;
}

int salt_b200_verify_submit_packed(salt_b200_t *h, int slot, const salt_packed_chunk_t *pc, int nogap_T0, int lv_T0,
                                   salt_verify_out_t *rec, int8_t *acc0, int8_t *acc1, char *cigars, int cigar_stride)
{
    if (int rc = use_device(h)) return rc;
    if (slot < 0 || slot >= SALT_SLOTS) return fail(SALT_ERR_ARG, "slot out of range");
    if (int rc = check_packed(pc, true)) return rc;
    if (!rec) return fail(SALT_ERR_ARG, "null buffer");
    Slot &s = h->slot[slot];
    if (s.pending) return fail(SALT_ERR_ARG, "slot has a verify in flight: call salt_b200_verify_wait first");
    PackedView v; v.pc = pc; v.base_pos = pc->base_start;
    while (v.np_lo < pc->n_n && pc->n_pos[v.np_lo] < v.base_pos) ++v.np_lo;
    if (int rc = grow_view(v, pc->n_reads, true)) return rc;
    if (int rc = load_packed(h, s, v, true)) return rc;
    if (!s.n_reads) return SALT_OK;
    return run_and_download(h, slot, v.n0, v.n1, nogap_T0, lv_T0, rec, acc0, acc1, cigars, cigar_stride);
}

int salt_b200_verify_batch_packed(salt_b200_t *h, const salt_packed_chunk_t *pc, uint32_t chunk_reads, int nogap_T0, int lv_T0,
                                  salt_verify_out_t *rec, int8_t *acc0, int8_t *acc1, char *cigars, int cigar_stride)
{
    if (int rc = use_device(h)) return rc;
    if (int rc = check_packed(pc, true)) return rc;
    if (!rec) return fail(SALT_ERR_ARG, "null buffer");
    if (chunk_reads == 0) chunk_reads = 100000;          // N_SEQS, aln.h:27
    int rc = SALT_OK;
    PackedView v; v.pc = pc; v.base_pos = pc->base_start;
    while (v.np_lo < pc->n_n && pc->n_pos[v.np_lo] < v.base_pos) ++v.np_lo;
    v.np_hi = v.np_lo;
    uint32_t k = 0;
    for (uint32_t b = 0; b < pc->n_reads && rc == SALT_OK; b += chunk_reads, ++k) {
        const int si = (int)(k % SALT_SLOTS);
        if ((rc = finish_verify(h, si)) != SALT_OK) break;
        const uint32_t m = pc->n_reads - b < chunk_reads ? pc->n_reads - b : chunk_reads;
        // the next view starts where the last one ended
        PackedView w; w.pc = pc; w.first = b; w.base_pos = v.base_pos + v.n_bases; w.np_lo = v.np_hi;
        w.c0 = v.c0 + v.n0; w.c1 = v.c1 + v.n1;
        if ((rc = grow_view(w, m, true)) != SALT_OK) break;
        v = w;
        Slot &sl = h->slot[si];
        if ((rc = load_packed(h, sl, v, true)) != SALT_OK) break;
        rc = run_and_download(h, si, v.n0, v.n1, nogap_T0, lv_T0, rec + b, acc0 ? acc0 + v.c0 : nullptr,
                              acc1 ? acc1 + v.c1 : nullptr, cigars ? cigars + (size_t)b * cigar_stride : nullptr, cigar_stride);
    }
    for (int si = 0; si < SALT_SLOTS; ++si) { const int r2 = finish_verify(h, si); if (rc == SALT_OK) rc = r2; }
    return rc;
}

// ------------------------------------------------------------------ seeding + locate
int salt_b200_set_index(salt_b200_t *h, const salt_fm_index_t *ix)
{
    if (int rc = use_device(h)) return rc;
    if (!ix || !ix->c_bwt || !ix->c_sa || !ix->lkt || !ix->r_bwt || !ix->r_occ || !ix->r_occ_major || !ix->r_sa_sharp)
        return fail(SALT_ERR_ARG, "an index array is null");
    if (ix->lkt_len < 1 || ix->lkt_len > 14) return fail(SALT_ERR_ARG, "lookup length out of range");
    if (ix->c_sa_intv < 1 || ix->c_n_sa < 1) return fail(SALT_ERR_ARG, "suffix-array sampling is empty");
    cudaStream_t st = h->slot[0].stream;
    const size_t lkt_items = ((size_t)1 << (2 * ix->lkt_len)) + 1;
    // the SNP-context BWT is read up to the next 256-character boundary past its end (rbwt.c:248 sizes it that way)
    const size_t rpad = (ix->r_bwt_words * 8 + 256) / 256 * 256 / 8 + 1;
    CU(h->ix_cbwt.need(ix->c_bwt_words * 4 + 64)); CU(h->ix_csa.need((size_t)ix->c_n_sa * 4));
    CU(h->ix_lkt.need(lkt_items * 4)); CU(h->ix_rbwt.need(rpad * 4 + 64)); CU(h->ix_rocc.need(ix->r_occ_words * 4 + 64));
    CU(h->ix_rmaj.need(ix->r_occ_major_words * 4 + 64)); CU(h->ix_rsa.need(ix->r_n_sa_sharp * 4 + 4));
    CU(cudaMemsetAsync(h->ix_rbwt.p, 0, rpad * 4 + 64, st));
    CU(cudaMemcpyAsync(h->ix_cbwt.p, ix->c_bwt, ix->c_bwt_words * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->ix_csa.p, ix->c_sa, (size_t)ix->c_n_sa * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->ix_lkt.p, ix->lkt, lkt_items * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->ix_rbwt.p, ix->r_bwt, ix->r_bwt_words * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->ix_rocc.p, ix->r_occ, ix->r_occ_words * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->ix_rmaj.p, ix->r_occ_major, ix->r_occ_major_words * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->ix_rsa.p, ix->r_sa_sharp, ix->r_n_sa_sharp * 4, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    FmIndexDev &f = h->fm;
    f.cbwt = h->ix_cbwt.as<uint32_t>(); f.c_primary = ix->c_primary; f.c_seq_len = ix->c_seq_len;
    for (int i = 0; i < 5; ++i) f.c_L2[i] = ix->c_L2[i];
    f.c_sa = h->ix_csa.as<uint32_t>(); f.c_sa_intv = ix->c_sa_intv; f.c_n_sa = ix->c_n_sa;
    f.lkt = h->ix_lkt.as<uint32_t>(); f.l_lkt = (int)ix->lkt_len;
    f.r_bwt = h->ix_rbwt.as<uint32_t>(); f.r_occ = h->ix_rocc.as<uint32_t>(); f.r_occ_major = h->ix_rmaj.as<uint32_t>();
    f.r_sa_sharp = h->ix_rsa.as<uint32_t>(); f.r_n_sa_sharp = (uint32_t)ix->r_n_sa_sharp;
    for (int i = 0; i < 6; ++i) f.r_cum[i] = ix->r_cum[i];
    f.r_inv_sa0 = ix->r_inv_sa0; f.r_text_len = ix->r_text_len;
    // the device's own layouts of the two BWTs (seed.cu): one aligned line per occurrence count
    CU(h->ix_c32.need(fm_c32_entries(ix->c_bwt_words) * 32 + 64));
    CU(h->ix_r64.need(fm_r64_entries(ix->r_text_len) * 64 + 64));
    f.c32 = h->ix_c32.as<uint4>(); f.r64 = h->ix_r64.as<uint4>();
    CU(launch_build_dense_index(f, ix->c_bwt_words, rpad, h->ix_c32.as<uint4>(), h->ix_r64.as<uint4>(), st));
    CU(cudaStreamSynchronize(st));
    // the file-layout BWT words and count tables have served their purpose (the sampled suffix arrays, the lookup table
    // and the 4-bit characters -- for Rbwt_bwt2nt -- stay)
    h->ix_cbwt.release(); h->ix_rocc.release(); h->ix_rmaj.release();
    f.cbwt = nullptr; f.r_occ = nullptr; f.r_occ_major = nullptr;
    h->launches += 2;
    h->have_index = true;
    return SALT_OK;
}

int salt_b200_seed_locate(salt_b200_t *h, int slot, const salt_seed_opt_t *opt, uint32_t *offs0, uint32_t *offs1,
                          uint32_t *loci0, size_t cap0, uint32_t *loci1, size_t cap1, size_t *n0, size_t *n1)
{
    if (int rc = use_device(h)) return rc;
    if (slot < 0 || slot >= SALT_SLOTS) return fail(SALT_ERR_ARG, "slot out of range");
    Slot &s = h->slot[slot];
    if (s.pending) return fail(SALT_ERR_ARG, "slot has a verify in flight: call salt_b200_verify_wait first");
    if (!s.n_reads) return fail(SALT_ERR_ARG, "no reads set");
    if (int rc = seed_enqueue(h, slot, opt)) return rc;
    if (int rc = seed_finish(h, slot)) return rc;
    if (n0) *n0 = s.seeded_n0;
    if (n1) *n1 = s.seeded_n1;
    if ((loci0 && cap0 < s.seeded_n0) || (loci1 && cap1 < s.seeded_n1)) return fail(SALT_ERR_NOMEM, "loci buffer too small for the lists");
    const size_t m1 = (size_t)s.n_reads + 1;
    if (offs0) CU(cudaMemcpyAsync(offs0, s.d_coffs(0), m1 * 4, cudaMemcpyDeviceToHost, s.stream));
    if (offs1) CU(cudaMemcpyAsync(offs1, s.d_coffs(1), m1 * 4, cudaMemcpyDeviceToHost, s.stream));
    if (loci0 && s.seeded_n0) CU(cudaMemcpyAsync(loci0, s.c_loci0.p, s.seeded_n0 * 4, cudaMemcpyDeviceToHost, s.stream));
    if (loci1 && s.seeded_n1) CU(cudaMemcpyAsync(loci1, s.c_loci1.p, s.seeded_n1 * 4, cudaMemcpyDeviceToHost, s.stream));
    CU(cudaStreamSynchronize(s.stream));
    return SALT_OK;
}

int salt_b200_seed_status(salt_b200_t *h, int slot, uint8_t *st0, uint8_t *st1)
{
    if (int rc = use_device(h)) return rc;
    if (slot < 0 || slot >= SALT_SLOTS) return fail(SALT_ERR_ARG, "slot out of range");
    Slot &s = h->slot[slot];
    if (s.seeded != 2 || !st0 || !st1) return fail(SALT_ERR_ARG, "slot holds no seeded lists, or null buffer");
    if (!s.n_reads) return SALT_OK;
    CU(cudaMemcpyAsync(st0, s.sd_status.p, s.n_reads, cudaMemcpyDeviceToHost, s.stream));
    CU(cudaMemcpyAsync(st1, s.sd_status.as<uint8_t>() + s.n_reads, s.n_reads, cudaMemcpyDeviceToHost, s.stream));
    CU(cudaStreamSynchronize(s.stream));
    return SALT_OK;
}

int salt_b200_verify_seeded(salt_b200_t *h, int slot, int nogap_T0, int lv_T0, salt_verify_out_t *rec,
                            int8_t *acc0, int8_t *acc1, char *cigars, int cigar_stride)
{
    if (int rc = use_device(h)) return rc;
    if (slot < 0 || slot >= SALT_SLOTS) return fail(SALT_ERR_ARG, "slot out of range");
    Slot &s = h->slot[slot];
    if (s.pending) return fail(SALT_ERR_ARG, "slot has a verify in flight: call salt_b200_verify_wait first");
    if (!rec) return fail(SALT_ERR_ARG, "null buffer");
    if (s.seeded != 2) return fail(SALT_ERR_ARG, "slot holds no seeded lists: call salt_b200_seed_locate first");
    if (!s.n_reads) return SALT_OK;
    if (int rc = run_and_download(h, slot, s.seeded_n0, s.seeded_n1, nogap_T0, lv_T0, rec, acc0, acc1, cigars, cigar_stride)) return rc;
    return finish_verify(h, slot);
}

int salt_b200_align_batch_packed(salt_b200_t *h, const salt_packed_chunk_t *pc, const salt_seed_opt_t *opt, uint32_t chunk_reads,
                                 int nogap_T0, int lv_T0, salt_verify_out_t *rec, char *cigars, int cigar_stride)
{
    if (int rc = use_device(h)) return rc;
    if (int rc = check_packed(pc, false)) return rc;
    if (!rec) return fail(SALT_ERR_ARG, "null buffer");
    if (chunk_reads == 0) chunk_reads = 100000;          // N_SEQS, aln.h:27
    int rc = SALT_OK;
    PackedView v; v.pc = pc; v.base_pos = pc->base_start;
    while (v.np_lo < pc->n_n && pc->n_pos[v.np_lo] < v.base_pos) ++v.np_lo;
    v.np_hi = v.np_lo;
    // chunk k: upload + seed + locate are queued first; the totals of chunk k-1 are awaited only then, so the device
    // always has the next chunk's seeding to do while the host sizes and queues the previous chunk's verification
    int prev_slot = -1; uint32_t prev_b = 0;
    auto verify_prev = [&]() -> int {
        if (prev_slot < 0) return SALT_OK;
        Slot &sp = h->slot[prev_slot];
        if (int r2 = seed_finish(h, prev_slot)) return r2;
        const int ps = prev_slot; prev_slot = -1;
        if (!sp.n_reads) return SALT_OK;
        return run_and_download(h, ps, sp.seeded_n0, sp.seeded_n1, nogap_T0, lv_T0, rec + prev_b, nullptr, nullptr,
                                cigars ? cigars + (size_t)prev_b * cigar_stride : nullptr, cigar_stride);
    };
    uint32_t k = 0;
    for (uint32_t b = 0; b < pc->n_reads && rc == SALT_OK; b += chunk_reads, ++k) {
        const int si = (int)(k % SALT_SLOTS);
        if ((rc = finish_verify(h, si)) != SALT_OK) break;
        const uint32_t m = pc->n_reads - b < chunk_reads ? pc->n_reads - b : chunk_reads;
        PackedView w; w.pc = pc; w.first = b; w.base_pos = v.base_pos + v.n_bases; w.np_lo = v.np_hi;
        if ((rc = grow_view(w, m, false)) != SALT_OK) break;
        v = w;
        if ((rc = load_packed(h, h->slot[si], v, false)) != SALT_OK) break;
        if ((rc = seed_enqueue(h, si, opt)) != SALT_OK) break;
        if ((rc = verify_prev()) != SALT_OK) break;
        prev_slot = si; prev_b = b;
    }
    if (rc == SALT_OK) rc = verify_prev();
    for (int si = 0; si < SALT_SLOTS; ++si) { const int r2 = finish_verify(h, si); if (rc == SALT_OK) rc = r2; }
    return rc;
}

int salt_b200_profile(salt_b200_t *h, int enable)
{
    if (int rc = use_device(h)) return rc;
    h->profiling = enable != 0;
    h->have_ssw_prof = false;
    for (int i = 0; i < SALT_SLOTS; ++i) h->slot[i].have_verify_prof = false;
    return SALT_OK;
}

int salt_b200_profile_read(salt_b200_t *h, float *ms /* [12] */)
{
    if (int rc = use_device(h)) return rc;
    if (!ms) return fail(SALT_ERR_ARG, "null buffer");
    Slot &s = h->slot[0];
    CU(cudaStreamSynchronize(s.stream));
    for (int i = 0; i < 12; ++i) ms[i] = -1.f;
    if (s.have_verify_prof)
        for (int i = 0; i < 6; ++i) CU(cudaEventElapsedTime(&ms[i], s.ev_verify[i], s.ev_verify[i + 1]));
    if (h->have_ssw_prof)
        for (int i = 0; i < 6; ++i) CU(cudaEventElapsedTime(&ms[6 + i], h->ev_ssw[i], h->ev_ssw[i + 1]));
    return SALT_OK;
}

}  // extern "C"
#pragma GCC visibility pop
