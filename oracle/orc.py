"""ctypes front-ends for the CPU checkers.  TEST INFRASTRUCTURE ONLY.

`Oracle` wraps oracle/liboracle.so (our plain-C restatement, oracle.c).
`Ref`    wraps oracle/_ref/libsaltref.so -- the reference's own unmodified
         editdistance.c / LandauVishkin.c / ssw.c compiled by oracle/Makefile.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  Nothing here reads /root/reference at
run time; the libraries are prebuilt files.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
u8p = C.POINTER(C.c_uint8)
i8p = C.POINTER(C.c_int8)
u32p = C.POINTER(C.c_uint32)


def build(force=False):
    """(Re)build liboracle.so and, when the reference tree is present, oracle/_ref."""
    if force or not os.path.exists(os.path.join(HERE, "liboracle.so")) or \
            os.path.getmtime(os.path.join(HERE, "liboracle.so")) < os.path.getmtime(os.path.join(HERE, "oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, os.path.join(HERE, "liboracle.so")])
    if os.path.isdir("/root/reference/Align_src") and (force or not all(
            os.path.exists(os.path.join(HERE, "_ref", f)) for f in ("libsaltref.so", "libsaltref_sam.so", "libsaltref_pair.so"))):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _p(a, t):
    return a.ctypes.data_as(t)


class AlignRec(C.Structure):
    _fields_ = [("score1", C.c_uint16), ("score2", C.c_uint16), ("ref_begin1", C.c_int32),
                ("ref_end1", C.c_int32), ("read_begin1", C.c_int32), ("read_end1", C.c_int32),
                ("ref_end2", C.c_int32), ("cigarLen", C.c_int32)]

    def astuple(self):
        return (self.score1, self.score2, self.ref_begin1, self.ref_end1, self.read_begin1,
                self.read_end1, self.ref_end2, self.cigarLen)


class Hit(C.Structure):
    _fields_ = [("pos", C.c_uint32), ("n_diff", C.c_uint8), ("is_gap", C.c_uint8), ("strand", C.c_uint16)]


class VerifyRec(C.Structure):
    _fields_ = [("pos", C.c_uint32), ("strand", C.c_int), ("n_diff", C.c_uint8), ("is_gap", C.c_uint8),
                ("b0", C.c_int), ("b1", C.c_int), ("mapq", C.c_uint32),
                ("n_hits", C.c_int * 2), ("n_alt", C.c_int * 2)]


class Oracle:
    def __init__(self):
        build()
        self.lib = L = C.CDLL(os.path.join(HERE, "liboracle.so"))
        L.orc_ed_mismatch.argtypes = [u32p, C.c_uint32, u8p, C.c_uint32, C.c_int]
        L.orc_ed_diff.argtypes = [u32p, C.c_uint32, C.c_uint32, C.c_uint32, u8p, C.c_uint32, C.c_int]
        L.orc_ed_diff_withcigar.argtypes = [u32p, C.c_uint32, C.c_uint32, u8p, C.c_uint32, C.c_int, C.c_char_p, C.c_int]
        L.orc_lv.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int]
        L.orc_lv_cigar.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.orc_ssw_align.argtypes = [i8p, C.c_int, i8p, C.c_int, i8p, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.POINTER(AlignRec), u32p, C.c_int]
        L.orc_rescue_mixref.argtypes = [u32p, C.c_uint32, C.c_uint32, u8p, C.c_int, i8p, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.POINTER(AlignRec), u32p, C.c_int]
        L.orc_rescue_pac.argtypes = [u8p, C.c_uint32, C.c_uint32, u8p, C.c_int, i8p, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.POINTER(AlignRec), u32p, C.c_int]
        L.orc_md_nm.argtypes = [u32p, u8p, C.c_uint32, u8p, C.c_int, C.c_uint32, C.c_uint32, C.c_char_p, C.c_char_p, C.c_int]
        L.orc_score_mat2.argtypes = [i8p]
        L.orc_score_mat.argtypes = [i8p]
        L.orc_mixref_put_seq.argtypes = [u32p, C.c_uint32, C.c_char_p, C.c_uint32]
        L.orc_allele_mask.argtypes = [C.c_char_p]
        L.orc_allele_mask.restype = C.c_uint
        L.orc_mixref_or_snp.argtypes = [u32p, C.c_uint32, C.c_uint]
        L.orc_verify_read.argtypes = [u32p, C.c_uint32, u8p, u8p, C.c_uint32, u32p, C.c_uint32, u32p, C.c_uint32,
                                      C.c_int, C.c_int, C.c_int, C.POINTER(VerifyRec),
                                      C.POINTER(Hit), C.POINTER(Hit), C.POINTER(Hit), C.POINTER(Hit)]

    # --- threaded CPU baselines (bench.py) ----------------------------------
    def verify_batch(self, mixref, l, codes, roffs, offs0, loci0, offs1, loci1, nogap_T0=3, lv_T0=-1,
                     n_threads=1, ref=None, want_acc=True, cigar_stride=128):
        """The reference's verification loop over a batch on n_threads host threads.
        ref=None times this file's port; ref=Ref() times the reference's own compiled functions.
        Returns (seconds, recs, acc0, acc1, cigars)."""
        L = self.lib
        L.orc_verify_batch.restype = C.c_double
        L.orc_verify_batch.argtypes = [u32p, C.c_uint32, u8p, u32p, C.c_uint32, u32p, u32p, u32p, u32p, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(VerifyRec), C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int]
        n = len(roffs) - 1
        recs = (VerifyRec * n)()
        acc0 = np.empty(len(loci0), np.int8) if want_acc else None
        acc1 = np.empty(len(loci1), np.int8) if want_acc else None
        cig = np.zeros((n, cigar_stride), np.uint8)
        if ref is not None:
            fn = [C.cast(getattr(ref.lib, f), C.c_void_p) for f in ("ed_mismatch", "ed_diff", "ed_diff_withcigar")]
        else:
            fn = [None, None, C.cast(L.orc_ed_diff_withcigar_ref_abi, C.c_void_p)]
        sec = L.orc_verify_batch(_p(mixref, u32p), int(l), _p(codes, u8p), _p(roffs, u32p), n,
                                 _p(offs0, u32p), _p(loci0, u32p), _p(offs1, u32p), _p(loci1, u32p),
                                 nogap_T0, lv_T0, fn[0], fn[1], fn[2], int(n_threads), recs,
                                 None if acc0 is None else acc0.ctypes.data, None if acc1 is None else acc1.ctypes.data,
                                 cig.ctypes.data, cigar_stride)
        return sec, recs, acc0, acc1, cig

    def ssw_batch(self, mixref, codes, L_read, starts, ends, mat, gapO=3, gapE=1, n_threads=1, ref=None):
        """ssw_init + ssw_align(flag=2) + destroys per task on n_threads host threads; returns (seconds, checksum)."""
        L = self.lib
        L.orc_ssw_batch.restype = C.c_double
        L.orc_ssw_batch.argtypes = [u32p, u8p, C.c_int, u32p, u32p, C.c_uint32, i8p, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]
        fn = [None] * 4
        if ref is not None:
            fn = [C.cast(getattr(ref.lib, f), C.c_void_p) for f in ("ssw_init", "ssw_align", "init_destroy", "align_destroy")]
        chk = C.c_uint64(0)
        mat = np.ascontiguousarray(mat, np.int8)
        sec = L.orc_ssw_batch(_p(mixref, u32p), _p(codes, u8p), int(L_read), _p(starts, u32p), _p(ends, u32p), len(starts),
                              _p(mat, i8p), gapO, gapE, fn[0], fn[1], fn[2], fn[3], int(n_threads), C.byref(chk))
        return sec, chk.value

    # --- scoring matrices -------------------------------------------------
    def score_mat2(self, pad=-3):
        """256 entries + one pad entry (the reference reads mat[256] for read-N vs mask 15)."""
        m = np.zeros(257, np.int8)
        self.lib.orc_score_mat2(_p(m, i8p))
        m[256] = pad
        return m

    def score_mat(self):
        m = np.zeros(25, np.int8)
        self.lib.orc_score_mat(_p(m, i8p))
        return m

    # --- per-pair kernels -------------------------------------------------
    def md_nm(self, mixref, pac, l_pac, seq, pos, seq_start, cigar, cap=4096):
        """text sam_add_md_nm appends for the read as aligned (`seq` = query->seq or ->rseq); -3 = reference assert"""
        buf = C.create_string_buffer(cap)
        n = self.lib.orc_md_nm(_p(mixref, u32p), _p(pac, u8p), int(l_pac), _p(seq, u8p), len(seq), int(pos), int(seq_start),
                               cigar.encode(), buf, cap)
        return n if n < 0 else buf.value.decode()

    def ed_mismatch(self, mixref, pos, seq, max_err):
        return self.lib.orc_ed_mismatch(_p(mixref, u32p), int(pos), _p(seq, u8p), len(seq), int(max_err))

    def ed_diff(self, mixref, l, pos, seq, k, l_ref=None):
        l_ref = len(seq) + 4 if l_ref is None else l_ref
        return self.lib.orc_ed_diff(_p(mixref, u32p), int(l), int(pos), int(l_ref), _p(seq, u8p), len(seq), int(k))

    def ed_diff_withcigar(self, mixref, pos, seq, k, buflen=128, l_ref=None):
        l_ref = len(seq) + 4 if l_ref is None else l_ref
        buf = C.create_string_buffer(b"\0" * (buflen + 8), buflen + 8)
        r = self.lib.orc_ed_diff_withcigar(_p(mixref, u32p), int(pos), int(l_ref), _p(seq, u8p), len(seq), int(k), buf, buflen)
        return r, buf.raw[:buflen].split(b"\0")[0].decode()

    def lv(self, text, pattern, k):
        t = np.concatenate([text, np.zeros(16, np.uint8)]); p = np.concatenate([pattern, np.zeros(16, np.uint8)])
        return self.lib.orc_lv(_p(t, u8p), len(text), _p(p, u8p), len(pattern), int(k))

    def lv_cigar(self, text, pattern, k, buflen=128):
        t = np.concatenate([text, np.zeros(16, np.uint8)]); p = np.concatenate([pattern, np.zeros(16, np.uint8)])
        buf = C.create_string_buffer(b"\0" * (buflen + 8), buflen + 8)
        r = self.lib.orc_lv_cigar(_p(t, u8p), len(text), _p(p, u8p), len(pattern), int(k), buf, buflen)
        return r, buf.raw[:buflen].split(b"\0")[0].decode()

    def ssw_align(self, read, mat, n, ref, gapO=3, gapE=1, flag=2, filters=0, filterd=20, maskLen=None, cigar_cap=512):
        read = np.ascontiguousarray(read, np.int8); ref = np.ascontiguousarray(ref, np.int8)
        mat = np.ascontiguousarray(mat, np.int8)
        maskLen = len(read) // 2 if maskLen is None else maskLen
        rec = AlignRec(); cig = np.zeros(cigar_cap, np.uint32)
        rc = self.lib.orc_ssw_align(_p(read, i8p), len(read), _p(mat, i8p), n, _p(ref, i8p), len(ref), gapO, gapE,
                                    flag, filters, filterd, maskLen, C.byref(rec), _p(cig, u32p), cigar_cap)
        return rc, rec.astuple(), cig[:min(rec.cigarLen, cigar_cap)].copy()

    def rescue_mixref(self, mixref, start, end, seq, mat, gapO=3, gapE=1, filters=0, filterd=20, cigar_cap=512):
        rec = AlignRec(); cig = np.zeros(cigar_cap, np.uint32)
        mat = np.ascontiguousarray(mat, np.int8)
        rc = self.lib.orc_rescue_mixref(_p(mixref, u32p), int(start), int(end), _p(seq, u8p), len(seq), _p(mat, i8p),
                                        gapO, gapE, filters, filterd, C.byref(rec), _p(cig, u32p), cigar_cap)
        return rc, rec.astuple(), cig[:min(rec.cigarLen, cigar_cap)].copy()

    def rescue_pac(self, pac, start, end, seq, mat, gapO=3, gapE=1, filters=0, filterd=20, cigar_cap=512):
        rec = AlignRec(); cig = np.zeros(cigar_cap, np.uint32)
        mat = np.ascontiguousarray(mat, np.int8)
        rc = self.lib.orc_rescue_pac(_p(pac, u8p), int(start), int(end), _p(seq, u8p), len(seq), _p(mat, i8p),
                                     gapO, gapE, filters, filterd, C.byref(rec), _p(cig, u32p), cigar_cap)
        return rc, rec.astuple(), cig[:min(rec.cigarLen, cigar_cap)].copy()

    # --- reference construction -------------------------------------------
    def build_mixref(self, records, snp_rows):
        """records: list of (name, bases str); snp_rows: list of (chrom, pos1, 'A/G', ref) in file order.
        SNP rows are consumed one contiguous same-chrom block per record, in order, without checking
        the name (Index_src/hapmap.c:65-91, mixRef.c:147)."""
        tot = sum(len(s) for _, s in records)
        words = np.zeros((tot + 7) // 8, np.uint32)
        off = 0; ri = 0
        for _, s in records:
            self.lib.orc_mixref_put_seq(_p(words, u32p), off, s.encode(), len(s))
            if ri < len(snp_rows):
                chrom = snp_rows[ri][0]
                while ri < len(snp_rows) and snp_rows[ri][0] == chrom:
                    _, pos1, alleles, _ = snp_rows[ri]
                    self.lib.orc_mixref_or_snp(_p(words, u32p), off + int(pos1) - 1, self.lib.orc_allele_mask(alleles.encode()))
                    ri += 1
            off += len(s)
        return words, tot

    # --- acceptance logic -------------------------------------------------
    def verify_read(self, mixref, l, seq, rseq, loci0, loci1, nogap_T0=3, lv_T0=None, max_hits=5):
        lv_T0 = len(seq) // 10 if lv_T0 is None else lv_T0
        loci0 = np.ascontiguousarray(loci0, np.uint32); loci1 = np.ascontiguousarray(loci1, np.uint32)
        rec = VerifyRec()
        h0 = (Hit * (2 * len(loci0) + 1))(); h1 = (Hit * (2 * len(loci1) + 1))()
        a0 = (Hit * (max_hits + 1))(); a1 = (Hit * (max_hits + 1))()
        self.lib.orc_verify_read(_p(mixref, u32p), int(l), _p(seq, u8p), _p(rseq, u8p), len(seq),
                                 _p(loci0, u32p), len(loci0), _p(loci1, u32p), len(loci1),
                                 nogap_T0, lv_T0, max_hits, C.byref(rec), h0, h1, a0, a1)
        hits = [[(h.pos, h.n_diff, h.is_gap, h.strand) for h in hh[:n]] for hh, n in ((h0, rec.n_hits[0]), (h1, rec.n_hits[1]))]
        alts = [[(h.pos, h.n_diff, h.is_gap, h.strand) for h in hh[:n]] for hh, n in ((a0, rec.n_alt[0]), (a1, rec.n_alt[1]))]
        prim = (rec.pos, rec.strand, rec.n_diff, rec.is_gap, rec.b0, rec.b1, rec.mapq)
        return prim, hits, alts


class SAlign(C.Structure):     # ssw.h:37-47
    _fields_ = [("score1", C.c_uint16), ("score2", C.c_uint16), ("ref_begin1", C.c_int32),
                ("ref_end1", C.c_int32), ("read_begin1", C.c_int32), ("read_end1", C.c_int32),
                ("ref_end2", C.c_int32), ("cigar", u32p), ("cigarLen", C.c_int32)]


def ref_available(o0=False):
    return os.path.exists(os.path.join(HERE, "_ref", "libsaltref_O0.so" if o0 else "libsaltref.so"))


class Ref:
    """The reference's own compiled hot-path functions (unmodified sources)."""

    def __init__(self, o0=False):
        build()
        self.lib = L = C.CDLL(os.path.join(HERE, "_ref", "libsaltref_O0.so" if o0 else "libsaltref.so"))
        L.ed_mismatch.argtypes = [u32p, C.c_uint32, u8p, C.c_uint32, C.c_int]
        L.ed_diff.argtypes = [u32p, C.c_uint32, C.c_uint32, C.c_uint32, u8p, C.c_uint32, C.c_int]
        L.ed_diff_withcigar.argtypes = [u32p, C.c_uint32, C.c_uint32, u8p, C.c_uint32, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int]
        L.computeEditDistance.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int]
        L.computeEditDistanceWithCigar.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int]
        L.ssw_init.restype = C.c_void_p
        L.ssw_init.argtypes = [i8p, C.c_int32, i8p, C.c_int32, C.c_int8]
        L.ssw_align.restype = C.POINTER(SAlign)
        L.ssw_align.argtypes = [C.c_void_p, i8p, C.c_int32, C.c_uint8, C.c_uint8, C.c_uint8, C.c_uint16, C.c_int32, C.c_int32]
        L.init_destroy.argtypes = [C.c_void_p]
        L.align_destroy.argtypes = [C.POINTER(SAlign)]
        self.score_mat2_ref = np.ctypeslib.as_array((C.c_int8 * 256).in_dll(L, "score_mat2")).copy()
        self.score_mat_ref = np.ctypeslib.as_array((C.c_int8 * 25).in_dll(L, "score_mat")).copy()

    def ed_mismatch(self, mixref, pos, seq, max_err):
        return self.lib.ed_mismatch(_p(mixref, u32p), int(pos), _p(seq, u8p), len(seq), int(max_err))

    def ed_diff(self, mixref, l, pos, seq, k, l_ref=None):
        l_ref = len(seq) + 4 if l_ref is None else l_ref
        return self.lib.ed_diff(_p(mixref, u32p), int(l), int(pos), int(l_ref), _p(seq, u8p), len(seq), int(k))

    def ed_diff_withcigar(self, mixref, pos, seq, k, buflen=128, l_ref=None):
        l_ref = len(seq) + 4 if l_ref is None else l_ref
        buf = C.create_string_buffer(b"\0" * (buflen + 8), buflen + 8)
        r = self.lib.ed_diff_withcigar(_p(mixref, u32p), int(pos), int(l_ref), _p(seq, u8p), len(seq), int(k), buf, buflen, 1, 0)
        return r, buf.raw[:buflen].split(b"\0")[0].decode()

    def lv(self, text, pattern, k):
        # the reference reads 8-byte words past both ends; give it the zero padding ed_diff's calloc gives
        t = np.concatenate([text, np.zeros(64, np.uint8)]); p = np.concatenate([pattern, np.zeros(64, np.uint8)])
        return self.lib.computeEditDistance(t.tobytes(), len(text), p.tobytes(), len(pattern), int(k))

    def lv_cigar(self, text, pattern, k, buflen=128):
        t = np.concatenate([text, np.zeros(64, np.uint8)]); p = np.concatenate([pattern, np.zeros(64, np.uint8)])
        buf = C.create_string_buffer(b"\0" * (buflen + 8), buflen + 8)
        r = self.lib.computeEditDistanceWithCigar(t.tobytes(), len(text), p.tobytes(), len(pattern), int(k), buf, buflen, 1, 0)
        return r, buf.raw[:buflen].split(b"\0")[0].decode()

    def ssw_align(self, read, mat, n, ref, gapO=3, gapE=1, flag=2, filters=0, filterd=20, maskLen=None):
        read = np.ascontiguousarray(read, np.int8); ref = np.ascontiguousarray(ref, np.int8)
        mat = np.ascontiguousarray(mat, np.int8)
        maskLen = len(read) // 2 if maskLen is None else maskLen
        prof = self.lib.ssw_init(_p(read, i8p), len(read), _p(mat, i8p), n, 1)
        # keep stderr quiet about maskLen < 15 -- the caller chooses
        res = self.lib.ssw_align(prof, _p(ref, i8p), len(ref), gapO, gapE, flag, filters, filterd, maskLen)
        if not res:
            self.lib.init_destroy(prof)
            return -1, None, None
        r = res.contents
        cig = np.array([r.cigar[i] for i in range(r.cigarLen)], np.uint32) if r.cigarLen > 0 else np.zeros(0, np.uint32)
        tup = (r.score1, r.score2, r.ref_begin1, r.ref_end1, r.read_begin1, r.read_end1, r.ref_end2, r.cigarLen)
        self.lib.align_destroy(res)
        self.lib.init_destroy(prof)
        return 0, tup, cig

    def rescue_mixref(self, mixref, start, end, seq, mat, gapO=3, gapE=1, filters=0, filterd=20):
        """snpaln_sw_snpaware's data preparation (alnpe.c:276-293) around the reference's ssw."""
        idx = np.arange(start, end + 1, dtype=np.int64)
        ref = ((mixref[idx >> 3] >> (4 * (idx & 7)).astype(np.uint32)) & 15).astype(np.int8)
        read = (1 << seq.astype(np.int32)).astype(np.int8)
        return self.ssw_align(read, mat, 16, ref, gapO, gapE, 2, filters, filterd, len(seq) // 2)

    def rescue_pac(self, pac, start, end, seq, mat, gapO=3, gapE=1, filters=0, filterd=20):
        idx = np.arange(start, end + 1, dtype=np.int64)
        ref = ((pac[idx >> 2] >> ((~idx & 3) << 1).astype(np.uint8)) & 3).astype(np.int8)
        return self.ssw_align(seq.astype(np.int8), mat, 5, ref, gapO, gapE, 2, filters, filterd, len(seq) // 2)


class RefSam:
    """The reference's own sam_add_md_nm (sam.c:246) behind oracle/dropin/sam_harness.c."""

    def __init__(self):
        build()
        self.lib = C.CDLL(os.path.join(HERE, "_ref", "libsaltref_sam.so"))
        self.lib.ref_sam_md_nm.argtypes = [u32p, C.c_uint32, u8p, u8p, u8p, C.c_int, C.c_uint32, C.c_int, C.c_uint32,
                                           C.c_char_p, C.c_char_p, C.c_int]

    def md_nm(self, mixref, l, pac, seq, rseq, pos, strand, seq_start, cigar, cap=4096):
        buf = C.create_string_buffer(cap)
        self.lib.ref_sam_md_nm(_p(mixref, u32p), int(l), _p(pac, u8p), _p(seq, u8p), _p(rseq, u8p), len(seq), int(pos),
                               int(strand), int(seq_start), cigar.encode(), buf, cap)
        return buf.value.decode()


class RefPair:
    """The reference's own pairing2 / pairing_singleton (alnpe.c) behind oracle/dropin/pair_harness.c; rescues always fail."""

    def __init__(self):
        build()
        self.lib = C.CDLL(os.path.join(HERE, "_ref", "libsaltref_pair.so"))
        self.lib.ref_pairing.argtypes = [C.c_uint32, C.c_int, C.c_int, u32p, C.POINTER(C.c_int), u32p, C.POINTER(C.c_int), C.c_int,
                                         u32p, C.POINTER(C.c_int), C.c_char_p]

    def pairing(self, l_pac, min_tlen, max_tlen, prim, l_seq, hits):
        """prim: [(pos, strand, n_diff, is_gap)] x2; hits[m][s] = [(pos, n_diff, is_gap), ...].
        Returns (paired, [(pos, strand, n_diff, is_gap, seq_start, seq_end)] x2, windows, cigars)."""
        mx = 16
        qf = np.array(prim, np.uint32).reshape(-1)
        ls = (C.c_int * 2)(*l_seq)
        hh = np.zeros((2, 2, mx, 3), np.uint32); nh = (C.c_int * 4)()
        for m in range(2):
            for s in range(2):
                nh[m * 2 + s] = len(hits[m][s])
                for i, h in enumerate(hits[m][s]):
                    hh[m, s, i] = h
        out_q = np.zeros(12, np.uint32); out_w = (C.c_int * 20)(); cg = C.create_string_buffer(128)
        r = self.lib.ref_pairing(int(l_pac), int(min_tlen), int(max_tlen), _p(qf, u32p), ls, _p(hh, u32p), nh, mx, _p(out_q, u32p), out_w, cg)
        n = r & 255
        wins = [tuple(out_w[i * 5 + j] & 0xFFFFFFFF if j >= 3 else out_w[i * 5 + j] for j in range(5)) for i in range(n)]
        cigs = [cg.raw[m * 64:(m + 1) * 64].split(b"\0")[0].decode() for m in range(2)]
        return r >> 8, [tuple(int(x) for x in out_q[m * 6:(m + 1) * 6]) for m in range(2)], wins, cigs


def ref_pair_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libsaltref_pair.so"))


def ref_sam_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libsaltref_sam.so"))


class RefIdx:
    """The reference's own build_mixRef (Index_src/mixRef.c:93), file based."""

    def __init__(self):
        build()
        self.lib = C.CDLL(os.path.join(HERE, "_ref", "libsaltref_idx.so"))
        self.lib.build_mixRef.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]

    def build_mixref(self, fa_path, snp_path, out_path):
        rc = self.lib.build_mixRef(fa_path.encode(), snp_path.encode(), out_path.encode())
        raw = np.fromfile(out_path, np.uint32)
        return raw[1:].copy(), int(raw[0]), rc


class SeedRef:
    """The reference's own seeding + locate and index loader (oracle/_ref/libsaltref_seed.so, built from the unmodified
    sources by oracle/Makefile around oracle/dropin/seed_harness.c): the oracle of row f1."""

    PATH = os.path.join(HERE, "_ref", "libsaltref_seed.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    def __init__(self, prefix):
        self.lib = C.CDLL(self.PATH)
        self.lib.seedref_open.restype = C.c_void_p
        self.lib.seedref_open.argtypes = [C.c_char_p]
        self.lib.seedref_close.argtypes = [C.c_void_p]
        self.lib.seedref_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]
        self.ix = self.lib.seedref_open(prefix.encode())

    def run(self, codes, roffs, l_seed, l_overlap, max_seed, max_locate, seed_only_ref=0, locate_mode=0, cap_per_read=None):
        """locate_mode 0: alnse_locate_alt (single-end program), 1: alnse_locate (paired-end program)"""
        codes = np.ascontiguousarray(codes, np.uint8).reshape(-1); roffs = np.ascontiguousarray(roffs, np.uint32)
        n = len(roffs) - 1
        self.lib.seedref_set_locate_mode(int(locate_mode))
        cap = n * (cap_per_read or max_locate) + 16
        offs0 = np.zeros(n + 1, np.uint32); offs1 = np.zeros(n + 1, np.uint32)
        loci0 = np.zeros(cap, np.uint32); loci1 = np.zeros(cap, np.uint32)
        rc = self.lib.seedref_run(self.ix, codes.ctypes.data, roffs.ctypes.data, n, l_seed, l_overlap if l_overlap > 0 else l_seed,
                                  max_seed, max_locate, seed_only_ref, offs0.ctypes.data, loci0.ctypes.data, cap,
                                  offs1.ctypes.data, loci1.ctypes.data, cap)
        self.lib.seedref_set_locate_mode(0)
        assert rc == 0
        return offs0, loci0[:offs0[-1]].copy(), offs1, loci1[:offs1[-1]].copy()

    def run_mt(self, codes, roffs, l_seed, l_overlap, max_seed, max_locate, n_threads, seed_only_ref=0):
        """the same on n_threads pthreads; returns (lists..., seconds)"""
        codes = np.ascontiguousarray(codes, np.uint8).reshape(-1); roffs = np.ascontiguousarray(roffs, np.uint32)
        n = len(roffs) - 1
        cap = n * 64 + 1024
        while True:
            offs0 = np.zeros(n + 1, np.uint32); offs1 = np.zeros(n + 1, np.uint32)
            loci0 = np.zeros(cap, np.uint32); loci1 = np.zeros(cap, np.uint32)
            sec = C.c_double(0.0)
            self.lib.seedref_run_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                                C.POINTER(C.c_double)]
            rc = self.lib.seedref_run_mt(self.ix, codes.ctypes.data, roffs.ctypes.data, n, l_seed, l_overlap if l_overlap > 0 else l_seed,
                                         max_seed, max_locate, seed_only_ref, n_threads, offs0.ctypes.data, loci0.ctypes.data, cap,
                                         offs1.ctypes.data, loci1.ctypes.data, cap, C.byref(sec))
            if rc == 0:
                return offs0, loci0[:offs0[-1]].copy(), offs1, loci1[:offs1[-1]].copy(), sec.value
            cap *= 4

    def close(self):
        if self.ix:
            self.lib.seedref_close(self.ix)
            self.ix = None
