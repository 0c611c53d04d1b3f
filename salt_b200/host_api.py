"""ctypes binding of libsalt_host.so (include/salt_host.h) -- the host-side C layer that re-stages
salt's per-chunk loop around the batched engine.  Forwarding only; used by tests and examples."""
import ctypes as C
import os

import numpy as np

from . import api

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(HERE, "libsalt_host.so")
SALT_MAX_HITS = 16


class HitT(C.Structure):          # hit_t, query.h:28-33
    _fields_ = [("pos", C.c_uint32), ("n_diff", C.c_uint8), ("is_gap", C.c_uint8), ("strand", C.c_uint16)]


class ReadResultT(C.Structure):
    _fields_ = [("pos", C.c_uint32), ("strand", C.c_uint8), ("n_diff", C.c_uint8), ("is_gap", C.c_uint8),
                ("b0", C.c_int), ("b1", C.c_int), ("mapq", C.c_uint32), ("n_alt", C.c_int * 2),
                ("alt", (HitT * SALT_MAX_HITS) * 2), ("cigar", C.c_char * 128)]


class RescueT(C.Structure):       # salt_rescue_t
    _fields_ = [("mate", C.c_int), ("strand", C.c_int), ("flavour", C.c_int), ("start", C.c_uint32), ("end", C.c_uint32)]


class PairPlanT(C.Structure):     # salt_pair_plan_t
    _fields_ = [("paired", C.c_int), ("hit", HitT * 2), ("n_win", C.c_int), ("win", RescueT * 2)]


class SswOutT(C.Structure):       # salt_ssw_out_t
    _fields_ = [("score1", C.c_uint16), ("score2", C.c_uint16), ("ref_begin1", C.c_int32), ("ref_end1", C.c_int32),
                ("read_begin1", C.c_int32), ("read_end1", C.c_int32), ("ref_end2", C.c_int32), ("cigarLen", C.c_int32)]


class MateFinalT(C.Structure):    # salt_mate_final_t
    _fields_ = [("pos", C.c_uint32), ("strand", C.c_uint8), ("n_diff", C.c_uint8), ("is_gap", C.c_uint8),
                ("seq_start", C.c_uint32), ("seq_end", C.c_uint32), ("b0", C.c_int), ("b1", C.c_int), ("mapq", C.c_uint32),
                ("cigar_kind", C.c_int), ("cigar", C.c_char * 256)]


class PairFinalT(C.Structure):    # salt_pair_final_t
    _fields_ = [("mate", MateFinalT * 2), ("paired", C.c_int)]


class PeStatsT(C.Structure):      # salt_pe_stats_t
    _fields_ = [("pairs", C.c_size_t), ("proper", C.c_size_t), ("windows16", C.c_size_t), ("windows5", C.c_size_t),
                ("rescued", C.c_size_t), ("promoted", C.c_size_t), ("declined", C.c_size_t),
                ("ms_plan", C.c_double), ("ms_ssw", C.c_double), ("ms_apply", C.c_double), ("ms_tail", C.c_double)]


def pair_apply(L, plan_c, r0, l0, r1, l1, ssw, cigars, filters=0, filterd=20, stride=8):
    """ssw: list of (score1, score2, ref_begin1, ref_end1, read_begin1, read_end1); cigars: list of [(len, op), ...]"""
    n = len(ssw)
    so = (SswOutT * max(n, 1))(); cg = (C.c_uint32 * (max(n, 1) * stride))()
    for w, (a, ops) in enumerate(zip(ssw, cigars)):
        so[w].score1, so[w].score2, so[w].ref_begin1, so[w].ref_end1, so[w].read_begin1, so[w].read_end1 = a
        so[w].cigarLen = len(ops)
        for j, (ln, op) in enumerate(ops):
            cg[w * stride + j] = (ln << 4) | op
    out = (MateFinalT * 2)()
    rc = L.salt_pair_apply(C.byref(plan_c), C.byref(r0), int(l0), C.byref(r1), int(l1), so, cg, stride, int(filters), int(filterd), out)
    return rc, [(o.pos, o.strand, o.n_diff, o.is_gap, o.seq_start, o.seq_end, o.b0, o.b1, o.mapq, o.cigar_kind, o.cigar.decode()) for o in out]


def make_result(pos=0xFFFFFFFF, strand=3, n_diff=255, is_gap=255, alt0=(), alt1=()):
    """a salt_read_result_t from (pos, strand, n_diff, is_gap) and alternates [(pos, n_diff, is_gap), ...] per strand"""
    r = ReadResultT()
    r.pos, r.strand, r.n_diff, r.is_gap = pos, strand, n_diff, is_gap
    for s, alts in enumerate((alt0, alt1)):
        r.n_alt[s] = len(alts)
        for j, (p, nd, g) in enumerate(alts):
            r.alt[s][j].pos, r.alt[s][j].n_diff, r.alt[s][j].is_gap, r.alt[s][j].strand = p, nd, g, s
    return r


def pair_plan(L, r0, l0, r1, l1, min_tlen, max_tlen, l_pac, raw=False):
    plan = PairPlanT()
    rc = L.salt_pair_plan(C.byref(r0), int(l0), C.byref(r1), int(l1), int(min_tlen), int(max_tlen), int(l_pac), C.byref(plan))
    wins = [(w.mate, w.strand, w.flavour, w.start, w.end) for w in plan.win[:plan.n_win]]
    hits = [(h.pos, h.strand, h.n_diff, h.is_gap) for h in plan.hit] if plan.paired else None
    if raw:
        return plan
    return rc, hits, wins


def declare(L):
    vp, i32, u32, sz = C.c_void_p, C.c_int, C.c_uint32, C.c_size_t
    L.salt_chunk_new.restype = vp
    L.salt_chunk_new.argtypes = [u32, sz, sz]
    L.salt_chunk_free.argtypes = [vp]; L.salt_chunk_free.restype = None
    L.salt_chunk_reset.argtypes = [vp]; L.salt_chunk_reset.restype = None
    L.salt_chunk_n_reads.argtypes = [vp]; L.salt_chunk_n_reads.restype = u32
    L.salt_chunk_add_read.argtypes = [vp, vp, u32, vp, u32, vp, u32]
    L.salt_chunk_submit.argtypes = [vp, i32, vp, i32, i32]
    L.salt_chunk_wait.argtypes = [vp, i32, vp]
    L.salt_chunk_result.argtypes = [vp, u32, i32, C.POINTER(ReadResultT)]
    L.salt_chunk_hits.argtypes = [vp, u32, i32, C.POINTER(HitT), i32]
    L.salt_chunk_tail.argtypes = [vp, i32, vp]
    L.salt_chunk_md.argtypes = [vp, u32, C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_uint16)), C.POINTER(C.c_int)]
    L.salt_chunk_md.restype = C.c_char_p
    if hasattr(L, "salt_chunk_pair"):
        from .index_io import SeedOptT
        L.salt_chunk_add_reads.argtypes = [vp, vp, vp, u32, vp, vp, vp, vp]
        L.salt_chunk_seed_verify.argtypes = [vp, vp, C.POINTER(SeedOptT), i32, i32]
        L.salt_chunk_pair.argtypes = [vp, i32, vp, u32, u32, u32, i32, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, i32, C.POINTER(PeStatsT)]
        L.salt_host_set_threads.argtypes = [i32]; L.salt_host_set_threads.restype = None
        L.salt_host_set_grain.argtypes = [C.c_uint32]; L.salt_host_set_grain.restype = None
        L.salt_chunk_results.argtypes = [vp, i32, vp]
        L.salt_multi_init.restype = vp
        L.salt_multi_init.argtypes = [vp, u32, vp, C.c_int64, vp, i32]
        L.salt_multi_destroy.argtypes = [vp]; L.salt_multi_destroy.restype = None
        L.salt_multi_n.argtypes = [vp]
        L.salt_multi_handle.restype = vp; L.salt_multi_handle.argtypes = [vp, i32]
        L.salt_multi_verify_batch_packed.argtypes = [vp, C.POINTER(api.PackedChunkT), u32, i32, i32, vp, vp, vp, vp, i32]
    L.salt_pair_plan.argtypes = [C.POINTER(ReadResultT), u32, C.POINTER(ReadResultT), u32, u32, u32, u32, C.POINTER(PairPlanT)]
    L.salt_pair_apply.argtypes = [C.POINTER(PairPlanT), C.POINTER(ReadResultT), u32, C.POINTER(ReadResultT), u32, vp, vp, i32, i32, i32, vp]
    return L


def load(path=HOST_LIB_PATH):
    if not os.path.exists(path):
        raise api.SaltError(-105, "libsalt_host.so is not built (python -m salt_b200.build)")
    return declare(C.CDLL(path))


class Chunk:
    """One pinned chunk queue (salt_chunk_t)."""

    def __init__(self, hostlib, max_reads, max_bases, max_cands):
        self.H = hostlib
        self.c = hostlib.salt_chunk_new(int(max_reads), int(max_bases), int(max_cands))
        if not self.c:
            raise api.SaltError(-103, "salt_chunk_new failed")

    def close(self):
        if self.c:
            self.H.salt_chunk_free(self.c); self.c = None

    def reset(self):
        self.H.salt_chunk_reset(self.c)

    def add_read(self, seq, loci0, loci1):
        seq = np.ascontiguousarray(seq, np.uint8)
        l0 = np.ascontiguousarray(loci0, np.uint32); l1 = np.ascontiguousarray(loci1, np.uint32)
        rc = self.H.salt_chunk_add_read(self.c, seq.ctypes.data, len(seq), l0.ctypes.data if len(l0) else None, len(l0),
                                        l1.ctypes.data if len(l1) else None, len(l1))
        if rc < 0:
            raise api.SaltError(rc, "chunk queue full")
        return rc

    def add_reads(self, codes, roffs, offs0, loci0, offs1, loci1):
        codes = np.ascontiguousarray(codes, np.uint8).reshape(-1); roffs = np.ascontiguousarray(roffs, np.uint32)
        offs0 = np.ascontiguousarray(offs0, np.uint32); offs1 = np.ascontiguousarray(offs1, np.uint32)
        loci0 = np.ascontiguousarray(loci0, np.uint32); loci1 = np.ascontiguousarray(loci1, np.uint32)
        rc = self.H.salt_chunk_add_reads(self.c, codes.ctypes.data, roffs.ctypes.data, len(roffs) - 1, offs0.ctypes.data,
                                         loci0.ctypes.data if len(loci0) else None, offs1.ctypes.data,
                                         loci1.ctypes.data if len(loci1) else None)
        if rc < 0:
            raise api.SaltError(rc, "chunk queue full")
        return rc

    def pair(self, eng, slot, n_pairs, min_tlen, max_tlen, l_pac, max_hits=5, gapO=3, gapE=1, filters=0, filterd=20, with_tail=True,
             md_stride=128, bufs=None):
        """salt_chunk_pair: the paired-end stage of the chunk.  Returns (finals, tail_out, tail_md, stats).
        bufs = (finals, tail_out, tail_md) from an earlier call re-uses the caller's output buffers."""
        if bufs is not None:
            out, tail_out, tail_md = bufs
        else:
            out = (PairFinalT * max(n_pairs, 1))()
            tail_out = np.zeros(2 * n_pairs, api.MDNM_OUT_DT); tail_md = np.zeros((2 * n_pairs, md_stride), np.uint8)
        st = PeStatsT()
        m16 = api.salt_score_mat2(); m5 = api.salt_score_mat()
        eng._ck(self.H.salt_chunk_pair(eng.h, int(slot), self.c, int(min_tlen), int(max_tlen), int(l_pac), int(max_hits),
                                       m16.ctypes.data, m5.ctypes.data, int(gapO), int(gapE), int(filters), int(filterd), int(with_tail),
                                       out, tail_out.ctypes.data, tail_md.ctypes.data, int(md_stride), C.byref(st)))
        return out, tail_out, tail_md, st

    def seed_verify(self, eng, opt, nogap_T0=3, lv_T0=-1):
        eng._ck(self.H.salt_chunk_seed_verify(eng.h, self.c, C.byref(opt), int(nogap_T0), int(lv_T0)))

    def submit(self, eng, slot, nogap_T0=3, lv_T0=-1):
        eng._ck(self.H.salt_chunk_submit(eng.h, int(slot), self.c, int(nogap_T0), int(lv_T0)))

    def wait(self, eng, slot):
        eng._ck(self.H.salt_chunk_wait(eng.h, int(slot), self.c))

    def result(self, i, max_hits=5):
        r = ReadResultT()
        rc = self.H.salt_chunk_result(self.c, int(i), int(max_hits), C.byref(r))
        if rc != 0:
            raise api.SaltError(rc, "salt_chunk_result")
        alts = [[(h.pos, h.n_diff, h.is_gap, h.strand) for h in r.alt[s][:r.n_alt[s]]] for s in (0, 1)]
        return (r.pos, r.strand, r.n_diff, r.is_gap, r.b0, r.b1, r.mapq), alts, r.cigar.decode()

    def tail(self, eng, slot):
        eng._ck(self.H.salt_chunk_tail(eng.h, int(slot), self.c))

    def md(self, i):
        """(MD value, NM, XV offsets) of read i after tail(); MD is '' for an unmapped read"""
        nm = C.c_int(); nx = C.c_int(); xv = C.POINTER(C.c_uint16)()
        md = self.H.salt_chunk_md(self.c, int(i), C.byref(nm), C.byref(xv), C.byref(nx))
        if md is None:
            raise api.SaltError(-101, "salt_chunk_md")
        return md.decode(), nm.value, [int(xv[k]) for k in range(nx.value)]

    def hits(self, i, strand, cap=4096):
        buf = (HitT * cap)()
        n = self.H.salt_chunk_hits(self.c, int(i), int(strand), buf, cap)
        if n < 0:
            raise api.SaltError(n, "salt_chunk_hits")
        return [(h.pos, h.n_diff, h.is_gap, h.strand) for h in buf[:n]]
