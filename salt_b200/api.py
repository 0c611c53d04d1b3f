"""ctypes binding of libsalt_b200.so (include/salt_b200.h) with numpy in/out.

This is the call a Python user makes; it only forwards to the C ABI.  `load()` opens the
CUDA library built in-tree and refuses to work without it or without a GPU -- there is no
CPU implementation behind this module.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsalt_b200.so")

SALT_OK = 0
ERRORS = {-101: "SALT_ERR_ARG", -102: "SALT_ERR_CUDA", -103: "SALT_ERR_NOMEM",
          -104: "SALT_ERR_UNSUPPORTED", -105: "SALT_ERR_NODEVICE"}

PAIR_DT = np.dtype([("rs", np.uint32), ("pos", np.uint32)])
WIN_DT = np.dtype([("rs", np.uint32), ("start", np.uint32), ("end", np.uint32)])
MDNM_IN_DT = np.dtype([("rs", np.uint32), ("pos", np.uint32), ("seq_start", np.uint32)])
MDNM_OUT_DT = np.dtype([("nm", np.int32), ("md_len", np.int16), ("n_xv", np.uint16)])
SSW_DT = np.dtype([("score1", np.uint16), ("score2", np.uint16), ("ref_begin1", np.int32), ("ref_end1", np.int32),
                   ("read_begin1", np.int32), ("read_end1", np.int32), ("ref_end2", np.int32), ("cigarLen", np.int32)])
VERIFY_DT = np.dtype([("pos", np.uint32), ("strand", np.uint8), ("n_diff", np.uint8), ("is_gap", np.uint8),
                      ("lv_ran", np.uint8), ("n_hits", np.int32, (2,))])
assert PAIR_DT.itemsize == 8 and WIN_DT.itemsize == 12 and SSW_DT.itemsize == 28 and VERIFY_DT.itemsize == 16


class SaltError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (ERRORS.get(code, "error"), code, msg))
        self.code = code


class ReadsT(C.Structure):
    _fields_ = [("codes", C.c_void_p), ("offs", C.c_void_p), ("n_reads", C.c_uint32)]


class CandsT(C.Structure):
    _fields_ = [("offs", C.c_void_p * 2), ("loci", C.c_void_p * 2)]


class PackedChunkT(C.Structure):
    """salt_packed_chunk_t (include/salt_b200.h): the compact host->device transport of a chunk."""
    _fields_ = [("n_reads", C.c_uint32), ("base_bits", C.c_int), ("bases", C.c_void_p), ("base_start", C.c_uint32),
                ("lens", C.c_void_p), ("l_seq", C.c_uint32), ("n_pos", C.c_void_p), ("n_n", C.c_size_t),
                ("count_bits", C.c_int), ("n_cand", C.c_void_p * 2), ("loci", C.c_void_p * 2)]


def pack_bases(codes, bits, base_start=0):
    """codes 0..4 (flat uint8 stream) -> (packed bytes, positions of N) in the transport layout: base p of the stream
    in bits `bits`*(p mod 8/bits) of byte p / (8/bits), lowest bits first.  The stream's first base sits at
    position base_start.  What a FASTQ parser would emit directly; numpy here because the tests are Python."""
    codes = np.ascontiguousarray(codes, np.uint8).reshape(-1)
    per = 8 // bits
    n_pos = np.nonzero(codes > 3)[0].astype(np.uint32) + np.uint32(base_start)
    c = np.minimum(codes, 4) if bits == 4 else np.where(codes > 3, 0, codes).astype(np.uint8)
    c = np.concatenate([np.zeros(base_start, np.uint8), c, np.zeros((-(len(c) + base_start)) % per + per, np.uint8)])
    c = c.reshape(-1, per).astype(np.uint32)
    out = np.zeros(len(c), np.uint32)
    for e in range(per):
        out |= c[:, e] << (bits * e)
    return out.astype(np.uint8), (n_pos if bits == 2 else np.zeros(0, np.uint32))


def _declare(L):
    vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
    L.salt_b200_last_error.restype = C.c_char_p
    L.salt_b200_init.restype = vp
    L.salt_b200_init.argtypes = [vp, C.c_uint32, vp, C.c_int64, i32]
    L.salt_b200_init_from_bases.restype = vp
    L.salt_b200_init_from_bases.argtypes = [vp, C.c_uint32, vp, vp, sz, i32]
    L.salt_b200_get_mixref.argtypes = [vp, vp, sz]
    L.salt_b200_destroy.argtypes = [vp]
    L.salt_b200_destroy.restype = None
    L.salt_b200_attach.argtypes = [vp]
    L.salt_b200_attach.restype = vp
    L.salt_b200_set_stream.argtypes = [vp, vp]
    L.salt_b200_sync.argtypes = [vp]
    L.salt_b200_host_alloc.restype = vp
    L.salt_b200_host_alloc.argtypes = [sz]
    L.salt_b200_host_free.argtypes = [vp]
    L.salt_b200_host_free.restype = None
    L.salt_b200_set_reads.argtypes = [vp, C.POINTER(ReadsT)]
    L.salt_b200_mismatch.argtypes = [vp, vp, sz, i32, vp]
    L.salt_b200_lv.argtypes = [vp, vp, sz, i32, vp]
    L.salt_b200_lv_cigar.argtypes = [vp, vp, vp, sz, vp, i32, vp]
    L.salt_b200_md_nm.argtypes = [vp, i32, vp, sz, vp, i32, vp, i32, vp, i32, vp]
    L.salt_b200_ssw.argtypes = [vp, vp, sz, i32, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, i32]
    L.salt_b200_verify.argtypes = [vp, C.POINTER(CandsT), i32, i32, vp, vp, vp, vp, i32]
    L.salt_b200_verify_submit.argtypes = [vp, i32, C.POINTER(ReadsT), C.POINTER(CandsT), i32, i32, vp, vp, vp, vp, i32]
    L.salt_b200_verify_wait.argtypes = [vp, i32]
    L.salt_b200_verify_batch.argtypes = [vp, C.POINTER(ReadsT), C.POINTER(CandsT), C.c_uint32, i32, i32, vp, vp, vp, vp, i32]
    if hasattr(L, "salt_b200_set_reads_packed"):
        L.salt_b200_set_reads_packed.argtypes = [vp, C.POINTER(PackedChunkT)]
        L.salt_b200_verify_submit_packed.argtypes = [vp, i32, C.POINTER(PackedChunkT), i32, i32, vp, vp, vp, vp, i32]
        L.salt_b200_verify_batch_packed.argtypes = [vp, C.POINTER(PackedChunkT), C.c_uint32, i32, i32, vp, vp, vp, vp, i32]
    if hasattr(L, "salt_b200_set_index"):
        from .index_io import FmIndexT, SeedOptT
        szp = C.POINTER(C.c_size_t)
        L.salt_b200_set_index.argtypes = [vp, C.POINTER(FmIndexT)]
        L.salt_b200_seed_locate.argtypes = [vp, i32, C.POINTER(SeedOptT), vp, vp, vp, sz, vp, sz, szp, szp]
        L.salt_b200_verify_seeded.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, i32]
        L.salt_b200_seed_status.argtypes = [vp, i32, vp, vp]
        L.salt_b200_align_batch_packed.argtypes = [vp, C.POINTER(PackedChunkT), C.POINTER(SeedOptT), C.c_uint32, i32, i32, vp, vp, i32]
    if hasattr(L, "salt_b200_tail_primaries"):
        L.salt_b200_tail_primaries.argtypes = [vp, i32, vp, vp, vp, sz, C.POINTER(C.c_size_t), vp, i32]
        L.salt_b200_use_slot.argtypes = [vp, i32]
        L.salt_b200_tail_submit.argtypes = [vp, i32, vp, vp, vp, sz, vp, i32]
        L.salt_b200_tail_wait.argtypes = [vp, i32, C.POINTER(C.c_size_t)]
    L.salt_b200_set_max_window.argtypes = [vp, i32]
    L.salt_b200_set_lv_mapping.argtypes = [vp, i32]
    L.salt_b200_set_lv_filter.argtypes = [vp, i32]
    L.salt_b200_mismatch_dev.argtypes = [vp, vp, sz, i32, vp]
    L.salt_b200_lv_dev.argtypes = [vp, vp, sz, i32, vp]
    L.salt_b200_verify_dev.argtypes = [vp, vp, vp, sz, vp, vp, sz, i32, i32, vp, vp, vp, vp, i32, vp, vp]
    L.salt_b200_profile.argtypes = [vp, i32]
    L.salt_b200_profile_read.argtypes = [vp, vp]
    L.salt_b200_ssw_dev.argtypes = [vp, vp, sz, i32, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, i32]
    L.salt_b200_launch_count.restype = C.c_uint64
    L.salt_b200_launch_count.argtypes = [vp, i32]
    return L


_lib = None


def load():
    """Open the CUDA library.  Raises if it was not built or no device is visible."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SaltError(-105, "libsalt_b200.so is not built (python -m salt_b200.build); there is no CPU fallback")
        _lib = _declare(C.CDLL(LIB_PATH))
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data


def salt_score_mat2():
    """salt's 16x16 SNP-aware SSW matrix as alnpe.c:58-73 lays it out (rule, not a copy):
    rows 1,2,4,8 score +1 where the column shares the row's bit, everything else -3."""
    m = np.full((16, 16), -3, np.int8)
    for r in (1, 2, 4, 8):
        for c in range(16):
            if c & r:
                m[r, c] = 1
    return m.ravel().copy()


def salt_score_mat():
    """salt's 5x5 matrix for the 2-bit reference (alnpe.c:52-56): +1 / -3, N scores -1."""
    m = np.full((5, 5), -3, np.int8)
    m[np.arange(4), np.arange(4)] = 1
    m[4, :] = -1
    m[:, 4] = -1
    return m.ravel().copy()


class Engine:
    """One handle = one GPU with the reference resident in HBM."""

    def __init__(self, mixref, l, pac=None, l_pac=0, device=0, lib=None):
        self.L = lib if lib is not None else load()
        self.mixref = np.ascontiguousarray(mixref, np.uint32)
        pac = None if pac is None else np.ascontiguousarray(pac, np.uint8)
        self.h = self.L.salt_b200_init(_ptr(self.mixref), int(l), _ptr(pac), int(l_pac), int(device))
        if not self.h:
            code = -105 if b"no CUDA device" in self.L.salt_b200_last_error() else -102
            raise SaltError(code, self.L.salt_b200_last_error().decode())
        self.l = int(l)
        self.n_reads = 0

    @classmethod
    def from_bases(cls, bases, snp_pos, snp_mask, device=0, lib=None):
        self = cls.__new__(cls)
        self.L = lib if lib is not None else load()
        b = np.frombuffer(bases.encode() if isinstance(bases, str) else bases, np.uint8)
        snp_pos = np.ascontiguousarray(snp_pos, np.uint32); snp_mask = np.ascontiguousarray(snp_mask, np.uint8)
        self.h = self.L.salt_b200_init_from_bases(_ptr(b), len(b), _ptr(snp_pos), _ptr(snp_mask), len(snp_pos), int(device))
        if not self.h:
            raise SaltError(-102, self.L.salt_b200_last_error().decode())
        self.l = len(b); self.n_reads = 0
        return self

    def attach(self):
        """A second handle for another host thread: own streams, slots and scratch, this handle's resident reference and
        indexes (salt_b200_attach).  This handle must outlive it."""
        other = Engine.__new__(Engine)
        other.L = self.L
        other.h = self.L.salt_b200_attach(self.h)
        if not other.h:
            raise SaltError(-102, self.L.salt_b200_last_error().decode())
        other.l = self.l; other.n_reads = 0; other._parent = self
        return other

    def _ck(self, rc):
        if rc != SALT_OK:
            raise SaltError(rc, self.L.salt_b200_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.salt_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_mixref(self):
        out = np.zeros((self.l + 7) // 8, np.uint32)
        self._ck(self.L.salt_b200_get_mixref(self.h, _ptr(out), len(out)))
        return out

    def set_stream(self, stream_ptr):
        self._ck(self.L.salt_b200_set_stream(self.h, stream_ptr))

    def sync(self):
        self._ck(self.L.salt_b200_sync(self.h))

    STAGES = ("_a", "nogap_fused", "_b", "lv", "scan_gap", "lv_cigar",
              "ssw_prep_fwd", "ssw_dp_fwd", "ssw_prep_rev", "ssw_dp_rev", "ssw_banded", "ssw_banded_ovf")

    def profile(self, enable=True):
        self._ck(self.L.salt_b200_profile(self.h, int(enable)))

    def profile_read(self):
        ms = np.zeros(12, np.float32)
        self._ck(self.L.salt_b200_profile_read(self.h, _ptr(ms)))
        return {k: float(v) for k, v in zip(self.STAGES, ms) if v >= 0}

    def set_lv_mapping(self, mapping):
        self._ck(self.L.salt_b200_set_lv_mapping(self.h, int(mapping)))

    def set_lv_filter(self, enable):
        self._ck(self.L.salt_b200_set_lv_filter(self.h, int(enable)))

    def launch_count(self, reset=False):
        return int(self.L.salt_b200_launch_count(self.h, int(reset)))

    # ---- reads -------------------------------------------------------------------
    def set_reads(self, codes, offs=None):
        """codes: [n, L] uint8 matrix, or flat uint8 with offs (n+1 uint32)."""
        codes = np.ascontiguousarray(codes, np.uint8)
        if offs is None:
            n, L = codes.shape
            offs = (np.arange(n + 1, dtype=np.uint64) * L).astype(np.uint32)
            codes = codes.reshape(-1)
        offs = np.ascontiguousarray(offs, np.uint32)
        r = ReadsT(_ptr(codes), _ptr(offs), len(offs) - 1)
        self._ck(self.L.salt_b200_set_reads(self.h, C.byref(r)))
        self.n_reads = len(offs) - 1

    # ---- per-pair kernels ----------------------------------------------------------
    @staticmethod
    def make_pairs(read_ids, strands, pos):
        p = np.zeros(len(pos), PAIR_DT)
        p["rs"] = (np.asarray(read_ids, np.uint32) << 1) | np.asarray(strands, np.uint32)
        p["pos"] = pos
        return p

    def mismatch(self, pairs, max_err=3):
        pairs = np.ascontiguousarray(pairs, PAIR_DT)
        out = np.empty(len(pairs), np.int8)
        self._ck(self.L.salt_b200_mismatch(self.h, _ptr(pairs), len(pairs), int(max_err), _ptr(out)))
        return out

    def lv(self, pairs, k=-1):
        pairs = np.ascontiguousarray(pairs, PAIR_DT)
        out = np.empty(len(pairs), np.int8)
        self._ck(self.L.salt_b200_lv(self.h, _ptr(pairs), len(pairs), int(k), _ptr(out)))
        return out

    def md_nm(self, rs, pos, seq_start, cigars, md_stride=128, xv_stride=64, cigar_stride=None, slot=0):
        """sam_add_md_nm for alignments of the current read set.  cigars: list of str.  Returns
        (out records, md strings as a [n, md_stride] uint8 array, xv [n, xv_stride] uint16)."""
        n = len(pos)
        items = np.zeros(n, MDNM_IN_DT)
        items["rs"] = rs; items["pos"] = pos; items["seq_start"] = seq_start
        if cigar_stride is None:
            cigar_stride = max([len(c) for c in cigars] + [1]) + 1
        cg = np.zeros((n, cigar_stride), np.uint8)
        for i, c in enumerate(cigars):
            b = np.frombuffer(c.encode(), np.uint8)
            cg[i, :len(b)] = b
        md = np.zeros((n, md_stride), np.uint8)
        xv = np.zeros((n, max(xv_stride, 1)), np.uint16)
        out = np.zeros(n, MDNM_OUT_DT)
        self._ck(self.L.salt_b200_md_nm(self.h, int(slot), _ptr(items), n, _ptr(cg), cigar_stride, _ptr(md), md_stride,
                                        _ptr(xv) if xv_stride > 0 else None, xv_stride, _ptr(out)))
        return out, md, xv

    def tail_primaries(self, slot=0, xv_stride=64, md_cap=None):
        """MD / NM / XV of the primaries of the chunk just verified in `slot` (salt_b200_tail_primaries).
        Returns (out records, md_offs, packed md bytes, xv rows)."""
        n = self.n_reads
        out = np.zeros(n, MDNM_OUT_DT); offs = np.zeros(n + 1, np.uint32)
        cap = md_cap if md_cap is not None else n * 300 + 16
        packed = np.zeros(cap, np.uint8); xv = np.zeros((n, max(xv_stride, 1)), np.uint16)
        nb = C.c_size_t(0)
        self._ck(self.L.salt_b200_tail_primaries(self.h, int(slot), _ptr(out), _ptr(offs), _ptr(packed), cap, C.byref(nb),
                                                 _ptr(xv) if xv_stride > 0 else None, int(xv_stride)))
        return out, offs, packed[:nb.value], xv

    @staticmethod
    def md_nm_text(out, md, xv, i):
        """the text sam_add_md_nm appends for item i, from the engine's record"""
        s = "\tMD:Z:" + cstr(md[i]) + "\tNM:i:%u" % out["nm"][i]
        if out["n_xv"][i]:
            s += "\tXV:i:" + ",".join(str(int(v)) for v in xv[i, :out["n_xv"][i]])
        return s

    def lv_cigar(self, pairs, k_each, stride=128, fill=0):
        pairs = np.ascontiguousarray(pairs, PAIR_DT)
        k_each = np.ascontiguousarray(k_each, np.uint8)
        out = np.empty(len(pairs), np.int8)
        buf = np.full((len(pairs), stride), fill, np.uint8)
        self._ck(self.L.salt_b200_lv_cigar(self.h, _ptr(pairs), _ptr(k_each), len(pairs), _ptr(buf), int(stride), _ptr(out)))
        return out, buf

    def ssw(self, wins, mat, n_sym=16, use_pac=False, gapO=3, gapE=1, flag=2, filters=0, filterd=20,
            mask_len=-1, cigar_stride=64):
        wins = np.ascontiguousarray(wins, WIN_DT)
        mat = np.ascontiguousarray(mat, np.int8)
        out = np.zeros(len(wins), SSW_DT)
        cig = np.zeros((len(wins), cigar_stride), np.uint32)
        self._ck(self.L.salt_b200_ssw(self.h, _ptr(wins), len(wins), int(use_pac), _ptr(mat), int(n_sym), int(gapO),
                                      int(gapE), int(flag), int(filters), int(filterd), int(mask_len),
                                      _ptr(out), _ptr(cig), int(cigar_stride)))
        return out, cig

    # ---- verification stage ----------------------------------------------------------
    def verify(self, offs0, loci0, offs1, loci1, nogap_T0=3, lv_T0=-1, cigar_stride=128, want_cigars=True):
        offs0 = np.ascontiguousarray(offs0, np.uint32); offs1 = np.ascontiguousarray(offs1, np.uint32)
        loci0 = np.ascontiguousarray(loci0, np.uint32); loci1 = np.ascontiguousarray(loci1, np.uint32)
        c = CandsT()
        c.offs[0], c.offs[1] = _ptr(offs0), _ptr(offs1)
        c.loci[0], c.loci[1] = _ptr(loci0) if len(loci0) else None, _ptr(loci1) if len(loci1) else None
        rec = np.zeros(self.n_reads, VERIFY_DT)
        acc0 = np.empty(len(loci0), np.int8); acc1 = np.empty(len(loci1), np.int8)
        cig = np.zeros((self.n_reads, cigar_stride), np.uint8) if want_cigars else None
        self._ck(self.L.salt_b200_verify(self.h, C.byref(c), int(nogap_T0), int(lv_T0), _ptr(rec), _ptr(acc0), _ptr(acc1),
                                         _ptr(cig), int(cigar_stride)))
        return rec, acc0, acc1, cig


    def verify_batch(self, codes, roffs, offs0, loci0, offs1, loci1, chunk_reads=100000, nogap_T0=3, lv_T0=-1,
                     cigar_stride=128, want_cigars=True):
        """Whole batch through the asynchronous chunk pipeline (salt_b200_verify_batch)."""
        codes = np.ascontiguousarray(codes, np.uint8).reshape(-1)
        roffs = np.ascontiguousarray(roffs, np.uint32)
        offs0 = np.ascontiguousarray(offs0, np.uint32); offs1 = np.ascontiguousarray(offs1, np.uint32)
        loci0 = np.ascontiguousarray(loci0, np.uint32); loci1 = np.ascontiguousarray(loci1, np.uint32)
        n = len(roffs) - 1
        r = ReadsT(_ptr(codes), _ptr(roffs), n)
        c = CandsT()
        c.offs[0], c.offs[1] = _ptr(offs0), _ptr(offs1)
        c.loci[0], c.loci[1] = _ptr(loci0) if len(loci0) else None, _ptr(loci1) if len(loci1) else None
        rec = np.zeros(n, VERIFY_DT)
        acc0 = np.empty(len(loci0), np.int8); acc1 = np.empty(len(loci1), np.int8)
        cig = np.zeros((n, cigar_stride), np.uint8) if want_cigars else None
        self._ck(self.L.salt_b200_verify_batch(self.h, C.byref(r), C.byref(c), int(chunk_reads), int(nogap_T0), int(lv_T0),
                                               _ptr(rec), _ptr(acc0), _ptr(acc1), _ptr(cig), int(cigar_stride)))
        return rec, acc0, acc1, cig


    def packed_chunk(self, codes, roffs, offs0, loci0, offs1, loci1, bits=2, count_bits=16, base_start=0, uniform=None):
        """Build a salt_packed_chunk_t (and the arrays it points into, which the caller must keep alive) from the
        plain CSR description of a batch."""
        codes = np.ascontiguousarray(codes, np.uint8).reshape(-1)
        roffs = np.ascontiguousarray(roffs, np.int64)
        keep = {}
        keep["bases"], keep["n_pos"] = pack_bases(codes, bits, base_start)
        lens = np.diff(roffs)
        if uniform is None:
            uniform = len(lens) > 0 and bool((lens == lens[0]).all())
        keep["lens"] = None if uniform else lens.astype(np.uint16)
        cdt = np.uint16 if count_bits == 16 else np.uint32
        keep["cnt0"] = np.diff(np.asarray(offs0, np.int64)).astype(cdt); keep["cnt1"] = np.diff(np.asarray(offs1, np.int64)).astype(cdt)
        keep["loci0"] = np.ascontiguousarray(loci0, np.uint32); keep["loci1"] = np.ascontiguousarray(loci1, np.uint32)
        pc = PackedChunkT()
        pc.n_reads = len(lens); pc.base_bits = bits; pc.bases = _ptr(keep["bases"]); pc.base_start = base_start
        pc.lens = _ptr(keep["lens"]); pc.l_seq = int(lens[0]) if uniform and len(lens) else 0
        pc.n_pos = _ptr(keep["n_pos"]) if len(keep["n_pos"]) else None; pc.n_n = len(keep["n_pos"])
        pc.count_bits = count_bits
        pc.n_cand[0], pc.n_cand[1] = _ptr(keep["cnt0"]), _ptr(keep["cnt1"])
        pc.loci[0] = _ptr(keep["loci0"]) if len(keep["loci0"]) else None
        pc.loci[1] = _ptr(keep["loci1"]) if len(keep["loci1"]) else None
        return pc, keep

    # ---- seeding + locate (row f1) --------------------------------------------------------
    def set_index(self, fm):
        """fm: index_io.FmIndex (kept alive by the caller until this returns; the engine keeps its own copy)."""
        st = fm.struct()
        self._ck(self.L.salt_b200_set_index(self.h, C.byref(st)))

    @staticmethod
    def seed_opt(l_seed, l_overlap=0, max_seed=50, max_locate=1000, seed_only_ref=0, locate_mode=0, list_cap=0):
        from .index_io import SeedOptT
        return SeedOptT(int(l_seed), int(l_overlap) if l_overlap > 0 else int(l_seed), int(max_seed), int(max_locate), int(seed_only_ref),
                        int(locate_mode), int(list_cap))

    def seed_status(self, slot=0):
        st0 = np.zeros(self.n_reads, np.uint8); st1 = np.zeros(self.n_reads, np.uint8)
        self._ck(self.L.salt_b200_seed_status(self.h, int(slot), _ptr(st0), _ptr(st1)))
        return st0, st1

    def seed_locate(self, opt, slot=0, download=True):
        """alnse_seed_overlap + alnse_locate_alt for the reads in `slot`: (offs0, loci0, offs1, loci1) or just the totals."""
        n0 = C.c_size_t(0); n1 = C.c_size_t(0)
        self._ck(self.L.salt_b200_seed_locate(self.h, int(slot), C.byref(opt), None, None, None, 0, None, 0, C.byref(n0), C.byref(n1)))
        if not download:
            return n0.value, n1.value
        offs0 = np.zeros(self.n_reads + 1, np.uint32); offs1 = np.zeros(self.n_reads + 1, np.uint32)
        loci0 = np.zeros(n0.value, np.uint32); loci1 = np.zeros(n1.value, np.uint32)
        self._ck(self.L.salt_b200_seed_locate(self.h, int(slot), C.byref(opt), _ptr(offs0), _ptr(offs1), _ptr(loci0), len(loci0),
                                              _ptr(loci1), len(loci1), C.byref(n0), C.byref(n1)))
        return offs0, loci0, offs1, loci1

    def verify_seeded(self, n0, n1, nogap_T0=3, lv_T0=-1, cigar_stride=128, slot=0):
        rec = np.zeros(self.n_reads, VERIFY_DT)
        acc0 = np.empty(n0, np.int8); acc1 = np.empty(n1, np.int8)
        cig = np.zeros((self.n_reads, cigar_stride), np.uint8)
        self._ck(self.L.salt_b200_verify_seeded(self.h, int(slot), int(nogap_T0), int(lv_T0), _ptr(rec), _ptr(acc0), _ptr(acc1),
                                                _ptr(cig), int(cigar_stride)))
        return rec, acc0, acc1, cig

    def align_batch_packed(self, pc, opt, chunk_reads=100000, nogap_T0=3, lv_T0=-1, cigar_stride=128):
        n = int(pc.n_reads)
        rec = np.zeros(n, VERIFY_DT)
        cig = np.zeros((n, cigar_stride), np.uint8)
        self._ck(self.L.salt_b200_align_batch_packed(self.h, C.byref(pc), C.byref(opt), int(chunk_reads), int(nogap_T0), int(lv_T0),
                                                     _ptr(rec), _ptr(cig), int(cigar_stride)))
        return rec, cig

    def set_reads_packed(self, pc):
        self._ck(self.L.salt_b200_set_reads_packed(self.h, C.byref(pc)))
        self.n_reads = int(pc.n_reads)

    def verify_batch_packed(self, pc, n0, n1, chunk_reads=100000, nogap_T0=3, lv_T0=-1, cigar_stride=128, want_cigars=True):
        """salt_b200_verify_batch_packed: the chunk pipeline fed in the compact transport format."""
        n = int(pc.n_reads)
        rec = np.zeros(n, VERIFY_DT)
        acc0 = np.empty(n0, np.int8); acc1 = np.empty(n1, np.int8)
        cig = np.zeros((n, cigar_stride), np.uint8) if want_cigars else None
        self._ck(self.L.salt_b200_verify_batch_packed(self.h, C.byref(pc), int(chunk_reads), int(nogap_T0), int(lv_T0),
                                                      _ptr(rec), _ptr(acc0), _ptr(acc1), _ptr(cig), int(cigar_stride)))
        return rec, acc0, acc1, cig


def cstr(row):
    """NUL-terminated bytes in a uint8 row -> str."""
    b = bytes(row)
    return b.split(b"\0")[0].decode()
