"""The reference's on-disk index (written by salt-idx, read by alnse_index_reload, indexio.c:23-50) as numpy arrays,
and the ctypes mirror of salt_fm_index_t.

    PREFIX.C.bwt            primary, L2[1..4], BWT words in BWA's 128-base blocks (bwtio.c:55-74)
    PREFIX.C.sa             primary, 4 skipped words, sa_intv, seq_len, sa[1..]; sa[0] = (uint32)-1 (bwtio.c:30-53)
    PREFIX.C.lkt            int maxLookupLen, then 4^len + 1 cumulative counts (lookup.c:47-65)
    PREFIX.R.backward.bwt   textLength, inverseSa0, cumulativeFreq[1..5], bwtSizeInWord, 4-bit BWT (rbwt.c:250-270)
    PREFIX.R.backward.occ   occSizeInWord + explicit counts, occMajorSizeInWord + major counts (rbwt.c:271-288)
    PREFIX.R.backward.sa    size, positions of the '#'-suffixes (rbwt.c:558-574)
    PREFIX.R.seedLen        int l_seed (aln.c:216-224)
    PREFIX.ref              uint32 l, then the 4-bit SNP-aware reference (metaref.c:61-93)
"""
import ctypes as C

import numpy as np


class FmIndexT(C.Structure):
    _fields_ = [("c_bwt", C.c_void_p), ("c_bwt_words", C.c_size_t),
                ("c_primary", C.c_uint32), ("c_seq_len", C.c_uint32), ("c_L2", C.c_uint32 * 5),
                ("c_sa", C.c_void_p), ("c_n_sa", C.c_uint32), ("c_sa_intv", C.c_uint32),
                ("lkt", C.c_void_p), ("lkt_len", C.c_uint32),
                ("r_bwt", C.c_void_p), ("r_bwt_words", C.c_size_t),
                ("r_occ", C.c_void_p), ("r_occ_words", C.c_size_t),
                ("r_occ_major", C.c_void_p), ("r_occ_major_words", C.c_size_t),
                ("r_sa_sharp", C.c_void_p), ("r_n_sa_sharp", C.c_size_t),
                ("r_cum", C.c_uint32 * 6), ("r_inv_sa0", C.c_uint32), ("r_text_len", C.c_uint32)]


class SeedOptT(C.Structure):
    _fields_ = [("l_seed", C.c_int), ("l_overlap", C.c_int), ("max_seed", C.c_int), ("max_locate", C.c_int),
                ("seed_only_ref", C.c_int), ("locate_mode", C.c_int), ("list_cap", C.c_int)]


class FmIndex:
    """Index files of PREFIX loaded into numpy arrays; .struct() is the salt_fm_index_t pointing into them."""

    def __init__(self, prefix):
        raw = np.fromfile(prefix + ".C.bwt", np.uint32)
        self.c_primary = int(raw[0]); self.c_L2 = [0] + [int(x) for x in raw[1:5]]
        self.c_bwt = np.ascontiguousarray(raw[5:]); self.c_seq_len = self.c_L2[4]
        raw = np.fromfile(prefix + ".C.sa", np.uint32)
        assert int(raw[0]) == self.c_primary and int(raw[6]) == self.c_seq_len, "SA-BWT inconsistency"
        self.c_sa_intv = int(raw[5])
        n_sa = (self.c_seq_len + self.c_sa_intv) // self.c_sa_intv
        self.c_sa = np.empty(n_sa, np.uint32); self.c_sa[0] = 0xFFFFFFFF; self.c_sa[1:] = raw[7:7 + n_sa - 1]
        raw = np.fromfile(prefix + ".C.lkt", np.uint32)
        self.lkt_len = int(raw[0]); self.lkt = np.ascontiguousarray(raw[1:1 + 4 ** self.lkt_len + 1])
        assert len(self.lkt) == 4 ** self.lkt_len + 1
        raw = np.fromfile(prefix + ".R.backward.bwt", np.uint32)
        self.r_text_len, self.r_inv_sa0 = int(raw[0]), int(raw[1])
        self.r_cum = [0] + [int(x) for x in raw[2:7]]
        nw = int(raw[7]); self.r_bwt = np.ascontiguousarray(raw[8:8 + nw])
        raw = np.fromfile(prefix + ".R.backward.occ", np.uint32)
        no = int(raw[0]); self.r_occ = np.ascontiguousarray(raw[1:1 + no])
        nm = int(raw[1 + no]); self.r_occ_major = np.ascontiguousarray(raw[2 + no:2 + no + nm])
        raw = np.fromfile(prefix + ".R.backward.sa", np.uint32)
        self.r_sa_sharp = np.ascontiguousarray(raw[1:1 + int(raw[0])])
        self.l_seed = int(np.fromfile(prefix + ".R.seedLen", np.int32)[0])
        raw = np.fromfile(prefix + ".ref", np.uint32)
        self.l = int(raw[0]); self.mixref = np.ascontiguousarray(raw[1:1 + (self.l + 7) // 8])

    def struct(self):
        s = FmIndexT()
        s.c_bwt = self.c_bwt.ctypes.data; s.c_bwt_words = len(self.c_bwt)
        s.c_primary = self.c_primary; s.c_seq_len = self.c_seq_len
        for i in range(5):
            s.c_L2[i] = self.c_L2[i]
        s.c_sa = self.c_sa.ctypes.data; s.c_n_sa = len(self.c_sa); s.c_sa_intv = self.c_sa_intv
        s.lkt = self.lkt.ctypes.data; s.lkt_len = self.lkt_len
        s.r_bwt = self.r_bwt.ctypes.data; s.r_bwt_words = len(self.r_bwt)
        s.r_occ = self.r_occ.ctypes.data; s.r_occ_words = len(self.r_occ)
        s.r_occ_major = self.r_occ_major.ctypes.data; s.r_occ_major_words = len(self.r_occ_major)
        s.r_sa_sharp = self.r_sa_sharp.ctypes.data; s.r_n_sa_sharp = len(self.r_sa_sharp)
        for i in range(6):
            s.r_cum[i] = self.r_cum[i]
        s.r_inv_sa0 = self.r_inv_sa0; s.r_text_len = self.r_text_len
        return s
