"""Row f2, the text itself: salt_sam_se / salt_sam_pe (include/salt_host.h) against the reference's own formatters aln_samse
(sam.c:86-180) and alnpe_sam (sam.c:331-455) behind oracle/dropin/sam_harness.c -- several reference records, both strands,
unmapped reads and mates, alternates with and without CIGARs (XA), MD / NM / XV, read groups, soft clips, template lengths
inside and outside the insert bounds.  Host code only: no GPU, no engine."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import parity_cases as pc
from salt_b200 import host_api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libsaltref_sam.so")


class HitT(C.Structure):
    _fields_ = [("pos", C.c_uint32), ("n_diff", C.c_uint8), ("is_gap", C.c_uint8), ("strand", C.c_uint16)]


class SamRefsT(C.Structure):
    _fields_ = [("n_seqs", C.c_int), ("names", C.POINTER(C.c_char_p)), ("offsets", C.POINTER(C.c_int64)), ("l_pac", C.c_int64)]


class SamReadT(C.Structure):
    _fields_ = [("name", C.c_char_p), ("seq", C.c_void_p), ("qual", C.c_char_p), ("l_seq", C.c_uint32), ("pos", C.c_uint32),
                ("strand", C.c_uint8), ("mapq", C.c_uint32), ("cigar", C.c_char_p), ("seq_start", C.c_uint32), ("seq_end", C.c_uint32),
                ("n_alt", C.c_int * 2), ("alt", C.POINTER(HitT) * 2), ("xa_cigars", C.POINTER(C.c_char_p)),
                ("md", C.c_char_p), ("nm", C.c_uint32), ("xv", C.c_void_p), ("n_xv", C.c_int)]


class RefReadT(C.Structure):
    _fields_ = [("name", C.c_char_p), ("seq", C.c_void_p), ("rseq", C.c_void_p), ("qual", C.c_char_p), ("l_seq", C.c_int),
                ("pos", C.c_uint32), ("strand", C.c_int), ("mapq", C.c_uint32), ("cigar", C.c_char_p), ("seq_start", C.c_uint32),
                ("seq_end", C.c_uint32), ("n_alt", C.c_int * 2), ("alt", C.c_void_p * 2)]


def _hostlib():
    try:
        return host_api.load()
    except Exception:
        import sys
        sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
        import build_emul
        return host_api.load(build_emul.build_host())


class Read:
    """one read with everything either formatter wants; keeps the numpy / ctypes objects alive"""

    def __init__(self, oracle, g, name, seq, qual, loci0, loci1, rng, unmapped=False):
        self.name = name.encode(); self.seq = np.ascontiguousarray(seq, np.uint8); self.rseq = np.ascontiguousarray(synth.revcomp(seq), np.uint8)
        self.qual = None if qual is None else qual.encode()
        L = len(seq)
        prim, hits, alts = oracle.verify_read(g.mixref, g.l, self.seq, self.rseq, loci0, loci1, 3, L // 10)
        self.pos, self.strand = (0xFFFFFFFF, 3) if unmapped or prim[0] == 0xFFFFFFFF else (prim[0], prim[1])
        self.mapq = int(prim[6]) & 255
        self.cigar = b""
        self.alts = [[], []]
        self.xa_cigars = []
        self.md = None; self.nm = 0; self.xv = np.zeros(0, np.uint16)
        self.seq_start, self.seq_end = 0, L - 1
        if self.pos != 0xFFFFFFFF:
            aligned = self.rseq if self.strand else self.seq
            self.cigar = (oracle.ed_diff_withcigar(g.mixref, self.pos, aligned, prim[2], 128)[1] if prim[3] == 1 else "%dM" % L).encode()
            self.alts = [[a for a in alts[0]], [a for a in alts[1]]]
            for s in (0, 1):
                for (p, nd, gap, _) in self.alts[s]:
                    if p != self.pos and gap:
                        self.xa_cigars.append(oracle.ed_diff_withcigar(g.mixref, p, self.rseq if s else self.seq, nd, 256)[1].encode())
            self.tags(oracle, g)

    def tags(self, oracle, g):
        """MD / NM / XV of the read as it stands (position, strand, CIGAR, seq_start): what the tail kernels deliver"""
        aligned = self.rseq if self.strand == 1 else self.seq
        txt = oracle.md_nm(g.mixref, g.pac, g.l, aligned, self.pos, self.seq_start, self.cigar.decode())
        m = re.match(r"\tMD:Z:([^\t]*)\tNM:i:(\d+)(?:\tXV:i:([\d,]+))?$", txt)
        assert m, txt
        self.md = m.group(1).encode(); self.nm = int(m.group(2))
        self.xv = np.array([int(x) for x in m.group(3).split(",")], np.uint16) if m.group(3) else np.zeros(0, np.uint16)

    def mine(self, with_tags):
        r = SamReadT()
        r.name = self.name; r.seq = self.seq.ctypes.data; r.qual = self.qual; r.l_seq = len(self.seq); r.pos = self.pos
        r.strand = self.strand & 255; r.mapq = self.mapq; r.cigar = self.cigar; r.seq_start = self.seq_start; r.seq_end = self.seq_end
        self._alt = [(HitT * max(1, len(a)))(*[HitT(p, nd, gap, s) for (p, nd, gap, s) in a]) for a in self.alts]
        for s in (0, 1):
            r.n_alt[s] = len(self.alts[s]); r.alt[s] = C.cast(self._alt[s], C.POINTER(HitT))
        self._xa = (C.c_char_p * max(1, len(self.xa_cigars)))(*self.xa_cigars)
        r.xa_cigars = C.cast(self._xa, C.POINTER(C.c_char_p))
        r.md = self.md if with_tags else None; r.nm = self.nm; r.xv = self.xv.ctypes.data if len(self.xv) else None; r.n_xv = len(self.xv)
        return r

    def theirs(self):
        r = RefReadT()
        r.name = self.name; r.seq = self.seq.ctypes.data; r.rseq = self.rseq.ctypes.data; r.qual = self.qual; r.l_seq = len(self.seq)
        r.pos = self.pos; r.strand = self.strand; r.mapq = self.mapq; r.cigar = self.cigar; r.seq_start = self.seq_start; r.seq_end = self.seq_end
        self._ralt = [np.array([x for (p, nd, gap, s) in a for x in (p, nd, gap)], np.uint32) for a in self.alts]
        for s in (0, 1):
            r.n_alt[s] = len(self.alts[s]); r.alt[s] = self._ralt[s].ctypes.data if len(self._ralt[s]) else None
        return r


@pytest.fixture(scope="module")
def world(oracle):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref not built (reference tree absent at build time)")
    g = synth.Genome(40000, snp_rate=0.02, n_rate=0.002, seed=321)
    # a second copy of [2000, 12000) at [22000, 32000): every read from the first copy has an alternate in the second (XA)
    A, B, span = 2000, 22000, 10000
    g.codes[B:B + span] = g.codes[A:A + span]; g.masks[B:B + span] = g.masks[A:A + span]
    g.mixref = synth.pack_mixref(g.masks); g.pac = synth.pack_pac(g.codes)
    reads, pos, strand = synth.sample_reads(g, 120, 100, seed=322, sub_rate=0.02, indel_frac=0.4, n_frac=0.02)
    offs0, loci0, offs1, loci1 = synth.make_candidates(g, pos, strand, 100, per_strand=5, seed=323)
    l0n, l1n, o0n, o1n = [], [], [0], [0]
    for i in range(len(reads)):                           # add the other copy's locus to the list of the read's own strand
        a = list(loci0[offs0[i]:offs0[i + 1]]); b = list(loci1[offs1[i]:offs1[i + 1]])
        twin = int(pos[i]) + (B - A) if A <= pos[i] < A + span - 110 else (int(pos[i]) - (B - A) if B <= pos[i] < B + span - 110 else None)
        if twin is not None:
            (b if strand[i] else a).append(twin)
        a = sorted(set(a)); b = sorted(set(b))
        l0n += a; l1n += b; o0n.append(len(l0n)); o1n.append(len(l1n))
    offs0, loci0, offs1, loci1 = np.array(o0n, np.uint32), np.array(l0n, np.uint32), np.array(o1n, np.uint32), np.array(l1n, np.uint32)
    rng = np.random.default_rng(8)
    rd = []
    for i in range(len(reads)):
        qual = "".join(chr(33 + int(k)) for k in rng.integers(0, 41, 100))
        if i % 11 == 0:
            qual = ""                                   # no qualities: '*'
        r = Read(oracle, g, "read%d" % i, reads[i], qual, loci0[offs0[i]:offs0[i + 1]], loci1[offs1[i]:offs1[i + 1]], rng,
                 unmapped=(i % 13 == 5))
        if r.qual == b"" and r.pos != 0xFFFFFFFF and r.strand == 1:
            r.qual = None                               # aln_samse prints l_seq bytes of an EMPTY quality string on the reverse strand
        rd.append(r)
    names = [b"chrA", b"chr_B", b"c3"]
    offsets = [0, 9001, 23456]
    return g, rd, names, offsets


def _refs(names, offsets, l):
    r = SamRefsT()
    nm = (C.c_char_p * len(names))(*names); of = (C.c_int64 * len(offsets))(*offsets)
    r.n_seqs = len(names); r.names = C.cast(nm, C.POINTER(C.c_char_p)); r.offsets = C.cast(of, C.POINTER(C.c_int64)); r.l_pac = l
    return r, (nm, of)


def test_sam_se_lines(world):
    g, rd, names, offsets = world
    H = _hostlib(); R = C.CDLL(REF)
    refs, keep = _refs(names, offsets, g.l)
    nm = (C.c_char_p * len(names))(*names); of = (C.c_int64 * len(offsets))(*offsets)
    out = C.create_string_buffer(8192); ref_out = C.create_string_buffer(8192)
    H.salt_sam_se.restype = C.c_int; R.ref_sam_se.restype = C.c_int
    n_xa = n_gap_xa = n_rev = n_unmapped = 0
    for i, r in enumerate(rd):
        for xa_cigar, tags, rg in ((1, 1, b"grp1"), (0, 1, None), (1, 0, None), (0, 0, b"x")):
            mine = r.mine(tags); theirs = r.theirs()
            n = H.salt_sam_se(C.byref(refs), C.byref(mine), xa_cigar, rg, out, len(out))
            m = R.ref_sam_se(g.mixref.ctypes.data, C.c_uint32(g.l), g.pac.ctypes.data, len(names), nm, of, C.byref(theirs), xa_cigar, tags, rg,
                             ref_out, len(ref_out))
            assert n == m and out.value == ref_out.value, (i, xa_cigar, tags, out.value, ref_out.value)
        n_xa += b"XA:Z:" in out.value or any(p != r.pos for a in r.alts for (p, *_) in a)
        n_gap_xa += len(r.xa_cigars) > 0
        n_rev += r.strand == 1; n_unmapped += r.pos == 0xFFFFFFFF
    assert n_xa >= 20 and n_gap_xa >= 3 and n_rev >= 30 and n_unmapped >= 5
    # a buffer that is too small is refused, a position beyond the reference too
    r = next(x for x in rd if x.pos != 0xFFFFFFFF)
    mine = r.mine(1)
    assert H.salt_sam_se(C.byref(refs), C.byref(mine), 1, None, out, 40) == -103
    mine.pos = g.l + 5
    assert H.salt_sam_se(C.byref(refs), C.byref(mine), 1, None, out, len(out)) == -101


def test_sam_pe_lines(world, oracle):
    g, rd, names, offsets = world
    H = _hostlib(); R = C.CDLL(REF)
    refs, keep = _refs(names, offsets, g.l)
    nm = (C.c_char_p * len(names))(*names); of = (C.c_int64 * len(offsets))(*offsets)
    o = [C.create_string_buffer(8192) for _ in range(4)]
    ln = (C.c_int * 2)(); rln = (C.c_int * 2)()
    rng = np.random.default_rng(3)
    seen = set()
    for k in range(0, len(rd) - 1, 2):
        a, b = rd[k], rd[k + 1]
        for trial in range(3):
            for x in (a, b):                                 # soft clips as a rescue leaves them
                x.seq_start, x.seq_end = (0, 99) if trial == 0 else (int(rng.integers(0, 8)), 99 - int(rng.integers(0, 8)))
            if trial == 2 and a.pos != 0xFFFFFFFF:           # a pair inside the insert bounds: mate b next to mate a, other strand
                b.pos = min(g.l - 120, a.pos + 300); b.strand = 1 - (a.strand & 1)
                if not b.cigar:
                    b.cigar = b"100M"
                b.alts = [[], []]; b.xa_cigars = []
            for x in (a, b):
                if x.qual is None:
                    x.qual = b""                          # alnpe_sam takes strlen() of it on either strand
                if x.pos != 0xFFFFFFFF:
                    x.cigar = b"%dM" % (x.seq_end - x.seq_start + 1) if trial else x.cigar
                    x.alts = [[h for h in hs if not h[2]] for hs in x.alts] if trial else x.alts     # keep the XA list consistent
                    x.xa_cigars = [] if trial else x.xa_cigars
                    x.tags(oracle, g)
            mine = (SamReadT * 2)(a.mine(1), b.mine(1)); theirs = (RefReadT * 2)(a.theirs(), b.theirs())
            for (lo, hi) in ((250, 550), (0, 100000)):
                rc = H.salt_sam_pe(C.byref(refs), mine, lo, hi, 1, b"rg", o[0], len(o[0]), o[1], len(o[1]), ln)
                assert rc == 0
                R.ref_sam_pe(g.mixref.ctypes.data, C.c_uint32(g.l), g.pac.ctypes.data, len(names), nm, of, theirs, lo, hi, 1, 1, b"rg",
                             o[2], len(o[2]), o[3], len(o[3]), rln)
                assert (ln[0], ln[1]) == (rln[0], rln[1]) and o[0].value == o[2].value and o[1].value == o[3].value, \
                    (k, trial, o[0].value, o[2].value, o[1].value, o[3].value)
                f = int(o[0].value.split(b"\t")[1])
                seen.add((bool(f & 2), bool(f & 4), bool(f & 8)))
    assert (True, False, False) in seen and (False, True, False) in seen and (False, False, True) in seen


def test_sam_lines_large_coordinates(oracle):
    """24 records, 3.1 Gbp: positions beyond 2^31 map to the right record and print as the reference prints them (no tags, no XA
    CIGARs here: those would need the 3.1 Gbp reference itself; they do not depend on the coordinate arithmetic)"""
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref not built (reference tree absent at build time)")
    H = _hostlib(); R = C.CDLL(REF)
    l = 3_100_000_000
    names = [b"chr%d" % (i + 1) for i in range(24)]
    offsets = [i * (l // 24) + (i * 7919) % 1000 for i in range(24)]; offsets[0] = 0
    refs, keep = _refs(names, offsets, l)
    nm = (C.c_char_p * 24)(*names); of = (C.c_int64 * 24)(*offsets)
    dummy = np.zeros(64, np.uint32); dpac = np.zeros(64, np.uint8)
    rng = np.random.default_rng(12)
    out = [C.create_string_buffer(4096) for _ in range(4)]
    ln = (C.c_int * 2)(); rln = (C.c_int * 2)()
    H.salt_sam_se.restype = C.c_int; R.ref_sam_se.restype = C.c_int

    class Bare:
        pass

    def mk(pos, strand, alts):
        b = Bare()
        b.seq = rng.integers(0, 5, 100).astype(np.uint8); b.rseq = np.ascontiguousarray(synth.revcomp(b.seq))
        b.qual = bytes(33 + int(k) for k in rng.integers(0, 41, 100)); b.name = b"q%d" % pos
        b.alts = alts; b.ralt = [np.array([x for (p, nd, gap) in a for x in (p, nd, gap)], np.uint32) for a in alts]
        b.halt = [(HitT * max(1, len(a)))(*[HitT(p, nd, gap, s) for (p, nd, gap) in a]) for s, a in enumerate(alts)]
        m = SamReadT(); t = RefReadT()
        for r in (m, t):
            r.name = b.name; r.seq = b.seq.ctypes.data; r.qual = b.qual; r.l_seq = 100; r.pos = pos; r.mapq = 60; r.cigar = b"100M"
            r.seq_start = 0; r.seq_end = 99
            for s in (0, 1):
                r.n_alt[s] = len(alts[s])
        m.strand = strand; t.strand = strand; t.rseq = b.rseq.ctypes.data
        for s in (0, 1):
            m.alt[s] = C.cast(b.halt[s], C.POINTER(HitT)); t.alt[s] = b.ralt[s].ctypes.data if len(b.ralt[s]) else None
        b.m, b.t = m, t
        return b
    edge = [offsets[k] for k in (1, 12, 23)] + [offsets[k] - 1 for k in (1, 12, 23)] + [l - 101, 2**31 - 1, 2**31, 2**31 + 5, 4_000_000 % l]
    poss = edge + [int(x) for x in rng.integers(0, l - 101, 40)]
    items = [mk(p, i % 2, [[(int(rng.integers(0, l - 101)), 3, 0)], [(int(rng.integers(2**31, l - 101)), 1, 0)]] if i % 3 else [[], []])
             for i, p in enumerate(poss)]
    for b in items:
        n = H.salt_sam_se(C.byref(refs), C.byref(b.m), 0, None, out[0], 4096)
        m = R.ref_sam_se(dummy.ctypes.data, C.c_uint32(l), dpac.ctypes.data, 24, nm, of, C.byref(b.t), 0, 0, None, out[1], 4096)
        assert n == m and out[0].value == out[1].value, (out[0].value, out[1].value)
    for a, b in zip(items[::2], items[1::2]):
        mine = (SamReadT * 2)(a.m, b.m); theirs = (RefReadT * 2)(a.t, b.t)
        assert H.salt_sam_pe(C.byref(refs), mine, 250, 550, 0, b"g", out[0], 4096, out[1], 4096, ln) == 0
        R.ref_sam_pe(dummy.ctypes.data, C.c_uint32(l), dpac.ctypes.data, 24, nm, of, theirs, 250, 550, 0, 0, b"g", out[2], 4096, out[3], 4096, rln)
        assert out[0].value == out[2].value and out[1].value == out[3].value, (out[0].value, out[2].value)
