#!/usr/bin/env python
"""The single-end program re-staged on this repository's own libraries only -- no reference code in the loop:

    FASTQ text --salt_fastq_pack--> reads --salt_chunk_seed_verify (seeding, locate, verification on the GPU)-->
    --salt_chunk_result (hit selection, mapq, CIGAR)--> --salt_chunk_tail / salt_b200_lv_cigar (MD NM XV, XA CIGARs)-->
    --salt_sam_se--> SAM

on an index written by the reference's salt-idx.  The counterpart of `salt [-r N] [-m N] [-s N] [-c] [-d] [-g RG] [-v] PREFIX reads.fq`
(aln.c:138-226; alnse_core, alnse.c:1353-1480); its output equals the reference program's except for the @PG line
(tests/test_native_pipeline.py runs it on the SIMT emulator against oracle/_ref/salt).

    python tools/salt_se.py -d -c -r 1 -m 500 PREFIX reads.fq > out.sam
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from salt_b200 import api, host_api, index_io          # noqa: E402

MAX_N_PERSEQ = 200          # alnse.c:1281
MAX_HITS = 5                # aln.h:139


class FastqT(C.Structure):
    _fields_ = [("bases", C.c_void_p), ("bases_cap", C.c_size_t), ("n_pos", C.c_void_p), ("n_pos_cap", C.c_size_t),
                ("lens", C.c_void_p), ("n_ambiguous", C.c_void_p), ("name_off", C.c_void_p), ("name_len", C.c_void_p),
                ("comment_off", C.c_void_p), ("comment_len", C.c_void_p), ("qual_off", C.c_void_p),
                ("n_reads", C.c_uint32), ("n_bases", C.c_size_t), ("n_n", C.c_size_t)]


class HitT(C.Structure):
    _fields_ = [("pos", C.c_uint32), ("n_diff", C.c_uint8), ("is_gap", C.c_uint8), ("strand", C.c_uint16)]


class SamRefsT(C.Structure):
    _fields_ = [("n_seqs", C.c_int), ("names", C.POINTER(C.c_char_p)), ("offsets", C.POINTER(C.c_int64)), ("l_pac", C.c_int64)]


class SamReadT(C.Structure):
    _fields_ = [("name", C.c_char_p), ("seq", C.c_void_p), ("qual", C.c_char_p), ("l_seq", C.c_uint32), ("pos", C.c_uint32),
                ("strand", C.c_uint8), ("mapq", C.c_uint32), ("cigar", C.c_char_p), ("seq_start", C.c_uint32), ("seq_end", C.c_uint32),
                ("n_alt", C.c_int * 2), ("alt", C.POINTER(HitT) * 2), ("xa_cigars", C.POINTER(C.c_char_p)),
                ("md", C.c_char_p), ("nm", C.c_uint32), ("xv", C.c_void_p), ("n_xv", C.c_int)]


def read_ann(path):
    """PREFIX.C.ann as bns_dump writes it (bntseq.c): 'l_pac n_seqs seed', then per record 'gi name anno' / 'offset len n_ambs'"""
    lines = open(path).read().split("\n")
    l_pac, n_seqs = int(lines[0].split()[0]), int(lines[0].split()[1])
    names, offsets = [], []
    for i in range(n_seqs):
        names.append(lines[1 + 2 * i].split(" ")[1].encode())
        offsets.append(int(lines[2 + 2 * i].split()[0]))
    return l_pac, names, offsets


def parse_fastq(H, text):
    """the whole text through salt_fastq_pack: codes per read, names, quality strings, ambiguity counts"""
    H.salt_fastq_pack.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_uint32, C.POINTER(FastqT), C.POINTER(C.c_size_t)]
    cap = len(text) + 8
    m = text.count(b"\n") // 2 + 2
    bases = np.zeros(cap // 4 + 2, np.uint8); n_pos = np.zeros(cap, np.uint32)
    arrs = {k: np.zeros(m, dt) for k, dt in (("lens", np.uint16), ("n_ambiguous", np.uint16), ("name_off", np.uint32), ("name_len", np.uint16),
                                              ("comment_off", np.uint32), ("comment_len", np.uint16), ("qual_off", np.uint32))}
    fq = FastqT(bases.ctypes.data, cap, n_pos.ctypes.data, cap, *(arrs[k].ctypes.data for k in
                ("lens", "n_ambiguous", "name_off", "name_len", "comment_off", "comment_len", "qual_off")), 0, 0, 0)
    used = C.c_size_t(0)
    n = H.salt_fastq_pack(text, len(text), 1, m, C.byref(fq), C.byref(used))
    if n < 0:
        raise RuntimeError("salt_fastq_pack: %d" % n)
    nb = fq.n_bases
    codes = ((bases[np.arange(nb) >> 2] >> (2 * (np.arange(nb) & 3)).astype(np.uint8)) & 3).astype(np.uint8)
    codes[n_pos[:fq.n_n]] = 4
    roffs = np.concatenate([[0], np.cumsum(arrs["lens"][:n].astype(np.int64))])
    reads = []
    for i in range(n):
        L = int(arrs["lens"][i])
        name = text[arrs["name_off"][i]:arrs["name_off"][i] + arrs["name_len"][i]]
        qo = int(arrs["qual_off"][i])
        qual = None
        if qo != 0xFFFFFFFF:                                   # L graphic characters from there (one line in what this tool reads)
            qual = bytes(text[qo:qo + L])
        reads.append((name, np.ascontiguousarray(codes[roffs[i]:roffs[i + 1]]), qual, int(arrs["n_ambiguous"][i])))
    return reads


def align(lib, H, prefix, fastq, l_overlap=0, max_seed=50, max_locate=1000, seed_only_ref=0, print_xa_cigar=False, print_nm_md=False,
          rg_id=None, chunk_reads=20000, device=0):
    """returns the SAM body (one bytes object per read, in input order)"""
    fm = index_io.FmIndex(prefix)
    l_pac, names, offsets = read_ann(prefix + ".C.ann")
    assert l_pac == fm.l
    pac = np.ascontiguousarray(np.fromfile(prefix + ".C.pac", np.uint8)[:(fm.l + 3) // 4])
    kw = {"lib": lib} if lib is not None else {"device": device}
    eng = api.Engine(fm.mixref, fm.l, pac, fm.l, **kw)
    eng.set_index(fm)
    opt = api.Engine.seed_opt(fm.l_seed, l_overlap, max_seed, max_locate, seed_only_ref)
    reads = parse_fastq(H, open(fastq, "rb").read())
    refs = SamRefsT()
    nm_arr = (C.c_char_p * len(names))(*names); of_arr = (C.c_int64 * len(offsets))(*offsets)
    refs.n_seqs = len(names); refs.names = C.cast(nm_arr, C.POINTER(C.c_char_p)); refs.offsets = C.cast(of_arr, C.POINTER(C.c_int64)); refs.l_pac = fm.l
    H.salt_sam_se.restype = C.c_int
    out = []
    buf = C.create_string_buffer(1 << 16)
    max_bases = sum(len(r[1]) for r in reads[:chunk_reads]) + 1024
    for b in range(0, len(reads), chunk_reads):
        part = reads[b:b + chunk_reads]
        ch = host_api.Chunk(H, len(part) + 8, max(max_bases, sum(len(r[1]) for r in part) + 1024), (len(part) + 8) * max_locate)
        slot_of = []
        for (name, codes, qual, amb) in part:
            slot_of.append(ch.add_read(codes, np.zeros(0, np.uint32), np.zeros(0, np.uint32)) if amb <= MAX_N_PERSEQ else -1)
        ch.seed_verify(eng, opt, 3, -1)                         # alnse.c:1079 / :1090 thresholds
        if print_nm_md:
            ch.tail(eng, 0)
        res = [ch.result(k, MAX_HITS) if k >= 0 else None for k in slot_of]
        # CIGARs of the gapped alternates that will be printed (sam.c:205, :215), one call for the chunk
        xa_pairs, xa_k, xa_owner = [], [], []
        if print_xa_cigar:
            for i, r in enumerate(res):
                if r is None or r[0][0] == 0xFFFFFFFF:
                    continue
                for s in (0, 1):
                    for (p, nd, gap, _) in r[1][s]:
                        if p != r[0][0] and gap:
                            xa_pairs.append(((slot_of[i] << 1) | s, p)); xa_k.append(nd); xa_owner.append(i)
        xa_cig = {}
        if xa_pairs:
            pairs = np.zeros(len(xa_pairs), api.PAIR_DT); pairs["rs"] = [x[0] for x in xa_pairs]; pairs["pos"] = [x[1] for x in xa_pairs]
            e, cg = eng.lv_cigar(pairs, np.array(xa_k, np.uint8), 256)
            for j, i in enumerate(xa_owner):
                assert int(e[j]) == xa_k[j], "XA CIGAR: edit distance changed"       # sam.c:219-223 exits there
                xa_cig.setdefault(i, []).append(api.cstr(cg[j]).encode())
        for i, (name, codes, qual, amb) in enumerate(part):
            r = res[i]
            if r is None:
                out.append(b"")                                  # alnse.c:1296: the read is skipped, its line stays empty
                continue
            (pos, strand, n_diff, is_gap, b0, b1, mapq), alts, cigar = r
            q = SamReadT()
            q.name = name; q.seq = codes.ctypes.data; q.qual = qual; q.l_seq = len(codes); q.pos = pos; q.strand = strand & 255
            q.mapq = mapq & 255; q.cigar = cigar.encode(); q.seq_start = 0; q.seq_end = len(codes) - 1
            keep_alt = [(HitT * max(1, len(alts[s])))(*[HitT(p, nd, gap, st) for (p, nd, gap, st) in alts[s]]) for s in (0, 1)]
            for s in (0, 1):
                q.n_alt[s] = len(alts[s]); q.alt[s] = C.cast(keep_alt[s], C.POINTER(HitT))
            xs = xa_cig.get(i, [])
            keep_xa = (C.c_char_p * max(1, len(xs)))(*xs)
            q.xa_cigars = C.cast(keep_xa, C.POINTER(C.c_char_p))
            keep_xv = None
            if print_nm_md and pos != 0xFFFFFFFF:
                md, nm, xv = ch.md(slot_of[i])
                keep_xv = np.array(xv, np.uint16)
                q.md = md.encode(); q.nm = nm; q.xv = keep_xv.ctypes.data if len(xv) else None; q.n_xv = len(xv)
            n = H.salt_sam_se(C.byref(refs), C.byref(q), int(print_xa_cigar), rg_id, buf, len(buf))
            if n < 0:
                raise RuntimeError("salt_sam_se: %d on %s" % (n, name))
            out.append(buf.raw[:n])
        ch.close()
    eng.close()
    return out, names, [(offsets[i + 1] if i + 1 < len(offsets) else fm.l) - offsets[i] for i in range(len(offsets))]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-r", type=int, default=0); ap.add_argument("-m", type=int, default=1000); ap.add_argument("-s", type=int, default=50)
    ap.add_argument("-c", action="store_true"); ap.add_argument("-d", action="store_true"); ap.add_argument("-v", action="store_true")
    ap.add_argument("-g", default=None)
    ap.add_argument("prefix"); ap.add_argument("fastq")
    a = ap.parse_args()
    H = host_api.load()
    body, names, lens = align(None, H, a.prefix, a.fastq, a.r, a.s, a.m, int(a.v), a.c, a.d, a.g.encode() if a.g else None)
    w = sys.stdout.buffer
    w.write(b"@HD\tVN:ec1fec2\tSO:unsorted\n")                   # aln_samhead, sam.c:55-84
    for nm, ln in zip(names, lens):
        w.write(b"@SQ\tSN:%s\tLN:%d\n" % (nm, ln))
    w.write(b"@RG\tID:%s\n" % (a.g.encode() if a.g else b"(null)"))
    w.write(b"@PG\tID:salt_b200\tPN:salt_se.py\n")
    for ln in body:
        w.write(ln + b"\n")


if __name__ == "__main__":
    main()
