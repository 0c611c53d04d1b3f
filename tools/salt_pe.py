#!/usr/bin/env python
"""The paired-end program re-staged on this repository's own libraries only -- no reference code in the loop:

    two FASTQ texts --salt_fastq_pack--> mates 2i, 2i+1 --salt_b200_seed_locate (alnse_seed_overlap + alnse_locate on the GPU)-->
    --salt_chunk_submit / _wait (verification, PE thresholds 3 / 3)--> --salt_chunk_pair (query_set_hits, pairing2 /
    pairing_singleton as plans, one Smith-Waterman batch per rescue flavour, apply, CIGARs of promoted alternates)-->
    --salt_b200_md_nm / salt_b200_lv_cigar (MD NM XV, XA CIGARs)--> --salt_sam_pe--> SAM

on an index written by the reference's salt-idx.  The counterpart of `salt -p [-a N] [-b N] [-r N] [-m N] [-s N] [-c] [-d] [-g RG]
PREFIX r1.fq r2.fq` (aln.c:138-226; alnpe_core / alnpe_core1, alnpe.c:478-615); its output equals the reference program's except
for the @PG line (tests/test_native_pipeline.py runs it on the SIMT emulator against oracle/_ref/salt).

One deliberate difference: where an SNP-context interval is wider than -m the reference locates a random subset of its rows
(srand(time(0)) / rand(), alnse.c:538-552), so two runs of the reference itself differ there; the device leaves that interval
out and this program counts the mates concerned (salt_b200_seed_status) instead of imitating a random draw.

    python tools/salt_pe.py -d -c -a 350 -b 650 -r 5 PREFIX r1.fq r2.fq > out.sam
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from salt_b200 import api, host_api, index_io          # noqa: E402
from salt_se import HitT, SamReadT, SamRefsT, parse_fastq, read_ann          # noqa: E402

MAX_N_PERSEQ = 5            # alnpe.c:477
MAX_HITS = 5                # aln.h:139
LIST_CAP = 4096             # room per candidate list on the device (the reference allows 262144, alnse.c:42)
UNMAPPED = 0xFFFFFFFF


def align(lib, H, prefix, fastq1, fastq2, min_tlen=250, max_tlen=550, l_overlap=0, max_seed=50, max_locate=1000, seed_only_ref=0,
          print_xa_cigar=False, print_nm_md=False, rg_id=None, chunk_pairs=10000, device=0, stats=None):
    """returns the SAM body (two bytes objects per pair, in input order, each with its trailing newline as alnpe_sam leaves them)"""
    fm = index_io.FmIndex(prefix)
    l_pac, names, offsets = read_ann(prefix + ".C.ann")
    assert l_pac == fm.l
    pac = np.ascontiguousarray(np.fromfile(prefix + ".C.pac", np.uint8)[:(fm.l + 3) // 4])
    kw = {"lib": lib} if lib is not None else {"device": device}
    eng = api.Engine(fm.mixref, fm.l, pac, fm.l, **kw)
    eng.set_index(fm)
    opt = api.Engine.seed_opt(fm.l_seed, l_overlap, max_seed, max_locate, seed_only_ref, locate_mode=1, list_cap=LIST_CAP)
    mates = [parse_fastq(H, open(f, "rb").read()) for f in (fastq1, fastq2)]
    n_pairs_all = min(len(mates[0]), len(mates[1]))             # query_read_multiPairedSeqs stops with the shorter file
    refs = SamRefsT()
    nm_arr = (C.c_char_p * len(names))(*names); of_arr = (C.c_int64 * len(offsets))(*offsets)
    refs.n_seqs = len(names); refs.names = C.cast(nm_arr, C.POINTER(C.c_char_p)); refs.offsets = C.cast(of_arr, C.POINTER(C.c_int64)); refs.l_pac = fm.l
    H.salt_sam_pe.restype = C.c_int
    out = []
    buf = [C.create_string_buffer(1 << 16), C.create_string_buffer(1 << 16)]
    ln = (C.c_int * 2)()
    tot = {"pairs": 0, "proper": 0, "windows16": 0, "windows5": 0, "rescued": 0, "promoted": 0, "declined": 0, "flagged_mates": 0}
    for b in range(0, n_pairs_all, chunk_pairs):
        n_pairs = min(chunk_pairs, n_pairs_all - b)
        part = [mates[m][b + p] for p in range(n_pairs) for m in (0, 1)]          # mates 2p, 2p+1 of pair p
        n = len(part)
        codes = np.concatenate([r[1] for r in part]) if n else np.zeros(0, np.uint8)
        roffs = np.concatenate([[0], np.cumsum([len(r[1]) for r in part])]).astype(np.uint32)
        # phase 1: candidate lists of every mate, both strands (alnse_seed_overlap + alnse_locate, alnse.c:1010-1013)
        eng.set_reads(codes, roffs)
        offs0, loci0, offs1, loci1 = eng.seed_locate(opt)
        st0, st1 = eng.seed_status()
        tot["flagged_mates"] += int(np.count_nonzero(st0 | st1))
        skip = np.array([r[3] > MAX_N_PERSEQ for r in part])                      # alnpe.c:491: the mate is not aligned (it can still be rescued)
        if skip.any():
            cnt0 = np.diff(offs0.astype(np.int64)); cnt1 = np.diff(offs1.astype(np.int64))
            keep0 = np.repeat(~skip, cnt0); keep1 = np.repeat(~skip, cnt1)
            loci0, loci1 = loci0[keep0], loci1[keep1]
            cnt0[skip] = 0; cnt1[skip] = 0
            offs0 = np.concatenate([[0], np.cumsum(cnt0)]).astype(np.uint32); offs1 = np.concatenate([[0], np.cumsum(cnt1)]).astype(np.uint32)
        # phase 2: verification with the paired-end thresholds (alnse.c:1016, :1027), then the pair stage of the whole chunk
        ch = host_api.Chunk(H, n + 8, len(codes) + 1024, max(len(loci0), len(loci1)) + 64)
        ch.add_reads(codes, roffs, offs0, loci0, offs1, loci1)
        ch.submit(eng, 0, 3, 3); ch.wait(eng, 0)
        finals, _, _, st = ch.pair(eng, 0, n_pairs, min_tlen, max_tlen, fm.l, max_hits=MAX_HITS, with_tail=False)
        for k in ("pairs", "proper", "windows16", "windows5", "rescued", "promoted", "declined"):
            tot[k] += getattr(st, k)
        fin = [finals[p].mate[m] for p in range(n_pairs) for m in (0, 1)]
        res = [ch.result(i, MAX_HITS) for i in range(n)]                          # query->hits of every mate (pairing does not change them)
        # tags of every mapped mate as it stands after pairing (sam_add_md_nm, sam.c:246-328), one call for the chunk
        tags = {}
        if print_nm_md:
            idx = [i for i in range(n) if fin[i].pos != UNMAPPED]
            if idx:
                o, md, xv = eng.md_nm(np.array([(i << 1) | (fin[i].strand & 1) for i in idx], np.uint32),
                                      np.array([fin[i].pos for i in idx], np.uint32), np.array([fin[i].seq_start for i in idx], np.uint32),
                                      [fin[i].cigar.decode() for i in idx], md_stride=512, xv_stride=64, slot=0)
                for j, i in enumerate(idx):
                    if int(o["md_len"][j]) < 0:
                        raise RuntimeError("MD string of %s: code %d" % (part[i][0], int(o["md_len"][j])))
                    tags[i] = (api.cstr(md[j]).encode(), int(o["nm"][j]), np.array(xv[j, :int(o["n_xv"][j])], np.uint16))
        # CIGARs of the gapped alternates that will be printed (sam_add_xa, sam.c:205, :215), one call for the chunk
        xa_pairs, xa_k, xa_owner = [], [], []
        if print_xa_cigar:
            for i in range(n):
                for s in (0, 1):
                    for (p, nd, gap, _) in res[i][1][s]:
                        if p != fin[i].pos and gap:
                            xa_pairs.append(((i << 1) | s, p)); xa_k.append(nd); xa_owner.append(i)
        xa_cig = {}
        if xa_pairs:
            pairs = np.zeros(len(xa_pairs), api.PAIR_DT); pairs["rs"] = [x[0] for x in xa_pairs]; pairs["pos"] = [x[1] for x in xa_pairs]
            e, cg = eng.lv_cigar(pairs, np.array(xa_k, np.uint8), 256)
            for j, i in enumerate(xa_owner):
                assert int(e[j]) == xa_k[j], "XA CIGAR: edit distance changed"       # sam.c:219-223 exits there
                xa_cig.setdefault(i, []).append(api.cstr(cg[j]).encode())
        for p in range(n_pairs):
            q = (SamReadT * 2)()
            keep = []
            for m in (0, 1):
                i = 2 * p + m
                name, rc, qual, _ = part[i]
                f = fin[i]; alts = res[i][1]
                q[m].name = name; q[m].seq = rc.ctypes.data; q[m].qual = qual if qual is not None else b""; q[m].l_seq = len(rc)
                q[m].pos = f.pos; q[m].strand = f.strand; q[m].mapq = f.mapq & 255
                cg = f.cigar; q[m].cigar = cg; q[m].seq_start = f.seq_start; q[m].seq_end = f.seq_end
                ka = [(HitT * max(1, len(alts[s])))(*[HitT(pp, nd, gap, stv) for (pp, nd, gap, stv) in alts[s]]) for s in (0, 1)]
                for s in (0, 1):
                    q[m].n_alt[s] = len(alts[s]); q[m].alt[s] = C.cast(ka[s], C.POINTER(HitT))
                xs = xa_cig.get(i, [])
                kx = (C.c_char_p * max(1, len(xs)))(*xs)
                q[m].xa_cigars = C.cast(kx, C.POINTER(C.c_char_p))
                kv = None
                if i in tags:
                    md, nmv, kv = tags[i]
                    q[m].md = md; q[m].nm = nmv; q[m].xv = kv.ctypes.data if len(kv) else None; q[m].n_xv = len(kv)
                keep.append((ka, kx, kv, cg))
            rc = H.salt_sam_pe(C.byref(refs), q, int(min_tlen), int(max_tlen), int(print_xa_cigar), rg_id, buf[0], len(buf[0]), buf[1], len(buf[1]), ln)
            if rc < 0:
                raise RuntimeError("salt_sam_pe: %d on %s" % (rc, part[2 * p][0]))
            out.append(buf[0].raw[:ln[0]]); out.append(buf[1].raw[:ln[1]])
        ch.close()
    eng.close()
    if stats is not None:
        stats.update(tot)
    return out, names, [(offsets[i + 1] if i + 1 < len(offsets) else fm.l) - offsets[i] for i in range(len(offsets))]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-a", type=int, default=250); ap.add_argument("-b", type=int, default=550)
    ap.add_argument("-r", type=int, default=0); ap.add_argument("-m", type=int, default=1000); ap.add_argument("-s", type=int, default=50)
    ap.add_argument("-c", action="store_true"); ap.add_argument("-d", action="store_true"); ap.add_argument("-v", action="store_true")
    ap.add_argument("-g", default=None)
    ap.add_argument("prefix"); ap.add_argument("fastq1"); ap.add_argument("fastq2")
    a = ap.parse_args()
    H = host_api.load()
    st = {}
    body, names, lens = align(None, H, a.prefix, a.fastq1, a.fastq2, a.a, a.b, a.r, a.s, a.m, int(a.v), a.c, a.d,
                              a.g.encode() if a.g else None, stats=st)
    w = sys.stdout.buffer
    w.write(b"@HD\tVN:ec1fec2\tSO:unsorted\n")                   # aln_samhead, sam.c:55-84
    for nm, ln in zip(names, lens):
        w.write(b"@SQ\tSN:%s\tLN:%d\n" % (nm, ln))
    w.write(b"@RG\tID:%s\n" % (a.g.encode() if a.g else b"(null)"))
    w.write(b"@PG\tID:salt_b200\tPN:salt_pe.py\n")
    for ln in body:
        w.write(ln + b"\n")                                      # alnpe.c:620: the line already ends in a newline, printf adds one
    sys.stderr.write("[salt_pe] %s\n" % st)


if __name__ == "__main__":
    main()
