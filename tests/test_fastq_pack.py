"""Row f4 (input side): salt_fastq_pack -- FASTQ text straight into the compact transport -- against the reference's own
reader (query_open / query_read_seq, query.c:66-239, through oracle/_ref/libsaltref_seed.so): same codes, lengths,
ambiguity counts, names (with the /1 trim), comments and quality strings; block-wise parsing with carry-over."""
import ctypes as C
import os

import numpy as np
import pytest

from salt_b200 import api, host_api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libsaltref_seed.so")


class FastqT(C.Structure):
    _fields_ = [("bases", C.c_void_p), ("bases_cap", C.c_size_t), ("n_pos", C.c_void_p), ("n_pos_cap", C.c_size_t),
                ("lens", C.c_void_p), ("n_ambiguous", C.c_void_p), ("name_off", C.c_void_p), ("name_len", C.c_void_p),
                ("comment_off", C.c_void_p), ("comment_len", C.c_void_p), ("qual_off", C.c_void_p),
                ("n_reads", C.c_uint32), ("n_bases", C.c_size_t), ("n_n", C.c_size_t)]


def _hostlib():
    try:
        return host_api.load()
    except Exception:
        import sys
        sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
        import build_emul
        return host_api.load(build_emul.build_host())


def _write_fastq(path, rng, n, clean=False):
    # clean: long runs of plain A/C/G/T (the parser's 8- and 16-at-a-time paths) with an odd character now and then
    alphabet = "ACGT" * 60 + "acgtNn.R" if clean else "ACGTACGTACGTacgtNRYn."
    with open(path, "w") as f:
        for i in range(n):
            L = int(rng.integers(20, 260))
            seq = "".join(alphabet[k] for k in rng.integers(0, len(alphabet), L))
            qual = "".join(chr(33 + int(k)) for k in rng.integers(0, 41, L))
            name = "read%d" % i + ("/1" if i % 3 == 0 else "") + ("/x" if i % 11 == 0 else "")
            com = " comment %d with blanks" % i if i % 4 == 0 else ""
            if i % 5 == 0:                                   # multi-line sequence and quality (kseq accepts both)
                cut = L // 2
                f.write("@%s%s\n%s\n%s\n+%s\n%s\n%s\n" % (name, com, seq[:cut], seq[cut:], name if i % 10 == 0 else "", qual[:cut], qual[cut:]))
            elif i % 7 == 0:                                 # FASTA-style record without qualities
                f.write(">%s%s\n%s\n" % (name, com, seq))
            else:
                f.write("@%s%s\n%s\n+\n%s\n" % (name, com, seq, qual))


def _parse_all(H, text, block, max_reads_per_call):
    """drive salt_fastq_pack the way a reader loop would: fixed-size blocks, the unconsumed tail carried over"""
    recs = []
    pos = 0
    carry = b""
    while True:
        chunk = text[pos:pos + block]; pos += len(chunk)
        final = pos >= len(text)
        buf = carry + chunk
        if not buf:
            break
        cap = len(buf) + 8
        bases = np.zeros(cap // 4 + 2, np.uint8); n_pos = np.zeros(cap, np.uint32)
        m = max_reads_per_call
        arrs = {k: np.zeros(m, dt) for k, dt in (("lens", np.uint16), ("n_ambiguous", np.uint16), ("name_off", np.uint32),
                                                  ("name_len", np.uint16), ("comment_off", np.uint32), ("comment_len", np.uint16),
                                                  ("qual_off", np.uint32))}
        fq = FastqT(bases.ctypes.data, cap, n_pos.ctypes.data, cap, *(arrs[k].ctypes.data for k in
                    ("lens", "n_ambiguous", "name_off", "name_len", "comment_off", "comment_len", "qual_off")), 0, 0, 0)
        used = C.c_size_t(0)
        H.salt_fastq_pack.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_uint32, C.POINTER(FastqT), C.POINTER(C.c_size_t)]
        n = H.salt_fastq_pack(buf, len(buf), int(final), m, C.byref(fq), C.byref(used))
        assert n >= 0, n
        # unpack what came back
        codes = ((bases[np.arange(fq.n_bases) >> 2] >> (2 * (np.arange(fq.n_bases) & 3)).astype(np.uint8)) & 3).astype(np.uint8)
        codes[n_pos[:fq.n_n]] = 4
        at = 0
        for i in range(n):
            L = int(arrs["lens"][i])
            name = buf[arrs["name_off"][i]:arrs["name_off"][i] + arrs["name_len"][i]].decode()
            com = buf[arrs["comment_off"][i]:arrs["comment_off"][i] + arrs["comment_len"][i]].decode()
            qo = int(arrs["qual_off"][i])
            qual = b"" if qo == 0xFFFFFFFF else bytes(c for c in buf[qo:qo + 2 * L + 8] if c > 32)[:L]
            recs.append((codes[at:at + L].copy(), int(arrs["n_ambiguous"][i]), name, com, qual.decode()))
            at += L
        assert at == fq.n_bases
        carry = buf[used.value:]
        if final and (n == 0 or not carry.strip()):
            break
        if n == 0 and not final and len(carry) > 4 * block + 100000:
            raise AssertionError("parser makes no progress")
    return recs


@pytest.mark.parametrize("block,per_call,clean", [(1 << 20, 100000, False), (4096, 7, False), (700, 3, False),
                                                  (1 << 20, 100000, True), (4096, 7, True), (900, 3, True)])
def test_fastq_pack_matches_reference_reader(tmp_path, block, per_call, clean):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref not built (reference tree absent at build time)")
    H = _hostlib()
    rng = np.random.default_rng(17)
    fn = os.path.join(str(tmp_path), "r.fq")
    n = 400
    _write_fastq(fn, rng, n, clean)
    R = C.CDLL(REF)
    codes = np.zeros(n * 300, np.uint8); roffs = np.zeros(n + 1, np.uint32); n_amb = np.zeros(n, np.uint16)
    stride = 320
    names = np.zeros((n, stride), np.uint8); coms = np.zeros((n, stride), np.uint8); quals = np.zeros((n, stride), np.uint8)
    R.seedref_read_fastq.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int]
    got_n = R.seedref_read_fastq(fn.encode(), n, codes.ctypes.data, len(codes), roffs.ctypes.data, n_amb.ctypes.data,
                                 names.ctypes.data, coms.ctypes.data, quals.ctypes.data, stride)
    assert got_n == n
    text = open(fn, "rb").read()
    recs = _parse_all(H, text, block, per_call)
    assert len(recs) == n
    for i, (c, amb, name, com, qual) in enumerate(recs):
        want = codes[roffs[i]:roffs[i + 1]]
        assert np.array_equal(c, np.minimum(want, 4)), i
        assert amb == int(n_amb[i]), i
        assert name == api.cstr(names[i]), (i, name, api.cstr(names[i]))
        # a record without a comment: kseq leaves the previous record's comment in its buffer and query_read_seq copies
        # that (query.c:160; never printed anywhere) -- salt_fastq_pack reports the record's own, empty, comment
        if com or i == 0:
            assert com == api.cstr(coms[i]), (i, com, api.cstr(coms[i]))
        assert qual == api.cstr(quals[i]), (i, qual[:20], api.cstr(quals[i])[:20])
    assert sum(r[1] for r in recs) > (100 if clean else 1000)


def test_fastq_split_parts_equal_the_whole(tmp_path):
    """salt_fastq_split: four-line records whose quality lines start with '@' or '+' as often as chance allows; the parts parsed
    one by one give the same records as the whole text; multi-line records are declined"""
    H = _hostlib()
    H.salt_fastq_split.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_size_t)]
    H.salt_fastq_split.restype = C.c_int
    rng = np.random.default_rng(23)
    lines = []
    for i in range(3000):
        L = int(rng.integers(30, 160))
        seq = "".join("ACGTN"[k] for k in rng.choice(5, L, p=[0.245, 0.245, 0.245, 0.245, 0.02]))
        qual = "".join(chr(33 + int(k)) for k in rng.integers(0, 41, L))
        if i % 3 == 0: qual = "@" + qual[1:]
        if i % 3 == 1: qual = "+" + qual[1:]
        if i % 7 == 0: qual = "@@" + qual[2:]
        lines.append("@r%d%s\n%s\n+%s\n%s\n" % (i, " c%d" % i if i % 5 == 0 else "", seq, "r%d" % i if i % 4 == 0 else "", qual))
    text = "".join(lines).encode()
    whole = _parse_all(H, text, 1 << 26, 1000000)
    assert len(whole) == 3000
    for parts in (2, 5, 16, 64):
        cuts = (C.c_size_t * (parts + 1))()
        made = H.salt_fastq_split(text, len(text), parts, cuts)
        assert 1 <= made <= parts and cuts[0] == 0 and cuts[made] == len(text)
        got = []
        for k in range(made):
            assert cuts[k] < cuts[k + 1] and (k == 0 or text[cuts[k]:cuts[k] + 1] == b"@")
            got += _parse_all(H, text[cuts[k]:cuts[k + 1]], 1 << 26, 1000000)
        assert len(got) == len(whole), parts
        for a, b in zip(got, whole):
            assert np.array_equal(a[0], b[0]) and a[1:] == b[1:]
    # records whose sequence spans lines have no such header pattern everywhere: declined, never cut wrongly
    multi = b"".join(b"@m%d\nACGTACGT\nACGT\n+\nIIIIIIII\nIIII\n" % i for i in range(200))
    cuts = (C.c_size_t * 9)()
    assert H.salt_fastq_split(multi, len(multi), 8, cuts) in (-104, 1)
    fasta = b"".join(b">f%d\nACGTACGTAC\n" % i for i in range(200))
    assert H.salt_fastq_split(fasta, len(fasta), 4, cuts) in (-104, 1)        # no header of that shape anywhere: one part
