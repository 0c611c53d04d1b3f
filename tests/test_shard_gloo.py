"""Host-side multi-GPU logic on CPU: two gloo ranks shard the chunks of a batch (chunk c -> rank
c mod 2), run the verification stage of their chunks (the SIMT-emulated engine stands in for the
GPU), and rank 0 merges the per-read records in input order; the merge must equal a single-process
run.  No data-path collective exists; gloo only carries the result records."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from salt_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_chunk_ownership():
    n, c = 1050, 100
    seen = np.zeros(n, int)
    for world in (1, 2, 3, 8):
        seen[:] = 0
        for r in range(world):
            for ci, b, e in shard.my_chunks(n, c, world, r):
                assert ci % world == r and e - b <= c
                seen[b:e] += 1
        assert (seen == 1).all()
    offs = np.array([0, 2, 2, 5, 9], np.uint32); loci = np.arange(9, dtype=np.uint32)
    o, l = shard.slice_csr(offs, loci, 1, 3)
    assert o.tolist() == [0, 0, 3] and l.tolist() == [2, 3, 4]


WORKER = textwrap.dedent("""
    import os, sys, ctypes as C
    import numpy as np
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests")); sys.path.insert(0, os.path.join({root!r}, "tests", "emul"))
    import torch.distributed as dist
    import parity_cases as pc, build_emul
    from salt_b200 import api, shard
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = api._declare(C.CDLL(build_emul.build()))
    g, reads, pos, strand, cands = pc.make_world(77, L=100, n_reads=46, per_strand=4, indel_frac=0.4, glen=30000)
    offs0, loci0, offs1, loci1 = cands
    eng = api.Engine(g.mixref, g.l, None, 0, lib=lib)
    n, chunk = len(reads), 8
    local = {{}}
    for c, b, e in shard.my_chunks(n, chunk, world, rank):
        o0, l0 = shard.slice_csr(offs0, loci0, b, e); o1, l1 = shard.slice_csr(offs1, loci1, b, e)
        eng.set_reads(reads[b:e])
        rec, a0, a1, cig = eng.verify(o0, l0, o1, l1, 3, -1)
        local[c] = rec
    merged = shard.gather_in_order(local, n, chunk, world, rank, dist)
    if rank == 0:
        eng.set_reads(reads)
        want = eng.verify(offs0, loci0, offs1, loci1, 3, -1)[0]
        assert merged.tobytes() == want.tobytes()
        assert (want["is_gap"] == 1).sum() >= 2
        print("MERGE_OK", len(merged))
    dist.barrier()
    dist.destroy_process_group()
""")


def test_two_rank_shard_and_merge(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
    import build_emul
    build_emul.build()                        # build once, not concurrently in both ranks
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "MERGE_OK 46" in outs[0]
