"""Multi-GPU sharding of the verification path: reads shard, the reference is
replicated, there is no collective on the data path.

salt works through its input in chunks of N_SEQS = 100000 reads and prints SAM in input order
(alnse.c:1414-1440).  With G GPUs (one process per GPU) chunk c goes to rank c mod G; each rank
runs its chunks through its own engine; only the fixed-size result records travel afterwards
(to rank 0, which emits them in input order).  `torch.distributed` is plumbing here: NCCL on the
GPU box, gloo in the CPU tests.
"""
import numpy as np


def chunk_ranges(n_reads, chunk_reads):
    """[(begin, end)] of consecutive chunks."""
    return [(b, min(n_reads, b + chunk_reads)) for b in range(0, n_reads, chunk_reads)]


def my_chunks(n_reads, chunk_reads, world, rank):
    """Chunk indices and read ranges this rank owns: chunk c -> rank c mod world."""
    return [(c, b, e) for c, (b, e) in enumerate(chunk_ranges(n_reads, chunk_reads)) if c % world == rank]


def slice_csr(offs, loci, b, e):
    """The CSR sub-lists of reads [b, e) with offsets rebased to 0."""
    o = offs[b:e + 1].astype(np.int64)
    return (o - o[0]).astype(np.uint32), loci[o[0]:o[-1]]


def gather_in_order(local, n_reads, chunk_reads, world, rank, dist=None, dst=0):
    """local: {chunk index: structured numpy array of that chunk's per-read records}.
    Returns, on rank dst, the records of all reads in input order (None elsewhere).
    The wire format is raw bytes of fixed-size records: one gather per call, sized by the widest rank."""
    import torch
    dt = None
    for v in local.values():
        dt = v.dtype
    mine = my_chunks(n_reads, chunk_reads, world, rank)
    for c, b, e in mine:
        if c not in local or len(local[c]) != e - b:
            raise ValueError("rank %d: chunk %d missing or of the wrong size" % (rank, c))
    if world == 1 or dist is None:
        out = np.empty(n_reads, dt)
        for c, b, e in mine:
            out[b:e] = local[c]
        return out
    # record size must agree across ranks (a rank may own no chunk)
    isz = torch.tensor([0 if dt is None else dt.itemsize], dtype=torch.int64)
    dist.all_reduce(isz, op=dist.ReduceOp.MAX)
    itemsize = int(isz.item())
    per_rank = [sum(e - b for _, b, e in my_chunks(n_reads, chunk_reads, world, r)) for r in range(world)]
    cap = max(per_rank) * itemsize
    buf = torch.zeros(max(cap, 1), dtype=torch.uint8)
    if mine:
        flat = np.concatenate([local[c] for c, _, _ in mine]).view(np.uint8)
        buf[:len(flat)] = torch.from_numpy(flat.copy())
    bufs = [torch.zeros_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, bufs, dst=dst)
    if rank != dst:
        return None
    if dt is None:
        raise ValueError("destination rank owns no chunk and cannot know the record type")
    out = np.empty(n_reads, dt)
    for r in range(world):
        raw = bufs[r].numpy()
        at = 0
        for c, b, e in my_chunks(n_reads, chunk_reads, world, r):
            nb = (e - b) * itemsize
            out[b:e] = raw[at:at + nb].view(dt)
            at += nb
    return out
