"""Multi-GPU plumbing for the two global points of the path, over torch.distributed
(NCCL on the GPUs, gloo in the CPU tests).  Reads are independent units: they are sharded over ranks with no
data-path collective.  The only exchanges are

  * `-M`: the per-rank k-mer count tables are summed (all-reduce, int32) between pass 1 and pass 2
    (the `#pragma omp for` barrier at reference src/rkmh.cpp:910);
  * sharded `hash` of one large record: each rank sketches a contiguous range of k-mer starts, the bottom-s
    multisets are all-gathered and merged (first s of the multiset union == minhashes of the whole record).
"""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous, exhaustive, non-overlapping partition of range(n)."""
    return n * rank // world, n * (rank + 1) // world


def kmer_shard(record_len, k, rank, world):
    """k-mer start range of this rank for a record of record_len bases (len - k windows exist) and the
    byte range of the record it needs: starts [b, e) need bytes [b, e + k)."""
    nk = max(record_len - k, 0)
    b, e = shard_range(nk, rank, world)
    return b, e, (b, min(e + k, record_len) if e > b else b)


def allreduce_count_table(table_tensor):
    """Sum the read count tables of all ranks in place (table_tensor: int32 tensor on the rank's device)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(table_tensor, op=dist.ReduceOp.SUM)
    return table_tensor


def allgather_sketches(sketch, s, device=None):
    """All-gather one bottom-s multiset per rank.  sketch: uint64 numpy array (len <= s).
    Returns (parts uint64 [world, s], lens int32 [world])."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    row = np.zeros(s + 1, np.int64)
    row[0] = len(sketch)
    row[1:1 + len(sketch)] = np.asarray(sketch, np.uint64).view(np.int64)
    mine = torch.from_numpy(row)
    if device is not None:
        mine = mine.to(device)
    if world == 1:
        rows = [mine]
    else:
        rows = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(rows, mine)
    mat = torch.stack(rows).cpu().numpy()
    lens = mat[:, 0].astype(np.int32)
    parts = np.ascontiguousarray(mat[:, 1:]).view(np.uint64)
    return parts, lens


def max_over_ranks(value, device=None):
    """Max of a float over ranks (timing is the max over ranks, never a wall clock of one rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
