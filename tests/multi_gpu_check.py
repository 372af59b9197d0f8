#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun on a box with N >= 2 B200s (the driver's `pytest -m gpu` runs on
one GPU; tests/test_multi_gpu.py launches this script when more are visible):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py

Each rank owns one GPU through the C ABI.  Checks against the oracle:
  1. config-3 style `-M`: pass 1 on the rank's read shard, NCCL all-reduce of the count table, pass 2
     -> the gathered hits equal the single-process oracle result;
  2. config-5 style sharded `hash`: per-rank rkg_sketch_shard, all-gather, rkg_merge_sketches
     -> equals minhashes of the whole record.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rkmh_b200  # noqa: E402
from rkmh_b200.dist import allgather_sketches, allreduce_count_table, kmer_shard, shard_range  # noqa: E402
from rkmh_b200.fastx import load_sequences  # noqa: E402

DATA = os.path.join(ROOT, "tests", "golden", "data")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    gpu = rkmh_b200.RkmhGpu(local)
    _, refs, _ = load_sequences([os.path.join(DATA, "new_refs.fa.gz")])
    _, reads, _ = load_sequences([os.path.join(DATA, "mix2k.fq.gz")])
    ks, s, size, m = [12], 1000, rkmh_b200.STREAM_TABLE_SIZE, 5

    # 1. -M with the count table all-reduced over NVLink
    gpu.set_params(ks, s, read_table_size=size)
    gpu.load_refs(*rkmh_b200.pack_sequences(refs))
    lo, hi = shard_range(len(reads), rank, world)
    gpu.count_reset()
    gpu.count_batch(*rkmh_b200.pack_sequences(reads[lo:hi]))
    table = gpu.count_table_tensor()
    local_sum = int(table.sum().item())
    allreduce_count_table(table)
    torch.cuda.synchronize()
    global_sum = int(table.sum().item())
    hits = gpu.classify_batch(*rkmh_b200.pack_sequences(reads[lo:hi]), min_occ=m)
    # the same exchange on the per-code partial counts (64 MB instead of 800 MB): identical table, identical records
    gpu.count_reset()
    gpu.count_batch(*rkmh_b200.pack_sequences(reads[lo:hi]))
    allreduce_count_table(gpu.count_partial_tensor())
    torch.cuda.synchronize()
    compact_sum = int(gpu.count_table_tensor().sum().item())
    hits_c = gpu.classify_batch(*rkmh_b200.pack_sequences(reads[lo:hi]), min_occ=m)
    compact_ok = compact_sum == global_sum and bool(np.array_equal(hits_c.view(np.int32), hits.view(np.int32)))
    mine = torch.from_numpy(np.stack([hits[f] for f in ("ref_id", "best", "prev", "sketch_len")], 1).astype(np.int32)).cuda()
    sizes = [shard_range(len(reads), r, world) for r in range(world)]
    bufs = [torch.empty((b - a, 4), dtype=torch.int32, device="cuda") for a, b in sizes]
    dist.all_gather(bufs, mine)
    got = torch.cat(bufs).cpu().numpy()

    # 2. one record sharded over the ranks
    k2, s2 = 21, 10000
    rng = np.random.default_rng(5)
    rec = rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=4_000_000, p=[.2475, .2475, .2475, .2475, .01])
    gpu.set_params([k2], s2)
    b, e, (blo, bhi) = kmer_shard(rec.size, k2, rank, world)
    part = gpu.sketch_shard(rec[blo:bhi], b, e, rec.size)
    parts, lens = allgather_sketches(part, s2, device="cuda")
    merged = gpu.merge_sketches(parts, lens)

    ok = compact_ok
    if rank == 0:
        from oracle.oracle import Oracle
        ora = Oracle()
        counts = np.zeros(size, np.int32)
        ora.count_batch(reads, ks, counts)
        osk, ol = ora.sketch_batch(refs, ks, s)
        exp = ora.classify_batch(reads, ks, s, osk, ol, counts=counts, min_occ=m)
        exp = np.stack([exp[f] for f in ("ref_id", "best", "prev", "sketch_len")], 1)
        ok &= global_sum == int(counts.sum())
        ok &= bool(np.array_equal(got, exp))
        whole = ora.minhashes(ora.calc_hashes(rec.tobytes(), [k2]), s2)
        ok &= bool(np.array_equal(merged, whole))
        print(f"multi_gpu_check world={world}: count sum local {local_sum} global {global_sum} (oracle {int(counts.sum())}), "
              f"-M hits equal {np.array_equal(got, exp)}, sharded sketch equal {np.array_equal(merged, whole)}")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    gpu.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
