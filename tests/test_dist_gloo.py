"""CPU tests of the N > 1 host logic with world_size-2 gloo: read sharding, the -M count-table all-reduce and
the all-gather + merge of sharded sketches give exactly the single-rank result (checked with the oracle)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rkmh_b200.dist import allgather_sketches, allreduce_count_table, kmer_shard, max_over_ranks, shard_range
from rkmh_b200.fastx import load_sequences
from tests.util import data_path


def test_shard_range_partitions():
    for n in (0, 1, 7, 10, 1000003):
        for world in (1, 2, 3, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_kmer_shard_bytes():
    for L, k, world in ((100, 21, 4), (21, 21, 2), (22, 21, 8), (3_100_000, 21, 8)):
        got = set()
        for r in range(world):
            b, e, (lo, hi) = kmer_shard(L, k, r, world)
            got.update(range(b, e) if e - b < 1000 else ())
            assert hi <= L and (e == b or hi == e + k)
        if L < 1000:
            assert got == set(range(max(L - k, 0)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import Oracle
    ora = Oracle()
    _, reads, _ = load_sequences([data_path("mix2k.fq.gz")])
    reads = reads[:400]
    _, refs, _ = load_sequences([data_path("new_refs.fa.gz")])
    size = 1_000_003
    # -M: pass 1 on this rank's shard, all-reduce, pass 2 on this rank's shard
    lo, hi = shard_range(len(reads), rank, world)
    counts = np.zeros(size, np.int32)
    ora.count_batch(reads[lo:hi], [12], counts)
    t = torch.from_numpy(counts)
    allreduce_count_table(t)
    sk, ln = ora.sketch_batch(refs, [12], 1000)
    hits = ora.classify_batch(reads[lo:hi], [12], 1000, sk, ln, counts=counts, min_occ=3)
    # sharded hash of one record
    rec = ora.to_upper(b"".join(refs))
    k, s = 21, 2000
    b, e, (blo, bhi) = kmer_shard(len(rec), k, rank, world)
    part = ora.minhashes(ora.calc_hashes(rec[blo:bhi], [k]), s) if e > b else np.zeros(0, np.uint64)
    parts, lens = allgather_sketches(part, s)
    merged = parts[0, :lens[0]]
    for r in range(1, world):
        merged = ora.merge_sketches(merged, parts[r, :lens[r]], s)
    slowest = max_over_ranks(10.0 + rank)
    q.put((rank, counts.copy(), np.stack([hits[f] for f in ("ref_id", "best", "prev", "sketch_len")], 1), merged, slowest))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_matches_single_rank(oracle):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    _, reads, _ = load_sequences([data_path("mix2k.fq.gz")])
    reads = reads[:400]
    _, refs, _ = load_sequences([data_path("new_refs.fa.gz")])
    size = 1_000_003
    counts = np.zeros(size, np.int32)
    oracle.count_batch(reads, [12], counts)
    sk, ln = oracle.sketch_batch(refs, [12], 1000)
    hits = oracle.classify_batch(reads, [12], 1000, sk, ln, counts=counts, min_occ=3)
    exp = np.stack([hits[f] for f in ("ref_id", "best", "prev", "sketch_len")], 1)
    for rank, c, h, merged, slowest in res:
        assert np.array_equal(c, counts)                       # every rank holds the global table
        assert slowest == 11.0
    assert np.array_equal(np.concatenate([r[2] for r in res]), exp)   # sharded results == unsharded
    rec = oracle.to_upper(b"".join(refs))
    whole = oracle.minhashes(oracle.calc_hashes(rec, [21]), 2000)
    for r in res:
        assert np.array_equal(r[3], whole)
