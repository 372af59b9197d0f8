#!/usr/bin/env python
"""Per-kernel timings of the BASELINE.json configurations that are NOT the headline bench line (configs 3, 4, 5)
on ONE B200, at a bounded sample size.  A tuning aid: bench.py stays the contract, this prints one JSON line per
configuration with the CUDA-event time of every kernel (rkg_profile_read).

    python tools/bench_configs.py --config 3 [--reads 50000]
    python tools/bench_configs.py --config 4 [--reads 5000000]
    python tools/bench_configs.py --config 5 [--bases 400000000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DATA = os.path.join(ROOT, "tests", "golden", "data")


def kernels(gpu):
    prof = gpu.profile_read(reset=True)
    return {k: {"launches": v["launches"], "ms": round(v["ms"], 3)} for k, v in prof.items()}


def timed(fn, reps):
    import torch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps


def config3(args, gpu):
    import rkmh_b200
    from rkmh_b200.fastx import load_sequences
    from rkmh_b200.synth import fast_ont_reads
    _, refs, _ = load_sequences([os.path.join(DATA, "hpv_16_allFasta.fa.gz")])
    n, L, k, s = args.reads or 50_000, 10_000, 12, 1000
    m = max(1, round(100 * n / 2_000_000))            # -M 100 at 2 M reads, scaled to the sample
    bases, offs = fast_ont_reads(3, refs, 0, n, L)
    gpu.set_params([k], s, read_table_size=rkmh_b200.STREAM_TABLE_SIZE)
    gpu.load_refs(*rkmh_b200.pack_sequences(refs))
    batch = gpu.upload(bases, offs)
    gpu.count_reset()
    batch.count(); batch.classify(min_occ=m)           # warm-up
    gpu.synchronize()
    gpu.profile_enable(True); gpu.profile_read(reset=True)
    ms_reset = timed(gpu.count_reset, 1)
    ms_count = timed(batch.count, args.reps)
    kc = kernels(gpu)
    ms_cls = timed(lambda: batch.classify(min_occ=m), args.reps)
    kk = kernels(gpu)
    hits = batch.results()
    ms_cls0 = timed(lambda: batch.classify(min_occ=-1), args.reps)
    k0 = kernels(gpu)
    gpu.profile_enable(False)
    batch.free()
    nk = n * (L - k)
    return {"config": 3, "reads": n, "read_len": L, "k": k, "s": s, "M": m, "refs": len(refs),
            "count_reset_ms": ms_reset, "count_pass_ms": ms_count, "classify_pass_ms": ms_cls,
            "classify_noM_ms": ms_cls0,
            "reads_per_s_two_pass": n / ((ms_count + ms_cls) / 1e3),
            "kmers_per_s_count": nk / (ms_count / 1e3), "kmers_per_s_classify": nk / (ms_cls / 1e3),
            "kernels_count": kc, "kernels_classify": kk, "kernels_classify_noM": k0,
            "mean_best": float(hits["best"].mean()), "mean_sketch_len": float(hits["sketch_len"].mean())}


def config4(args, gpu):
    import rkmh_b200
    from rkmh_b200.fastx import load_sequences
    from rkmh_b200.synth import mixed_reads
    refs = []
    for f in ("zika.fa.gz", "dengue.fa.gz", "yellow_fever.fa.gz", "hpv_16.fa.gz"):
        refs += load_sequences([os.path.join(DATA, f)])[1]
    n, L, k, s = args.reads or 5_000_000, 150, 20, 2000
    bases, offs, viral = mixed_reads(np.random.default_rng(4), refs, n, L)
    gpu.set_params([k], s)
    gpu.load_refs(*rkmh_b200.pack_sequences(refs), per_sample_set=True)
    batch = gpu.upload(bases, offs)
    batch.ingest(); batch.classify()
    gpu.synchronize()
    gpu.profile_enable(True); gpu.profile_read(reset=True)
    ms = timed(lambda: (batch.ingest(), batch.classify()), args.reps)
    kk = kernels(gpu)
    gpu.profile_enable(False)
    hits = batch.results()
    passing = (hits["sketch_len"] > 0) & (np.maximum(hits["best"], 0) >= 10)
    batch.free()
    return {"config": 4, "reads": n, "read_len": L, "k": k, "s": s, "refs": len(refs), "step_ms": ms,
            "reads_per_s": n / (ms / 1e3), "kmers_per_s": n * (L - k) / (ms / 1e3), "kernels": kk,
            "viral_frac": float(viral.mean()), "pass_N10": float(passing.mean()),
            "viral_recall_N10": float(passing[viral].mean()), "background_pass_N10": float(passing[~viral].mean())}


def config5(args, gpu):
    from rkmh_b200.synth import random_genome
    nb, k, s = args.bases or 400_000_000, 21, 10_000
    seq = random_genome(np.random.default_rng(5), nb)
    from rkmh_b200.api import PinnedBuffer
    gpu.set_params([k], s)
    gpu.sketch_shard(seq[:5_000_000], 0, 5_000_000 - k, nb)     # warm-up
    gpu.sketch_shard(seq, 0, nb - k, nb)                         # sizes the device buffers
    gpu.profile_enable(True); gpu.profile_read(reset=True)
    t0 = time.perf_counter()
    sk = gpu.sketch_shard(seq, 0, nb - k, nb)
    ms = 1e3 * (time.perf_counter() - t0)
    kk = kernels(gpu)
    pin = PinnedBuffer(gpu, nb)
    pv = pin.view(np.uint8, nb)
    pv[:] = seq
    gpu.profile_read(reset=True)
    t0 = time.perf_counter()
    sk2 = gpu.sketch_shard(pv, 0, nb - k, nb)
    ms_pinned = 1e3 * (time.perf_counter() - t0)
    same = bool(np.array_equal(sk, sk2))
    pin.free()
    gpu.profile_enable(False)
    return {"config": 5, "bases": nb, "k": k, "s": s, "shard_ms_pageable_host_buffer": ms, "shard_ms_pinned_host_buffer": ms_pinned,
            "kmers_per_s_pinned": (nb - k) / (ms_pinned / 1e3), "kernels": kk, "sketch_len": int(len(sk)), "same": same,
            "sketch_head": [int(x) for x in sk[:3]]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[3, 4, 5])
    ap.add_argument("--reads", type=int, default=0)
    ap.add_argument("--bases", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import rkmh_b200
    with rkmh_b200.RkmhGpu(0) as gpu:
        out = {3: config3, 4: config4, 5: config5}[args.config](args, gpu)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
