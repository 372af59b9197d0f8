#!/usr/bin/env python
"""bench.py -- rkmh classify throughput on B200 (BASELINE.json config 2:
`rkmh stream -r all_pave_ref.fa -k 12 -s 1000` on synthetic 150 bp Illumina-like reads, 182 references).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU implementation

One "step" = one pass of the hot path (ingest -> hash -> sketch -> probe -> argmax) over one batch of
`--reads` synthetic reads per GPU.  `value` is timed with the batch resident in HBM; `e2e` is timed through
rkg_classify_batch with pinned HOST buffers (H2D of the bases/offsets and D2H of the hit records inside
the timed region).  Prints ONE JSON line on rank 0.

Beside the headline (config 2, weak scaling) the line carries, each under the same barrier + max-over-ranks clock:
  strong_scaling   config 2 with 10 M reads IN TOTAL split over the ranks
  murmur_path      config 2 with the memoised hashing off: every k-mer is hashed (2 x MurmurHash3_x64_128)
  configs.cfg3     config 3 (`-M`, 10 kb ONT-like reads): count -> all-reduce (NCCL) -> classify, per phase
  configs.cfg4     config 4 (`filter`, k = 20: every k-mer hashed, 95 % background reads), kernel-only and end to end
  configs.cfg5     config 5 (`hash`, one record sharded over the ranks): sketch shard -> all-gather -> merge
  collectives      all-reduce / all-gather+merge times and bus bandwidth of those two sections
  parity_vs_n1     (inside cfg3 / cfg5) the N-rank result equals the same inputs run on ONE rank
  parity_spot      1000 reads of the timed batch checked against the oracle (untimed tail)
"""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
DATA = os.path.join(ROOT, "tests", "golden", "data")
K, S, READ_LEN = 12, 1000, 150
METRIC = "classify reads/s (rkmh stream, k=12, s=1000, 182 refs, 150 bp reads)"
WORKLOAD = ("BASELINE config 2: rkmh stream -r all_pave_ref.fa -k 12 -s 1000, synthetic 150 bp Illumina-like reads "
            "(90% sampled from the 182 refs, 0.5% subs, 0.1% N; 10% random)")
SEED2, SEED3, SEED5 = 2, 3, 5
# ncu captures of the kernels of THIS build (profiles/ncu_current.json, written by profiles/ncu_summary.py --current on
# the GPU box): DRAM bytes and warp instructions per launch with the number of reads of that launch, plus the hash of
# the kernel sources they were taken from.  A capture of other sources is reported as stale, never silently reused.
NCU_CURRENT = os.path.join(ROOT, "profiles", "ncu_current.json")
CSRC = os.path.join(ROOT, "rkmh_b200", "csrc")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per step (config 2: 10 M)")
    ap.add_argument("--cpu-sample", type=int, default=1_000_000, help="reads in the CPU baseline sample (about 10 s of the reference on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cfg3-reads", type=int, default=50_000, help="10 kb reads per GPU in the config-3 section (0 = skip)")
    ap.add_argument("--cfg4-reads", type=int, default=5_000_000, help="150 bp reads per GPU in the config-4 section (0 = skip)")
    ap.add_argument("--cfg5-bases", type=int, default=387_500_000, help="bases per GPU in the config-5 section (3.1 Gbp / 8; 0 = skip)")
    ap.add_argument("--only-headline", action="store_true", help="skip the strong-scaling / Murmur / config 3 / config 5 sections")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        # the median over the samples taken while the GPU was busy (clock at or above half of max): the bench has
        # host-only stretches (workload generation) during which an idle GPU clocks down
        busy = [x for x, m in zip(sm, mx) if x >= 0.5 * m]
        return {"sm_mhz": float(np.median(busy or sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "samples_under_load": len(busy), "reasons": sorted(reasons)}


def load_seqs(*files):
    from rkmh_b200.fastx import load_sequences
    names, seqs = [], []
    for f in files:
        n, s, _ = load_sequences([os.path.join(DATA, f)])
        names += n; seqs += s
    return names, seqs


def load_refs():
    return load_seqs("all_pave_ref.fa.gz")


def bench_config(args, world, nrefs):
    """The `config` object of the JSON line: identical for both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "reads_per_gpu_per_step": args.reads, "k": K, "s": S, "refs": nrefs,
            "parallelism": f"reads sharded x{world}", "generator": "std::mt19937_64, seed 2 (rkmh_b200/host/synth.cpp)",
            "l2": "inputs (1.5 GB of bases per step) are larger than L2; no flush needed"}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def config2_reads(refs, first, n, out=None):
    """Reads [first, first + n) of the config-2 stream (seed 2, std::mt19937_64: rkmh_b200/host/synth.cpp)."""
    from rkmh_b200.synth import fast_sample_reads
    return fast_sample_reads(SEED2, refs, first, n, READ_LEN, out=out)


def source_hash():
    h = hashlib.sha1()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh")):
            with open(os.path.join(CSRC, f), "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()[:16]


def ncu_current():
    """{kernel: {"dram_bytes", "warp_instructions", "reads", ...}} of the committed capture when it was taken from the
    kernel sources of this tree, else ({}, reason)."""
    try:
        with open(NCU_CURRENT) as fh:
            cur = json.load(fh)
    except (OSError, ValueError):
        return {}, "no capture (profiles/ncu_current.json)"
    if cur.get("source_hash") != source_hash():
        return {}, f"stale capture: taken from sources {cur.get('source_hash')}, this tree is {source_hash()}"
    return cur.get("kernels", {}), "profiles/ncu_current.json (same kernel sources as this build)"


# ---------------------------------------------------------------------------------------------------------
def run_reference_cli(bases, offs, nreads, threads, workdir):
    """Time the reference's own CPU implementation (oracle/_ref/rkmh_ref, built from /root/reference by
    oracle/Makefile) on `nreads` reads: wall clock of `rkmh stream -t threads` minus a references-only run."""
    from oracle.oracle import ref_binary
    from rkmh_b200.synth import write_fastq
    exe = ref_binary()
    if exe is None:
        return None
    import gzip
    ref_fa = os.path.join(workdir, "refs.fa")
    if not os.path.exists(ref_fa):
        with gzip.open(os.path.join(DATA, "all_pave_ref.fa.gz"), "rb") as fi, open(ref_fa, "wb") as fo:
            shutil.copyfileobj(fi, fo)
    fq = os.path.join(workdir, "sample.fq")
    write_fastq(fq, bases, offs[:nreads + 1])
    one = os.path.join(workdir, "one.fq")
    write_fastq(one, bases, offs[:2])
    cmd = [exe, "stream", "-r", ref_fa, "-k", str(K), "-s", str(S), "-t", str(threads)]

    def wall(reads_file):
        t0 = time.perf_counter()
        with open(os.devnull, "wb") as dn:
            subprocess.run(cmd + ["-f", reads_file], stdout=dn, stderr=dn, check=True)
        return time.perf_counter() - t0

    t_fixed = min(wall(one), wall(one))
    t_full = wall(fq)
    return {"seconds": t_full, "fixed_seconds": t_fixed, "reads": nreads,
            "reads_per_s": nreads / max(t_full - t_fixed, 1e-9)}


def run_oracle_port(bases, offs, nreads):
    """Fallback CPU baseline when oracle/_ref is absent: the single-threaded C restatement."""
    from oracle.oracle import Oracle
    _, refs = load_refs()
    ora = Oracle()
    sk, ln = ora.sketch_batch(refs, [K], S)
    reads = [bases[offs[i]:offs[i + 1]].tobytes() for i in range(nreads)]
    t0 = time.perf_counter()
    ora.classify_batch(reads, [K], S, sk, ln)
    dt = time.perf_counter() - t0
    return {"seconds": dt, "reads": nreads, "reads_per_s": nreads / dt}


def cpu_baseline(bases, offs, sample, workdir):
    cores = host_cores()
    r = run_reference_cli(bases, offs, sample, cores, workdir)
    if r is not None:
        return {"value": r["reads_per_s"], "unit": "reads/s", "cores": cores, "kind": "reference",
                "sample": f"{sample} reads of this workload through oracle/_ref/rkmh_ref stream -t {cores}: "
                          f"{r['seconds']:.2f} s wall minus {r['fixed_seconds']:.2f} s refs-only run",
                "kmers_per_s": r["reads_per_s"] * (READ_LEN - K)}
    sample = min(sample, 20000)
    r = run_oracle_port(bases, offs, sample)
    return {"value": r["reads_per_s"], "unit": "reads/s", "cores": 1, "kind": "port",
            "sample": f"{sample} reads through oracle/rkmh_oracle.c (single thread), {r['seconds']:.2f} s",
            "kmers_per_s": r["reads_per_s"] * (READ_LEN - K)}


def parity_spot(gpu_hits, bases, offs, refs, nspot=1000):
    """`nspot` reads of the timed batch (evenly spaced) through the oracle (test infrastructure: used here only as the
    checker, in the untimed tail) against the records the timed GPU run produced."""
    from oracle.oracle import Oracle
    ora = Oracle()
    n = len(offs) - 1
    idx = np.unique(np.linspace(0, n - 1, min(nspot, n)).astype(np.int64))
    sk, ln = ora.sketch_batch(refs, [K], S)
    reads = [bases[offs[i]:offs[i + 1]].tobytes() for i in idx]
    exp = ora.classify_batch(reads, [K], S, sk, ln)
    equal = all(bool(np.array_equal(gpu_hits[f][idx], exp[f])) for f in ("ref_id", "best", "prev", "sketch_len"))
    return {"reads": int(len(idx)), "equal": bool(equal), "checker": "oracle/rkmh_oracle.c (pinned to the reference's vectors)",
            "classified_fraction": float((exp["best"] > 0).mean())}


# ---------------------------------------------------------------------------------------------------------
def main_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation on the host cores, rank 0 only."""
    if rank != 0:
        return
    _, refs = load_refs()
    sample = args.cpu_sample
    bases, offs = config2_reads(refs, 0, sample)
    workdir = tempfile.mkdtemp(prefix="rkmh_ref_")
    cores = host_cores()
    try:
        from oracle.oracle import ref_binary
        kind = "reference" if ref_binary() else "port"
        times = []
        for i in range(args.warmup + args.steps):
            if kind == "reference":
                r = run_reference_cli(bases, offs, sample, cores, workdir)
                dt = r["seconds"] - r["fixed_seconds"]
            else:
                r = run_oracle_port(bases, offs, min(sample, 20000))
                dt, sample, cores = r["seconds"], r["reads"], 1
            if i >= args.warmup:
                times.append(dt)
    finally:
        shutil.rmtree(workdir, ignore_errors=True)
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": bench_config(args, args.gpus, len(refs)),
            "kmers_per_s": value * (READ_LEN - K),
            "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": kind,
                             "sample": f"each step is a bounded sample of the workload: the first {sample} reads of the config-2 "
                                       f"stream through the reference's own CPU implementation on the {cores} host cores of "
                                       "rank 0 (wall clock of the CLI minus a refs-only run)"},
            "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
class Ranks:
    """torch.distributed plumbing of one rank: barrier + synchronize, max over ranks, gathers."""

    def __init__(self, rank, world, local_rank):
        import torch
        self.torch, self.rank, self.world, self.local = torch, rank, world, local_rank
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            torch.cuda.set_device(local_rank)
            # NCCL may print its version banner on stdout while the communicator is built; stdout carries exactly one
            # JSON line, so anything printed during initialisation goes to stderr instead
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
                warm = torch.zeros(1, device=f"cuda:{local_rank}")
                dist.all_reduce(warm)
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)
            self.dist = dist
        torch.cuda.set_device(local_rank)
        self.dev = torch.device(f"cuda:{local_rank}")

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, v):
        if self.dist is None:
            return float(v)
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(self, flag):
        if self.dist is None:
            return bool(flag)
        t = self.torch.tensor([1 if flag else 0], dtype=self.torch.int32, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())

    def timed(self, fn, steps, warmup):
        """`steps` calls of fn bracketed by barrier + synchronize on both sides, wall clock, max over ranks -> ms per step.
        (For host-driven sections whose work spans several streams and collectives.)"""
        for _ in range(warmup):
            fn()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        self.torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0)
        self.barrier()
        return self.max(ms) / steps

    def gather_to_rank0(self, arr):
        """Concatenation over ranks (rank order) of equally shaped numpy arrays, on rank 0 (None elsewhere)."""
        if self.dist is None:
            return arr
        t = self.torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1)).to(self.dev)
        bufs = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(bufs, t)
        if self.rank != 0:
            return None
        return np.concatenate([b.cpu().numpy() for b in bufs]).view(arr.dtype)


def digest(a):
    return hashlib.sha1(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()[:16]


def time_kernels(R, batch, fn, steps, warmup):
    """CUDA-event time of `steps` calls of fn on the batch's launching stream (a real stream handed to the library),
    barrier + synchronize on both sides, max over ranks -> ms per step."""
    torch = R.torch
    stream = torch.cuda.Stream(device=R.dev)
    torch.cuda.set_stream(stream)
    batch.set_stream(stream.cuda_stream)
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    R.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        fn()
    ev1.record(stream)
    R.barrier()
    ms = R.max(ev0.elapsed_time(ev1)) / steps
    batch.set_stream(None)
    torch.cuda.set_stream(torch.cuda.default_stream(R.dev))
    return ms


def kernel_table(prof):
    tot = max(sum(x["ms"] for x in prof.values()), 1e-9)
    return {k: {"launches": v["launches"], "ms_per_launch": v["ms"] / max(v["launches"], 1), "share": v["ms"] / tot}
            for k, v in prof.items()}


# ---------------------------------------------------------------------------------------------------------
def section_config2(args, R, gpu, refs, pins):
    """Headline: config 2, weak scaling (every rank classifies its own `--reads` reads of the global stream)."""
    from rkmh_b200.api import HIT_DTYPE
    n = args.reads
    pin_bases, pin_offs, pin_hits = pins
    hb = pin_bases.view(np.uint8, n * READ_LEN)
    ho = pin_offs.view(np.int64, n + 1)
    config2_reads(refs, R.rank * n, n, out=hb)
    ho[:] = np.arange(n + 1, dtype=np.int64) * READ_LEN
    hits_host = pin_hits.view(HIT_DTYPE, n)

    # ---- kernel-only: batch resident in HBM, CUDA events on the launching stream ----------------------------
    batch = gpu.upload(hb, ho)
    gpu.profile_enable(False)

    def step():
        batch.ingest(); batch.classify()
    for _ in range(args.warmup):
        step()
    gpu.synchronize()
    gpu.profile_enable(True)
    gpu.profile_read(reset=True)
    launches0 = gpu.launch_count()
    kernel_ms = time_kernels(R, batch, step, args.steps, 0)
    launches = gpu.launch_count() - launches0
    prof = gpu.profile_read(reset=True)
    gpu.profile_enable(False)
    dev_hits = batch.results().copy()
    batch.free()

    # ---- end to end: pinned host buffers through the C ABI (H2D + kernels + D2H per step) ---------------------
    def e2e_offsets():
        gpu.classify_raw(pin_bases.ptr, pin_offs.ptr, n, -1, pin_hits.ptr)

    def e2e_fixed():
        gpu.classify_fixed_raw(pin_bases.ptr, READ_LEN, n, -1, pin_hits.ptr)

    def copies_only():
        gpu.copy_probe_raw(pin_bases.ptr, None, READ_LEN, n, pin_hits.ptr)

    w = max(1, min(args.warmup, 2))
    e2e_offs_ms = R.timed(e2e_offsets, args.steps, w)
    same_offs = bool(np.array_equal(hits_host.view(np.int32), dev_hits.view(np.int32)))
    hits_host[:] = 0
    e2e_ms = R.timed(e2e_fixed, args.steps, w)
    same = bool(np.array_equal(hits_host.view(np.int32), dev_hits.view(np.int32)))
    ceiling_ms = R.timed(copies_only, args.steps, 1)
    # the same call with the packed transfer: the library packs every chunk to 2-bit codes + mask on this rank's host
    # cores (inside the timed call) and ships 0.375 B/base; the records must be the same
    threads = max(1, host_cores() // max(R.world, 1))
    gpu.set_host_pack(True, threads)
    hits_host[:] = 0
    packed_ms = R.timed(e2e_fixed, args.steps, w)
    same_packed = bool(np.array_equal(hits_host.view(np.int32), dev_hits.view(np.int32)))
    # adaptive: a chunk is packed only while the link is busy with earlier chunks (link and cores both kept busy)
    gpu.set_host_pack(2, threads)
    hits_host[:] = 0
    for _ in range(max(w, 2)):
        e2e_fixed()
    gpu.transfer_stats(reset=True)
    adaptive_ms = R.timed(e2e_fixed, args.steps, 0)
    xfer = gpu.transfer_stats(reset=True)
    same_adaptive = bool(np.array_equal(hits_host.view(np.int32), dev_hits.view(np.int32)))
    gpu.set_host_pack(False)
    return {"adaptive_ms": adaptive_ms, "same_adaptive": same_adaptive,
            "adaptive_h2d": (xfer["plain_bytes"] + xfer["packed_bytes"]) // max(args.steps, 1),
            "adaptive_packed_share": xfer["packed_chunks"] / max(xfer["packed_chunks"] + xfer["plain_chunks"], 1), "n": n, "hb": hb, "ho": ho, "kernel_ms": kernel_ms, "launches": launches, "prof": prof, "dev_hits": dev_hits,
            "e2e_ms": e2e_ms, "e2e_offs_ms": e2e_offs_ms, "same": same and same_offs, "ceiling_ms": ceiling_ms,
            "packed_ms": packed_ms, "same_packed": same_packed, "pack_threads": threads}


def section_strong(args, R, gpu, c2, pin_hits):
    """Config 2 as BASELINE states it: 10 M reads IN TOTAL, split over the ranks (strong scaling)."""
    total = args.reads
    lo, hi = total * R.rank // R.world, total * (R.rank + 1) // R.world
    n = hi - lo                                     # this rank's slice: the first n reads of its pinned batch
    hb, ho = c2["hb"][: n * READ_LEN], c2["ho"][: n + 1]
    batch = gpu.upload(hb, ho)
    kernel_ms = time_kernels(R, batch, lambda: (batch.ingest(), batch.classify()), args.steps, args.warmup)
    batch.free()
    hits = pin_hits.view(c2["dev_hits"].dtype, max(n, 1))       # pinned, like the bases: both copies run at link speed
    hits[:] = 0
    e2e_ms = R.timed(lambda: gpu.classify_fixed_raw(hb.ctypes.data, READ_LEN, n, -1, pin_hits.ptr), args.steps, 2)
    ok = R.all_true(bool(np.array_equal(hits[:n].view(np.int32), c2["dev_hits"][:n].view(np.int32))))
    # the same slice with the adaptive packed transfer; the faster of the two is the section's end-to-end figure
    e2e_plain_ms = e2e_ms
    gpu.set_host_pack(2, c2["pack_threads"])
    hits[:] = 0
    e2e_adaptive_ms = R.timed(lambda: gpu.classify_fixed_raw(hb.ctypes.data, READ_LEN, n, -1, pin_hits.ptr), args.steps, 3)
    ok_adaptive = R.all_true(bool(np.array_equal(hits[:n].view(np.int32), c2["dev_hits"][:n].view(np.int32))))
    gpu.set_host_pack(0)
    if ok_adaptive and e2e_adaptive_ms < e2e_ms:
        e2e_ms = e2e_adaptive_ms
    return {"total_reads": total, "reads_per_gpu": n, "kernel_ms_per_step": kernel_ms, "e2e_ms_per_step": e2e_ms,
            "e2e_plain_ms_per_step": e2e_plain_ms, "e2e_adaptive_ms_per_step": e2e_adaptive_ms,
            "value": total / (kernel_ms / 1e3), "e2e_value": total / (e2e_ms / 1e3), "unit": "reads/s", "scaling": "strong",
            "matches_weak_run": ok and ok_adaptive}


def section_murmur(args, R, refs, c2, ncu, clocks_mhz, sm_count):
    """Config 2 with the memoised path off: the kernel north_star describes -- every window is hashed (2 x Murmur)."""
    import rkmh_b200
    n = min(c2["n"], 4_000_000)
    with rkmh_b200.RkmhGpu(R.local) as g:
        g.set_memo(False)
        g.set_params([K], S)
        g.load_refs(*rkmh_b200.pack_sequences(refs))
        batch = g.upload(c2["hb"][: n * READ_LEN], c2["ho"][: n + 1])
        g.profile_enable(True)
        ms = time_kernels(R, batch, lambda: (batch.ingest(), batch.classify()), args.steps, args.warmup)
        prof = g.profile_read(reset=True)
        g.profile_enable(False)
        hits = batch.results().copy()
        batch.free()
    same = R.all_true(bool(np.array_equal(hits.view(np.int32), c2["dev_hits"][:n].view(np.int32))))
    kt = kernel_table(prof)
    dom = max(kt.items(), key=lambda kv: kv[1]["share"])[0] if kt else None
    out = {"reads_per_gpu": n, "ms_per_step": ms, "kmers_hashed_per_s": R.world * n * (READ_LEN - K) / (ms / 1e3),
           "reads_per_s": R.world * n / (ms / 1e3), "kernels": kt, "equals_memoised_path": same,
           "murmur_per_kmer": 2, "int_pipe_frac": None}
    cap = ncu.get(dom) if dom else None
    if cap and cap.get("reads") and clocks_mhz:
        # issue-slot occupancy of the hashing kernel from the instruction count of the capture and the LIVE launch time
        wi = cap["warp_instructions"] / cap["reads"]
        dom_ms = kt[dom]["ms_per_launch"]
        out["int_pipe_frac"] = wi * n / (dom_ms / 1e3) / (sm_count * 4 * clocks_mhz * 1e6)
        out["thread_instructions_per_kmer"] = 32 * wi / (READ_LEN - K)
        out["alu_pipe_pct_ncu"] = cap.get("alu_pct")
        out["fma_pipe_pct_ncu"] = cap.get("fma_pct")
    return out


def section_cfg3(args, R):
    """Config 3: `classify -k 12 -s 1000 -M m -N 10` on 10 kb ONT-like reads of the 10 HPV16 lineages.  Each rank owns
    `--cfg3-reads` reads of the global stream; one step = reset -> count (pass 1) -> all-reduce -> classify (pass 2)."""
    import rkmh_b200
    from rkmh_b200.synth import fast_ont_reads
    torch, dist = R.torch, R.dist
    _, refs = load_seqs("hpv_16_allFasta.fa.gz")
    n, L = args.cfg3_reads, 10_000
    total = n * R.world
    m = max(1, round(100 * total / 2_000_000))           # -M 100 at the 2 M reads of config 3, scaled to the sample
    bases, offs = fast_ont_reads(SEED3, refs, R.rank * n, n, L)
    out = {"reads_per_gpu": n, "read_len": L, "k": K, "s": S, "refs": len(refs), "M": m, "M_note": "100 x total reads / 2 M"}
    with rkmh_b200.RkmhGpu(R.local) as g:
        g.set_params([K], S, read_table_size=rkmh_b200.STREAM_TABLE_SIZE)
        g.load_refs(*rkmh_b200.pack_sequences(refs))
        batch = g.upload(bases, offs)
        phase = {"count": 0.0, "allreduce": 0.0, "classify": 0.0}

        def sync_clock():
            torch.cuda.synchronize()
            return time.perf_counter()

        def step(compact, record):
            g.count_reset()
            t0 = sync_clock()
            batch.count()
            t1 = sync_clock()
            table = g.count_partial_tensor() if compact else g.count_table_tensor()
            if dist is not None:
                dist.all_reduce(table, op=dist.ReduceOp.SUM)
            t2 = sync_clock()
            batch.classify(min_occ=m)                     # folds the summed per-code array, bakes the mask, classifies
            t3 = sync_clock()
            if record:
                phase["count"] += t1 - t0; phase["allreduce"] += t2 - t1; phase["classify"] += t3 - t2
            return table.numel() * 4

        steps, warm = max(args.steps, 2), 1
        res = {}
        # the per-code exchange exists on the memoised path (single k <= 12); the 200 M-slot table always does
        for compact in ((True, False) if g.memo_active() else (False,)):
            for k in phase:
                phase[k] = 0.0
            nbytes = 0
            for _ in range(warm):
                step(compact, False)
            R.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                nbytes = step(compact, True)
            torch.cuda.synchronize()
            step_ms = R.max(1e3 * (time.perf_counter() - t0)) / steps
            R.barrier()
            ph = {k: R.max(1e3 * v) / steps for k, v in phase.items()}
            busbw = (2.0 * (R.world - 1) / R.world) * nbytes / (ph["allreduce"] / 1e3) / 1e9 if R.world > 1 else None
            res["compact" if compact else "full_table"] = {
                "allreduce_bytes": nbytes, "step_ms": step_ms, "count_ms": ph["count"], "allreduce_ms": ph["allreduce"],
                "classify_ms": ph["classify"], "allreduce_busbw_gbs": busbw}
            hits = batch.results().copy()
            res["compact" if compact else "full_table"]["hits_digest"] = digest(hits)
        # profile of one more step for the per-kernel split
        g.profile_enable(True); g.profile_read(reset=True)
        step(g.memo_active(), False)
        g.synchronize()
        out["kernels"] = kernel_table(g.profile_read(reset=True))
        g.profile_enable(False)
        hits = batch.results().copy()
        batch.free()
        same_exchange = res.get("compact", res["full_table"])["hits_digest"] == res["full_table"]["hits_digest"]
        # ---- parity against ONE rank: rank 0 runs the reads of every rank through both passes alone -----------------
        all_hits = R.gather_to_rank0(hits)
        all_bases = R.gather_to_rank0(bases) if R.world > 1 else None
        parity = None
        if R.world > 1 and R.rank == 0:
            aoffs = np.arange(total + 1, dtype=np.int64) * L
            g.count_reset()
            g.count_batch(all_bases, aoffs)
            one = g.classify_batch(all_bases, aoffs, min_occ=m)
            parity = bool(np.array_equal(one.view(np.int32), all_hits.view(np.int32)))
        if R.world > 1:
            parity = R.all_true(parity if R.rank == 0 else True)
    best = res.get("compact", res["full_table"])
    nk = total * (L - K)
    out.update({"total_reads": total, "step_ms": best["step_ms"], "reads_per_s": total / (best["step_ms"] / 1e3),
                "kmers_per_s_count": nk / (best["count_ms"] / 1e3), "kmers_per_s_classify": nk / (best["classify_ms"] / 1e3),
                "exchange": res, "compact_equals_full_table": bool(R.all_true(same_exchange)),
                "parity_vs_n1": parity, "mean_best": float(hits["best"].mean()),
                "mean_sketch_len": float(hits["sketch_len"].mean())})
    return out


def section_cfg4(args, R):
    """Config 4: `filter -k 20 -s 2000` against zika / dengue / yellow fever / HPV16 on 150 bp reads, 5 % viral and 95 %
    human-like background (order-0, 41 % GC).  k = 20 is beyond the memoised path: every k-mer is hashed (Murmur path).
    Each rank owns `--cfg4-reads` reads; kernel-only and end to end (pinned host buffers, fixed-length entry point)."""
    import rkmh_b200
    from rkmh_b200.api import HIT_DTYPE, PinnedBuffer
    from rkmh_b200.synth import fast_background, fast_sample_reads
    _, refs = load_seqs("zika.fa.gz", "dengue.fa.gz", "yellow_fever.fa.gz", "hpv_16.fa.gz")
    n, L, k4, s4, N4 = args.cfg4_reads, 150, 20, 2000, 10
    first = R.rank * n
    out = {"reads_per_gpu": n, "read_len": L, "k": k4, "s": s4, "refs": len(refs), "N": N4}
    with rkmh_b200.RkmhGpu(R.local) as g:
        g.set_params([k4], s4)
        g.load_refs(*rkmh_b200.pack_sequences(refs), per_sample_set=True)
        pin_b, pin_h = PinnedBuffer(g, n * L), PinnedBuffer(g, HIT_DTYPE.itemsize * n)
        hb = pin_b.view(np.uint8, n * L)
        fast_background(4, first * L, n * L, out=hb)
        nv = (n + 19) // 20                             # every 20th read of the stream is viral (5 %)
        vb, _ = fast_sample_reads(4, refs, first // 20, nv, L, random_frac=0.0)
        hb.reshape(n, L)[::20] = vb.reshape(nv, L)[: len(range(0, n, 20))]
        offs = np.arange(n + 1, dtype=np.int64) * L
        batch = g.upload(hb, offs)
        g.profile_enable(True)
        ms = time_kernels(R, batch, lambda: (batch.ingest(), batch.classify()), args.steps, args.warmup)
        out["kernels"] = kernel_table(g.profile_read(reset=True))
        g.profile_enable(False)
        dev_hits = batch.results().copy()
        batch.free()
        hits = pin_h.view(HIT_DTYPE, n)
        e2e_ms = R.timed(lambda: g.classify_fixed_raw(pin_b.ptr, L, n, -1, pin_h.ptr), args.steps, 2)
        same = R.all_true(bool(np.array_equal(hits.view(np.int32), dev_hits.view(np.int32))))
        # the same call with the adaptive packed transfer (on the Murmur path a packed chunk also skips the ingest kernel)
        e2e_plain_ms = e2e_ms
        g.set_host_pack(2, max(1, host_cores() // max(R.world, 1)))
        hits[:] = 0
        e2e_adaptive_ms = R.timed(lambda: g.classify_fixed_raw(pin_b.ptr, L, n, -1, pin_h.ptr), args.steps, 3)
        same_adaptive = R.all_true(bool(np.array_equal(hits.view(np.int32), dev_hits.view(np.int32))))
        g.set_host_pack(0)
        if same_adaptive and e2e_adaptive_ms < e2e_ms:
            e2e_ms = e2e_adaptive_ms
        out["e2e_transfers"] = {"plain_ms": e2e_plain_ms, "adaptive_ms": e2e_adaptive_ms, "adaptive_matches_device_run": same_adaptive}
        # rkmh filter keeps a read iff sketch_len > 0, best >= N and best - prev > D (src/rkmh.cpp:1292-1298)
        best, prev = np.maximum(dev_hits["best"], 0), np.maximum(dev_hits["prev"], 0)
        keep = (dev_hits["sketch_len"] > 0) & (best >= N4) & (best - prev > 0)
        viral = np.zeros(n, bool); viral[::20] = True
        pin_b.free(); pin_h.free()
    total = n * R.world
    out.update({"total_reads": total, "ms_per_step": ms, "reads_per_s": total / (ms / 1e3),
                "kmers_hashed_per_s": total * (L - k4) / (ms / 1e3), "e2e_ms_per_step": e2e_ms,
                "e2e_reads_per_s": total / (e2e_ms / 1e3), "e2e_matches_device_run": same,
                "pass_fraction": float(keep.mean()), "viral_recall": float(keep[viral].mean()),
                "background_pass": float(keep[~viral].mean())})
    return out


def section_cfg5(args, R):
    """Config 5: `hash -k 21 -s 10000` of ONE record whose k-mer starts are sharded over the ranks (each rank holds
    `--cfg5-bases` of it): rkg_sketch_shard -> all-gather of the bottom-s multisets -> rkg_merge_sketches."""
    import rkmh_b200
    from rkmh_b200.api import PinnedBuffer
    from rkmh_b200.dist import allgather_sketches
    from rkmh_b200.synth import fast_genome
    k5, s5 = 21, 10_000
    per = args.cfg5_bases
    L = per * R.world                                   # the whole record
    nk = L - k5                                          # k-mer starts [0, L - k) exist
    b, e = nk * R.rank // R.world, nk * (R.rank + 1) // R.world
    lo, hi = b, min(e + k5, L)
    out = {"record_bases": L, "bases_per_gpu": per, "k": k5, "s": s5}
    with rkmh_b200.RkmhGpu(R.local) as g:
        g.set_params([k5], s5)
        pin = PinnedBuffer(g, hi - lo)
        shard = fast_genome(SEED5, lo, hi - lo, out=pin.view(np.uint8, hi - lo))
        phase = {"shard": 0.0, "gather_merge": 0.0}
        result = {}

        def step(record=True):
            t0 = time.perf_counter()
            part = g.sketch_shard(shard, b, e, L)         # synchronous: H2D in pieces overlapped with ingest + one hashing pass, sort
            t1 = time.perf_counter()
            parts, lens = allgather_sketches(part, s5, device=R.dev if R.world > 1 else None)
            merged = g.merge_sketches(parts, lens) if R.world > 1 else part
            R.torch.cuda.synchronize()
            t2 = time.perf_counter()
            if record:
                phase["shard"] += t1 - t0; phase["gather_merge"] += t2 - t1
            result["merged"] = merged

        step(False)
        g.profile_enable(True); g.profile_read(reset=True)
        steps = max(args.steps, 2)
        ms = R.timed(step, steps, 0)
        out["kernels"] = kernel_table(g.profile_read(reset=True))
        g.profile_enable(False)
        merged = result["merged"]
        shard_plain_ms = R.max(1e3 * phase["shard"]) / steps
        gather_plain = phase["gather_merge"]
        # the same with the adaptive packed transfer (pieces packed on the host while the link is busy); the faster one is
        # the section's figure, both are reported
        threads = max(1, host_cores() // max(R.world, 1))
        g.set_host_pack(2, threads)
        step(False); step(False)
        g.transfer_stats(reset=True)
        phase["shard"] = 0.0; phase["gather_merge"] = 0.0
        ms_adaptive = R.timed(step, steps, 0)
        xfer = g.transfer_stats(reset=True)
        same_adaptive = R.all_true(bool(np.array_equal(result["merged"], merged)))
        g.set_host_pack(0)
        out["plain_transfer"] = {"step_ms": ms, "shard_ms": shard_plain_ms}
        out["adaptive_transfer"] = {"step_ms": ms_adaptive, "shard_ms": R.max(1e3 * phase["shard"]) / steps, "same_sketch": same_adaptive,
                                    "packed_share_of_pieces": xfer["packed_chunks"] / max(xfer["packed_chunks"] + xfer["plain_chunks"], 1),
                                    "host_threads_per_rank": threads}
        if same_adaptive and ms_adaptive < ms:
            ms = ms_adaptive
        else:
            phase["shard"] = shard_plain_ms * steps / 1e3
            phase["gather_merge"] = gather_plain
        parity = None
        if R.world > 1:
            # ONE rank, whole record: rank 0 regenerates every shard's bases (a pure function of the seed and the
            # position) and sketches the record with a single rkg_sketch_shard call
            ok = True
            if R.rank == 0:
                whole = fast_genome(SEED5, 0, L)
                one = g.sketch_shard(whole, 0, nk, L)
                ok = bool(np.array_equal(one, merged))
                del whole
            parity = R.all_true(ok)
        pin.free()
    out.update({"step_ms": ms, "shard_ms": R.max(1e3 * phase["shard"]) / steps,
                "allgather_merge_ms": R.max(1e3 * phase["gather_merge"]) / steps,
                "kmers_per_s": nk / (ms / 1e3), "sketch_len": int(len(merged)), "sketch_digest": digest(merged),
                "parity_vs_n1": parity, "h2d_bytes_per_step_per_gpu": int(hi - lo),
                "allgather_bytes": int(R.world * (s5 + 1) * 8)})
    return out


def guarded(name, fn, R):
    """A side section must never take the headline line down with it; all ranks agree on whether it ran."""
    try:
        return fn()
    except Exception as exc:                              # noqa: BLE001
        sys.stderr.write(f"bench.py: section {name} failed on rank {R.rank}: {exc!r}\n")
        return {"error": repr(exc)}


def main_ours(args, rank, world, local_rank):
    import rkmh_b200
    from rkmh_b200.api import HIT_DTYPE, PinnedBuffer

    R = Ranks(rank, world, local_rank)
    torch = R.torch
    # each rank sits on its GPU's NUMA node before it allocates pinned memory (matters from 4 ranks up: the reads of
    # one step are 1.6 GB per GPU, and crossing sockets halves the host-to-device rate)
    numa_cpus = rkmh_b200.bind_host_thread_to_gpu(local_rank) if world > 1 else None
    ref_names, refs = load_refs()
    gpu = rkmh_b200.RkmhGpu(local_rank)
    gpu.set_params([K], S)
    gpu.load_refs(*rkmh_b200.pack_sequences(refs))

    n = args.reads
    pins = (PinnedBuffer(gpu, n * READ_LEN), PinnedBuffer(gpu, 8 * (n + 1)), PinnedBuffer(gpu, HIT_DTYPE.itemsize * n))
    # clocks are sampled from before the warm-up to the end of the timed sections (the timed regions are tens of
    # milliseconds; nvidia-smi cannot sample inside them alone)
    sampler = ClockSampler(local_rank)
    sampler.start()
    c2 = section_config2(args, R, gpu, refs, pins)
    clocks = sampler.stop()
    try:
        sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    except Exception:                                     # noqa: BLE001
        sm_count = 148
    ncu, ncu_note = ncu_current()
    clk = float(clocks.get("sm_mhz") or 1965.0)

    extra = {}
    if not args.only_headline:
        extra["strong_scaling"] = guarded("strong_scaling", lambda: section_strong(args, R, gpu, c2, pins[2]), R)
        extra["murmur_path"] = guarded("murmur_path", lambda: section_murmur(args, R, refs, c2, ncu, clk, sm_count), R)
    spot = None
    if rank == 0:
        spot = guarded("parity_spot", lambda: parity_spot(c2["dev_hits"], c2["hb"], c2["ho"], refs), R)
    hb_sample = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        m = min(args.cpu_sample, n)
        hb_sample = (c2["hb"][: m * READ_LEN].copy(), c2["ho"][: m + 1].copy())
    for p in pins:
        p.free()
    gpu.close()
    configs = {}
    if not args.only_headline:
        if args.cfg3_reads > 0:
            configs["cfg3"] = guarded("cfg3", lambda: section_cfg3(args, R), R)
        if args.cfg4_reads > 0:
            configs["cfg4"] = guarded("cfg4", lambda: section_cfg4(args, R), R)
        if args.cfg5_bases > 0:
            configs["cfg5"] = guarded("cfg5", lambda: section_cfg5(args, R), R)

    if rank == 0:
        total_reads = n * world
        kernel_ms = c2["kernel_ms"]
        # the end-to-end figure is the faster of the two transfers the public call offers (plain bytes, or packed on the
        # host inside the call); both are reported below
        plain_ms = c2["e2e_ms"]
        use_packed = c2["same_packed"] and c2["packed_ms"] < plain_ms
        e2e_ms = c2["packed_ms"] if use_packed else plain_ms
        use_adaptive = c2["same_adaptive"] and c2["adaptive_ms"] < e2e_ms
        if use_adaptive:
            e2e_ms = c2["adaptive_ms"]
        value = total_reads / (kernel_ms / 1e3)
        e2e_value = total_reads / (e2e_ms / 1e3)
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            with open(pk) as fh:
                peaks = json.load(fh)
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        prof = c2["prof"]
        # dominant kernel = the one with the largest share of the timed region
        dom = max(prof.items(), key=lambda kv: kv[1]["ms"]) if prof else ("none", {"ms": 0.0, "launches": 1})
        dom_ms = dom[1]["ms"] / max(dom[1]["launches"], 1)
        # ALGORITHMIC bytes per read of the dominant kernel.  Memoised kernels (single k <= 12)
        # read the ASCII bases as parsed -- ingest is fused -- : 150 B bases + 8 B offset + 16 B hit record.  The Murmur
        # kernels read the packed batch: 2-bit codes + 1-bit mask per base + offset + hit record.
        memo_path = dom[0].startswith("classify_memo")
        bytes_per_read = (READ_LEN + 8 + 16) if memo_path else (READ_LEN * 3 / 8 + 8 + 16)
        achieved = n * bytes_per_read / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else 0.0
        cap = ncu.get(dom[0])
        traffic = cap["dram_bytes"] / cap["reads"] * n if cap and cap.get("reads") else None
        issue_slots = None
        if cap and cap.get("reads") and dom_ms > 0:
            wi = cap["warp_instructions"] / cap["reads"]
            got = wi * n / (dom_ms / 1e3)
            peak = sm_count * 4 * clk * 1e6
            issue_slots = {"kernel": dom[0], "warp_instructions_per_read": wi, "thread_instructions_per_window": 32 * wi / (READ_LEN - K),
                           "achieved": got / 1e12, "peak": peak / 1e12, "unit": "T warp-instructions/s", "frac": got / peak,
                           "alu_pipe_pct_ncu": cap.get("alu_pct"), "issue_active_pct_ncu": cap.get("issue_pct"),
                           "source": "instruction count from " + ncu_note + "; launch time and SM clock live"}
        if memo_path:
            note = ("memoised path: every window is LOOKED UP, not hashed (shared-memory sieve, then L2-resident code map "
                    "and bitset rows); the kernel is bound by instruction issue, not by HBM -- see issue_slots and "
                    "DESIGN.md section 5")
        else:
            note = ("the kernel is INT-pipe bound (2 x MurmurHash3_x64_128 per k-mer), not HBM bound; see DESIGN.md "
                    "section 5")
        h2d = int(n * READ_LEN)
        line = {
            "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic (std::mt19937_64, seed 2; rkmh_b200/host/synth.cpp)",
            "config": bench_config(args, world, len(refs)),
            "host": {"cores": host_cores(), "numa_binding": (f"{len(numa_cpus)} local cores per rank" if numa_cpus else "none")},
            "kmers_per_s": value * (READ_LEN - K),
            "kmers_note": "windows classified per second; on the memoised path they are looked up, murmur_path hashes them",
            "e2e": {"value": e2e_value, "unit": "reads/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": (int(c2["adaptive_h2d"]) if use_adaptive else int(12 * ((n * READ_LEN + 31) // 32)) if use_packed else h2d),
                    "d2h_bytes_per_step": int(16 * n),
                    "api": ("rkg_classify_fixed after rkg_set_host_pack(2): pinned host buffers of ASCII bases in; the library packs a chunk "
                            "to 2-bit codes + mask on the host cores INSIDE the timed call while the link is busy with earlier chunks and "
                            "sends it as it is otherwise (h2d bytes are this rank's measured mix)"
                            if use_adaptive else
                            "rkg_classify_fixed after rkg_set_host_pack: pinned host buffers of ASCII bases in, packed to 2-bit codes + "
                            "mask by the library on the host cores INSIDE the timed call, 0.375 B/base over the link"
                            if use_packed else
                            "rkg_classify_fixed (pinned host buffers, fixed-length reads: no offsets cross the link)"),
                    "matches_device_run": c2["same"] and (c2["same_packed"] or not use_packed) and (c2["same_adaptive"] or not use_adaptive),
                    "adaptive_transfer": {"ms_per_step": c2["adaptive_ms"], "value": total_reads / (c2["adaptive_ms"] / 1e3),
                                          "h2d_bytes_per_step": int(c2["adaptive_h2d"]), "packed_share_of_chunks": c2["adaptive_packed_share"],
                                          "host_threads_per_rank": c2["pack_threads"], "matches_device_run": c2["same_adaptive"]},
                    "plain_transfer": {"ms_per_step": plain_ms, "value": total_reads / (plain_ms / 1e3), "h2d_bytes_per_step": h2d},
                    "packed_transfer": {"ms_per_step": c2["packed_ms"], "value": total_reads / (c2["packed_ms"] / 1e3),
                                        "h2d_bytes_per_step": int(12 * ((n * READ_LEN + 31) // 32)), "host_threads_per_rank": c2["pack_threads"],
                                        "matches_device_run": c2["same_packed"]},
                    "with_offsets": {"api": "rkg_classify_batch", "ms_per_step": c2["e2e_offs_ms"],
                                     "value": total_reads / (c2["e2e_offs_ms"] / 1e3), "h2d_bytes_per_step": h2d + 8 * (n + 1)},
                    "copy_ceiling_ms": c2["ceiling_ms"], "frac_of_copy_ceiling": c2["ceiling_ms"] / plain_ms,
                    "copy_ceiling_gbs_per_gpu": (h2d + 16 * n) / (c2["ceiling_ms"] / 1e3) / 1e9,
                    "copy_ceiling_note": "rkg_copy_probe: the chunked H2D + D2H copies of the PLAIN transfer on the same streams "
                                         "without the kernels, all ranks at once, max over ranks; frac = ceiling / plain e2e"},
            "gpu_launches": int(c2["launches"]),
            "clocks": clocks,
            "kernels": kernel_table(prof),
            "roofline": {"kernel": dom[0], "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": ncu_note, "peak_source": peak_src,
                         "algorithmic_bytes_per_read": bytes_per_read, "note": note},
            "issue_slots": issue_slots,
            "dominant_kernel": {"kernel": dom[0], "ms_per_launch": dom_ms,
                                "windows_per_s_per_gpu": (n * (READ_LEN - K)) / (dom_ms / 1e3) if dom_ms > 0 else 0.0},
            "parity_spot": spot,
        }
        line.update(extra)
        if configs:
            line["configs"] = configs
            c3, c5 = configs.get("cfg3", {}), configs.get("cfg5", {})
            ex = c3.get("exchange", {})
            line["collectives"] = {
                "allreduce_ms": ex.get("compact", {}).get("allreduce_ms"),
                "allreduce_bytes": ex.get("compact", {}).get("allreduce_bytes"),
                "allreduce_busbw_gbs": ex.get("compact", {}).get("allreduce_busbw_gbs"),
                "allreduce_full_table_ms": ex.get("full_table", {}).get("allreduce_ms"),
                "allreduce_full_table_bytes": ex.get("full_table", {}).get("allreduce_bytes"),
                "allreduce_full_table_busbw_gbs": ex.get("full_table", {}).get("allreduce_busbw_gbs"),
                "allgather_merge_ms": c5.get("allgather_merge_ms"),
                "parity_vs_n1": {"cfg3": c3.get("parity_vs_n1"), "cfg5": c5.get("parity_vs_n1")},
                "note": "N = 1: no exchange takes place (times are the host-side no-op), parity_vs_n1 is null"}
        if hb_sample is not None:
            workdir = tempfile.mkdtemp(prefix="rkmh_cpu_")
            try:
                line["cpu_baseline"] = cpu_baseline(hb_sample[0], hb_sample[1], len(hb_sample[1]) - 1, workdir)
            finally:
                shutil.rmtree(workdir, ignore_errors=True)
        print(json.dumps(line), flush=True)

    if R.dist is not None:
        R.dist.barrier()
        R.dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        main_reference(args, rank, world)
    else:
        main_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
