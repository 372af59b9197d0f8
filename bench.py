#!/usr/bin/env python
"""bench.py -- rkmh classify throughput on B200 (BASELINE.json config 2:
`rkmh stream -r all_pave_ref.fa -k 12 -s 1000` on synthetic 150 bp Illumina-like reads, 182 references).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU implementation

One "step" = one pass of the hot path (ingest -> hash -> sketch -> probe -> argmax) over one batch of
`--reads` synthetic reads per GPU.  `value` is timed with the batch resident in HBM; `e2e` is timed through
rkg_classify_batch with pinned HOST buffers (H2D of the bases/offsets and D2H of the hit records inside
the timed region).  Prints ONE JSON line on rank 0.
"""
import argparse
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
# DRAM bytes per read (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of a 2 M-read launch,
# divided by the reads); `roofline.traffic` scales it to the reads of one launch.  Kernels without a capture report null.
# Warp instructions per read of each kernel (smsp__inst_executed.sum of the same captures / reads): with the live launch time
# and SM clock they give the fraction of the chip's issue slots (148 SMs x 4 schedulers x 1 instruction per clock) in use,
# which is the resource these kernels are bound by (the HBM `roofline` above is reported because the contract asks for it).
WARP_INSTR_PER_READ = {"classify_memo_sieve_kernel": 657.6, "classify_memo_warp_kernel": 803.6, "classify_group_kernel": 1805.8}
TRAFFIC_PER_READ = {"classify_memo_sieve_kernel": 188.6,      # profiles/r01_sieve_final_ncu_raw.json (377.2 MB / 2 M reads)
                    "classify_memo_warp_kernel": 1831.4,      # profiles/r01_memo_warp_ncu_raw.json
                    "classify_group_kernel": 91.0}            # profiles/r01_group_kernel_ncu_raw.json


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per step (config 2: 10 M)")
    ap.add_argument("--cpu-sample", type=int, default=1_000_000, help="reads in the CPU baseline sample (about 10 s of the reference on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def load_refs():
    from rkmh_b200.fastx import load_sequences
    names, seqs, _ = load_sequences([os.path.join(DATA, "all_pave_ref.fa.gz")])
    return names, seqs


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


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


# ---------------------------------------------------------------------------------------------------------
def main_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation on the host cores, rank 0 only."""
    if rank != 0:
        return
    from rkmh_b200.synth import sample_reads
    _, refs = load_refs()
    sample = args.cpu_sample
    rng = np.random.default_rng(2)
    bases, offs = sample_reads(rng, refs, sample, READ_LEN)
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
            "config": {"workload": WORKLOAD, "reads_per_gpu_per_step": args.reads, "k": K, "s": S, "refs": len(refs),
                       "parallelism": "host cores of rank 0", "sample_reads_per_step": sample,
                       "note": "each step is a bounded sample of the workload through the reference's own CPU implementation"},
            "kmers_per_s": value * (READ_LEN - K),
            "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": kind,
                             "sample": f"{sample} reads per step, wall clock of the CLI minus a refs-only run"},
            "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main_ours(args, rank, world, local_rank):
    import torch
    import rkmh_b200
    from rkmh_b200.api import HIT_DTYPE, PinnedBuffer
    from rkmh_b200.synth import sample_reads

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        # NCCL may print its version banner on stdout while the communicator is built; stdout carries exactly one JSON
        # line, so anything printed during initialisation goes to stderr instead
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
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")

    # each rank sits on its GPU's NUMA node before it allocates pinned memory (matters from 4 ranks up: the reads of
    # one step are 1.6 GB per GPU, and crossing sockets halves the host-to-device rate)
    numa_cpus = rkmh_b200.bind_host_thread_to_gpu(local_rank) if world > 1 else None
    ref_names, refs = load_refs()
    gpu = rkmh_b200.RkmhGpu(local_rank)
    gpu.set_params([K], S)
    gpu.load_refs(*rkmh_b200.pack_sequences(refs))

    # synthetic reads straight into pinned host memory (seeded per rank: weak scaling, reads sharded by rank)
    n = args.reads
    pin_bases = PinnedBuffer(gpu, n * READ_LEN)
    pin_offs = PinnedBuffer(gpu, 8 * (n + 1))
    pin_hits = PinnedBuffer(gpu, HIT_DTYPE.itemsize * n)
    rng = np.random.default_rng(2 + rank)
    bases, offs = sample_reads(rng, refs, n, READ_LEN)
    pin_bases.view(np.uint8, n * READ_LEN)[:] = bases
    pin_offs.view(np.int64, n + 1)[:] = offs
    hb, ho = pin_bases.view(np.uint8, n * READ_LEN), pin_offs.view(np.int64, n + 1)
    hits_host = pin_hits.view(HIT_DTYPE, n)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # clocks are sampled from before the warm-up to the end of the end-to-end loop (the timed regions are
    # tens of milliseconds; nvidia-smi cannot sample inside them alone)
    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- kernel-only: batch resident in HBM, timed with CUDA events on the launching stream -------------
    batch = gpu.upload(hb, ho)
    stream = torch.cuda.Stream(device=dev)          # a real (non-default) stream: its handle is what the library launches on
    torch.cuda.set_stream(stream)
    batch.set_stream(stream.cuda_stream)
    for _ in range(args.warmup):
        batch.ingest(); batch.classify()
    torch.cuda.synchronize()
    gpu.profile_enable(True)
    gpu.profile_read(reset=True)
    launches0 = gpu.launch_count()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        batch.ingest(); batch.classify()
    ev1.record(stream)
    barrier()
    kernel_ms = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    launches = gpu.launch_count() - launches0
    prof = gpu.profile_read(reset=True)
    gpu.profile_enable(False)
    dev_hits = batch.results().copy()
    batch.set_stream(None)
    batch.free()

    # ---- end to end: pinned host buffers through rkg_classify_batch (H2D + kernels + D2H per step) ----------
    for _ in range(max(1, min(args.warmup, 2))):
        gpu.classify_raw(pin_bases.ptr, pin_offs.ptr, n, -1, pin_hits.ptr)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        gpu.classify_raw(pin_bases.ptr, pin_offs.ptr, n, -1, pin_hits.ptr)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0)) / args.steps
    barrier()
    clocks = sampler.stop()
    same = bool(np.array_equal(hits_host.view(np.int32), dev_hits.view(np.int32)))

    if rank == 0:
        total_reads = n * world
        value = total_reads / (kernel_ms / 1e3)
        e2e_value = total_reads / (e2e_ms / 1e3)
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            with open(pk) as fh:
                peaks = json.load(fh)
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        try:
            sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
        except Exception:
            sm_count = 148
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        # dominant kernel = the one with the largest share of the timed region
        dom = max(prof.items(), key=lambda kv: kv[1]["ms"]) if prof else ("none", {"ms": 0.0, "launches": 1})
        dom_ms = dom[1]["ms"] / max(dom[1]["launches"], 1)
        # ALGORITHMIC bytes per read of the dominant kernel.  Memoised kernels (single k <= 12)
        # read the ASCII bases as parsed -- ingest is fused -- : 150 B bases + 8 B offset + 16 B hit record.  The Murmur
        # group kernel reads the packed batch: 2-bit codes + 1-bit mask per base + offset + hit record.
        memo_path = dom[0].startswith("classify_memo")
        bytes_per_read = (READ_LEN + 8 + 16) if memo_path else (READ_LEN * 3 / 8 + 8 + 16)
        alg_bytes = n * bytes_per_read
        achieved = alg_bytes / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else 0.0
        if memo_path:
            note = ("memoised path: every window is one lookup (shared-memory sieve, then L2-resident code map and bitset "
                    "rows); the kernel is issue/latency bound, not HBM bound -- ncu counters in profiles/, DESIGN.md section 5")
        else:
            note = ("the kernel is INT-pipe bound (2 x MurmurHash3_x64_128 per k-mer), not HBM bound; see DESIGN.md "
                    "section 5")
        issue_slots = None
        try:
            if dom[0] in WARP_INSTR_PER_READ and dom_ms > 0:
                got = WARP_INSTR_PER_READ[dom[0]] * n / (dom_ms / 1e3)
                peak = sm_count * 4 * float(clocks.get("sm_mhz") or 1965.0) * 1e6
                issue_slots = {"kernel": dom[0], "warp_instructions_per_read": WARP_INSTR_PER_READ[dom[0]],
                               "achieved": got / 1e12, "peak": peak / 1e12, "unit": "T warp-instructions/s", "frac": got / peak,
                               "source": "instruction count from ncu (profiles/), launch time and SM clock live"}
        except Exception:                                 # explanatory only: never in the way of the bench line
            issue_slots = None
        line = {
            "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "reads_per_gpu_per_step": n, "k": K, "s": S, "refs": len(refs), "parallelism": f"reads sharded x{world}",
                       "host_numa_binding": (f"{len(numa_cpus)} local cores per rank" if numa_cpus else "none"),
                       "l2": "inputs (1.5 GB of bases per step) are larger than L2; no flush needed"},
            "kmers_per_s": value * (READ_LEN - K),
            "e2e": {"value": e2e_value, "unit": "reads/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(n * READ_LEN + 8 * (n + 1)), "d2h_bytes_per_step": int(16 * n),
                    "api": "rkg_classify_batch (pinned host buffers)", "matches_device_run": same},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "kernels": {k: {"launches": v["launches"], "ms_per_launch": v["ms"] / max(v["launches"], 1),
                            "share": v["ms"] / max(sum(x["ms"] for x in prof.values()), 1e-9)} for k, v in prof.items()},
            "roofline": {"kernel": dom[0], "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": (TRAFFIC_PER_READ[dom[0]] * n if dom[0] in TRAFFIC_PER_READ else None), "peak_source": peak_src,
                         "algorithmic_bytes_per_read": bytes_per_read, "note": note},
            "issue_slots": issue_slots,
            "dominant_kernel": {"kernel": dom[0], "ms_per_launch": dom_ms,
                                "kmers_per_s_per_gpu": (n * (READ_LEN - K)) / (dom_ms / 1e3) if dom_ms > 0 else 0.0},
        }
        if not args.no_cpu_baseline and world == 1:
            workdir = tempfile.mkdtemp(prefix="rkmh_cpu_")
            try:
                line["cpu_baseline"] = cpu_baseline(bases, offs, min(args.cpu_sample, n), workdir)
            finally:
                shutil.rmtree(workdir, ignore_errors=True)
        print(json.dumps(line), flush=True)

    for p in (pin_bases, pin_offs, pin_hits):
        p.free()
    gpu.close()
    if dist is not None:
        dist.destroy_process_group()


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
