#!/usr/bin/env python
"""Wall-clock throughput of the C++ `rkmh stream` command line on a synthetic config-2 FASTQ (parsing included),
next to the reference binary on the same file when oracle/_ref/rkmh_ref exists.  A tuning aid, not the bench contract.

    python tools/cli_throughput.py [--reads 2000000] [--gpus 1]
"""
import argparse
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DATA = os.path.join(ROOT, "tests", "golden", "data")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=2_000_000)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--ref-reads", type=int, default=200_000, help="reads given to the reference binary")
    args = ap.parse_args()
    from rkmh_b200.fastx import load_sequences
    from rkmh_b200.synth import sample_reads, write_fastq
    _, refs, _ = load_sequences([os.path.join(DATA, "all_pave_ref.fa.gz")])
    bases, offs = sample_reads(np.random.default_rng(2), refs, args.reads, 150)
    work = tempfile.mkdtemp(prefix="rkmh_cli_")
    try:
        ref_fa = os.path.join(work, "refs.fa")
        with gzip.open(os.path.join(DATA, "all_pave_ref.fa.gz"), "rb") as fi, open(ref_fa, "wb") as fo:
            shutil.copyfileobj(fi, fo)
        fq = os.path.join(work, "reads.fq")
        write_fastq(fq, bases, offs)
        small = os.path.join(work, "small.fq")
        write_fastq(small, bases, offs[:args.ref_reads + 1])
        exe = os.path.join(ROOT, "rkmh_b200", "bin", "rkmh")
        out = {"reads": args.reads, "fastq_bytes": os.path.getsize(fq), "gpus": args.gpus}

        def run(cmd, path):
            t0 = time.perf_counter()
            with open(os.path.join(work, "out.txt"), "wb") as fo:
                subprocess.run(cmd + ["-f", path], stdout=fo, stderr=subprocess.DEVNULL, check=True)
            return time.perf_counter() - t0

        ours = [exe, "stream", "-r", ref_fa, "-k", "12", "-s", "1000", "-g", str(args.gpus)]
        run(ours, small)                                   # warm the page cache and the driver
        t_small, t_full = run(ours, small), run(ours, fq)
        out["ours_seconds"] = t_full
        out["ours_reads_per_s_wall"] = args.reads / t_full
        out["ours_reads_per_s_marginal"] = (args.reads - args.ref_reads) / max(t_full - t_small, 1e-9)
        md5_ours = subprocess.run(["md5sum", os.path.join(work, "out.txt")], capture_output=True, text=True).stdout.split()[0]
        run(ours, small)
        with open(os.path.join(work, "out.txt"), "rb") as fh:
            ours_small = fh.read().splitlines()
        ref_exe = os.path.join(ROOT, "oracle", "_ref", "rkmh_ref")
        if os.path.exists(ref_exe):
            cores = len(os.sched_getaffinity(0))
            t_ref = run([ref_exe, "stream", "-r", ref_fa, "-k", "12", "-s", "1000", "-t", str(cores)], small)
            with open(os.path.join(work, "out.txt"), "rb") as fh:
                ref_small = fh.read().splitlines()
            out["reference_reads"] = args.ref_reads
            out["reference_seconds"] = t_ref
            out["reference_reads_per_s_wall"] = args.ref_reads / t_ref
            out["reference_cores"] = cores
            # the reference prints from inside its OpenMP loop: compare the lines as a multiset, not in order
            out["lines"] = [len(ours_small), len(ref_small)]
            out["same_lines_as_reference"] = sorted(ours_small) == sorted(ref_small)
            if not out["same_lines_as_reference"]:
                out["first_differences"] = [x.decode() for x in sorted(set(ours_small) ^ set(ref_small))[:6]]
        out["stdout_md5"] = md5_ours
        print(json.dumps(out))
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
