#!/usr/bin/env python
"""Run ONE workload a few times so that a profiler can capture its kernels (ncu -k regex:<kernel> -c 1):

    python tools/run_kernel.py sieve  [reads]    config 2, memoised path   (classify_memo_sieve_kernel)
    python tools/run_kernel.py murmur [reads]    config 2, memoised path off (classify_warp_kernel, ingest_kernel)
    python tools/run_kernel.py cfg3   [reads]    config 3: count_memo_kernel, classify_memo_block_kernel
    python tools/run_kernel.py cfg4   [reads]    config 4: classify_warp_kernel at k = 20
Prints the reads per launch as JSON (profiles/ncu_summary.py --current needs it).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rkmh_b200  # noqa: E402
from rkmh_b200 import synth  # noqa: E402
from rkmh_b200.fastx import load_sequences  # noqa: E402

DATA = os.path.join(ROOT, "tests", "golden", "data")


def refs_of(*files):
    out = []
    for f in files:
        out += load_sequences([os.path.join(DATA, f)])[1]
    return out


def main():
    mode = sys.argv[1]
    reps = 3
    with rkmh_b200.RkmhGpu(0) as g:
        if mode in ("sieve", "murmur"):
            n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
            refs = refs_of("all_pave_ref.fa.gz")
            g.set_memo(mode == "sieve")
            g.set_params([12], 1000)
            g.load_refs(*rkmh_b200.pack_sequences(refs))
            b = g.upload(*synth.fast_sample_reads(2, refs, 0, n, 150))
            for _ in range(reps):
                b.ingest(); b.classify()
        elif mode == "cfg3":
            n = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000
            refs = refs_of("hpv_16_allFasta.fa.gz")
            g.set_params([12], 1000, read_table_size=rkmh_b200.STREAM_TABLE_SIZE)
            g.load_refs(*rkmh_b200.pack_sequences(refs))
            b = g.upload(*synth.fast_ont_reads(3, refs, 0, n, 10_000))
            for _ in range(reps):
                g.count_reset(); b.count(); b.classify(min_occ=2)
        elif mode == "cfg4":
            n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
            refs = refs_of("zika.fa.gz", "dengue.fa.gz", "yellow_fever.fa.gz", "hpv_16.fa.gz")
            g.set_params([20], 2000)
            g.load_refs(*rkmh_b200.pack_sequences(refs), per_sample_set=True)
            hb = synth.fast_background(4, 0, n * 150)
            nv = (n + 19) // 20
            vb, _ = synth.fast_sample_reads(4, refs, 0, nv, 150, random_frac=0.0)
            hb.reshape(n, 150)[::20] = vb.reshape(nv, 150)[: len(range(0, n, 20))]
            b = g.upload(hb, np.arange(n + 1, dtype=np.int64) * 150)
            for _ in range(reps):
                b.ingest(); b.classify()
        else:
            raise SystemExit(__doc__)
        g.synchronize()
        b.free()
    print(json.dumps({"mode": mode, "reads_per_launch": n}))


if __name__ == "__main__":
    main()
