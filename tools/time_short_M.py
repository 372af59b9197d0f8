#!/usr/bin/env python
"""Kernel time of `stream -M` on short reads (config-2 reads with a -M mask: the rank-table warp kernel)."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rkmh_b200
from rkmh_b200.fastx import load_sequences
from rkmh_b200.synth import sample_reads
_, refs, _ = load_sequences([os.path.join(ROOT, "tests/golden/data/all_pave_ref.fa.gz")])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
bases, offs = sample_reads(np.random.default_rng(2), refs, n, 150)
with rkmh_b200.RkmhGpu(0) as g:
    g.set_params([12], 1000, read_table_size=rkmh_b200.STREAM_TABLE_SIZE)
    g.load_refs(*rkmh_b200.pack_sequences(refs))
    b = g.upload(bases, offs)
    g.count_reset(); b.count(); b.classify(min_occ=2); g.synchronize()
    g.profile_enable(True); g.profile_read(reset=True)
    for _ in range(3):
        g.count_reset(); b.count(); b.classify(min_occ=2)
    g.synchronize()
    print(json.dumps({"reads": n, "kernels_ms_per_pass": {k: round(v["ms"] / v["launches"], 3) for k, v in g.profile_read().items()}}))
    b.free()
