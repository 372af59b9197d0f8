import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import rkmh_b200
from rkmh_b200.fastx import load_sequences
from rkmh_b200.synth import ont_reads
_, refs, _ = load_sequences(['tests/golden/data/hpv_16_allFasta.fa.gz'])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
bases, offs = ont_reads(np.random.default_rng(3), refs, n, 10000)
with rkmh_b200.RkmhGpu(0) as gpu:
    gpu.set_params([12], 1000, read_table_size=rkmh_b200.STREAM_TABLE_SIZE)
    gpu.load_refs(*rkmh_b200.pack_sequences(refs))
    b = gpu.upload(bases, offs)
    gpu.count_reset()
    for _ in range(2):
        b.count(); b.classify(min_occ=1)
    gpu.synchronize()
