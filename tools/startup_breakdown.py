import time, os, sys
sys.path.insert(0, os.getcwd())
t0=time.perf_counter()
import numpy as np
import rkmh_b200
from rkmh_b200.fastx import load_sequences
t1=time.perf_counter()
_, refs, _ = load_sequences(["tests/golden/data/all_pave_ref.fa.gz"])
t2=time.perf_counter()
g = rkmh_b200.RkmhGpu(0)
t3=time.perf_counter()
g.set_params([12], 1000)
t4=time.perf_counter()
g.load_refs(*rkmh_b200.pack_sequences(refs))
t5=time.perf_counter()
g.load_refs(*rkmh_b200.pack_sequences(refs))
t6=time.perf_counter()
g.profile_enable(True)
g.load_refs(*rkmh_b200.pack_sequences(refs))
print({k: round(v["ms"],2) for k,v in g.profile_read().items()})
print(f"import {t1-t0:.2f} refs-parse {t2-t1:.2f} create {t3-t2:.2f} set_params {t4-t3:.3f} load_refs(first) {t5-t4:.3f} load_refs(second) {t6-t5:.3f}")
