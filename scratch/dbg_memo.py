import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import rkmh_b200
from rkmh_b200.fastx import load_sequences
from oracle.oracle import Oracle
D='tests/golden/data/'
_, refs, _ = load_sequences([D+'hpv_16_allFasta.fa.gz'])
_, reads, _ = load_sequences([D+'minION25.fq.gz'])
ora = Oracle()
k=12
allh = np.concatenate([ora.calc_hashes(ora.to_upper(q), [k]) for q in reads])
probe = np.unique(allh[allh != 0])[::5]
res={}
for memo in (True, False):
    with rkmh_b200.RkmhGpu(0) as g:
        g.set_memo(memo)
        g.set_params([k], 1000, read_table_size=200_000_000)
        g.load_refs(*rkmh_b200.pack_sequences(refs))
        g.count_reset()
        g.count_batch(*rkmh_b200.pack_sequences(reads))
        t = g.count_table_tensor()
        res[memo] = (g.count_lookup(probe), int(t.sum().item()), g.classify_batch(*rkmh_b200.pack_sequences(reads), min_occ=2))
print("sum", res[True][1], res[False][1], "expected", int((allh!=0).sum()))
d = res[True][0] != res[False][0]
print("lookup diffs", d.sum(), "of", len(probe), res[True][0][d][:10], res[False][0][d][:10])
a,b = res[True][2], res[False][2]
for f in ("ref_id","best","prev","sketch_len"):
    print(f, (a[f]!=b[f]).sum(), a[f][:6], b[f][:6])
