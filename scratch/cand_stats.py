import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
from rkmh_b200.fastx import load_sequences
from rkmh_b200.synth import sample_reads
from oracle.oracle import Oracle
D='tests/golden/data/'
_, refs, _ = load_sequences([D+'all_pave_ref.fa.gz'])
ora = Oracle()
sk, ln = ora.sketch_batch(refs, [12], 1000)
allref = np.unique(np.concatenate([sk[i,:ln[i]] for i in range(len(refs))]))
print("distinct ref hashes", len(allref))
# bits per row
from collections import Counter
cnt = Counter()
for i in range(len(refs)):
    for h in np.unique(sk[i,:ln[i]]): cnt[int(h)] += 1
v = np.array(list(cnt.values())); print("refs per hash: mean", v.mean(), "max", v.max(), np.percentile(v,[50,90,99]))
rng = np.random.default_rng(2)
b, o = sample_reads(rng, refs, 4000, 150)
nc=[]; nd=[]; dup=0
for i in range(4000):
    h = ora.calc_hashes(b[o[i]:o[i+1]].tobytes(), [12]); h = h[h!=0]
    m = np.isin(h, allref); c = h[m]
    nc.append(m.sum()); dup += (len(np.unique(c)) != len(c))
    refs_hit = set()
nc = np.array(nc)
print("cands/read mean", nc.mean(), "pcts", np.percentile(nc,[50,90,99,99.9]), "max", nc.max(), ">32:", (nc>32).mean(), ">64:", (nc>64).mean(), "reads with dup cands", dup/4000)
