"""CPU tests of the kernel LOGIC: rkmh_b200/csrc/rk_kernels.cuh compiled by g++ against the CUDA stand-in
of tests/emu/cuda_emu.h, compared with the oracle.  The same kernels are checked on the B200 by the
`-m gpu` tests; this keeps logic regressions visible in the GPU-less container."""
import random

import numpy as np
import pytest

from rkmh_b200.fastx import load_sequences
from tests.util import data_path


@pytest.fixture(scope="module")
def emu():
    from tests.emu.emu import Emu
    return Emu()


def test_fastmod_exact(emu):
    rnd = random.Random(1)
    for d in [200_000_000, 10_000_000, 3, 7, 1_000_003, 2**31 - 1, 2**40 + 15, 1 << 20, 640_000]:
        for n in [0, 1, d - 1, d, d + 1, 2**64 - 1, 2**63] + [rnd.getrandbits(64) for _ in range(50)]:
            assert emu.lib.emu_fastmod(n, d) == n % d


def test_rolling_murmur_all_k(emu, oracle):
    rnd = random.Random(2)
    for k in range(1, 33):
        for _ in range(10):
            km = "".join(rnd.choice("ACGT") for _ in range(k)).encode()
            assert emu.lib.emu_kmer_hash(km, k) == oracle.calc_hash(km)
            assert emu.lib.emu_window_hash(km, k) == oracle.calc_hash(km)       # lane-per-window path (rk_hash.cuh)


@pytest.mark.parametrize("ks", [[8], [4], [9, 16], [17], [25], [32], [12, 20]])
def test_hash_kernel(emu, oracle, ks):
    _, seqs, _ = load_sequences([data_path("awkward_reads.fq.gz")])
    h, _ = emu.hash(seqs, ks)
    exp = np.concatenate([oracle.calc_hashes(oracle.to_upper(s), ks) for s in seqs])
    assert np.array_equal(h, exp)


def _classify_case(emu, oracle, ref_files, read_file, ks, s, nreads=None, nrefs=None, **kw):
    _, refs, _ = load_sequences([data_path(f) for f in ref_files])
    _, reads, _ = load_sequences([data_path(read_file)])
    reads = reads[:nreads] if nreads else reads
    refs = refs[:nrefs] if nrefs else refs
    hits, rsk, rl, redo = emu.classify(refs, reads, ks, s, **kw)
    osk, ol = oracle.sketch_batch(refs, ks, s)
    oh = oracle.classify_batch(reads, ks, s, osk, ol)
    assert np.array_equal(rsk, osk) and np.array_equal(rl, ol)
    for f in ("ref_id", "best", "prev", "sketch_len"):
        assert np.array_equal(hits[f], oh[f]), f
    return redo


@pytest.mark.parametrize("kw", [dict(path=0), dict(path=1, glog2=3), dict(path=1, glog2=5), dict(path=2),
                                dict(path=0, hbuf_global=True, hs_global=True)])
def test_classify_awkward(emu, oracle, kw):
    _classify_case(emu, oracle, ["awkward_refs.fa.gz"], "awkward_reads.fq.gz", [8], 50, **kw)


def test_classify_small_sketch_and_multik(emu, oracle):
    _classify_case(emu, oracle, ["awkward_refs.fa.gz"], "awkward_reads.fq.gz", [8], 10, path=1, glog2=4)
    _classify_case(emu, oracle, ["awkward_refs.fa.gz"], "awkward_reads.fq.gz", [6, 17], 20, path=1, glog2=4)
    _classify_case(emu, oracle, ["awkward_refs.fa.gz"], "awkward_reads.fq.gz", [8], 10, path=2)
    _classify_case(emu, oracle, ["awkward_refs.fa.gz"], "awkward_reads.fq.gz", [6, 17], 20, path=2)
    _classify_case(emu, oracle, ["awkward_refs.fa.gz"], "awkward_reads.fq.gz", [3, 21, 13, 32], 3000, path=2)


@pytest.mark.parametrize("k", [4, 12, 16, 21, 32])
def test_classify_warp_kernel_every_variant(emu, oracle, k):
    """classify_warp_kernel (rk_warp.cuh) for every key-length variant: sketch = every hash, no read on the redo list"""
    redo = _classify_case(emu, oracle, ["awkward_refs.fa.gz"], "awkward_reads.fq.gz", [k], 5000, path=2)
    assert redo == 0 or k < 10          # tiny k: more than 32 hits per read is the rule, those reads take the exact path


def test_classify_fixtures(emu, oracle):
    assert _classify_case(emu, oracle, ["hpv_16_allFasta.fa.gz"], "minION25.fq.gz", [10], 1000, nreads=4, path=1) == 4
    assert _classify_case(emu, oracle, ["all_pave_ref.fa.gz"], "z1.fq.gz", [12], 1000, nreads=48, nrefs=70, path=1,
                          glog2=4) == 0
    _classify_case(emu, oracle, ["zika.fa.gz", "dengue.fa.gz", "yellow_fever.fa.gz", "hpv_16.fa.gz"], "z1.fq.gz", [20],
                   2000, nreads=64, path=1, glog2=4)
    assert _classify_case(emu, oracle, ["all_pave_ref.fa.gz"], "z1.fq.gz", [12], 1000, nreads=48, nrefs=70, path=2) == 0
    _classify_case(emu, oracle, ["zika.fa.gz", "dengue.fa.gz", "yellow_fever.fa.gz", "hpv_16.fa.gz"], "z1.fq.gz", [20],
                   2000, nreads=64, path=2)


def test_count_mask_and_filters(emu, oracle):
    _, refs, _ = load_sequences([data_path("new_refs.fa.gz")])
    _, reads, _ = load_sequences([data_path("mix2k.fq.gz")])
    reads = reads[:120]
    size = 1_000_003
    c1 = np.zeros(size, np.int32)
    emu.count(reads, [12], c1)
    c2 = np.zeros(size, np.int32)
    oracle.count_batch(reads, [12], c2)
    assert np.array_equal(c1, c2)
    for path in (0, 1):
        hits, _, _, _ = emu.classify(refs, reads, [12], 1000, counts=c1, min_occ=2, path=path, glog2=5)
        osk, ol = oracle.sketch_batch(refs, [12], 1000)
        oh = oracle.classify_batch(reads, [12], 1000, osk, ol, counts=c2, min_occ=2)
        for f in ("ref_id", "best", "prev", "sketch_len"):
            assert np.array_equal(hits[f], oh[f]), (path, f)
    sk, ln = emu.sketch(refs, [12], 1000, counts=c1, mode=1, occ=3)
    osk, ol = oracle.sketch_batch(refs, [12], 1000, counts=c2, mode=1, occ=3)
    assert np.array_equal(sk, osk) and np.array_equal(ln, ol)


def test_sketch_and_merge(emu, oracle):
    _, refs, _ = load_sequences([data_path("new_refs.fa.gz")])
    refs = refs[:4]
    for s, kw in ((300, dict(force_global_sort=True)), (20000, {}), (1000, {})):
        sk, ln = emu.sketch(refs, [12], s, **kw)
        osk, ol = oracle.sketch_batch(refs, [12], s)
        assert np.array_equal(sk, osk) and np.array_equal(ln, ol)
    parts, pl = oracle.sketch_batch(refs, [12], 1000)
    whole = oracle.minhashes(np.concatenate([oracle.calc_hashes(oracle.to_upper(r), [12]) for r in refs]), 1000)
    assert np.array_equal(emu.merge(parts, pl, 1000), whole)
    assert np.array_equal(emu.merge(parts, pl, 1000, in_smem=False), whole)


# ---- memoised path (rk_memo.cuh): same results as the oracle, through the table build + streaming kernels ----------
def _memo_case(emu, oracle, refs, reads, k, s, counts=None, min_occ=-1, **kw):
    hits, redo = emu.classify_memo(refs, reads, k, s, counts=counts, min_occ=min_occ, **kw)
    osk, ol = oracle.sketch_batch(refs, [k], s)
    oh = oracle.classify_batch(reads, [k], s, osk, ol, counts=counts, min_occ=min_occ if counts is not None else 0)
    for f in ("ref_id", "best", "prev", "sketch_len"):
        assert np.array_equal(hits[f], oh[f]), (f, kw)
    return redo


@pytest.mark.parametrize("kw", [dict(path=0), dict(path=1), dict(path=2), dict(path=0, vcap_limit=16),
                                dict(path=1, block=128), dict(path=4), dict(path=4, block=128)])
def test_memo_classify_awkward(emu, oracle, kw):
    _, refs, _ = load_sequences([data_path("awkward_refs.fa.gz")])
    _, reads, _ = load_sequences([data_path("awkward_reads.fq.gz")])
    reads = reads + [b"A" * 70, b"ACGT" * 20, b"N" * 40, b"ACGTACGTA", b"ACGTACGT", b"", b"acgtacgtacgtnacgtacgtacgt",
                     b"G" * 9 + b"{}|~" + b"C" * 20, refs[0], refs[1] + refs[1]]
    combos = ((8, 50), (3, 7)) if kw.get("block") else ((8, 50), (8, 5), (5, 1000), (3, 7))    # keep the CPU suite short
    for k, s in combos:
        _memo_case(emu, oracle, refs, reads, k, s, **kw)


@pytest.mark.parametrize("block", [64, 256])
def test_memo_block_kernel_bound_cases(emu, oracle, block):
    """classify_memo_block_kernel keeps only the windows under a GUESSED rank bound (from the previous read of the block)
    and falls back to keeping everything when the guess fails: reads that make consecutive guesses wrong in both
    directions (mostly-N after random, one repeated k-mer, tandem repeats, s close to the number of windows), with and
    without a -M mask, must still give the oracle's records."""
    rng = np.random.default_rng(11)
    _, refs, _ = load_sequences([data_path("awkward_refs.fa.gz")])
    def rnd(n, pn=0.0):
        return rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=n, p=[(1 - pn) / 4] * 4 + [pn]).tobytes()
    reads = [rnd(3000), rnd(3000, 0.15), b"A" * 2500, rnd(2600), (b"ACGGTCA" * 400), rnd(900) + b"N" * 1500 + rnd(700),
             refs[0] * 3, rnd(1100), b"N" * 700, rnd(40), rnd(5000), refs[1] + rnd(2000) + refs[2]]
    for s in (60, 1000):
        _memo_case(emu, oracle, refs, reads, 8, s, path=0, block=block)
    c = np.zeros(10_007, np.int32)
    oracle.count_batch(reads, [8], c)
    _memo_case(emu, oracle, refs, reads, 8, 300, counts=c, min_occ=40, path=0, block=block)
    _memo_case(emu, oracle, refs, reads, 6, 500, path=0, block=block, vcap_limit=32)


def test_memo_classify_fixtures(emu, oracle):
    _, refs, _ = load_sequences([data_path("hpv_16_allFasta.fa.gz")])
    _, reads, _ = load_sequences([data_path("minION25.fq.gz")])
    assert _memo_case(emu, oracle, refs, reads[:3], 8, 300, path=1) == 3        # long reads: all through the block kernel
    _, short, _ = load_sequences([data_path("z1.fq.gz")])
    assert _memo_case(emu, oracle, refs, short[:40], 8, 1000, path=1) == 0
    _memo_case(emu, oracle, refs, short[:300] + reads[:1], 8, 1000, path=4)     # sieve kernel (+ its redo list)
    _memo_case(emu, oracle, refs, short[:120], 6, 1000, path=4)                 # dense sieve: many false positives


def test_memo_count_and_min_occ(emu, oracle):
    _, refs, _ = load_sequences([data_path("new_refs.fa.gz")])
    _, reads, _ = load_sequences([data_path("mix2k.fq.gz")])
    reads = reads[:60] + [b"ACGTNNACGTTTGACCA" * 9, b"N" * 50]
    size = 10_007
    for zeros in (False, True):
        c1 = np.zeros(size, np.int32)
        emu.count_memo(reads, 8, c1, count_zeros=zeros, wpr=1)
        c2 = np.zeros(size, np.int32)
        oracle.count_batch(reads, [8], c2, valid_only=not zeros)
        assert np.array_equal(c1, c2), zeros
        c3 = np.zeros(size, np.int32)
        emu.count_memo(reads, 8, c3, count_zeros=zeros, wpr=2)
        assert np.array_equal(c3, c2), zeros
    c = np.zeros(size, np.int32)
    oracle.count_batch(reads, [8], c)
    for m in (1, 30):
        for path in (0, 1):                       # with a -M mask baked the sieve kernel is not eligible
            _memo_case(emu, oracle, refs[:4], reads, 8, 200, counts=c, min_occ=m, path=path)
    _memo_case(emu, oracle, refs[:4], reads, 8, 200, counts=c, min_occ=30, path=4)     # falls back to the warp kernel


@pytest.mark.parametrize("hbuf_global", [False, True])
def test_classify_sets_kernels(emu, oracle, hbuf_global):
    """The engine of `rkmh hpv16` (rkg_load_ref_sets + rkg_classify_sets): a set table built from FULL hash sets (with the
    zeros and duplicates the reference's arrays carry) and the exact block kernel in set mode == the oracle's restatement of
    src/rkmh.cpp:2643-2706 (hash_set_intersection_size for the type sets, per-read counts of the lineage sets), with and
    without a -M mask, hash buffer in shared memory and in global scratch."""
    k = 10
    _, refs, _ = load_sequences([data_path("hpv_16_allFasta.fa.gz")])
    _, reads, _ = load_sequences([data_path("minION25.fq.gz")])
    reads = reads[:6] + [b"N" * 50, b"ACGT", refs[0][100:400]]
    sets = []
    for r in refs[:5]:
        h = np.sort(oracle.calc_hashes(oracle.to_upper(r[:3000]), [k]))
        sets.append(h)                                             # ascending, zeros and duplicates kept
    lineage = [np.unique(s[s != 0])[::7] for s in sets[:3]]        # "specific hash" style subsets as the extra sets
    allsets = sets + lineage
    ntypes = len(sets)
    hits, extra = emu.classify_sets(allsets, reads, k, ntypes, hbuf_global=hbuf_global)
    oh, oe = oracle.classify_sets(reads, [k], allsets, ntypes)
    for f in ("ref_id", "best"):                                   # (main_hpv16 keeps no runner-up: `prev` is not part of it)
        assert np.array_equal(hits[f], oh[f]), f
    assert np.array_equal(extra, oe)
    c = np.zeros(10_007, np.int32)
    oracle.count_batch(reads, [k], c)
    hits, extra = emu.classify_sets(allsets, reads, k, ntypes, counts=c, min_occ=2, hbuf_global=hbuf_global)
    oh, oe = oracle.classify_sets(reads, [k], allsets, ntypes, counts=c, min_occ=2)
    for f in ("ref_id", "best"):
        assert np.array_equal(hits[f], oh[f]), ("M", f)
    assert np.array_equal(extra, oe)


@pytest.mark.parametrize("k", [9, 16, 24])
def test_call_depth_kernels(emu, oracle, k):
    """rk_call.cuh (`rkmh call`, src/rkmh.cpp:1570-1858): the exact k-mer depth table of the reads, the depth of every window
    of the reference, and the depth of every single-base substitution (rotate_snps order) and deletion of chosen windows,
    against a dictionary built from the oracle's hashes."""
    from collections import Counter
    rng = np.random.default_rng(k)
    ref = bytearray(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=600).tobytes())
    ref[100] = ord("N"); ref[300:303] = b"acg"
    ref = bytes(ref)
    up = oracle.to_upper(ref)
    reads = []
    for _ in range(30):
        a = int(rng.integers(0, 450)); r = bytearray(up[a:a + int(rng.integers(k + 1, 150))])
        if rng.random() < 0.5:
            r[int(rng.integers(0, len(r)))] = ord("ACGT"[int(rng.integers(0, 4))])
        reads.append(bytes(r))
    reads += [b"N" * 40, b"ACG"]
    table = Counter()
    for r in reads:
        table.update(int(h) for h in oracle.calc_hashes(r, [k]))
    windows = np.array([0, 1, 57, 99, 250, 299, len(ref) - k - 1], np.int64)
    depth, resc = emu.call_depth(reads, up, k, windows)
    want = np.array([table.get(int(h), 0) for h in oracle.calc_hashes(up, [k])], np.int32)
    assert np.array_equal(depth, want)
    rot = {"A": "CTG", "C": "TGA", "T": "CGA", "G": "ACT"}
    def dep(kmer):
        return table.get(oracle.calc_hash(kmer) if oracle.canonical(kmer) else 0, 0)
    for wi, j in enumerate(windows):
        j = int(j)
        for a in range(4 * k):
            if a < 3 * k:
                pos, which = divmod(a, 3)
                orig = chr(up[j + pos])
                if orig not in rot:
                    exp = -1
                else:
                    km = bytearray(up[j:j + k]); km[pos] = ord(rot[orig][which]); exp = dep(bytes(km))
            else:
                d = a - 3 * k + 1
                if j == 0:
                    exp = -1
                else:
                    kp = up[j - 1:j + k]
                    exp = dep(kp[:d] + kp[d + 1:])
            assert resc[wi, a] == exp, (k, j, a)


def test_sketch_shard_kernels(emu, oracle):
    """rkg_sketch_shard's histogram / pick / emit / sort kernels == minhashes(calc_hashes()) of the oracle: more windows
    than s, fewer than s, invalid bases, a period-4 repeat whose few hashes occur thousands of times, and a middle shard."""
    rng = np.random.default_rng(3)
    seq = rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=30000, p=[.245, .245, .245, .245, .02]).tobytes()
    for k, s in ((21, 500), (12, 64), (5, 3000)):
        n = len(seq) - k
        assert np.array_equal(emu.sketch_shard(seq, n, k, s), oracle.minhashes(oracle.calc_hashes(seq, [k]), s)), (k, s)
    mid = seq[5000:9000 + 21]
    assert np.array_equal(emu.sketch_shard(mid, 4000, 21, 300), oracle.minhashes(oracle.calc_hashes(mid, [21]), 300))
    tiny = seq[:400]
    assert np.array_equal(emu.sketch_shard(tiny, len(tiny) - 21, 21, 1000), oracle.minhashes(oracle.calc_hashes(tiny, [21]), 1000))
    assert len(emu.sketch_shard(b"N" * 300, 279, 21, 50)) == 0
    rep = b"ACGT" * 2000
    assert np.array_equal(emu.sketch_shard(rep, len(rep) - 21, 21, 100), oracle.minhashes(oracle.calc_hashes(rep, [21]), 100))
    # the one-pass route (hashes at or under a limit guessed from s / n, piece by piece) and its fall-back to the exact one
    routes = {}
    for name, sq, k, s in (("random", seq, 21, 500), ("random_k12", seq, 12, 64), ("k5", seq, 5, 3000), ("tiny", tiny, 21, 1000),
                           ("allN", b"N" * 300, 21, 50), ("repeat", rep, 21, 100), ("mostlyN", b"N" * 9000 + seq[:1500], 21, 400),
                           ("one_kmer", b"A" * 5000, 12, 10)):
        got, route = emu.sketch_shard(sq, len(sq) - k, k, s, scan_first=True, pieces=4, want_route=True)
        assert np.array_equal(got, oracle.minhashes(oracle.calc_hashes(sq, [k]), s)), name
        routes[name] = route
    assert routes["random"] == 1 and routes["tiny"] == 1 and routes["allN"] == 1     # guess accepted
    assert routes["repeat"] == 2 and routes["mostlyN"] == 2                          # guess refused: exact route
