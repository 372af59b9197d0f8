"""GPU parity tests (run on the B200 with `-m gpu`): every call goes through the C ABI
(include/rkmh_gpu.h) and is compared with the oracle / golden vectors of the real reference."""
import hashlib

import numpy as np
import pytest

from rkmh_b200.fastx import load_sequences
from tests.golden.make_golden import CLI_CASES, LIB_CASES, sketch_checksum
from tests.util import data_path, golden_cli_output

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["memo", "murmur"])
def gpu(request):
    """Every test runs twice: with the memoised path (single k <= 12, include/rkmh_gpu.h) and with the Murmur
    kernels forced for every k."""
    import rkmh_b200
    ctx = rkmh_b200.RkmhGpu(0)
    ctx.set_memo(request.param == "memo")
    yield ctx
    ctx.close()


def _pack(seqs):
    from rkmh_b200 import pack_sequences
    return pack_sequences(seqs)


def _fields(h):
    return np.stack([h["ref_id"], h["best"], h["prev"], h["sketch_len"]], axis=1)


# ---- hashing -------------------------------------------------------------------------------------------
def test_calc_hashes_full_dump(gpu, lib_vectors):
    _, qs, _ = load_sequences([data_path("minION25.fq.gz")])
    gpu.set_params([10], 1000)
    h, ho = gpu.calc_hashes(*_pack(qs[:1]))
    assert np.array_equal(h, lib_vectors["dump/minion0_k10"])
    _, rs, _ = load_sequences([data_path("all_pave_ref.fa.gz")])
    gpu.set_params([12], 1000)
    h, ho = gpu.calc_hashes(*_pack(rs[:1]))
    assert np.array_equal(h, lib_vectors["dump/pave0_k12"])
    sk, ln = gpu.sketch_batch(*_pack(rs[:1]))
    assert np.array_equal(sk[0, :ln[0]], lib_vectors["dump/pave0_k12_sketch"])


def test_calc_hash_kats(gpu, kats):
    for kat in kats["calc_hash"]:
        km = kat["kmer"].encode()
        if len(km) > 32:
            continue
        gpu.set_params([len(km)], 10)
        h, _ = gpu.calc_hashes(*_pack([km + b"A"]))
        assert int(h[0]) == int(kat["hash"]), kat["kmer"]


@pytest.mark.parametrize("k", list(range(1, 33)))
def test_calc_hashes_every_k_vs_oracle(gpu, oracle, k):
    rng = np.random.default_rng(100 + k)
    seqs = []
    for n in (k, k + 1, k + 2, 40, 151, 1000):
        s = rng.choice(np.frombuffer(b"ACGTacgtNnRY-*", np.uint8), size=n, p=[.22, .22, .22, .22, .02, .02, .02, .02, .01, .01, .005, .005, .005, .005])
        seqs.append(s.tobytes())
    gpu.set_params([k], 100)
    h, ho = gpu.calc_hashes(*_pack(seqs))
    exp = np.concatenate([oracle.calc_hashes(oracle.to_upper(s), [k]) for s in seqs])
    assert np.array_equal(h, exp)


def test_multi_k_order(gpu, oracle):
    _, qs, _ = load_sequences([data_path("awkward_reads.fq.gz")])
    ks = [12, 5, 20, 16, 32, 9]
    gpu.set_params(ks, 100)
    h, ho = gpu.calc_hashes(*_pack(qs))
    exp = np.concatenate([oracle.calc_hashes(oracle.to_upper(s), ks) for s in qs])
    assert np.array_equal(h, exp)


# ---- sketches and classification vs the real mkmh vectors ---------------------------------------------------
@pytest.mark.parametrize("case", sorted(LIB_CASES))
def test_against_real_mkmh_vectors(gpu, lib_vectors, case):
    ref_files, read_file, ks, s = LIB_CASES[case]
    _, rs, _ = load_sequences([data_path(f) for f in ref_files])
    _, qs, _ = load_sequences([data_path(read_file)])
    gpu.set_params(ks, s)
    gpu.load_refs(*_pack(rs))
    sk, lens = gpu.ref_sketches()
    assert np.array_equal(lens, lib_vectors[case + "/ref_len"])
    x = np.array([sketch_checksum(sk[i, :lens[i]]) for i in range(len(rs))], np.uint64)
    assert np.array_equal(x, lib_vectors[case + "/ref_sketch_x"])
    if case + "/ref_sketches" in lib_vectors:
        assert np.array_equal(sk, lib_vectors[case + "/ref_sketches"])
    hits = gpu.classify_batch(*_pack(qs))
    assert np.array_equal(_fields(hits), lib_vectors[case + "/hits"])
    # the resident-batch path gives the same records
    b = gpu.upload(*_pack(qs))
    b.classify()
    assert np.array_equal(_fields(b.results()), lib_vectors[case + "/hits"])
    b.free()
    # read sketches through the `hash` entry point
    rsk, rl = gpu.sketch_batch(*_pack(qs))
    assert np.array_equal(rl, lib_vectors[case + "/hits"][:, 3])
    rx = np.array([sketch_checksum(rsk[i, :rl[i]]) for i in range(len(qs))], np.uint64)
    assert np.array_equal(rx, lib_vectors[case + "/read_sketch_x"])


# ---- whole commands vs the real CLI's stdout ------------------------------------------------------------------
def _parse(argv):
    opt = {"r": [], "f": [], "k": [], "s": 1000, "M": -1, "N": -1, "D": 0, "I": None}
    it = iter(argv[1:])
    for a in it:
        v = next(it)
        key = a[1]
        if key in "rf":
            opt[key].append(data_path(v))
        elif key == "k":
            opt["k"].append(int(v))
        else:
            opt[key] = int(v)
    return opt


@pytest.mark.parametrize("name", sorted(CLI_CASES))
def test_cli_output_md5(gpu, cli_manifest, name):
    import rkmh_b200.api as api
    which, argv = CLI_CASES[name]
    o = _parse(argv)
    if which == "filter":
        out = api.run_filter(o["r"], o["f"], o["k"], o["s"], o["M"], o["N"], o["D"],
                             100000 if o["I"] is None else o["I"], ctx=gpu)
    else:
        out = api.run_stream(o["r"], o["f"], o["k"], o["s"], o["M"], o["N"], o["D"], o["I"], ctx=gpu)
    gold = golden_cli_output(name)
    if out != gold:
        a, b = out.splitlines(), gold.splitlines()
        diff = [(i, x, y) for i, (x, y) in enumerate(zip(a, b)) if x != y][:3]
        pytest.fail(f"{name}: {len(a)} vs {len(b)} lines; first diffs {diff}")
    assert hashlib.md5(out).hexdigest() == cli_manifest[name]["md5"]


# ---- edge cases ----------------------------------------------------------------------------------------------
def test_edge_cases_vs_oracle(gpu, oracle):
    _, rs, _ = load_sequences([data_path("awkward_refs.fa.gz")])
    reads = [b"A" * 200, b"ACGT" * 50, b"N" * 60, b"ACGTACGTA", b"ACGTACGT", b"ACG", b"", b"acgtacgtacgtnacgtacgtacgt",
             b"GATTACA" * 30, b"G" * 9 + b"{}|~" + b"C" * 20, rs[0], rs[1] + rs[1]]
    for ks, s in (([8], 50), ([8], 5), ([8, 4], 1000), ([31], 16), ([2], 3)):
        gpu.set_params(ks, s)
        gpu.load_refs(*_pack(rs))
        osk, ol = oracle.sketch_batch(rs, ks, s)
        sk, ln = gpu.ref_sketches()
        assert np.array_equal(sk, osk) and np.array_equal(ln, ol)
        hits = gpu.classify_batch(*_pack(reads))
        exp = oracle.classify_batch(reads, ks, s, osk, ol)
        assert np.array_equal(_fields(hits), _fields(exp)), (ks, s)
    # empty batch
    assert len(gpu.classify_batch(np.zeros(1, np.uint8), np.zeros(1, np.int64))) == 0


def test_errors_are_loud(gpu):
    import rkmh_b200
    with pytest.raises(rkmh_b200.RkmhError):
        gpu.set_params([33], 100)
    with pytest.raises(rkmh_b200.RkmhError):
        gpu.set_params([], 100)
    gpu.set_params([12], 100)
    with pytest.raises(rkmh_b200.RkmhError):
        gpu.classify_batch(*_pack([b"ACGTACGTACGTACGTAC"]))       # no references loaded
    with pytest.raises(rkmh_b200.RkmhError):
        gpu.count_batch(*_pack([b"ACGTACGTACGTACGTAC"]))          # no count table configured


def _synthetic_reads(rng, refs, n, length, sub=0.005, nrate=0.001, random_frac=0.1):
    from rkmh_b200.synth import sample_reads
    return sample_reads(rng, refs, n, length, sub, nrate, random_frac)


def test_random_reads_vs_oracle(gpu, oracle):
    """seeded cfg2-like reads at a size the oracle finishes in seconds; both kernels are exercised"""
    _, rs, _ = load_sequences([data_path("all_pave_ref.fa.gz")])
    rng = np.random.default_rng(2)
    bases, offs = _synthetic_reads(rng, rs, 3000, 150)
    reads = [bases[offs[i]:offs[i + 1]].tobytes() for i in range(len(offs) - 1)]
    reads[5] = b"A" * 150                      # duplicates inside one read
    reads[6] = (b"GTTAACTACCATCATTCATT" * 8)[:150]
    bases, offs = _pack(reads)
    gpu.set_params([12], 1000)
    gpu.load_refs(*_pack(rs))
    sk, ln = gpu.ref_sketches()
    hits = gpu.classify_batch(bases, offs)
    exp = oracle.classify_batch(reads, [12], 1000, sk, ln)
    assert np.array_equal(_fields(hits), _fields(exp))
    # long reads, sketch smaller than the read
    genome = b"".join(rs[:3])                   # 24 kb of concatenated genomes to cut 10 kb reads from
    bases, offs = _synthetic_reads(rng, [genome, genome[5000:] + genome[:5000]], 40, 10000, sub=0.04)
    reads = [bases[offs[i]:offs[i + 1]].tobytes() for i in range(len(offs) - 1)]
    hits = gpu.classify_batch(bases, offs)
    exp = oracle.classify_batch(reads, [12], 1000, sk, ln)
    assert np.array_equal(_fields(hits), _fields(exp))


def test_count_table_vs_oracle(gpu, oracle):
    _, rs, _ = load_sequences([data_path("new_refs.fa.gz")])
    _, qs, _ = load_sequences([data_path("mix2k.fq.gz")])
    qs = qs[:600]
    size = 200_000_000
    gpu.set_params([12], 1000, read_table_size=size)
    gpu.load_refs(*_pack(rs))
    gpu.count_reset()
    gpu.count_batch(*_pack(qs))
    counts = np.zeros(size, np.int32)
    oracle.count_batch(qs, [12], counts)
    allh = np.concatenate([oracle.calc_hashes(oracle.to_upper(q), [12]) for q in qs[:50]])
    probe = allh[allh != 0][::7]
    assert np.array_equal(gpu.count_lookup(probe), counts[(probe % np.uint64(size)).astype(np.int64)])
    t = gpu.count_table_tensor()
    assert int(t.sum().item()) == int(counts.sum())
    sk, ln = gpu.ref_sketches()
    for m in (1, 3, 10):
        hits = gpu.classify_batch(*_pack(qs), min_occ=m)
        exp = oracle.classify_batch(qs, [12], 1000, sk, ln, counts=counts, min_occ=m)
        assert np.array_equal(_fields(hits), _fields(exp)), m


def test_sharded_sketch_equals_whole(gpu, oracle):
    """`hash` of one record split over shards + merge == minhashes of the whole record."""
    rng = np.random.default_rng(5)
    seq = rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=3_000_000, p=[.2475, .2475, .2475, .2475, .01])
    k, s = 21, 10000
    gpu.set_params([k], s)
    whole, wl = gpu.sketch_batch(seq, np.array([0, seq.size], np.int64))
    nk = seq.size - k
    parts, lens = [], []
    nshard = 4
    for r in range(nshard):
        b, e = nk * r // nshard, nk * (r + 1) // nshard
        sk = gpu.sketch_shard(seq[b:min(e + k, seq.size)], b, e, seq.size)
        row = np.zeros(s, np.uint64)
        row[:len(sk)] = sk
        parts.append(row)
        lens.append(len(sk))
    merged = gpu.merge_sketches(np.stack(parts), lens)
    assert np.array_equal(merged, whole[0, :wl[0]])
    ref = oracle.minhashes(oracle.calc_hashes(seq[:400_000].tobytes(), [k]), s)
    part, pl = gpu.sketch_batch(seq[:400_000], np.array([0, 400_000], np.int64))
    assert np.array_equal(part[0, :pl[0]], ref)
    # the shard path itself against the oracle: a whole record as one shard, a middle shard, fewer windows than s,
    # a record of invalid bases only, and a period-4 repeat (a handful of distinct hashes, thousands of copies each)
    assert np.array_equal(gpu.sketch_shard(seq[:400_000], 0, 400_000 - k, 400_000), ref)
    mid = oracle.minhashes(oracle.calc_hashes(seq[100_000:250_000 + k].tobytes(), [k]), s)     # windows [100000, 250000)
    assert np.array_equal(gpu.sketch_shard(seq[100_000:250_000 + k], 100_000, 250_000, seq.size), mid)
    tiny = seq[:3000]
    assert np.array_equal(gpu.sketch_shard(tiny, 0, tiny.size - k, tiny.size), oracle.minhashes(oracle.calc_hashes(tiny.tobytes(), [k]), s))
    allN = np.full(5000, ord("N"), np.uint8)
    assert len(gpu.sketch_shard(allN, 0, allN.size - k, allN.size)) == 0
    rep = np.frombuffer(b"ACGT" * 50_000, np.uint8)
    gpu.set_params([k], 1000)
    assert np.array_equal(gpu.sketch_shard(rep, 0, rep.size - k, rep.size), oracle.minhashes(oracle.calc_hashes(rep.tobytes(), [k]), 1000))
    gpu.set_params([12], 777)                      # a memoisable k goes through the same kernels (Murmur hashing)
    m12 = oracle.minhashes(oracle.calc_hashes(seq[:300_000].tobytes(), [12]), 777)
    assert np.array_equal(gpu.sketch_shard(seq[:300_000], 0, 300_000 - 12, 300_000), m12)


def test_shard_sketch_in_pieces(oracle, monkeypatch):
    """rkg_sketch_shard copies a shard in 32 MB pieces and hashes behind the copies: a 110 Mbp record (four pieces, windows
    across every piece boundary, invalid bases right at the boundaries) gives the same sketch through the one-pass route,
    through the exact two-pass route (RKMH_SHARD_TWO_PASS), with pieces packed on the host (rkg_set_host_pack 2), and as
    four shards merged; its first 2 Mbp are checked against the oracle."""
    import rkmh_b200
    rng = np.random.default_rng(17)
    n = 110_000_000
    seq = rng.integers(0, 4, size=n, dtype=np.uint8)
    seq = np.frombuffer(b"ACGT", np.uint8)[seq]
    piece = 32 << 20
    for j in (1, 2, 3):
        seq[j * piece - 3:j * piece + 2] = ord("N")
        seq[j * piece - 40] = ord("a")
    k, s = 21, 10000
    nk = n - k
    out = {}
    for name, env, pack in (("one_pass", None, 0), ("two_pass", "1", 0), ("adaptive", None, 2)):
        if env:
            monkeypatch.setenv("RKMH_SHARD_TWO_PASS", env)
        else:
            monkeypatch.delenv("RKMH_SHARD_TWO_PASS", raising=False)
        with rkmh_b200.RkmhGpu(0) as g:
            g.set_params([k], s)
            g.set_host_pack(pack, 4)
            g.profile_enable(True)
            res = [g.sketch_shard(seq, 0, nk, n) for _ in range(3 if pack else 1)]
            launches = g.profile_read(reset=True)["sketch_kernel"]["launches"]
            assert all(np.array_equal(r, res[0]) for r in res)
            out[name] = res[0]
            if name == "one_pass":
                assert launches == 4 + 1, launches                     # one hashing launch per piece + the sort
                parts, lens = [], []
                for r in range(4):
                    b, e = nk * r // 4, nk * (r + 1) // 4
                    sk = g.sketch_shard(seq[b:min(e + k, n)], b, e, n)
                    row = np.zeros(s, np.uint64); row[:len(sk)] = sk
                    parts.append(row); lens.append(len(sk))
                out["merged"] = g.merge_sketches(np.stack(parts), lens)
                out["head"] = g.sketch_shard(seq[:2_000_000], 0, 2_000_000 - k, 2_000_000)
            if name == "two_pass":
                assert launches > 4 + 1, launches
            if pack:
                st = g.transfer_stats()
                assert st["packed_chunks"] > 0 and st["plain_chunks"] > 0, st
    assert len(out["one_pass"]) == s and np.all(np.diff(out["one_pass"].astype(np.float64)) >= 0)
    for name in ("two_pass", "adaptive", "merged"):
        assert np.array_equal(out[name], out["one_pass"]), name
    assert np.array_equal(out["head"], oracle.minhashes(oracle.calc_hashes(seq[:2_000_000].tobytes(), [k]), s))


def test_full_size_properties(gpu, oracle):
    """BASELINE config 2 at a tenth of its size (1 M reads) through size-independent properties: results do not
    depend on batch order or chunking, a spot sample equals the oracle, and every best <= sketch_len <= s."""
    _, rs, _ = load_sequences([data_path("all_pave_ref.fa.gz")])
    rng = np.random.default_rng(22)
    n = 1_000_000
    bases, offs = _synthetic_reads(rng, rs, n, 150)
    gpu.set_params([12], 1000)
    gpu.load_refs(*_pack(rs))
    hits = gpu.classify_batch(bases, offs)
    assert (hits["best"] <= hits["sketch_len"]).all() and (hits["sketch_len"] <= 138).all()
    assert (hits["prev"] < np.maximum(hits["best"], 0) + (hits["ref_id"] == 0)).all()
    # reversed read order gives reversed results
    mat = bases.reshape(n, 150)[::-1].copy()
    hits_r = gpu.classify_batch(mat.reshape(-1), offs)
    assert np.array_equal(_fields(hits_r), _fields(hits)[::-1])
    # a spot sample against the oracle
    idx = rng.choice(n, 400, replace=False)
    reads = [bases[offs[i]:offs[i + 1]].tobytes() for i in idx]
    sk, ln = gpu.ref_sketches()
    exp = oracle.classify_batch(reads, [12], 1000, sk, ln)
    assert np.array_equal(_fields(hits[idx]), _fields(exp))


@pytest.mark.parametrize("variant", ["sieve", "warp"])
def test_memo_path_is_taken_and_matches_murmur(oracle, variant, monkeypatch):
    """The memoised kernels run for a single k <= 12 and give the same records as the Murmur kernels: short reads
    (warp kernel), long reads (block kernel, bottom-s by rank), with and without -M."""
    import rkmh_b200
    from rkmh_b200.synth import ont_reads
    _, rs, _ = load_sequences([data_path("all_pave_ref.fa.gz")])
    _, hpv, _ = load_sequences([data_path("hpv_16_allFasta.fa.gz")])
    rng = np.random.default_rng(77)
    sb, so = _synthetic_reads(rng, rs, 20000, 150)
    lb, lo = ont_reads(rng, hpv, 300, 10000)
    res = {}
    # the two short-read kernels: shared-memory sieve (default), rank-table warp kernel
    monkeypatch.setenv("RKMH_NO_SIEVE", "1" if variant == "warp" else "0")
    short_kernel = f"classify_memo_{variant}_kernel"
    for memo in (True, False):
        with rkmh_b200.RkmhGpu(0) as g:
            g.set_memo(memo)
            g.set_params([12], 1000, read_table_size=rkmh_b200.STREAM_TABLE_SIZE)
            g.load_refs(*_pack(rs))
            assert g.memo_active() == memo
            g.profile_enable(True)
            out = [_fields(g.classify_batch(sb, so)), _fields(g.classify_batch(lb, lo))]
            g.count_reset()
            g.count_batch(lb, lo)
            g.count_batch(sb, so)
            t = g.count_table_tensor()
            out.append(int(t.sum().item()))
            for m in (2, 40):
                out.append(_fields(g.classify_batch(lb, lo, min_occ=m)))
                out.append(_fields(g.classify_batch(sb, so, min_occ=m)))
            out.append(_fields(g.classify_batch(sb, so)))          # -M off again: FAILM flags cleared
            names = set(g.profile_read())
            assert (short_kernel in names) == memo and ("classify_warp_kernel" in names) == (not memo)
            if memo and variant != "warp":                     # -M bakes FAILM flags: those passes take the rank-table kernel
                assert "classify_memo_warp_kernel" in names
            assert ("count_memo_kernel" in names) == memo
            res[memo] = out
    for a, b in zip(res[True], res[False]):
        assert np.array_equal(a, b)
    assert np.array_equal(res[True][0], res[True][-1])
    sk, ln = oracle.sketch_batch(rs, [12], 1000)
    reads = [lb[lo[i]:lo[i + 1]].tobytes() for i in range(12)]
    exp = oracle.classify_batch(reads, [12], 1000, sk, ln)
    assert np.array_equal(res[True][1][:12], _fields(exp))


@pytest.mark.parametrize("nrefs", [40, 100, 182, 250, 300])
def test_reference_count_variants(oracle, nrefs):
    """Every width of the per-reference bitset rows: <= 64 references (2 words), <= 128 (4), <= 192 (6), <= 256 (8:
    the widest register-row kernels) and more than 256 (shared-memory counters), each against the oracle and against
    the Murmur path.  References beyond the 182 fixture sequences are mutated copies, so that hits are shared by many."""
    import rkmh_b200
    _, rs, _ = load_sequences([data_path("all_pave_ref.fa.gz")])
    rng = np.random.default_rng(nrefs)
    refs = list(rs[:min(nrefs, len(rs))])
    while len(refs) < nrefs:
        src = np.frombuffer(rs[int(rng.integers(len(rs)))], np.uint8).copy()
        pos = rng.random(src.size) < 0.03
        src[pos] = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=int(pos.sum()))
        refs.append(src.tobytes())
    bases, offs = _synthetic_reads(rng, refs, 4000, 150)
    res = {}
    for memo in (True, False):
        with rkmh_b200.RkmhGpu(0) as g:
            g.set_memo(memo)
            g.set_params([12], 1000)
            g.load_refs(*_pack(refs))
            g.profile_enable(True)
            res[memo] = _fields(g.classify_batch(bases, offs))
            names = set(g.profile_read())
            if memo:
                assert ("classify_memo_sieve_kernel" in names) == (nrefs <= 256), names
    assert np.array_equal(res[True], res[False])
    sk, ln = oracle.sketch_batch(refs, [12], 1000)
    idx = rng.choice(4000, size=250, replace=False)
    reads = [bases[offs[i]:offs[i + 1]].tobytes() for i in idx]
    exp = oracle.classify_batch(reads, [12], 1000, sk, ln)
    assert np.array_equal(res[True][idx], _fields(exp))


# ---- round 2: fixed-length entry points, per-code partial counts, copy probe --------------------------------
def test_fixed_length_entry_points_equal_offsets(oracle):
    """rkg_classify_fixed / rkg_count_fixed (no offsets cross the link) == rkg_classify_batch / rkg_count_batch on
    the equivalent offsets, with and without -M, on both hashing paths; spot-checked against the oracle."""
    import rkmh_b200
    _, rs, _ = load_sequences([data_path("all_pave_ref.fa.gz")])
    rng = np.random.default_rng(11)
    bases, offs = _synthetic_reads(rng, rs, 30000, 150)
    for memo in (True, False):
        with rkmh_b200.RkmhGpu(0) as g:
            g.set_memo(memo)
            g.set_params([12], 1000, read_table_size=rkmh_b200.STREAM_TABLE_SIZE)
            g.load_refs(*_pack(rs))
            a = _fields(g.classify_batch(bases, offs))
            b = _fields(g.classify_fixed(bases, 150))
            assert np.array_equal(a, b)
            g.count_reset(); g.count_batch(bases, offs)
            t1 = g.count_table_tensor().clone()
            am = _fields(g.classify_batch(bases, offs, min_occ=3))
            g.count_reset(); g.count_fixed(bases, 150)
            t2 = g.count_table_tensor()
            assert bool((t1 == t2).all().item())
            bm = _fields(g.classify_fixed(bases, 150, min_occ=3))
            assert np.array_equal(am, bm)
            # the copies alone (rkg_copy_probe) run on the same buffers without touching the results of later calls
            out = np.zeros(30000, rkmh_b200.HIT_DTYPE)
            g.copy_probe_raw(bases.ctypes.data, None, 150, 30000, out.ctypes.data)
            g.copy_probe_raw(bases.ctypes.data, offs.ctypes.data, 0, 30000, out.ctypes.data)
            assert np.array_equal(_fields(g.classify_fixed(bases, 150)), a)
    sk, ln = oracle.sketch_batch(rs, [12], 1000)
    reads = [bases[offs[i]:offs[i + 1]].tobytes() for i in range(200)]
    assert np.array_equal(a[:200], _fields(oracle.classify_batch(reads, [12], 1000, sk, ln)))


@pytest.mark.parametrize("count_zeros", [False, True])
def test_count_partial_exchange_equals_one_table(count_zeros):
    """The -M exchange on the per-code partial counts (rkg_count_partial, 64 MB at k = 12): two contexts stand in for
    two ranks, each counts half of the reads, the partial arrays are summed into both (what the all-reduce does), and
    every context then holds the table -- and gives the records -- that one context counting all reads does.  One of
    the two "ranks" counts NO read in a second round (a rank with an empty shard still folds the sum)."""
    import rkmh_b200
    import torch
    from rkmh_b200.synth import ont_reads
    _, hpv, _ = load_sequences([data_path("hpv_16_allFasta.fa.gz")])
    rng = np.random.default_rng(5)
    bases, offs = ont_reads(rng, hpv, 120, 6000)
    bases = bases.copy()
    bases[rng.integers(0, bases.size, 500)] = ord("N")       # invalid windows: the count_zeros mode counts them in slot 0
    size = rkmh_b200.FILTER_TABLE_SIZE if count_zeros else rkmh_b200.STREAM_TABLE_SIZE
    half = 60
    lo = (bases[:offs[half]], offs[:half + 1])
    hi = (bases[offs[half]:], offs[half:] - offs[half])
    ctxs = []
    for _ in range(3):
        g = rkmh_b200.RkmhGpu(0)
        g.set_params([12], 1000, read_table_size=size)
        g.load_refs(*_pack(hpv))
        g.count_reset()
        ctxs.append(g)
    one, ra, rb = ctxs
    try:
        one.count_batch(bases, offs, count_zeros=count_zeros)
        ref_table = one.count_table_tensor().clone()
        ref_hits = _fields(one.classify_batch(bases, offs, min_occ=4))
        for shards in ((lo, hi), ((bases, offs), None)):
            for g, sh in zip((ra, rb), shards):
                g.count_reset()
                if sh is not None:
                    g.count_batch(*sh, count_zeros=count_zeros)
            pa, pb = ra.count_partial_tensor(count_zeros), rb.count_partial_tensor(count_zeros)
            assert pa.numel() == 4 ** 12 + 32
            total = pa + pb
            pa.copy_(total); pb.copy_(total)
            torch.cuda.synchronize()
            for g in (ra, rb):
                assert bool((g.count_table_tensor() == ref_table).all().item())
                assert np.array_equal(_fields(g.classify_batch(bases, offs, min_occ=4)), ref_hits)
    finally:
        for g in ctxs:
            g.close()


def test_limits_fail_loudly():
    """Arguments past the documented limits are refused with an error, never looped on or silently truncated."""
    import rkmh_b200
    with rkmh_b200.RkmhGpu(0) as g:
        with pytest.raises(rkmh_b200.RkmhError, match="sketch size"):
            g.set_params([12], (1 << 26) + 1)
        g.set_params([12], 1 << 26)
        with pytest.raises(rkmh_b200.RkmhError, match="2\\^29"):
            g.load_refs(*_pack([b"ACGTACGTACGTACGTACGT"] * 9))
        g.set_memo(False)
        g.set_params([20], 100, read_table_size=1000)
        g.load_refs(*_pack([b"ACGTACGTACGTACGTACGTTTGACA" * 3]))
        with pytest.raises(rkmh_b200.RkmhError, match="memoised path is not active"):
            g.count_partial_tensor()


# ---- hpv16: full hash sets ----------------------------------------------------------------------
@pytest.mark.parametrize("ks,min_occ", [([8], -1), ([6, 17], -1), ([12], 2), ([21], -1)])
def test_classify_sets_vs_oracle(oracle, ks, min_occ):
    """rkg_load_ref_sets + rkg_classify_sets == hash_set_intersection_size / hash_intersection_size of the oracle
    (mkmh/mkmh.hpp:842-863, 822-840 as main_hpv16 uses them): sets with zeros and duplicates, reads with repeated
    k-mers, invalid bases and reads shorter than k, several k-mer sizes, -M masking, both hashing paths."""
    import rkmh_b200
    _, refs, _ = load_sequences([data_path("awkward_refs.fa.gz"), data_path("new_refs.fa.gz")])
    _, reads, _ = load_sequences([data_path("awkward_reads.fq.gz"), data_path("mix2k.fq.gz")])
    reads = reads[:300] + [b"ACGT" * 30, b"A" * 80, b"N" * 30, b"ACGTAC"]
    k0 = ks[0]
    sets = [oracle.calc_hashes(oracle.to_upper(r), [k0]) for r in refs]          # unsorted, zeros and duplicates inside
    uniq = [np.setdiff1d(np.unique(sets[i]), np.concatenate(sets[:i] + sets[i + 1:])) for i in range(4)]
    all_sets = sets + uniq                                                       # types first, then "specific" tables
    ntypes = len(sets)
    size = 1_000_003
    counts = None
    if min_occ >= 0:
        counts = np.zeros(size, np.int32)
        oracle.count_batch(reads, ks, counts)
    exp_hits, exp_extra = oracle.classify_sets(reads, ks, [np.sort(x) for x in all_sets], ntypes, counts=counts, min_occ=max(min_occ, 0))
    for memo in (True, False):
        with rkmh_b200.RkmhGpu(0) as g:
            g.set_memo(memo)
            g.set_params(ks, 1, read_table_size=size if min_occ >= 0 else 0)
            g.load_ref_sets(all_sets)
            if min_occ >= 0:
                g.count_reset()
                g.count_batch(*_pack(reads))
            hits, extra = g.classify_sets(*_pack(reads), ntypes, min_occ=min_occ)
            with pytest.raises(rkmh_b200.RkmhError):
                g.classify_batch(*_pack(reads))          # a set table is not a sketch table
        assert np.array_equal(hits["ref_id"], exp_hits["ref_id"]) and np.array_equal(hits["best"], exp_hits["best"])
        assert np.array_equal(extra, exp_extra)


@pytest.mark.parametrize("name", ["hpv16cmd_minion_k16", "hpv16cmd_mix2k_k12"])
def test_hpv16_through_the_c_abi_equals_the_real_binary(cli_manifest, name):
    """the Python mirror of main_hpv16 with the GPU as its engine reproduces the real binary's stdout"""
    import rkmh_b200
    from tests.golden.make_golden import HPV16_CASES
    from tests.hpv16_mirror import GpuEngine, hpv16_stdout
    argv = HPV16_CASES[name]
    ks = [int(argv[i + 1]) for i, a in enumerate(argv) if a == "-k"]
    reads = [data_path(argv[i + 1]) for i, a in enumerate(argv) if a == "-f"]
    with rkmh_b200.RkmhGpu(0) as g:
        out, side = hpv16_stdout(GpuEngine(g), data_path("all_pave_ref.fa.gz"), data_path("new_refs.fa.gz"), reads, ks)
    assert out == golden_cli_output(name)
    assert hashlib.md5(side).hexdigest() == cli_manifest[name]["side_file_md5"]


def test_pre_hashed_references_and_count_table_round_trip(oracle):
    """rkg_load_ref_sketches on the sketches rkg_get_ref_sketches returns == rkg_load_refs on the sequences (records of a
    read batch identical, memoised and Murmur path); rkg_count_export -> rkg_count_import restores the read count table
    in a fresh context, whose -M classification then equals the original's.  Malformed sketches are refused."""
    import rkmh_b200
    _, rs, _ = load_sequences([data_path("all_pave_ref.fa.gz")])
    rng = np.random.default_rng(23)
    bases, offs = _synthetic_reads(rng, rs, 8000, 150)
    for memo in (True, False):
        with rkmh_b200.RkmhGpu(0) as a, rkmh_b200.RkmhGpu(0) as b:
            for g in (a, b):
                g.set_memo(memo)
                g.set_params([12], 1000, read_table_size=rkmh_b200.STREAM_TABLE_SIZE)
            a.load_refs(*_pack(rs))
            sk, ln = a.ref_sketches()
            b.load_ref_sketches(sk, ln)
            assert b.memo_active() == memo
            ha = _fields(a.classify_batch(bases, offs))
            assert np.array_equal(ha, _fields(b.classify_batch(bases, offs)))
            a.count_reset(); a.count_batch(bases, offs)
            slots, vals = a.count_export()
            assert len(slots) > 1000 and int(vals.sum()) == int(a.count_table_tensor().sum().item())
            b.count_import(slots, vals)
            assert bool((a.count_table_tensor() == b.count_table_tensor()).all().item())
            assert np.array_equal(_fields(a.classify_batch(bases, offs, min_occ=3)), _fields(b.classify_batch(bases, offs, min_occ=3)))
            bad = sk.copy(); bad[3, 5] = bad[3, 4] - np.uint64(1)
            with pytest.raises(rkmh_b200.RkmhError, match="ascending"):
                b.load_ref_sketches(bad, ln)
    exp = oracle.classify_batch([bases[offs[i]:offs[i + 1]].tobytes() for i in range(100)], [12], 1000, *oracle.sketch_batch(rs, [12], 1000))
    assert np.array_equal(ha[:100], _fields(exp))


def test_call_depth_entry_points_vs_oracle(oracle):
    """rkg_depth_*: the exact multiplicity of every read hash (zeros = invalid windows included, as the reference's
    unordered_map counts them, src/rkmh.cpp:1617-1622), the depth of every reference window, and the depths of the
    substitution / deletion alternatives of listed windows, against counts made with the oracle's calc_hash."""
    import collections
    import rkmh_b200
    _, refs, _ = load_sequences([data_path("hpv_16.fa.gz")])
    _, reads, _ = load_sequences([data_path("mix2k.fq.gz")])
    reads = reads[:400] + [b"ACGTNNNNACGTACGTAGGATCCANNNN", b"ACGTACG"]
    ref = oracle.to_upper(refs[0])[:1500] + b"NNACGTRYACGT" + oracle.to_upper(refs[0])[1500:1700]
    for k in (11, 16, 23):
        depth = collections.Counter()
        for r in reads:
            depth.update(int(h) for h in oracle.calc_hashes(oracle.to_upper(r), [k]))
        with rkmh_b200.RkmhGpu(0) as g:
            g.set_params([k], 1)
            g.depth_begin(sum(max(len(r) - k, 0) for r in reads))
            g.depth_add(*_pack(reads))
            with pytest.raises(rkmh_b200.RkmhError, match="expect"):
                g.depth_add(*_pack(reads[:1]))
            got = g.depth_of_sequence(ref)
            exp = np.array([depth.get(int(h), 0) for h in oracle.calc_hashes(ref, [k])], np.int32)
            assert np.array_equal(got, exp)
            wins = np.array([0, 1, 7, 1490, 1495, 1502, len(ref) - k - 1], np.int64)
            alt = g.depth_rescue(ref, wins)
        rot = {ord("A"): b"CTG", ord("C"): b"TGA", ord("T"): b"CGA", ord("G"): b"ACT"}
        for w, j in enumerate(wins):
            j = int(j)
            for p in range(k):
                for x in range(3):
                    orig = ref[j + p]
                    if orig not in rot:
                        assert alt[w, 3 * p + x] == -1
                        continue
                    km = ref[j:j + p] + bytes([rot[orig][x]]) + ref[j + p + 1:j + k]
                    assert alt[w, 3 * p + x] == depth.get(oracle.calc_hash(km), 0), (k, j, p, x)
            for d in range(1, k + 1):
                if j == 0:
                    assert alt[w, 3 * k + d - 1] == -1
                    continue
                big = ref[j - 1:j + k]
                km = big[:d] + big[d + 1:]
                assert alt[w, 3 * k + d - 1] == depth.get(oracle.calc_hash(km), 0), (k, j, d)


def test_packed_transfer_gives_the_same_records(oracle):
    """rkg_set_host_pack: the bases are packed to 2-bit codes + an invalid-base mask on the HOST and cross the link at
    0.375 B/base; the memoised kernels get them unpacked on the device (every invalid byte as N, every letter upper-cased).
    Records, count tables and sketches are those of the plain transfer: ragged reads with lower case, N, IUPAC and
    punctuation bytes, reads shorter than k, fixed-length batches, both hashing paths, with and without -M; and the same
    with the adaptive mode (2), which packs some chunks and sends others as they are."""
    import rkmh_b200
    _, rs, _ = load_sequences([data_path("all_pave_ref.fa.gz")])
    _, aw, _ = load_sequences([data_path("awkward_reads.fq.gz"), data_path("minION25.fq.gz")])
    rng = np.random.default_rng(31)
    sb, so = _synthetic_reads(rng, rs, 20000, 150)
    weird = aw + [b"acgtnnacgtRYKMacgtacgt{}|~ACGTACGTACGTAC", b"ACGTACG", b"", b"N" * 70, rs[3][:3001].lower(), rs[5][100:9000]]
    wb, wo = _pack(weird)
    for memo in (True, False):
        res = {}
        for pack in (0, 1, 2):
            with rkmh_b200.RkmhGpu(0) as g:
                g.set_memo(memo)
                g.set_host_pack(pack, 3)
                g.set_params([12], 1000, read_table_size=rkmh_b200.STREAM_TABLE_SIZE)
                g.load_refs(*_pack(rs))
                out = [_fields(g.classify_batch(sb, so)), _fields(g.classify_fixed(sb, 150)), _fields(g.classify_batch(wb, wo))]
                g.count_reset(); g.count_batch(wb, wo); g.count_fixed(sb, 150)
                out.append(g.count_table_tensor().cpu().numpy().copy())
                out.append(_fields(g.classify_batch(wb, wo, min_occ=2)))
                out.append(_fields(g.classify_batch(sb, so, min_occ=3)))
                sk, ln = g.sketch_batch(wb, wo)
                out += [sk, ln]
                res[pack] = out
        for mode in (1, 2):
            for a, b in zip(res[0], res[mode]):
                assert np.array_equal(a, b)
    sk, ln = oracle.sketch_batch(rs, [12], 1000)
    exp = oracle.classify_batch(weird, [12], 1000, sk, ln)
    assert np.array_equal(res[2][2], _fields(exp))
    # several 64 MB chunks in flight: the adaptive mode mixes packed and plain chunks inside one call
    big, bo = _synthetic_reads(rng, rs, 1_500_000, 150)
    big[::997] = ord("N"); big[5::1201] = ord("a")
    with rkmh_b200.RkmhGpu(0) as g:
        g.set_params([12], 1000, read_table_size=rkmh_b200.STREAM_TABLE_SIZE)
        g.load_refs(*_pack(rs))
        plain = _fields(g.classify_fixed(big, 150))
        g.count_reset(); g.count_fixed(big, 150)
        ct = g.count_table_tensor().cpu().numpy().copy()
        g.set_host_pack(2, 4)
        for _ in range(3):
            assert np.array_equal(_fields(g.classify_fixed(big, 150)), plain)
        assert np.array_equal(_fields(g.classify_batch(big, bo)), plain)
        g.count_reset(); g.count_fixed(big, 150)
        assert np.array_equal(g.count_table_tensor().cpu().numpy(), ct)
