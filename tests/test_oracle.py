"""CPU tests: pin the C restatement (oracle/rkmh_oracle.c) against the reference's own known-answer
vectors and against golden vectors produced by the real reference (tests/golden/make_golden.py)."""
import os
import struct

import numpy as np
import pytest

from tests.util import data_path
from rkmh_b200.fastx import load_sequences
from tests.golden.make_golden import LIB_CASES, sketch_checksum


def test_murmur3_kats(oracle, kats):
    # literals of the reference's own test, mkmh/murmur3/test.c:42-45
    for kat in kats["murmur3_x64_128"]:
        d = oracle.murmur3_x64_128(kat["key"].encode(), kat["seed"])
        words = struct.unpack("<4I", d)
        assert "".join(f"{w:08x}" for w in words) == kat["hex_u32x4"]


def test_calc_hash_kats(oracle, kats):
    for kat in kats["calc_hash"]:
        assert oracle.calc_hash(kat["kmer"].encode()) == int(kat["hash"]), kat["kmer"]


def test_relations_of_reference_unit_tests(oracle):
    # mkmh/mkmh_test.cpp:12-16
    assert oracle.reverse_complement(b"ACTGGCC") == b"GGCCAGT"
    # :43-64
    assert not oracle.canonical(b"ACTGN") and not oracle.canonical(b"ACTG!")
    assert oracle.canonical(b"acgtACGT")
    # :141-147, :158-168 strand invariance
    assert oracle.calc_hash(b"AAAAAAA") == oracle.calc_hash(b"TTTTTTT")
    assert oracle.calc_hash(b"ACTGGCTTGCC") == oracle.calc_hash(b"GGCAAGCCAGT")
    # :180-183
    assert oracle.calc_hash(b"ACGTNTTA") == 0
    # :90-134  calc_hashes == calc_hash per window, and only len-k windows
    h = oracle.calc_hashes(b"ACCCCTG", [4])
    assert len(h) == 3
    assert [int(x) for x in h] == [oracle.calc_hash(b"ACCC"), oracle.calc_hash(b"CCCC"), oracle.calc_hash(b"CCCT")]
    # :186-209 the literal intersection example equals 2
    a = np.array([1, 2, 3, 4], np.uint64)
    b = np.array([2, 3, 5, 6, 7, 8], np.uint64)
    assert oracle.intersection_size(a, b) == 2


def test_multiset_semantics(oracle, kats):
    mh = kats["multiset"]["minhashes"]
    assert [int(x) for x in oracle.minhashes(np.array(mh["in"], np.uint64), mh["s"])] == mh["out"] == [3, 3, 5, 7]
    it = kats["multiset"]["intersection"]
    assert oracle.intersection_size(it["a"], it["b"]) == it["out"] == 3


def test_to_upper_quirk(oracle):
    # mkmh.hpp:231-236: every byte above 91 loses 32, not only letters
    assert oracle.to_upper(b"acgtn{|}~`_^]\\[Z") == bytes((c - 32) if c > 91 else c for c in b"acgtn{|}~`_^]\\[Z")


def test_merge_equals_global_bottom_s(oracle):
    rng = np.random.default_rng(7)
    x = rng.integers(1, 50, size=400, dtype=np.uint64)  # plenty of duplicates
    s = 64
    whole = oracle.minhashes(x, s)
    a, b = oracle.minhashes(x[:150], s), oracle.minhashes(x[150:], s)
    assert np.array_equal(oracle.merge_sketches(a, b, s), whole)


def test_argmax_prefix_diff(oracle):
    # src/rkmh.cpp:874-883
    hit = oracle.argmax([3, 7, 7, 2, 9, 9], 11)
    assert (hit["ref_id"], hit["best"], hit["prev"], hit["sketch_len"]) == (4, 9, 7, 11)
    hit = oracle.argmax([0, 0, 0], 0)
    assert (hit["ref_id"], hit["best"], hit["prev"]) == (0, 0, -1)
    hit = oracle.argmax([5, 1], 0)
    assert (hit["ref_id"], hit["best"], hit["prev"]) == (0, 5, -1)


@pytest.mark.parametrize("case", sorted(LIB_CASES))
def test_against_real_mkmh_vectors(oracle, lib_vectors, case):
    """hashes, sketches and per-read hit records equal those of the real mkmh functions."""
    ref_files, read_file, ks, s = LIB_CASES[case]
    _, rs, _ = load_sequences([data_path(f) for f in ref_files])
    _, qs, _ = load_sequences([data_path(read_file)])
    sk, lens = oracle.sketch_batch(rs, ks, s)
    assert np.array_equal(lens, lib_vectors[case + "/ref_len"])
    x = np.array([sketch_checksum(sk[i, :lens[i]]) for i in range(len(rs))], np.uint64)
    assert np.array_equal(x, lib_vectors[case + "/ref_sketch_x"])
    if case + "/ref_sketches" in lib_vectors:
        assert np.array_equal(sk, lib_vectors[case + "/ref_sketches"])
    for i in range(min(len(rs), 8)):
        h = oracle.calc_hashes(oracle.to_upper(rs[i]), ks)
        with np.errstate(over="ignore"):
            assert np.add.reduce(h, dtype=np.uint64) == lib_vectors[case + "/ref_hsum"][i]
        assert np.bitwise_xor.reduce(h) == lib_vectors[case + "/ref_hxor"][i]
    n = min(len(qs), 300)
    hits = oracle.classify_batch(qs[:n], ks, s, sk, lens)
    gold = lib_vectors[case + "/hits"][:n]
    got = np.stack([hits["ref_id"], hits["best"], hits["prev"], hits["sketch_len"]], axis=1)
    assert np.array_equal(got, gold)


def test_full_hash_dump(oracle, lib_vectors):
    _, qs, _ = load_sequences([data_path("minION25.fq.gz")])
    assert np.array_equal(oracle.calc_hashes(oracle.to_upper(qs[0]), [10]), lib_vectors["dump/minion0_k10"])
    _, rs, _ = load_sequences([data_path("all_pave_ref.fa.gz")])
    h = oracle.calc_hashes(oracle.to_upper(rs[0]), [12])
    assert np.array_equal(h, lib_vectors["dump/pave0_k12"])
    assert np.array_equal(oracle.minhashes(h, 1000), lib_vectors["dump/pave0_k12_sketch"])


def test_count_and_mask_against_real_counter(oracle):
    """-M path: counts and masking equal the real HASHTCounter / mask_by_frequency (needs oracle/_ref)."""
    from oracle.oracle import RefMkmh
    try:
        ref = RefMkmh()
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    _, qs, _ = load_sequences([data_path("minION25.fq.gz")])
    size = 1_000_003
    counts = np.zeros(size, np.int32)
    oracle.count_batch(qs[:6], [12], counts, valid_only=True)
    ctr = ref.lib.ref_counter_new(size)
    allh = []
    for sq in qs[:6]:
        h = ref.calc_hashes(ref.to_upper(sq), [12])
        allh.append(h)
        for v in h[h != 0]:
            ref.lib.ref_counter_increment(ctr, int(v))
    probe = np.concatenate(allh)[::17]
    got = np.array([ref.lib.ref_counter_get(ctr, int(v)) for v in probe], np.int32)
    assert np.array_equal(got, counts[(probe % np.uint64(size)).astype(np.int64)])
    masked_ref = allh[0].copy()
    ref.lib.ref_mask_by_frequency(masked_ref, masked_ref.size, ctr, 2)
    masked = allh[0].copy()
    oracle.lib.rko_mask_by_frequency(masked, masked.size, counts, size, 2)
    assert np.array_equal(masked, masked_ref)
    ref.lib.ref_counter_free(ctr)


# ---- hpv16: the oracle's restatement pinned to the real binary's stdout --------------------------
def test_set_intersection_size_semantics(oracle):
    """mkmh/mkmh.hpp:842-863: distinct non-zero common values of two ascending arrays"""
    f = oracle.set_intersection_size
    assert f([3, 3, 5, 7], [3, 5, 5, 7, 7]) == 3
    assert f([0, 0, 3, 3, 3], [0, 3, 3]) == 1
    assert f([1, 2, 3], [4, 5]) == 0
    assert f([], [1]) == 0
    assert f([5, 5, 7], [5, 7]) == 2
    # the multiset variant (mkmh/mkmh.hpp:822-840) counts min(multiplicity) instead
    assert oracle.intersection_size(np.array([3, 3, 5, 7], np.uint64), np.array([3, 3, 5, 5], np.uint64)) == 3


@pytest.mark.parametrize("name", ["hpv16cmd_minion_k16", "hpv16cmd_minion_k10_k14", "hpv16cmd_z1_k20"])
def test_hpv16_mirror_on_the_oracle_equals_the_real_binary(oracle, cli_manifest, name):
    """`rkmh hpv16` restated on the oracle's functions gives the real binary's stdout and side file byte for byte."""
    import hashlib
    from tests.golden.make_golden import HPV16_CASES
    from tests.hpv16_mirror import OracleEngine, hpv16_stdout
    from tests.util import data_path, golden_cli_output
    argv = HPV16_CASES[name]
    ks = [int(argv[i + 1]) for i, a in enumerate(argv) if a == "-k"]
    reads = [data_path(argv[i + 1]) for i, a in enumerate(argv) if a == "-f"]
    out, side = hpv16_stdout(OracleEngine(oracle), data_path("all_pave_ref.fa.gz"), data_path("new_refs.fa.gz"), reads, ks)
    assert out == golden_cli_output(name)
    assert hashlib.md5(side).hexdigest() == cli_manifest[name]["side_file_md5"]
