#!/usr/bin/env python
"""tests/golden/make_golden.py -- regenerates every committed golden vector.

Run in the BUILD container (needs /root/reference and oracle/_ref built by `make -C oracle`):

    python tests/golden/make_golden.py

It produces, from the REAL reference (never from our own code):
  tests/golden/data/*.gz        the reference's shipped fixtures (reference data/*), gzip-compressed so
                                they can travel to the GPU box, plus two small hand-made awkward files
  tests/golden/cli/*.out.gz     stdout of the real rkmh CLI (oracle/_ref/rkmh_ref, and rkmh_ref_filter
                                for `filter`) for each case in CLI_CASES, with md5s in cli/manifest.json
  tests/golden/kats.json        literal Murmur3 KATs quoted from the reference's own test
                                (mkmh/murmur3/test.c:42-45) and canonical-hash / multiset KATs computed
                                with the real mkmh library (oracle/_ref/libmkmh_ref.so)
  tests/golden/lib_vectors.npz  per-record hash checksums, sketches and per-read hit records computed
                                with the real mkmh functions for several (refs, reads, k, s) cases

The GPU box has no /root/reference: tests read only the files written here.
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.oracle import RefMkmh, ref_binary  # noqa: E402
from rkmh_b200.fastx import load_sequences  # noqa: E402

REF = "/root/reference"
DATA = os.path.join(HERE, "data")
CLI = os.path.join(HERE, "cli")

FIXTURES = ["all_pave_ref.fa", "hpv_16_allFasta.fa", "minION25.fq", "new_refs.fa", "z1.fq", "zika.refs.fa",
            "zika.fa", "dengue.fa", "yellow_fever.fa", "hpv_16.fa"]

# name -> (binary, argv with fixture names relative to tests/golden/data)
CLI_CASES = {
    "cfg1_classify_k10": ("stream", ["classify", "-r", "hpv_16_allFasta.fa.gz", "-f", "minION25.fq.gz", "-k", "10", "-s", "1000"]),
    "pave_minion_k12": ("stream", ["stream", "-r", "all_pave_ref.fa.gz", "-f", "minION25.fq.gz", "-k", "12", "-s", "1000"]),
    "pave_z1_k12": ("stream", ["stream", "-r", "all_pave_ref.fa.gz", "-f", "z1.fq.gz", "-k", "12", "-s", "1000"]),
    "hpv16_minion_M100_N10": ("stream", ["classify", "-r", "hpv_16_allFasta.fa.gz", "-f", "minION25.fq.gz", "-k", "12", "-s", "1000", "-M", "100", "-N", "10"]),
    "hpv16_minion_M2_N10": ("stream", ["classify", "-r", "hpv_16_allFasta.fa.gz", "-f", "minION25.fq.gz", "-k", "12", "-s", "1000", "-M", "2", "-N", "10"]),
    "hpv16_minion_M2_N10_D5": ("stream", ["classify", "-r", "hpv_16_allFasta.fa.gz", "-f", "minION25.fq.gz", "-k", "12", "-s", "1000", "-M", "2", "-N", "10", "-D", "5"]),
    "pave_minion_I20": ("stream", ["stream", "-r", "all_pave_ref.fa.gz", "-f", "minION25.fq.gz", "-k", "12", "-s", "1000", "-I", "20"]),
    "hpv16_minion_multik": ("stream", ["stream", "-r", "hpv_16_allFasta.fa.gz", "-f", "minION25.fq.gz", "-k", "10", "-k", "12", "-s", "1000"]),
    "viral4_z1_k20": ("stream", ["stream", "-r", "zika.fa.gz", "-r", "dengue.fa.gz", "-r", "yellow_fever.fa.gz", "-r", "hpv_16.fa.gz", "-f", "z1.fq.gz", "-k", "20", "-s", "2000"]),
    "zikarefs_z1long_k16_N5": ("stream", ["stream", "-r", "zika.refs.fa.gz", "-f", "z1_long_200.fq.gz", "-k", "16", "-s", "1000", "-N", "5"]),
    "newrefs_mix2k_M20_N10": ("stream", ["classify", "-r", "new_refs.fa.gz", "-f", "mix2k.fq.gz", "-k", "12", "-s", "1000", "-M", "20", "-N", "10"]),
    "pave_mix2k_k12_s100": ("stream", ["stream", "-r", "all_pave_ref.fa.gz", "-f", "mix2k.fq.gz", "-k", "12", "-s", "100", "-N", "3", "-D", "1"]),
    "awkward_k8": ("stream", ["stream", "-r", "awkward_refs.fa.gz", "-f", "awkward_reads.fq.gz", "-k", "8", "-s", "50"]),
    "filter_viral4_z1": ("filter", ["filter", "-r", "zika.fa.gz", "-r", "dengue.fa.gz", "-r", "yellow_fever.fa.gz", "-r", "hpv_16.fa.gz", "-f", "z1.fq.gz", "-k", "20", "-s", "2000"]),
    "filter_zika_z1_N5": ("filter", ["filter", "-r", "zika.fa.gz", "-f", "z1.fq.gz", "-k", "20", "-s", "2000", "-N", "5"]),
    "filter_hpv16_minion_N50": ("filter", ["filter", "-r", "hpv_16.fa.gz", "-f", "minION25.fq.gz", "-k", "12", "-s", "1000", "-N", "50"]),
    "filter_dengue_z1": ("filter", ["filter", "-r", "dengue.fa.gz", "-f", "z1.fq.gz", "-k", "20", "-s", "2000"]),
    "pave_two_read_files_k12": ("stream", ["stream", "-r", "all_pave_ref.fa.gz", "-f", "z1.fq.gz", "-f", "mix2k.fq.gz", "-k", "12", "-s", "1000"]),
    "pave_fasta_reads_k12": ("stream", ["stream", "-r", "all_pave_ref.fa.gz", "-f", "hpv_16_allFasta.fa.gz", "-k", "12", "-s", "1000"]),
    "two_ref_files_multik_M2": ("stream", ["classify", "-r", "hpv_16_allFasta.fa.gz", "-r", "new_refs.fa.gz", "-f", "minION25.fq.gz", "-k", "10", "-k", "14", "-s", "500", "-M", "2", "-N", "5", "-D", "2"]),
    "filter_viral4_z1_M2_D2": ("filter", ["filter", "-r", "zika.fa.gz", "-r", "dengue.fa.gz", "-r", "yellow_fever.fa.gz", "-r", "hpv_16.fa.gz", "-f", "z1.fq.gz", "-k", "20", "-s", "2000", "-M", "2", "-D", "2"]),
}

# `rkmh hpv16` (src/rkmh.cpp:2366-2723): name -> argv without -R (the generator and the tests add it; the command reads
# <refpath>/all_pave_ref.fa and <refpath>/new_refs.fa and writes lineage_specific_hashes.<k>.tst into its working directory)
HPV16_CASES = {
    "hpv16cmd_minion_k16": ["hpv16", "-f", "minION25.fq.gz", "-k", "16"],
    "hpv16cmd_mix2k_k12": ["hpv16", "-f", "mix2k.fq.gz", "-k", "12"],
    "hpv16cmd_minion_k10_k14": ["hpv16", "-f", "minION25.fq.gz", "-k", "10", "-k", "14"],
    "hpv16cmd_mix2k_k16_M2": ["hpv16", "-f", "mix2k.fq.gz", "-k", "16", "-M", "2"],
    "hpv16cmd_z1_k20": ["hpv16", "-f", "z1.fq.gz", "-k", "20"],
    "hpv16cmd_two_files_k14_M3": ["hpv16", "-f", "minION25.fq.gz", "-f", "mix2k.fq.gz", "-k", "14", "-M", "3"],
}

# `rkmh call` (src/rkmh.cpp:1455-1904): VCF of k-mer-depth calls; argv relative to tests/golden/data like CLI_CASES
CALL_CASES = {
    "call_hpv16_minion_k12": ["call", "-r", "hpv_16.fa.gz", "-f", "minION25.fq.gz", "-k", "12"],
    "call_hpv16_mix2k_k16": ["call", "-r", "hpv_16.fa.gz", "-f", "mix2k.fq.gz", "-k", "16"],
    "call_zika_z1_k20_w50": ["call", "-r", "zika.fa.gz", "-f", "z1.fq.gz", "-k", "20", "-w", "50"],
    "call_newrefs_mix2k_k14": ["call", "-r", "new_refs.fa.gz", "-f", "mix2k.fq.gz", "-f", "minION25.fq.gz", "-k", "14"],
    "call_zika_z1_k9_w10": ["call", "-r", "zika.fa.gz", "-f", "z1.fq.gz", "-k", "9", "-w", "10"],
    "call_show_depth": ["call", "-r", "hpv_16.fa.gz", "-f", "minION25.fq.gz", "-k", "12", "-d"],
}

# library-level cases: name -> (ref fixtures, read fixture, k list, s)
LIB_CASES = {
    "hpv16_minion_k10": (["hpv_16_allFasta.fa.gz"], "minION25.fq.gz", [10], 1000),
    "hpv16_minion_k10k12": (["hpv_16_allFasta.fa.gz"], "minION25.fq.gz", [10, 12], 1000),
    "pave_z1_k12": (["all_pave_ref.fa.gz"], "z1.fq.gz", [12], 1000),
    "pave_minion_k12": (["all_pave_ref.fa.gz"], "minION25.fq.gz", [12], 1000),
    "viral4_z1_k20": (["zika.fa.gz", "dengue.fa.gz", "yellow_fever.fa.gz", "hpv_16.fa.gz"], "z1.fq.gz", [20], 2000),
    "zikarefs_z1long_k16": (["zika.refs.fa.gz"], "z1_long_200.fq.gz", [16], 1000),
    "pave_mix2k_k12_s100": (["all_pave_ref.fa.gz"], "mix2k.fq.gz", [12], 100),
    "awkward_k8": (["awkward_refs.fa.gz"], "awkward_reads.fq.gz", [8], 50),
}


def gz_write(path, payload: bytes):
    with open(path, "wb") as raw:  # mtime=0 keeps the files reproducible
        with gzip.GzipFile(fileobj=raw, mode="wb", compresslevel=9, mtime=0) as fh:
            fh.write(payload)


def make_data():
    os.makedirs(DATA, exist_ok=True)
    for f in FIXTURES:
        with open(os.path.join(REF, "data", f), "rb") as fh:
            gz_write(os.path.join(DATA, f + ".gz"), fh.read())
    # first 200 records of z1_long.fq (1500 bp wgsim reads)
    with open(os.path.join(REF, "data", "z1_long.fq"), "rb") as fh:
        lines = fh.read().split(b"\n")
    gz_write(os.path.join(DATA, "z1_long_200.fq.gz"), b"\n".join(lines[:800]) + b"\n")
    # first 2000 records of the simulated HPV16 co-infection mix
    big = [f for f in os.listdir(os.path.join(REF, "data")) if f.startswith("1088_")][0]
    with gzip.open(os.path.join(REF, "data", big), "rb") as fh:
        lines = []
        for _ in range(8000):
            lines.append(fh.readline())
    gz_write(os.path.join(DATA, "mix2k.fq.gz"), b"".join(lines))
    # hand-made awkward inputs: multi-line FASTA, lower case, IUPAC codes, blank lines, CRLF, comments,
    # homopolymers and exact repeats (multiset sketches), a palindromic k-mer, reads barely longer than k
    refs = (b">refA first reference\nACGTACGTACGTTTGACCAGTAGGACCATTAGACCAGGTTTTAAACCCGGGTTTAAACC\nacgtacgtaggatcca\n\n"
            b">refB\tdup-heavy\nAAAAAAAAAAAAAAAAAAAAAAAAAAAAAACCCCCCCCCCCCCCCCCCCCGGGATCCGGGATCCGGGATCCGGGATCC\n"
            b">refC iupac\r\nACGTNACGTRYACGTACGGATTACAGATTACAGATTACAcagattacagattaca\r\nGGGTTTAAACCC\r\n"
            b">refD\nTTGACCAGTAGGACCATTAGACCAGGTTTTAAACCCGGGTTTAAACCACGTACGTAC\n")
    reads = (b"@r1 comment\nACGTACGTACGTTTGACCAGTAGGACCATTAGACC\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n"
             b"@r2\nAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA\n+r2\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n"
             b"@r3\nGGGATCCGGGATCCGGGATCCGGGATCCGGGATCCGGGATCC\n+\n@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@@\n"
             b"@r4\nacgtnnnnacgtacgtaggatccaNNNNNNNNNNNN\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n"
             b"@r5\nACGTACGTA\n+\nIIIIIIIII\n"
             b"@r6\nNNNNNNNNNNNNNNNNNNNNNNNNN\n+\nIIIIIIIIIIIIIIIIIIIIIIIII\n"
             b"@r7\nGATTACAGATTACAGATTACAGATTACAGATTACATTGACCAGTAGG\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n"
             b"@r8\nGGTTTAAACCCGGGTTTAAACCCGGGTTTAAACC\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n")
    gz_write(os.path.join(DATA, "awkward_refs.fa.gz"), refs)
    gz_write(os.path.join(DATA, "awkward_reads.fq.gz"), reads)


def make_cli():
    os.makedirs(CLI, exist_ok=True)
    manifest = {}
    for name, (which, argv) in CLI_CASES.items():
        exe = ref_binary(filter_build=(which == "filter"))
        assert exe, "build oracle/_ref first (make -C oracle)"
        out = subprocess.run([exe] + argv + ["-t", "1"], cwd=DATA, stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, check=True).stdout
        gz_write(os.path.join(CLI, name + ".out.gz"), out)
        manifest[name] = {"argv": argv, "md5": hashlib.md5(out).hexdigest(),
                          "md5_sorted": hashlib.md5(b"".join(sorted(out.splitlines(True)))).hexdigest(),
                          "lines": out.count(b"\n"), "binary": "rkmh_ref_filter" if which == "filter" else "rkmh_ref"}
        print(f"cli  {name:28s} {manifest[name]['lines']:6d} lines  md5 {manifest[name]['md5']}")
    with open(os.path.join(CLI, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)


def make_call(manifest_path=None):
    manifest_path = manifest_path or os.path.join(CLI, "manifest.json")
    with open(manifest_path) as fh:
        manifest = json.load(fh)
    exe = ref_binary()
    assert exe, "build oracle/_ref first (make -C oracle)"
    for name, argv in CALL_CASES.items():
        out = subprocess.run([exe] + argv + ["-t", "1"], cwd=DATA, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
        gz_write(os.path.join(CLI, name + ".out.gz"), out)
        manifest[name] = {"argv": argv, "md5": hashlib.md5(out).hexdigest(), "lines": out.count(b"\n"), "binary": "rkmh_ref"}
        print(f"cli  {name:28s} {manifest[name]['lines']:6d} lines  md5 {manifest[name]['md5']}")
    with open(manifest_path, "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)


def make_hpv16(manifest_path=None):
    """stdout (and the side file) of the real binary's `hpv16` command.  It wants plain file NAMES under -R; gzopen reads
    gzip content whatever the name, so a scratch directory of symlinks onto the committed .gz fixtures serves."""
    import tempfile
    manifest_path = manifest_path or os.path.join(CLI, "manifest.json")
    with open(manifest_path) as fh:
        manifest = json.load(fh)
    exe = ref_binary()
    assert exe, "build oracle/_ref first (make -C oracle)"
    with tempfile.TemporaryDirectory() as tmp:
        for f in ("all_pave_ref.fa", "new_refs.fa"):
            os.symlink(os.path.join(DATA, f + ".gz"), os.path.join(tmp, f))
        for name, argv in HPV16_CASES.items():
            full = [exe] + [os.path.join(DATA, a) if a.endswith(".gz") else a for a in argv] + ["-R", tmp, "-t", "1"]
            res = subprocess.run(full, cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
            k0 = argv[argv.index("-k") + 1]
            with open(os.path.join(tmp, f"lineage_specific_hashes.{k0}.tst"), "rb") as fh:
                side = fh.read()
            gz_write(os.path.join(CLI, name + ".out.gz"), res.stdout)
            manifest[name] = {"argv": argv, "md5": hashlib.md5(res.stdout).hexdigest(), "lines": res.stdout.count(b"\n"),
                              "binary": "rkmh_ref", "side_file_md5": hashlib.md5(side).hexdigest(),
                              "stderr_md5": hashlib.md5(res.stderr).hexdigest()}
            print(f"cli  {name:28s} {manifest[name]['lines']:6d} lines  md5 {manifest[name]['md5']}")
    with open(manifest_path, "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)


def sketch_checksum(m):
    idx = np.arange(1, len(m) + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return int(np.bitwise_xor.reduce(m * idx)) if len(m) else 0


def make_kats(ref: RefMkmh):
    kats = {
        # quoted from the reference's own known-answer test, mkmh/murmur3/test.c:42-45 (hex of uint32[4])
        "murmur3_x64_128": [
            {"key": "Hello, world!", "seed": 123, "hex_u32x4": "8743acad421c8c73d373c3f5f19732fd"},
            {"key": "Hello, world!", "seed": 321, "hex_u32x4": "f86d4004ca47f42bb9546c7979200aee"},
            {"key": "x" * 28, "seed": 123, "hex_u32x4": "becf7e04dbcf74637751664ef66e73e0"},
            {"key": "", "seed": 123, "hex_u32x4": "4cd9597081679d1abd92f8784bace33d"},
        ],
        "calc_hash": [],
        "multiset": {
            "minhashes": {"in": [5, 0, 3, 3, 9, 0, 7], "s": 4},
            "intersection": {"a": [3, 3, 5, 7], "b": [3, 5, 5, 7, 7]},
        },
    }
    kmers = ["ACCC", "CCCC", "CCCT", "AAAAAA", "TTTTTT", "ACTGGCTTGCC", "GGCAAGCCAGT", "ACGTNTTA", "ACGTACGTAC",
             "ACTACAATAATT", "GTTAACTACCAT", "ACTACAATAATTCATG", "ACTACAATAATTCATGTATA", "ACTACAATAATTCATGTATAA",
             "ACTACAATAATTCATGTATAAAACTAAGGGCG", "ACGT", "AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA",
             "GATTACAGATTACAGATTACAGATTACAGATTACAGATTACAGATTACA", "ACGTACGTACGTACGTA", "ACGTACGTACGTACGTACGTACGTACGTACGTAC"]
    for km in kmers:
        # calc_hashes hashes len-k windows, so append one base to get exactly the one window we want
        h = ref.calc_hashes((km + "A").encode(), [len(km)])
        kats["calc_hash"].append({"kmer": km, "hash": str(int(h[0]))})
    mh = ref.minhashes(np.array(kats["multiset"]["minhashes"]["in"], np.uint64), 4)
    kats["multiset"]["minhashes"]["out"] = [int(x) for x in mh]
    kats["multiset"]["intersection"]["out"] = ref.intersection_size(
        np.array(kats["multiset"]["intersection"]["a"], np.uint64),
        np.array(kats["multiset"]["intersection"]["b"], np.uint64))
    # the reference unit test's literal example (mkmh/mkmh_test.cpp:186-209) expects 2
    with open(os.path.join(HERE, "kats.json"), "w") as fh:
        json.dump(kats, fh, indent=1)


def make_lib_vectors(ref: RefMkmh):
    out = {}
    for name, (ref_files, read_file, ks, s) in LIB_CASES.items():
        rn, rs, _ = load_sequences([os.path.join(DATA, f) for f in ref_files])
        qn, qs, _ = load_sequences([os.path.join(DATA, read_file)])
        rs = [ref.to_upper(x) for x in rs]
        qs = [ref.to_upper(x) for x in qs]
        ref_sk = np.zeros((len(rs), s), np.uint64)
        ref_len = np.zeros(len(rs), np.int32)
        ref_hsum = np.zeros(len(rs), np.uint64)
        ref_hxor = np.zeros(len(rs), np.uint64)
        for i, sq in enumerate(rs):
            h = ref.calc_hashes(sq, ks)
            with np.errstate(over="ignore"):
                ref_hsum[i] = np.add.reduce(h, dtype=np.uint64)
            ref_hxor[i] = np.bitwise_xor.reduce(h)
            m = ref.minhashes(h, s)
            ref_sk[i, :len(m)] = m
            ref_len[i] = len(m)
        hits = np.zeros((len(qs), 4), np.int32)
        q_hsum = np.zeros(len(qs), np.uint64)
        q_sx = np.zeros(len(qs), np.uint64)
        for i, sq in enumerate(qs):
            h = ref.calc_hashes(sq, ks)
            with np.errstate(over="ignore"):
                q_hsum[i] = np.add.reduce(h, dtype=np.uint64)
            m = ref.minhashes(h, s)
            q_sx[i] = sketch_checksum(m)
            best, rid, prev = -1, 0, -1
            for j in range(len(rs)):
                c = ref.intersection_size(m, ref_sk[j, :ref_len[j]]) if len(m) and ref_len[j] else 0
                if c > best:
                    prev, best, rid = best, c, j
            hits[i] = (rid, best, prev, len(m))
        out[name + "/ref_len"] = ref_len
        out[name + "/ref_hsum"] = ref_hsum
        out[name + "/ref_hxor"] = ref_hxor
        out[name + "/ref_sketch_x"] = np.array([sketch_checksum(ref_sk[i, :ref_len[i]]) for i in range(len(rs))], np.uint64)
        if len(rs) <= 16:
            out[name + "/ref_sketches"] = ref_sk
        out[name + "/hits"] = hits
        out[name + "/read_hsum"] = q_hsum
        out[name + "/read_sketch_x"] = q_sx
        print(f"lib  {name:28s} refs {len(rs):4d} reads {len(qs):5d}  best-sum {int(hits[:, 1].sum())}")
    # one full hash dump for direct comparison: first minION read, k=10 and first pave ref, k=12
    qn, qs, _ = load_sequences([os.path.join(DATA, "minION25.fq.gz")])
    out["dump/minion0_k10"] = ref.calc_hashes(ref.to_upper(qs[0]), [10])
    rn, rs, _ = load_sequences([os.path.join(DATA, "all_pave_ref.fa.gz")])
    out["dump/pave0_k12"] = ref.calc_hashes(ref.to_upper(rs[0]), [12])
    out["dump/pave0_k12_sketch"] = ref.minhashes(out["dump/pave0_k12"], 1000)
    np.savez_compressed(os.path.join(HERE, "lib_vectors.npz"), **out)


if __name__ == "__main__":
    if sys.argv[1:] == ["hpv16"]:               # only (re)generate the hpv16 command's goldens
        make_hpv16()
        sys.exit(0)
    if sys.argv[1:] == ["call"]:
        make_call()
        sys.exit(0)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    ref = RefMkmh()
    make_data()
    make_cli()
    make_hpv16()
    make_call()
    make_kats(ref)
    make_lib_vectors(ref)
    print("golden vectors written under", HERE)
