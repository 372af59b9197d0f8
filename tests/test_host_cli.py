"""The C++ `rkmh` front-end: its FASTA/FASTQ parser (CPU) and, on the GPU box, its stdout against the real
reference CLI's golden outputs."""
import gzip
import hashlib
import os
import subprocess

import pytest

from rkmh_b200.fastx import read_fastx
from tests.golden.make_golden import CLI_CASES
from tests.util import DATA, ROOT, golden_cli_output

BIN = os.path.join(ROOT, "rkmh_b200", "bin", "rkmh")


@pytest.fixture(scope="module")
def rkmh_bin():
    if not os.path.exists(BIN):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "rkmh_b200", "host")], stdout=subprocess.DEVNULL)
    return BIN


def _dump(path):
    return b"".join(n + b"\t" + s + b"\t" + q + b"\n" for n, s, q in read_fastx(path))


@pytest.mark.parametrize("name", ["awkward_refs.fa.gz", "awkward_reads.fq.gz", "minION25.fq.gz", "hpv_16_allFasta.fa.gz",
                                  "z1.fq.gz", "zika.refs.fa.gz"])
def test_cpp_parser_matches_python_parser(rkmh_bin, name):
    path = os.path.join(DATA, name)
    out = subprocess.run([rkmh_bin, "parse", path], stdout=subprocess.PIPE, check=True).stdout
    assert out == _dump(path)


def test_parser_odd_inputs(rkmh_bin, tmp_path):
    """header without newline, '>' inside a sequence line, CRLF, truncated quality, empty file, no final newline"""
    cases = {
        "a.fa": b">s1 desc\nACGT\nAC>s2\nGG\r\nTT\r\n>s3",
        "b.fq": b"@r1\nACGT\n+\nIIII\n@r2\nAC\nGT\n+r2\nII\nII\n@r3\nACGTA\n+\nIII",
        "c.fa": b"",
        "d.fa": b"junk before\n>x\nacgtn\n\n\n>y\tcomment\n\nAC GT\n",
        "e.fq": b"@q\nACGT\n+\n@@@@\n@q2\nAC\n+\n@I",
        # a header byte as the LAST byte of the input starts no record: kseq_read returns -1 there (the real reference
        # binary prints two lines for g.fq's shape, checked by hand against oracle/_ref/rkmh_ref); FASTA likewise
        "g.fq": b"@r1\nACGT\n+\nIIII\n@r2\nGGCA\n+\nIIII\n@",
        "h.fa": b">s1\nACGT\n>",
        "i.fa": b">",
    }
    for fn, payload in cases.items():
        p = tmp_path / fn
        p.write_bytes(payload)
        out = subprocess.run([rkmh_bin, "parse", str(p)], stdout=subprocess.PIPE, check=True).stdout
        assert out == _dump(str(p)), fn
        if fn in ("g.fq", "h.fa", "i.fa"):
            assert out.count(b"\n") == {"g.fq": 2, "h.fa": 1, "i.fa": 0}[fn]
        pz = tmp_path / (fn + ".gz")
        with gzip.open(pz, "wb") as fh:
            fh.write(payload)
        outz = subprocess.run([rkmh_bin, "parse", str(pz)], stdout=subprocess.PIPE, check=True).stdout
        assert outz == out, fn


def _bgzf(payload, block=60000, eof=True, level=6):
    """`payload` as blocked gzip (BGZF, what bgzip / samtools write): members of at most `block` input bytes, each with its
    compressed size in a 'BC' extra field, and the empty end-of-file member."""
    import struct
    import zlib
    out = bytearray()
    chunks = [payload[i:i + block] for i in range(0, len(payload), block)] + ([b""] if eof else [])
    for c in chunks:
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        data = co.compress(c) + co.flush()
        bsize = 12 + 6 + len(data) + 8
        out += struct.pack("<4BI2BH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
        out += data + struct.pack("<II", zlib.crc32(c) & 0xFFFFFFFF, len(c) & 0xFFFFFFFF)
    return bytes(out)


def test_bgzf_input_is_inflated_in_parallel_and_reads_the_same(rkmh_bin, tmp_path):
    """Blocked gzip (BGZF) input is recognised by its first member and inflated by several threads, member batches in
    order; every other gzip file still goes through zlib's single stream.  The records must be those of the plain file:
    members that end inside records, tiny members, one thread, a tiny parse buffer, a file without the end-of-file member,
    a file truncated inside a member (the records before the damage, like zlib), and RKMH_NO_BGZF=1 (zlib reads BGZF too)."""
    import gzip
    import random
    rnd = random.Random(5)
    recs = []
    for i in range(3000):
        L = rnd.choice([30, 150, 150, 700, 9000])
        recs.append(f"@r{i} x\n{''.join(rnd.choice('ACGTN') for _ in range(L))}\n+\n{''.join(chr(rnd.randint(33, 126)) for _ in range(L))}\n")
    payload = "".join(recs).encode()
    plain = tmp_path / "p.fq"
    plain.write_bytes(payload)
    exp = subprocess.run([rkmh_bin, "parse", str(plain)], stdout=subprocess.PIPE, check=True).stdout
    assert exp.count(b"\n") == 3000
    # Python's gzip module reads multi-member files: the helper above writes valid gzip
    assert gzip.decompress(_bgzf(payload, 777)) == payload
    for name, blob, envs in (("b60k", _bgzf(payload), [{}, {"RKMH_PARSE_THREADS": "1"}, {"RKMH_PARSE_BUF": "100"}, {"RKMH_NO_BGZF": "1"}]),
                             ("b1k", _bgzf(payload, 1000, level=1), [{"RKMH_PARSE_THREADS": "5"}]),
                             ("noeof", _bgzf(payload, 50000, eof=False), [{}]),
                             ("gz", gzip.compress(payload), [{}])):
        f = tmp_path / (name + ".fq.gz")
        f.write_bytes(blob)
        for env in envs:
            out = subprocess.run([rkmh_bin, "parse", str(f)], stdout=subprocess.PIPE, check=True, env=dict(os.environ, **env)).stdout
            assert out == exp, (name, env)
        out = subprocess.run([rkmh_bin, "parse-stream", str(f)], stdout=subprocess.PIPE, check=True,
                             env=dict(os.environ, RKMH_CHUNK_BYTES="20000")).stdout
        assert out == exp, name
    # a FIFO is never probed for BGZF (the probe would eat its first bytes): zlib reads it, BGZF or not
    import threading
    fifo = tmp_path / "in.fifo"
    os.mkfifo(fifo)
    blob = _bgzf(payload)
    t = threading.Thread(target=lambda: open(fifo, "wb").write(blob), daemon=True)
    t.start()
    out = subprocess.run([rkmh_bin, "parse", str(fifo)], stdout=subprocess.PIPE, check=True, timeout=60).stdout
    t.join(timeout=10)
    assert out == exp
    cut = tmp_path / "cut.fq.gz"
    cut.write_bytes(blob[:len(blob) * 2 // 3])
    got = subprocess.run([rkmh_bin, "parse", str(cut)], stdout=subprocess.PIPE, check=True).stdout
    ref = subprocess.run([rkmh_bin, "parse", str(cut)], stdout=subprocess.PIPE, check=True, env=dict(os.environ, RKMH_NO_BGZF="1")).stdout
    lines, rlines = got.split(b"\n"), ref.split(b"\n")
    assert 1000 < len(lines) < 3000 and exp.startswith(b"\n".join(lines[:-2]))
    # zlib stops at the same damaged member, give or take the partial member it may have delivered
    assert abs(len(lines) - len(rlines)) < 400


def test_parser_fast_path_fuzz(rkmh_bin, tmp_path):
    """Well-formed four-line FASTQ takes a memchr fast path; records that deviate in any way fall back to the kseq
    state machine.  Both must agree with the Python restatement of the state machine on mutated inputs, also when the
    input buffer is tiny (records straddling refills, records longer than the buffer)."""
    import random
    rnd = random.Random(11)

    def record(i):
        L = rnd.choice([1, 2, 30, 150, 150, 150, 700, 5000])
        seq = "".join(rnd.choice("ACGTacgtN") for _ in range(L))
        qual = "".join(chr(rnd.randint(33, 126)) for _ in range(L))
        return f"@r{i} c{i}\n{seq}\n+\n{qual}\n"

    def mutate(rec):
        t = rnd.random()
        if t < 0.70:
            return rec
        ops = [lambda r: r.replace("\n", "\r\n"), lambda r: r[:len(r) // 2] + "\n" + r[len(r) // 2:],
               lambda r: r.replace("+\n", "+ again\n", 1), lambda r: r[:-1], lambda r: "\n" + r, lambda r: r + "\n",
               lambda r: r[:10] + "@" + r[10:], lambda r: r[:12] + "+" + r[12:], lambda r: r[:14] + ">" + r[14:],
               lambda r: r[:9] + " " + r[9:], lambda r: r[:11] + "1*-." + r[11:], lambda r: ">fa" + r[1:r.index("\n+")] + "\n",
               lambda r: r[:-3], lambda r: r.replace("\n", "\n\n", 1), lambda r: r[:r.rindex("\n", 0, -1)] + " \n"]
        return rnd.choice(ops)(rec)

    for case in range(12):
        n = rnd.choice([1, 3, 40, 300])
        payload = "".join(mutate(record(i)) for i in range(n)).encode()
        if case % 4 == 3:
            payload = payload[:rnd.randint(0, len(payload))]             # truncated anywhere
        f = tmp_path / f"fuzz{case}.fq"
        f.write_bytes(payload)
        exp = _dump(str(f))
        for buf in ("64", "1000", "4194304"):
            out = subprocess.run([rkmh_bin, "parse", str(f)], stdout=subprocess.PIPE, check=True,
                                 env=dict(os.environ, RKMH_PARSE_BUF=buf)).stdout
            assert out == exp, (case, buf)
        # the streaming front-end's chunking with the parallel segment parser (tiny segments and chunks: segments that
        # do not chain, records that are not well-formed and records larger than a chunk all hand over to the state machine)
        for seg, chunk, thr in (("300", "2000", "4"), ("5000", "1500", "3"), ("700", "100000", "8")):
            env = dict(os.environ, RKMH_PARSE_SEG=seg, RKMH_CHUNK_BYTES=chunk, RKMH_PARSE_THREADS=thr)
            out = subprocess.run([rkmh_bin, "parse-stream", str(f)], stdout=subprocess.PIPE, check=True, env=env, timeout=60).stdout
            assert out == exp, (case, seg, chunk, thr)


def test_parallel_parser_on_fixtures(rkmh_bin, tmp_path):
    """plain-text copies of the FASTQ fixtures through the parallel segment parser == the sequential parser"""
    for name in ("z1.fq.gz", "mix2k.fq.gz", "minION25.fq.gz", "awkward_reads.fq.gz"):
        plain = tmp_path / name[:-3]
        plain.write_bytes(gzip.open(os.path.join(DATA, name), "rb").read())
        exp = subprocess.run([rkmh_bin, "parse", str(plain)], stdout=subprocess.PIPE, check=True).stdout
        for seg, chunk in (("4096", "30000"), ("20000", "8000"), ("100000", "1000000")):
            env = dict(os.environ, RKMH_PARSE_SEG=seg, RKMH_CHUNK_BYTES=chunk, RKMH_PARSE_THREADS="4")
            out = subprocess.run([rkmh_bin, "parse-stream", str(plain), os.path.join(DATA, name)], stdout=subprocess.PIPE,
                                 check=True, env=env, timeout=120).stdout
            assert out == exp + exp, (name, seg, chunk)


def test_usage_and_exit_codes(rkmh_bin):
    assert subprocess.run([rkmh_bin], stderr=subprocess.PIPE).returncode == 1
    assert subprocess.run([rkmh_bin, "stream"], stderr=subprocess.PIPE).returncode == 1
    assert subprocess.run([rkmh_bin, "stream", "-h"], stderr=subprocess.PIPE).returncode == 1
    # an option letter without a case aborts, as the reference does (src/rkmh.cpp:716-718)
    assert subprocess.run([rkmh_bin, "stream", "-d"], stderr=subprocess.PIPE).returncode < 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CLI_CASES))
def test_cli_stdout_equals_reference(rkmh_bin, cli_manifest, name):
    _, argv = CLI_CASES[name]
    res = subprocess.run([rkmh_bin] + argv + ["-t", "1"], cwd=DATA, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert res.returncode == 0, res.stderr.decode()[-400:]
    gold = golden_cli_output(name)
    if res.stdout != gold:
        a, b = res.stdout.splitlines(), gold.splitlines()
        diff = [(i, x, y) for i, (x, y) in enumerate(zip(a, b)) if x != y][:3]
        pytest.fail(f"{name}: {len(a)} vs {len(b)} lines; first diffs {diff}")
    assert hashlib.md5(res.stdout).hexdigest() == cli_manifest[name]["md5"]


@pytest.mark.gpu
def test_hash_command(rkmh_bin, oracle):
    """`rkmh hash -s` prints name + sorted sketch; without -s every hash in window order (HEAD's behaviour)."""
    import numpy as np
    from rkmh_b200.fastx import load_sequences
    path = os.path.join(DATA, "new_refs.fa.gz")
    names, seqs, _ = load_sequences([path])
    res = subprocess.run([rkmh_bin, "hash", "-f", path, "-k", "21", "-s", "500"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    lines = res.stdout.splitlines()
    assert len(lines) == len(names)
    for ln, nm, sq in zip(lines, names, seqs):
        f = ln.split(b"\t")
        assert f[0] == nm
        exp = oracle.minhashes(oracle.calc_hashes(oracle.to_upper(sq), [21]), 500)
        assert np.array_equal(np.array([int(x) for x in f[1:]], np.uint64), exp)
    res = subprocess.run([rkmh_bin, "hash", "-f", path, "-k", "12"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    f = res.stdout.splitlines()[0].split(b"\t")
    exp = oracle.calc_hashes(oracle.to_upper(seqs[0]), [12])
    assert np.array_equal(np.array([int(x) for x in f[1:]], np.uint64), exp)


@pytest.mark.gpu
def test_two_gpus_same_output(rkmh_bin, cli_manifest):
    """read sharding over 2 GPUs (incl. the NCCL all-reduce of the -M table) does not change stdout"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    for name in ("pave_z1_k12", "newrefs_mix2k_M20_N10"):
        _, argv = CLI_CASES[name]
        res = subprocess.run([rkmh_bin] + argv + ["-g", "2"], cwd=DATA, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert res.returncode == 0, res.stderr.decode()[-400:]
        assert hashlib.md5(res.stdout).hexdigest() == cli_manifest[name]["md5"]
        # many small chunks finishing out of order on two workers: the writer still emits them in input order
        env = dict(os.environ, RKMH_CHUNK_BYTES="20000")
        res = subprocess.run([rkmh_bin] + argv + ["-g", "2"], cwd=DATA, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
        assert res.returncode == 0, res.stderr.decode()[-400:]
        assert hashlib.md5(res.stdout).hexdigest() == cli_manifest[name]["md5"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["pave_mix2k_k12_s100", "newrefs_mix2k_M20_N10", "filter_viral4_z1_M2_D2", "pave_minion_k12",
                                  "awkward_k8"])
def test_streaming_chunks_do_not_change_stdout(rkmh_bin, cli_manifest, name):
    """The reads are streamed through small pinned chunks (here 20 kB: hundreds of chunks, records larger than a
    chunk, two passes over the files with -M) and written in input order: stdout is the reference's, byte for byte."""
    _, argv = CLI_CASES[name]
    env = dict(os.environ, RKMH_CHUNK_BYTES="20000")
    res = subprocess.run([rkmh_bin] + argv, cwd=DATA, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    assert res.returncode == 0, res.stderr.decode()[-400:]
    assert hashlib.md5(res.stdout).hexdigest() == cli_manifest[name]["md5"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["pave_z1_k12", "filter_zika_z1_N5", "hpv16_minion_M2_N10"])
def test_reads_from_standard_input(rkmh_bin, cli_manifest, name):
    """`-f -` streams the reads from standard input (gzip or plain); with -M they are buffered for the second pass."""
    _, argv = CLI_CASES[name]
    i = argv.index("-f")
    reads = os.path.join(DATA, argv[i + 1])
    argv = argv[:i + 1] + ["-"] + argv[i + 2:]
    with open(reads, "rb") as fh:
        res = subprocess.run([rkmh_bin] + argv, cwd=DATA, stdin=fh, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert res.returncode == 0, res.stderr.decode()[-400:]
    assert hashlib.md5(res.stdout).hexdigest() == cli_manifest[name]["md5"]
    plain = gzip.open(reads, "rb").read()
    res = subprocess.run([rkmh_bin] + argv, cwd=DATA, input=plain, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert res.returncode == 0 and hashlib.md5(res.stdout).hexdigest() == cli_manifest[name]["md5"]


@pytest.mark.gpu
def test_hash_command_sharded_records(rkmh_bin):
    """Records above RKMH_SHARD_MIN go through rkg_sketch_shard (split over the GPUs with -g, partial sketches merged):
    stdout does not change."""
    import torch
    path = os.path.join(DATA, "hpv_16_allFasta.fa.gz")
    base = subprocess.run([rkmh_bin, "hash", "-f", path, "-k", "21", "-s", "300"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout
    assert len(base.splitlines()) == 10
    env = dict(os.environ, RKMH_SHARD_MIN="1000")
    for g in ([1, 2] if torch.cuda.device_count() >= 2 else [1]):
        res = subprocess.run([rkmh_bin, "hash", "-f", path, "-k", "21", "-s", "300", "-g", str(g)], stdout=subprocess.PIPE,
                             stderr=subprocess.PIPE, env=env)
        assert res.returncode == 0, res.stderr.decode()[-400:]
        assert res.stdout == base, g


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_minus_M_on_a_fifo_and_an_empty_input(rkmh_bin, cli_manifest, tmp_path):
    """-M reads the input twice; an input that cannot be re-opened (a FIFO, /dev/stdin, `<(zcat ...)`) is buffered
    like standard input instead of being drained by pass 1 (the reference loads all reads once).  An input without
    any record terminates with empty stdout (the end-of-stream notification cannot be lost)."""
    import threading
    name = "hpv16_minion_M2_N10"
    _, argv = CLI_CASES[name]
    i = argv.index("-f")
    payload = gzip.open(os.path.join(DATA, argv[i + 1]), "rb").read()
    fifo = tmp_path / "reads.fifo"
    os.mkfifo(fifo)

    def feed():
        with open(fifo, "wb") as fh:
            fh.write(payload)
    t = threading.Thread(target=feed, daemon=True)
    t.start()
    res = subprocess.run([rkmh_bin] + argv[:i + 1] + [str(fifo)] + argv[i + 2:], cwd=DATA, stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, timeout=120)
    t.join(timeout=10)
    assert res.returncode == 0, res.stderr.decode()[-400:]
    assert hashlib.md5(res.stdout).hexdigest() == cli_manifest[name]["md5"]
    empty = tmp_path / "empty.fq"
    empty.write_bytes(b"")
    for extra in ([], ["-M", "2"]):
        for _ in range(3):
            res = subprocess.run([rkmh_bin, "stream", "-r", "hpv_16.fa.gz", "-f", str(empty), "-k", "12", "-s", "100"] + extra,
                                 cwd=DATA, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
            assert res.returncode == 0 and res.stdout == b""


@pytest.mark.gpu
@pytest.mark.timeout(600)
@pytest.mark.parametrize("name", ["hpv16cmd_minion_k16", "hpv16cmd_mix2k_k12", "hpv16cmd_minion_k10_k14", "hpv16cmd_mix2k_k16_M2",
                                  "hpv16cmd_z1_k20", "hpv16cmd_two_files_k14_M3"])
def test_hpv16_command_equals_reference(rkmh_bin, cli_manifest, tmp_path, name):
    """`rkmh hpv16`: stdout, the messages on stderr and the side file lineage_specific_hashes.<k>.tst are the
    real binary's, byte for byte; small chunks (many batches, -M streamed twice) and a FIFO input do not change them."""
    from tests.golden.make_golden import HPV16_CASES
    argv = [os.path.join(DATA, a) if a.endswith(".gz") else a for a in HPV16_CASES[name]] + ["-R", DATA]
    k0 = argv[argv.index("-k") + 1]
    res = subprocess.run([rkmh_bin] + argv, cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert res.returncode == 0, res.stderr.decode()[-400:]
    gold = golden_cli_output(name)
    if res.stdout != gold:
        a, b = res.stdout.splitlines(), gold.splitlines()
        diff = [(i, x, y) for i, (x, y) in enumerate(zip(a, b)) if x != y][:2]
        pytest.fail(f"{name}: {len(a)} vs {len(b)} lines; first diffs {diff}")
    assert hashlib.md5(res.stderr).hexdigest() == cli_manifest[name]["stderr_md5"], res.stderr.decode()[-300:]
    side = (tmp_path / f"lineage_specific_hashes.{k0}.tst").read_bytes()
    assert hashlib.md5(side).hexdigest() == cli_manifest[name]["side_file_md5"]
    env = dict(os.environ, RKMH_CHUNK_BYTES="30000")
    res = subprocess.run([rkmh_bin] + argv, cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    assert res.returncode == 0 and res.stdout == gold


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [[], ["-N", "5", "-D", "1"], ["-M", "2", "-N", "3"]])
def test_filter_in_stream_mode(rkmh_bin, oracle, extra):
    """`rkmh filter ... -i` (src/rkmh.cpp:1329-1414): after the read files, every record of standard input is classified
    and REPORTED as `Sample: <name>\\tResult: <ref>\\t<shared>\\t<min sketch length>\\t[FAIL:DEPTH]\\t[FAIL:MATCHES]\\t[FAIL:DIFF]`.
    The reference's own code path cannot run (use after free in its OpenMP tasks, see rkmh_main.cpp), so the expected lines
    are restated on the pinned oracle: parity for this mode is pinned by the oracle, not by the real binary."""
    import numpy as np
    from rkmh_b200.fastx import load_sequences
    refs_f = ["zika.fa.gz", "dengue.fa.gz", "hpv_16.fa.gz"]
    rn, rs, _ = load_sequences([os.path.join(DATA, f) for f in refs_f])
    fn, fs, _ = load_sequences([os.path.join(DATA, "z1.fq.gz")])
    payload = gzip.open(os.path.join(DATA, "z1.fq.gz"), "rb").read() + gzip.open(os.path.join(DATA, "minION25.fq.gz"), "rb").read() + \
        b"@short\nACGTACGTACGTACGTACGTACGTAC\n+\nIIIIIIIIIIIIIIIIIIIIIIIIII\n@allN\nNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNN\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n"
    k, s = 20, 2000
    opt = dict(zip(extra[::2], map(int, extra[1::2])))
    n_min, d_min, m = opt.get("-N", -1), opt.get("-D", 0), opt.get("-M", -1)
    argv = [rkmh_bin, "filter"] + sum((["-r", f] for f in refs_f), []) + ["-f", "z1.fq.gz", "-k", str(k), "-s", str(s), "-i"] + extra
    res = subprocess.run(argv, cwd=DATA, input=payload, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert res.returncode == 0, res.stderr.decode()[-400:]
    # expected: the filtered records of the read file first (the ordinary filter output), then one line per stdin record
    base = subprocess.run([a for a in argv if a != "-i"], cwd=DATA, stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout
    assert res.stdout.startswith(base)
    tmp_names, tmp_seqs = [], []
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".fq") as fh:
        fh.write(payload); fh.flush()
        tmp_names, tmp_seqs, _ = load_sequences([fh.name])
    sk, ln = oracle.sketch_batch(rs, [k], s)
    counts = None
    if m > 0:
        counts = np.zeros(10_000_000, np.int32)
        oracle.count_batch(fs, [k], counts, valid_only=False)           # filter counts every hash of the read FILES, zeros too
    hits = oracle.classify_batch(tmp_seqs, [k], s, sk, ln, counts=counts, min_occ=m if m > 0 else 0)
    lines = []
    for nm, h in zip(tmp_names, hits):
        best, prev = max(int(h["best"]), 0), max(int(h["prev"]), 0)
        found = best > 0
        lines.append(b"Sample: " + nm + b"\tResult: " + (rn[int(h["ref_id"])] if found else b"") + b"\t%d\t%d\t" % (
            best, min(int(h["sketch_len"]), int(ln[int(h["ref_id"])])) if found else 0) +
            (b"FAIL:DEPTH" if int(h["sketch_len"]) <= 0 else b"") + b"\t" + (b"FAIL:MATCHES" if best < n_min else b"") + b"\t" +
            (b"" if best - prev > d_min else b"FAIL:DIFF") + b"\n")
    assert res.stdout[len(base):] == b"".join(lines)


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_persistence_pre_references_and_count_table(rkmh_bin, cli_manifest, tmp_path):
    """SURVEY 8f N3.  `hash -s N -o refs.json` writes the reference sketches in the reference's JSON schema; `stream -R
    refs.json` (no -r) then prints exactly what `stream -r refs.fa` prints (the real binary's golden).  With -M, `-p FILE`
    writes the read k-mer count table after pass 1 and a second run loads it instead of counting: same stdout."""
    import json
    name = "pave_mix2k_k12_s100"
    _, argv = CLI_CASES[name]
    js = tmp_path / "refs.json"
    res = subprocess.run([rkmh_bin, "hash", "-r", "all_pave_ref.fa.gz", "-k", "12", "-s", "100", "-o", str(js)], cwd=DATA,
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert res.returncode == 0 and res.stdout == b"", res.stderr.decode()[-300:]
    doc = json.loads(js.read_text())
    assert len(doc) == 182 and doc[0]["kmer"] == "12" and doc[0]["hashType"] == "MurmurHash3_x64_128" and doc[0]["hashSeed"] == 42
    assert doc[0]["sketches"]["length"] == 100 and len(doc[0]["sketches"]["hashes"]) == 100
    assert doc[0]["sketches"]["hashes"] == sorted(doc[0]["sketches"]["hashes"]) and doc[0]["seqLen"] > 7000
    i = argv.index("-r")
    pre = argv[:i] + ["-R", str(js)] + argv[i + 2:]
    res = subprocess.run([rkmh_bin] + pre, cwd=DATA, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert res.returncode == 0, res.stderr.decode()[-300:]
    assert hashlib.md5(res.stdout).hexdigest() == cli_manifest[name]["md5"]
    # sequences and pre-hashed references mixed: the two HPV16 reference files of this golden, the second one pre-hashed
    name2 = "two_ref_files_multik_M2"
    _, argv2 = CLI_CASES[name2]
    js2 = tmp_path / "new_refs.json"
    subprocess.run([rkmh_bin, "hash", "-r", "new_refs.fa.gz", "-k", "10", "-k", "14", "-s", "500", "-o", str(js2)], cwd=DATA, check=True,
                   stderr=subprocess.PIPE)
    j = [k for k, a in enumerate(argv2) if a == "-r"][1]
    mixed = argv2[:j] + ["-R", str(js2)] + argv2[j + 2:]
    res = subprocess.run([rkmh_bin] + mixed, cwd=DATA, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert res.returncode == 0, res.stderr.decode()[-300:]
    assert hashlib.md5(res.stdout).hexdigest() == cli_manifest[name2]["md5"]
    # a file for other parameters is refused
    bad = subprocess.run([rkmh_bin, "stream", "-R", str(js), "-f", "z1.fq.gz", "-k", "12", "-s", "1000"], cwd=DATA, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert bad.returncode == 2 and b"-s 100" in bad.stderr
    # the read k-mer count table
    for name3 in ("newrefs_mix2k_M20_N10", "filter_viral4_z1_M2_D2"):
        _, argv3 = CLI_CASES[name3]
        tab = tmp_path / (name3 + ".cnt")
        first = subprocess.run([rkmh_bin] + argv3 + ["-p", str(tab)], cwd=DATA, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert first.returncode == 0 and tab.stat().st_size > 64
        again = subprocess.run([rkmh_bin] + argv3 + ["-p", str(tab)], cwd=DATA, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                               env=dict(os.environ, RKMH_TIMING="1"))
        assert again.returncode == 0
        assert hashlib.md5(first.stdout).hexdigest() == cli_manifest[name3]["md5"] == hashlib.md5(again.stdout).hexdigest()
    # Vowpal Wabbit lines (print_wabbit): label, key with '|' -> '_', hash:1 pairs, the sketch namespace
    res = subprocess.run([rkmh_bin, "hash", "-f", "hpv_16_allFasta.fa.gz", "-k", "12", "-s", "5", "-w"], cwd=DATA, stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, check=True)
    line = res.stdout.splitlines()[0].decode()
    assert line.startswith("XYX 1.0 `gi_") and "|vir " in line and line.endswith(" |sketch k:12 s:5") and line.count(":1 ") == 5


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["call_hpv16_minion_k12", "call_hpv16_mix2k_k16", "call_zika_z1_k20_w50", "call_newrefs_mix2k_k14",
                                  "call_zika_z1_k9_w10", "call_show_depth"])
def test_call_command_equals_reference(rkmh_bin, cli_manifest, name):
    """`rkmh call`: the VCF the real binary prints, byte for byte -- exact k-mer depths, sliding-window
    average carried across references, substitution and deletion rescues, calls in std::map order."""
    from tests.golden.make_golden import CALL_CASES
    res = subprocess.run([rkmh_bin] + CALL_CASES[name], cwd=DATA, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert res.returncode == 0, res.stderr.decode()[-400:]
    gold = golden_cli_output(name)
    if res.stdout != gold:
        a, b = res.stdout.splitlines(), gold.splitlines()
        diff = [(i, x, y) for i, (x, y) in enumerate(zip(a, b)) if x != y][:3]
        pytest.fail(f"{name}: {len(a)} vs {len(b)} lines; first diffs {diff}")
    assert hashlib.md5(res.stdout).hexdigest() == cli_manifest[name]["md5"]
