"""Seeded synthetic read generators for the configurations of BASELINE.json.

Vectorised numpy; used by bench.py and the tests.  Reads come back in the layout the C ABI takes:
a uint8 base buffer plus int64 offsets.
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, np.uint8)
_COMP[:] = np.arange(256, dtype=np.uint8)
for a, b in (b"AT", b"TA", b"CG", b"GC"):
    _COMP[a] = b


def upper_refs(refs):
    """list[bytes] -> list[np.uint8 arrays], upper-cased ACGT letters (other bytes kept)."""
    out = []
    for r in refs:
        a = np.frombuffer(r, np.uint8).copy()
        low = (a >= 97) & (a <= 122)
        a[low] -= 32
        out.append(a)
    return out


def sample_reads(rng, refs, n, length, sub=0.005, nrate=0.001, random_frac=0.1, chunk=1 << 18):
    """cfg2-style reads: (1 - random_frac) sampled from a uniformly chosen reference / strand / offset with
    `sub` substitutions and `nrate` N bases, the rest uniform random ACGT.  Fixed length.
    Returns (bases uint8 [n * length], offs int64 [n + 1])."""
    refs_u = upper_refs(refs)
    usable = [r for r in refs_u if r.size > length]
    cat = np.concatenate(usable)
    starts = np.concatenate([[0], np.cumsum([r.size for r in usable])[:-1]])
    sizes = np.array([r.size for r in usable])
    bases = np.empty(n * length, np.uint8)
    ar = np.arange(length, dtype=np.int64)
    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        ridx = rng.integers(0, len(usable), m)
        off = (rng.random(m) * (sizes[ridx] - length)).astype(np.int64)
        pos = (starts[ridx] + off)[:, None] + ar[None, :]
        blk = cat[pos]
        rev = rng.random(m) < 0.5
        blk[rev] = _COMP[blk[rev][:, ::-1]]
        rnd = rng.random(m) < random_frac
        nr = int(rnd.sum())
        if nr:
            blk[rnd] = _ACGT[rng.integers(0, 4, (nr, length))]
        if sub > 0:
            mask = rng.random((m, length)) < sub
            blk[mask] = _ACGT[rng.integers(0, 4, int(mask.sum()))]
        if nrate > 0:
            mask = rng.random((m, length)) < nrate
            blk[mask] = ord("N")
        bases[c0 * length:(c0 + m) * length] = blk.reshape(-1)
    offs = np.arange(n + 1, dtype=np.int64) * length
    return bases, offs


def write_fastq(path, bases, offs, prefix="r"):
    """Write the batch as FASTQ with constant quality 'I' (for the reference CLI baseline)."""
    n = len(offs) - 1
    length = int(offs[1] - offs[0]) if n else 0
    assert n == 0 or (np.diff(offs) == length).all(), "fixed-length batches only"
    mat = bases[: n * length].reshape(n, length)
    pre = prefix.encode()
    rec = 1 + len(pre) + 9 + 1 + length + 3 + length + 1          # "@" prefix 9 digits "\n" bases "\n+\n" quality "\n"
    with open(path, "wb") as fh:
        step = 200000
        for c0 in range(0, n, step):
            rows = mat[c0:c0 + step]
            m = rows.shape[0]
            out = np.empty((m, rec), np.uint8)
            out[:, 0] = ord("@")
            out[:, 1:1 + len(pre)] = np.frombuffer(pre, np.uint8)
            idx = np.arange(c0, c0 + m, dtype=np.int64)
            for d in range(9):
                out[:, 1 + len(pre) + 8 - d] = 48 + (idx // 10 ** d) % 10
            p = 1 + len(pre) + 9
            out[:, p] = 10
            out[:, p + 1:p + 1 + length] = rows
            out[:, p + 1 + length:p + 4 + length] = np.frombuffer(b"\n+\n", np.uint8)
            out[:, p + 4 + length:p + 4 + 2 * length] = ord("I")
            out[:, rec - 1] = 10
            fh.write(out.tobytes())


def ont_reads(rng, genomes, n, length, sub=0.04, ins=0.03, dele=0.03, chunk=2048):
    """cfg3-style long reads: each read is a circular walk on one uniformly chosen genome (random start, random
    strand) with ONT-like errors: `sub` substitutions, `ins` inserted random bases, `dele` skipped genome bases.
    Fixed output length.  Returns (bases uint8 [n * length], offs int64 [n + 1])."""
    gs = upper_refs(genomes)
    bases = np.empty(n * length, np.uint8)
    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        gi = rng.integers(0, len(gs), m)
        u = rng.random((m, length), dtype=np.float32)
        is_ins = u < ins
        step = np.where(is_ins, 0, 1 + (u > 1.0 - dele)).astype(np.int32)      # 0 = inserted base, 2 = one base skipped
        walk = np.cumsum(step, axis=1, dtype=np.int32) - step
        blk = np.empty((m, length), np.uint8)
        for g in np.unique(gi):
            rows = np.nonzero(gi == g)[0]
            G = gs[g]
            start = rng.integers(0, G.size, rows.size)
            src = (start[:, None] + walk[rows]) % G.size
            seg = G[src]
            rev = rng.random(rows.size) < 0.5
            seg[rev] = _COMP[seg[rev][:, ::-1]]
            blk[rows] = seg
        noise = is_ins | (rng.random((m, length), dtype=np.float32) < sub)
        blk[noise] = _ACGT[rng.integers(0, 4, int(noise.sum()))]
        bases[c0 * length:(c0 + m) * length] = blk.reshape(-1)
    return bases, np.arange(n + 1, dtype=np.int64) * length


def background_reads(rng, n, length, gc=0.41, chunk=1 << 18):
    """cfg4-style "human-like" background: order-0 random bases with the given GC fraction."""
    p = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])     # A C G T
    bases = np.empty(n * length, np.uint8)
    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        bases[c0 * length:(c0 + m) * length] = _ACGT[rng.choice(4, size=m * length, p=p)]
    return bases, np.arange(n + 1, dtype=np.int64) * length


def mixed_reads(rng, refs, n, length, viral_frac=0.05, gc=0.41):
    """cfg4: `viral_frac` of the reads sampled from the references as in cfg2, the rest background; shuffled."""
    nv = int(round(n * viral_frac))
    vb, _ = sample_reads(rng, refs, nv, length, random_frac=0.0)
    bb, _ = background_reads(rng, n - nv, length, gc)
    mat = np.concatenate([vb.reshape(nv, length), bb.reshape(n - nv, length)])
    perm = rng.permutation(n)
    return np.ascontiguousarray(mat[perm]).reshape(-1), np.arange(n + 1, dtype=np.int64) * length, perm < nv


def random_genome(rng, nbases, n_run_rate=1e-5, n_run_len=1000, chunk=1 << 26):
    """cfg5: one uniform ACGT sequence with ~1 % of its bases inside runs of N."""
    out = np.empty(nbases, np.uint8)
    for c0 in range(0, nbases, chunk):
        m = min(chunk, nbases - c0)
        out[c0:c0 + m] = _ACGT[rng.integers(0, 4, m, dtype=np.uint8)]
    nruns = int(nbases * n_run_rate)
    for st in rng.integers(0, max(nbases - n_run_len, 1), nruns):
        out[st:st + n_run_len] = ord("N")
    return out


# ---------------------------------------------------------------------------------------------------------------------
# The same workloads from the C++ generator (rkmh_b200/host/synth.cpp: std::mt19937_64, block-seeded, multi-threaded):
# 20-50x faster than the numpy versions above, and a pure function of (seed, unit index), so rank r of N can generate
# its slice [first, first + n) of the one global stream a single process would generate whole.
_SYN = None


def _syn():
    global _SYN
    if _SYN is None:
        import ctypes as C
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin", "libsynth.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `make -C rkmh_b200/host`")
        lib = C.CDLL(path)
        P, I, L, D, U = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_uint64
        lib.syn_sample_reads.argtypes = [P, P, I, L, L, I, D, D, D, U, P, I]
        lib.syn_ont_reads.argtypes = [P, P, I, L, L, I, D, D, D, U, P, I]
        lib.syn_background.argtypes = [L, L, D, U, P, I]
        lib.syn_genome.argtypes = [L, L, D, I, U, P, I]
        _SYN = lib
    return _SYN


def _threads():
    import os
    try:
        return max(1, min(len(os.sched_getaffinity(0)), 16))
    except AttributeError:
        return 4


def _cat(seqs):
    offs = np.zeros(len(seqs) + 1, np.int64)
    offs[1:] = np.cumsum([len(s) for s in seqs])
    return np.frombuffer(b"".join(seqs), np.uint8), offs


def _out(out, nbytes):
    if out is None:
        out = np.empty(nbytes, np.uint8)
    assert out.dtype == np.uint8 and out.size >= nbytes and out.flags.c_contiguous
    return out


def fast_sample_reads(seed, refs, first, n, length, sub=0.005, nrate=0.001, random_frac=0.1, out=None):
    """Reads [first, first + n) of the config-2 stream of `seed` (see sample_reads); writes into `out` when given
    (e.g. a pinned buffer).  Returns (bases uint8 [n * length], offs int64 [n + 1])."""
    cat, offs = _cat(refs)
    out = _out(out, n * length)
    rc = _syn().syn_sample_reads(cat.ctypes.data, offs.ctypes.data, len(refs), int(first), int(n), int(length), sub, nrate,
                                 random_frac, int(seed), out.ctypes.data, _threads())
    assert rc == 0
    return out[:n * length], np.arange(n + 1, dtype=np.int64) * length


def fast_ont_reads(seed, genomes, first, n, length, sub=0.04, ins=0.03, dele=0.03, out=None):
    """Reads [first, first + n) of the config-3 stream of `seed` (see ont_reads)."""
    cat, offs = _cat(genomes)
    out = _out(out, n * length)
    rc = _syn().syn_ont_reads(cat.ctypes.data, offs.ctypes.data, len(genomes), int(first), int(n), int(length), sub, ins,
                              dele, int(seed), out.ctypes.data, _threads())
    assert rc == 0
    return out[:n * length], np.arange(n + 1, dtype=np.int64) * length


def fast_background(seed, first, n, gc=0.41, out=None):
    """Bases [first, first + n) of the config-4 background stream of `seed`."""
    out = _out(out, n)
    assert _syn().syn_background(int(first), int(n), gc, int(seed), out.ctypes.data, _threads()) == 0
    return out[:n]


def fast_genome(seed, first, n, n_frac=0.01, n_run_len=1000, out=None):
    """Bases [first, first + n) of the config-5 sequence of `seed`: uniform ACGT, `n_frac` of it inside runs of N."""
    out = _out(out, n)
    assert _syn().syn_genome(int(first), int(n), n_frac, int(n_run_len), int(seed), out.ctypes.data, _threads()) == 0
    return out[:n]
