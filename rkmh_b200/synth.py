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
    with open(path, "wb") as fh:
        step = 100000
        qual = b"I" * length
        for c0 in range(0, n, step):
            rows = mat[c0:c0 + step]
            parts = []
            for i in range(rows.shape[0]):
                parts.append(b"@%s%09d\n" % (prefix.encode(), c0 + i))
                parts.append(rows[i].tobytes())
                parts.append(b"\n+\n")
                parts.append(qual)
                parts.append(b"\n")
            fh.write(b"".join(parts))
