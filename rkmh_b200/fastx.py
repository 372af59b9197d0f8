"""FASTA/FASTQ record reader with the record rules of the reference's parser.

Mirrors ``kseq_read`` as used by ``parse_fastas`` (reference kseq_reader/kseq.hpp:170-208,
src/rkmh.cpp:238-292): a record starts at '>' or '@'; the name is the header up to the first
white-space byte; the sequence is every ``isgraph`` byte up to the next '>', '@' or '+' byte
(anywhere, not only at line starts); for FASTQ the rest of the '+' line is skipped and quality
bytes in [33,127] are collected until as many as sequence bytes were read.  gzip input is
transparent.  This is host-side convenience for the Python mirror (tests, bench, golden scripts);
the C++ CLI has its own implementation of the same rules (rkmh_b200/host/fastx.hpp).
"""
import gzip
from typing import Iterator, Tuple


def _open(path):
    with open(path, "rb") as fh:
        magic = fh.read(2)
    return gzip.open(path, "rb") if magic == b"\x1f\x8b" else open(path, "rb")


def read_fastx(path) -> Iterator[Tuple[bytes, bytes, bytes]]:
    """Yield (name, sequence, quality) with sequence exactly as stored (not upper-cased)."""
    with _open(path) as fh:
        data = fh.read()
    n = len(data)
    i = 0
    # jump to the first header
    while i < n and data[i] not in b">@":
        i += 1
    while i < n:
        i += 1  # past '>' / '@'
        if i >= n:
            return  # a header byte that is the last byte of the input starts no record (ks_getuntil returns -1, kseq.hpp:97)
        j = i
        while j < n and not chr(data[j]).isspace():
            j += 1
        name = data[i:j]
        # rest of the header line is the comment
        if j < n and data[j] != 0x0A:
            nl = data.find(b"\n", j)
            j = n if nl < 0 else nl
        i = j + 1
        seq = bytearray()
        c = -1
        while i < n:
            c = data[i]
            i += 1
            if c in b">+@":
                break
            if 33 <= c <= 126:
                seq.append(c)
            c = -1
        if c != 0x2B:  # FASTA (or EOF): c is '>' / '@' of the next record, already consumed
            yield name, bytes(seq), b""
            if c == -1:
                return
            i -= 1  # step back onto the header byte
            continue
        nl = data.find(b"\n", i)  # skip the rest of the '+' line
        if nl < 0:
            return  # truncated: the reference stops here (kseq returns -2)
        i = nl + 1
        qual = bytearray()
        while i < n and len(qual) < len(seq):
            c = data[i]
            i += 1
            if 33 <= c <= 127:
                qual.append(c)
        if len(qual) != len(seq):
            return
        # kseq consumes one more byte after the last quality byte before testing the length
        if i < n:
            i += 1
        yield name, bytes(seq), bytes(qual)
        while i < n and data[i] not in b">@":
            i += 1


def load_sequences(paths):
    """-> (names[list[bytes]], seqs[list[bytes]], quals[list[bytes]]) over all files in order."""
    names, seqs, quals = [], [], []
    for p in paths:
        for nm, sq, ql in read_fastx(p):
            names.append(nm)
            seqs.append(sq)
            quals.append(ql)
    return names, seqs, quals
