"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes loaders for the CPU checkers:
  * ``Oracle``     -> oracle/liboracle.so          (our C restatement, oracle/rkmh_oracle.c)
  * ``RefMkmh``    -> oracle/_ref/libmkmh_ref.so   (the real mkmh library behind oracle/ref_shim.cpp)
  * ``ref_binary`` -> oracle/_ref/rkmh_ref[_filter] (the real rkmh CLI, see oracle/Makefile)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module; the product package rkmh_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_HIT = np.dtype([("ref_id", "<i4"), ("best", "<i4"), ("prev", "<i4"), ("sketch_len", "<i4")])


def build(force=False):
    """(Re)build liboracle.so and, when /root/reference is present, oracle/_ref/*."""
    so = os.path.join(HERE, "liboracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(HERE, "rkmh_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)
    return so


def _as_bases(seqs):
    """list of bytes -> (uint8 concatenation, int64 offsets[n+1])"""
    offs = np.zeros(len(seqs) + 1, dtype=np.int64)
    if seqs:
        offs[1:] = np.cumsum([len(s) for s in seqs])
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if seqs else np.zeros(0, np.uint8)
    return bases, offs


class Oracle:
    """The C restatement (oracle/rkmh_oracle.c)."""

    def __init__(self):
        self.lib = lib = C.CDLL(build())
        lib.rko_murmur3_x64_128.argtypes = [C.c_char_p, C.c_int, C.c_uint32, C.c_void_p]
        lib.rko_calc_hash.argtypes = [C.c_char_p, C.c_int]
        lib.rko_calc_hash.restype = C.c_uint64
        lib.rko_calc_hashes_multi.argtypes = [C.c_char_p, C.c_int64, _i32p, C.c_int, _u64p]
        lib.rko_calc_hashes_multi.restype = C.c_int64
        lib.rko_num_hashes.argtypes = [C.c_int64, _i32p, C.c_int]
        lib.rko_num_hashes.restype = C.c_int64
        lib.rko_minhashes.argtypes = [_u64p, C.c_int64, C.c_int, _u64p]
        lib.rko_minhashes.restype = C.c_int
        lib.rko_intersection_size.argtypes = [_u64p, C.c_int, _u64p, C.c_int]
        lib.rko_intersection_size.restype = C.c_int
        lib.rko_merge_sketches.argtypes = [_u64p, C.c_int, _u64p, C.c_int, C.c_int, _u64p]
        lib.rko_merge_sketches.restype = C.c_int
        lib.rko_to_upper.argtypes = [C.c_char_p, C.c_int64]
        lib.rko_canonical.argtypes = [C.c_char_p, C.c_int]
        lib.rko_canonical.restype = C.c_int
        lib.rko_reverse_complement.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        lib.rko_argmax.argtypes = [_i32p, C.c_int, C.c_int, C.c_void_p]
        lib.rko_count_increment.argtypes = [_i32p, C.c_uint64, _u64p, C.c_int64, C.c_int]
        lib.rko_mask_by_frequency.argtypes = [_u64p, C.c_int64, _i32p, C.c_uint64, C.c_int]
        lib.rko_sketch_batch.argtypes = [C.c_void_p, _i64p, C.c_int64, _i32p, C.c_int, C.c_int,
                                         C.c_void_p, C.c_uint64, C.c_int, C.c_int, _u64p, _i32p]
        lib.rko_count_batch.argtypes = [C.c_void_p, _i64p, C.c_int64, _i32p, C.c_int,
                                        _i32p, C.c_uint64, C.c_int, C.c_int]
        lib.rko_classify_batch.argtypes = [C.c_void_p, _i64p, C.c_int64, _i32p, C.c_int, C.c_int,
                                           _u64p, _i32p, C.c_int, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        lib.rko_set_intersection_size.argtypes = [_u64p, C.c_int64, _u64p, C.c_int64]
        lib.rko_set_intersection_size.restype = C.c_int
        lib.rko_classify_sets.argtypes = [C.c_void_p, _i64p, C.c_int64, _i32p, C.c_int, _u64p, _i64p, C.c_int, C.c_int,
                                          C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]

    # -- primitives -------------------------------------------------------------------------
    def murmur3_x64_128(self, key: bytes, seed: int) -> bytes:
        out = C.create_string_buffer(16)
        self.lib.rko_murmur3_x64_128(key, len(key), seed, out)
        return out.raw

    def calc_hash(self, kmer: bytes) -> int:
        return int(self.lib.rko_calc_hash(kmer, len(kmer)))

    def to_upper(self, seq: bytes) -> bytes:
        buf = C.create_string_buffer(seq, len(seq))
        self.lib.rko_to_upper(buf, len(seq))
        return buf.raw

    def canonical(self, seq: bytes) -> bool:
        return bool(self.lib.rko_canonical(seq, len(seq)))

    def reverse_complement(self, seq: bytes) -> bytes:
        out = C.create_string_buffer(len(seq) + 1)
        self.lib.rko_reverse_complement(seq, out, len(seq))
        return out.raw[:len(seq)]

    def calc_hashes(self, seq: bytes, ks) -> np.ndarray:
        """seq must already be upper-cased, as in the reference call sites."""
        ks = np.atleast_1d(np.asarray(ks, dtype=np.int32))
        n = int(self.lib.rko_num_hashes(len(seq), ks, len(ks)))
        out = np.zeros(max(n, 1), dtype=np.uint64)
        got = self.lib.rko_calc_hashes_multi(seq, len(seq), ks, len(ks), out)
        assert got == n
        return out[:n]

    def minhashes(self, hashes: np.ndarray, s: int) -> np.ndarray:
        h = np.ascontiguousarray(hashes, dtype=np.uint64).copy()
        if h.size == 0:
            return np.zeros(0, np.uint64)
        ret = np.zeros(max(min(h.size, s), 1), dtype=np.uint64)
        n = self.lib.rko_minhashes(h, h.size, s, ret)
        return ret[:n]

    def intersection_size(self, a, b) -> int:
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = np.ascontiguousarray(b, dtype=np.uint64)
        if a.size == 0 or b.size == 0:
            return 0
        return int(self.lib.rko_intersection_size(a, a.size, b, b.size))

    def merge_sketches(self, a, b, s) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = np.ascontiguousarray(b, dtype=np.uint64)
        out = np.zeros(max(s, 1), np.uint64)
        aa = a if a.size else np.zeros(1, np.uint64)
        bb = b if b.size else np.zeros(1, np.uint64)
        n = self.lib.rko_merge_sketches(aa, a.size, bb, b.size, s, out)
        return out[:n]

    def argmax(self, counts, sketch_len=0):
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        out = np.zeros(1, _HIT)
        self.lib.rko_argmax(counts if counts.size else np.zeros(1, np.int32), counts.size, sketch_len,
                            out.ctypes.data)
        return out[0]

    # -- batch drivers ----------------------------------------------------------------------
    def count_batch(self, seqs, ks, counts, valid_only=True, per_sample_set=False):
        bases, offs = _as_bases(seqs)
        ks = np.atleast_1d(np.asarray(ks, dtype=np.int32))
        self.lib.rko_count_batch(bases.ctypes.data, offs, len(seqs), ks, len(ks), counts, counts.size,
                                 int(valid_only), int(per_sample_set))

    def sketch_batch(self, seqs, ks, s, counts=None, mode=0, occ=0):
        bases, offs = _as_bases(seqs)
        ks = np.atleast_1d(np.asarray(ks, dtype=np.int32))
        sk = np.zeros((max(len(seqs), 1), s), dtype=np.uint64)
        lens = np.zeros(max(len(seqs), 1), dtype=np.int32)
        cptr = counts.ctypes.data if counts is not None else None
        csz = counts.size if counts is not None else 0
        self.lib.rko_sketch_batch(bases.ctypes.data, offs, len(seqs), ks, len(ks), s, cptr, csz, mode, occ,
                                  sk.reshape(-1), lens)
        return sk[:len(seqs)], lens[:len(seqs)]

    def classify_batch(self, seqs, ks, s, ref_sketches, ref_lens, counts=None, min_occ=0):
        bases, offs = _as_bases(seqs)
        ks = np.atleast_1d(np.asarray(ks, dtype=np.int32))
        out = np.zeros(max(len(seqs), 1), _HIT)
        rs = np.ascontiguousarray(ref_sketches, dtype=np.uint64).reshape(-1)
        rl = np.ascontiguousarray(ref_lens, dtype=np.int32)
        cptr = counts.ctypes.data if counts is not None else None
        csz = counts.size if counts is not None else 0
        self.lib.rko_classify_batch(bases.ctypes.data, offs, len(seqs), ks, len(ks), s, rs, rl, len(rl),
                                    cptr, csz, min_occ, out.ctypes.data)
        return out[:len(seqs)]

    # -- hpv16 ---------------------------------------------------------------------
    def set_intersection_size(self, a, b) -> int:
        """mkmh/mkmh.hpp:842-863 on two ascending arrays"""
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = np.ascontiguousarray(b, dtype=np.uint64)
        return int(self.lib.rko_set_intersection_size(a, len(a), b, len(b)))

    def classify_sets(self, seqs, ks, sets, ntypes, counts=None, min_occ=0):
        """src/rkmh.cpp:2643-2706: (hits over sets[:ntypes], extra counts [n, len(sets) - ntypes]).  `sets` is a list of
        ASCENDING uint64 arrays (zeros and duplicates allowed, as the reference's type arrays have them)."""
        bases, offs = _as_bases(seqs)
        ks = np.atleast_1d(np.asarray(ks, dtype=np.int32))
        so = np.zeros(len(sets) + 1, np.int64)
        so[1:] = np.cumsum([len(x) for x in sets])
        sh = np.ascontiguousarray(np.concatenate([np.asarray(x, np.uint64) for x in sets]) if sets else np.zeros(1, np.uint64))
        out = np.zeros(max(len(seqs), 1), _HIT)
        extra = np.zeros((max(len(seqs), 1), max(len(sets) - ntypes, 1)), np.int32)
        cptr = counts.ctypes.data if counts is not None else None
        csz = counts.size if counts is not None else 0
        self.lib.rko_classify_sets(bases.ctypes.data, offs, len(seqs), ks, len(ks), sh, so, len(sets), ntypes,
                                   cptr, csz, min_occ, out.ctypes.data, extra.ctypes.data)
        return out[:len(seqs)], extra[:len(seqs), :len(sets) - ntypes]


class RefMkmh:
    """The real mkmh library (oracle/_ref/libmkmh_ref.so).  Raises FileNotFoundError when it was not built."""

    def __init__(self):
        path = os.path.join(REF_DIR, "libmkmh_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = lib = C.CDLL(path)
        lib.ref_murmur3_x64_128.argtypes = [C.c_char_p, C.c_int, C.c_uint32, C.c_void_p]
        lib.ref_calc_hashes_multi.argtypes = [C.c_char_p, C.c_int64, _i32p, C.c_int, _u64p]
        lib.ref_calc_hashes_multi.restype = C.c_int64
        lib.ref_minhashes.argtypes = [_u64p, C.c_int64, C.c_int, _u64p]
        lib.ref_minhashes.restype = C.c_int
        lib.ref_intersection_size.argtypes = [_u64p, C.c_int, _u64p, C.c_int]
        lib.ref_intersection_size.restype = C.c_int
        lib.ref_to_upper.argtypes = [C.c_char_p, C.c_int64]
        lib.ref_counter_new.argtypes = [C.c_uint64]
        lib.ref_counter_new.restype = C.c_void_p
        lib.ref_counter_free.argtypes = [C.c_void_p]
        lib.ref_counter_increment.argtypes = [C.c_void_p, C.c_uint64]
        lib.ref_counter_get.argtypes = [C.c_void_p, C.c_uint64]
        lib.ref_counter_get.restype = C.c_int
        lib.ref_mask_by_frequency.argtypes = [_u64p, C.c_int64, C.c_void_p, C.c_int]

    def murmur3_x64_128(self, key: bytes, seed: int) -> bytes:
        out = C.create_string_buffer(16)
        self.lib.ref_murmur3_x64_128(key, len(key), seed, out)
        return out.raw

    def to_upper(self, seq: bytes) -> bytes:
        buf = C.create_string_buffer(seq, len(seq))
        self.lib.ref_to_upper(buf, len(seq))
        return buf.raw

    def calc_hashes(self, seq: bytes, ks) -> np.ndarray:
        ks = np.atleast_1d(np.asarray(ks, dtype=np.int32))
        n = sum(len(seq) - int(k) for k in ks)
        out = np.zeros(max(n, 1), dtype=np.uint64)
        got = self.lib.ref_calc_hashes_multi(seq, len(seq), ks, len(ks), out)
        assert got == n
        return out[:n]

    def minhashes(self, hashes, s):
        h = np.ascontiguousarray(hashes, dtype=np.uint64).copy()
        ret = np.zeros(max(min(h.size, s), 1), dtype=np.uint64)
        n = self.lib.ref_minhashes(h, h.size, s, ret)
        return ret[:n]

    def intersection_size(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = np.ascontiguousarray(b, dtype=np.uint64)
        return int(self.lib.ref_intersection_size(a, a.size, b, b.size))


def ref_binary(filter_build=False):
    """Path of the real rkmh CLI built by oracle/Makefile, or None."""
    p = os.path.join(REF_DIR, "rkmh_ref_filter" if filter_build else "rkmh_ref")
    return p if os.path.exists(p) else None


HIT_DTYPE = _HIT
