"""tests/emu/emu.py -- ctypes face of tests/emu/libemu.so (TEST INFRASTRUCTURE ONLY): the product's kernel
source compiled for the CPU by g++ against tests/emu/cuda_emu.h."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
HIT = np.dtype([("ref_id", "<i4"), ("best", "<i4"), ("prev", "<i4"), ("sketch_len", "<i4")])


def pack(seqs):
    offs = np.zeros(len(seqs) + 1, dtype=np.int64)
    if seqs:
        offs[1:] = np.cumsum([len(s) for s in seqs])
    return np.frombuffer(b"".join(seqs) + b"\0", dtype=np.uint8).copy(), offs


class Emu:
    def __init__(self):
        subprocess.check_call(["make", "-C", HERE, "libemu.so"], stdout=subprocess.DEVNULL)
        self.lib = lib = C.CDLL(os.path.join(HERE, "libemu.so"))
        lib.emu_fastmod.argtypes = [C.c_uint64, C.c_uint64]
        lib.emu_fastmod.restype = C.c_uint64
        lib.emu_kmer_hash.argtypes = [C.c_char_p, C.c_int]
        lib.emu_kmer_hash.restype = C.c_uint64
        lib.emu_window_hash.argtypes = [C.c_char_p, C.c_int]
        lib.emu_window_hash.restype = C.c_uint64
        P, I, L, U = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
        lib.emu_hash.argtypes = [P, P, L, P, I, P, P, I]
        lib.emu_count.argtypes = [P, P, L, P, I, P, U, I, I]
        lib.emu_sketch.argtypes = [P, P, L, P, I, I, P, U, I, I, P, P, I, I]
        lib.emu_merge.argtypes = [P, P, I, I, P, P, I, I]
        lib.emu_sketch_shard.argtypes = [P, L, L, I, I, P, P, I, I, I, P]
        lib.emu_unpack.argtypes = [P, P, P, L, I]
        lib.emu_classify_sets.argtypes = [P, P, I, P, P, L, I, P, U, I, I, I, I, P, P]
        lib.emu_call_depth.argtypes = [P, P, L, P, L, I, P, P, L, P, I]
        lib.emu_classify.argtypes = [P, P, I, P, P, L, P, I, I, P, U, I, I, I, I, I, I, P, P, P, P]
        lib.emu_count_memo.argtypes = [P, P, L, I, P, U, I, I, I]
        lib.emu_classify_memo.argtypes = [P, P, I, P, P, L, I, I, P, U, I, I, I, I, P, P]

    def hash(self, seqs, ks, block=64):
        bases, offs = pack(seqs)
        ks = np.atleast_1d(np.asarray(ks, np.int32))
        n = sum(max(len(s) - int(k), 0) for s in seqs for k in ks)
        out = np.zeros(max(n, 1), np.uint64)
        ho = np.zeros(len(seqs) + 1, np.int64)
        self.lib.emu_hash(bases.ctypes.data, offs.ctypes.data, len(seqs), ks.ctypes.data, len(ks), out.ctypes.data,
                          ho.ctypes.data, block)
        return out[:n], ho

    def count(self, seqs, ks, counts, count_zeros=False, block=64):
        bases, offs = pack(seqs)
        ks = np.atleast_1d(np.asarray(ks, np.int32))
        self.lib.emu_count(bases.ctypes.data, offs.ctypes.data, len(seqs), ks.ctypes.data, len(ks), counts.ctypes.data,
                           C.c_uint64(counts.size), int(count_zeros), block)

    def sketch(self, seqs, ks, s, counts=None, mode=0, occ=0, block=64, force_global_sort=False):
        bases, offs = pack(seqs)
        ks = np.atleast_1d(np.asarray(ks, np.int32))
        sk = np.zeros((len(seqs), s), np.uint64)
        lens = np.zeros(len(seqs), np.int32)
        self.lib.emu_sketch(bases.ctypes.data, offs.ctypes.data, len(seqs), ks.ctypes.data, len(ks), s,
                            counts.ctypes.data if counts is not None else None,
                            C.c_uint64(counts.size if counts is not None else 0), mode, occ,
                            sk.ctypes.data, lens.ctypes.data, block, int(force_global_sort))
        return sk, lens

    def sketch_shard(self, seq, n, k, s, block=64, scan_first=False, pieces=3, want_route=False):
        """rkg_sketch_shard's kernels; scan_first: the one-pass route with the guessed limit first (route 1 when accepted)"""
        seq = np.ascontiguousarray(np.frombuffer(seq, np.uint8))
        out = np.zeros(s, np.uint64)
        m = np.zeros(1, np.int32)
        route = np.zeros(1, np.int32)
        rc = self.lib.emu_sketch_shard(seq.ctypes.data, C.c_int64(seq.size), C.c_int64(n), int(k), int(s), out.ctypes.data,
                                       m.ctypes.data, block, int(scan_first), int(pieces), route.ctypes.data)
        assert rc == 0, rc
        return (out[:m[0]], int(route[0])) if want_route else out[:m[0]]

    def classify_sets(self, sets, reads, k, ntypes, counts=None, min_occ=-1, block=64, hbuf_global=False):
        """rkg_load_ref_sets + rkg_classify_sets: (hits over sets[:ntypes], counts [n, len(sets) - ntypes] of the rest)"""
        so = np.zeros(len(sets) + 1, np.int64)
        so[1:] = np.cumsum([len(x) for x in sets])
        sh = np.ascontiguousarray(np.concatenate([np.asarray(x, np.uint64) for x in sets] + [np.zeros(1, np.uint64)]))
        bases, offs = pack(reads)
        out = np.zeros(max(len(reads), 1), HIT)
        nextra = len(sets) - ntypes
        extra = np.zeros((max(len(reads), 1), max(nextra, 1)), np.int32)
        rc = self.lib.emu_classify_sets(sh.ctypes.data, so.ctypes.data, len(sets), bases.ctypes.data, offs.ctypes.data, len(reads), int(k),
                                        counts.ctypes.data if counts is not None else None,
                                        C.c_uint64(counts.size if counts is not None else 0), int(min_occ), int(ntypes), block,
                                        int(hbuf_global), out.ctypes.data, extra.ctypes.data)
        assert rc == 0
        return out[:len(reads)], extra[:len(reads), :nextra]

    def call_depth(self, reads, ref, k, windows, block=64):
        """rk_call.cuh: (depth of every window of ref, depth of the 4 k single-base alternatives of the listed windows)"""
        bases, offs = pack(reads)
        refa = np.ascontiguousarray(np.frombuffer(ref, np.uint8))
        n = max(len(ref) - k, 0)
        depth = np.zeros(max(n, 1), np.int32)
        win = np.ascontiguousarray(windows, np.int64)
        resc = np.zeros(max(win.size * 4 * k, 1), np.int32)
        rc = self.lib.emu_call_depth(bases.ctypes.data, offs.ctypes.data, len(reads), refa.ctypes.data, C.c_int64(refa.size), int(k),
                                     depth.ctypes.data, win.ctypes.data, C.c_int64(win.size), resc.ctypes.data, block)
        assert rc == 0
        return depth[:n], resc[:win.size * 4 * k].reshape(win.size, 4 * k)

    def unpack(self, codes, nmask, block=64):
        """unpack_kernel: 32 ASCII bytes per code pair / mask word"""
        codes = np.ascontiguousarray(codes, np.uint32); nmask = np.ascontiguousarray(nmask, np.uint32)
        out = np.zeros(32 * nmask.size, np.uint8)
        self.lib.emu_unpack(codes.ctypes.data, nmask.ctypes.data, out.ctypes.data, C.c_int64(nmask.size), block)
        return out

    def merge(self, parts, lens, s, block=64, in_smem=True):
        parts = np.ascontiguousarray(parts, np.uint64)
        lens = np.ascontiguousarray(lens, np.int32)
        out = np.zeros(s, np.uint64)
        n = np.zeros(1, np.int32)
        self.lib.emu_merge(parts.ctypes.data, lens.ctypes.data, len(lens), s, out.ctypes.data, n.ctypes.data, block,
                           int(in_smem))
        return out[:n[0]]

    def classify(self, refs, reads, ks, s, counts=None, min_occ=-1, path=1, glog2=4, block=64, hbuf_global=False,
                 hs_global=False):
        rb, ro = pack(refs)
        qb, qo = pack(reads)
        ks = np.atleast_1d(np.asarray(ks, np.int32))
        hits = np.zeros(max(len(reads), 1), HIT)
        rsk = np.zeros((len(refs), s), np.uint64)
        rl = np.zeros(len(refs), np.int32)
        redo = np.zeros(1, np.int32)
        self.lib.emu_classify(rb.ctypes.data, ro.ctypes.data, len(refs), qb.ctypes.data, qo.ctypes.data, len(reads),
                              ks.ctypes.data, len(ks), s, counts.ctypes.data if counts is not None else None,
                              C.c_uint64(counts.size if counts is not None else 0), min_occ, path, glog2, block,
                              int(hbuf_global), int(hs_global), hits.ctypes.data, rsk.ctypes.data, rl.ctypes.data,
                              redo.ctypes.data)
        return hits[:len(reads)], rsk, rl, int(redo[0])

    # ---- memoised path (rk_memo.cuh) ----
    def count_memo(self, seqs, k, counts, count_zeros=False, wpr=1, block=64):
        bases, offs = pack(seqs)
        rc = self.lib.emu_count_memo(bases.ctypes.data, offs.ctypes.data, len(seqs), int(k), counts.ctypes.data,
                                     C.c_uint64(counts.size), int(count_zeros), int(wpr), block)
        assert rc == 0

    def classify_memo(self, refs, reads, k, s, counts=None, min_occ=-1, path=1, block=64, vcap_limit=0):
        rb, ro = pack(refs)
        qb, qo = pack(reads)
        hits = np.zeros(max(len(reads), 1), HIT)
        redo = np.zeros(1, np.int32)
        rc = self.lib.emu_classify_memo(rb.ctypes.data, ro.ctypes.data, len(refs), qb.ctypes.data, qo.ctypes.data, len(reads),
                                        int(k), int(s), counts.ctypes.data if counts is not None else None,
                                        C.c_uint64(counts.size if counts is not None else 0), min_occ, path, block,
                                        vcap_limit, hits.ctypes.data, redo.ctypes.data)
        assert rc == 0
        return hits[:len(reads)], int(redo[0])
