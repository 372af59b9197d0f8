"""Python host-side mirror of the rkmh hot path over the C ABI (include/rkmh_gpu.h).

The reference's host is C++ (src/rkmh.cpp); the product's C++ front-end lives in rkmh_b200/host/.  This
module is the same call sequence for Python callers (tests, bench.py): names follow the reference
(`calc_hashes`, `minhashes`, `stream` ...), the arithmetic all happens in the CUDA library.
"""
import ctypes as C

import numpy as np

from . import _lib

HIT_DTYPE = np.dtype([("ref_id", "<i4"), ("best", "<i4"), ("prev", "<i4"), ("sketch_len", "<i4")])

STREAM_TABLE_SIZE = 200_000_000   # reference src/rkmh.cpp:739,742
FILTER_TABLE_SIZE = 10_000_000    # reference src/rkmh.cpp:1187-1188


class RkmhError(RuntimeError):
    pass


def pack_sequences(seqs):
    """list[bytes] -> (uint8 bases, int64 offsets[n+1]) -- the layout every entry point takes."""
    offs = np.zeros(len(seqs) + 1, dtype=np.int64)
    if len(seqs):
        offs[1:] = np.cumsum([len(s) for s in seqs])
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    if bases.size == 0:
        bases = np.zeros(1, np.uint8)
    return np.ascontiguousarray(bases), offs


class PinnedBuffer:
    """A pinned host allocation owned by a context, exposed as a numpy array."""

    def __init__(self, ctx, nbytes):
        self._ctx = ctx
        ptr = C.c_void_p()
        ctx._check(ctx.lib.rkg_host_alloc(ctx.handle, max(int(nbytes), 1), C.byref(ptr)))
        self.ptr = ptr.value
        self.nbytes = int(nbytes)
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr))

    def view(self, dtype, count=None):
        a = self.array.view(dtype)
        return a if count is None else a[:count]

    def free(self):
        if self.ptr:
            self._ctx.lib.rkg_host_free(self._ctx.handle, self.ptr)
            self.ptr = None


class Batch:
    """A read batch resident in device memory (rkg_batch)."""

    def __init__(self, ctx, handle, nreads):
        self.ctx, self.handle, self.nreads = ctx, handle, nreads

    def classify(self, min_occ=-1):
        self.ctx._check(self.ctx.lib.rkg_batch_classify(self.ctx.handle, self.handle, min_occ))

    def count(self, count_zeros=False):
        self.ctx._check(self.ctx.lib.rkg_batch_count(self.ctx.handle, self.handle, int(count_zeros)))

    def ingest(self):
        self.ctx._check(self.ctx.lib.rkg_batch_ingest(self.ctx.handle, self.handle))

    def set_stream(self, cuda_stream_ptr):
        self.ctx._check(self.ctx.lib.rkg_batch_set_stream(self.ctx.handle, self.handle, cuda_stream_ptr))

    def results(self, out=None):
        if out is None:
            out = np.zeros(max(self.nreads, 1), HIT_DTYPE)
        self.ctx._check(self.ctx.lib.rkg_batch_results(self.ctx.handle, self.handle, out.ctypes.data))
        return out[:self.nreads]

    def free(self):
        if self.handle:
            self.ctx.lib.rkg_batch_free(self.ctx.handle, self.handle)
            self.handle = None


class RkmhGpu:
    """One GPU context.  Mirrors the state main_stream / main_filter keep (k list, s, tables, refs)."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.rkg_create(int(device), C.byref(h))
        if rc != 0:
            raise RkmhError(f"rkg_create({device}) failed: {self.lib.rkg_last_error(None).decode()}")
        self.handle = h
        self.device = device
        self.ks, self.s, self.nrefs = None, None, 0

    # -- plumbing ---------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise RkmhError(self.lib.rkg_last_error(self.handle).decode())

    def close(self):
        if self.handle:
            self.lib.rkg_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @staticmethod
    def _buffers(bases, offs):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        return bases, offs

    # -- parameters / references ------------------------------------------------------------------
    def set_params(self, ks, sketch_size=1000, read_table_size=0, ref_table_size=0):
        ks = np.atleast_1d(np.asarray(ks, dtype=np.int32))
        self._check(self.lib.rkg_set_params(self.handle, ks.ctypes.data, len(ks), int(sketch_size),
                                            int(read_table_size), int(ref_table_size)))
        self.ks, self.s = [int(k) for k in ks], int(sketch_size)

    def load_refs(self, bases, offs, max_ref_occ=-1, per_sample_set=False):
        bases, offs = self._buffers(bases, offs)
        self._check(self.lib.rkg_load_refs(self.handle, bases.ctypes.data, offs.ctypes.data, len(offs) - 1,
                                           int(max_ref_occ), int(per_sample_set)))
        self.nrefs = len(offs) - 1

    def load_ref_sketches(self, sketches, lens):
        """Pre-hashed references: [nrefs, s] ascending non-zero rows + lengths (rkg_load_ref_sketches)."""
        sk = np.ascontiguousarray(sketches, dtype=np.uint64)
        ln = np.ascontiguousarray(lens, dtype=np.int32)
        assert sk.ndim == 2 and sk.shape == (len(ln), self.s)
        self._check(self.lib.rkg_load_ref_sketches(self.handle, sk.ctypes.data, ln.ctypes.data, len(ln)))
        self.nrefs = len(ln)

    def count_export(self, which=0):
        """Non-zero slots of a count table: (uint32 slots, int32 counts), unordered."""
        n = C.c_uint64()
        self._check(self.lib.rkg_count_export(self.handle, int(which), None, None, 0, C.byref(n)))
        slots, vals = np.zeros(max(n.value, 1), np.uint32), np.zeros(max(n.value, 1), np.int32)
        self._check(self.lib.rkg_count_export(self.handle, int(which), slots.ctypes.data, vals.ctypes.data, n.value, C.byref(n)))
        return slots[:n.value], vals[:n.value]

    def count_import(self, slots, vals, which=0):
        slots = np.ascontiguousarray(slots, dtype=np.uint32)
        vals = np.ascontiguousarray(vals, dtype=np.int32)
        self._check(self.lib.rkg_count_import(self.handle, int(which), slots.ctypes.data, vals.ctypes.data, len(slots)))

    def ref_sketches(self):
        sk = np.zeros((self.nrefs, self.s), np.uint64)
        lens = np.zeros(self.nrefs, np.int32)
        self._check(self.lib.rkg_get_ref_sketches(self.handle, sk.ctypes.data, lens.ctypes.data))
        return sk, lens

    def transfer_stats(self, reset=False):
        """Bytes / chunks the host-buffer calls sent as they are and packed (rkg_transfer_stats)."""
        st = (C.c_int64 * 4)()
        self._check(self.lib.rkg_transfer_stats(self.handle, st, int(reset)))
        return {"plain_bytes": int(st[0]), "packed_bytes": int(st[1]), "plain_chunks": int(st[2]), "packed_chunks": int(st[3])}

    def set_host_pack(self, enable=True, threads=0):
        """Pack bases to 2 bits + mask on the host before the host-to-device copies (rkg_set_host_pack):
        0 / False = off, 1 / True = every chunk, 2 = adaptive (a chunk is packed while the link is busy anyway)."""
        self._check(self.lib.rkg_set_host_pack(self.handle, int(enable), int(threads)))

    def set_memo(self, enable=True):
        """Memoised hashing for a single k <= 12 (include/rkmh_gpu.h); takes effect at the next load_refs."""
        self._check(self.lib.rkg_set_memo(self.handle, int(enable)))

    def memo_active(self):
        return bool(self.lib.rkg_memo_active(self.handle))

    # -- -M count table ---------------------------------------------------------------------------
    def count_reset(self):
        self._check(self.lib.rkg_count_reset(self.handle))

    def count_batch(self, bases, offs, count_zeros=False):
        bases, offs = self._buffers(bases, offs)
        self._check(self.lib.rkg_count_batch(self.handle, bases.ctypes.data, offs.ctypes.data, len(offs) - 1,
                                             int(count_zeros)))

    def count_table(self):
        """(device pointer, number of int32 slots) of the read count table."""
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self.lib.rkg_count_table(self.handle, C.byref(p), C.byref(n)))
        return p.value, n.value

    def count_table_tensor(self):
        """The read count table as a torch int32 CUDA tensor sharing the library's memory (for NCCL)."""
        import torch
        ptr, n = self.count_table()

        class _Mem:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 3}

        return torch.as_tensor(_Mem(), device=f"cuda:{self.device}")

    def count_partial_tensor(self, count_zeros=False):
        """The per-code partial counts of pass 1 (memoised path, 4^k + 32 int32 slots) as a torch int32 CUDA tensor
        sharing the library's memory: all-reduce THIS instead of the 200 M-slot table (include/rkmh_gpu.h)."""
        import torch
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self.lib.rkg_count_partial(self.handle, int(count_zeros), C.byref(p), C.byref(n)))
        ptr, n = p.value, n.value

        class _Mem:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 3}

        return torch.as_tensor(_Mem(), device=f"cuda:{self.device}")

    def count_lookup(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        out = np.zeros(max(keys.size, 1), np.int32)
        self._check(self.lib.rkg_count_lookup(self.handle, keys.ctypes.data, keys.size, out.ctypes.data))
        return out[:keys.size]

    # -- classification ---------------------------------------------------------------------------
    def classify_batch(self, bases, offs, min_occ=-1, out=None):
        """hash -> mask -> sketch -> intersect -> argmax for every read (host buffers in and out)."""
        bases, offs = self._buffers(bases, offs)
        n = len(offs) - 1
        if out is None:
            out = np.zeros(max(n, 1), HIT_DTYPE)
        self._check(self.lib.rkg_classify_batch(self.handle, bases.ctypes.data, offs.ctypes.data, n, int(min_occ),
                                                out.ctypes.data))
        return out[:n]

    def classify_raw(self, bases_ptr, offs_ptr, nreads, min_occ, out_ptr):
        self._check(self.lib.rkg_classify_batch(self.handle, bases_ptr, offs_ptr, nreads, int(min_occ), out_ptr))

    def classify_fixed_raw(self, bases_ptr, read_len, nreads, min_occ, out_ptr):
        self._check(self.lib.rkg_classify_fixed(self.handle, bases_ptr, int(read_len), int(nreads), int(min_occ), out_ptr))

    def classify_fixed(self, bases, read_len, min_occ=-1):
        """rkg_classify_fixed: every read has read_len bases (no offsets cross the link)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        n = bases.size // read_len if read_len else 0
        out = np.zeros(max(n, 1), HIT_DTYPE)
        self.classify_fixed_raw(bases.ctypes.data, read_len, n, min_occ, out.ctypes.data)
        return out[:n]

    def count_fixed(self, bases, read_len, count_zeros=False):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        n = bases.size // read_len if read_len else 0
        self._check(self.lib.rkg_count_fixed(self.handle, bases.ctypes.data, int(read_len), n, int(count_zeros)))

    def copy_probe_raw(self, bases_ptr, offs_ptr, read_len, nreads, out_ptr):
        self._check(self.lib.rkg_copy_probe(self.handle, bases_ptr, offs_ptr, int(read_len), int(nreads), out_ptr))

    def upload(self, bases, offs):
        bases, offs = self._buffers(bases, offs)
        h = C.c_void_p()
        self._check(self.lib.rkg_batch_upload(self.handle, bases.ctypes.data, offs.ctypes.data, len(offs) - 1, C.byref(h)))
        return Batch(self, h, len(offs) - 1)

    def synchronize(self):
        self._check(self.lib.rkg_synchronize(self.handle))

    # -- hpv16: full hash sets ---------------------------------------------------------------------
    def load_ref_sets(self, sets):
        """sets: list of uint64 arrays (any order; zeros / duplicates ignored) -> the hash -> set-bitset table."""
        so = np.zeros(len(sets) + 1, np.int64)
        so[1:] = np.cumsum([len(x) for x in sets])
        sh = np.ascontiguousarray(np.concatenate([np.asarray(x, np.uint64) for x in sets])) if so[-1] else np.zeros(1, np.uint64)
        self._check(self.lib.rkg_load_ref_sets(self.handle, sh.ctypes.data, so.ctypes.data, len(sets)))
        self.nrefs = len(sets)

    def classify_sets(self, bases, offs, ntypes, min_occ=-1):
        """-> (hits over sets [0, ntypes), counts [n, nsets - ntypes] of the remaining sets)"""
        bases, offs = self._buffers(bases, offs)
        n = len(offs) - 1
        nextra = self.nrefs - ntypes
        out = np.zeros(max(n, 1), HIT_DTYPE)
        extra = np.zeros((max(n, 1), max(nextra, 1)), np.int32)
        self._check(self.lib.rkg_classify_sets(self.handle, bases.ctypes.data, offs.ctypes.data, n, int(min_occ), int(ntypes),
                                               out.ctypes.data, extra.ctypes.data))
        return out[:n], extra[:n, :nextra]

    # -- call: exact k-mer depth -------------------------------------------------------------------
    def depth_begin(self, expected_windows):
        self._check(self.lib.rkg_depth_begin(self.handle, int(expected_windows)))

    def depth_add(self, bases, offs):
        bases, offs = self._buffers(bases, offs)
        self._check(self.lib.rkg_depth_add(self.handle, bases.ctypes.data, offs.ctypes.data, len(offs) - 1))

    def depth_of_sequence(self, seq):
        seq = np.ascontiguousarray(np.frombuffer(seq, np.uint8) if isinstance(seq, (bytes, bytearray)) else seq, dtype=np.uint8)
        out = np.zeros(max(seq.size - self.ks[0], 1), np.int32)
        self._check(self.lib.rkg_depth_of_sequence(self.handle, seq.ctypes.data, seq.size, out.ctypes.data))
        return out[:max(seq.size - self.ks[0], 0)]

    def depth_rescue(self, seq, windows):
        seq = np.ascontiguousarray(np.frombuffer(seq, np.uint8) if isinstance(seq, (bytes, bytearray)) else seq, dtype=np.uint8)
        windows = np.ascontiguousarray(windows, dtype=np.int64)
        out = np.zeros((max(len(windows), 1), 4 * self.ks[0]), np.int32)
        self._check(self.lib.rkg_depth_rescue(self.handle, seq.ctypes.data, seq.size, windows.ctypes.data, len(windows), out.ctypes.data))
        return out[:len(windows)]

    # -- hash command -----------------------------------------------------------------------------
    def sketch_batch(self, bases, offs, min_occ=-1):
        """minhashes(calc_hashes(seq)) per record: ([n, s] uint64, [n] int32 lengths)."""
        bases, offs = self._buffers(bases, offs)
        n = len(offs) - 1
        sk = np.zeros((max(n, 1), self.s), np.uint64)
        lens = np.zeros(max(n, 1), np.int32)
        self._check(self.lib.rkg_sketch_batch(self.handle, bases.ctypes.data, offs.ctypes.data, n, int(min_occ),
                                              sk.ctypes.data, lens.ctypes.data))
        return sk[:n], lens[:n]

    def calc_hashes(self, bases, offs):
        """Raw calc_hashes output of every record: (uint64 hashes, int64 offsets[n+1])."""
        bases, offs = self._buffers(bases, offs)
        n = len(offs) - 1
        lens = np.diff(offs)
        total = int(sum(np.maximum(lens - k, 0).sum() for k in self.ks))
        hashes = np.zeros(max(total, 1), np.uint64)
        hoffs = np.zeros(n + 1, np.int64)
        self._check(self.lib.rkg_hash_batch(self.handle, bases.ctypes.data, offs.ctypes.data, n, hashes.ctypes.data,
                                            hoffs.ctypes.data))
        return hashes[:total], hoffs

    def sketch_shard(self, shard_bases, kmer_begin, kmer_end, record_len):
        shard_bases = np.ascontiguousarray(shard_bases, dtype=np.uint8)
        sk = np.zeros(self.s, np.uint64)
        n = C.c_int32()
        self._check(self.lib.rkg_sketch_shard(self.handle, shard_bases.ctypes.data, shard_bases.size, int(kmer_begin),
                                              int(kmer_end), int(record_len), sk.ctypes.data, C.byref(n)))
        return sk[:n.value]

    def merge_sketches(self, parts, lens):
        parts = np.ascontiguousarray(parts, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        sk = np.zeros(self.s, np.uint64)
        n = C.c_int32()
        self._check(self.lib.rkg_merge_sketches(self.handle, parts.ctypes.data, lens.ctypes.data, len(lens),
                                                sk.ctypes.data, C.byref(n)))
        return sk[:n.value]

    # -- measurement ------------------------------------------------------------------------------
    def launch_count(self):
        return int(self.lib.rkg_launch_count(self.handle))

    def profile_enable(self, on=True):
        self._check(self.lib.rkg_profile_enable(self.handle, int(on)))

    def profile_read(self, reset=True):
        arr = (_lib.KernelTime * 16)()
        n = self.lib.rkg_profile_read(self.handle, arr, 16, int(reset))
        return {arr[i].name.decode(): {"launches": int(arr[i].launches), "ms": float(arr[i].ms)} for i in range(min(n, 16))}


# -- output formatting: the reference's line formats ------------------------------------------------
def stream_line(ref_name, read_name, hit, sketch_size, min_matches=-1, min_diff=0):
    """One output line of `rkmh stream` / `classify` (reference src/rkmh.cpp:884-893)."""
    best, diff = int(hit["best"]), int(hit["best"]) - int(hit["prev"])
    depth_fail = int(hit["sketch_len"]) <= min_matches
    match_fail = best < min_matches
    diff_ok = diff > min_diff
    return (ref_name + b"\t" + read_name + b"\t" + str(best).encode() + b"\t" + str(sketch_size).encode() +
            (b"FAIL:DEPTH" if depth_fail else b"") + b"\t" + (b"FAIL:MATCHES" if match_fail else b"") + b"\t" +
            (b"" if diff_ok else b"FAIL:DIFF") + b"\n")


def filter_pass(hit, min_matches=-1, min_diff=0):
    """`rkmh filter` keeps a read iff this holds (reference src/rkmh.cpp:1289-1298, src/equiv.hpp:328-352)."""
    best, prev = max(int(hit["best"]), 0), max(int(hit["prev"]), 0)
    return int(hit["sketch_len"]) > 0 and not (best < min_matches) and (best - prev > min_diff)


# -- whole commands (Python mirror of main_stream / main_filter / main_hash) ---------------------------
def run_stream(ref_files, read_files, ks, sketch_size=1000, min_kmer_occ=-1, min_matches=-1, min_diff=0,
               max_samples=None, device=0, ctx=None):
    """`rkmh stream|classify -r ... -f ... -k ... -s ... [-M -N -D -I]` -> stdout bytes
    (reference src/rkmh.cpp:584-989)."""
    from .fastx import load_sequences
    ref_names, ref_seqs, _ = load_sequences(ref_files)
    read_names, read_seqs, _ = load_sequences(read_files)
    own = ctx is None
    ctx = ctx or RkmhGpu(device)
    try:
        ctx.set_params(ks or [16], sketch_size,
                       read_table_size=STREAM_TABLE_SIZE if min_kmer_occ != -1 else 0,
                       ref_table_size=STREAM_TABLE_SIZE if max_samples is not None else 0)
        rb, ro = pack_sequences(ref_seqs)
        ctx.load_refs(rb, ro, max_ref_occ=-1 if max_samples is None else max_samples, per_sample_set=False)
        qb, qo = pack_sequences(read_seqs)
        use_m = min_kmer_occ != -1          # -M given (doReadDepth, src/rkmh.cpp:696-699)
        if use_m:
            ctx.count_reset()
            ctx.count_batch(qb, qo, count_zeros=False)
        hits = ctx.classify_batch(qb, qo, min_occ=min_kmer_occ if use_m else -1)
        out = [stream_line(ref_names[int(h["ref_id"])], read_names[i], h, sketch_size, min_matches, min_diff)
               for i, h in enumerate(hits)]
        return b"".join(out)
    finally:
        if own:
            ctx.close()


def run_filter(ref_files, read_files, ks, sketch_size=1000, min_kmer_occ=-1, min_matches=-1, min_diff=0,
               max_samples=100000, device=0, ctx=None):
    """`rkmh filter ...` -> stdout bytes (reference src/rkmh.cpp:996-1424, intended flow)."""
    from .fastx import load_sequences
    ref_names, ref_seqs, _ = load_sequences(ref_files)
    read_names, read_seqs, read_quals = load_sequences(read_files)
    own = ctx is None
    ctx = ctx or RkmhGpu(device)
    try:
        use_m = min_kmer_occ != -1
        use_i = max_samples < 100000
        ctx.set_params(ks or [16], sketch_size, read_table_size=FILTER_TABLE_SIZE if use_m else 0,
                       ref_table_size=FILTER_TABLE_SIZE if use_i else 0)
        rb, ro = pack_sequences(ref_seqs)
        ctx.load_refs(rb, ro, max_ref_occ=max_samples if use_i else -1, per_sample_set=True)
        qb, qo = pack_sequences(read_seqs)
        if use_m:
            ctx.count_reset()
            ctx.count_batch(qb, qo, count_zeros=True)
        hits = ctx.classify_batch(qb, qo, min_occ=min_kmer_occ if use_m else -1)
        out = []
        for i, h in enumerate(hits):
            if filter_pass(h, min_matches, min_diff):
                seq = bytes((c - 32) if c > 91 and c < 128 else c for c in read_seqs[i])  # to_upper, mkmh.hpp:231
                out.append(b">" + read_names[i] + b"\n" + seq + b"\n+\n" + read_quals[i] + b"\n")
        return b"".join(out)
    finally:
        if own:
            ctx.close()


def bind_host_thread_to_gpu(device):
    """Pin the calling process to the CPU cores that are local to CUDA device `device` (NVML's ideal CPU set), so that
    the pinned host buffers allocated afterwards -- and the threads that fill them -- sit on the GPU's own NUMA node.
    On an 8-GPU box every rank otherwise pulls its reads across the socket interconnect.  Returns the CPU list, or
    None when NVML or the topology information is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device)
        handle = None
        if all(hasattr(props, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bus = "%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
            try:
                handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                handle = None
        if handle is None:
            handle = pynvml.nvmlDeviceGetHandleByIndex(int(device))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None
