"""ctypes binding of rkmh_b200/librkmh_gpu.so (the C ABI declared in include/rkmh_gpu.h).

There is no fallback of any kind: if the shared library is missing or no B200 is visible, loading or
context creation raises.  Build the library with ``python -c 'import __graft_entry__ as g; g.build()'``
or ``make -C rkmh_b200/csrc``.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librkmh_gpu.so")

RKG_SYMBOLS = [
    "rkg_create", "rkg_destroy", "rkg_last_error", "rkg_set_params", "rkg_host_alloc", "rkg_host_free",
    "rkg_load_refs", "rkg_get_ref_sketches", "rkg_count_reset", "rkg_count_batch", "rkg_count_table",
    "rkg_count_lookup", "rkg_classify_batch", "rkg_batch_upload", "rkg_batch_classify", "rkg_batch_count",
    "rkg_batch_results", "rkg_batch_ingest", "rkg_batch_set_stream", "rkg_batch_free", "rkg_synchronize", "rkg_sketch_batch", "rkg_hash_batch",
    "rkg_sketch_shard", "rkg_merge_sketches", "rkg_launch_count", "rkg_profile_enable", "rkg_profile_read",
    "rkg_set_memo", "rkg_memo_active", "rkg_count_partial", "rkg_classify_fixed", "rkg_count_fixed", "rkg_copy_probe",
    "rkg_load_ref_sets", "rkg_classify_sets", "rkg_load_ref_sketches", "rkg_count_export", "rkg_count_import",
    "rkg_depth_begin", "rkg_depth_add", "rkg_depth_of_sequence", "rkg_depth_rescue", "rkg_set_host_pack", "rkg_transfer_stats",
]


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char_p), ("launches", C.c_int64), ("ms", C.c_double)]


_lib = None


def load():
    """Load the shared library once and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C rkmh_b200/csrc` "
                           "(there is no CPU fallback for the rkmh hot path)")
    lib = C.CDLL(LIB_PATH)
    P, I, L, U, Z = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_size_t
    PP = C.POINTER(C.c_void_p)
    lib.rkg_create.argtypes = [I, PP]
    lib.rkg_destroy.argtypes = [P]
    lib.rkg_destroy.restype = None
    lib.rkg_last_error.argtypes = [P]
    lib.rkg_last_error.restype = C.c_char_p
    lib.rkg_set_params.argtypes = [P, P, I, I, U, U]
    lib.rkg_host_alloc.argtypes = [P, Z, PP]
    lib.rkg_host_free.argtypes = [P, P]
    lib.rkg_load_refs.argtypes = [P, P, P, I, I, I]
    lib.rkg_get_ref_sketches.argtypes = [P, P, P]
    lib.rkg_count_reset.argtypes = [P]
    lib.rkg_count_batch.argtypes = [P, P, P, L, I]
    lib.rkg_count_table.argtypes = [P, PP, C.POINTER(U)]
    lib.rkg_count_partial.argtypes = [P, I, PP, C.POINTER(U)]
    lib.rkg_count_lookup.argtypes = [P, P, L, P]
    lib.rkg_classify_fixed.argtypes = [P, P, L, L, I, P]
    lib.rkg_count_fixed.argtypes = [P, P, L, L, I]
    lib.rkg_copy_probe.argtypes = [P, P, P, L, L, P]
    lib.rkg_transfer_stats.argtypes = [P, P, C.c_int]
    lib.rkg_load_ref_sets.argtypes = [P, P, P, I]
    lib.rkg_load_ref_sketches.argtypes = [P, P, P, I]
    lib.rkg_count_export.argtypes = [P, I, P, P, U, C.POINTER(U)]
    lib.rkg_count_import.argtypes = [P, I, P, P, U]
    lib.rkg_set_host_pack.argtypes = [P, I, I]
    lib.rkg_depth_begin.argtypes = [P, U]
    lib.rkg_depth_add.argtypes = [P, P, P, L]
    lib.rkg_depth_of_sequence.argtypes = [P, P, L, P]
    lib.rkg_depth_rescue.argtypes = [P, P, L, P, L, P]
    lib.rkg_classify_sets.argtypes = [P, P, P, L, I, I, P, P]
    lib.rkg_classify_batch.argtypes = [P, P, P, L, I, P]
    lib.rkg_batch_upload.argtypes = [P, P, P, L, PP]
    lib.rkg_batch_classify.argtypes = [P, P, I]
    lib.rkg_batch_count.argtypes = [P, P, I]
    lib.rkg_batch_ingest.argtypes = [P, P]
    lib.rkg_batch_set_stream.argtypes = [P, P, P]
    lib.rkg_batch_results.argtypes = [P, P, P]
    lib.rkg_batch_free.argtypes = [P, P]
    lib.rkg_synchronize.argtypes = [P]
    lib.rkg_sketch_batch.argtypes = [P, P, P, L, I, P, P]
    lib.rkg_hash_batch.argtypes = [P, P, P, L, P, P]
    lib.rkg_sketch_shard.argtypes = [P, P, L, L, L, L, P, P]
    lib.rkg_merge_sketches.argtypes = [P, P, P, I, P, P]
    lib.rkg_set_memo.argtypes = [P, I]
    lib.rkg_memo_active.argtypes = [P]
    lib.rkg_launch_count.argtypes = [P]
    lib.rkg_launch_count.restype = L
    lib.rkg_profile_enable.argtypes = [P, I]
    lib.rkg_profile_read.argtypes = [P, C.POINTER(KernelTime), I, I]
    _lib = lib
    return lib
