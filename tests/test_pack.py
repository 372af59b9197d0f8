"""Host side of the packed transfer (rkmh_b200/csrc/rk_pack.cpp, exported by librkmh_gpu.so as rk_pack_bases): the 2-bit
codes and the invalid-base mask the host makes must be the ones ingest_kernel makes on the device, whatever the vector
path (AVX-512BW / AVX2 / scalar), the thread count and the alignment of the tail; and unpack_kernel (through the kernel
emulator) must turn them back into the upper-cased sequence with every invalid byte as N.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

from rkmh_b200 import _lib


def _expected(b):
    """codes (2 x u32 per 32 bases, base i at bits 2 i) and mask (bit i set: not one of ACGTacgt), bases beyond len invalid"""
    n = b.size
    nw = (n + 31) // 32
    pad = np.zeros(nw * 32, np.uint8)
    pad[:n] = b
    code = ((pad >> 1) & 3).astype(np.uint64).reshape(nw, 32)
    valid = np.isin(pad, np.frombuffer(b"ACGTacgt", np.uint8)).reshape(nw, 32)
    sh = (2 * np.arange(32, dtype=np.uint64))[None, :]
    c64 = np.bitwise_or.reduce(code << sh, axis=1)
    codes = np.empty(2 * nw, np.uint32)
    codes[0::2] = (c64 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    codes[1::2] = (c64 >> np.uint64(32)).astype(np.uint32)
    mask = np.bitwise_or.reduce((~valid).astype(np.uint64) << np.arange(32, dtype=np.uint64)[None, :], axis=1).astype(np.uint32)
    return codes, mask


@pytest.fixture(scope="module")
def pack():
    lib = C.CDLL(_lib.LIB_PATH)
    lib.rk_pack_bases.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
    lib.rk_pack_bases.restype = None

    def run(b, threads):
        nw = (b.size + 31) // 32
        # 64-byte aligned destinations with the mask at a 32-byte multiple, as the library's staging buffers are
        raw = np.zeros(3 * nw + 64, np.uint32)
        off = (-raw.ctypes.data // 4) % 16
        mask_at = (2 * nw + 7) & ~7
        codes = raw[off:off + 2 * nw]
        mask = raw[off + mask_at:off + mask_at + nw]
        lib.rk_pack_bases(b.ctypes.data, C.c_int64(b.size), codes.ctypes.data, mask.ctypes.data, threads)
        return codes.copy(), mask.copy()
    return run


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 63, 64, 65, 1000, 4096 * 32 + 17, 1_000_003])
def test_pack_matches_ingest_semantics(pack, n):
    rng = np.random.default_rng(n)
    alphabet = np.frombuffer(b"ACGTacgtNnRYKM-*{}|~@[`\x00\xff\x80", np.uint8)
    b = rng.choice(alphabet, size=n, p=None) if n else np.zeros(0, np.uint8)
    if n > 100:
        b[:n // 2] = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=n // 2)      # long valid stretches too
    ec, em = _expected(b)
    for threads in (1, 3, 8):
        codes, mask = pack(b, threads)
        assert np.array_equal(mask, em), (n, threads)
        # the code of an invalid base is not specified beyond "two bits of the byte": compare where the kernels look
        assert np.array_equal(codes, ec), (n, threads)


def test_unpack_kernel_restores_upper_case_with_n(pack):
    from tests.emu.emu import Emu
    rng = np.random.default_rng(5)
    b = rng.choice(np.frombuffer(b"ACGTacgtNnXy-", np.uint8), size=5000)
    codes, mask = pack(b, 2)
    got = Emu().unpack(codes, mask)[:b.size]
    up = b.copy()
    low = (up >= ord("a")) & (up <= ord("z"))
    up[low] -= 32
    want = np.where(np.isin(up, np.frombuffer(b"ACGT", np.uint8)), up, ord("N")).astype(np.uint8)
    assert np.array_equal(got, want)
