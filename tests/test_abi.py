"""CPU tests of the boundary: the C-ABI library builds, loads and exports every symbol that
include/rkmh_gpu.h declares; without a GPU every compute entry fails loudly (no fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

from tests.util import ROOT


@pytest.fixture(scope="module")
def lib_path():
    path = os.path.join(ROOT, "rkmh_b200", "librkmh_gpu.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "rkmh_b200", "csrc")], stdout=subprocess.DEVNULL)
    return path


def declared_symbols():
    with open(os.path.join(ROOT, "include", "rkmh_gpu.h")) as fh:
        text = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(rkg_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rkmh_gpu.h but not exported"


def test_python_binding_declares_every_symbol(lib_path):
    from rkmh_b200 import _lib
    assert sorted(_lib.RKG_SYMBOLS) == declared_symbols()
    lib = _lib.load()
    for n in _lib.RKG_SYMBOLS:
        assert getattr(lib, n).argtypes is not None or n == "rkg_destroy"


def test_sass_is_sm100a_only(lib_path):
    out = subprocess.run(["cuobjdump", "--list-elf", lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_silent_cpu_fallback(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    import rkmh_b200
    with pytest.raises(rkmh_b200.RkmhError, match="no CPU fallback"):
        rkmh_b200.RkmhGpu(0)


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "rkmh_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                with open(os.path.join(base, f), errors="ignore") as fh:
                    txt = fh.read()
                if re.search(r"oracle[/.]|rkmh_oracle|liboracle|rko_", txt):
                    bad.append(os.path.join(base, f))
    assert not bad, bad


def test_numa_binding_helper_is_harmless_without_a_gpu():
    """bind_host_thread_to_gpu never raises and never empties the affinity mask (None when NVML / the GPU is missing)."""
    import os
    import rkmh_b200
    before = os.sched_getaffinity(0)
    got = rkmh_b200.bind_host_thread_to_gpu(0)
    after = os.sched_getaffinity(0)
    assert after and after <= before
    assert got is None or set(got) == after
    os.sched_setaffinity(0, before)
