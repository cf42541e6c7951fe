"""Every function declared in include/*.h is exported by the matching shared library (no compute
calls: this runs without a GPU), and the product never falls back to the CPU."""
import ctypes
import os
import re

import pytest
from conftest import ROOT

import p2p_b200

LIBS = {
    "p2p_b200.h": "libp2p_b200.so",
    "p2p_host.h": "libp2p_host.so",
    "photoNs_CUDA_indexing.h": "libphotoNs_CUDA_indexing.so",
    "photoNs_CUDA_redundant.h": "libphotoNs_CUDA_redundant.so",
}


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    body = src[src.index('extern "C"'):]
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", body)
    return sorted(set(n for n in names if n not in ("defined",)))


@pytest.mark.parametrize("header", sorted(LIBS))
def test_exports(header):
    lib = ctypes.CDLL(os.path.join(p2p_b200.LIB_DIR, LIBS[header]))
    names = _declared(header)
    assert len(names) >= 7, names
    for n in names:
        assert hasattr(lib, n), f"{LIBS[header]} does not export {n}"


def test_reference_symbol_lists():
    """The compat headers declare exactly the reference's entry points."""
    idx = set(_declared("photoNs_CUDA_indexing.h"))
    assert {"initGPU", "getGPUMemoryState", "allocMemGPU", "copyMemGPU", "LaunchKernelP2PIndexing", "readResultsGPU"} <= idx
    red = set(_declared("photoNs_CUDA_redundant.h"))
    assert {"initGPU", "getGPUMemoryState", "allocMemGPU", "copyMemGPU", "readResultsGPU", "LaunchKernelP2PDualNaive",
            "allocAndCopySelfInteractionsGPU", "LaunchKernelP2PSelfInteractions", "readResultsGPUSelfInteractions"} <= red


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product must fail loudly (P2P_ERR_NODEVICE), never compute on the CPU."""
    if p2p_b200.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(p2p_b200.P2PError) as e:
        p2p_b200.P2PContext(0)
    assert e.value.code == -4


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "photons-2.0_gpu-p2p-redundancy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f)).read()
                assert "p2p_oracle" not in txt and "import oracle" not in txt and "import flow" not in txt, os.path.join(dirpath, f)
