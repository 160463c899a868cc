"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/flexi_gpu.h declares,
agrees with the ctypes mirror of flexi_gpu_desc, and fails loudly (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re

import pytest

from flexi_b200 import gpu as G
from flexi_b200.desc import ABI_VERSION, FlexiGpuDesc
from flexi_b200.hostmesh import basis as B
from flexi_b200.hostmesh.case import build_case, tgv_box
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(G.LIB_PATH):
        G.build()
    return G.lib()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "flexi_gpu.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(flexi_gpu_[a-z_0-9]+)\s*\(", hdr, flags=re.M))
    assert declared == set(G.EXPORTS), declared ^ set(G.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None, name


def test_abi_version_and_desc_size(lib):
    assert lib.flexi_gpu_abi_version() == ABI_VERSION
    # the oracle is compiled from the same header: its sizeof(flexi_gpu_desc) pins the ctypes mirror
    assert O._load().oracle_desc_bytes() == C.sizeof(FlexiGpuDesc)


def test_null_and_mismatched_descriptors_are_rejected(lib):
    h = C.c_void_p()
    assert lib.flexi_gpu_init(None, C.byref(h)) == -1
    d = FlexiGpuDesc()
    d.abi_version, d.struct_bytes = ABI_VERSION + 1, C.sizeof(FlexiGpuDesc)
    assert lib.flexi_gpu_init(C.byref(d), C.byref(h)) == -1
    assert b"ABI mismatch" in lib.flexi_gpu_last_error()


def test_no_cpu_fallback(lib):
    """Without a CUDA device init must fail with a CUDA error instead of computing anything on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = build_case(tgv_box(2), 2, B.NODE_GAUSS, dict(parabolic=False))[0]
    with pytest.raises(G.FlexiGpuError):
        G.FlexiGpu(r)


def test_unsupported_options_are_refused(lib):
    r = build_case(tgv_box(2), 2, B.NODE_GAUSS, dict(parabolic=False))[0]
    r.desc.doWeakLifting = 1
    h = C.c_void_p()
    assert lib.flexi_gpu_init(C.byref(r.desc), C.byref(h)) == -4
