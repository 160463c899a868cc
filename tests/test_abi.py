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


@pytest.mark.parametrize("riemann", ["hll", "hlle", "hllem", "hllc"])
def test_split_dg_rejects_hll_family(lib, riemann):
    """SPLIT_DG builds do not offer the HLL family (riemann.f90:204-246): flexi_gpu_init refuses the descriptor before it touches
    the device, as the reference's CollectiveStop('Riemann solver not defined!') does."""
    from flexi_b200.gpu import FlexiGpu, FlexiGpuError
    eqn = dict(parabolic=False, split_dg=True, split_dg_type="pi", riemann=riemann)
    r = build_case(tgv_box(2), 3, B.NODE_GAUSS_LOBATTO, eqn)[0]
    with pytest.raises(FlexiGpuError, match="not selectable with SPLIT_DG"):
        FlexiGpu(r)


@pytest.mark.parametrize("riemann", ["average", "ch", "ir"])
def test_split_only_solvers_rejected_without_split_dg(lib, riemann):
    """Riemann_FluxAverage / _CH / _IR exist in SPLIT_DG builds only (riemann.f90:54-58,211-220)."""
    from flexi_b200.gpu import FlexiGpu, FlexiGpuError
    r = build_case(tgv_box(2), 3, B.NODE_GAUSS_LOBATTO, dict(parabolic=False, riemann=riemann))[0]
    with pytest.raises(FlexiGpuError, match="unknown Riemann solver"):
        FlexiGpu(r)


def test_undefined_split_formulation_is_rejected(lib):
    """splitflux.f90:150-152: 'SplitDG formulation not defined!'"""
    from flexi_b200.gpu import FlexiGpu, FlexiGpuError
    r = build_case(tgv_box(2), 3, B.NODE_GAUSS_LOBATTO, dict(parabolic=False, split_dg=True, split_dg_type="pi", riemann="lf"))[0]
    r.desc.splitDGType = 7
    with pytest.raises(FlexiGpuError, match="SplitDG formulation not defined"):
        FlexiGpu(r)
