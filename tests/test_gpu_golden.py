"""The reference's own unit tests (unitTests/ProlongToFace.f90, SurfInt.f90) replayed through the C ABI on the GPU:
flexi_gpu_prolong_to_face / flexi_gpu_surf_int against the committed .bin fixtures, tolerance 100*eps like the reference."""
import types

import numpy as np
import pytest

from fortran_bin import UnittestElementData3D, f64, read_records
from flexi_b200.gpu import FlexiGpu
from test_oracle_golden import almost_equal_abs_or_rel, unit_element_desc

pytestmark = pytest.mark.gpu


def _gpu(node_type):
    d = UnittestElementData3D()
    desc, keep = unit_element_desc(node_type, d)
    return FlexiGpu(types.SimpleNamespace(desc=desc, keep=keep)), d


@pytest.mark.parametrize("node_type,fname", [(1, "ProlongToFace_G3D.bin"), (2, "ProlongToFace_GL3D.bin")])
def test_prolong_to_face_bin_gpu(node_type, fname):
    g, d = _gpu(node_type)
    (u,) = read_records("ProlongToFaceUvol.bin")
    Uvol = f64(u, (10, 10, 10, 1))
    Uvol_nVar = np.zeros((1, 10, 10, 10, 5))
    for v in range(5):
        Uvol_nVar[0, :, :, :, v] = np.transpose(Uvol[:, :, :, 0], (2, 1, 0))
    ref = f64(read_records(fname)[0], (10, 10, 6))
    um, us = np.zeros((6, 10, 10, 5)), np.full((6, 10, 10, 5), 7.0)
    g.ProlongToFace(Uvol_nVar, um, us, doMPISides=False)
    g.close()
    for v in range(5):
        assert almost_equal_abs_or_rel(np.transpose(um[:, :, :, v], (2, 1, 0)), ref), (fname, v)
    assert np.all(us == 7.0)  # no slave element on these sides: untouched, like in the reference


# SurfInt_G3D_Split.bin: the Gauss split-form fixture with a constant state, where the two-point correction vanishes (see
# tests/test_oracle_golden.py): expected Ut = the plain Gauss surface integral
@pytest.mark.parametrize("node_type,fname", [(1, "SurfInt_G3D.bin"), (2, "SurfInt_GL3D.bin"), (1, "SurfInt_G3D_Split.bin")])
def test_surf_int_bin_gpu(node_type, fname):
    g, d = _gpu(node_type)
    (fl,) = read_records("SurfIntFlux.bin")
    Flux = f64(fl, (10, 10, 6))
    Flux_nVar = np.zeros((6, 10, 10, 5))
    for v in range(5):
        Flux_nVar[:, :, :, v] = np.transpose(Flux, (2, 1, 0))
    ref = f64(read_records(fname)[0], (10, 10, 10, 1))[:, :, :, 0]
    Ut = np.zeros((1, 10, 10, 10, 5))
    g.SurfInt(Flux_nVar, Flux_nVar, Ut, doMPISides=False)
    g.close()
    for v in range(5):
        assert almost_equal_abs_or_rel(np.transpose(Ut[0, :, :, :, v], (2, 1, 0)), ref), (fname, v)


def test_domisides_true_touches_nothing_on_single_rank():
    g, d = _gpu(1)
    um, us = np.full((6, 10, 10, 5), 3.0), np.full((6, 10, 10, 5), 4.0)
    g.ProlongToFace(np.ones((1, 10, 10, 10, 5)), um, us, doMPISides=True)
    g.close()
    assert np.all(um == 3.0) and np.all(us == 4.0)
