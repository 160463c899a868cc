"""Pins the synthetic host's basis + mappings against the reference's own unit-test fixtures.

Mirrors unitTests/NodesAndWeights.f90, DerivativeMatrix.f90, Vandermonde.f90, ChangeBasis.f90 and the L_*/S2V/S2V2
payload of UnittestElementData3D.bin (unitTests/unittest.f90:29-101). Tolerance is the reference's own
ALMOSTEQUALABSORREL(x, ref, 100*eps) (unitTests/NodesAndWeights.f90:68-70).
"""
import numpy as np
import pytest

from fortran_bin import UnittestElementData3D, f64, read_records
from flexi_b200.hostmesh import basis as B
from flexi_b200.hostmesh import mappings as M

TYPES = [B.NODE_GAUSS, B.NODE_GAUSS_LOBATTO, B.NODE_CL, B.NODE_VISU]
TOL = 100.0 * np.finfo(float).eps


def almost_equal_abs_or_rel(x, ref, tol=TOL):
    return np.all((np.abs(x - ref) <= tol) | (np.abs(x - ref) <= tol * np.abs(ref)))


def test_nodes_and_weights_bin():
    (rec,) = read_records("NodesAndWeights.bin")
    a = f64(rec)
    n = 11 * 10 * 4
    xi = a[:n].reshape((11, 10, 4), order="F")
    w = a[n : 2 * n].reshape((11, 10, 4), order="F")
    wb = a[2 * n :].reshape((11, 10, 4), order="F")
    for it, t in enumerate(TYPES):
        for N in range(1, 11):
            x_, w_, wb_ = B.get_nodes_and_weights(N, t)
            assert almost_equal_abs_or_rel(x_, xi[: N + 1, N - 1, it]), (t, N)
            assert almost_equal_abs_or_rel(w_, w[: N + 1, N - 1, it]), (t, N)
            assert almost_equal_abs_or_rel(wb_, wb[: N + 1, N - 1, it]), (t, N)


def test_derivative_matrix_bin():
    (rec,) = read_records("DerivativeMatrix.bin")
    D = f64(rec, (11, 11, 40))
    k = 0
    for t in TYPES:
        for N in range(1, 11):
            x_, _, _ = B.get_nodes_and_weights(N, t)
            D_ = B.polynomial_derivative_matrix(x_)
            assert almost_equal_abs_or_rel(D_, D[: N + 1, : N + 1, k]), (t, N)
            k += 1


def test_vandermonde_bin():
    # unitTests/Vandermonde.f90: loops iIn=2..5, iOut=3..6, typeIn, typeOut, modal; stores In->Out and Out->In
    recs = read_records("Vandermonde.bin")
    a = f64(recs[0])
    nio = 7 * 6 * 512
    # WRITE(10) Vdm_Out_In,Vdm_In_Out (unitTests/Vandermonde.f90:86)
    V_oi = a[:nio].reshape((6, 7, 512), order="F")
    V_io = a[nio:].reshape((7, 6, 512), order="F")
    k = 0
    checked = 0
    for iIn in range(2, 6):
        for iOut in range(3, 7):
            for tin in TYPES:
                for tout in TYPES:
                    for modal in (False, True):
                        V = B.get_vandermonde(iIn, tin, iOut, tout, modal=modal)
                        ref = V_io[: iOut + 1, : iIn + 1, k]
                        assert almost_equal_abs_or_rel(V, ref, 1000 * TOL), (iIn, iOut, tin, tout, modal)
                        # Out->In (interpolation.f90:352-360): modal projection only if N_In < N_Out
                        Vb = B.get_vandermonde(iOut, tout, iIn, tin, modal=modal)
                        if tin.lower() == tout.lower() and iIn == iOut:
                            Vb = np.eye(iIn + 1)
                        refb = V_oi[: iIn + 1, : iOut + 1, k]
                        assert almost_equal_abs_or_rel(Vb, refb, 1000 * TOL), ("oi", iIn, iOut, tin, tout, modal)
                        k += 1
                        checked += 1
    assert checked == 512


def test_change_basis_bin():
    # unitTests/ChangeBasis.f90:45-100: synthetic Vdm/UIn, UOut(:,:,:,:,:,1) = ChangeBasis3D(nVar,NIn,NOut,Vdm,UIn)
    nVar, NIn, NOut, nElems = 3, 4, 5, 6
    Vdm = np.zeros((NOut + 1, NIn + 1))
    z = 1
    for q in range(NIn + 1):
        for p in range(NOut + 1):
            z += 1
            Vdm[p, q] = 1.0 + 1.0 / z
    UIn = np.zeros((nElems, NIn + 1, NIn + 1, NIn + 1, nVar))
    z = 1
    for e in range(nElems):
        for k in range(NIn + 1):
            for j in range(NIn + 1):
                for i in range(NIn + 1):
                    for v in range(nVar):
                        z += 1
                        UIn[e, k, j, i, v] = 1.0 + 0.1 * z
    recs = read_records("ChangeBasis.bin")
    a = f64(recs[0])
    n3 = nVar * (NOut + 1) ** 3 * nElems * 4
    UOut = a[:n3].reshape((nVar, NOut + 1, NOut + 1, NOut + 1, nElems, 4), order="F")
    # slot 3, element 1: CALL ChangeBasis3D(nVar,NIn,NOut,Vdm,UIn(:,:,:,:,1),UOut(:,:,:,:,1,3)) (ChangeBasis.f90:83)
    mine = B.change_basis_3d(Vdm, UIn[0])  # [k,j,i,v]
    ref = np.transpose(UOut[:, :, :, :, 0, 2], (3, 2, 1, 0))
    assert np.max(np.abs(mine - ref) / np.abs(ref)) < 50 * np.finfo(float).eps


def test_unittest_element_data_basis_and_mappings():
    d = UnittestElementData3D()
    assert d.nElems == 1 and d.nSides == 6
    # the fixture was generated with Gauss nodes, N=9 (L_Minus(0) = 1.588...)
    b = B.DGBasis(9, B.NODE_GAUSS)
    assert almost_equal_abs_or_rel(b.L_Minus, d.L_Minus)
    assert almost_equal_abs_or_rel(b.L_Plus, d.L_Plus)
    assert almost_equal_abs_or_rel(b.L_HatMinus, d.L_HatMinus)
    assert almost_equal_abs_or_rel(b.L_HatPlus, d.L_HatPlus)
    V2S, S2V, S2V2 = M.build_mappings(9)
    assert np.array_equal(S2V2, d.S2V2)
    assert np.array_equal(S2V, d.S2V)


@pytest.mark.parametrize("N", [1, 3, 7])
def test_mappings_inverse(N):
    V2S, S2V, S2V2 = M.build_mappings(N)
    for f in range(5):
        for s in range(6):
            for i in range(N + 1):
                for j in range(N + 1):
                    for k in range(N + 1):
                        p, q, l = V2S[:, i, j, k, f, s]
                        assert tuple(S2V[:, l, p, q, f, s]) == (i, j, k)


@pytest.mark.parametrize("node", [B.NODE_GAUSS, B.NODE_GAUSS_LOBATTO])
def test_dg_basis_identities(node):
    b = B.DGBasis(7, node, split_dg=True)
    # derivative of a constant vanishes; D_Hat = -M^-1 D^T M  (dg.f90:177)
    assert np.max(np.abs(b.D.sum(axis=1))) < 1e-12
    for i in range(8):
        for j in range(8):
            assert abs(b.D_Hat[i, j] + b.D[j, i] * b.wGP[j] / b.wGP[i]) < 1e-13
    if node == B.NODE_GAUSS_LOBATTO:
        # dg.f90:196-201: DVolSurf has a zero diagonal on LGL nodes
        assert np.max(np.abs(np.diag(b.DVolSurf))) < 1e-12
        assert b.L_Minus[0] == 1.0 and np.all(b.L_Minus[1:] == 0.0)


def test_overintegration_mat_and_cfl_scaling():
    """Host mirror of InitOverintegration (cut-off, dg/overintegration.f90:104-116) and of the time-step scaling it implies:
    NEff = MIN(PP_N, NUnder) in the advective part only, Gauss table on Gauss-Lobatto nodes (timedisc_func.f90:142-144,
    timedisc_vars.f90:153-160), viscous part on PP_N (timedisc_func.f90:481)."""
    from flexi_b200.hostmesh.equation import Equation

    N, NU = 7, 5
    x, w, _ = B.get_nodes_and_weights(N, B.NODE_GAUSS_LOBATTO)
    M = B.overintegration_mat(x, NU)
    V, sV = B.build_legendre_vdm(x)
    # a polynomial of degree <= NUnder passes unchanged, Legendre modes above NUnder are removed
    for deg in range(N + 1):
        f = V[:, deg]
        assert np.max(np.abs(M @ f - (f if deg <= NU else 0.0))) < 1e-13
    # the projector is L2-orthogonal with respect to Gauss quadrature, not to the LGL quadrature; it always keeps the mean
    xg, wg, _ = B.get_nodes_and_weights(N, B.NODE_GAUSS)
    Mg = B.overintegration_mat(xg, NU)
    assert np.max(np.abs(wg @ Mg - wg)) < 1e-14
    assert np.max(np.abs(Mg.T * wg[None, :] - (Mg.T * wg[None, :]).T)) < 1e-14  # W M symmetric on Gauss nodes
    plain = Equation(N, B.NODE_GAUSS_LOBATTO, cfl_scale=0.9, dfl_scale=0.9)
    over = Equation(N, B.NODE_GAUSS_LOBATTO, cfl_scale=0.9, dfl_scale=0.9, overintegration="cutoff", n_under=NU)
    assert over.overintegration == 1 and plain.overintegration == 0
    assert abs(plain.CFLScale - 0.9 * 1.8225 / 15.0) < 1e-15      # GL table of carpenterrk4-5, N = 7 (timedisc_vars.f90:175-190)
    assert abs(over.CFLScale - 0.9 * 1.2375 / 11.0) < 1e-15       # Gauss table, NEff = 5
    assert over.DFLScale == plain.DFLScale
