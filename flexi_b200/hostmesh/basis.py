"""1-D interpolation basis of the synthetic FLEXI host (stand-in for MOD_Basis / MOD_Interpolation / InitDGBasis).

The GPU library never builds a basis itself: exactly like the Fortran host it receives the finished operators
(D, D_Hat_T, L_Minus, L_HatMinus, DVolSurf, ...) through the init descriptor.  This module fabricates them the
way the reference does so that the descriptor is bit-compatible with what `InitInterpolation`/`InitDGBasis`
would hand over.

Reference (all under /root/reference/src):
  interpolation/basis.f90:142-183   LegendrePolynomialAndDerivative (normalised by sqrt(n+1/2))
  interpolation/basis.f90:370-438   LegendreGaussNodesAndWeights
  interpolation/basis.f90:521-586   LegGaussLobNodesAndWeights (+ qAndLEvaluation :487-512)
  interpolation/basis.f90:217-240   ChebyGaussLobNodesAndWeights
  interpolation/basis.f90:592-613   BarycentricWeights
  interpolation/basis.f90:619-643   PolynomialDerivativeMatrix
  interpolation/basis.f90:651-716   ALMOSTEQUAL, LagrangeInterpolationPolys
  interpolation/basis.f90:60-110    buildLegendreVdm
  interpolation/interpolation.f90:200-288,293-366  GetNodesAndWeights, GetVandermonde
  interpolation/changeBasis.t90     ChangeBasis3D (xi, then eta, then zeta)
  dg/dg.f90:132-209                 InitDGBasis
"""
import math

import numpy as np

EPS = np.finfo(np.float64).eps  # PP_RealTolerance, globals/preprocessing.f90:27

NODE_GAUSS = "gauss"
NODE_GAUSS_LOBATTO = "gauss-lobatto"
NODE_CL = "chebyshev-gauss-lobatto"
NODE_VISU = "visu"


def legendre_polynomial_and_derivative(n, x):
    """basis.f90:142-183. Returns (L_n(x), L_n'(x)) scaled by sqrt(n+1/2)."""
    if n == 0:
        L, Lder = 1.0, 0.0
    elif n == 1:
        L, Lder = x, 1.0
    else:
        L_Nm2, L_Nm1 = 1.0, x
        Lder_Nm2, Lder_Nm1 = 0.0, 1.0
        for i in range(2, n + 1):
            L = (float(2 * i - 1) * x * L_Nm1 - float(i - 1) * L_Nm2) / float(i)
            Lder = i * L_Nm1 + x * Lder_Nm1
            L_Nm2, L_Nm1 = L_Nm1, L
            Lder_Nm2, Lder_Nm1 = Lder_Nm1, Lder
    s = math.sqrt(float(n) + 0.5)
    return L * s, Lder * s


def legendre_gauss_nodes_and_weights(n):
    """basis.f90:370-438."""
    x = np.zeros(n + 1)
    w = np.zeros(n + 1)
    if n == 0:
        x[0] = 0.0
        w[0] = 2.0
        return x, w
    if n == 1:
        x[0] = -math.sqrt(1.0 / 3.0)
        x[1] = -x[0]
        w[:] = 1.0
        return x, w
    cheb_tmp = 2.0 * math.atan(1.0) / float(n + 1)
    for i in range((n + 1) // 2):
        xi = -math.cos(cheb_tmp * float(2 * i + 1))
        for _ in range(11):
            L, Ld = legendre_polynomial_and_derivative(n + 1, xi)
            dx = -L / Ld
            xi += dx
            if abs(dx) < 1.0e-15 * abs(xi):
                break
        L, Ld = legendre_polynomial_and_derivative(n + 1, xi)
        x[i] = xi
        x[n - i] = -xi
        w[i] = (2.0 * n + 3) / ((1.0 - xi * xi) * Ld * Ld)
        w[n - i] = w[i]
    if n % 2 == 0:
        x[n // 2] = 0.0
        L, Ld = legendre_polynomial_and_derivative(n + 1, 0.0)
        w[n // 2] = (2.0 * n + 3) / (Ld * Ld)
    return x, w


def _q_and_l_evaluation(n, x):
    """basis.f90:487-512 (Kopriva Alg. 24)."""
    L_Nm2, L_Nm1 = 1.0, x
    L = x
    for i in range(2, n + 1):
        L = (float(2 * i - 1) * x * L_Nm1 - float(i - 1) * L_Nm2) / float(i)
        L_Nm2, L_Nm1 = L_Nm1, L
    q = float(2 * n + 1) / float(n + 1) * (x * L - L_Nm2)
    qder = float(2 * n + 1) * L
    return q, qder, L


def leg_gauss_lob_nodes_and_weights(n):
    """basis.f90:521-586 (Kopriva Alg. 25)."""
    x = np.zeros(n + 1)
    w = np.zeros(n + 1)
    x[0], x[n] = -1.0, 1.0
    w[0] = 2.0 / float(n * (n + 1))
    w[n] = w[0]
    if n > 1:
        cont1 = math.pi / float(n)
        cont2 = 3.0 / (float(8 * n) * math.pi)
        for i in range(1, (n + 1) // 2):
            xi = -math.cos(cont1 * (float(i) + 0.25) - cont2 / (float(i) + 0.25))
            for _ in range(11):
                q, qder, L = _q_and_l_evaluation(n, xi)
                dx = -q / qder
                xi += dx
                if abs(dx) < 1.0e-15 * abs(xi):
                    break
            q, qder, L = _q_and_l_evaluation(n, xi)
            x[i] = xi
            x[n - i] = -xi
            w[i] = w[0] / (L * L)
            w[n - i] = w[i]
    if n % 2 == 0:
        x[n // 2] = 0.0
        q, qder, L = _q_and_l_evaluation(n, 0.0)
        w[n // 2] = w[0] / (L * L)
    return x, w


def cheby_gauss_lob_nodes_and_weights(n):
    """basis.f90:217-240."""
    x = np.array([-math.cos(i / float(n) * math.pi) for i in range(n + 1)])
    w = np.full(n + 1, math.pi / float(n))
    w[0] *= 0.5
    w[n] *= 0.5
    return x, w


def barycentric_weights(x):
    """basis.f90:592-613 (Kopriva Alg. 30)."""
    n = len(x) - 1
    wb = np.ones(n + 1)
    for i in range(1, n + 1):
        for j in range(i):
            wb[j] = wb[j] * (x[j] - x[i])
            wb[i] = wb[i] * (x[i] - x[j])
    return 1.0 / wb


def get_nodes_and_weights(n, node_type):
    """interpolation.f90:127-288. Returns (x, w, wBary)."""
    t = node_type.lower()
    if t == NODE_GAUSS:
        x, w = legendre_gauss_nodes_and_weights(n)
    elif t == NODE_GAUSS_LOBATTO:
        x, w = leg_gauss_lob_nodes_and_weights(n)
    elif t == NODE_CL:
        x, w = cheby_gauss_lob_nodes_and_weights(n)
    elif t == NODE_VISU:
        x = np.array([2.0 * float(i) / float(n) - 1.0 for i in range(n + 1)])
        w = np.full(n + 1, 2.0 / float(n))
        w[0] *= 0.5
        w[n] *= 0.5
    else:
        raise ValueError('NodeType "%s" in GetNodesAndWeights not found!' % node_type)
    return x, w, barycentric_weights(x)


def polynomial_derivative_matrix(x):
    """basis.f90:619-643 (Kopriva Alg. 37). D[i, j] = l_j'(x_i)."""
    n = len(x) - 1
    wb = barycentric_weights(x)
    D = np.zeros((n + 1, n + 1))
    for il in range(n + 1):
        for ig in range(n + 1):
            if il != ig:
                D[ig, il] = wb[il] / (wb[ig] * (x[ig] - x[il]))
                D[ig, ig] = D[ig, ig] - D[ig, il]
    return D


def almost_equal(x, y):
    """basis.f90:651-668 (Kopriva Alg. 139)."""
    if x == 0.0 or y == 0.0:
        return abs(x - y) <= 2.0 * EPS
    return abs(x - y) <= EPS * abs(x) and abs(x - y) <= EPS * abs(y)


def lagrange_interpolation_polys(x, xgp, wbary):
    """basis.f90:676-716 (Kopriva Alg. 34)."""
    n = len(xgp) - 1
    L = np.zeros(n + 1)
    hit = False
    for i in range(n + 1):
        if almost_equal(x, xgp[i]):
            L[i] = 1.0
            hit = True
    if hit:
        return L
    s = 0.0
    for i in range(n + 1):
        L[i] = wbary[i] / (x - xgp[i])
        s += L[i]
    return L / s


def initialize_vandermonde(x_in, wbary_in, x_out):
    """basis.f90:112-134. Vdm[iOut, iIn]."""
    return np.array([lagrange_interpolation_polys(xo, x_in, wbary_in) for xo in x_out])


def build_legendre_vdm(x_in):
    """basis.f90:60-110 (LAPACK branch). Returns (Vdm_Leg, sVdm_Leg)."""
    n = len(x_in) - 1
    V = np.zeros((n + 1, n + 1))
    for i in range(n + 1):
        for j in range(n + 1):
            V[i, j] = legendre_polynomial_and_derivative(j, x_in[i])[0]
    return V, np.linalg.inv(V)


def overintegration_mat(x_gp, n_under):
    """dg/overintegration.f90:104-116 (OVERINTEGRATIONTYPE_CUTOFF): Vdm_Leg * diag(1 for modes 0..NUnder, else 0) * sVdm_Leg,
    the modal cut-off projector as a nodal matrix OverintegrationMat[out, in]."""
    V, sV = build_legendre_vdm(x_gp)
    d = np.zeros(len(x_gp))
    d[: n_under + 1] = 1.0
    return (V * d[None, :]) @ sV


def get_vandermonde(n_in, type_in, n_out, type_out, modal=False):
    """interpolation.f90:293-366. Returns Vdm_In_Out[0:n_out, 0:n_in]."""
    if type_in.lower() == type_out.lower() and n_in == n_out:
        return np.eye(n_in + 1)
    x_in, _, wb_in = get_nodes_and_weights(n_in, type_in)
    x_out, _, _ = get_nodes_and_weights(n_out, type_out)
    if modal and n_out < n_in:
        _, sV_in = build_legendre_vdm(x_in)
        V_out, _ = build_legendre_vdm(x_out)
        return V_out[: n_out + 1, : n_out + 1] @ sV_in[: n_out + 1, : n_in + 1]
    return initialize_vandermonde(x_in, wb_in, x_out)


def change_basis_3d(vdm, x_in):
    """changeBasis.t90 ChangeBasis3D for arrays shaped [..., k, j, i, var] (C order == Fortran (var,i,j,k,...)).

    Applies the same 1-D matrix Vdm[out, in] along xi, eta, zeta.
    """
    y = np.einsum("ai,...kjiv->...kjav", vdm, x_in)
    y = np.einsum("bj,...kjav->...kbav", vdm, y)
    y = np.einsum("ck,...kbav->...cbav", vdm, y)
    return y


def change_basis_2d(vdm, x_in):
    """ChangeBasis2D for arrays shaped [..., q, p, var]."""
    y = np.einsum("ai,...qiv->...qav", vdm, x_in)
    y = np.einsum("bj,...jav->...bav", vdm, y)
    return y


class DGBasis:
    """Everything InitInterpolation + InitDGBasis leave in MOD_Interpolation_Vars / MOD_DG_Vars.

    All matrices are stored exactly as Fortran would index them: M[a, b] here == M(a, b) there.
    """

    def __init__(self, n, node_type, split_dg=False):
        self.N = n
        self.node_type = node_type.lower()
        self.split_dg = bool(split_dg)
        self.xGP, self.wGP, self.wBary = get_nodes_and_weights(n, node_type)
        # interpolation.f90:155-158  L_Minus/L_Plus = Lagrange polynomials at -1/+1
        self.L_Minus = lagrange_interpolation_polys(-1.0, self.xGP, self.wBary)
        self.L_Plus = lagrange_interpolation_polys(1.0, self.xGP, self.wBary)
        # dg.f90:166-209
        self.D = polynomial_derivative_matrix(self.xGP)
        self.D_T = self.D.T.copy()
        M = np.diag(self.wGP)
        Minv = np.diag(1.0 / self.wGP)
        self.D_Hat = -(Minv @ (self.D.T @ M))
        self.D_Hat_T = self.D_Hat.T.copy()
        self.L_HatPlus = Minv @ self.L_Plus
        self.L_HatMinus = Minv @ self.L_Minus
        if self.split_dg:
            if self.node_type == NODE_GAUSS:
                Vf = np.stack([self.L_Minus, self.L_Plus])
                B = np.array([[-1.0, 0.0], [0.0, 1.0]])
                dvs = self.D - 0.5 * (Minv @ ((Vf.T @ B) @ Vf))
                self.DVolSurf = dvs.T.copy()
            else:
                dvs = self.D_T.copy()
                dvs[0, 0] = dvs[0, 0] + 1.0 / (2.0 * self.wGP[0])
                dvs[n, n] = dvs[n, n] - 1.0 / (2.0 * self.wGP[n])
                self.DVolSurf = dvs
        else:
            self.DVolSurf = np.zeros((n + 1, n + 1))


def overintegration_cons(n, node_type, n_under, sJ):
    """InitOverintegration for OVERINTEGRATIONTYPE_CONSCUTOFF (dg/overintegration.f90:118-135): the modal projection
    Vdm_N_NUnder [NUnder+1, N+1], the interpolation Vdm_NUnder_N [N+1, NUnder+1] (GetVandermonde(..., modal=.TRUE.)) and
    sJNUnder [nElems, kU, jU, iU] = 1 / (Jacobian projected to NUnder).  The reference projects DetJac_Ref from NGeoRef; this host
    stand-in projects the Jacobian 1/sJ from N -- the library and the oracle take sJNUnder as an input either way."""
    vdm_down = get_vandermonde(n, node_type, n_under, node_type, modal=True)
    vdm_up = get_vandermonde(n_under, node_type, n, node_type, modal=False)
    jac = 1.0 / np.asarray(sJ)
    jac_u = np.einsum("li,mj,nk,ekji->enml", vdm_down, vdm_down, vdm_down, jac)
    return vdm_down, vdm_up, np.ascontiguousarray(1.0 / jac_u)


def laf_params(n, n_filter, w_gp, sJ):
    """InitFilter for FILTERTYPE_LAF (filter/filter.f90:149-186): IntegrationWeight [nElems,k,j,i] = wGP(i)wGP(j)wGP(k)/sJ normalised to
    sum 1 per element, and normMod."""
    iw = np.einsum("k,j,i->kji", w_gp, w_gp, w_gp)[None] / np.asarray(sJ)
    iw = iw / iw.reshape(len(iw), -1).sum(axis=1)[:, None, None, None]
    norm_mod = ((n_filter + 1.0) ** (-2.0 / 3.0) - 2.0 ** (-2.0 / 3.0)) / ((n + 1.0) ** (-2.0 / 3.0) - (n_filter + 1.0) ** (-2.0 / 3.0))
    return np.ascontiguousarray(iw), norm_mod
