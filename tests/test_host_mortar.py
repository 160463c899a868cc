"""Mortar interface operators (mortar/mortar.f90, mortar/fillmortar.t90) restated on the host: first step of the mortar plan of
DESIGN.md section 7.  The reference holds no fixture for them; InitMortar carries its own known-answer check, restated here, the
rest are the defining properties: interpolation is exact for polynomials, projection is the L2 adjoint, a projected interpolant
is the original, a constant flux through four small sides is the same flux through the big side (conservation)."""
import numpy as np
import pytest

from flexi_b200.hostmesh import basis as B
from flexi_b200.hostmesh.mortar import flux_mortar, init_mortar_self_check, mortar_basis, u_mortar

NODES = [B.NODE_GAUSS, B.NODE_GAUSS_LOBATTO]


@pytest.mark.parametrize("n", [1, 2, 3, 5, 7, 9, 11])
def test_init_mortar_self_check(n):
    """mortar.f90:77-88: error <= 100 * PP_RealTolerance (= 100 * epsilon)."""
    assert init_mortar_self_check(n) <= 100 * np.finfo(float).eps


@pytest.mark.parametrize("node", NODES)
@pytest.mark.parametrize("n", [2, 4, 7])
def test_big_to_small_is_exact_interpolation(n, node):
    xi, _, _ = B.get_nodes_and_weights(n, node)
    m01, m02, _, _ = mortar_basis(n, node)
    c = np.random.default_rng(n).normal(size=n + 1)
    f = lambda x: np.polyval(c, x)  # noqa: E731  degree n
    assert np.allclose(m01.T @ f(xi), f(0.5 * (xi - 1.0)), rtol=0, atol=1e-12)
    assert np.allclose(m02.T @ f(xi), f(0.5 * (xi + 1.0)), rtol=0, atol=1e-12)


@pytest.mark.parametrize("n", [2, 4, 7])
def test_small_to_big_is_the_l2_projection(n):
    """Gauss nodes (exact mass matrix): sum_p w_p v_p (M_1_0^T a + M_2_0^T b)_p = sum_p w_p ((M_0_1^T v)_p a_p + (M_0_2^T v)_p b_p)
    for all v, a, b -- the projection is the adjoint of the interpolation WITHOUT the length ratio 1/2 of the intervals, i.e. twice
    the L2 projection per split direction: the fluxes Flux_Mortar projects carry SurfElem of the small sides, which is half
    (a quarter) of the big side's (the same convention as InitMortar's own check: 0.25 * sum(...) = 1 for the constants 0.5 / 1.5).
    Hence projecting the interpolant of a polynomial gives twice the polynomial."""
    node = B.NODE_GAUSS
    _, w, _ = B.get_nodes_and_weights(n, node)
    m01, m02, m10, m20 = mortar_basis(n, node)
    rng = np.random.default_rng(3 * n)
    v, a, b = rng.normal(size=(3, n + 1))
    lhs = np.sum(w * v * (m10.T @ a + m20.T @ b))
    rhs = np.sum(w * ((m01.T @ v) * a + (m02.T @ v) * b))
    assert abs(lhs - rhs) <= 1e-13 * max(1.0, abs(rhs))
    assert np.allclose(m10.T @ (m01.T @ v) + m20.T @ (m02.T @ v), 2.0 * v, rtol=0, atol=1e-13)


@pytest.mark.parametrize("node", NODES)
@pytest.mark.parametrize("mtype", [1, 2, 3])
def test_face_operators(mtype, node):
    """U_Mortar on a bivariate polynomial is its evaluation on the small faces; Flux_Mortar after U_Mortar is 2x (types 2, 3) / 4x
    (type 1) the identity (SurfElem convention, see above); the big-side quadrature of the projected flux is the sum of the
    small-side quadratures (discrete conservation of SurfElem-weighted fluxes, Gauss weights)."""
    n = 4
    xi, w, _ = B.get_nodes_and_weights(n, node)
    m01, m02, m10, m20 = mortar_basis(n, node)
    rng = np.random.default_rng(11 + mtype)
    c = rng.normal(size=(n + 1, n + 1))
    poly = lambda x, y: np.polynomial.polynomial.polyval2d(x, y, c)  # noqa: E731
    P, Q = np.meshgrid(xi, xi, indexing="xy")                          # arrays [q, p]: xi along p, eta along q
    u_big = np.stack([poly(P, Q), 2.0 * poly(P, Q) + 1.0], axis=-1)
    small = u_mortar(u_big, mtype, m01, m02)
    lo, hi = (lambda x: 0.5 * (x - 1.0)), (lambda x: 0.5 * (x + 1.0))
    maps = {1: [(lo, lo), (hi, lo), (lo, hi), (hi, hi)], 2: [(None, lo), (None, hi)], 3: [(lo, None), (hi, None)]}[mtype]
    assert small.shape[0] == len(maps)
    for s, (fx, fy) in zip(small, maps):
        X = fx(P) if fx else P
        Y = fy(Q) if fy else Q
        assert np.allclose(s[..., 0], poly(X, Y), rtol=0, atol=1e-11)
    back = flux_mortar(small, mtype, m10, m20)
    tol = 1e-12 if node == B.NODE_GAUSS else 1e-11
    assert np.allclose(back, (4.0 if mtype == 1 else 2.0) * u_big, rtol=0, atol=tol)
    if node == B.NODE_GAUSS:
        f_small = rng.normal(size=small.shape)
        big = flux_mortar(f_small, mtype, m10, m20)
        w2 = w[:, None] * w[None, :]
        assert np.allclose(np.einsum("qp,qpv->v", w2, big), np.einsum("qp,sqpv->v", w2, f_small), rtol=0, atol=1e-12)


def test_mortar_numbering_against_a_reference_mesh():
    """tutorials/convtest/CART_HEX_PERIODIC_MORTAR_002_mesh.h5 (28 elements, mortar types 1, 2 and 3): the face nodes of every
    small element coincide with the big side's face nodes interpolated by U_Mortar -- i.e. the small-side numbering of
    fillmortar.t90 (type 1: (xi-,eta-), (xi+,eta-), (xi-,eta+), (xi+,eta+) in the big side's master orientation), the file layout
    read by mesh_readin.f90:333-376 and the flip of the small (slave) sides, on the reference's own mesh file."""
    import os

    from flexi_b200.hostmesh import mappings as MP
    from flexi_b200.hostmesh.hopr import mortar_topology
    from test_host_mesh import face_of

    path = os.path.join(os.path.dirname(__file__), "golden", "reference_naca3d", "CART_HEX_PERIODIC_MORTAR_002_mesh.h5")
    X, ngeo, mortars = mortar_topology(path)
    assert ngeo == 1 and sorted({m["type"] for m in mortars}) == [1, 2, 3]
    n = ngeo
    _, _, S2V2 = MP.build_mappings(n)
    m01, m02, _, _ = mortar_basis(n, B.NODE_VISU)
    L = X.reshape(-1, 3).max(axis=0) - X.reshape(-1, 3).min(axis=0)
    pq = np.arange(n + 1)
    P, Q = np.meshgrid(pq, pq, indexing="ij")

    def side_points(e, loc, flip):   # [q, p, 3] in the master orientation of the side
        a, c = S2V2[:, P, Q, flip, loc - 1]
        return np.transpose(face_of(X[e], loc, n)[c, a], (1, 0, 2))

    for m in mortars:
        big = side_points(m["elem"], m["loc_side"], 0)
        small = u_mortar(big, m["type"], m01, m02)
        assert len(m["small"]) == small.shape[0]
        for s, (se, sloc, flip) in zip(small, m["small"]):
            assert flip != 0, "small elements are the slaves of their mortar sides"
            d = side_points(se, sloc, flip) - s
            d -= L * np.round(d / L)   # periodic box
            assert np.max(np.abs(d)) < 1e-13, (m, se)


def test_mortar_surface_metrics_against_the_small_elements():
    """Mortar_CalcSurfMetrics on the reference's mortar mesh: the mortarized metrics of a big side (contravariant vector of the
    side's normal direction, i.e. NormVec * SurfElem up to the sign) equal what the small element computes for its own face from
    its own geometry -- with the opposite sign, it is the slave -- and so do the face points.  Gauss-Lobatto nodes, N = 2."""
    import os

    from flexi_b200.hostmesh import mappings as MP
    from flexi_b200.hostmesh.hopr import mortar_topology
    from flexi_b200.hostmesh.mesh import _element_metrics, _metric_operators
    from flexi_b200.hostmesh.mortar import mortar_surf_metrics
    from test_host_mesh import face_of

    path = os.path.join(os.path.dirname(__file__), "golden", "reference_naca3d", "CART_HEX_PERIODIC_MORTAR_002_mesh.h5")
    X, ngeo, mortars = mortar_topology(path)
    N = 2
    geo = _element_metrics(X, _metric_operators(N, B.NODE_GAUSS_LOBATTO, ngeo))
    _, _, S2V2 = MP.build_mappings(N)
    m01, m02, _, _ = mortar_basis(N, B.NODE_GAUSS_LOBATTO)
    L = X.reshape(-1, 3).max(axis=0) - X.reshape(-1, 3).min(axis=0)
    pq = np.arange(N + 1)
    P, Q = np.meshgrid(pq, pq, indexing="ij")
    met_of = {MP.XI_MINUS: "Mf", MP.XI_PLUS: "Mf", MP.ETA_MINUS: "Mg", MP.ETA_PLUS: "Mg", MP.ZETA_MINUS: "Mh", MP.ZETA_PLUS: "Mh"}
    sign_of = {MP.XI_MINUS: -1.0, MP.ETA_MINUS: -1.0, MP.ZETA_MINUS: -1.0, MP.XI_PLUS: 1.0, MP.ETA_PLUS: 1.0, MP.ZETA_PLUS: 1.0}

    def side(A, e, loc, flip):   # [q, p, c] in the master orientation of the side
        a, c = S2V2[:, P, Q, flip, loc - 1]
        return np.transpose(face_of(A[e], loc, N)[c, a], (1, 0, 2))

    n_checked = 0
    for m in mortars:
        e, loc = m["elem"], m["loc_side"]
        nS_big = sign_of[loc] * side(geo[met_of[loc]], e, loc, 0)          # outward normal * SurfElem of the big side
        ja, xgp = mortar_surf_metrics(nS_big, side(geo["xGP"], e, loc, 0), m["type"], m01, m02)
        for k, (se, sloc, flip) in enumerate(m["small"]):
            nS_small = sign_of[sloc] * side(geo[met_of[sloc]], se, sloc, flip)
            assert np.max(np.abs(nS_small + ja[k])) < 1e-13, (m, k)        # the slave's outward normal is the opposite one
            d = side(geo["xGP"], se, sloc, flip) - xgp[k]
            d -= L * np.round(d / L)
            assert np.max(np.abs(d)) < 1e-13
            n_checked += 1
    assert n_checked == 24


@pytest.mark.parametrize("node", NODES)
@pytest.mark.parametrize("mtype", [1, 2, 3])
def test_free_stream_across_a_mortar(mtype, node):
    """For a uniform state the numerical flux of a small side is F(U) . (n SurfElem)_small; projected to the big side by Flux_Mortar
    it must be F(U) . (n SurfElem)_big, or the big element would not preserve the free stream: Flux_Mortar(Mortar_CalcSurfMetrics(Ja))
    = Ja for any polynomial Ja of degree N (curved big side)."""
    n = 5
    xi, _, _ = B.get_nodes_and_weights(n, node)
    m01, m02, m10, m20 = mortar_basis(n, node)
    from flexi_b200.hostmesh.mortar import mortar_surf_metrics

    rng = np.random.default_rng(5 + mtype)
    c = rng.normal(size=(3, n + 1, n + 1))
    P, Q = np.meshgrid(xi, xi, indexing="xy")
    nS_big = np.stack([np.polynomial.polynomial.polyval2d(P, Q, c[d]) for d in range(3)], axis=-1)      # [q, p, 3]
    nS_small, _ = mortar_surf_metrics(nS_big, nS_big, mtype, m01, m02)
    F = rng.normal(size=(5, 3))                                                                         # F(U): one row per variable
    flux_small = np.einsum("vd,sqpd->sqpv", F, nS_small)
    flux_big = flux_mortar(flux_small, mtype, m10, m20)
    assert np.allclose(flux_big, np.einsum("vd,qpd->qpv", F, nS_big), rtol=0, atol=1e-11)
