"""Self-validation of the oracle's boundary conditions, BR2 lifting and manufactured-solution source (SURVEY.md 8c identities):
free-stream preservation with Dirichlet / slip-wall BCs (run_basic/freestream_3D: L2 <= 1e-12), exact gradients of polynomials,
BR1 == BR2 for continuous fields, wall-flux properties and the p-convergence of ExactFunc 4 + CalcSource (convtest/p_3D)."""
import numpy as np
import pytest

from flexi_b200.hostmesh import basis as B
from flexi_b200.hostmesh.case import build_case
from flexi_b200.hostmesh.mesh import BoxSpec
from oracle import oracle as O

G, GL = B.NODE_GAUSS, B.NODE_GAUSS_LOBATTO
REF = (1.0, 0.3, 0.0, 0.0, 0.71428571)   # convtest RefState
BOX = dict(x0=(-1.0, -1.0, -1.0), length=(2.0, 2.0, 2.0))


def world(spec, N, node, eqn, nranks=1):
    ranks = build_case(spec, N, node, eqn, nranks=nranks)
    return ranks, O.OracleWorld([(r.desc, r.keep) for r in ranks])


def dirichlet_box(nel, curved=False, periodic=(False, False, False)):
    bc = {(a, s): 1 for a in range(3) for s in (-1, 1) if not periodic[a]}
    kw = dict(ngeo=2, deform=0.1) if curved else {}
    return BoxSpec(nel, periodic=periodic, bc_of_face=bc, **BOX, **kw)


@pytest.mark.parametrize("node", [G, GL])
@pytest.mark.parametrize("lifting", [1, 2])
def test_freestream_with_dirichlet_bcs_on_curved_mesh(node, lifting):
    eqn = dict(parabolic=True, lifting=lifting, riemann="roe", mu0=1e-2, boundary_type=[(2, 1, 0)], ref_state_prim=[REF])
    ranks, w = world(dirichlet_box((2, 2, 2), curved=True), 4, node, eqn, nranks=2)
    w.set_state([r.eqn.const_state(r.mesh.Elem_xGP.shape[:-1], REF) for r in ranks])
    w.dg_time_derivative(0.3)
    for o in w.ranks:
        assert np.max(np.abs(o.array("Ut"))) < 1e-11
        assert np.max(np.abs(o.array("gradUx"))) < 1e-11
    w.close()


@pytest.mark.parametrize("bctype", [9, 3, 4])
def test_wall_bcs_no_mass_flux_and_tangential_freestream(bctype):
    """Walls at x = +-1: the mass flux through wall sides is exactly zero; for the slip wall a wall-parallel constant flow is a
    steady state (Ut = 0), for no-slip walls only the density equation stays steady at t=0."""
    eqn = dict(parabolic=True, riemann="roe", mu0=1e-2, boundary_type=[(bctype, 1, 0)], ref_state_prim=[(1.0, 0.0, 0.2, -0.1, 0.8)])
    spec = BoxSpec((2, 2, 2), periodic=(False, True, True), bc_of_face={(0, -1): 1, (0, 1): 1}, **BOX)
    ranks, w = world(spec, 3, GL, eqn)
    r = ranks[0]
    w.set_state([r.eqn.const_state(r.mesh.Elem_xGP.shape[:-1], (1.0, 0.0, 0.2, -0.1, 0.8))])
    w.dg_time_derivative(0.0)
    o = w.ranks[0]
    nbc = r.mesh.nBCSides
    assert np.max(np.abs(o.array("Flux")[:nbc, ..., 0])) == 0.0
    if bctype == 9:
        assert np.max(np.abs(o.array("Ut"))) < 1e-12
    else:
        assert np.max(np.abs(o.array("Ut")[..., 0])) < 1e-12
        assert np.max(np.abs(o.array("Ut")[..., 2])) > 1e-4   # the wall decelerates the tangential flow
    w.close()


@pytest.mark.parametrize("node", [G, GL])
def test_br2_equals_br1_for_continuous_polynomial_and_is_exact(node):
    """A globally polynomial (continuous) field has no face jumps: BR1 and BR2 give the same, exact gradient."""
    N = 4
    res = {}
    for lifting in (1, 2):
        eqn = dict(parabolic=True, lifting=lifting, riemann="lf", mu0=1e-2, R=1.0, boundary_type=[(2, 0, 0)], ref_state_prim=[REF],
                   ini_exact_func=1)
        ranks, w = world(dirichlet_box((3, 3, 3)), N, node, eqn)
        r = ranks[0]
        x = r.mesh.Elem_xGP
        prim = np.zeros(x.shape[:-1] + (6,))
        prim[..., 0] = 1.0 + 0.1 * x[..., 0] * x[..., 1]
        prim[..., 1] = 0.2 * x[..., 1] ** 2
        prim[..., 2] = 0.1 * x[..., 2] * x[..., 0]
        prim[..., 3] = -0.3 * x[..., 0] ** 2
        prim[..., 5] = 1.0 + 0.2 * x[..., 2]                      # T, R = 1
        prim[..., 4] = prim[..., 0] * prim[..., 5]
        w.set_state([r.eqn.prim_to_cons(prim)])
        w.dg_time_derivative(0.0)
        res[lifting] = [w.ranks[0].array(n).copy() for n in ("gradUx", "gradUy", "gradUz")]
        w.close()
    c = 13  # the central element of the 3x3x3 box does not see the (inexact) Dirichlet boundary state
    for a, b in zip(res[1], res[2]):
        assert np.max(np.abs(a[c] - b[c])) < 1e-11
    gx = res[2][0]
    assert np.max(np.abs(gx[c, ..., 3] - (-0.6 * x[c, ..., 0]))) < 1e-10


def test_br2_partition_invariance_and_conservation():
    spec = BoxSpec((3, 2, 2), ngeo=2, deform=0.1, **BOX)
    eqn = dict(parabolic=True, lifting=2, riemann="roe", mu0=1e-2)
    ranks1, w1 = world(spec, 3, G, eqn)
    ranks3, w3 = world(spec, 3, G, eqn, nranks=3)
    w1.set_state([ranks1[0].eqn.smooth_state(ranks1[0].mesh.Elem_xGP, length=2.0)])
    w3.set_state([r.eqn.smooth_state(r.mesh.Elem_xGP, length=2.0) for r in ranks3])
    w1.dg_time_derivative(0.0)
    w3.dg_time_derivative(0.0)
    full = w1.ranks[0].array("Ut")
    for r, o in zip(ranks3, w3.ranks):
        assert np.max(np.abs(o.array("Ut") - full[r.mesh.elem_gid])) < 1e-12 * np.max(np.abs(full))
    w3d = np.einsum("k,j,i->kji", ranks1[0].basis.wGP, ranks1[0].basis.wGP, ranks1[0].basis.wGP)
    cons = np.einsum("ekjiv,ekji->v", full, w3d[None] / ranks1[0].mesh.sJ)
    assert np.max(np.abs(cons)) < 1e-12
    w1.close()
    w3.close()


@pytest.mark.parametrize("node,lifting", [(G, 1), (GL, 2)])
def test_manufactured_solution_p_convergence(node, lifting):
    """convtest/p_3D: ExactFunc 4 + CalcSource, Dirichlet exact BCs: the residual Ut - dU_exact/dt converges spectrally."""
    errs = []
    for N in (3, 5, 7):
        eqn = dict(parabolic=True, lifting=lifting, riemann="roeentropyfix", mu0=1e-3, kappa=1.4, R=287.058, boundary_type=[(2, 0, 0)],
                   ref_state_prim=[REF], ini_exact_func=4, do_calc_source=True, adv_vel=(0.3, 0.3, 0.3))
        ranks, w = world(dirichlet_box((2, 2, 2)), N, node, eqn)
        r = ranks[0]
        t = 0.1
        w.set_state([r.eqn.exact_func4(r.mesh.Elem_xGP, t)])
        w.dg_time_derivative(t)
        h = 1e-6
        dUdt = (r.eqn.exact_func4(r.mesh.Elem_xGP, t + h) - r.eqn.exact_func4(r.mesh.Elem_xGP, t - h)) / (2 * h)
        errs.append(np.max(np.abs(w.ranks[0].array("Ut") - dUdt)))
        w.close()
    # spectral decay of the pointwise residual (it contains first and second derivatives of the degree-N interpolant)
    assert errs[1] < errs[0] / 8 and errs[2] < errs[1] / 8, errs
    assert errs[2] < 1e-2, errs
