"""Synthetic host mesh: closed forms, side bookkeeping (A.3), watertight flips, partition invariance."""
import numpy as np
import pytest

from flexi_b200.hostmesh import basis as B
from flexi_b200.hostmesh import mappings as MP
from flexi_b200.hostmesh.mesh import BoxSpec, build_mesh


def face_of(A, loc_side, N):
    """Volume [k,j,i,c] -> face [b,a,c] as EvalElemFaceGL (prolongtoface.t90:270-303)."""
    if loc_side == MP.XI_MINUS:
        return A[:, :, 0]
    if loc_side == MP.XI_PLUS:
        return A[:, :, N]
    if loc_side == MP.ETA_MINUS:
        return A[:, 0, :]
    if loc_side == MP.ETA_PLUS:
        return A[:, N, :]
    if loc_side == MP.ZETA_MINUS:
        return A[0]
    return A[N]


def test_cartesian_closed_forms():
    N = 4
    spec = BoxSpec((2, 3, 4), length=(2.0, 3.0, 8.0))
    m = build_mesh(spec, N, B.NODE_GAUSS)
    hx, hy, hz = 1.0, 1.0, 2.0
    assert np.allclose(m.Metrics_fTilde, np.array([hy * hz / 4, 0, 0]), atol=1e-13)
    assert np.allclose(m.Metrics_gTilde, np.array([0, hx * hz / 4, 0]), atol=1e-13)
    assert np.allclose(m.Metrics_hTilde, np.array([0, 0, hx * hy / 4]), atol=1e-13)
    assert np.allclose(m.sJ, 8.0 / (hx * hy * hz), rtol=1e-13)
    assert m.nSides == 3 * 24 and m.nBCSides == 0 and m.nInnerSides == 72
    # volume = sum w w w / sJ (analyze.f90:150-158)
    b = B.DGBasis(N, B.NODE_GAUSS)
    w3 = np.einsum("k,j,i->kji", b.wGP, b.wGP, b.wGP)
    assert abs(np.sum(w3[None] / m.sJ) - 48.0) < 1e-11


@pytest.mark.parametrize("node", [B.NODE_GAUSS, B.NODE_GAUSS_LOBATTO])
@pytest.mark.parametrize("deform,ngeo", [(0.0, 1), (0.1, 3)])
def test_watertight_sides_and_normals(node, deform, ngeo):
    N = 4
    spec = BoxSpec((3, 2, 2), length=(2.0, 2.0, 2.0), x0=(-1.0, -1.0, -1.0), ngeo=ngeo, deform=deform)
    m = build_mesh(spec, N, node)
    _, _, S2V2 = MP.build_mappings(N)
    b = B.DGBasis(N, node)
    if node == B.NODE_GAUSS:
        # prolong coordinates to the faces like EvalElemFaceG
        def trace(X, s):
            Lm, Lp = b.L_Minus, b.L_Plus
            L = Lm if s in (MP.XI_MINUS, MP.ETA_MINUS, MP.ZETA_MINUS) else Lp
            if s in (MP.XI_MINUS, MP.XI_PLUS):
                return np.einsum("kjic,i->kjc", X, L)
            if s in (MP.ETA_MINUS, MP.ETA_PLUS):
                return np.einsum("kjic,j->kic", X, L)
            return np.einsum("kjic,k->jic", X, L)
    else:
        def trace(X, s):
            return face_of(X, s, N)
    L = np.array([2.0, 2.0, 2.0])
    for sid in range(m.nSides):
        me, se, ms, ss, fl = m.SideToElem[sid]
        assert me > 0 and se > 0
        xm = trace(m.Elem_xGP[me - 1], ms)
        xs = trace(m.Elem_xGP[se - 1], ss)
        for p in range(N + 1):
            for q in range(N + 1):
                a, c = S2V2[:, p, q, 0, ms - 1]
                a2, c2 = S2V2[:, p, q, fl, ss - 1]
                d = xm[c, a] - xs[c2, a2]
                d = d - L * np.round(d / L)          # periodic image
                assert np.max(np.abs(d)) < 1e-12
                if ngeo == 1 or N >= ngeo:
                    dm = xm[c, a] - m.Face_xGP[sid, q, p]
                    assert np.max(np.abs(dm)) < 1e-12
        # normal points out of the master element: (face centre - element centre) . n > 0
        ctr = m.Elem_xGP[me - 1].reshape(-1, 3).mean(axis=0)
        fc = m.Face_xGP[sid].reshape(-1, 3).mean(axis=0)
        nv = m.NormVec[sid].reshape(-1, 3).mean(axis=0)
        assert np.dot(fc - ctr, nv) > 0
        # orthonormal frame, right handed (metrics.f90:751-787)
        n_, t1, t2 = m.NormVec[sid], m.TangVec1[sid], m.TangVec2[sid]
        assert np.allclose(np.sum(n_ * n_, -1), 1) and np.allclose(np.sum(t1 * t1, -1), 1)
        assert np.allclose(np.sum(n_ * t1, -1), 0, atol=1e-13)
        assert np.allclose(np.cross(n_, t1), t2)


def test_discrete_metric_identities_curved():
    """Curl-form metrics are free-stream preserving: sum_d D_d Ja^d_n = 0 on the solution nodes (metrics.f90:441-475)."""
    N = 5
    spec = BoxSpec((2, 2, 2), length=(2.0, 2.0, 2.0), x0=(-1.0, -1.0, -1.0), ngeo=2, deform=0.1)
    for node in (B.NODE_GAUSS, B.NODE_GAUSS_LOBATTO):
        m = build_mesh(spec, N, node)
        b = B.DGBasis(N, node)
        div = (np.einsum("il,ekjlc->ekjic", b.D, m.Metrics_fTilde) + np.einsum("jl,eklic->ekjic", b.D, m.Metrics_gTilde)
               + np.einsum("kl,eljic->ekjic", b.D, m.Metrics_hTilde))
        assert np.max(np.abs(div)) < 1e-12


def test_side_ranges_with_bcs():
    spec = BoxSpec((2, 2, 2), periodic=(False, True, True), bc_of_face={(0, -1): 1, (0, 1): 2})
    m = build_mesh(spec, 2, B.NODE_GAUSS)
    assert m.nBCSides == 8 and m.nInnerSides == 4 + 8 + 8
    assert set(m.BC[:8].tolist()) == {1, 2}
    assert np.all(m.SideToElem[:8, 1] == -1) and np.all(m.SideToElem[:8, 0] > 0)
    # first-seen numbering: first BC side belongs to element 1, XI_MINUS
    assert m.SideToElem[0, 0] == 1 and m.SideToElem[0, 2] == MP.XI_MINUS


@pytest.mark.parametrize("partition,blocks,nranks", [("linear", None, 2), ("linear", None, 3), ("block", (2, 2, 1), 4)])
def test_partition_bookkeeping(partition, blocks, nranks):
    N = 2
    spec = BoxSpec((4, 4, 2), ngeo=2, deform=0.05)
    glob = build_mesh(spec, N, B.NODE_GAUSS_LOBATTO)
    meshes = [build_mesh(spec, N, B.NODE_GAUSS_LOBATTO, rank=r, nranks=nranks, partition=partition, blocks=blocks)
              for r in range(nranks)]
    assert sum(m.nElems for m in meshes) == glob.nElems
    for r, m in enumerate(meshes):
        assert m.nSides == m.nBCSides + m.nInnerSides + m.nMPISides_MINE + m.nMPISides_YOUR
        # MINE: master present, slave absent; YOUR: the other way round
        mine = slice(m.firstMPISide_MINE - 1, m.lastMPISide_MINE)
        your = slice(m.firstMPISide_YOUR - 1, m.lastMPISide_YOUR)
        assert np.all(m.SideToElem[mine, 0] > 0) and np.all(m.SideToElem[mine, 1] == -1)
        assert np.all(m.SideToElem[your, 0] == -1) and np.all(m.SideToElem[your, 1] > 0)
        for inb, nb in enumerate(m.NbProc):
            o = meshes[nb]
            jnb = list(o.NbProc).index(r)
            # my MINE block towards nb pairs with nb's YOUR block towards me, slice by slice
            a0, a1 = m.offsetMPISides_MINE[inb], m.offsetMPISides_MINE[inb + 1]
            b0, b1 = o.offsetMPISides_YOUR[jnb], o.offsetMPISides_YOUR[jnb + 1]
            assert a1 - a0 == b1 - b0
            # same geometry on both ends (this is what the Geo exchange guarantees in the reference)
            assert np.allclose(m.NormVec[a0:a1], o.NormVec[b0:b1], atol=1e-13)
            assert np.allclose(m.SurfElem[a0:a1], o.SurfElem[b0:b1], atol=1e-13)
            # lower rank owns floor(n/2) MINE sides (prepare_mesh.f90:247-254)
            n = (a1 - a0) + (m.offsetMPISides_YOUR[inb + 1] - m.offsetMPISides_YOUR[inb])
            assert (a1 - a0) == (n // 2 if r < nb else n - n // 2)


def test_host_body_forces_of_a_uniform_state_are_pressure_times_wall_area():
    """Checker of the GPU CalcBodyForces test: on a Cartesian box with walls at y = -1 / +1 (2 x 2 each) a uniform state at rest
    gives Fp = p * n * A per wall BC, no friction, and a moment about the wall centre of zero."""
    from flexi_b200.hostmesh.analyze import calc_body_forces
    from flexi_b200.hostmesh.case import build_case
    from flexi_b200.hostmesh.mesh import BoxSpec

    btype = [(4, 1, 0), (9, 0, 0), (2, 1, 0)]
    bc = {(1, -1): 1, (1, 1): 2, (0, -1): 3, (0, 1): 3}
    eqn = dict(parabolic=True, riemann="lf", mu0=1e-2, boundary_type=btype, ref_state_prim=[(1.0, 0.0, 0.0, 0.0, 0.9)])
    r = build_case(BoxSpec((3, 2, 2), periodic=(False, False, True), bc_of_face=bc, x0=(-1.0, -1.0, -1.0), length=(2.0, 2.0, 2.0)),
                   3, "gauss-lobatto", eqn)[0]
    m = r.mesh
    p0 = 0.9
    U = np.zeros((m.nSides, 4, 4, 5))
    U[..., 0] = 1.0
    U[..., 4] = p0 / (r.eqn.kappa - 1.0)
    g0 = np.zeros((m.nSides, 4, 4, 5))
    Fp, Fv, Mp, Mv = calc_body_forces(r.basis, r.eqn, m, btype, U, g0, g0, g0, mom_origin=(0.0, -1.0, 0.0))
    assert np.allclose(Fp[0], (0.0, -p0 * 4.0, 0.0), atol=1e-13) and np.allclose(Fp[1], (0.0, p0 * 4.0, 0.0), atol=1e-13)
    assert np.all(Fp[2] == 0.0) and np.all(Fv == 0.0) and np.all(Mv == 0.0)
    assert np.allclose(Mp[0], 0.0, atol=1e-13)  # symmetric wall about the moment origin


def test_entropy_variables_round_trip_and_symmetrise_the_flux_jacobian():
    """PrimToEntropy / EntropyToCons (eos.f90:401-531; groundwork for the Gauss split form): the maps are inverse to each other,
    and dU/dw is symmetric positive definite (what makes w the entropy variables of s = ln p - kappa ln rho)."""
    from flexi_b200.hostmesh.case import TGV_EQN
    from flexi_b200.hostmesh.equation import Equation

    eq = Equation(3, "gauss", parabolic=False, **TGV_EQN)
    rng = np.random.default_rng(7)
    prim = np.empty((50, 6))
    prim[:, 0] = rng.uniform(0.5, 1.5, 50)
    prim[:, 1:4] = rng.uniform(-0.5, 0.5, (50, 3))
    prim[:, 4] = rng.uniform(0.5, 1.5, 50)
    prim[:, 5] = prim[:, 4] / (prim[:, 0] * eq.R)
    cons = np.empty((50, 5))
    cons[:, 0] = prim[:, 0]
    cons[:, 1:4] = prim[:, 0:1] * prim[:, 1:4]
    cons[:, 4] = prim[:, 4] / (eq.kappa - 1.0) + 0.5 * prim[:, 0] * np.sum(prim[:, 1:4] ** 2, axis=1)
    w = eq.prim_to_entropy(prim)
    assert np.max(np.abs(eq.entropy_to_cons(w) - cons)) < 1e-13
    # finite-difference Jacobian dU/dw at one state
    w0, h = w[0], 1e-6
    J = np.stack([(eq.entropy_to_cons(w0 + h * e) - eq.entropy_to_cons(w0 - h * e)) / (2 * h) for e in np.eye(5)], axis=1)
    assert np.max(np.abs(J - J.T)) < 1e-7 * np.max(np.abs(J))
    assert np.all(np.linalg.eigvalsh(0.5 * (J + J.T)) > 0)


# ---- "unstructured" connectivity: every element in its own local frame (all flips 0..4, all local-side pairings) ---------------
def test_oriented_flip_reduces_to_the_aligned_table():
    from flexi_b200.hostmesh import mesh as M

    for s in range(1, 7):
        assert M.oriented_flip(0, s, 0, M._OPPOSITE[s]) == M._FLIP_TABLE[(s, M._OPPOSITE[s])]
    seen = set()
    for o_m in range(M.N_ORIENT):
        for o_s in range(M.N_ORIENT):
            for ls_m in range(1, 7):
                gax, gsg = M._ORI_GAXIS[o_m, ls_m - 1], M._ORI_GSIGN[o_m, ls_m - 1]
                ls_s = M._ORI_LOCSIDE[o_s, gax, (1 - gsg) // 2]
                seen.add((ls_m, ls_s, M.oriented_flip(o_m, ls_m, o_s, ls_s)))
    # every (master side, slave side) pair occurs, and every flip 1..4 (flip 0 is the master's own)
    assert {(a, b) for a, b, _ in seen} == {(a, b) for a in range(1, 7) for b in range(1, 7)}
    assert {f for _, _, f in seen} == {1, 2, 3, 4}


@pytest.mark.parametrize("node,eqn,N", [
    (B.NODE_GAUSS, dict(parabolic=True, riemann="hllc", mu0=1e-2, lifting=2), 4),
    (B.NODE_GAUSS_LOBATTO, dict(parabolic=True, riemann="lf", mu0=1e-2, split_dg=True, split_dg_type="pi"), 5),
    (B.NODE_GAUSS, dict(parabolic=False, riemann="roe"), 3),
])
def test_randomly_oriented_elements_give_the_same_operator(node, eqn, N):
    """The DG operator does not depend on the local frames of the elements: Ut and the lifted gradients of the oracle on a curved
    periodic mesh whose elements are rotated at random (all five flips and mixed local-side pairings occur) equal those on the
    block-aligned mesh at the same physical nodes.  Pins hostmesh's orientation logic AND the oracle's S2V/S2V2/flip handling
    (mappings.f90:188-450, prolongtoface.t90:110-140, surfint.t90:541-583)."""
    from flexi_b200.hostmesh.case import build_case
    from oracle import oracle as O

    res = {}
    for seed in (None, 11):
        spec = BoxSpec((3, 3, 2), orient_seed=seed, x0=(-1.0, -1.0, -1.0), length=(2.0, 2.0, 2.0), ngeo=2, deform=0.1)
        r = build_case(spec, N, node, eqn)[0]
        U = r.eqn.smooth_state(r.mesh.Elem_xGP, length=2.0)
        w = O.OracleWorld([(r.desc, r.keep)])
        w.set_state([U])
        w.dg_time_derivative(0.0)
        res[seed] = (spec, r, w.ranks[0].array("Ut").copy(), w.ranks[0].array("gradUx").copy() if eqn["parabolic"] else None)
        w.close()
    spec, r, ut_o, gx_o = res[11]
    assert np.all(np.bincount(r.mesh.ElemToSide[:, :, 1].ravel(), minlength=5) > 0), "all flips 0..4 must occur"
    ut_a = BoxSpec._to_local_frames(res[None][2], spec.orient)
    assert np.abs(ut_o - ut_a).max() <= 2e-12 * np.abs(ut_a).max()
    if gx_o is not None:
        gx_a = BoxSpec._to_local_frames(res[None][3], spec.orient)
        assert np.abs(gx_o - gx_a).max() <= 1e-12 * np.abs(gx_a).max()


def test_oriented_mesh_partition_invariance():
    """1 rank vs 3 ranks on the randomly oriented curved mesh (MPI sides with every flip): same Ut."""
    from flexi_b200.hostmesh.case import build_case
    from oracle import oracle as O

    eqn = dict(parabolic=True, riemann="hllc", mu0=1e-2, lifting=2)
    spec = BoxSpec((3, 3, 2), orient_seed=5, x0=(-1.0, -1.0, -1.0), length=(2.0, 2.0, 2.0), ngeo=2, deform=0.1)
    outs = []
    for nr in (1, 3):
        ranks = build_case(spec, 4, B.NODE_GAUSS, eqn, nranks=nr)
        w = O.OracleWorld([(r.desc, r.keep) for r in ranks])
        w.set_state([r.eqn.smooth_state(r.mesh.Elem_xGP, length=2.0) for r in ranks])
        w.dg_time_derivative(0.0)
        outs.append(np.concatenate([r.array("Ut") for r in w.ranks]))
        w.close()
    assert np.abs(outs[0] - outs[1]).max() <= 1e-12 * np.abs(outs[0]).max()
