"""BASELINE configs[3]: NACA0012-style synthetic curved O-mesh with randomly oriented elements (all flips / local-side pairings of
an unstructured HOPR mesh), N=5, HLLC + BR2, isothermal no-slip wall + Dirichlet far field.  CPU: geometry and free-stream
preservation through the oracle; GPU: the CUDA path against the oracle (one Ut, all side arrays, RK steps)."""
import numpy as np
import pytest

from flexi_b200.hostmesh import basis as B
from flexi_b200.hostmesh.case import NACA_EQN, build_case, naca_spec, naca_state
from oracle import oracle as O


def test_naca_mesh_geometry_and_connectivity():
    r = build_case(naca_spec((16, 4, 2)), 5, B.NODE_GAUSS, NACA_EQN)[0]
    m = r.mesh
    assert m.nElems == 128 and m.nBCSides == 2 * 16 * 2 and np.all(m.sJ > 0.0) and np.all(np.isfinite(m.sJ))
    # every flip 0..4 and both BC types occur; wall sides carry BC 1, far-field sides BC 2
    flips = np.unique(m.ElemToSide[..., 1])
    assert set(flips.tolist()) == {0, 1, 2, 3, 4}
    assert sorted(np.unique(m.BC[:m.nBCSides]).tolist()) == [1, 2]
    # the wall sides lie on the section (|y| <= 0.06 c + curvature), the far-field sides on the circle of radius 8
    xf = m.Face_xGP[:m.nBCSides]
    wall = m.BC[:m.nBCSides] == 1
    assert 0.0595 < np.max(np.abs(xf[wall][..., 1])) < 0.0601 and np.all((xf[wall][..., 0] > -1e-12) & (xf[wall][..., 0] < 1 + 1e-12))
    rad = np.hypot(xf[~wall][..., 0] - 0.5, xf[~wall][..., 1])
    assert np.max(np.abs(rad - 8.0)) < 2e-3  # NGeo = 4 interpolant of the circle


def test_naca_mesh_freestream_preservation():
    """Uniform flow with a far-field state on every boundary: the curved, randomly oriented mesh must be watertight (metric
    identities, master/slave normals, S2V2 of every flip) for Ut to vanish."""
    eqn = dict(NACA_EQN, boundary_type=[(2, 1, 0), (2, 1, 0)])
    r = build_case(naca_spec((12, 3, 2)), 5, B.NODE_GAUSS, eqn)[0]
    rs = r.eqn.ref_state_prim[0]
    prim = np.zeros(r.mesh.Elem_xGP.shape[:-1] + (6,))
    prim[...] = rs
    w = O.OracleWorld([(r.desc, r.keep)])
    w.set_state([r.eqn.prim_to_cons(prim)])
    w.dg_time_derivative(0.0)
    ut = w.ranks[0].array("Ut")
    w.close()
    # scale: a unit flux through an element face of the finest layer, |Ut| ~ sJ * SurfElem ~ 1e3
    assert np.max(np.abs(ut)) < 2e-9, np.max(np.abs(ut))


@pytest.mark.gpu
@pytest.mark.parametrize("node,extra", [(B.NODE_GAUSS, {}), (B.NODE_GAUSS_LOBATTO, {}),
                                        (B.NODE_GAUSS_LOBATTO, dict(split_dg=True, split_dg_type="pi", riemann="lf", lifting=1))])
def test_naca_gpu_matches_oracle(node, extra):
    from flexi_b200.gpu import FlexiGpu

    r = build_case(naca_spec((16, 4, 2)), 5, node, dict(NACA_EQN, **extra))[0]
    U0 = naca_state(r.eqn, r.mesh.Elem_xGP)
    w = O.OracleWorld([(r.desc, r.keep)])
    w.set_state([U0])
    w.dg_time_derivative(0.0)
    o = w.ranks[0]
    with FlexiGpu(r) as g:
        g.set_state(U0)
        ut = g.DGTimeDerivative_weakForm(0.0)
        ref = o.array("Ut")
        scale = np.max(np.abs(ref).reshape(-1, 5), axis=0)
        scale = np.maximum(scale, 1e-3 * scale.max())
        noise, _ = O.ut_noise_floor([(r.desc, r.keep)], [U0], 0.0)
        d = np.max(np.abs(ut - ref).reshape(-1, 5), axis=0)
        assert np.all(d <= np.maximum(1e-12 * scale, 4.0 * noise)), (d, scale, noise)
        assert np.max(np.abs(g.get_array("Flux") - o.array("Flux"))) <= 1e-12 * np.max(np.abs(o.array("Flux")))
        t = 0.0
        for _ in range(3):
            dt = w.calc_timestep()[0]
            assert abs(g.CalcTimeStep()[0] - dt) <= 1e-13 * dt
            w.timestep(t, dt)
            g.TimeStep(t, dt)
            t += dt
        Ug, Uo = g.get_state(), o.array("U")
        sc = np.max(np.abs(Uo).reshape(-1, 5), axis=0)
        sc[1:4] = sc[1:4].max()
        assert np.all(np.max(np.abs(Ug - Uo).reshape(-1, 5), axis=0) <= 1e-12 * sc)
    w.close()
