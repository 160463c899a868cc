"""The CUDA path on the reference's own NACA0012 mesh file and state (tests/test_hopr_mesh.py describes the files and what they
pin on the host side): N = 3, Gauss nodes, weak form, RoeEntropyFix, BR1, adiabatic wall + far field -- the settings of
regressioncheck/checks/naca/3D -- and the same mesh with the split form on Gauss-Lobatto nodes and with HLLC + BR2 (BASELINE configs[3])."""
import numpy as np
import pytest

from flexi_b200.gpu import FlexiGpu
from flexi_b200.hostmesh import basis as B
from oracle import oracle as O
from test_hopr_mesh import STATE, naca_rank, read_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("node,extra", [
    (B.NODE_GAUSS, {}),
    (B.NODE_GAUSS, dict(riemann="hllc", lifting=2)),
    (B.NODE_GAUSS_LOBATTO, dict(split_dg=True, split_dg_type="pi", riemann="lf")),
], ids=["G-weak-roeef-br1", "G-weak-hllc-br2", "GL-split-lf"])
def test_reference_naca_mesh_and_state_match_oracle(node, extra):
    r = naca_rank(node=node, **extra)[0]
    U0, _, ed = read_state(STATE)  # DG_Solution of the reference at t = 10 (Gauss nodes; taken as nodal values on GL nodes too)
    w = O.OracleWorld([(r.desc, r.keep)])
    o = w.ranks[0]
    w.set_state([U0])
    w.dg_time_derivative(10.0)
    with FlexiGpu(r) as g:
        g.set_state(U0)
        ut, ref = g.DGTimeDerivative_weakForm(10.0), o.array("Ut")
        # gate: rel 1e-12, or 4x the oracle's own strict-vs-FMA build noise where that is larger (1/J spans five decades here)
        noise, _ = O.ut_noise_floor([(r.desc, r.keep)], [U0])
        scale = np.max(np.abs(ref).reshape(-1, 5), axis=0)
        scale[1:4] = scale[1:4].max()  # the flow is two-dimensional: w-momentum ~ 0, judged against the common momentum scale
        d = np.max(np.abs(ut - ref).reshape(-1, 5), axis=0)
        assert np.all(d <= np.maximum(1e-12 * scale, 4.0 * noise)), (d, scale, noise)
        nbc = r.mesh.nBCSides  # (the device keeps the lifting boundary state as a virtual slave trace on BC sides; the oracle has none)
        for nm, lo in (("U_master", 0), ("U_slave", nbc)):
            a, b = g.get_array(nm)[lo:], o.array(nm)[lo:]
            assert np.max(np.abs(a - b)) <= 1e-12 * np.max(np.abs(b)), nm
        dt_o = w.calc_timestep()[0]
        dt_g = g.CalcTimeStep()[0]
        assert abs(dt_g - dt_o) <= 1e-13 * dt_o
        if node == B.NODE_GAUSS and not extra:
            # the reference's own per-element time step (ElemData 'dt', one step earlier): see test_hopr_mesh.py
            assert abs(dt_g / ed["dt"].min() - 1.0) < 1e-5
        t = 10.0
        for _ in range(3):
            dt = w.calc_timestep()[0]
            assert abs(g.CalcTimeStep()[0] - dt) <= 1e-13 * dt
            w.timestep(t, dt)
            g.TimeStep(t, dt)
            t += dt
        Ug, Uo = g.get_state(), o.array("U")
        sc = np.max(np.abs(Uo).reshape(-1, 5), axis=0)
        sc[1:4] = sc[1:4].max()
        assert np.all(np.max(np.abs(Ug - Uo).reshape(-1, 5), axis=0) <= 1e-11 * sc)
    w.close()
