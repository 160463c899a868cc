"""Split form on Gauss nodes (SPLIT_DG with PP_NodeType = GAUSS => PP_EntropyVars = 1) on the CUDA path:
entropy-projected traces (prolongtoface.t90:151-204), two-point volume fluxes with the Gauss DVolSurf (dg.f90:184-194) and the
two-point surface correction (surfint.t90:90-137,190-209), against the oracle (pinned on regressioncheck/checks/tgv/splitGauss in
tests/test_oracle_tgv_regression.py) and against the same reference CSV rows directly."""
import numpy as np
import pytest

from flexi_b200.gpu import FlexiGpu
from flexi_b200.hostmesh import basis as B
from flexi_b200.hostmesh.analyze import AnalyzeBasis, tgv_analyze
from flexi_b200.hostmesh.case import TGV_EQN, build_case, tgv_box
from flexi_b200.hostmesh.mesh import BoxSpec
from oracle import oracle as O

pytestmark = pytest.mark.gpu
G = B.NODE_GAUSS
CURVED = dict(x0=(-1.0, -1.0, -1.0), length=(2.0, 2.0, 2.0), ngeo=2, deform=0.1)


@pytest.mark.parametrize("N,eqn", [
    (3, dict(parabolic=False, split_dg=True, split_dg_type="sd", riemann="lf")),
    (4, dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="roeentropyfix", mu0=1e-2)),
    (5, dict(parabolic=True, split_dg=True, split_dg_type="kg", riemann="lf", mu0=1e-2, n_var_lifting=4)),
    (7, dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="lf", mu0=1e-2, lifting=2)),
])
def test_gauss_split_matches_oracle(N, eqn):
    r = build_case(BoxSpec((3, 2, 2), **CURVED), N, G, eqn, with_ja_face=True)[0]
    U0 = r.eqn.smooth_state(r.mesh.Elem_xGP, length=2.0)
    w = O.OracleWorld([(r.desc, r.keep)])
    w.ranks[0].set_face_metrics(r.mesh.Ja_Face, r.mesh.Ja_slave)
    w.set_state([U0])
    g = FlexiGpu(r)
    try:
        g.set_face_metrics(r.mesh.Ja_Face, r.mesh.Ja_slave)
        g.set_state(U0)
        w.dg_time_derivative(0.0)
        ut = g.DGTimeDerivative_weakForm(0.0)
        o = w.ranks[0]
        ref = o.array("Ut")
        for name in ("U_master", "U_slave"):
            a, b = g.get_array(name), o.array(name)
            assert np.max(np.abs(a - b)) <= 1e-12 * np.max(np.abs(b)), name
        d = np.max(np.abs(ut - ref).reshape(-1, 5), axis=0)
        scale = np.max(np.abs(ref).reshape(-1, 5), axis=0)
        scale[1:4] = scale[1:4].max()
        assert np.all(d <= 1e-12 * scale), (d, scale)
        t = 0.0
        for _ in range(2):
            dt = w.calc_timestep()[0]
            assert abs(g.CalcTimeStep()[0] - dt) <= 1e-13 * dt
            w.timestep(t, dt)
            g.TimeStep(t, dt)
            t += dt
        Ug, Uo = g.get_state(), o.array("U")
        sc = np.max(np.abs(Uo).reshape(-1, 5), axis=0)
        assert np.all(np.max(np.abs(Ug - Uo).reshape(-1, 5), axis=0) <= 1e-12 * sc)
    finally:
        g.close()
        w.close()


def test_gauss_split_face_metrics_through_the_descriptor():
    """ABI 2: Ja_Face / Ja_slave are descriptor fields -- neither the library nor the oracle needs the extra setter call."""
    eqn = dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="lf", mu0=1e-2)
    r = build_case(BoxSpec((3, 2, 2), **CURVED), 4, G, eqn, with_ja_face=True)[0]
    assert bool(r.desc.Ja_Face) and bool(r.desc.Ja_slave)
    U0 = r.eqn.smooth_state(r.mesh.Elem_xGP, length=2.0)
    w = O.OracleWorld([(r.desc, r.keep)])
    w.set_state([U0])
    w.dg_time_derivative(0.0)
    ref = np.array(w.ranks[0].array("Ut"))  # a copy: the oracle's arrays are gone after close()
    w.close()
    with FlexiGpu(r) as g:
        g.set_state(U0)
        ut = g.DGTimeDerivative_weakForm(0.0)
    scale = np.max(np.abs(ref).reshape(-1, 5), axis=0)
    scale[1:4] = scale[1:4].max()
    assert np.all(np.max(np.abs(ut - ref).reshape(-1, 5), axis=0) <= 1e-12 * scale)


def test_gauss_split_needs_face_metrics():
    eqn = dict(parabolic=False, split_dg=True, split_dg_type="pi", riemann="lf")
    r = build_case(BoxSpec((2, 2, 2), **CURVED), 2, G, eqn)[0]
    g = FlexiGpu(r)
    try:
        g.set_state(r.eqn.smooth_state(r.mesh.Elem_xGP, length=2.0))
        with pytest.raises(Exception, match="Ja_Face"):
            g.DGTimeDerivative_weakForm(0.0)
    finally:
        g.close()


def test_gpu_reproduces_reference_tgv_split_gauss_csv():
    """regressioncheck/checks/tgv/splitGauss/TGV_Re1600_Split_TGVAnalysis_Reference.csv rows 1 and 2 on the CUDA path."""
    from test_oracle_tgv_regression import GROW1, GROW2

    eqn = dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="roeentropyfix", cfl_scale=0.8, dfl_scale=0.8,
               n_var_lifting=4, **TGV_EQN)
    r = build_case(tgv_box(8), 5, G, eqn, with_ja_face=True)[0]
    ab = AnalyzeBasis(r.basis, n_analyze=10)
    vol = ab.volume(r.mesh.sJ)
    g = FlexiGpu(r)
    try:
        g.set_face_metrics(r.mesh.Ja_Face, r.mesh.Ja_slave)
        g.keep_gradients(True)
        g.set_state(r.eqn.tgv_state(r.mesh.Elem_xGP))

        def analyze():
            out = tgv_analyze(ab, r.eqn, r.mesh, g.get_state(), g.get_array("gradUx"), g.get_array("gradUy"), g.get_array("gradUz"), vol)
            out["uPrime"] = np.sqrt(2.0 / 3.0 * out["Ekin"])
            return out

        g.DGTimeDerivative_weakForm(0.0)
        row1 = analyze()
        for k, ref in GROW1.items():
            assert abs(row1[k] - ref) <= 2e-13 * max(1.0, abs(ref)) + 1e-13 * abs(ref), (k, row1[k], ref)
        t = 0.0
        for _ in range(10):
            dt, _, _, err = g.CalcTimeStep()
            assert err == 0
            g.TimeStep(t, dt)
            t += dt
        row2 = analyze()
        row2["Time"] = t
        for k, ref in GROW2.items():
            tol = {"ED_D": 1e-8, "DR_p": 1e-8, "Vorticity_max": 1e-11}.get(k, 5e-13)
            assert abs(row2[k] - ref) <= tol * max(abs(ref), 1e-30) + 1e-18, (k, row2[k], ref)
    finally:
        g.close()
