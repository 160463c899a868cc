"""N = 11 on the CUDA path: the degree of the reference's regressioncheck/checks/tgv/oInt (4^3 elements, Gauss nodes, weak form,
cut-off overintegration NUnder = 7, RoeEntropyFix, BR1).  1728 nodes per element do not fit one CTA of one thread per node: the
element kernels run on a thread-block cluster of two CTAs whose shared memories together hold the element's work planes
(layout.cuh: Dim::PAIR, PairPtr).  Same kernels, same source -- checked against the oracle and against the reference's CSV rows."""
import numpy as np
import pytest

from flexi_b200.gpu import FlexiGpu
from flexi_b200.hostmesh import basis as B
from flexi_b200.hostmesh.analyze import AnalyzeBasis, tgv_analyze
from flexi_b200.hostmesh.case import TGV_EQN, build_case, tgv_box
from flexi_b200.hostmesh.mesh import BoxSpec
from oracle import oracle as O

pytestmark = pytest.mark.gpu
G, GL = B.NODE_GAUSS, B.NODE_GAUSS_LOBATTO
CURVED = dict(x0=(-1.0, -1.0, -1.0), length=(2.0, 2.0, 2.0), ngeo=2, deform=0.1)

CASES = [
    ("G-weak-ns-oint", G, dict(parabolic=True, riemann="roeentropyfix", mu0=1e-2, n_var_lifting=4), 7),
    ("G-weak-euler", G, dict(parabolic=False, riemann="hllc"), None),
    ("GL-split-ns", GL, dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="lf", mu0=1e-2), None),
    ("GL-weak-ns-br2", GL, dict(parabolic=True, riemann="roe", mu0=1e-2, lifting=2), None),
    ("G-split-ns", G, dict(parabolic=True, split_dg=True, split_dg_type="kg", riemann="lf", mu0=1e-2), None),
]


@pytest.mark.parametrize("name,node,eqn,n_under", CASES, ids=[c[0] for c in CASES])
def test_n11_matches_oracle(name, node, eqn, n_under):
    """One DG operator (Ut and the side arrays, rel 1e-12), CalcTimeStep, one LSERK step on the sine-deformed periodic box."""
    N = 11
    gsplit = bool(eqn.get("split_dg")) and node == G
    r = build_case(BoxSpec((2, 2, 2), **CURVED), N, node, eqn, with_ja_face=gsplit)[0]
    U0 = r.eqn.smooth_state(r.mesh.Elem_xGP, length=2.0)
    w = O.OracleWorld([(r.desc, r.keep)])
    o = w.ranks[0]
    if gsplit:
        o.set_face_metrics(r.mesh.Ja_Face, r.mesh.Ja_slave)
    mat = B.overintegration_mat(r.basis.xGP, n_under) if n_under else None
    if mat is not None:
        o.set_overintegration(1, mat)
    w.set_state([U0])
    g = FlexiGpu(r)
    try:
        if mat is not None:
            g.set_overintegration(1, mat)
        g.set_state(U0)
        w.dg_time_derivative(0.0)
        ut, ref = g.DGTimeDerivative_weakForm(0.0), o.array("Ut")
        for nm in ("U_master", "U_slave", "Flux"):
            a, b = g.get_array(nm), o.array(nm)
            assert np.max(np.abs(a - b)) <= 1e-12 * np.max(np.abs(b)), nm
        scale = np.max(np.abs(ref).reshape(-1, 5), axis=0)
        assert np.all(np.max(np.abs(ut - ref).reshape(-1, 5), axis=0) <= 2e-12 * scale)
        dt = w.calc_timestep()[0]
        assert abs(g.CalcTimeStep()[0] - dt) <= 1e-13 * dt
        w.timestep(0.0, dt)
        g.TimeStep(0.0, dt)
        Ug, Uo = g.get_state(), o.array("U")
        sc = np.max(np.abs(Uo).reshape(-1, 5), axis=0)
        assert np.all(np.max(np.abs(Ug - Uo).reshape(-1, 5), axis=0) <= 1e-12 * sc)
    finally:
        g.close()
        w.close()


def test_n10_is_refused():
    r = build_case(BoxSpec((2, 2, 2), **CURVED), 10, G, dict(parabolic=False, riemann="lf"))[0]
    with pytest.raises(Exception, match="not instantiated"):
        FlexiGpu(r)


def test_gpu_reproduces_reference_tgv_oint_csv():
    """regressioncheck/checks/tgv/oInt/TGV_Re1600_OInt_TGVAnalysis_Reference.csv rows 1 and 2 (state after 10 steps = 50 RHS
    evaluations with the time step recomputed every step) on the CUDA path; the settings are those of the case's parameter.ini
    (see tests/test_oracle_tgv_regression.py, where the oracle reproduces the same rows)."""
    from test_oracle_tgv_regression import OINT_EQN, OROW1, OROW2

    r = build_case(tgv_box(4), 11, G, OINT_EQN)[0]
    mat = B.overintegration_mat(r.basis.xGP, 7)
    ab = AnalyzeBasis(r.basis, n_analyze=10)
    vol = ab.volume(r.mesh.sJ)
    # timedisc.f90:115-123: with OverintegrationType=1 the host filters the initial solution once before the first DG operator call
    U0 = np.ascontiguousarray(np.einsum("li,mj,nk,ekjiv->enmlv", mat, mat, mat, r.eqn.tgv_state(r.mesh.Elem_xGP)))
    g = FlexiGpu(r)
    try:
        g.set_overintegration(1, mat)
        g.keep_gradients(True)
        g.set_state(U0)

        def analyze():
            out = tgv_analyze(ab, r.eqn, r.mesh, g.get_state(), g.get_array("gradUx"), g.get_array("gradUy"), g.get_array("gradUz"), vol)
            out["uPrime"] = np.sqrt(2.0 / 3.0 * out["Ekin"])
            return out

        g.DGTimeDerivative_weakForm(0.0)
        row1 = analyze()
        for k, ref in OROW1.items():
            assert abs(row1[k] - ref) <= 5e-13 * max(1.0, abs(ref)), (k, row1[k], ref)
        t = 0.0
        for _ in range(10):
            dt, _, _, err = g.CalcTimeStep()
            assert err == 0
            g.TimeStep(t, dt)
            t += dt
        row2 = analyze()
        row2["Time"] = t
        for k, ref in OROW2.items():
            tol = {"ED_D": 1e-8, "DR_p": 1e-8, "Vorticity_max": 1e-11}.get(k, 5e-13)
            assert abs(row2[k] - ref) <= tol * max(abs(ref), 1e-30) + 1e-18, (k, row2[k], ref)
    finally:
        g.close()
