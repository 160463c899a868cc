"""Parity of the CUDA path (through the C ABI) with the CPU oracle on identical synthetic meshes and states.

north_star gate: one Ut evaluation within relative 1e-12 (max|diff| / max|ref| per variable); the state after a fixed
number of RK steps and the L2/Linf norms within 1e-10.
"""
import numpy as np
import pytest

from flexi_b200.gpu import FlexiGpu
from flexi_b200.hostmesh import basis as B
from flexi_b200.hostmesh.analyze import AnalyzeBasis, calc_error_norms
from flexi_b200.hostmesh.case import TGV_EQN, build_case, tgv_box
from flexi_b200.hostmesh.mesh import BoxSpec
from oracle import oracle as O

pytestmark = pytest.mark.gpu

CURVED = dict(x0=(-1.0, -1.0, -1.0), length=(2.0, 2.0, 2.0), ngeo=2, deform=0.1)
TOL_UT = 1e-12
G, GL = B.NODE_GAUSS, B.NODE_GAUSS_LOBATTO


def rel(a, ref):
    """max|a-ref| / max|ref| per last-axis variable."""
    a2, r2 = a.reshape(-1, a.shape[-1]), ref.reshape(-1, ref.shape[-1])
    den = np.maximum(np.max(np.abs(r2), axis=0), 1e-300)
    return np.max(np.abs(a2 - r2), axis=0) / den


def assert_ut_parity(ut, ref, noise=None):
    """rel 1e-12 per variable; where the reference's own strict-vs-FMA builds differ by more (ill-conditioned TGV at Ma 0.1),
    the gate is 4x that noise floor (oracle.ut_noise_floor)."""
    d = np.max(np.abs(ut - ref).reshape(-1, 5), axis=0)
    tol = TOL_UT * np.max(np.abs(ref).reshape(-1, 5), axis=0)
    if noise is not None:
        tol = np.maximum(tol, 4.0 * noise)
    assert np.all(d <= tol), (d, tol)


def run_pair(spec, N, node, eqn, state="smooth"):
    ranks = build_case(spec, N, node, eqn)
    r = ranks[0]
    if state == "smooth":
        U0 = r.eqn.smooth_state(r.mesh.Elem_xGP, length=spec.length[0])
    elif state == "tgv":
        U0 = r.eqn.tgv_state(r.mesh.Elem_xGP)
    else:
        U0 = r.eqn.random_state(r.mesh.Elem_xGP.shape[:-1])
    w = O.OracleWorld([(r.desc, r.keep)])
    w.set_state([U0])
    g = FlexiGpu(r)
    g.set_state(U0)
    return r, w, g, U0


# (N, node, eqn overrides) -- weak form
WEAK = [
    (3, G, dict(parabolic=False, riemann="lf")),
    (4, G, dict(parabolic=False, riemann="roe")),
    (4, GL, dict(parabolic=False, riemann="hllc")),
    (2, G, dict(parabolic=True, riemann="roeentropyfix", mu0=1e-2)),
    (3, GL, dict(parabolic=True, riemann="lf", mu0=1e-2)),
    (5, G, dict(parabolic=True, riemann="hllc", mu0=1e-2)),
    (5, GL, dict(parabolic=True, riemann="roe", mu0=1e-2, n_var_lifting=4)),
    (7, G, dict(parabolic=True, riemann="roe", mu0=1e-2, n_var_lifting=4)),
    (7, GL, dict(parabolic=True, riemann="roeentropyfix", mu0=1e-2)),
    (9, G, dict(parabolic=True, riemann="hll", mu0=1e-2)),
    (9, GL, dict(parabolic=True, riemann="hllem", mu0=1e-2)),
    (1, G, dict(parabolic=True, riemann="roel2", mu0=1e-2)),
    (6, GL, dict(parabolic=True, riemann="hlle", mu0=1e-2)),
    (8, G, dict(parabolic=False, riemann="roeentropyfix")),
]
SPLIT = [
    (3, GL, dict(parabolic=False, split_dg=True, split_dg_type="sd", riemann="lf")),
    (4, GL, dict(parabolic=False, split_dg=True, split_dg_type="pi", riemann="roe")),
    (5, GL, dict(parabolic=True, split_dg=True, split_dg_type="du", riemann="roeentropyfix", mu0=1e-2)),
    (5, GL, dict(parabolic=True, split_dg=True, split_dg_type="kg", riemann="lf", mu0=1e-2, n_var_lifting=4)),
    (7, GL, dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="lf", mu0=1e-2)),
    (7, GL, dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="roeentropyfix", mu0=1e-2, n_var_lifting=4)),
    (9, GL, dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="roel2", mu0=1e-2)),
    (1, GL, dict(parabolic=False, split_dg=True, split_dg_type="sd", riemann="lf")),
    (1, GL, dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="lf", mu0=1e-2)),
    (7, GL, dict(parabolic=False, split_dg=True, split_dg_type="du", riemann="roe")),
    (7, GL, dict(parabolic=True, split_dg=True, split_dg_type="kg", riemann="roe", mu0=1e-2)),
    (7, GL, dict(parabolic=True, split_dg=True, split_dg_type="sd", riemann="lf", mu0=1e-2, n_var_lifting=4)),
]


def _ids(cases):
    return [f"N{n}-{'G' if nd == G else 'GL'}-" + "-".join(f"{k}={v}" for k, v in e.items() if k != "mu0") for n, nd, e in cases]


@pytest.mark.parametrize("N,node,eqn", WEAK + SPLIT, ids=_ids(WEAK + SPLIT))
def test_ut_and_side_arrays_match_oracle_on_curved_mesh(N, node, eqn):
    """One DGTimeDerivative_weakForm on the sine-deformed periodic box: Ut, traces, fluxes and lifted gradients."""
    r, w, g, _ = run_pair(BoxSpec((3, 2, 2), **CURVED), N, node, eqn)
    try:
        w.dg_time_derivative(0.0)
        ut = g.DGTimeDerivative_weakForm(0.0)
        o = w.ranks[0]
        assert np.all(rel(ut, o.array("Ut")) < TOL_UT), rel(ut, o.array("Ut"))
        for name in ("U_master", "U_slave"):
            assert np.all(rel(g.get_array(name), o.array(name)) < 1e-14), name
        assert np.all(rel(g.get_array("Flux"), o.array("Flux")) < TOL_UT)
        if eqn.get("parabolic"):
            for name in ("gradUx", "gradUy", "gradUz", "gradUx_master", "gradUy_master", "gradUz_master", "gradUx_slave",
                         "gradUy_slave", "gradUz_slave"):
                a, ref = g.get_array(name), o.array(name)
                scale = max(np.max(np.abs(o.array(n2))) for n2 in ("gradUx", "gradUy", "gradUz"))
                assert np.max(np.abs(a - ref)) < 1e-12 * scale, name
    finally:
        g.close()
        w.close()


ORIENTED = [WEAK[1], WEAK[2], WEAK[5], WEAK[6], WEAK[10], SPLIT[1], SPLIT[3], SPLIT[4], SPLIT[5], SPLIT[6],
            (5, G, dict(parabolic=True, riemann="hllc", mu0=1e-2, lifting=2, eta_br2=2.0)),          # BASELINE configs[3] flavour
            (5, GL, dict(parabolic=True, riemann="roe", mu0=1e-2, lifting=2, eta_br2=2.0, split_dg=True, split_dg_type="pi"))]


@pytest.mark.parametrize("N,node,eqn", ORIENTED, ids=_ids(ORIENTED))
def test_unstructured_orientations_match_oracle(N, node, eqn):
    """BASELINE configs[3] says "unstructured": every element of the curved periodic box sits in its own randomly rotated local
    frame, so all five flips and every (master side, slave side) pairing occur (mappings.f90:188-450).  Ut, traces, fluxes and
    lifted gradients against the oracle, whose own handling of these meshes is pinned by tests/test_host_mesh.py."""
    r, w, g, _ = run_pair(BoxSpec((3, 3, 2), orient_seed=3, **CURVED), N, node, eqn)
    try:
        assert np.all(np.bincount(r.mesh.ElemToSide[:, :, 1].ravel(), minlength=5) > 0)
        w.dg_time_derivative(0.0)
        ut = g.DGTimeDerivative_weakForm(0.0)
        o = w.ranks[0]
        assert np.all(rel(ut, o.array("Ut")) < TOL_UT), rel(ut, o.array("Ut"))
        for name in ("U_master", "U_slave"):
            assert np.all(rel(g.get_array(name), o.array(name)) < 1e-14), name
        assert np.all(rel(g.get_array("Flux"), o.array("Flux")) < TOL_UT)
        if eqn.get("parabolic"):
            scale = max(np.max(np.abs(o.array(n2))) for n2 in ("gradUx", "gradUy", "gradUz"))
            for name in ("gradUx", "gradUy", "gradUz", "gradUx_master", "gradUy_master", "gradUz_master", "gradUx_slave",
                         "gradUy_slave", "gradUz_slave"):
                assert np.max(np.abs(g.get_array(name) - o.array(name))) < 1e-12 * scale, name
        # and three LSERK steps on top
        t = 0.0
        for _ in range(3):
            dt_o = w.calc_timestep()[0]
            assert abs(g.CalcTimeStep()[0] - dt_o) <= 1e-13 * dt_o
            w.timestep(t, dt_o)
            g.TimeStep(t, dt_o)
            t += dt_o
        assert np.all(rel(g.get_state(), o.array("U")) < 1e-12)
    finally:
        g.close()
        w.close()


@pytest.mark.parametrize("N,node,eqn", [WEAK[4], WEAK[7], SPLIT[4], SPLIT[5]], ids=_ids([WEAK[4], WEAK[7], SPLIT[4], SPLIT[5]]))
def test_rk_steps_and_timestep_match_oracle(N, node, eqn):
    """CalcTimeStep + 3 LSERK steps: dt to 1e-14, state to 1e-12, L2/Linf norms against the initial state to 1e-10."""
    r, w, g, U0 = run_pair(BoxSpec((3, 2, 2), **CURVED), N, node, eqn)
    try:
        t = 0.0
        for _ in range(3):
            dt_o, c_o, v_o, e_o = w.calc_timestep()
            dt_g, c_g, v_g, e_g = g.CalcTimeStep()
            assert e_o == 0 and e_g == 0
            assert abs(dt_g - dt_o) <= 1e-13 * dt_o and abs(c_g - c_o) <= 1e-13 * c_o
            if eqn.get("parabolic"):
                assert abs(v_g - v_o) <= 1e-13 * v_o
            w.timestep(t, dt_o)
            g.TimeStep(t, dt_o)
            t += dt_o
        Ug, Uo = g.get_state(), w.ranks[0].array("U")
        assert np.all(rel(Ug, Uo) < 1e-12), rel(Ug, Uo)
        ab = AnalyzeBasis(r.basis)
        vol = ab.volume(r.mesh.sJ)
        exact = lambda x: r.eqn.smooth_state(x, length=2.0)  # noqa: E731
        l2g, lig = calc_error_norms(ab, r.mesh, Ug, exact, vol)
        l2o, lio = calc_error_norms(ab, r.mesh, Uo, exact, vol)
        assert np.all(np.abs(np.sqrt(l2g / vol) - np.sqrt(l2o / vol)) <= 1e-10 * np.maximum(np.sqrt(l2o / vol), 1e-30))
        assert np.all(np.abs(lig - lio) <= 1e-10 * np.maximum(lio, 1e-30))
    finally:
        g.close()
        w.close()


def test_tgv_config1_flavour_small():
    """BASELINE config 1 flavour (TGV, N=7, Gauss, Roe, BR1, weak form) on a 4^3 box: Ut to 1e-12, 2 steps to 1e-12."""
    eqn = dict(parabolic=True, riemann="roe", cfl_scale=0.8, dfl_scale=0.8, **TGV_EQN)
    r, w, g, _ = run_pair(tgv_box(4), 7, G, eqn, state="tgv")
    try:
        w.dg_time_derivative(0.0)
        ut = g.DGTimeDerivative_weakForm(0.0)
        noise, _ = O.ut_noise_floor([(r.desc, r.keep)], [r.eqn.tgv_state(r.mesh.Elem_xGP)])
        assert_ut_parity(ut, w.ranks[0].array("Ut"), noise)
        t = 0.0
        for _ in range(2):
            dt, *_ = w.calc_timestep()
            w.timestep(t, dt)
            g.TimeStep(t, dt)
            t += dt
        # TGV: w-momentum is ~0 (max 1e-5), so momenta are judged against the common momentum scale
        Ug, Uo = g.get_state(), w.ranks[0].array("U")
        scale = np.array([np.abs(Uo[..., 0]).max(), np.abs(Uo[..., 1:4]).max(), np.abs(Uo[..., 1:4]).max(),
                          np.abs(Uo[..., 1:4]).max(), np.abs(Uo[..., 4]).max()])
        assert np.all(np.max(np.abs(Ug - Uo).reshape(-1, 5), axis=0) <= 1e-12 * scale)
    finally:
        g.close()
        w.close()


def test_tgv_config2_flavour_small():
    """BASELINE config 2 flavour (TGV, N=7, split-form Gauss-Lobatto Pirozzoli + Lax-Friedrichs) on a 4^3 box."""
    eqn = dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="lf", cfl_scale=0.8, dfl_scale=0.8, **TGV_EQN)
    r, w, g, _ = run_pair(tgv_box(4), 7, GL, eqn, state="tgv")
    try:
        w.dg_time_derivative(0.0)
        ut = g.DGTimeDerivative_weakForm(0.0)
        noise, _ = O.ut_noise_floor([(r.desc, r.keep)], [r.eqn.tgv_state(r.mesh.Elem_xGP)])
        assert_ut_parity(ut, w.ranks[0].array("Ut"), noise)
        t = 0.0
        for _ in range(2):
            dt, *_ = w.calc_timestep()
            w.timestep(t, dt)
            g.TimeStep(t, dt)
            t += dt
        # TGV: w-momentum is ~0 (max 1e-5), so momenta are judged against the common momentum scale
        Ug, Uo = g.get_state(), w.ranks[0].array("U")
        scale = np.array([np.abs(Uo[..., 0]).max(), np.abs(Uo[..., 1:4]).max(), np.abs(Uo[..., 1:4]).max(),
                          np.abs(Uo[..., 1:4]).max(), np.abs(Uo[..., 4]).max()])
        assert np.all(np.max(np.abs(Ug - Uo).reshape(-1, 5), axis=0) <= 1e-12 * scale)
    finally:
        g.close()
        w.close()


def test_gradient_traces_after_time_steps():
    """The RK stages of the split-form Gauss-Lobatto path hand n.F_visc per side to the side kernel and leave gradient traces on
    BC sides only (viscous-trace protocol); a gradient-trace request after time steps must still return the
    traces of the current state (recomputed), like the volume gradients."""
    eqn = dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="lf", mu0=1e-3)
    r, w, g, _ = run_pair(BoxSpec((3, 2, 2), **CURVED), 5, GL, eqn)
    try:
        t = 0.0
        for _ in range(2):
            dt, *_ = w.calc_timestep()
            w.timestep(t, dt)
            g.TimeStep(t, dt)
            t += dt
        o = w.ranks[0]
        assert np.all(rel(g.get_state(), o.array("U")) < 1e-12)
        w.dg_time_derivative(t)  # periodic box, no source: the operator does not depend on t
        scale = max(np.max(np.abs(o.array(n2))) for n2 in ("gradUx", "gradUy", "gradUz"))
        for name in ("gradUx_master", "gradUy_master", "gradUz_master", "gradUx_slave", "gradUy_slave", "gradUz_slave", "gradUx",
                     "gradUy", "gradUz"):
            assert np.max(np.abs(g.get_array(name) - o.array(name))) < 1e-12 * scale, name
        # and the time stepping goes on unharmed
        dt, *_ = w.calc_timestep()
        w.timestep(t, dt)
        g.TimeStep(t, dt)
        assert np.all(rel(g.get_state(), o.array("U")) < 1e-12)
    finally:
        g.close()
        w.close()


def test_random_field_stress():
    """Discontinuous random state (SURVEY 8d stress input): every face has an O(1) jump, all Riemann branches are exercised."""
    for riemann in ("hllc", "roeentropyfix", "hll"):
        # Gauss-Lobatto: face values are nodal values, so the random traces stay physical (no interpolation overshoot)
        r, w, g, _ = run_pair(BoxSpec((2, 2, 3), **CURVED), 4, GL, dict(parabolic=True, riemann=riemann, mu0=1e-3), state="random")
        try:
            w.dg_time_derivative(0.0)
            ut = g.DGTimeDerivative_weakForm(0.0)
            assert np.all(rel(ut, w.ranks[0].array("Ut")) < TOL_UT), (riemann, rel(ut, w.ranks[0].array("Ut")))
        finally:
            g.close()
            w.close()


def test_freestream_preservation_on_device():
    """Constant state on the curved mesh => Ut = 0 (run_basic/freestream_3D: L2 <= 1e-12)."""
    ranks = build_case(BoxSpec((3, 2, 2), **CURVED), 4, G, dict(parabolic=True, riemann="roe", mu0=1e-2))
    r = ranks[0]
    prim = np.zeros(r.mesh.Elem_xGP.shape[:-1] + (6,))
    prim[...] = (1.0, 0.3, -0.2, 0.1, 0.7, 0.0)
    U0 = r.eqn.prim_to_cons(prim)
    with FlexiGpu(r) as g:
        g.set_state(U0)
        ut = g.DGTimeDerivative_weakForm(0.0)
        assert np.max(np.abs(ut)) < 1e-11


def test_nan_state_reports_errtype():
    """calctimestep.f90:157-160: NaN density => errType != 0 instead of a silent dt."""
    ranks = build_case(tgv_box(2), 3, G, dict(parabolic=True, **TGV_EQN))
    r = ranks[0]
    U0 = r.eqn.tgv_state(r.mesh.Elem_xGP)
    U0[1, 2, 1, 0, 0] = np.nan
    with FlexiGpu(r) as g:
        g.set_state(U0)
        assert g.CalcTimeStep()[3] != 0


def test_gpu_reproduces_reference_tgv_regression_csv():
    """The reference's own end-to-end pin (regressioncheck/checks/tgv/split/TGVAnalysis_Reference.csv) replayed on
    the CUDA path: 8^3 elements, LGL N=7, split form (PI), RoeEntropyFix, BR1 with nVarLifting=4, carpenterrk4-5, CFL/DFL 0.8,
    NAnalyze=10, dt recomputed every step.  Row 1 = t=0 (analytic), row 2 = after 10 steps; the intermediate AnalyzeTestCase sees
    the gradients left by the last RK stage (timedisc_func.f90:323-338), hence keep_gradients."""
    from flexi_b200.hostmesh.analyze import AnalyzeBasis, tgv_analyze
    from test_oracle_tgv_regression import ROW1, ROW2

    eqn = dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="roeentropyfix", cfl_scale=0.8, dfl_scale=0.8,
               n_var_lifting=4, **TGV_EQN)
    r = build_case(tgv_box(8), 7, GL, eqn)[0]
    ab = AnalyzeBasis(r.basis, n_analyze=10)
    vol = ab.volume(r.mesh.sJ)
    g = FlexiGpu(r)
    try:
        g.keep_gradients(True)
        g.set_state(r.eqn.tgv_state(r.mesh.Elem_xGP))

        def analyze():
            # the host-side analysis of the downloaded arrays, and the on-device analysis (flexi_gpu_analyze_tgv): both must
            # reproduce the reference rows, and each other
            out = tgv_analyze(ab, r.eqn, r.mesh, g.get_state(), g.get_array("gradUx"), g.get_array("gradUy"), g.get_array("gradUz"), vol)
            out["uPrime"] = np.sqrt(2.0 / 3.0 * out["Ekin"])
            dev = g.tgv_normalise([g.AnalyzeTestcase_sums(ab.Vdm, ab.w3)], vol, r.eqn.mu0)
            for k, v in out.items():
                tol = {"ED_D": 1e-7, "DR_p": 1e-7}.get(k, 1e-12)  # differences of nearly equal numbers ...
                floor = {"DR_p": 1e-13, "ED_D": 1e-16}.get(k, 1e-18)  # ... that vanish at t=0 (p ~ 71 times div u ~ 1e-15)
                assert abs(dev[k] - v) <= tol * max(abs(v), 1e-30) + floor, (k, dev[k], v)
            return dev

        g.DGTimeDerivative_weakForm(0.0)
        row1 = analyze()
        for k, ref in ROW1.items():
            assert abs(row1[k] - ref) <= 2e-13 * max(1.0, abs(ref)) + 1e-13 * abs(ref), (k, row1[k], ref)
        t = 0.0
        for _ in range(10):
            dt, _, _, err = g.CalcTimeStep()
            assert err == 0
            g.TimeStep(t, dt)
            t += dt
        row2 = analyze()
        row2["Time"] = t
        for k, ref in ROW2.items():
            tol = {"ED_D": 1e-9, "DR_p": 1e-9, "Vorticity_max": 1e-11}.get(k, 5e-13)
            assert abs(row2[k] - ref) <= tol * max(abs(ref), 1e-30) + 1e-18, (k, row2[k], ref)
    finally:
        g.close()


@pytest.mark.parametrize("nel,N,node,eqn", [
    ((2, 2, 2), 3, GL, dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="lf", mu0=1e-2)),
    ((2, 2, 2), 7, GL, dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="roe", mu0=1e-2)),
    ((2, 2, 2), 4, G, dict(parabolic=True, riemann="hllc", mu0=1e-2)),
    ((2, 2, 2), 2, GL, dict(parabolic=False, riemann="lf")),
])
def test_smallest_periodic_meshes(nel, N, node, eqn):
    """Edge case of the side bookkeeping: two elements per periodic direction (the smallest box the host mesh, like HOPR,
    allows), so both neighbours of an element in a direction are the same element, joined by two different sides; the grid is
    far smaller than one CTA wave."""
    r, w, g, _ = run_pair(BoxSpec(nel, **CURVED), N, node, eqn)
    try:
        w.dg_time_derivative(0.0)
        ut = g.DGTimeDerivative_weakForm(0.0)
        assert np.all(rel(ut, w.ranks[0].array("Ut")) < TOL_UT), rel(ut, w.ranks[0].array("Ut"))
        dt, *_ = w.calc_timestep()
        assert abs(g.CalcTimeStep()[0] - dt) <= 1e-13 * dt
        for k in range(2):
            w.timestep(k * dt, dt)
            g.TimeStep(k * dt, dt)
        assert np.all(rel(g.get_state(), w.ranks[0].array("U")) < 1e-12)
    finally:
        g.close()
        w.close()


def test_p_convergence_through_the_gpu_path():
    """convtest/p_3D flavour (ExactFunc 4 with source, exact Dirichlet BCs, LF, mu0=1e-3): the L2 error of the density after a
    short run drops spectrally with N (the reference expects EOC >= 0.8 (N+1) in its h/p convergence checks); norms computed by
    flexi_gpu_calc_error_norms."""
    from flexi_b200.hostmesh.analyze import AnalyzeBasis

    bc = {(a, s): 1 for a in range(3) for s in (-1, 1)}
    errs = []
    for N in (2, 3, 4, 5, 6):
        eqn = dict(parabolic=True, riemann="lf", mu0=1e-3, boundary_type=[(2, 0, 0)], ref_state_prim=[(1.0, 0.3, 0.0, 0.0, 0.71428571)],
                   ini_exact_func=4, do_calc_source=True, adv_vel=(0.3, 0.3, 0.3), split_dg=True, split_dg_type="pi")
        r = build_case(BoxSpec((2, 2, 2), periodic=(False,) * 3, bc_of_face=bc, x0=(-1.0, -1.0, -1.0), length=(2.0, 2.0, 2.0)), N, GL, eqn)[0]
        g = FlexiGpu(r)
        try:
            g.set_state(r.eqn.exact_func4(r.mesh.Elem_xGP, 0.0))
            t, tend = 0.0, 0.05
            while t < tend - 1e-14:
                dt = min(g.CalcTimeStep()[0], tend - t)
                g.TimeStep(t, dt)
                t += dt
            ab = AnalyzeBasis(r.basis)
            l2, _ = g.CalcErrorNorms_sums(t, 4, ab.Vdm, ab.w3, r.mesh.Elem_xGP)
            errs.append(float(np.sqrt(l2[0] / ab.volume(r.mesh.sJ))))
        finally:
            g.close()
    ratios = [errs[i] / errs[i + 1] for i in range(len(errs) - 1)]
    assert all(x > 2.5 for x in ratios), (errs, ratios)
    assert errs[-1] < 1e-5, errs


def test_pinned_host_array_roundtrip():
    """flexi_gpu_pin_host_array (cudaHostRegister of the host's U array): set_state / get_state through a page-locked numpy
    array give back the same bits; double registration is reported, not ignored."""
    import ctypes as C

    from flexi_b200.gpu import FlexiGpuError, lib

    eqn = dict(parabolic=False, riemann="lf")
    r = build_case(BoxSpec((2, 2, 2), **CURVED), 3, GL, eqn)[0]
    U0 = np.ascontiguousarray(r.eqn.smooth_state(r.mesh.Elem_xGP, length=2.0))
    out = np.empty_like(U0)
    L = lib()
    g = FlexiGpu(r)
    try:
        for a in (U0, out):
            assert L.flexi_gpu_pin_host_array(C.c_void_p(a.ctypes.data), a.nbytes) == 0, L.flexi_gpu_last_error()
        assert L.flexi_gpu_pin_host_array(C.c_void_p(U0.ctypes.data), U0.nbytes) != 0  # already registered
        g.set_state(U0)
        g.get_state(out)
        assert np.array_equal(out, U0)
    finally:
        for a in (U0, out):
            L.flexi_gpu_unpin_host_array(C.c_void_p(a.ctypes.data))
        g.close()
