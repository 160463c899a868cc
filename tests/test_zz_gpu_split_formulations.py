"""The remaining SplitDG formulations (MO, CH, IR: splitflux.f90:471-593,711-989) and the SPLIT_DG-only Riemann solvers
(Average, CH, IR: riemann.f90:1236-1402) on the CUDA path against the oracle: Gauss-Lobatto kernels (`k_volume_gl` for even N,
`k_volume_gl2` for odd N) and the Gauss-node split form.  The point functions themselves are pinned on the CPU
(tests/test_host_physics.py compiles physics.cuh as host code); this file checks the kernels around them.

Sorted last on purpose: written after the round's GPU budget was spent, so it has not run on a B200 yet."""
import numpy as np
import pytest

from flexi_b200.gpu import FlexiGpu
from flexi_b200.hostmesh import basis as B
from flexi_b200.hostmesh.case import build_case, synthetic_sponge
from flexi_b200.hostmesh.mesh import BoxSpec
from oracle import oracle as O

pytestmark = pytest.mark.gpu
G, GL = B.NODE_GAUSS, B.NODE_GAUSS_LOBATTO
CURVED = dict(x0=(-1.0, -1.0, -1.0), length=(2.0, 2.0, 2.0), ngeo=2, deform=0.1)

CASES = [
    (4, GL, dict(parabolic=False, split_dg=True, split_dg_type="ch", riemann="ch")),
    (7, GL, dict(parabolic=True, split_dg=True, split_dg_type="ch", riemann="ch", mu0=1e-2)),
    (3, GL, dict(parabolic=False, split_dg=True, split_dg_type="ir", riemann="ir")),
    (6, GL, dict(parabolic=True, split_dg=True, split_dg_type="ir", riemann="ir", mu0=1e-2, n_var_lifting=4)),
    (5, GL, dict(parabolic=True, split_dg=True, split_dg_type="mo", riemann="lf", mu0=1e-2)),
    (2, GL, dict(parabolic=False, split_dg=True, split_dg_type="mo", riemann="roe")),
    (7, GL, dict(parabolic=False, split_dg=True, split_dg_type="pi", riemann="average")),
    (7, GL, dict(parabolic=True, split_dg=True, split_dg_type="kg", riemann="ch", mu0=1e-2)),
    (3, G, dict(parabolic=False, split_dg=True, split_dg_type="ch", riemann="ch")),
    (4, G, dict(parabolic=True, split_dg=True, split_dg_type="ir", riemann="ir", mu0=1e-2)),
    (5, G, dict(parabolic=True, split_dg=True, split_dg_type="mo", riemann="average", mu0=1e-2)),
]


def _id(c):
    N, node, e = c
    return f"N{N}-{'G' if node == G else 'GL'}-{e['split_dg_type']}-{e['riemann']}-{'ns' if e['parabolic'] else 'euler'}"


@pytest.mark.parametrize("N,node,eqn", CASES, ids=[_id(c) for c in CASES])
def test_split_formulation_matches_oracle(N, node, eqn):
    """One DG operator (rel 1e-12), then CalcTimeStep + 2 LSERK steps (state rel 1e-12) on the sine-deformed periodic box."""
    gauss = node == G
    r = build_case(BoxSpec((3, 2, 2), **CURVED), N, node, eqn, with_ja_face=gauss)[0]
    U0 = r.eqn.smooth_state(r.mesh.Elem_xGP, length=2.0)
    w = O.OracleWorld([(r.desc, r.keep)])
    if gauss:
        w.ranks[0].set_face_metrics(r.mesh.Ja_Face, r.mesh.Ja_slave)
    w.set_state([U0])
    g = FlexiGpu(r)
    try:
        if gauss:
            g.set_face_metrics(r.mesh.Ja_Face, r.mesh.Ja_slave)
        g.set_state(U0)
        w.dg_time_derivative(0.0)
        ut = g.DGTimeDerivative_weakForm(0.0)
        o = w.ranks[0]
        ref = o.array("Ut")
        a, b = g.get_array("Flux"), o.array("Flux")
        assert np.max(np.abs(a - b)) <= 1e-12 * np.max(np.abs(b)), "Flux"
        d = np.max(np.abs(ut - ref).reshape(-1, 5), axis=0)
        scale = np.max(np.abs(ref).reshape(-1, 5), axis=0)
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


@pytest.mark.parametrize("N,stype,riemann,lifting,wall", [(5, "ch", "ch", 1, 3), (7, "ch", "ir", 1, 3), (5, "ch", "ir", 2, 4), (4, "mo", "average", 1, 3)])
def test_split_formulations_with_walls_and_farfield(N, stype, riemann, lifting, wall):
    """RiemannBC = same: the SPLIT_DG-only solvers also serve the Dirichlet far-field sides (getboundaryflux.f90:561-578)."""
    from test_gpu_bc_br2 import REF, check, wall_farfield_box
    eqn = dict(parabolic=True, lifting=lifting, riemann=riemann, mu0=1e-2, eta_br2=2.0, eta_br2_wall=3.0, split_dg=True,
               split_dg_type=stype, boundary_type=[(wall, 2, 0), (9, 0, 0), (2, 1, 0)], ref_state_prim=[REF, (1.0, 0.0, 0.0, 0.0, 0.9)])
    ranks = build_case(wall_farfield_box(), N, GL, eqn)
    r = ranks[0]
    check(ranks, r.eqn.smooth_state(r.mesh.Elem_xGP, length=2.0))


@pytest.mark.parametrize("N,NU,node,eqn", [
    (7, 5, GL, dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="lf", mu0=1e-2)),
    (5, 3, G, dict(parabolic=True, riemann="roe", mu0=1e-2)),
    (4, 2, GL, dict(parabolic=False, riemann="hllc")),
    (4, 3, G, dict(parabolic=False, split_dg=True, split_dg_type="kg", riemann="lf")),
], ids=["N7-GL-split", "N5-G-weak", "N4-GL-euler", "N4-G-split"])
def test_overintegration_cutoff_matches_oracle(N, NU, node, eqn):
    """OverintegrationType = cutoff (flexi_gpu_set_overintegration; dg.f90:663-676): Ut filtered before ApplyJacobian inside the
    update kernel, one operator evaluation and 2 LSERK steps on the curved mesh; switching it off restores the plain operator."""
    gauss_split = node == G and eqn.get("split_dg")
    r = build_case(BoxSpec((3, 2, 2), **CURVED), N, node, eqn, with_ja_face=bool(gauss_split))[0]
    U0 = r.eqn.smooth_state(r.mesh.Elem_xGP, length=2.0)
    mat = B.overintegration_mat(r.basis.xGP, NU)
    w = O.OracleWorld([(r.desc, r.keep)])
    o = w.ranks[0]
    if gauss_split:
        o.set_face_metrics(r.mesh.Ja_Face, r.mesh.Ja_slave)
    w.set_state([U0])
    g = FlexiGpu(r)
    try:
        if gauss_split:
            g.set_face_metrics(r.mesh.Ja_Face, r.mesh.Ja_slave)
        g.set_state(U0)
        w.dg_time_derivative(0.0)
        plain = np.array(o.array("Ut"))
        o.set_overintegration(1, mat)
        g.set_overintegration(1, mat)
        w.dg_time_derivative(0.0)
        ut, ref = g.DGTimeDerivative_weakForm(0.0), o.array("Ut")
        scale = np.max(np.abs(ref).reshape(-1, 5), axis=0)
        assert np.all(np.max(np.abs(ut - ref).reshape(-1, 5), axis=0) <= 1e-12 * scale)
        assert np.max(np.abs(ref - plain)) > 1e-6 * np.max(np.abs(plain))  # the filter does act on this state
        t = 0.0
        for _ in range(2):
            dt = w.calc_timestep()[0]
            w.timestep(t, dt)
            g.TimeStep(t, dt)
            t += dt
        Ug, Uo = g.get_state(), o.array("U")
        sc = np.max(np.abs(Uo).reshape(-1, 5), axis=0)
        assert np.all(np.max(np.abs(Ug - Uo).reshape(-1, 5), axis=0) <= 1e-12 * sc)
        g.set_overintegration(0)
        g.set_state(U0)
        ut0 = g.DGTimeDerivative_weakForm(0.0)
        assert np.all(np.max(np.abs(ut0 - plain).reshape(-1, 5), axis=0) <= 1e-12 * np.max(np.abs(plain).reshape(-1, 5), axis=0))
        with pytest.raises(Exception, match="OverintegrationType 2"):
            g.set_overintegration(2, mat)
    finally:
        g.close()
        w.close()


@pytest.mark.parametrize("N,NF,node,eqn", [
    (7, 4, GL, dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="lf", mu0=1e-2)),
    (4, 2, G, dict(parabolic=True, riemann="roe", mu0=1e-2)),
], ids=["N7-GL-split", "N4-G-weak"])
def test_solution_filter_matches_oracle(N, NF, node, eqn):
    """FilterType cutoff (flexi_gpu_set_filter; dg.f90:329-330): U filtered in place at the start of every operator evaluation
    (stand-alone k_filter pass + trace refresh); one operator, then 2 LSERK steps."""
    r = build_case(BoxSpec((3, 2, 2), **CURVED), N, node, eqn)[0]
    U0 = r.eqn.smooth_state(r.mesh.Elem_xGP, length=2.0)
    mat = B.overintegration_mat(r.basis.xGP, NF)
    w = O.OracleWorld([(r.desc, r.keep)])
    o = w.ranks[0]
    o.set_filter(1, mat)
    w.set_state([U0])
    g = FlexiGpu(r)
    try:
        g.set_filter(1, mat)
        g.set_state(U0)
        w.dg_time_derivative(0.0)
        ut, ref = g.DGTimeDerivative_weakForm(0.0), o.array("Ut")
        scale = np.max(np.abs(ref).reshape(-1, 5), axis=0)
        assert np.all(np.max(np.abs(ut - ref).reshape(-1, 5), axis=0) <= 1e-12 * scale)
        Ug, Uo = g.get_state(), o.array("U")
        assert np.max(np.abs(Ug - Uo)) <= 1e-13 * np.max(np.abs(Uo))  # the state itself is filtered
        assert np.max(np.abs(Uo - U0)) > 1e-6
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
        with pytest.raises(Exception, match="FilterType 3"):
            g.set_filter(3, mat)
    finally:
        g.close()
        w.close()


@pytest.mark.parametrize("N,node,eqn", [
    (7, GL, dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="lf", mu0=1e-2)),
    (4, G, dict(parabolic=True, riemann="roe", mu0=1e-2)),
    (5, GL, dict(parabolic=False, riemann="hllc")),
], ids=["N7-GL-split", "N4-G-weak", "N5-GL-euler"])
def test_sponge_matches_oracle(N, node, eqn):
    """Sponge(Ut) (flexi_gpu_set_sponge; sponge.f90:612-678, dg.f90:659-661) inside the update kernel: one operator evaluation (the
    kernel has to fetch U itself there) and 2 LSERK steps on the curved mesh with a sponge zone over the last 40 % in x; also
    together with the cut-off overintegration (the sponge term is filtered with Ut); switching it off restores the plain operator."""
    r = build_case(BoxSpec((4, 2, 2), **CURVED), N, node, eqn)[0]
    U0 = r.eqn.smooth_state(r.mesh.Elem_xGP, length=2.0)
    smap, smat, base = synthetic_sponge(r)
    w = O.OracleWorld([(r.desc, r.keep)])
    o = w.ranks[0]
    w.set_state([U0])
    g = FlexiGpu(r)
    try:
        g.set_state(U0)
        w.dg_time_derivative(0.0)
        plain = np.array(o.array("Ut"))
        o.set_sponge(smap, smat, base)
        g.set_sponge(smap, smat, base)
        w.dg_time_derivative(0.0)
        ut, ref = g.DGTimeDerivative_weakForm(0.0), o.array("Ut")
        scale = np.max(np.abs(ref).reshape(-1, 5), axis=0)
        assert np.all(np.max(np.abs(ut - ref).reshape(-1, 5), axis=0) <= 1e-12 * scale)
        assert np.max(np.abs(ref - plain)) > 1e-3 * np.max(np.abs(plain))
        mat = B.overintegration_mat(r.basis.xGP, N - 2)
        o.set_overintegration(1, mat)
        g.set_overintegration(1, mat)
        w.dg_time_derivative(0.0)
        ut, ref = g.DGTimeDerivative_weakForm(0.0), o.array("Ut")
        assert np.all(np.max(np.abs(ut - ref).reshape(-1, 5), axis=0) <= 1e-12 * np.max(np.abs(ref).reshape(-1, 5), axis=0))
        o.set_overintegration(0)
        g.set_overintegration(0)
        t = 0.0
        for _ in range(2):
            dt = w.calc_timestep()[0]
            w.timestep(t, dt)
            g.TimeStep(t, dt)
            t += dt
        Ug, Uo = g.get_state(), o.array("U")
        sc = np.max(np.abs(Uo).reshape(-1, 5), axis=0)
        assert np.all(np.max(np.abs(Ug - Uo).reshape(-1, 5), axis=0) <= 1e-12 * sc)
        g.set_sponge()
        g.set_state(U0)
        ut0 = g.DGTimeDerivative_weakForm(0.0)
        assert np.all(np.max(np.abs(ut0 - plain).reshape(-1, 5), axis=0) <= 1e-12 * np.max(np.abs(plain).reshape(-1, 5), axis=0))
    finally:
        g.close()
        w.close()


WALLS = dict(boundary_type=[(4, 2, 0), (9, 0, 0), (2, 1, 0)], ref_state_prim=[(1.0, 0.3, 0.0, 0.0, 0.71428571), (1.0, 0.0, 0.0, 0.0, 0.9)])


@pytest.mark.parametrize("visc", [1, 2], ids=["sutherland", "powerlaw"])
@pytest.mark.parametrize("N,node,eqn,walls", [
    (7, GL, dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="lf", mu0=1e-2), False),
    (5, GL, dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="lf", mu0=1e-2, lifting=1), True),
    (4, G, dict(parabolic=True, riemann="roe", mu0=1e-2, lifting=2), True),
    (3, GL, dict(parabolic=True, riemann="hllc", mu0=1e-2), False),
], ids=["N7-GL-split", "N5-GL-split-walls", "N4-G-weak-BR2-walls", "N3-GL-weak"])
def test_viscosity_laws_match_oracle(N, node, eqn, walls, visc):
    """PP_VISC = 1 / 2 (flexi_gpu_set_viscosity; viscosity.f90:57-90, eos.h:73-87) in EvalDiffFlux3D of the element kernels, the side
    kernel, the boundary fluxes, and in CalcTimeStep: one operator evaluation (all side arrays), dt, 2 LSERK steps on the curved mesh.
    The temperature of the test state straddles Ts, so both branches of Sutherland's formula are taken."""
    from flexi_b200.hostmesh.equation import sutherland_params

    wall_bc = {(1, -1): 1, (1, 1): 2, (0, -1): 3, (0, 1): 3}
    spec = BoxSpec((3, 2, 2), periodic=(False, False, True), bc_of_face=wall_bc, **CURVED) if walls else BoxSpec((3, 2, 2), **CURVED)
    r = build_case(spec, N, node, dict(eqn, **(WALLS if walls else {})))[0]
    U0 = r.eqn.smooth_state(r.mesh.Elem_xGP, length=2.0)
    T = r.eqn.cons_to_prim(U0)[..., 5]
    law = sutherland_params(Ts_ini=float(np.median(T)), Tref_ini=float(np.mean(T)), expo=1.5) if visc == 1 else dict(ExpoSuth=0.7)
    w = O.OracleWorld([(r.desc, r.keep)])
    o = w.ranks[0]
    w.set_state([U0])
    g = FlexiGpu(r)
    try:
        g.set_state(U0)
        w.dg_time_derivative(0.0)
        plain = np.array(o.array("Ut"))
        o.set_viscosity(visc, **law)
        g.set_viscosity(visc, **law)
        w.dg_time_derivative(0.0)
        ut, ref = g.DGTimeDerivative_weakForm(0.0), o.array("Ut")
        scale = np.max(np.abs(ref).reshape(-1, 5), axis=0)
        assert np.all(np.max(np.abs(ut - ref).reshape(-1, 5), axis=0) <= 1e-12 * scale)
        assert np.max(np.abs(ref - plain)) > 1e-4 * np.max(np.abs(plain))  # the law does change the operator
        assert np.max(np.abs(g.get_array("Flux") - o.array("Flux"))) <= 1e-12 * np.max(np.abs(o.array("Flux")))
        t = 0.0
        for _ in range(2):
            dt = w.calc_timestep()
            dtg = g.CalcTimeStep()
            assert abs(dtg[0] - dt[0]) <= 1e-13 * dt[0] and abs(dtg[2] - dt[2]) <= 1e-13 * dt[2]
            w.timestep(t, dt[0])
            g.TimeStep(t, dt[0])
            t += dt[0]
        Ug, Uo = g.get_state(), o.array("U")
        sc = np.max(np.abs(Uo).reshape(-1, 5), axis=0)
        assert np.all(np.max(np.abs(Ug - Uo).reshape(-1, 5), axis=0) <= 1e-12 * sc)
    finally:
        g.close()
        w.close()


@pytest.mark.parametrize("N,NU,node,eqn", [
    (7, 4, GL, dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="lf", mu0=1e-2)),
    (5, 3, G, dict(parabolic=True, riemann="roe", mu0=1e-2)),
    (4, 2, GL, dict(parabolic=False, riemann="hllc")),
], ids=["N7-GL-split", "N5-G-weak", "N4-GL-euler"])
def test_overintegration_conservative_cutoff_matches_oracle(N, NU, node, eqn):
    """OverintegrationType = 2 (flexi_gpu_set_overintegration_cons; overintegration.f90:181-330, dg.f90:664-668): projection to NUnder,
    Jacobian on NUnder, interpolation back inside the update kernel, no ApplyJacobian: one operator evaluation and 2 LSERK steps on the
    curved mesh; the plain cut-off and 'off' still work afterwards."""
    r = build_case(BoxSpec((3, 2, 2), **CURVED), N, node, eqn)[0]
    U0 = r.eqn.smooth_state(r.mesh.Elem_xGP, length=2.0)
    vd, vu, sjnu = B.overintegration_cons(N, node, NU, r.mesh.sJ)
    w = O.OracleWorld([(r.desc, r.keep)])
    o = w.ranks[0]
    w.set_state([U0])
    g = FlexiGpu(r)
    try:
        g.set_state(U0)
        w.dg_time_derivative(0.0)
        plain = np.array(o.array("Ut"))
        o.set_overintegration_cons(NU, vd, vu, sjnu)
        g.set_overintegration_cons(NU, vd, vu, sjnu)
        w.dg_time_derivative(0.0)
        ut, ref = g.DGTimeDerivative_weakForm(0.0), o.array("Ut")
        scale = np.max(np.abs(ref).reshape(-1, 5), axis=0)
        assert np.all(np.max(np.abs(ut - ref).reshape(-1, 5), axis=0) <= 1e-12 * scale)
        assert np.max(np.abs(ref - plain)) > 1e-6 * np.max(np.abs(plain))
        t = 0.0
        for _ in range(2):
            dt = w.calc_timestep()[0]
            w.timestep(t, dt)
            g.TimeStep(t, dt)
            t += dt
        Ug, Uo = g.get_state(), o.array("U")
        sc = np.max(np.abs(Uo).reshape(-1, 5), axis=0)
        assert np.all(np.max(np.abs(Ug - Uo).reshape(-1, 5), axis=0) <= 1e-12 * sc)
        # back to the plain cut-off, then off
        mat = B.overintegration_mat(r.basis.xGP, NU)
        o.set_overintegration_cons(-1)
        o.set_overintegration(1, mat)
        g.set_overintegration(1, mat)
        g.set_state(U0)
        w.set_state([U0])
        w.dg_time_derivative(0.0)
        ut, ref = g.DGTimeDerivative_weakForm(0.0), o.array("Ut")
        assert np.all(np.max(np.abs(ut - ref).reshape(-1, 5), axis=0) <= 1e-12 * np.max(np.abs(ref).reshape(-1, 5), axis=0))
        g.set_overintegration(0)
        ut0 = g.DGTimeDerivative_weakForm(0.0)
        assert np.all(np.max(np.abs(ut0 - plain).reshape(-1, 5), axis=0) <= 1e-12 * np.max(np.abs(plain).reshape(-1, 5), axis=0))
    finally:
        g.close()
        w.close()


@pytest.mark.parametrize("N,NF,node,eqn", [
    (7, 3, GL, dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="lf", mu0=1e-2)),
    (5, 2, G, dict(parabolic=True, riemann="roe", mu0=1e-2)),
], ids=["N7-GL-split", "N5-G-weak"])
def test_laf_filter_matches_oracle(N, NF, node, eqn):
    """FilterType 3, LAF (flexi_gpu_set_filter_laf; filter.f90:295-448): per element the cut-off filtered state replaces U only where
    the energy ratio exceeds the adaptive limit.  Two operators (the first initialises the limits, before the second some elements get new fluctuations: both branches
    occur), then 2 LSERK steps (5 filter calls each; the limits adapt from call to call); eRatio / lim / r and the state against the oracle."""
    r = build_case(BoxSpec((3, 2, 2), **CURVED), N, node, eqn)[0]
    U0 = r.eqn.smooth_state(r.mesh.Elem_xGP, length=2.0)
    rng = np.random.default_rng(5)
    U0[::2, ..., 1:4] += 0.05 * rng.normal(size=U0[::2, ..., 1:4].shape)
    mat = B.overintegration_mat(r.basis.xGP, NF)
    iw, norm_mod = B.laf_params(N, NF, r.basis.wGP, r.mesh.sJ)
    w = O.OracleWorld([(r.desc, r.keep)])
    o = w.ranks[0]
    o.set_filter(1, mat)
    o.set_filter_laf(iw, norm_mod, 1.0)
    w.set_state([U0])
    g = FlexiGpu(r)
    try:
        g.set_filter_laf(mat, iw, norm_mod, 1.0)
        g.set_state(U0)
        w.dg_time_derivative(0.0)
        ut, ref = g.DGTimeDerivative_weakForm(0.0), o.array("Ut")
        scale = np.max(np.abs(ref).reshape(-1, 5), axis=0)
        assert np.all(np.max(np.abs(ut - ref).reshape(-1, 5), axis=0) <= 1e-12 * scale)
        Ug, Uo = g.get_state(), o.array("U")
        assert np.max(np.abs(Ug - Uo)) <= 1e-13 * np.max(np.abs(Uo))
        (ea, la, ra), (eb, lb, rb) = g.get_filter_laf(), o.get_filter_laf()
        assert np.max(np.abs(ea - eb)) <= 1e-11 * np.max(np.abs(eb))
        assert np.max(np.abs(la - lb)) <= 1e-11 * np.max(np.abs(lb))
        assert np.array_equal(ra, rb)
        # the first call only initialises the limits (old energies 1e-16); new fluctuations in every third element, second call
        V = np.array(Uo)
        V[1::3, ..., 1:4] += 0.02 * rng.normal(size=V[1::3, ..., 1:4].shape)
        w.set_state([V])
        g.set_state(V)
        w.dg_time_derivative(0.0)
        ut, ref = g.DGTimeDerivative_weakForm(0.0), o.array("Ut")
        scale = np.max(np.abs(ref).reshape(-1, 5), axis=0)
        assert np.all(np.max(np.abs(ut - ref).reshape(-1, 5), axis=0) <= 1e-12 * scale)
        (ea, la, ra), (eb, lb, rb) = g.get_filter_laf(), o.get_filter_laf()
        assert np.array_equal(ra, rb)
        assert 0 < rb.sum() < len(rb), "both branches of the LAF decision have to occur"
        assert np.max(np.abs(la - lb)) <= 1e-11 * np.max(np.abs(lb))
        Ug, Uo = g.get_state(), o.array("U")
        assert np.max(np.abs(Ug - Uo)) <= 1e-13 * np.max(np.abs(Uo))
        t = 0.0
        for _ in range(2):
            dt = w.calc_timestep()[0]
            assert abs(g.CalcTimeStep()[0] - dt) <= 1e-13 * dt
            w.timestep(t, dt)
            g.TimeStep(t, dt)
            t += dt
        (ea, la, ra), (eb, lb, rb) = g.get_filter_laf(), o.get_filter_laf()
        assert np.array_equal(ra, rb)
        assert np.max(np.abs(ea - eb)) <= 1e-9 * np.max(np.abs(eb))
        Ug, Uo = g.get_state(), o.array("U")
        sc = np.max(np.abs(Uo).reshape(-1, 5), axis=0)
        assert np.all(np.max(np.abs(Ug - Uo).reshape(-1, 5), axis=0) <= 1e-12 * sc)
        g.set_filter(0)  # off again: the plain operator
        o.set_filter_laf(None)
        o.set_filter(0)
        w.dg_time_derivative(t)
        ut, ref = g.DGTimeDerivative_weakForm(t), o.array("Ut")
        scale = np.max(np.abs(ref).reshape(-1, 5), axis=0)
        assert np.all(np.max(np.abs(ut - ref).reshape(-1, 5), axis=0) <= 1e-12 * scale)
    finally:
        g.close()
        w.close()
