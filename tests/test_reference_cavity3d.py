"""The reference's own end-to-end golden vector for the Navier-Stokes path: regressioncheck/checks/parabolic/cavity_3D.

The reference's regression suite runs the lid-driven cavity (4^3 elements, N = 2, Gauss nodes, weak form, BR1 with all five lifting
variables, the default Riemann solver RoeEntropyFix, isothermal walls BC type 4 + a Dirichlet lid BC type 2, CarpenterRK4-5,
CFLscale 0.99 / DFLscale 0.4, time step recomputed every step) from rest to t = 1 and compares `DG_Solution` of the final state with
`reggie_cavity_Re100_State_0000001.000000000.h5` at an ABSOLUTE tolerance of 1e-12 (analyze.ini: h5diff).  Mesh and reference
state are copied byte-identically into tests/golden/reference_cavity3d/ and read with flexi_b200/hostmesh/hdf5_min.py.

Replaying the same run (353 time steps = 1765 operator evaluations, the end-of-run step logic of UpdateTimeStep,
timedisc_func.f90:236-266) pins, against numbers the reference produced: the mesh reader and metrics, ConsToPrim, ProlongToFace
and SurfInt on Gauss nodes, BR1 lifting incl. its boundary fluxes, the viscous fluxes, Riemann_Side + RoeEntropyFix,
GetBoundaryFlux types 2 and 4, the weak-form VolInt, CalcTimeStep and the LSERK update -- for the oracle (CPU suite) and for
the CUDA path (GPU suite)."""
import os

import numpy as np
import pytest

from flexi_b200.hostmesh import basis as B
from flexi_b200.hostmesh.case import build_case
from flexi_b200.hostmesh.hopr import HoprMesh, read_state
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_cavity3d")
MESH = os.path.join(GOLD, "cavity4x4x4_mesh.h5")
STATE = os.path.join(GOLD, "reggie_cavity_Re100_State_0000001.000000000.h5")
REF1, REF2 = (1.0, 1.0, 0.0, 0.0, 71.4285714286), (1.0, 0.0, 0.0, 0.0, 71.4285714286)
T_END = 1.0


def cavity_rank(nranks=1):
    # parameter.ini: BoundaryName / BoundaryType overrides of the mesh file's BCs
    hm = HoprMesh(MESH, boundary_type={"BC_wall_left": (4, 1), "BC_wall_right": (4, 1), "BC_free": (2, 1)})
    eqn = dict(parabolic=True, riemann="roeentropyfix", kappa=1.4, R=1.0, Pr=0.72, mu0=0.01, cfl_scale=0.99, dfl_scale=0.4,
               ref_state_prim=[REF1, REF2], boundary_type=hm.boundary_type, n_var_lifting=5)
    return build_case(hm, 2, B.NODE_GAUSS, eqn, nranks=nranks)


def run_to_t_end(calc_timestep, timestep, t_end=T_END):
    """The time loop of TimeDisc (timedisc.f90:191-212) with UpdateTimeStep (timedisc_func.f90:236-266); tAnalyze = tEnd here."""
    t, n = 0.0, 0
    while True:
        dt_min = (calc_timestep(), t_end - t, t_end - t)       # DT_MIN, DT_ANALYZE, DT_END
        dt = min(dt_min)
        finalize = dt == dt_min[2]
        dt = min(x for x in dt_min if x > 0.0)
        if dt_min[2] - dt < dt / 100.0 and dt_min[2] > 0.0:    # the LAST step would be smaller than dt/100: stretch this one
            dt, finalize = dt_min[2], True
        timestep(t, dt)
        t += dt
        n += 1
        if finalize:
            return t, n


def check_against_reference(U, tol):
    ref, attrs, _ = read_state(STATE)
    assert float(attrs["Time"][0]) == T_END and int(attrs["N"][0]) == 2 and attrs["NodeType"][0].strip() == b"GAUSS"
    d = np.max(np.abs(U - ref).reshape(-1, 5), axis=0)
    assert np.all(d <= tol), d
    return d


def initial_state(r):
    # IniExactFunc = 1, IniRefState = 2: the fluid at rest
    return np.broadcast_to(r.eqn.prim_to_cons(np.array(REF2)), r.mesh.Elem_xGP.shape[:-1] + (5,)).copy()


@pytest.mark.parametrize("nranks", [1, 2])
def test_oracle_reproduces_the_reference_state(nranks):
    """The reference checks its own reruns (MPI = 1, 2) at 1e-12 absolute; the strict-IEEE oracle lands within 2e-13 of it
    (energy ~ 180: 1e-15 relative)."""
    ranks = cavity_rank(nranks)
    w = O.OracleWorld([(r.desc, r.keep) for r in ranks])
    w.set_state([initial_state(r) for r in ranks])

    def calc():
        dt, _, _, err = w.calc_timestep()
        assert err == 0
        return dt

    t, n = run_to_t_end(calc, w.timestep)
    assert abs(t - T_END) < 1e-14 and n == 353
    U = np.zeros((64, 3, 3, 3, 5))
    for r, o in zip(ranks, w.ranks):
        U[r.mesh.elem_gid] = o.array("U")
    w.close()
    check_against_reference(U, 1e-12)


@pytest.mark.gpu
def test_gpu_reproduces_the_reference_state():
    """The same run through the C ABI on a B200 (CalcTimeStep + TimeStep per step): the reference's own acceptance tolerance."""
    from flexi_b200.gpu import FlexiGpu

    r = cavity_rank()[0]
    with FlexiGpu(r) as g:
        g.set_state(initial_state(r))

        def calc():
            dt, _, _, err = g.CalcTimeStep()
            assert err == 0
            return dt

        t, n = run_to_t_end(calc, g.TimeStep)
        assert abs(t - T_END) < 1e-14 and n == 353
        check_against_reference(g.get_state(), 1e-12)
