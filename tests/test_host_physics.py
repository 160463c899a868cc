"""The product's point physics (flexi_b200/csrc/physics.cuh, __host__ __device__) compiled as HOST code and compared with the
oracle on the CPU: every two-point volume flux (splitflux.f90) and every Riemann solver (riemann.f90), with and without
SPLIT_DG.  This pins the source the CUDA kernels inline against the restated reference without needing a GPU; the -m gpu
parity tests then check the kernels around it."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from flexi_b200.desc import RIEMANN, SPLITDG
from flexi_b200.hostmesh import basis as B
from flexi_b200.hostmesh.case import build_case, tgv_box
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
PD = C.POINTER(C.c_double)


def _pd(a):
    return a.ctypes.data_as(PD)


@pytest.fixture(scope="module")
def hp(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("hp") / "libhost_physics.so")
    # host code only: -O2 without -march, so gcc contracts no FMAs (same arithmetic as the strict oracle build)
    subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC", "-shared", "-o", out,
                    os.path.join(HERE, "host_physics.cu")], check=True)
    lib = C.CDLL(out)
    lib.hp_split_volume_flux.argtypes = [C.c_int, C.c_double, C.c_double, PD, PD, PD, PD, PD]
    lib.hp_riemann_point.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, PD, PD, PD, PD, PD, PD]
    return lib


def frame(rng):
    n = rng.normal(size=3)
    n /= np.linalg.norm(n)
    a = rng.normal(size=3)
    t1 = a - np.dot(a, n) * n
    t1 /= np.linalg.norm(t1)
    return n, t1, np.cross(n, t1)


def states(eq, rng, close):
    U = eq.random_state((1,), seed=int(rng.integers(1 << 30)))[0]
    if close:  # exercises the series branch of GetLogMean (|f| < 0.1)
        V = U * (1.0 + 0.02 * rng.normal(size=5))
    else:
        V = eq.random_state((1,), seed=int(rng.integers(1 << 30)))[0]
    return np.ascontiguousarray(U), np.ascontiguousarray(V)


@pytest.mark.parametrize("stype", ["sd", "mo", "du", "kg", "pi", "ch", "ir"])
def test_split_volume_flux_matches_oracle(hp, stype):
    ranks = build_case(tgv_box(2), 2, B.NODE_GAUSS_LOBATTO, dict(parabolic=False, split_dg=True, split_dg_type=stype))
    w = O.OracleWorld([(r.desc, r.keep) for r in ranks])
    eq = ranks[0].eqn
    rng = np.random.default_rng(5)
    for it in range(40):
        U, V = states(eq, rng, close=it % 2 == 1)
        M, M2 = rng.normal(size=3), rng.normal(size=3)
        ref = O.split_volume_flux(w.ranks[0], SPLITDG[stype], U, V, M, M2)
        F = np.zeros(5)
        assert hp.hp_split_volume_flux(SPLITDG[stype], eq.kappa, eq.R, _pd(U), _pd(V), _pd(M), _pd(M2), _pd(F)) == 0
        assert np.max(np.abs(F - ref)) <= 4e-15 * np.max(np.abs(ref)), (stype, F, ref)
    w.close()


SPLIT_SOLVERS = ["lf", "roe", "roel2", "roeentropyfix", "average", "ch", "ir"]  # riemann.f90:204-222


@pytest.mark.parametrize("stype", ["sd", "mo", "du", "kg", "pi", "ch", "ir"])
def test_split_riemann_matches_oracle(hp, stype):
    """SPLIT_DG builds: central part = SplitSurfaceFlux of the chosen formulation, all selectable solvers (riemann.f90:206-227)."""
    ranks = build_case(tgv_box(2), 2, B.NODE_GAUSS_LOBATTO, dict(parabolic=False, split_dg=True, split_dg_type=stype))
    w = O.OracleWorld([(r.desc, r.keep) for r in ranks])
    eq = ranks[0].eqn
    rng = np.random.default_rng(9)
    for solver in SPLIT_SOLVERS:
        for it in range(20):
            UL, UR = states(eq, rng, close=it % 2 == 1)
            n, t1, t2 = frame(rng)
            ref = O.riemann_point(w.ranks[0], RIEMANN[solver], UL, UR, n, t1, t2)
            F = np.zeros(5)
            hp.hp_riemann_point(eq.kappa, eq.R, 1, SPLITDG[stype], RIEMANN[solver], _pd(UL), _pd(UR), _pd(n), _pd(t1), _pd(t2), _pd(F))
            assert np.max(np.abs(F - ref)) <= 2e-14 * max(1.0, np.max(np.abs(ref))), (stype, solver, F, ref)
    w.close()


@pytest.mark.parametrize("solver", ["lf", "hllc", "roe", "roel2", "roeentropyfix", "hll", "hlle", "hllem"])
def test_riemann_matches_oracle(hp, solver):
    ranks = build_case(tgv_box(2), 2, B.NODE_GAUSS, dict(parabolic=False))
    w = O.OracleWorld([(r.desc, r.keep) for r in ranks])
    eq = ranks[0].eqn
    rng = np.random.default_rng(13)
    for it in range(30):
        UL, UR = states(eq, rng, close=it % 2 == 1)
        n, t1, t2 = frame(rng)
        ref = O.riemann_point(w.ranks[0], RIEMANN[solver], UL, UR, n, t1, t2)
        F = np.zeros(5)
        hp.hp_riemann_point(eq.kappa, eq.R, 0, 0, RIEMANN[solver], _pd(UL), _pd(UR), _pd(n), _pd(t1), _pd(t2), _pd(F))
        assert np.max(np.abs(F - ref)) <= 2e-14 * max(1.0, np.max(np.abs(ref))), (solver, F, ref)
    w.close()


@pytest.mark.parametrize("nvl", [5, 4])
def test_viscous_flux_matches_oracle(hp, nvl):
    """EvalDiffFlux3D_Point (flux.f90:149-241) of physics.cuh vs the oracle, for both lifting-variable layouts
    (PP_nVarLifting = 5: full prim set, 4: (v1,v2,v3,T); eos.h:110-131)."""
    hp.hp_eval_diff_flux3d.argtypes = [C.c_double] * 4 + [PD] * 7
    ranks = build_case(tgv_box(2), 2, B.NODE_GAUSS, dict(parabolic=True, mu0=1.3e-2, n_var_lifting=nvl))
    w = O.OracleWorld([(r.desc, r.keep) for r in ranks])
    o, eq = w.ranks[0], ranks[0].eqn
    v1, tt = o.lib.oracle_lv_vel1(o.h), o.lib.oracle_lv_temp(o.h)
    rng = np.random.default_rng(31)
    for _ in range(30):
        U = eq.random_state((1,), seed=int(rng.integers(1 << 30)))[0]
        prim = np.ascontiguousarray(eq.cons_to_prim(U))
        G = [rng.normal(size=nvl) for _ in range(3)]  # oracle layout
        ref = [np.zeros(5) for _ in range(3)]
        o.lib.oracle_eval_diff_flux3d(o.h, _pd(prim), *[_pd(g) for g in G], *[_pd(r) for r in ref])
        g4 = [np.ascontiguousarray(np.array([g[v1], g[v1 + 1], g[v1 + 2], g[tt]])) for g in G]
        out = [np.zeros(4) for _ in range(3)]
        hp.hp_eval_diff_flux3d(eq.kappa, eq.R, eq.mu0, eq.Pr, _pd(prim), *[_pd(g) for g in g4], *[_pd(x) for x in out])
        for a, b in zip(out, ref):
            assert b[0] == 0.0
            assert np.max(np.abs(a - b[1:])) <= 4e-15 * max(1.0, np.max(np.abs(b)))
    w.close()


def test_entropy_variables_match_oracle(hp):
    """PrimToEntropy / EntropyToCons (eos.f90:401-420,499-531) of physics.cuh vs the oracle (split form on Gauss nodes)."""
    hp.hp_prim_to_entropy.argtypes = [C.c_double, C.c_double, PD, PD]
    hp.hp_entropy_to_cons.argtypes = [C.c_double, C.c_double, PD, PD]
    ranks = build_case(tgv_box(2), 2, B.NODE_GAUSS, dict(parabolic=False))
    w = O.OracleWorld([(r.desc, r.keep) for r in ranks])
    o, eq = w.ranks[0], ranks[0].eqn
    rng = np.random.default_rng(37)
    for _ in range(30):
        U = eq.random_state((1,), seed=int(rng.integers(1 << 30)))[0]
        prim = np.ascontiguousarray(eq.cons_to_prim(U))
        wa, wb, ca, cb = np.zeros(5), np.zeros(5), np.zeros(5), np.zeros(5)
        o.lib.oracle_prim_to_entropy(o.h, _pd(prim), _pd(wb))
        hp.hp_prim_to_entropy(eq.kappa, eq.R, _pd(prim), _pd(wa))
        assert np.max(np.abs(wa - wb)) <= 4e-15 * np.max(np.abs(wb))
        o.lib.oracle_entropy_to_cons(o.h, _pd(wb), _pd(cb))
        hp.hp_entropy_to_cons(eq.kappa, eq.R, _pd(wb), _pd(ca))
        assert np.max(np.abs(ca - cb)) <= 4e-15 * np.max(np.abs(cb))
        assert np.max(np.abs(cb - U)) <= 1e-13 * np.max(np.abs(U))
    w.close()
