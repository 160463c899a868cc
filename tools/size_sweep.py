"""Parity + discrete conservation of one Ut (GPU vs single-rank oracle) over box sizes: looks for size-dependent defects.
usage: python tools/size_sweep.py 8 16 24 32 40"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flexi_b200 import gpu as G  # noqa: E402
from flexi_b200.hostmesh.case import TGV_EQN, build_case, tgv_box  # noqa: E402
from oracle import oracle as O  # noqa: E402

eqn = dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="lf", cfl_scale=0.8, dfl_scale=0.8, n_var_lifting=5, **TGV_EQN)
for nel in [int(a) for a in sys.argv[1:]] or [8, 16]:
    r = build_case(tgv_box(nel), 7, "gauss-lobatto", eqn)[0]
    U0 = r.eqn.tgv_state(r.mesh.Elem_xGP)
    w3 = np.einsum("k,j,i->kji", r.basis.wGP, r.basis.wGP, r.basis.wGP)[None] / r.mesh.sJ

    def cons(ut):
        m = np.einsum("ekjiv,ekji->v", np.abs(ut), w3)
        m[0:4] = m[1:4].max()
        return np.max(np.abs(np.einsum("ekjiv,ekji->v", ut, w3)) / m)

    with G.FlexiGpu(r) as g:
        g.set_state(U0)
        ut = g.DGTimeDerivative_weakForm(0.0)
        um = g.get_array("U_master")
    t0 = time.perf_counter()
    w = O.OracleWorld([(r.desc, r.keep)], fast=True)
    w.set_state([U0])
    w.dg_time_derivative(0.0)
    ref = w.ranks[0].array("Ut").copy()
    umo = w.ranks[0].array("U_master").copy()
    w.close()
    sc = np.max(np.abs(ref).reshape(-1, 5), axis=0)
    sc[0:4] = sc[1:4].max()  # TGV: div(rho u) = 0, judge density and momenta against the common momentum scale
    d = np.abs(ut - ref).reshape(ut.shape[0], -1, 5).max(axis=1) / sc
    bad = np.nonzero(d.max(axis=1) > 1e-10)[0]
    print(f"nel {nel}: ut_rel {d.max():.3e} cons gpu {cons(ut):.3e} oracle {cons(ref):.3e} U_master diff {np.abs(um - umo).max():.2e} "
          f"bad elems {len(bad)} of {ut.shape[0]} first {bad[:8]} last {bad[-4:]} oracle {time.perf_counter() - t0:.1f}s", flush=True)
