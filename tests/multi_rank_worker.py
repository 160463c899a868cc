"""Worker run under torchrun (world_size >= 2) by tests/test_multi_rank.py.

mode "cpu": gloo only -- checks the host-side plumbing of the partitioned path: neighbour tables are symmetric, slice k of my
            MINE<->nb block pairs with slice k of nb's YOUR<->me block (same master geometry on both ranks), the NCCL-id style
            byte broadcast works, and the oracle world over the same partition is partition invariant.
mode "gpu": one GPU per rank -- flexi_gpu with the NCCL side exchange against the oracle world (all ranks emulated in-process)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from flexi_b200.hostmesh import basis as B  # noqa: E402
from flexi_b200.hostmesh.case import NACA_EQN, build_case, naca_spec, naca_state  # noqa: E402
from flexi_b200.hostmesh.mesh import BoxSpec  # noqa: E402
from oracle import oracle as O  # noqa: E402

CURVED = dict(x0=(-1.0, -1.0, -1.0), length=(2.0, 2.0, 2.0), ngeo=2, deform=0.1)
CASES = {
    "split": (5, B.NODE_GAUSS_LOBATTO, dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="lf", mu0=1e-2)),
    "weakG": (4, B.NODE_GAUSS, dict(parabolic=True, riemann="roe", mu0=1e-2, n_var_lifting=4)),
    "euler": (3, B.NODE_GAUSS_LOBATTO, dict(parabolic=False, riemann="hllc")),
    # BASELINE configs[3] flavour: curved mesh with no-slip / slip walls and a Dirichlet far field, N=5, HLLC + BR2, partitioned
    "config4": (5, B.NODE_GAUSS, dict(parabolic=True, lifting=2, riemann="hllc", mu0=1e-2, eta_br2=2.0, eta_br2_wall=3.0,
                                      boundary_type=[(3, 2, 0), (9, 0, 0), (2, 1, 0)],
                                      ref_state_prim=[(1.0, 0.3, 0.0, 0.0, 0.71428571), (1.0, 0.0, 0.0, 0.0, 0.9)])),
    # the same boundaries under the split-form LGL kernel (viscous-trace protocol: boundary sides keep gradient traces)
    "splitwall": (5, B.NODE_GAUSS_LOBATTO, dict(parabolic=True, lifting=1, riemann="lf", mu0=1e-2, split_dg=True, split_dg_type="pi",
                                                boundary_type=[(4, 2, 0), (9, 0, 0), (2, 1, 0)],
                                                ref_state_prim=[(1.0, 0.3, 0.0, 0.0, 0.71428571), (1.0, 0.0, 0.0, 0.0, 0.9)])),
}
CASES["splitG"] = (4, B.NODE_GAUSS, dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="lf", mu0=1e-2))  # PP_EntropyVars=1
# entropy-conserving CH formulation with its own dissipation on inner, MPI and boundary sides (RiemannBC = same)
CASES["splitCH"] = (5, B.NODE_GAUSS_LOBATTO, dict(parabolic=True, lifting=1, riemann="ch", mu0=1e-2, split_dg=True, split_dg_type="ch",
                                                  boundary_type=[(4, 2, 0), (9, 0, 0), (2, 1, 0)],
                                                  ref_state_prim=[(1.0, 0.3, 0.0, 0.0, 0.71428571), (1.0, 0.0, 0.0, 0.0, 0.9)]))
# BASELINE configs[3] proper: NACA0012-style O-mesh, every element in a random local frame, N=5, HLLC + BR2, wall + far field
CASES["naca"] = (5, B.NODE_GAUSS, NACA_EQN)
# the reference's own NACA0012 mesh file (652 curved elements, flips 0/1/2/4) and its state at t = 10 (tests/test_hopr_mesh.py)
from test_hopr_mesh import FARFIELD as HOPR_BC, MESH as HOPR_MESH, NACA_EQN as HOPR_EQN, STATE as HOPR_STATE  # noqa: E402
from flexi_b200.hostmesh.hopr import HoprMesh, read_state  # noqa: E402
CASES["hopr"] = (3, B.NODE_GAUSS, HOPR_EQN)
# N = 11: every element on a cluster of two CTAs (layout.cuh: Dim::PAIR), inner / boundary element lists and the side exchange
CASES["n11"] = (11, B.NODE_GAUSS, dict(parabolic=True, riemann="roeentropyfix", mu0=1e-2, n_var_lifting=4))
# the solution filter (FilterType = cutoff, dg.f90:329-330) on a partitioned run: the traces of the filtered state are exchanged at the
# start of every operator evaluation
CASES["splitfilt"] = CASES["split"]
CASES["config4_unstr"] = CASES["config4"]
CASES["split_unstr"] = CASES["split"]
WALL_BC = {(1, -1): 1, (1, 1): 2, (0, -1): 3, (0, 1): 3}  # walls in y, far field in x, periodic in z


def main():
    mode, case = sys.argv[1], sys.argv[2]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if mode == "gpu":
        torch.cuda.set_device(local)
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    N, node, eqn = CASES[case]
    oseed = 9 if case.endswith("_unstr") else None  # "_unstr": every element in a random local frame (all flips on MPI sides too)
    if case == "hopr":
        spec = HoprMesh(HOPR_MESH, boundary_type=HOPR_BC)
        eqn = dict(eqn, boundary_type=spec.boundary_type)
        u_file = read_state(HOPR_STATE)[0]
    elif case == "naca":
        spec = naca_spec((12, 3, 2))
    elif "boundary_type" in eqn:
        spec = BoxSpec((4, 3, 2), periodic=(False, False, True), bc_of_face=WALL_BC, orient_seed=oseed, **CURVED)
    else:
        spec = BoxSpec((4, 3, 2), orient_seed=oseed, **CURVED)
    part = dict(partition="linear") if world != 4 else dict(partition="block", blocks=(2, 2, 1))
    gauss_split = bool(eqn.get("split_dg")) and node == B.NODE_GAUSS  # needs Ja_Face / Ja_slave
    ranks = build_case(spec, N, node, eqn, nranks=world, with_ja_face=gauss_split, **part)  # every rank builds all partitions (tiny mesh)
    me = ranks[rank]
    m = me.mesh

    # ---- host plumbing checks (both modes) ----
    tables = [None] * world
    mine = dict(nb=m.NbProc.tolist(), nM=m.nMPISides_MINE_Proc.tolist(), nY=m.nMPISides_YOUR_Proc.tolist(),
                oM=m.offsetMPISides_MINE.tolist(), oY=m.offsetMPISides_YOUR.tolist())
    dist.all_gather_object(tables, mine)
    for i, nb in enumerate(mine["nb"]):
        t = tables[nb]
        assert rank in t["nb"], "asymmetric neighbour tables"
        j = t["nb"].index(rank)
        assert mine["nM"][i] == t["nY"][j] and mine["nY"][i] == t["nM"][j], "message sizes differ"
        # geometry of my MINE slice must be what the neighbour holds for its YOUR slice (master geometry on both ranks)
        mb = ranks[nb].mesh
        a = m.NormVec[mine["oM"][i]:mine["oM"][i] + mine["nM"][i]]
        b = mb.NormVec[t["oY"][j]:t["oY"][j] + t["nY"][j]]
        assert np.array_equal(a, b), "MINE/YOUR slices are not paired side by side"
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        idt = torch.arange(128, dtype=torch.uint8)
    dist.broadcast(idt, src=0)
    assert bytes(idt.numpy().tobytes()) == bytes(range(128))

    # ---- oracle world on this partition vs the single-rank oracle (partition invariance, SURVEY 8c identity 7) ----
    smooth = (lambda r: naca_state(r.eqn, r.mesh.Elem_xGP)) if case == "naca" else (lambda r: r.eqn.smooth_state(r.mesh.Elem_xGP, length=2.0))
    if case == "hopr":
        smooth = lambda r: np.ascontiguousarray(u_file[r.mesh.elem_gid])  # noqa: E731
    U0s = [smooth(r) for r in ranks]
    w = O.OracleWorld([(r.desc, r.keep) for r in ranks])
    if gauss_split:
        for r, o in zip(ranks, w.ranks):
            o.set_face_metrics(r.mesh.Ja_Face, r.mesh.Ja_slave)
    filt = B.overintegration_mat(me.basis.xGP, N - 2) if case == "splitfilt" else None  # the cut-off FilterMat has this form
    if filt is not None:
        for o in w.ranks:
            o.set_filter(1, filt)
    w.set_state(U0s)
    w.dg_time_derivative(0.0)
    ut_ref = w.ranks[rank].array("Ut").copy()
    if rank == 0:
        one = build_case(spec, N, node, eqn, with_ja_face=gauss_split)[0]
        w1 = O.OracleWorld([(one.desc, one.keep)])
        if gauss_split:
            w1.ranks[0].set_face_metrics(one.mesh.Ja_Face, one.mesh.Ja_slave)
        if filt is not None:
            w1.ranks[0].set_filter(1, filt)
        w1.set_state([smooth(one)])
        w1.dg_time_derivative(0.0)
        full = w1.ranks[0].array("Ut")
        for r, ow in zip(ranks, w.ranks):
            # naca: 1/J varies over four decades (wall layer), the flux sums cancel to 1e-4 of their terms -> round-off 5e-11
            ptol = 2e-11 if case in ("naca", "hopr") else 1e-12
            assert np.max(np.abs(ow.array("Ut") - full[r.mesh.elem_gid])) < ptol * np.max(np.abs(full))
        w1.close()

    if mode == "gpu":
        from flexi_b200 import gpu as G

        if rank == 0:
            idt = torch.frombuffer(bytearray(G.nccl_unique_id()), dtype=torch.uint8).clone()
        dist.broadcast(idt, src=0)
        me.remake_desc(nccl_id=bytes(idt.numpy().tobytes()), device=local)
        g = G.FlexiGpu(me)
        if gauss_split:
            g.set_face_metrics(m.Ja_Face, m.Ja_slave)
        if filt is not None:
            g.set_filter(1, filt)
        g.set_state(U0s[rank])
        ut = g.DGTimeDerivative_weakForm(0.0)
        scale = np.max(np.abs(ut_ref).reshape(-1, 5), axis=0)
        # gate: rel 1e-12, or 4x the oracle's own strict-vs-FMA build noise where that is larger (naca: 1/J spans four decades)
        noise = O.ut_noise_floor([(r.desc, r.keep) for r in ranks], U0s)[0] if case in ("naca", "hopr") else 0.0
        assert np.all(np.max(np.abs(ut - ut_ref).reshape(-1, 5), axis=0) <= np.maximum(1e-12 * scale, 4.0 * noise)), "multi-rank Ut mismatch"
        if eqn.get("parabolic"):
            # side arrays after a full DG operator call, incl. what arrives from the neighbour rank: gradU?_slave on MPI_MINE sides
            o = w.ranks[rank]
            gsc = max(np.max(np.abs(o.array(n))) for n in ("gradUx", "gradUy", "gradUz"))
            lo, hi = m.nBCSides, m.firstMPISide_YOUR - 1  # slave arrays of inner + MPI_MINE sides (0-based [lo, hi))
            gtol = 1e-11 if case in ("naca", "hopr") else 1e-12  # naca: 1/J of the wall layer amplifies the round-off of the lifting (see ptol)
            for name in ("gradUx_slave", "gradUy_slave", "gradUz_slave"):
                assert np.max(np.abs(g.get_array(name)[lo:hi] - o.array(name)[lo:hi])) <= gtol * gsc, name
            hi_m = m.lastMPISide_MINE
            for name in ("gradUx_master", "gradUy_master", "gradUz_master"):
                assert np.max(np.abs(g.get_array(name)[:hi_m] - o.array(name)[:hi_m])) <= gtol * gsc, name
        t = 0.0
        for _ in range(3):
            dt_o = w.calc_timestep()[0]
            dt_g = g.CalcTimeStep()[0]
            assert abs(dt_o - dt_g) <= 1e-13 * dt_o, (dt_o, dt_g)
            w.timestep(t, dt_o)
            g.TimeStep(t, dt_o)
            t += dt_o
        Ug, Uo = g.get_state(), w.ranks[rank].array("U")
        sc = np.max(np.abs(Uo).reshape(-1, 5), axis=0)
        stol = 1e-11 if case in ("naca", "hopr") else 1e-12  # naca: round-off of Ut is amplified by 1/J of the wall layer (see ptol above)
        if case == "hopr":
            sc[1:4] = sc[1:4].max()  # two-dimensional flow: w-momentum ~ 1e-10, judged against the common momentum scale
        assert np.all(np.max(np.abs(Ug - Uo).reshape(-1, 5), axis=0) <= stol * sc), "multi-rank state mismatch after 3 steps"
        if eqn.get("parabolic") and "boundary_type" not in eqn and N <= 9:  # (the on-device analysis holds an element in one CTA: N <= 9)
            # on-device TGV AnalyzeTestcase (steps 1-4) per rank + the host's reduction (steps 5-6) against the host analysis of
            # the oracle world: the integrals of the partitioned run
            from flexi_b200.hostmesh.analyze import AnalyzeBasis, tgv_analyze

            ab = AnalyzeBasis(me.basis)
            vol = sum(ab.volume(r.mesh.sJ) for r in ranks)
            g.keep_gradients(True)
            g.DGTimeDerivative_weakForm(t, want_ut=False)
            w.dg_time_derivative(t)
            parts = [None] * world
            dist.all_gather_object(parts, {k: float(v) for k, v in g.AnalyzeTestcase_sums(ab.Vdm, ab.w3).items()})
            dev = G.FlexiGpu.tgv_normalise(parts, vol, me.eqn.mu0)
            host = [tgv_analyze(ab, r.eqn, r.mesh, o.array("U"), o.array("gradUx"), o.array("gradUy"), o.array("gradUz"), vol)
                    for r, o in zip(ranks, w.ranks)]
            for k in host[0]:
                if k == "uPrime":
                    continue
                ref = max(h[k] for h in host) if k == "Vorticity_max" else sum(h[k] for h in host)
                tol = 1e-8 if k in ("DR_p", "ED_D") else 1e-11
                assert abs(dev[k] - ref) <= tol * max(abs(ref), 1e-30) + 1e-14, (k, dev[k], ref)
        g.close()
    w.close()
    dist.barrier()
    if rank == 0:
        print(f"MULTI_RANK_OK mode={mode} case={case} world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
