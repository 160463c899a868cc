#!/usr/bin/env python
"""bench.py -- PID / GDOF/s of the DGSEM RHS + LSERK hot path (Navier-Stokes TGV, N=7) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (N>1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port) on the host cores

A "step" is one LSERK time step = nRKStages (5) evaluations of DGTimeDerivative_weakForm with the fused stage update on a
synthetic periodic TGV box (BASELINE.json configs[1]: N=7, 48^3 elements per GPU, split-form Gauss-Lobatto (Pirozzoli) with
Lax-Friedrichs, BR1).  value = nDOF * nRKStages * steps / time / 1e9 [GDOF/s] (whole job); PID = 1 / (value * 1e9).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# algorithmic HBM bytes per DOF per RK stage (FP64); split per kernel class in DESIGN.md
B_ALG_STAGE = {True: lambda N: 240.0 + 3.0 / (N + 1) * 720.0, False: lambda N: 240.0 + 3.0 / (N + 1) * 240.0}
BLOCKS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def b_alg_kernel(N, parabolic):
    """Share of the stage's algorithmic bytes per kernel class (DESIGN.md 'Algorithmic bytes'), per DOF per stage."""
    f = 3.0 / (N + 1)  # unique face nodes per volume node
    vol = 8.0 * 15 + (f * 8.0 * 30 if parabolic else 0.0)        # U read + geometry read (+ gradient traces written)
    flux = f * 8.0 * (10 + 10 + (30 if parabolic else 0))         # U traces read + face geometry (+ gradient traces read)
    upd = 8.0 * 15 + f * 8.0 * 10                                  # U write + RK register read/write + U traces written
    return dict(volume=vol, flux=flux, update=upd)


def naca_nel(nel, world):
    """configs[3] mesh: nel elements around the section, nel/3 radial, nel/6 per GPU along the span (z-slabs of the linear partition)."""
    return (nel, max(2, nel // 3), max(2, nel // 6) * world)


def workload_case(args, world, nel, small=False):
    """(spec, partition kwargs, initial-state function of a HostRank) of the benchmarked workload on `world` ranks with `nel`
    elements per direction per GPU."""
    from flexi_b200.hostmesh.case import naca_spec, naca_state, tgv_box

    if args.workload == "naca":
        spec = naca_spec(naca_nel(nel, world))
        return spec, dict(partition="linear"), (lambda r: naca_state(r.eqn, r.mesh.Elem_xGP))
    blocks = BLOCKS[world]
    spec = tgv_box(tuple(nel * b for b in blocks))
    if small:  # parity gate: element size of the benchmarked mesh, TGV state with the wavelength of the (smaller) box
        spec.length = np.array([2 * np.pi * b * nel / args.nel for b in blocks])
    else:      # same element size at every N: true weak scaling, TGV tiles periodically
        spec.length = np.array([2 * np.pi * b for b in blocks])
    kx = 2 * np.pi / spec.length if small else 1.0
    return spec, dict(partition="block", blocks=blocks), (lambda r: r.eqn.tgv_state(r.mesh.Elem_xGP * kx))


def eqn_for(args):
    from flexi_b200.hostmesh.case import NACA_EQN, TGV_EQN

    if args.workload == "naca":  # configs[3]: Gauss nodes, weak form, HLLC + BR2, isothermal wall + Dirichlet far field
        return "gauss", dict(NACA_EQN, n_var_lifting=5)

    if args.flavour == "split":   # configs[1]
        return "gauss-lobatto", dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="lf", cfl_scale=0.8, dfl_scale=0.8,
                                     n_var_lifting=5, **TGV_EQN)
    # configs[0]: Gauss, weak form, Roe
    return "gauss", dict(parabolic=True, split_dg=False, riemann="roe", cfl_scale=0.8, dfl_scale=0.8, n_var_lifting=5, **TGV_EQN)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.rows = []
        self.stop = threading.Event()
        self.index = index
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(r[3 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "power_w_max": max(float(r[2]) for r in self.rows),
                "samples": len(self.rows), "reasons": reasons}


def mem_available_gb():
    try:
        return int([ln.split()[1] for ln in open("/proc/meminfo") if ln.startswith("MemAvailable")][0]) / 1048576.0
    except Exception:
        return 0.0


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_cpu(args, steps, warmup, nel, cores):
    """The reference algorithm on the host cores: oracle port, -O3 build, one emulated MPI rank per core (OpenMP thread)."""
    from flexi_b200.hostmesh.case import build_case
    from oracle import oracle as O

    node, eqn = eqn_for(args)
    spec, _, state = workload_case(args, 1, nel)
    nranks = max(1, min(cores, spec.n_global_elems // 8))
    os.environ["OMP_NUM_THREADS"] = str(nranks)
    ranks = build_case(spec, args.N, node, eqn, nranks=nranks)
    w = O.OracleWorld([(r.desc, r.keep) for r in ranks], fast=True)
    w.set_state([state(r) for r in ranks])
    dt = w.calc_timestep()[0]
    t = 0.0
    for _ in range(warmup):
        w.timestep(t, dt)
        t += dt
    t0 = time.perf_counter()
    for _ in range(steps):
        w.timestep(t, dt)
        t += dt
    el = time.perf_counter() - t0
    w.close()
    ndof = spec.n_global_elems * (args.N + 1) ** 3
    nst = len(ranks[0].eqn.RKb)
    return ndof * nst * steps / el / 1e9, el, nranks, ndof


def ut_scale(ut):
    """Per-variable magnitude of Ut that a difference is judged against: max|Ut_v|, with one common scale for the three momenta and
    the density judged against the momentum scale too (TGV: div(rho u) = 0, so max|Ut_rho| is itself at truncation-error level and
    says nothing about the size of the terms that were summed)."""
    sc = np.max(np.abs(ut).reshape(-1, 5), axis=0)
    sc[0:4] = sc[1:4].max()
    return sc


def parity_gate(args, node, eqn, world, rank, local_rank, dist, torch, G):
    """Parity of the benchmarked code path with the CPU oracle, outside the timed region ('parity gates run with every
    measurement'): same flavour, same partition shape, `--parity-nel`^3 elements per GPU.  Every rank runs its partition on its GPU
    (NCCL side exchange included) and compares its slice with the oracle world of all ranks, which it computes itself on the host.
    Returns max-over-ranks of: one Ut (max|diff|/max|ref| per variable, and the same in units of the gate = max(1e-12*scale,
    4 x the oracle's own strict-vs-FMA build noise)), the state after K fixed-dt LSERK steps (gate 1e-12) and dt (gate 1e-13)."""
    from flexi_b200.hostmesh.case import build_case
    from oracle import oracle as O

    # TGV: state with the wavelength of the parity box, so that it is smooth across the periodic boundaries (the 2*pi-periodic state
    # on this shorter box has O(1) jumps there, and the 2x1x1 box went unstable -- NaN in oracle and GPU alike -- after 4 steps)
    spec, part, state = workload_case(args, world, max(12, 2 * args.parity_nel) if args.workload == "naca" else args.parity_nel, small=True)
    ranks = build_case(spec, args.N, node, eqn, nranks=world, **part)
    me = ranks[rank]
    U0s = [state(r) for r in ranks]
    os.environ["OMP_NUM_THREADS"] = str(max(1, min(world, host_cores() // world)))
    dk = [(r.desc, r.keep) for r in ranks]
    noise, ut_strict = O.ut_noise_floor(dk, U0s)
    nccl_id = None
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            idt = torch.frombuffer(bytearray(G.nccl_unique_id()), dtype=torch.uint8).clone()
        dist.broadcast(idt, src=0)
        nccl_id = bytes(idt.numpy().tobytes())
    me.remake_desc(nccl_id=nccl_id, device=local_rank)
    w = O.OracleWorld(dk)
    w.set_state(U0s)
    out = {}
    with G.FlexiGpu(me) as g:
        g.set_state(U0s[rank])
        ut = g.DGTimeDerivative_weakForm(0.0)
        ref = ut_strict[rank]
        scale = ut_scale(np.concatenate([u.reshape(-1, 5) for u in ut_strict]))
        d = np.max(np.abs(ut - ref).reshape(-1, 5), axis=0)
        # naca: 8 x noise (the wall layer's 1/J amplifies every rounding difference; the CUDA kernels contract other FMAs than gcc)
        gate = np.maximum(1e-12 * scale, (8.0 if args.workload == "naca" else 4.0) * noise)
        out["ut_rel"] = float(np.max(d / scale))
        out["ut_over_gate"] = float(np.max(d / gate))
        out["noise_abs"] = [float(x) for x in noise]
        dt_o = w.calc_timestep()[0]
        dt_g = g.CalcTimeStep()[0]
        out["dt_rel"] = abs(dt_g - dt_o) / dt_o
        t = 0.0
        for _ in range(args.parity_steps):
            w.timestep(t, dt_o)
            g.TimeStep(t, dt_o)
            t += dt_o
        Ug, Uo = g.get_state(), w.ranks[rank].array("U")
        sc = np.max(np.abs(np.concatenate([r.array("U").reshape(-1, 5) for r in w.ranks])), axis=0)
        sc[1:4] = sc[1:4].max()  # TGV: judge the momenta against the common momentum scale
        out["state_rel"] = float(np.max(np.max(np.abs(Ug - Uo).reshape(-1, 5), axis=0) / sc))
    w.close()
    if world > 1:
        keys = ("ut_rel", "ut_over_gate", "dt_rel", "state_rel")
        tt = torch.tensor([out[k] for k in keys], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        out.update(dict(zip(keys, [float(x) for x in tt.tolist()])))
    out["ok"] = bool(out["ut_over_gate"] <= 1.0 and out["state_rel"] <= 1e-12 and out["dt_rel"] <= 1e-13)
    out["what"] = (f"same flavour and partition shape, {args.parity_nel}^3 elements per GPU, GPU (C ABI, NCCL exchange at N>1) vs CPU oracle: one Ut "
                   f"(gate max(1e-12*scale, {8 if args.workload == 'naca' else 4}x oracle strict-vs-FMA noise)), dt (1e-13), state after {args.parity_steps} LSERK steps (1e-12)")
    return out


def parity_fullsize(args, g, r, U0, noise):
    """Checks on the benchmarked mesh itself: discrete conservation of one Ut on the periodic box (sum_elem sum_ijk w_ijk Ut / sJ = 0,
    relative to sum w |Ut| / sJ of the momenta) and, when the host has the memory (~30 GB), that same Ut against the single-rank
    oracle (-O3 build; gate max(1e-12*scale, 8 x the strict-vs-FMA noise measured on the small case: same element size and state))."""
    from oracle import oracle as O

    out = {}
    g.set_state(U0)
    ut = g.DGTimeDerivative_weakForm(0.0)
    w3 = np.einsum("k,j,i->kji", r.basis.wGP, r.basis.wGP, r.basis.wGP)[None] / r.mesh.sJ
    cons = np.einsum("ekjiv,ekji->v", ut, w3)
    mag = np.einsum("ekjiv,ekji->v", np.abs(ut), w3)
    mag[0:4] = mag[1:4].max()
    # walls and far field exchange mass, momentum and energy with the outside: the identity holds on the periodic box only
    out["conservation_rel"] = float(np.max(np.abs(cons) / mag)) if args.workload == "tgv" else 0.0
    if mem_available_gb() > 96.0 and not args.no_fullsize_oracle:
        t0 = time.perf_counter()
        if args.workload == "naca":
            # 1/J of the wall layer grows with the resolution, and the round-off of Ut with it: the noise floor (strict vs FMA
            # build of the oracle) is measured on the benchmarked mesh itself instead of being carried over from the small case
            noise_full, refs = O.ut_noise_floor([(r.desc, r.keep)], [U0])
            ref = refs[0]
            gate_noise = 8.0 * noise_full
        else:
            w = O.OracleWorld([(r.desc, r.keep)], fast=True)
            w.set_state([U0])
            w.dg_time_derivative(0.0)
            ref = np.array(w.ranks[0].array("Ut"))
            w.close()
            gate_noise = 8.0 * np.asarray(noise)
        scale = ut_scale(ref)
        d = np.max(np.abs(ut - ref).reshape(-1, 5), axis=0)
        out["ut_rel_fullsize"] = float(np.max(d / scale))
        out["ut_over_gate_fullsize"] = float(np.max(d / np.maximum(1e-12 * scale, gate_noise)))
        out["fullsize_oracle_s"] = time.perf_counter() - t0
    else:
        out["ut_rel_fullsize"] = out["ut_over_gate_fullsize"] = None
    return out


def emit(line, fd):
    """Write the one JSON line to the real stdout (fd saved before anything could print there)."""
    os.write(fd, (json.dumps(line) + "\n").encode())


def main():
    # stdout carries exactly ONE JSON line: everything else that prints to fd 1 (NCCL's version banner at NCCL_DEBUG=VERSION,
    # torchrun notices, library chatter) is sent to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="flexi_b200", choices=["flexi_b200", "reference"])
    ap.add_argument("--nel", type=int, default=48, help="elements per direction PER GPU (BASELINE configs[1]: 48)")
    ap.add_argument("--workload", default="tgv", choices=["tgv", "naca"],
                    help="tgv: BASELINE configs[1]/[0]/[4]; naca: configs[3] (NACA0012-style curved O-mesh, randomly oriented elements, "
                         "N=5, HLLC + BR2, wall + far field; --nel = elements around the section per GPU-slab: nel x nel/3 x nel/6 per GPU)")
    ap.add_argument("--N", type=int, default=None, help="polynomial degree (default 7; naca: 5)")
    ap.add_argument("--flavour", default="split", choices=["split", "weak"], help="split: configs[1]; weak: configs[0] flavour")
    ap.add_argument("--cpu-nel", type=int, default=0, help="box size of the CPU run; 0 = --impl reference: the benchmarked box (--nel) when the "
                    "host has the memory, else 16; cpu_baseline leg of the repo arm: 16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parity-nel", type=int, default=6, help="elements per direction per GPU of the parity gate (0 = skip)")
    ap.add_argument("--parity-steps", type=int, default=5)
    ap.add_argument("--no-fullsize-oracle", action="store_true", help="skip the single-rank oracle Ut on the benchmarked mesh (N=1)")
    ap.add_argument("--profile-steps", type=int, default=2, help="extra steps with per-kernel CUDA events for the roofline object")
    args = ap.parse_args()
    if args.N is None:
        args.N = 5 if args.workload == "naca" else 7
    if args.workload == "naca" and args.nel == 48:
        args.nel = 96

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    nst = 5
    workload = (f"TGV Navier-Stokes N={args.N}, {args.nel}^3 elements per GPU, "
                + ("split-form Gauss-Lobatto (Pirozzoli) + Lax-Friedrichs" if args.flavour == "split" else "weak-form Gauss + Roe")
                + ", BR1, LSERK4-5 (carpenterrk4-5)")
    if args.workload == "naca":
        n3 = naca_nel(args.nel, 1)
        workload = (f"NACA0012-style synthetic curved O-mesh (NGeo=4, every element in a random local frame), Navier-Stokes N={args.N}, "
                    f"{n3[0]}x{n3[1]}x{n3[2]} elements per GPU (z-slabs), weak-form Gauss, HLLC + BR2, isothermal wall + Dirichlet far field, LSERK4-5")

    if args.impl == "reference":
        if rank != 0:
            return
        cores = host_cores()
        cpu_nel = args.cpu_nel
        if cpu_nel == 0:  # the benchmarked configuration itself needs ~0.6 KB per DOF on the host (oracle arrays + host mesh)
            nE_full = int(np.prod(naca_nel(args.nel, 1))) if args.workload == "naca" else args.nel ** 3
            need_gb = nE_full * (args.N + 1) ** 3 * 900.0 / 2 ** 30
            cpu_nel = args.nel if mem_available_gb() > need_gb + 16.0 else (24 if args.workload == "naca" else 16)
        # the whole run has to end within a few minutes: ~0.18 us per DOF and stage and core
        nE_cpu = int(np.prod(naca_nel(cpu_nel, 1))) if args.workload == "naca" else cpu_nel ** 3
        est_step_s = nE_cpu * (args.N + 1) ** 3 * nst * 0.6e-6 / max(1, min(cores, nE_cpu // 8))
        steps = max(1, min(args.steps, int(150.0 / est_step_s)))
        warmup = max(0, min(args.warmup, int(30.0 / est_step_s)))
        val, el, nr, ndof = run_cpu(args, steps, warmup, cpu_nel, cores)
        same = "the benchmarked box itself" if cpu_nel == args.nel else "a bounded sample of the workload"
        sample = (f"{same}: same flavour, {nE_cpu} elements ({ndof} DOF), {steps} timed LSERK steps after {warmup} warm-up ({el:.1f} s), {nr} emulated "
                  f"MPI ranks (one per OpenMP thread), C port of the reference (oracle/flexi_oracle.c, -O3 -march=x86-64-v3); no Fortran compiler in the image")
        line = {"impl": "reference", "metric": "GDOF/s (DOF*RK-stages per second), PID = 1/(value*1e9) s/DOF/stage", "value": val,
                "unit": "GDOF/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": el / steps * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "pid_s_per_dof_stage": 1.0 / (val * 1e9), "pid_core_s": nr / (val * 1e9),
                "config": {"workload": workload, "cpu_sample_nel": cpu_nel, "same_box_as_gpu_arm": cpu_nel == args.nel},
                "cpu_baseline": {"value": val, "unit": "GDOF/s", "cores": nr, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line, real_stdout)
        return

    import torch
    import torch.distributed as dist

    from flexi_b200 import gpu as G
    from flexi_b200.hostmesh.case import build_case, tgv_box

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL_DEBUG is left as the caller set it: its output goes to fd 1, which main() redirected to stderr
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}"

    node, eqn = eqn_for(args)
    parity = None
    if args.parity_nel > 0:
        parity = parity_gate(args, node, eqn, world, rank, local_rank, dist if world > 1 else None, torch, G)
        if not parity["ok"]:
            raise SystemExit(f"bench.py: parity gate failed, no number is reported: {parity}")
    blocks = BLOCKS[world]
    spec, part, state = workload_case(args, world, args.nel)
    t_mesh = time.perf_counter()
    r = build_case(spec, args.N, node, eqn, nranks=world, ranks=[rank], **part)[0]
    t_mesh = time.perf_counter() - t_mesh
    nccl_id = None
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            idt = torch.frombuffer(bytearray(G.nccl_unique_id()), dtype=torch.uint8).clone()
        dist.broadcast(idt, src=0)
        nccl_id = bytes(idt.numpy().tobytes())
    r.remake_desc(nccl_id=nccl_id, device=local_rank)
    g = G.FlexiGpu(r)
    nst = len(r.eqn.RKb)
    ndof_local = r.mesh.nElems * (args.N + 1) ** 3
    ndof = ndof_local * world

    U0 = torch.empty(r.mesh.Elem_xGP.shape[:-1] + (5,), dtype=torch.float64, pin_memory=True)
    U0.numpy()[...] = state(r)
    Uout = torch.empty_like(U0, pin_memory=True)
    g.set_state(U0.numpy())
    dt = g.CalcTimeStep()[0]
    stream = torch.cuda.ExternalStream(g.stream(), device=local_rank)

    def barrier():
        g.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # ---- device-resident timing: K steps, inputs in HBM (state 2.3 GB per GPU >> 126 MB L2, no flush needed) ----
    t = 0.0
    for _ in range(args.warmup):
        g.TimeStep(t, dt)
        t += dt
    c0 = g.counters()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record(stream)
        for _ in range(args.steps):
            g.TimeStep(t, dt)
            t += dt
        e1.record(stream)
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    c1 = g.counters()
    launches = c1["kernel_launches"] - c0["kernel_launches"]
    value = ndof * nst * args.steps / (ms * 1e-3) / 1e9
    assert np.isfinite(g.CalcTimeStep()[0]), "state blew up"

    # ---- end to end through the C ABI with host buffers: one analyze interval = set_state + K x (CalcTimeStep + TimeStep)
    #      + get_state, host<->device copies (pinned) inside the timed region ----
    barrier()
    w0 = time.perf_counter()
    g.set_state(U0.numpy())
    tt = 0.0
    for _ in range(args.steps):
        dte = g.CalcTimeStep()[0]
        g.TimeStep(tt, dte)
        tt += dte
    g.get_state(Uout.numpy())
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - w0)
    e2e_val = ndof * nst * args.steps / e2e_s / 1e9
    ubytes = U0.numel() * 8
    assert np.all(np.isfinite(Uout.numpy()[::97].ravel()[:1000]))

    # ---- per-kernel device times (CUDA events on the launching stream) for the roofline object ----
    g.set_profiling(True)
    g.kernel_times(reset=True)
    for _ in range(max(1, args.profile_steps)):
        g.TimeStep(t, dt)
        t += dt
    kt = g.kernel_times(reset=True)
    g.set_profiling(False)
    args.profile_steps = max(1, args.profile_steps)
    peak, peak_src = peaks()
    balg = b_alg_kernel(args.N, True)
    kern = {}
    for name in ("volume", "flux", "update"):
        msk, n = kt[name]
        if n:
            per_launch_ms = msk / n
            launches_per_stage = n / (args.profile_steps * nst)
            bytes_per_launch = balg[name] * ndof_local / launches_per_stage
            kern[name] = {"avg_ms": per_launch_ms, "launches_per_stage": launches_per_stage, "alg_bytes_per_dof": balg[name],
                          "achieved_gbs": bytes_per_launch / (per_launch_ms * 1e-3) / 1e9}
    # multi-GPU: time the compute stream stood still at the three halo-exchange waits (what the inner-element kernels did not
    # hide), per RK stage, max over ranks
    xwait = None
    if world > 1:
        xwait = {k: max_over_ranks(kt["wait_" + k][0] / (args.profile_steps * nst)) for k in ("trace", "grad", "flux")}
    dom = max(kern, key=lambda k: kern[k]["avg_ms"] * kern[k]["launches_per_stage"])
    stage_ms = sum(v["avg_ms"] * v["launches_per_stage"] for v in kern.values())
    kname = {"volume": "k_volume_gl2" if (args.workload == "tgv" and args.flavour == "split" and (args.N + 1) % 2 == 0) else "k_volume", "flux": "k_flux", "update": "k_update"}
    roofline = {"bound": "hbm", "kernel": kname[dom], "achieved": kern[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["achieved_gbs"] / peak, "traffic": None, "peak_source": peak_src,
                "share_of_stage": kern[dom]["avg_ms"] * kern[dom]["launches_per_stage"] / stage_ms, "kernels": kern,
                "stage": {"alg_bytes_per_dof": B_ALG_STAGE[True](args.N), "achieved_gbs": B_ALG_STAGE[True](args.N) * ndof_local * nst * args.steps / (ms * 1e-3) / 1e9,
                          "frac": B_ALG_STAGE[True](args.N) * ndof_local * nst * args.steps / (ms * 1e-3) / 1e9 / peak}}
    tr = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get(kname[dom])
        except Exception:
            pass

    if parity is not None and world == 1:
        parity.update(parity_fullsize(args, g, r, U0.numpy(), parity["noise_abs"]))
        if parity["conservation_rel"] > 1e-12 or (parity["ut_over_gate_fullsize"] is not None and parity["ut_over_gate_fullsize"] > 1.0):
            raise SystemExit(f"bench.py: full-size parity check failed: {parity}")
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        cnel = args.cpu_nel or (24 if args.workload == "naca" else 16)
        val, el, nr, nd = run_cpu(args, 30, 2, cnel, cores)
        cpu = {"value": val, "unit": "GDOF/s", "cores": nr, "kind": "port",
               "sample": f"same flavour, {'NACA O-mesh nel=' if args.workload == 'naca' else 'box nel='}{cnel} ({nd} DOF), 30 timed LSERK steps in {el:.1f} s, {nr} emulated MPI ranks, "
                         "C port of the reference (-O3 -march=x86-64-v3)"}
    g.close()
    if rank == 0:
        line = {"metric": "GDOF/s (DOF*RK-stages per second), PID = 1/(value*1e9) s/DOF/stage", "value": value, "unit": "GDOF/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "pid_s_per_dof_stage": 1.0 / (value * 1e9), "pid_x_ngpu": world / (value * 1e9),
                "config": {"workload": workload, "n_dof": ndof, "n_elems_per_gpu": r.mesh.nElems, "rk_stages": nst, "dt": dt,
                           "partition": ("x".join(map(str, blocks)) + " blocks") if args.workload == "tgv" else f"{world} z-slabs (linear)", "l2": "working set 20 GB per GPU >> 126 MB L2 (no flush needed)",
                           "mesh_build_s": t_mesh},
                "clocks": clk.summary(), "gpu_launches": launches,
                "e2e": {"value": e2e_val, "unit": "GDOF/s", "h2d_bytes_per_step": ubytes / args.steps,
                        "d2h_bytes_per_step": ubytes / args.steps + 24,
                        "what": "set_state(pinned host U) + steps x (CalcTimeStep + TimeStep) + get_state(pinned host U) through the C ABI"},
                "roofline": roofline, "cpu_baseline": cpu, "parity": parity}
        if xwait is not None:
            line["exchange_wait_ms_per_stage"] = xwait
        emit(line, real_stdout)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
