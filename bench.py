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


def eqn_for(args):
    from flexi_b200.hostmesh.case import TGV_EQN

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


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_cpu(args, steps, warmup, nel, cores):
    """The reference algorithm on the host cores: oracle port, -O3 build, one emulated MPI rank per core (OpenMP thread)."""
    from flexi_b200.hostmesh.case import build_case, tgv_box
    from oracle import oracle as O

    node, eqn = eqn_for(args)
    nranks = max(1, min(cores, nel ** 3 // 8))
    os.environ["OMP_NUM_THREADS"] = str(nranks)
    ranks = build_case(tgv_box(nel), args.N, node, eqn, nranks=nranks)
    w = O.OracleWorld([(r.desc, r.keep) for r in ranks], fast=True)
    w.set_state([r.eqn.tgv_state(r.mesh.Elem_xGP) for r in ranks])
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
    ndof = nel ** 3 * (args.N + 1) ** 3
    nst = len(ranks[0].eqn.RKb)
    return ndof * nst * steps / el / 1e9, el, nranks, ndof


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
    ap.add_argument("--N", type=int, default=7)
    ap.add_argument("--flavour", default="split", choices=["split", "weak"], help="split: configs[1]; weak: configs[0] flavour")
    ap.add_argument("--cpu-nel", type=int, default=16, help="box size of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-steps", type=int, default=2, help="extra steps with per-kernel CUDA events for the roofline object")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    nst = 5
    workload = (f"TGV Navier-Stokes N={args.N}, {args.nel}^3 elements per GPU, "
                + ("split-form Gauss-Lobatto (Pirozzoli) + Lax-Friedrichs" if args.flavour == "split" else "weak-form Gauss + Roe")
                + ", BR1, LSERK4-5 (carpenterrk4-5)")

    if args.impl == "reference":
        if rank != 0:
            return
        cores = host_cores()
        steps = max(1, min(args.steps, 40))  # 16^3 elements at N=7: ~0.4 s per LSERK step on 16 cores
        val, el, nr, ndof = run_cpu(args, steps, min(args.warmup, 3), args.cpu_nel, cores)
        sample = (f"same flavour on a {args.cpu_nel}^3-element box ({ndof} DOF), {steps} timed LSERK steps, {nr} emulated MPI ranks "
                  f"(one per OpenMP thread), C port of the reference (oracle/flexi_oracle.c, -O3 -march=x86-64-v3); no Fortran compiler in the image")
        line = {"impl": "reference", "metric": "GDOF/s (DOF*RK-stages per second), PID = 1/(value*1e9) s/DOF/stage", "value": val,
                "unit": "GDOF/s", "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 3), "ms_per_step": el / steps * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "pid_s_per_dof_stage": 1.0 / (val * 1e9), "pid_core_s": nr / (val * 1e9),
                "config": {"workload": workload, "cpu_sample_nel": args.cpu_nel},
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
        # stdout carries exactly one JSON line: NCCL prints its version banner there at NCCL_DEBUG=VERSION/INFO
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE") and not os.environ.get("FLEXI_BENCH_KEEP_NCCL_DEBUG"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}"

    node, eqn = eqn_for(args)
    blocks = BLOCKS[world]
    nel = tuple(args.nel * b for b in blocks)
    spec = tgv_box(nel)
    spec.length = np.array([2 * np.pi * b for b in blocks])  # same element size at every N: true weak scaling, TGV tiles periodically
    t_mesh = time.perf_counter()
    r = build_case(spec, args.N, node, eqn, nranks=world, partition="block", blocks=blocks, ranks=[rank])[0]
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
    U0.numpy()[...] = r.eqn.tgv_state(r.mesh.Elem_xGP)
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
    dom = max(kern, key=lambda k: kern[k]["avg_ms"] * kern[k]["launches_per_stage"])
    stage_ms = sum(v["avg_ms"] * v["launches_per_stage"] for v in kern.values())
    kname = {"volume": "k_volume_gl2" if (args.flavour == "split" and (args.N + 1) % 2 == 0) else "k_volume", "flux": "k_flux", "update": "k_update"}
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

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        val, el, nr, nd = run_cpu(args, 30, 2, args.cpu_nel, cores)
        cpu = {"value": val, "unit": "GDOF/s", "cores": nr, "kind": "port",
               "sample": f"same flavour on a {args.cpu_nel}^3-element box ({nd} DOF), 30 timed LSERK steps in {el:.1f} s, {nr} emulated MPI ranks, "
                         "C port of the reference (-O3 -march=x86-64-v3)"}
    g.close()
    if rank == 0:
        line = {"metric": "GDOF/s (DOF*RK-stages per second), PID = 1/(value*1e9) s/DOF/stage", "value": value, "unit": "GDOF/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "pid_s_per_dof_stage": 1.0 / (value * 1e9), "pid_x_ngpu": world / (value * 1e9),
                "config": {"workload": workload, "n_dof": ndof, "n_elems_per_gpu": r.mesh.nElems, "rk_stages": nst, "dt": dt,
                           "partition": "x".join(map(str, blocks)) + " blocks", "l2": "working set 20 GB per GPU >> 126 MB L2 (no flush needed)",
                           "mesh_build_s": t_mesh},
                "clocks": clk.summary(), "gpu_launches": launches,
                "e2e": {"value": e2e_val, "unit": "GDOF/s", "h2d_bytes_per_step": ubytes / args.steps,
                        "d2h_bytes_per_step": ubytes / args.steps + 24,
                        "what": "set_state(pinned host U) + steps x (CalcTimeStep + TimeStep) + get_state(pinned host U) through the C ABI"},
                "roofline": roofline, "cpu_baseline": cpu}
        emit(line, real_stdout)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
