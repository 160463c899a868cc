"""Multi-rank debugging aid (torchrun): the bench.py parity gate with per-step diagnostics.
usage: torchrun --nproc-per-node 2 tools/mr_debug.py [N] [nel_per_gpu] [steps]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flexi_b200 import gpu as G  # noqa: E402
from flexi_b200.hostmesh.case import TGV_EQN, build_case, tgv_box  # noqa: E402
from oracle import oracle as O  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 7
pn = int(sys.argv[2]) if len(sys.argv) > 2 else 6
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)
BLOCKS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
blocks = BLOCKS[world]
eqn = dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="lf", cfl_scale=0.8, dfl_scale=0.8, **TGV_EQN)
spec = tgv_box(tuple(pn * b for b in blocks))
if len(sys.argv) > 4:  # as bench.py's parity gate: element size of the benchmarked mesh (argv[4] = its elements per direction)
    spec.length = np.array([2 * np.pi * b * pn / int(sys.argv[4]) for b in blocks])
ranks = build_case(spec, N, "gauss-lobatto", eqn, nranks=world, partition="block", blocks=blocks)
me = ranks[rank]
U0s = [r.eqn.tgv_state(r.mesh.Elem_xGP) for r in ranks]
idt = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    idt = torch.frombuffer(bytearray(G.nccl_unique_id()), dtype=torch.uint8).clone()
dist.broadcast(idt, src=0)
me.remake_desc(nccl_id=bytes(idt.numpy().tobytes()), device=local)
w = O.OracleWorld([(r.desc, r.keep) for r in ranks])
w.set_state(U0s)
g = G.FlexiGpu(me)
g.set_state(U0s[rank])
if len(sys.argv) > 4:
    ut = g.DGTimeDerivative_weakForm(0.0)
    w.dg_time_derivative(0.0)
    print(f"rank {rank}: Ut diff {np.max(np.abs(ut - w.ranks[rank].array('Ut'))):.3e} nan {np.isnan(ut).sum()}", flush=True)
dt = w.calc_timestep()[0]
print(f"rank {rank}: dt oracle {dt:.6e} gpu {g.CalcTimeStep()[0]:.6e}", flush=True)
t = 0.0
for s in range(steps):
    w.timestep(t, dt)
    g.TimeStep(t, dt)
    t += dt
    Ug, Uo = g.get_state(), w.ranks[rank].array("U")
    d = np.abs(Ug - Uo)
    bad = np.argwhere(~np.isfinite(Ug))
    print(f"rank {rank} N={N} step {s}: max diff {np.nanmax(d):.3e} nan(gpu) {np.isnan(Ug).sum()} nan(oracle) {np.isnan(Uo).sum()} "
          f"first bad {bad[0].tolist() if len(bad) else None} shape {Ug.shape}", flush=True)
g.close()
w.close()
dist.destroy_process_group()
