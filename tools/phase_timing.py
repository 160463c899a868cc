"""Timing experiment for the element kernel (not a benchmark): repeated DG operator evaluations (mode 0: U is never changed, so
the FLEXI_GPU_DBG phase-skipping switches, which produce wrong Ut, are harmless) with per-kernel CUDA events.
usage: FLEXI_GPU_DBG=<bits> FLEXI_GPU_SMEM_PAD=<bytes> python tools/phase_timing.py [nel]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flexi_b200 import gpu as G  # noqa: E402
from flexi_b200.hostmesh.case import TGV_EQN, build_case, tgv_box  # noqa: E402

nel = int(sys.argv[1]) if len(sys.argv) > 1 else 32
eqn = dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="lf", cfl_scale=0.8, dfl_scale=0.8, n_var_lifting=5, **TGV_EQN)
r = build_case(tgv_box(nel), 7, "gauss-lobatto", eqn)[0]
g = G.FlexiGpu(r)
g.set_state(r.eqn.tgv_state(r.mesh.Elem_xGP))
for _ in range(2):
    g.DGTimeDerivative_weakForm(0.0, want_ut=False)
g.set_profiling(True)
g.kernel_times(reset=True)
for _ in range(5):
    g.DGTimeDerivative_weakForm(0.0, want_ut=False)
kt = g.kernel_times(reset=True)
print("dbg=%s pad=%s volume %.3f ms" % (os.environ.get("FLEXI_GPU_DBG", "0"), os.environ.get("FLEXI_GPU_SMEM_PAD", "0"), kt["volume"][0] / kt["volume"][1]))
g.close()
