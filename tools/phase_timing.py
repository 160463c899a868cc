"""Timing experiment for the element kernel (not a benchmark): repeated DG operator evaluations (mode 0: U is never changed, so
the FLEXI_GPU_DBG phase-skipping switches, which produce wrong Ut, are harmless) with per-kernel CUDA events.
usage: FLEXI_GPU_DBG=<bits> FLEXI_GPU_SMEM_PAD=<bytes> python tools/phase_timing.py [nel] [euler]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flexi_b200 import gpu as G  # noqa: E402
from flexi_b200.hostmesh.case import TGV_EQN, build_case, tgv_box  # noqa: E402

nel = int(sys.argv[1]) if len(sys.argv) > 1 else 32
para = not (len(sys.argv) > 2 and sys.argv[2] == "euler")  # "euler": split-form Euler, i.e. the two-point sweeps without lifting
eqn = dict(parabolic=para, split_dg=True, split_dg_type="pi", riemann="lf", cfl_scale=0.8, dfl_scale=0.8, n_var_lifting=5, **TGV_EQN)
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
print("%s dbg=%s pad=%s volume %.3f ms flux %.3f update %.3f" % ("NS" if para else "Euler", os.environ.get("FLEXI_GPU_DBG", "0"), os.environ.get("FLEXI_GPU_SMEM_PAD", "0"), kt["volume"][0] / kt["volume"][1], kt["flux"][0] / max(kt["flux"][1], 1), kt["update"][0] / max(kt["update"][1], 1)))
g.close()
