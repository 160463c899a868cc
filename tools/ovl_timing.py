"""Timing experiments on the development library (not a benchmark; FLEXI_GPU_OVL / FX_DEV_* variants produce wrong results):
wall time of repeated DG operator evaluations (mode 0: U is never changed).
usage: FLEXI_GPU_LIB=flexi_b200/lib/libflexi_gpu_dev.so [FLEXI_GPU_OVL=1|2] python tools/ovl_timing.py [nel]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flexi_b200 import gpu as G  # noqa: E402
from flexi_b200.hostmesh.case import TGV_EQN, build_case, tgv_box  # noqa: E402

nel = int(sys.argv[1]) if len(sys.argv) > 1 else 32
eqn = dict(parabolic=True, split_dg=True, split_dg_type="pi", riemann="lf", cfl_scale=0.8, dfl_scale=0.8, n_var_lifting=5, **TGV_EQN)
r = build_case(tgv_box(nel), 7, "gauss-lobatto", eqn)[0]
g = G.FlexiGpu(r)
g.set_state(r.eqn.tgv_state(r.mesh.Elem_xGP))
dt = g.CalcTimeStep()[0]
for _ in range(2):
    g.TimeStep(0.0, dt)
g.synchronize()
n = 4
t0 = time.perf_counter()
for _ in range(n):
    g.TimeStep(0.0, dt)
g.synchronize()
wall = (time.perf_counter() - t0) / n / 5
g.set_profiling(True)
g.kernel_times(reset=True)
for _ in range(2):
    g.TimeStep(0.0, dt)
kt = g.kernel_times(reset=True)
print("lib=%s ovl=%s nel=%d: wall per stage %.3f ms | events: volume %.3f flux %.3f update %.3f" % (
    os.path.basename(G.LIB_PATH), os.environ.get("FLEXI_GPU_OVL", "0"), nel, wall * 1e3, kt["volume"][0] / kt["volume"][1],
    kt["flux"][0] / max(kt["flux"][1], 1), kt["update"][0] / max(kt["update"][1], 1)))
g.close()
