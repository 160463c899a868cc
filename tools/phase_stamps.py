"""Per-phase warp time of k_volume_gl2 from the clock64() stamps of the development build (make dev).
usage: FLEXI_GPU_LIB=flexi_b200/lib/libflexi_gpu_dev.so python tools/phase_stamps.py [nel]"""
import ctypes as C
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
dt = g.CalcTimeStep()[0]
g.TimeStep(0.0, dt)
L = G.lib()
buf = (C.c_ulonglong * 32)()
L.flexi_gpu_dev_stamps(buf)
n = 2
g.set_profiling(True)
g.kernel_times(reset=True)
for _ in range(n):
    g.TimeStep(0.0, dt)
kt = g.kernel_times(reset=True)
L.flexi_gpu_dev_stamps(buf)
names = {0: "A: loads + ConsToPrim", 1: "S1 wait", 2: "B: normals, lifting sweeps, face lifting flux", 3: "cp.async wait + S2", 4: "node phase",
         5: "S3 wait", 6: "cpart + S3b", 7: "sweep zeta", 10: "S4 wait", 8: "sweep xi", 11: "S5 wait", 9: "sweep eta", 12: "tail"}
nw = nel ** 3 * 4 * n * 5
tot = sum(buf[i] for i in range(32))
print("lib=%s nel=%d volume %.3f ms; per-warp lifetime %.0f cycles" % (os.path.basename(G.LIB_PATH), nel, kt["volume"][0] / kt["volume"][1], tot / nw))
for i, nm in names.items():
    print("  %-36s %8.0f cycles %5.1f %%" % (nm, buf[i] / nw, 100. * buf[i] / tot))
g.close()
