"""p-sweep of the hot path on one GPU (BASELINE configs[0] and configs[2]): N = 3,5,7,9 x Gauss / Gauss-Lobatto x Euler / Navier-Stokes
x weak / split form on a TGV box sized to ~17 MDOF (state 0.7 GB >> 126 MB L2).  Device-timed LSERK4-5 steps with the state
resident in HBM; per flavour: GDOF/s, PID, ms per stage per kernel class, and the stage roofline fraction against B_alg(N) of
BASELINE.md section 3 and the measured HBM peak.  One JSON object per line.
usage: python tools/psweep.py [out.json] [--quick] [--only=N] [--flavour=substring]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from flexi_b200 import gpu as G  # noqa: E402
from flexi_b200.hostmesh import basis as B  # noqa: E402
from flexi_b200.hostmesh.case import TGV_EQN, HostRank, build_case, tgv_box  # noqa: E402
from flexi_b200.hostmesh.equation import Equation  # noqa: E402
from flexi_b200.hostmesh import mappings as MP  # noqa: E402

B_ALG = {True: lambda N: 240.0 + 3.0 / (N + 1) * 720.0, False: lambda N: 240.0 + 3.0 / (N + 1) * 240.0}
NEL = {3: 64, 5: 42, 7: 32, 9: 26}
quick = "--quick" in sys.argv
out_path = next((a for a in sys.argv[1:] if not a.startswith("--")), None)
only_n = next((int(a.split("=")[1]) for a in sys.argv[1:] if a.startswith("--only=")), None)
only_f = next((a.split("=")[1] for a in sys.argv[1:] if a.startswith("--flavour=")), None)
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6650.0
FLAVOURS = [  # (name, node, parabolic, split)
    ("LGL split NS", B.NODE_GAUSS_LOBATTO, True, True), ("LGL split Euler", B.NODE_GAUSS_LOBATTO, False, True),
    ("LGL weak NS", B.NODE_GAUSS_LOBATTO, True, False), ("LGL weak Euler", B.NODE_GAUSS_LOBATTO, False, False),
    ("Gauss weak NS", B.NODE_GAUSS, True, False), ("Gauss weak Euler", B.NODE_GAUSS, False, False),
    ("Gauss split NS", B.NODE_GAUSS, True, True),
]
rows = []
for N in ((only_n,) if only_n else (7,) if quick else (3, 5, 7, 9)):
    nel = NEL[N] // (2 if quick else 1)
    meshes = {}
    for name, node, para, split in FLAVOURS:
        if only_f and only_f.lower() not in name.lower():
            continue
        eqn = dict(parabolic=para, split_dg=split, riemann="lf" if split else "roe", cfl_scale=0.8, dfl_scale=0.8, n_var_lifting=5, **TGV_EQN)
        if split:
            eqn["split_dg_type"] = "pi"
        ja = split and node == B.NODE_GAUSS
        key = (node, ja)
        t0 = time.perf_counter()
        if key not in meshes:
            meshes[key] = build_case(tgv_box(nel), N, node, eqn, with_ja_face=ja)[0].mesh
        r = HostRank(meshes[key], B.DGBasis(N, node, split_dg=split), Equation(N, node, **eqn), MP.build_mappings(N)[2])
        ndof = r.mesh.nElems * (N + 1) ** 3
        g = G.FlexiGpu(r)
        if ja:
            g.set_face_metrics(r.mesh.Ja_Face, r.mesh.Ja_slave)
        g.set_state(r.eqn.tgv_state(r.mesh.Elem_xGP))
        dt = g.CalcTimeStep()[0]
        t = 0.0
        for _ in range(3):
            g.TimeStep(t, dt)
            t += dt
        g.synchronize()
        stream = torch.cuda.ExternalStream(g.stream())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 5
        e0.record(stream)
        for _ in range(steps):
            g.TimeStep(t, dt)
            t += dt
        e1.record(stream)
        g.synchronize()
        ms = e0.elapsed_time(e1)
        g.set_profiling(True)
        g.kernel_times(reset=True)
        for _ in range(2):
            g.TimeStep(t, dt)
            t += dt
        kt = g.kernel_times(reset=True)
        U = g.get_state()
        assert np.all(np.isfinite(U[::37].ravel()[:2000]))
        g.close()
        val = ndof * 5 * steps / (ms * 1e-3) / 1e9
        balg = B_ALG[para](N)
        row = {"N": N, "flavour": name, "n_elems": r.mesh.nElems, "n_dof": ndof, "GDOF_s": val, "PID_s": 1.0 / (val * 1e9),
               "ms_per_stage": ms / steps / 5, "kernel_ms_per_stage": {k: kt[k][0] / 10 for k in ("volume", "flux", "update")},
               "B_alg_bytes_per_dof": balg, "stage_roofline_frac": balg * val / peak, "hbm_peak_gbs": peak}
        rows.append(row)
        print(json.dumps(row), flush=True)
if out_path:
    json.dump(rows, open(out_path, "w"), indent=1)
