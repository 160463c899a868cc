#!/bin/bash
set -x
mkdir -p gpurun_out
export FLEXI_GPU_LIB=flexi_b200/lib/libflexi_gpu_dev_pi.so
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > gpurun_out/r02g_smoke.txt
cat gpurun_out/r02g_smoke.txt
python tools/phase_stamps.py 32 2>&1 | tail -16 > gpurun_out/r02g_stamps.txt
cat gpurun_out/r02g_stamps.txt
