#!/bin/bash
set -x
mkdir -p gpurun_out
export FLEXI_GPU_LIB=flexi_b200/lib/libflexi_gpu_dev_pi.so
python tools/phase_stamps.py 32 2>&1 | tail -30 > gpurun_out/r02f_stamps.txt
cat gpurun_out/r02f_stamps.txt
