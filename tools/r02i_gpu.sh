#!/bin/bash
mkdir -p gpurun_out
export FLEXI_GPU_LIB=flexi_b200/lib/libflexi_gpu_dev_pi.so
ncu --set full --import-source on --clock-control none --kernel-name regex:k_volume_gl2 --launch-skip 12 --launch-count 1 -f -o gpurun_out/r02i_gl2 python tools/ovl_timing.py 32 > gpurun_out/r02i_ncu.log 2>&1
tail -3 gpurun_out/r02i_ncu.log
ls -la gpurun_out/
