#!/bin/bash
mkdir -p gpurun_out
export FLEXI_GPU_LIB=flexi_b200/lib/libflexi_gpu_dev_st.so
python tools/phase_stamps.py 32 2>&1 | tail -14 > gpurun_out/r02s.txt
FLEXI_GPU_PERSIST=0 python tools/phase_stamps.py 32 2>&1 | tail -14 >> gpurun_out/r02s.txt
cat gpurun_out/r02s.txt
