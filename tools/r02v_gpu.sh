#!/bin/bash
mkdir -p gpurun_out
export FLEXI_GPU_LIB=flexi_b200/lib/libflexi_gpu_dev_pi.so
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 > gpurun_out/r02v.txt
python tools/ovl_timing.py 32 2>&1 | tail -1 >> gpurun_out/r02v.txt
python tools/ovl_timing.py 48 2>&1 | tail -1 >> gpurun_out/r02v.txt
cat gpurun_out/r02v.txt
