#!/bin/bash
# r02e: v3 element kernel (metric planes in shared memory) on the development library: smoke parity + stage timing at 32^3
set -x
mkdir -p gpurun_out
export FLEXI_GPU_LIB=flexi_b200/lib/libflexi_gpu_dev_pi.so
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02e_smoke.txt 2>&1
tail -3 gpurun_out/r02e_smoke.txt
python tools/ovl_timing.py 32 2>&1 | tail -1 > gpurun_out/r02e_timing.txt
python tools/ovl_timing.py 32 2>&1 | tail -1 >> gpurun_out/r02e_timing.txt
cat gpurun_out/r02e_timing.txt
