#!/bin/bash
# r02r: failing tests of r02q again + persistent element kernel (dev library): smoke parity, timing with and without
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_naca_mesh.py tests/test_zz_gpu_split_formulations.py -q -m gpu -k "naca or sponge" 2>&1 | tail -5 > gpurun_out/r02r.txt
export FLEXI_GPU_LIB=flexi_b200/lib/libflexi_gpu_dev_pi.so
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 >> gpurun_out/r02r.txt
python tools/ovl_timing.py 32 2>&1 | tail -1 >> gpurun_out/r02r.txt
FLEXI_GPU_PERSIST=0 python tools/ovl_timing.py 32 2>&1 | tail -1 | sed 's/^/persist=0 /' >> gpurun_out/r02r.txt
python tools/ovl_timing.py 48 2>&1 | tail -1 >> gpurun_out/r02r.txt
cat gpurun_out/r02r.txt
