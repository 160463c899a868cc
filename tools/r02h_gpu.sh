#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r02h.txt
for v in pi mg2 mg3; do
  export FLEXI_GPU_LIB=flexi_b200/lib/libflexi_gpu_dev_$v.so
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 >> gpurun_out/r02h.txt
  python tools/phase_stamps.py 32 2>&1 | tail -14 >> gpurun_out/r02h.txt
done
cat gpurun_out/r02h.txt
