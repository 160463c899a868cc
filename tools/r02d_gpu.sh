#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02d_pytest.log 2>&1
tail -5 gpurun_out/r02d_pytest.log
FLEXI_GPU_LIB=flexi_b200/lib/libflexi_gpu_dev.so python tools/ovl_timing.py 32 2>&1 | tail -1 > gpurun_out/r02d_timing.txt
for p in 74 222 296 444; do FLEXI_GPU_PREFETCH=$p python tools/ovl_timing.py 32 2>&1 | tail -1 | sed "s/^/prefetch $p: /" >> gpurun_out/r02d_timing.txt; done
cat gpurun_out/r02d_timing.txt
python bench.py --no-cpu-baseline > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
tail -3 gpurun_out/r02d_bench.err
cat gpurun_out/r02d_bench.json
