#!/bin/bash
# r02b: pipe micro-benchmark, kernel-concurrency and occupancy experiments (development library), bench with the parity gate
set -x
mkdir -p gpurun_out
./tools/ubench/fp64_pipes > gpurun_out/r02b_fp64_pipes.txt 2>&1
D=flexi_b200/lib/libflexi_gpu_dev.so
for o in 0 1 2; do FLEXI_GPU_LIB=$D FLEXI_GPU_OVL=$o python tools/ovl_timing.py 32 2>&1 | tail -1 >> gpurun_out/r02b_ovl.txt; done
FLEXI_GPU_LIB=flexi_b200/lib/libflexi_gpu_dev_w4m3.so python tools/ovl_timing.py 32 2>&1 | tail -1 >> gpurun_out/r02b_ovl.txt
FLEXI_GPU_LIB=$D FLEXI_GPU_SMEM_PAD=16384 python tools/ovl_timing.py 32 2>&1 | tail -1 >> gpurun_out/r02b_ovl.txt
cat gpurun_out/r02b_fp64_pipes.txt gpurun_out/r02b_ovl.txt
python bench.py --steps 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
tail -3 gpurun_out/r02b_bench.err
cat gpurun_out/r02b_bench.json
