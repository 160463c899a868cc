#!/bin/bash
# r02aa (2 GPUs): solution filter on partitioned runs, naca multi-rank case, descriptor path of the face metrics, filter cases
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02aa_pytest_2gpu.log
timeout 1200 python -m pytest tests/test_multi_rank.py tests/test_gpu_gauss_split.py tests/test_zz_gpu_split_formulations.py -q -m gpu -rA -k "splitfilt or naca or descriptor or filter or needs_face" >> gpurun_out/r02aa_pytest_2gpu.log 2>&1
grep "PASSED\|FAILED\|passed\|failed" gpurun_out/r02aa_pytest_2gpu.log | cut -c1-160
