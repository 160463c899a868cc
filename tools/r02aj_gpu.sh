#!/bin/bash
# r02aj (2 GPUs): N = 11 (CTA pairs) on a partitioned run with the NCCL side exchange
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02aj_pytest_2gpu.log
timeout 400 python -m pytest tests/test_multi_rank.py -q -m gpu -rA -k "nccl and n11" >> gpurun_out/r02aj_pytest_2gpu.log 2>&1
grep "PASSED\|FAILED\|ERROR\|passed\|failed\|Error\|assert" gpurun_out/r02aj_pytest_2gpu.log | cut -c1-200 | tail -12
