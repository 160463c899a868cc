#!/bin/bash
# r02ai (2 GPUs): all NCCL side-exchange parity cases with the final library of round 2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02ai_pytest_2gpu.log
timeout 900 python -m pytest tests/test_multi_rank.py -q -m gpu -rA >> gpurun_out/r02ai_pytest_2gpu.log 2>&1
grep "PASSED\|FAILED\|ERROR\|passed\|failed" gpurun_out/r02ai_pytest_2gpu.log | cut -c1-160 | tail -16
