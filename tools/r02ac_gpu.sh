#!/bin/bash
# r02ac (2 GPUs): the reference's NACA0012 mesh file + state on the CUDA path (1 rank and NCCL), then the whole GPU suite on 2 GPUs
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02ac_pytest.log
timeout 600 python -m pytest tests/test_gpu_hopr.py tests/test_multi_rank.py -q -m gpu -rA -k "hopr or reference_naca" >> gpurun_out/r02ac_pytest.log 2>&1
grep "PASSED\|FAILED\|ERROR\|passed\|failed\|Error\|assert" gpurun_out/r02ac_pytest.log | cut -c1-220 | tail -30
