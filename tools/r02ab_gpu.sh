#!/bin/bash
# r02ab (1 GPU): LAF filter cases, N = 11 on CTA pairs (operator parity + the reference's tgv/oInt rows), filter regression cases
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02ab_pytest.log
timeout 900 python -m pytest tests/test_gpu_n11.py tests/test_zz_gpu_split_formulations.py -q -m gpu -rA --durations=10 -k "laf" >> gpurun_out/r02ab_pytest.log 2>&1
grep "PASSED\|FAILED\|ERROR\|passed\|failed\|Error\|assert" gpurun_out/r02ab_pytest.log | cut -c1-220 | tail -40
