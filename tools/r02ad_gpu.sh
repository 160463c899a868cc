#!/bin/bash
# r02ad (1 GPU): the whole GPU suite at the final state of round 2 (incl. the reference's cavity_3D golden state through the C ABI)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02ad_pytest.log
timeout 1500 python -m pytest tests -q -m gpu -rA --durations=15 >> gpurun_out/r02ad_pytest.log 2>&1
grep "FAILED\|ERROR\|passed\|failed\|Error" gpurun_out/r02ad_pytest.log | cut -c1-220 | tail -30
