#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget was spent, then the usual evidence.
#   gpurun --timeout 1500 -- 'bash tools/next_round_gpu.sh'
# Outputs land in gpurun_out/ (scratch); copy what is to be judged into profiles/.
set -x
mkdir -p gpurun_out
# 1. the unrun parity file first (MO/CH/IR formulations, Average/CH/IR solvers, cut-off overintegration, solution filter)
timeout 900 python -m pytest tests/test_zz_gpu_split_formulations.py -q -m gpu -x > gpurun_out/r02a_zz.log 2>&1
tail -5 gpurun_out/r02a_zz.log
# 2. the whole GPU suite
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02a_pytest.log 2>&1
tail -3 gpurun_out/r02a_pytest.log
# 3. bench (own arm, default flags) + launch list of a short run
python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/r02a_launch_bench.log 2>&1
cat gpurun_out/r02a_bench.json
