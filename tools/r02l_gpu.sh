#!/bin/bash
# r02l: full GPU suite + bench of the v3 element kernel (release library)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu -x > gpurun_out/r02l_pytest.log 2>&1
tail -4 gpurun_out/r02l_pytest.log
python bench.py > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err
tail -3 gpurun_out/r02l_bench.err
cat gpurun_out/r02l_bench.json
