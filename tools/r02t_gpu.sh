#!/bin/bash
# r02t: full GPU suite + bench (TGV, naca) of the release library
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu > gpurun_out/r02t_pytest.log 2>&1
tail -4 gpurun_out/r02t_pytest.log
python bench.py > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err
tail -2 gpurun_out/r02t_bench.err | cut -c1-400; cat gpurun_out/r02t_bench.json
python bench.py --workload naca > gpurun_out/r02t_naca1.json 2> gpurun_out/r02t_naca1.err
tail -2 gpurun_out/r02t_naca1.err | cut -c1-600; cat gpurun_out/r02t_naca1.json
