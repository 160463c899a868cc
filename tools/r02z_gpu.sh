#!/bin/bash
# r02z: ABI 2 (Ja_Face / Ja_slave in the descriptor): full GPU suite, smoke, bench
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu > gpurun_out/r02z_pytest.log 2>&1
tail -3 gpurun_out/r02z_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --no-cpu-baseline > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err
tail -2 gpurun_out/r02z_bench.err | cut -c1-300; cut -c1-700 gpurun_out/r02z_bench.json
