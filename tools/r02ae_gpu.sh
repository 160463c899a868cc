#!/bin/bash
# r02ae: smoke + default bench (with cpu_baseline) at the final state of round 2
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r02ae_bench.json 2> gpurun_out/r02ae_bench.err
tail -2 gpurun_out/r02ae_bench.err | cut -c1-300; cut -c1-900 gpurun_out/r02ae_bench.json
