#!/bin/bash
# r02w: full GPU suite, bench, ncu launch list and one full capture per stage kernel (release library)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu > gpurun_out/r02w_pytest.log 2>&1
tail -3 gpurun_out/r02w_pytest.log
python bench.py > gpurun_out/r02w_bench.json 2> gpurun_out/r02w_bench.err
tail -2 gpurun_out/r02w_bench.err | cut -c1-300; cat gpurun_out/r02w_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02w_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fullsize-oracle --parity-nel 0 > gpurun_out/r02w_launch_bench.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:k_volume_gl2 --launch-skip 12 --launch-count 1 -f -o gpurun_out/r02w_vol python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fullsize-oracle --parity-nel 0 > gpurun_out/r02w_ncu.log 2>&1
ncu --set full --clock-control none --kernel-name regex:"k_flux|k_update" --launch-skip 24 --launch-count 2 -f -o gpurun_out/r02w_fu python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fullsize-oracle --parity-nel 0 >> gpurun_out/r02w_ncu.log 2>&1
ls -la gpurun_out/
