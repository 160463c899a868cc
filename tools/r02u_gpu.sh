#!/bin/bash
# r02u: viscosity-law GPU cases, smoke, quick timing of the release library, ncu launch list + full capture of the stage kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gpu_split_formulations.py tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -4 > gpurun_out/r02u.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 >> gpurun_out/r02u.txt
python tools/ovl_timing.py 48 2>&1 | tail -1 >> gpurun_out/r02u.txt
cat gpurun_out/r02u.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02u_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fullsize-oracle --parity-nel 0 > gpurun_out/r02u_launch_bench.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:"k_volume_gl2|k_flux|k_update" --launch-skip 30 --launch-count 3 -f -o gpurun_out/r02u_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fullsize-oracle --parity-nel 0 > gpurun_out/r02u_ncu.log 2>&1
ls -la gpurun_out/r02u*
