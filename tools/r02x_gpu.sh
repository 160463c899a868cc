#!/bin/bash
# r02x: final state of round 2 on one B200: full GPU suite, bench (TGV, naca), ncu launch list, full captures digested on the box
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu > gpurun_out/r02x_pytest.log 2>&1
tail -3 gpurun_out/r02x_pytest.log
python bench.py > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err
tail -2 gpurun_out/r02x_bench.err | cut -c1-300; cat gpurun_out/r02x_bench.json
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02x_bench_ref.json 2> gpurun_out/r02x_bench_ref.err
cat gpurun_out/r02x_bench_ref.json | cut -c1-600
python bench.py --workload naca --no-cpu-baseline > gpurun_out/r02x_naca1.json 2> gpurun_out/r02x_naca1.err
cat gpurun_out/r02x_naca1.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02x_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fullsize-oracle --parity-nel 0 > gpurun_out/r02x_launch_bench.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:"k_volume_gl2|k_flux|k_update" --launch-skip 30 --launch-count 3 -f -o /tmp/r02x_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fullsize-oracle --parity-nel 0 > gpurun_out/r02x_ncu.log 2>&1
python profiles/ncu_digest.py /tmp/r02x_full.ncu-rep k_volume_gl2 > gpurun_out/r02x_ncu_full.txt 2>&1
python profiles/ncu_phases.py /tmp/r02x_full.ncu-rep k_volume_gl2 >> gpurun_out/r02x_ncu_full.txt 2>&1
tail -12 gpurun_out/r02x_ncu_full.txt | cut -c1-200
ls -la gpurun_out/ /tmp/r02x_full.ncu-rep
