#!/bin/bash
# r02q (2 GPUs): full GPU suite incl. sponge + naca, multi-rank NCCL suite, naca bench at 1 and 2 GPUs
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02q_pytest_2gpu.log
timeout 2000 python -m pytest tests -q -m gpu -rA >> gpurun_out/r02q_pytest_2gpu.log 2>&1
tail -5 gpurun_out/r02q_pytest_2gpu.log; grep -c PASSED gpurun_out/r02q_pytest_2gpu.log; grep "FAILED\|ERROR" gpurun_out/r02q_pytest_2gpu.log | head
python bench.py --workload naca --no-cpu-baseline > gpurun_out/r02q_naca1.json 2> gpurun_out/r02q_naca1.err
tail -2 gpurun_out/r02q_naca1.err | cut -c1-600; cat gpurun_out/r02q_naca1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload naca --gpus 2 > gpurun_out/r02q_naca2.json 2> gpurun_out/r02q_naca2.err
grep "bench.py:" gpurun_out/r02q_naca2.err | cut -c1-600; cat gpurun_out/r02q_naca2.json
