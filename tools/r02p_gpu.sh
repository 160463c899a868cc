#!/bin/bash
# r02p (2 GPUs): naca GPU tests, full multi-rank NCCL tests, TGV bench at 2 GPUs (exposed waits), naca bench at 1 and 2 GPUs
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02p_pytest_2gpu.log
timeout 1500 python -m pytest tests/test_naca_mesh.py tests/test_multi_rank.py -q -m gpu -rA >> gpurun_out/r02p_pytest_2gpu.log 2>&1
tail -18 gpurun_out/r02p_pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02p_bench2.json 2> gpurun_out/r02p_bench2.err
tail -2 gpurun_out/r02p_bench2.err; cat gpurun_out/r02p_bench2.json
python bench.py --workload naca --no-cpu-baseline > gpurun_out/r02p_naca1.json 2> gpurun_out/r02p_naca1.err
tail -2 gpurun_out/r02p_naca1.err; cat gpurun_out/r02p_naca1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload naca --gpus 2 > gpurun_out/r02p_naca2.json 2> gpurun_out/r02p_naca2.err
tail -2 gpurun_out/r02p_naca2.err; cat gpurun_out/r02p_naca2.json
