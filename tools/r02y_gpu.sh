#!/bin/bash
# r02y (4 GPUs): NCCL parity cases on 4 ranks (block partition 2x2x1), TGV weak scaling at 4, naca (configs[3]) at 2 and 4 GPUs
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02y_pytest_4gpu.log
timeout 600 python -m pytest tests/test_zz_gpu_split_formulations.py -q -m gpu -k "conservative or overintegration" 2>&1 | tail -2 >> gpurun_out/r02y_pytest_4gpu.log
timeout 1500 python -m pytest tests/test_multi_rank.py -q -m gpu -rA >> gpurun_out/r02y_pytest_4gpu.log 2>&1
tail -14 gpurun_out/r02y_pytest_4gpu.log | cut -c1-160
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 > gpurun_out/r02y_bench4.json 2> gpurun_out/r02y_bench4.err
grep "bench.py:" gpurun_out/r02y_bench4.err | cut -c1-400; grep -c "NCCL INFO" gpurun_out/r02y_bench4.err; grep -m3 "NVLS\|NVLink\|via P2P" gpurun_out/r02y_bench4.err | cut -c1-200; cat gpurun_out/r02y_bench4.json
CUDA_VISIBLE_DEVICES=0,1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --workload naca --gpus 2 > gpurun_out/r02y_naca2.json 2> gpurun_out/r02y_naca2.err
grep "bench.py:" gpurun_out/r02y_naca2.err | cut -c1-400; cat gpurun_out/r02y_naca2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --workload naca --gpus 4 > gpurun_out/r02y_naca4.json 2> gpurun_out/r02y_naca4.err
grep "bench.py:" gpurun_out/r02y_naca4.err | cut -c1-400; cat gpurun_out/r02y_naca4.json
head -c 20000 gpurun_out/r02y_bench4.err > gpurun_out/r02y_bench4_head.err; rm -f gpurun_out/r02y_bench4.err
