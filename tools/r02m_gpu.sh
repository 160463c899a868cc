#!/bin/bash
# r02m (2 GPUs): NCCL side-exchange parity tests + 2-GPU bench with the exposed exchange waits
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02m_pytest_2gpu.log
timeout 1200 python -m pytest tests/test_multi_rank.py -q -m gpu -rA >> gpurun_out/r02m_pytest_2gpu.log 2>&1
tail -15 gpurun_out/r02m_pytest_2gpu.log
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02m_bench2.json 2> gpurun_out/r02m_bench2.err
grep -c "NCCL INFO" gpurun_out/r02m_bench2.err
tail -2 gpurun_out/r02m_bench2.err
cat gpurun_out/r02m_bench2.json
