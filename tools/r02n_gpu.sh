#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r02n.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/mr_debug.py 7 6 5 48 2>&1 | grep "rank " >> gpurun_out/r02n.txt
cat gpurun_out/r02n.txt
