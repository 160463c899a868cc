#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/psweep.py gpurun_out/r02o_psweep.json > gpurun_out/r02o_psweep.log 2>&1
tail -30 gpurun_out/r02o_psweep.log | cut -c1-260
