#!/bin/bash
# r02ag: generic Navier-Stokes element kernel (Gauss nodes, configs[0] flavour) with one vs two resident elements per SM at N = 7
mkdir -p gpurun_out
echo "release (1 CTA/SM by registers)" > gpurun_out/r02ag_minb.log
python tools/psweep.py --only=7 --flavour=Gauss >> gpurun_out/r02ag_minb.log 2>&1
echo "dev FX_KV_MINB=2" >> gpurun_out/r02ag_minb.log
FLEXI_GPU_LIB=$PWD/flexi_b200/lib/libflexi_gpu_dev_mb.so python tools/psweep.py --only=7 --flavour=Gauss >> gpurun_out/r02ag_minb.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r02ag_minb.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['flavour'], round(d['GDOF_s'],3), {k:round(v,3) for k,v in d['kernel_ms_per_stage'].items()}, round(d['stage_roofline_frac'],3))
    else: print(l.strip()[:200])
PY
