#!/bin/bash
# r02af: LSERK fold (element kernel writes -RKA*Ut_tmp - sJ*VolInt into the LSERK register): whole GPU suite, smoke, bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02af_pytest.log
timeout 1500 python -m pytest tests -q -m gpu -rA >> gpurun_out/r02af_pytest.log 2>&1
grep "FAILED\|ERROR\|passed\|failed\|Error" gpurun_out/r02af_pytest.log | cut -c1-220 | tail -20
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --no-cpu-baseline > gpurun_out/r02af_bench.json 2> gpurun_out/r02af_bench.err
tail -2 gpurun_out/r02af_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02af_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['stage']['frac'], {k:round(v['avg_ms'],3) for k,v in d['roofline']['kernels'].items()}, d['parity'])
PY
