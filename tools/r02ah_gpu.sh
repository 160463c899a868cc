#!/bin/bash
# r02ah: final library of round 2 on one B200: whole GPU suite, smoke, ncu launch list of a short bench run
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02ah_pytest.log
timeout 1500 python -m pytest tests -q -m gpu -rA --durations=10 >> gpurun_out/r02ah_pytest.log 2>&1
grep "FAILED\|ERROR\|passed\|failed\|Error" gpurun_out/r02ah_pytest.log | cut -c1-220 | tail -10
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02ah_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fullsize-oracle --parity-nel 0 > gpurun_out/r02ah_launch_bench.log 2>&1
tail -2 gpurun_out/r02ah_launch_bench.log | cut -c1-300
python - <<'PY'
import csv, collections
t=collections.defaultdict(float); c=collections.Counter()
rows=[r for r in csv.reader(l for l in open('gpurun_out/r02ah_launches.csv') if l.startswith('"'))]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
for r in rows[1:]:
    n=r[ki].split('<')[0].split('(')[0]; t[n]+=float(r[vi].replace(',','')); c[n]+=1
tot=sum(t.values())
for n,v in sorted(t.items(), key=lambda x:-x[1])[:8]: print(f"{n:28s} {c[n]:4d} launches {v/1e6:9.2f} ms {100*v/tot:5.1f} %")
PY
