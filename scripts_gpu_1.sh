#!/bin/bash
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader || { echo "GPU UNHEALTHY after $1"; exit 7; }; }
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/g1_pytest.log; tail -5 gpurun_out/g1_pytest.log; health pytest
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench1d.json; health bench
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench1d.json')); print(round(d["value"]), d["ms_per_step"], "host_enqueue_ms", d.get("host_enqueue_ms_per_step"), "launches", d["gpu_launches"], "e2e", d["e2e"]["value"], d["clocks"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 -c 260 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_launch_b.log 2>&1; health ncu
python - <<'PY'
import csv, collections, re
rows=[r for r in csv.DictReader(l for l in open('gpurun_out/launches_r1b.csv') if l.startswith('"'))]
names=[r['Kernel Name'] for r in rows]
idx=[i for i,n in enumerate(names) if 'tbe_pooled_fwd' in n]
a,b=idx[0],idx[1]
agg=collections.OrderedDict(); tot=0
for r in rows[a:b]:
    k=re.sub(r'\(.*','',re.sub(r'^void ','',r['Kernel Name']).replace('<unnamed>::',''))[:70]; d=float(r['Metric Value'])/1e3
    agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=d; tot+=d
print("kernel us per step", tot, "launches", b-a)
for k,(c,d) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:18]: print(f"{d:8.1f} {c:3d} {k}")
PY
