#!/bin/bash
# per-launch device times of one bench_configs.py configuration (args: config names)
mkdir -p gpurun_out
C="python tools/bench_configs.py $@"
$C > gpurun_out/plain_cfg.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_cfg.csv $C > gpurun_out/ncu_cfg.log 2>&1
echo "rc=$?"; cat gpurun_out/plain_cfg.log | cut -c1-400
python - <<'PY'
import csv, collections, re
rows = [r for r in csv.reader(open("gpurun_out/launches_cfg.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r[4]); val = float(r[-1].replace(",", ""))
    unit = r[-2]
    if unit == "us": val /= 1e3
    elif unit == "ns": val /= 1e6
    agg[name][0] += 1; agg[name][1] += val
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.3f} ms {100*v[1]/tot:5.1f}%  n={v[0]:4d}  avg {v[1]/v[0]*1e3:9.1f} us  {k[:90]}")
PY
