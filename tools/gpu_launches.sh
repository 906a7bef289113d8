#!/bin/bash
# per-launch device times of one short bench run (cold-cache, serialised: compare shares)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline $@"
$B > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu.log 2>&1
echo "rc=$?"
python - <<'PY'
import csv, collections, re
rows = [r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r) > 10 and r[0].isdigit()]
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
