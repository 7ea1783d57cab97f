#!/bin/bash
mkdir -p gpurun_out/r2
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/launches_v6.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gate --e2e-steps 2 --expanded-steps 0 > gpurun_out/r2/ncu_launch.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2/launches_v6.csv")) if len(r) > 5 and r[0].isdigit()]
# columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section Name, Metric Name, Metric Unit, Metric Value
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0]
    v = float(r[-1].replace(",", ""))
    unit = r[-2]
    if unit == "ns": v /= 1000.0
    elif unit == "ms": v *= 1000.0
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-60s n=%4d total %9.1f us  avg %8.2f us  %5.1f %%" % (k[:60], n, t, t / n, 100 * t / tot))
PY
