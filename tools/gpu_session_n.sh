#!/bin/bash
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
tools/run_bench.sh v13a --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | head -6
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2/bench_v13a.json"))
print("e2e", json.dumps({k: v for k, v in d["e2e"].items() if k in ("value", "ms_per_step", "sync_fetch", "f64_upload", "host_ms_per_step")}), d.get("host"), "host_enq", d.get("host_enqueue_ms_per_step"))
PY
CHD_EXPERIMENT_LIB=$PWD/tools/_bin/libchd_mb6.so tools/run_bench.sh v13a_mb6 --steps 100 --warmup 5 --no-cpu-baseline --no-gate --e2e-steps 4 --expanded-steps 0 2>&1 | head -4
tools/run_bench.sh c5 --config handover --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 4 --expanded-steps 0 2>&1 | head -4
tools/run_bench.sh c3 --config 10m --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 4 --expanded-steps 0 2>&1 | head -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"emit_visible|fanout" -c 40 --csv --log-file gpurun_out/r2/launches_v13.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gate --e2e-steps 2 --expanded-steps 0 > gpurun_out/r2/ncu_v13.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2/launches_v13.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0]; v = float(r[-1].replace(",", "")); unit = r[-2]
    if unit == "ns": v /= 1000.0
    elif unit == "ms": v *= 1000.0
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, t) in agg.items(): print("%-50s n=%3d avg %8.2f us" % (k[:50], n, t / n))
PY
