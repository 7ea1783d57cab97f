#!/bin/bash
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
tools/run_bench.sh v9a --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | head -6
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2/bench_v9a.json"))
print("e2e", json.dumps({k: v for k, v in d["e2e"].items() if k in ("value", "ms_per_step", "sync_fetch")}), d.get("host"))
PY
cat /sys/devices/system/node/node*/cpulist
