#!/bin/bash
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
tools/run_bench.sh v12a --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | head -8
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2/bench_v12a.json"))
print("e2e", json.dumps({k: v for k, v in d["e2e"].items() if k in ("value", "ms_per_step", "sync_fetch", "f64_upload", "serial", "h2d_bytes_per_step", "d2h_bytes_per_step")}), d.get("host"))
PY
python tools/pcie_probe.py 2>&1 | tail -3
