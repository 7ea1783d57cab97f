#!/bin/bash
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
tools/run_bench.sh v10a --steps 100 --warmup 5 2>&1 | head -12
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2/bench_v10a.json"))
print("e2e", json.dumps({k: v for k, v in d["e2e"].items() if k in ("value", "ms_per_step", "sync_fetch", "f64_upload", "serial", "h2d_bytes_per_step", "d2h_bytes_per_step")}), d.get("host"))
PY
