#!/bin/bash
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
tools/run_bench.sh v14a --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | head -6
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2/bench_v14a.json"))
print("e2e", json.dumps({k: v for k, v in d["e2e"].items() if k in ("value", "ms_per_step", "sync_fetch", "f64_upload", "host_ms_per_step")}), d.get("host"), "host_enq", d.get("host_enqueue_ms_per_step"))
PY
tools/run_bench.sh c5 --config handover --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 4 --expanded-steps 0 2>&1 | head -4
