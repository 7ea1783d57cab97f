#!/bin/bash
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_robustness.py -m gpu -x -q 2>&1 | tail -6
for v in p4_c6 p4_c4 p3_c5; do
  CHD_EXPERIMENT_LIB=tools/_bin/libchd_$v.so tools/run_bench.sh var_$v --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 4 --expanded-steps 0 2>&1 | head -3
done
tools/run_bench.sh v7b --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | head -6
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2/bench_v7b.json"))
print("e2e", json.dumps({k: v for k, v in d["e2e"].items() if k in ("value", "ms_per_step", "sync_fetch")}))
PY
