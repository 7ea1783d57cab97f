#!/bin/bash
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_robustness.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -6
for v in T8_c5 T16_c5 T64_c5 T32_c4 T32_c6; do
  CHD_EXPERIMENT_LIB=tools/_bin/libchd_$v.so tools/run_bench.sh var_$v --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 4 --expanded-steps 0 2>&1 | head -3
done
tools/run_bench.sh v7c --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | head -6
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2/bench_v7c.json"))
print("e2e", json.dumps({k: v for k, v in d["e2e"].items() if k in ("value", "ms_per_step", "sync_fetch")}), d.get("host"))
PY
tools/run_bench.sh c3 --config 10m --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 --expanded-steps 0 2>&1 | head -4
tools/run_bench.sh c5 --config handover --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 --expanded-steps 0 2>&1 | head -4
