#!/bin/bash
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
CHD_EXPERIMENT_LIB=tools/_bin/libchd_nofuse.so tools/run_bench.sh var_nofuse --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 4 --expanded-steps 0 2>&1 | head -3
tools/run_bench.sh v8a --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | head -6
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2/bench_v8a.json"))
print("e2e", json.dumps({k: v for k, v in d["e2e"].items() if k in ("value", "ms_per_step", "sync_fetch")}), d.get("host"))
PY
tools/run_bench.sh c3 --config 10m --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 --expanded-steps 0 2>&1 | head -4
tools/run_bench.sh c5 --config handover --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 --expanded-steps 0 2>&1 | head -4
python tools/pcie_probe.py 2>&1 | tail -2
taskset -c 0-31 python tools/pcie_probe.py 2>&1 | tail -1
taskset -c 32-63 python tools/pcie_probe.py 2>&1 | tail -1
