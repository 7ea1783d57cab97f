#!/bin/bash
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
tools/_bin/write_probe > gpurun_out/r2/write_probe3.json 2> gpurun_out/r2/write_probe3.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2/write_probe3.json"))
for k, v in d.items():
    if isinstance(v, dict) and ("shift" in k or k in ("copy_grid_OP_CS_r4", "copy_grid_desc_OP_CS_r4")):
        print("%-32s %.4f ms %6.0f GB/s" % (k, v["ms"], v["gbs"]))
PY
tools/run_bench.sh v6e --steps 100 --warmup 5 2>&1 | head -8
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2/bench_v6e.json"))
print("e2e", json.dumps({k: v for k, v in d["e2e"].items() if k in ("value", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step", "sync_fetch", "serial")}))
PY
