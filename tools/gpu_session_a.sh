#!/bin/bash
# one GPU session: tests, store-pattern probe, emit tile-shape variants, ncu capture of the emit kernel
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
tools/_bin/write_probe > gpurun_out/r2/write_probe2.json 2> gpurun_out/r2/write_probe2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2/write_probe2.json"))
for k, v in d.items():
    if isinstance(v, dict) and ("ticket" in k or "desc" in k or k in ("copy_grid_OP_CS_r4", "copy_grid_OP_CS_r2", "fill_grid_OP_WB_r4")):
        print("%-32s %.4f ms %6.0f GB/s" % (k, v["ms"], v["gbs"]))
PY
for v in r8_b4 r8_b3 r2_b8 r4_b4 r4_b8; do
  CHD_EXPERIMENT_LIB=tools/_bin/libchd_$v.so tools/run_bench.sh var_$v --steps 60 --warmup 5 --no-cpu-baseline --no-gate --e2e-steps 4 --expanded-steps 0 2>&1 | head -3
done
tools/run_bench.sh v6c --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | head -6
timeout 600 ncu --set full --clock-control none --import-source on -k regex:emit_visible_kernel -c 2 -f -o gpurun_out/r2/emit_v6c python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gate --e2e-steps 2 --expanded-steps 0 > gpurun_out/r2/ncu_emit.log 2>&1
tail -3 gpurun_out/r2/ncu_emit.log
ls -la gpurun_out/r2/ | tail -5
