#!/bin/bash
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for v in t1_b8 t2_b4 t2_b3 t1_b5; do
  CHD_EXPERIMENT_LIB=tools/_bin/libchd_$v.so tools/run_bench.sh var_$v --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 4 --expanded-steps 0 2>&1 | head -3
done
tools/run_bench.sh v6d --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | head -6
