#!/bin/bash
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_robustness.py tests/test_gpu_payload.py -m gpu -x -q 2>&1 | tail -12
for v in p0 p3 p5 p6; do
  CHD_EXPERIMENT_LIB=tools/_bin/libchd_$v.so tools/run_bench.sh var_$v --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 4 --expanded-steps 0 2>&1 | head -3
done
tools/run_bench.sh v7a --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | head -6
