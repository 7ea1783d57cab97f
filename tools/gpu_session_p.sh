#!/bin/bash
mkdir -p gpurun_out/r2
for v in s30 s36; do
CHD_EXPERIMENT_LIB=$PWD/tools/_bin/libchd_$v.so tools/run_bench.sh v15_$v --steps 100 --warmup 5 --no-cpu-baseline --no-gate --e2e-steps 4 --expanded-steps 0 2>&1 | head -3
done
tools/run_bench.sh v15_base --steps 100 --warmup 5 --no-cpu-baseline --no-gate --expanded-steps 0 --trace-e2e 2>&1 | head -3
grep "\[bench\]" gpurun_out/r2/bench_v15_base.err
