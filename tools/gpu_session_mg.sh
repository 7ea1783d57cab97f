#!/bin/bash
# multi-GPU session: usage tools/gpu_session_mg.sh <N> [skip-parity]
N=${1:-2}
mkdir -p gpurun_out/r2
nvidia-smi -L | head -8
if [ -z "$2" ]; then
CHD_PARITY_TICKS=${CHD_PARITY_TICKS:-6} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/run_multigpu_parity.py > gpurun_out/r2/mg_parity_n$N.txt 2>&1
echo "parity rc=$?"; grep -E "parity|mismatches=[1-9]|Error|error|exchange mode|collectives in" gpurun_out/r2/mg_parity_n$N.txt | tail -14; tail -3 gpurun_out/r2/mg_parity_n$N.txt
fi
run() {  # tag, extra args
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2/bench_${tag}_n${N}.json 2> gpurun_out/r2/bench_${tag}_n${N}.err
  echo "$tag rc=$?"; tail -2 gpurun_out/r2/bench_${tag}_n${N}.err | cut -c1-300
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2/bench_${tag}_n${N}.json"))
    print("${tag} N=${N} value %.1f M/s ms %.4f e2e %.1f M/s gate %s\n  stage %s\n  timeline %s\n  coll %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["parity_gate"]["mismatches"], {k: round(v, 4) for k, v in d["stage_ms"].items()}, d["last_tick_timeline_ms"], d.get("collectives")))
except Exception as ex:
    print("no bench json", ex)
PY
}
run peer --exchange peer
if [ "$N" = "2" ]; then run nccl --exchange nccl; fi
run weak --exchange peer --scaling weak
