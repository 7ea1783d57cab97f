#!/bin/bash
# multi-GPU session: usage tools/gpu_session_mg.sh <N>
N=${1:-2}
mkdir -p gpurun_out/r2
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/run_multigpu_parity.py > gpurun_out/r2/mg_parity_n$N.txt 2>&1
echo "parity rc=$?"; grep -E "parity|mismatches=[1-9]|Error|error" gpurun_out/r2/mg_parity_n$N.txt | tail -8; tail -3 gpurun_out/r2/mg_parity_n$N.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r2/bench_n${N}.json 2> gpurun_out/r2/bench_n${N}.err
echo "bench rc=$?"; tail -3 gpurun_out/r2/bench_n${N}.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2/bench_n${N}.json"))
    print("N=${N} value %.1f M/s ms %.4f e2e %.1f M/s gate %s stage %s timeline %s coll %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["parity_gate"], d["stage_ms"], d["last_tick_timeline_ms"], d.get("collectives")))
except Exception as ex:
    print("no bench json", ex)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 60 --warmup 5 --scaling weak --no-cpu-baseline > gpurun_out/r2/bench_weak_n${N}.json 2> gpurun_out/r2/bench_weak_n${N}.err
echo "weak rc=$?"; tail -2 gpurun_out/r2/bench_weak_n${N}.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2/bench_weak_n${N}.json"))
    print("weak N=${N} value %.1f M/s ms %.4f e2e %.1f M/s gate %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["parity_gate"]))
except Exception as ex:
    print("no weak json", ex)
PY
