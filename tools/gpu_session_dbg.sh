#!/bin/bash
# multi-GPU debug: bench shapes that produce 1-column slabs / two-pass builds with a halo, on N GPUs
N=${1:-2}
mkdir -p gpurun_out/r2
try() {
  tag=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 4 --expanded-steps 0 "$@" > gpurun_out/r2/dbg_${tag}.json 2> gpurun_out/r2/dbg_${tag}.err
  rc=$?
  echo "$tag rc=$rc $(head -c 200 gpurun_out/r2/dbg_${tag}.json)"
  if [ $rc -ne 0 ]; then
    grep -h "ChdError\|SystemExit\|Error" gpurun_out/r2/dbg_${tag}.err | sort | uniq -c | head -6 | cut -c1-400
    CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 4 --expanded-steps 0 "$@" > /dev/null 2> gpurun_out/r2/dbg_${tag}_blocking.err
    echo "  with CUDA_LAUNCH_BLOCKING=1:"; grep -h "ChdError\|SystemExit" gpurun_out/r2/dbg_${tag}_blocking.err | sort | uniq -c | head -6 | cut -c1-500
  fi
}
try 2x2 --config 2x2 --entities 1000000 --subscribers 100000
try c3s --config 10m --entities 2000000 --subscribers 200000
try c5s --config handover --entities 2000000 --subscribers 200000
try c3s_weak --config 10m --entities 1000000 --subscribers 100000 --scaling weak
try bench_small --entities 400000 --subscribers 20000
