#!/bin/bash
# usage (on the GPU box): tools/run_bench.sh <tag> [bench args...]   -> gpurun_out/r2/bench_<tag>.json/.err + a one-line digest
mkdir -p gpurun_out/r2
tag=$1; shift
python bench.py "$@" > gpurun_out/r2/bench_$tag.json 2> gpurun_out/r2/bench_$tag.err || { echo "bench failed"; tail -20 gpurun_out/r2/bench_$tag.err; exit 1; }
python - <<PY
import json
d = json.load(open("gpurun_out/r2/bench_$tag.json"))
print("$tag", "value %.1f M/s" % (d["value"] / 1e6), "ms/step %.4f" % d["ms_per_step"], "e2e %.1f M/s" % (d["e2e"]["value"] / 1e6),
      "launches", d.get("gpu_launches"), "\nstage", {k: round(v, 4) for k, v in d.get("stage_ms", {}).items()}, "\ntimeline", d.get("last_tick_timeline_ms"),
      "\nroofline", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.get("roofline", {}).items() if k in ("achieved", "frac", "kernel_ms", "l2_side_gbs", "write_only_peak_gbs_this_run", "frac_of_write_only_peak")},
      "\ngate", d.get("parity_gate"), "\ncpu", d.get("cpu_baseline"))
PY
