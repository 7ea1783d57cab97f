#!/bin/bash
# profiling session of the final tree: launch list of a tick, one --set full capture of the emit kernel, sanitizer passes
mkdir -p gpurun_out/r2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/launches_final.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gate --e2e-steps 2 --expanded-steps 0 > gpurun_out/r2/ncu_launch.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2/launches_final.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0]
    v = float(r[-1].replace(",", ""))
    unit = r[-2]
    if unit == "ns": v /= 1000.0
    elif unit == "ms": v *= 1000.0
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
with open("gpurun_out/r2/launches_final_summary.txt", "w") as f:
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        line = "%-60s n=%4d total %9.1f us  avg %8.2f us  %5.1f %%" % (k[:60], n, t, t / n, 100 * t / tot)
        print(line); f.write(line + "\n")
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:emit_visible_kernel -s 3 -c 1 -f -o gpurun_out/r2/emit_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gate --e2e-steps 2 --expanded-steps 0 > gpurun_out/r2/ncu_full.log 2>&1
ncu -i gpurun_out/r2/emit_full.ncu-rep --page details > gpurun_out/r2/emit_full_details.txt 2>&1
grep -E "dram__bytes_(read|write).sum |gpu__time_duration.sum|Duration|DRAM Throughput|Memory Throughput|Registers Per|Achieved Occupancy|Theoretical Occupancy|L2 Cache Throughput|Issue Slots Busy" gpurun_out/r2/emit_full_details.txt | head -20
ncu -i gpurun_out/r2/emit_full.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum > gpurun_out/r2/emit_full_raw.csv 2>&1
tail -3 gpurun_out/r2/emit_full_raw.csv
# sanitizer over the tests that drive the kernels added this round
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_robustness.py tests/test_gpu_payload.py -m gpu -x -q > gpurun_out/r2/memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/r2/memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_robustness.py tests/test_gpu_payload.py -m gpu -x -q > gpurun_out/r2/racecheck.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/r2/racecheck.log
