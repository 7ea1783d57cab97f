"""bench.py's output contract, checked on the arm that runs without a GPU (`--impl reference` = the CPU restatement):
exactly one JSON line on stdout, every key the driver reads, the reference-arm conventions (impl, zero-byte e2e that
repeats the value, cpu_baseline describing this run), and a non-rank-0 process that exits quietly under torchrun."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                           "--entities", "40000", "--subscribers", "4000"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)


def test_reference_arm_prints_one_contract_line():
    out = _run()
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "subscriber-AOI-queries/s" and d["unit"] == "queries/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["data"] == "synthetic"
    assert "workload" in d["config"] and d["config"]["entities"] == 40000 and d["config"]["subscribers"] == 4000
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # value = subscribers x steps / time
    assert abs(d["value"] - 4000 / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6


def test_reference_arm_other_ranks_stay_silent():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip() == ""
