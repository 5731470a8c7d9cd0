"""bench.py's reference arm (the CPU restatement timed on the host cores) honours the JSON-line contract; runs on the
CPU with a tiny sample.  Under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--impl", "reference", "--steps", "3", "--warmup", "1", "--cpu-instances", "2", "--cpu-steps", "8", "--prefill", "40"]


def _run(extra_env=None):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + SMALL, capture_output=True, text=True,
                          env=env, timeout=600)


def test_reference_arm_json_line():
    out = _run()
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert key in line, key
    assert line["unit"] == "instance-steps/s" and line["higher_is_better"] is True and line["dtype"] == "f64"
    assert line["config"]["workload"] == "rm3_irregular_ensemble"
    assert line["value"] > 0 and line["cpu_baseline"]["value"] == line["value"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip() == ""
