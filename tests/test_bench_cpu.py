"""bench.py contract, the part that runs without a GPU: the `--impl reference` arm (the CPU oracle timed on this host)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip().splitlines()


def test_reference_arm_prints_one_contract_line():
    lines = _run(["--impl", "reference", "--steps", "1", "--warmup", "3"])
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("frames/s YOLOv9-c") and d["unit"] == "frames/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["cpu_baseline"]["value"] - d["value"]) < 1e-9
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "BASELINE configs[1]" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_without_work():
    # under torchrun only rank 0 runs the CPU arm; the others print nothing and exit 0
    lines = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "3"],
                 env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29533"})
    assert lines == []
