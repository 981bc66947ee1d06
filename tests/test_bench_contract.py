"""bench.py's output contract: exactly ONE JSON line on stdout (library chatter goes to stderr), with the keys the driver
reads.  The reference arm runs anywhere (it times the oracle's CPU port, the only thing bench.py may execute from oracle/);
the default arm needs a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, timeout=timeout, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line():
    d = _run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0", "--cpu-seconds", "1"], 600)
    assert BASE_KEYS <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "mcts_simulations_per_sec" and d["unit"] == "sims/s"
    assert d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]


@pytest.mark.gpu
def test_default_arm_prints_one_json_line_with_rooflines():
    d = _run(["--gpus", "1", "--steps", "2", "--warmup", "3", "--trees", "2048", "--cpu-seconds", "1"], 900)
    assert BASE_KEYS | {"roofline", "clocks"} <= set(d)
    assert d["metric"] == "mcts_simulations_per_sec" and d["value"] > 1e6 and d["gpu_launches"] > 0
    rf = d["roofline"]
    assert rf["bound"] in ("hbm", "tensor") and 0 < rf["frac"] <= 1.05 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert d["overflow"] == 0
