"""CPU checks of bench.py's output contract: the committed bench lines carry every key the driver reads, their
derived fields are consistent, and the reference arm is silent on non-zero ranks."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"}


def _lines():
    out = []
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r01[v-z]_bench*.json"))):
        txt = open(f).read().strip().splitlines()[-1]
        out.append((os.path.basename(f), json.loads(txt)))
    return out


@pytest.mark.parametrize("name,line", _lines(), ids=[n for n, _ in _lines()])
def test_committed_bench_lines_follow_the_contract(name, line):
    assert BASE_KEYS <= set(line), BASE_KEYS - set(line)
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert line["metric"].split(" (")[0] == base["metric"].split(" (")[0]
    assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    # value = clips * 30 s / step time, whole job
    clips = line["config"]["global_batch_clips"]
    assert abs(line["value"] - clips * 30.0 / (line["ms_per_step"] / 1e3)) <= 1e-6 * line["value"]
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0 and e2e["value"] != line["value"]
    r = line["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert line["gpu_launches"] > 0
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(line["clocks"]["reasons"])
    if line["n_gpus"] == 1:
        c = line["cpu_baseline"]
        assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.strip() == ""
