"""bench.py's reference arm runs on the CPU (the oracle port on the host cores): its JSON line is the driver's contract for
BOTH arms (same metric / unit / config keys), so it is checked here without a GPU."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")   # what torchrun exports; the arm must ignore it and use the whole box
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["dtype"] == "f64" and line["gpu_launches"] == 0 and line["value"] > 0
    assert line["config"]["workload"].startswith("synthetic 1000 taxa x 100k patterns")
    assert {"taxa", "patterns_per_gpu", "categories", "draws_per_step", "parallelism"} <= set(line["config"])
    cpu = line["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["value"] == line["value"] and cpu["cores"] == len(os.sched_getaffinity(0))
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
