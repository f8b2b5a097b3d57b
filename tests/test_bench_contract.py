"""bench.py's reference arm (the reference's CPU renderer on the host cores) prints one JSON line with the keys the
driver reads; runs here without a GPU on a shortened sample."""
import json
import os
import subprocess
import sys

from conftest import REPO


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, RT_BENCH_REF_SECONDS="1.5", OMP_NUM_THREADS="1")   # torchrun exports OMP_NUM_THREADS=1
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().split("\n")[-1])
    assert d["impl"] == "reference" and d["metric"] == "Msamples/s" and d["unit"] == "Msamples/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["config"]["width"] == 1200 and d["config"]["height"] == 800 and d["config"]["spp"] == 500
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] == d["value"] and "sample" in cb
    assert cb["cores"] == len(os.sched_getaffinity(0))          # every core, despite OMP_NUM_THREADS=1
    assert d["e2e"] == {"value": d["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
