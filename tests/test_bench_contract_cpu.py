"""CPU: the reference arm of bench.py (the reference's own objects on the host cores) prints one JSON line with the
keys the driver reads; the b200 arm refuses to run without a GPU instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "channel_samples_per_s" and line["unit"] == "channel-samples/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["value"] > 0
    # default workload = the shape north_star quotes its target on
    assert line["config"]["channels"] == 256 and line["config"]["taps"] == 127 and line["config"]["decimation"] == 100
    assert line["config"]["workload"].startswith("headline") and line["scaling"] == "weak"
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_b200_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                         timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "no CPU path" in (out.stderr + out.stdout)


def test_both_arms_describe_the_same_config():
    """the driver compares the two arms' `config` objects: they come from one function, same keys and values"""
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    for name in bench.CONFIGS:
        for n_gpus in (1, 2, 8):
            args = argparse.Namespace(config=name, batch_log2=25, submits=16)
            cfg = bench.shape(args, n_gpus)
            d = bench.config_dict(args, cfg, n_gpus)
            assert set(d) == {"workload", "channels", "channels_per_gpu", "taps", "decimation", "fs", "batch_complex_samples",
                              "submits_per_step", "l2_policy", "parallelism"}
            assert d["channels"] == cfg["c_gpu"] * n_gpus
    c4 = bench.shape(argparse.Namespace(config="c4"), 8)
    assert c4["scaling"] == "strong" and c4["c_gpu"] == 128 and c4["c_total"] == 1024
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["config"]["channels"] == 64 and line["config"]["workload"].startswith("c2")
