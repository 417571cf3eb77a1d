"""bench.py's CPU-runnable arm (`--impl reference`: the C port of the reference's threaded analyzer) prints the
contract's JSON line; the GPU arm's line is checked for the same keys on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def _run(args, timeout):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


@pytest.mark.timeout(300)
def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"], 280)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "stft_feature_frames_per_sec_fft2048_hop512" and d["unit"] == "frames/s"
    assert d["value"] > 100 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=60, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_gpu_arm_line():
    d = _run(["--steps", "3", "--warmup", "3", "--files", "1", "--e2e-files", "1"], 580)
    assert BASE_KEYS | {"roofline", "clocks", "audio_seconds_per_sec"} <= set(d)
    assert d["n_gpus"] == 1 and d["dtype"] == "f64" and d["scaling"] == "weak" and d["data"] == "synthetic"
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] < 1.2 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert 0.2 < d["roofline_fp64"]["frac"] < 1.0                   # extra key: the pipe that really bounds the f64 kernel
    assert d["gpu_launches"] == 2 * d["steps"]                      # stft_kernel + finalize_kernel per step
    assert d["e2e"]["h2d_bytes_per_step"] == 3600 * 44100 * 8 and d["e2e"]["value"] < d["value"]
    assert d["e2e_pcm16"]["h2d_bytes_per_step"] == 3600 * 44100 * 2 and d["e2e_pcm16"]["value"] > d["e2e"]["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert d["value"] > 100 * d["cpu_baseline"]["value"]
