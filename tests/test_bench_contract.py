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
    assert d["config"]["files_per_gpu"] == 4 and "1-hour" in d["cpu_baseline"]["sample"]      # same `config` object as the GPU arm


def test_reference_arm_never_loads_the_product():
    """the reference arm mapped libaus_b200.so through g.build().  It must load the oracle library only."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    ref = src[src.index("def run_reference"):src.index("# --------------------------------------------------------------------------------------------- GPU plumbing")]
    assert "__graft_entry__" not in ref and "aus_rust_b200" not in ref and "load_library" not in ref
    port = src[src.index("def _oracle_lib"):src.index("def cpu_port_rate")]
    assert "__graft_entry__" not in port and "libaus_oracle.so" in port
    gpu = src[src.index("def run_ours"):src.index("def run_config")]
    assert "aus_oracle" not in gpu and "import aus_oracle" not in src            # the GPU arm imports nothing from oracle/


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
    e = d["e2e"]
    assert e["platform_h2d_gbs"] > 1 and 0 < e["frac_of_platform_h2d"] <= 1.05        # the call cannot beat a bare pinned copy
    assert d["e2e_pageable"]["value"] > 0 and d["e2e_pageable"]["value"] <= 1.05 * e["value"]
    assert d["other_precision"]["dtype"] == "f32" and d["other_precision"]["value"] > d["value"]


@pytest.mark.gpu
@pytest.mark.timeout(600)
@pytest.mark.parametrize("cfg", [1, 2, 3, 4])
def test_gpu_arm_other_baseline_configs(cfg):
    d = _run(["--config", str(cfg), "--steps", "3", "--warmup", "3"], 580)
    assert d["config"]["baseline_config"] == cfg and d["value"] > 0 and 0 < d["roofline"]["frac"] < 1.2
    assert d["clocks"] is not None and d["gpu_launches"] >= 3
    if cfg == 2:
        assert d["reconstruction_max_abs_error_body"] < 1e-9 and abs(d["peak_after_normalisation"] - 10 ** (-6 / 20)) < 1e-12


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_gpu_corpus_leg():
    d = _run(["--corpus", "--corpus-files", "6"], 580)
    assert d["finite"] and d["analyse_stream"] > 1e7 and 0 < d["generate_analyse_stream"] <= 1.05 * d["analyse_stream"]
    assert d["feature_d2h_gbs"] > 1


def test_build_is_quiet_on_stdout_and_safe_to_call_concurrently():
    """bench.py calls build() in every torchrun rank and prints ONE JSON line on stdout: build() must serialise on its lock
    (two `make`s in one directory race) and send all tool output to stderr."""
    code = "import __graft_entry__ as g; g.build()"
    procs = [subprocess.Popen([sys.executable, "-c", code], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
             for _ in range(3)]
    for p in procs:
        out, err = p.communicate(timeout=1500)
        assert p.returncode == 0, err.decode()[-2000:]
        assert out == b"", out[:500]
