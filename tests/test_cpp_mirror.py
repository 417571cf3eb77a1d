"""The C++ host mirror (aus-rust_b200/cpp/aus.hpp) compiles and links against the C ABI on CPU, and on a GPU gives
the oracle's numbers."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "aus-rust_b200", "csrc")
EXE = os.path.join(ROOT, "tests", "cpp_mirror_main.bin")


def _build():
    src = os.path.join(ROOT, "tests", "cpp_mirror_main.cpp")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-o", EXE, src, "-L" + CSRC, "-laus_b200",
                    "-Wl,-rpath," + CSRC], check=True)


def test_cpp_mirror_compiles_and_links(aus):
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_mirror_matches_oracle(aus, orc, tmp_path):
    _build()
    x = orc.synth_signal(30000, 44100, seed=16)
    f = tmp_path / "audio.f64"
    x.tofile(f)
    out = subprocess.run([EXE, str(f)], capture_output=True, text=True, check=True).stdout.splitlines()
    ref = orc.stft_analysis(x, 2048, 44100, 8)
    assert out[0] == f"frames {ref.shape[0]}"
    rows = [l.split() for l in out if l.startswith("row")]
    assert len(rows) >= 2
    for r in rows:
        i, vals = int(r[1]), np.array([float(v) for v in r[2:]])
        assert np.all(np.abs(vals - ref[i]) <= 1e-9 * np.abs(ref[i]))
    ir = [l for l in out if l.startswith("irstft")][0].split()
    assert int(ir[1]) == 2048 + 1024 * (orc.num_frames(x.size, 2048, 1024) - 1)
    assert float(ir[2]) == pytest.approx(10 ** (-6 / 20), rel=1e-12)
    w = [float(v) for v in [l for l in out if l.startswith("wrappers")][0].split()[1:]]
    mag = np.abs(orc.rstft(x, 2048, 1024, orc.HAMMING)[1])
    fr = orc.rfftfreq(2048, 44100)
    assert w[0] == orc.spectral_roll_off_point(mag, fr, 0.85)
    assert w[1] == pytest.approx(orc.spectral_slope_region(mag, fr, 105.0, 852.0, 44100), rel=1e-7)
    assert w[2] == pytest.approx(orc.spectral_centroid(mag, fr), rel=1e-9)
    err = [l for l in out if l.startswith("error")][0]
    assert "The spectrum at frame 1 has an incorrect length. It should be 1025, but it is actually 1024." in err
