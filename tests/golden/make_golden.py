"""Generates tests/golden/golden_v1.npz from the numpy oracle (oracle/aus_oracle.py).

The reference holds no golden vector for this path (every hot-path test reads an absent WAV and asserts
nothing) and cannot be run here (Rust, no toolchain), so these fixtures are oracle outputs on config-shaped but
short inputs.  Inputs are stored explicitly (numpy's sin may differ by an ulp between CPU SIMD paths).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import aus_oracle as o  # noqa: E402

SR = 44100
out = {}

# config 1 shape: fft 2048 / hop 512 / Hann, full feature set (short last frame included)
x = o.synth_signal(9000, SR, seed=0, file_id=0)
out["c1_x"] = x
out["c1_feat"] = o.stft_features(x, 2048, 512, o.HANNING, SR)
out["c1_spec"] = o.rstft(x, 2048, 512, o.HANNING)[[0, 7, -1]]

# mp::stft_analysis signature: hop = N/2, Hamming (mp.rs:42)
x = o.synth_signal(7000, SR, seed=0, file_id=1)
out["mp_x"] = x
out["mp_feat"] = o.stft_analysis(x, 2048, SR, 8)

# config 2 shape: rstft -> irstft round trip, fft 4096 / hop 1024, two channels
for ch in range(2):
    x = o.synth_signal(3 * 4096 + 100, SR, seed=0, file_id=10 + ch)
    S = o.rstft(x, 4096, 1024, o.HAMMING)
    out[f"c2_x{ch}"] = x
    out[f"c2_spec{ch}_first"] = S[0]
    out[f"c2_audio{ch}"] = o.irstft(S, 4096, 1024, o.HAMMING)

# config 3 shape: batch of clips, fft 1024 / hop 256, features only
lens = [3000, 1024, 2500]
xs = [o.synth_signal(n, SR, freq=o.file_freq(i), seed=0, file_id=20 + i) for i, n in enumerate(lens)]
out["c3_x"] = np.concatenate(xs)
out["c3_lens"] = np.array(lens)
out["c3_feat"] = np.concatenate([o.stft_features(xi, 1024, 256, o.HANNING, SR) for xi in xs])

# config 4 shape: large window fft 16384 / hop 4096
x = o.synth_signal(20000, SR, seed=0, file_id=30)
out["c4_x"] = x
out["c4_feat"] = o.stft_features(x, 16384, 4096, o.HANNING, SR)

# analyzer on a fixed magnitude row + windows
rng = np.random.default_rng(1234)
m = np.abs(rng.standard_normal((3, 1025))) + 0.01
out["an_mag"] = m
out["an_feat"] = o.analyze_spectrogram(m, 2048, SR)
for w in range(5):
    out[f"win{w}_16"] = o.generate_window(w, 16)
    out[f"win{w}_17"] = o.generate_window(w, 17)

np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
print("wrote", os.path.join(HERE, "golden_v1.npz"), sum(v.nbytes for v in out.values()) // 1024, "KiB raw")
