#!/usr/bin/env python
"""One-off robustness check on a B200: 300 000 ragged files (lengths 0..3000, gaps between them) in ONE host-pointer call and
one 12-hour file cut by the host pipeline; sampled files / frame ranges are compared with the numpy oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import aus_oracle as orc
import aus_rust_b200 as aus

SR = 44100
ctx = aus.Context(0)
rng = np.random.default_rng(11)

# ---- many tiny files
nf, N, hop = 300000, 256, 64
lens = rng.integers(0, 3001, nf).astype(np.uint64)
gaps = rng.integers(0, 4, nf).astype(np.uint64)
offs = np.concatenate([[0], np.cumsum(lens + gaps)[:-1]]).astype(np.uint64)
audio = orc.synth_signal(int(offs[-1] + lens[-1]), SR, seed=3)
out, starts = ctx.stft_features_batch(audio, offs, lens, N, hop, orc.HAMMING, SR, "f64")
assert starts[-1] == out.shape[0] == sum(orc.num_frames(int(n), N, hop) for n in lens)
worst = 0.0
for i in list(rng.integers(0, nf, 300)) + [0, nf - 1]:
    ref = orc.stft_features(audio[int(offs[i]):int(offs[i] + lens[i])], N, hop, orc.HAMMING, SR)
    got = out[starts[i]:starts[i + 1]]
    assert got.shape == ref.shape, (i, got.shape, ref.shape)
    assert np.array_equal(np.isnan(got), np.isnan(ref)), i
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
    r = np.where(np.isnan(ref) | (got == ref), 0.0, r)
    r[:, 10:] = 0.0          # slopes: differences of large terms, compared with their cancellation scale in the tests
    r[:, 2] = 0.0
    worst = max(worst, float(r.max()) if r.size else 0.0)
print(f"{nf} files, {out.shape[0]} frames: worst relative difference on sampled files {worst:.2e}", flush=True)
assert worst < 1e-9

# ---- one 12-hour file through the chunked host pipeline
n = 12 * 3600 * SR
x = orc.synth_signal(n, SR, seed=4)
N, hop = 2048, 512
got = ctx.stft_features(x, N, hop, orc.HANNING, SR, "f64")
F = orc.num_frames(n, N, hop)
assert got.shape == (F, 14)
worst = 0.0
for f0 in (0, F // 3, F - 40):
    seg = x[f0 * hop:min(n, (f0 + 39) * hop + N)] if f0 + 40 < F else x[f0 * hop:]
    ref = orc.stft_features(seg, N, hop, orc.HANNING, SR)
    k = min(ref.shape[0], 40) if f0 + 40 >= F else 39      # an interior slice ends with an artificial short frame: skip it
    r = np.abs(got[f0:f0 + k, :10] - ref[:k, :10]) / np.maximum(np.abs(ref[:k, :10]), 1e-300)
    worst = max(worst, float(r.max()))
print(f"12 h file, {F} frames: worst relative difference on sampled ranges {worst:.2e}")
assert worst < 1e-9
print("many-files check ok")
