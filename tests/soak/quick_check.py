#!/usr/bin/env python
"""Quick oracle comparison used while tuning kernels (tools/kbench.py --check): a short signal through K1 (f64, f32), K2, K3
and K4 at one fft_size, printed as max errors.  Test infrastructure: lives under tests/ because it uses oracle/.
AUS_B200_LIB selects another build of the library."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np

import aus_oracle as orc
import aus_rust_b200 as aus

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2048)
ap.add_argument("--hop", type=int, default=512)
args = ap.parse_args()
SR = 44100
ctx = aus.Context(0)
if True:
    x = orc.synth_signal(3 * SR)
    ref = orc.stft_features(x, args.n, args.hop, orc.HANNING, SR)
    got = ctx.stft_features(x, args.n, args.hop, aus.WindowType.Hanning, SR, "f64")
    rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
    print("check f64 features max rel", rel.max(axis=0).tolist(), flush=True)
    got32 = ctx.stft_features(x, args.n, args.hop, aus.WindowType.Hanning, SR, "f32")
    rel32 = np.abs(got32 - ref) / np.maximum(np.abs(ref), 1e-300)
    print("check f32 features max rel", rel32.max(axis=0).tolist(), flush=True)
    sref = orc.rstft(x, args.n, args.hop, orc.HANNING)
    sgot = np.asarray(ctx.rstft(x, args.n, args.hop, int(aus.WindowType.Hanning)))
    print("check f64 spectra max |d| / max|X|", float(np.abs(sgot - sref).max() / np.abs(sref).max()), flush=True)

if True:
    rng = np.random.default_rng(3)
    mags = np.abs(rng.normal(size=(40, args.n // 2 + 1))) + 1e-3
    mags[5, :] = 0.0
    mags[6, 10] = 0.0
    ref = orc.analyze_spectrogram(mags, args.n, SR)
    got = ctx.analyzer_batch(mags, args.n, SR)
    with np.errstate(all="ignore"):
        rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
    rel = np.where(np.isnan(got) & np.isnan(ref), 0.0, rel)
    print("check k3 f64 max rel per feature", np.nanmax(rel, axis=0).tolist(), "nan pattern equal",
          bool(np.array_equal(np.isnan(got), np.isnan(ref))), flush=True)

if True:
    xs = orc.synth_signal(40 * args.n + 123, SR, seed=8, file_id=args.n)
    for wt in (orc.HAMMING, orc.HANNING):
        S = orc.rstft(xs, args.n, args.hop, wt)
        ref = orc.irstft(S, args.n, args.hop, wt)
        got = ctx.irstft(S, args.n, args.hop, wt)
        fin = np.isfinite(ref)
        print("check k4 f64 window", wt, "max |d| / peak", float(np.max(np.abs(got[fin] - ref[fin])) / np.max(np.abs(ref[fin]))),
              "nan pattern equal", bool(np.array_equal(np.isnan(got), np.isnan(ref))), flush=True)

