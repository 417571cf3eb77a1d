#!/usr/bin/env python
"""Tiny invocations of every kernel (K1/K2/K3/K4/K5, f64 and f32, one size per FFT structure) for compute-sanitizer."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def synth(n, seed=0):
    """sine + noise; any signal will do here (results are not checked: the sanitizer watches the accesses)"""
    r = np.random.default_rng(seed)
    return 0.5 * np.sin(0.0627 * np.arange(n)) + 0.1 * r.uniform(-1, 1, n)

import aus_rust_b200 as aus

ctx = aus.Context(0)
SR = 44100
sizes = [int(a) for a in sys.argv[1:]] or [64, 512, 1024, 2048, 4096, 16384]
for N in sizes:
    hop = N // 4
    x = synth(3 * N + 101, N)           # 10 frames, the last one short; odd length
    for prec in ("f64", "f32"):
        f = ctx.stft_features(x, N, hop, 2, SR, prec)
        S = ctx.rstft(x, N, hop, 3, prec)
        y = ctx.irstft(S, N, hop, 3, prec)
        a = ctx.analyzer_batch(np.abs(S), N, SR, prec)
        ctx.stft_features(x[1:], N, hop + 1 if hop + 1 <= N else hop, 1, SR, prec)   # odd hop / offset: non-TMA staging
        print(N, prec, f.shape, S.shape, y.shape, a.shape, float(np.nanmax(np.abs(f))), flush=True)
lens = [0, 5, 700, 3000]
offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
print(ctx.stft_features_batch(synth(sum(lens)), offs, lens, 256, 64, 2, SR)[0].shape)
print("sanitize smoke done")
# round-2 entry points: producer/consumer variant, irstft batch / shards, per-frame time features, PCM + mixdown, polar
c2 = aus.Context(0)
c2.set_forward_variant(2)
x = synth(9000, 5)
print("pc", c2.stft_features(x, 2048, 512, 2, SR).shape, c2.stft_features(x[1:], 2048, 511, 1, SR).shape)
S = [ctx.rstft(synth(n, 6), 1024, 256, 3) for n in (5000, 1024, 9001)]
print("irstft_batch", [y.shape for y in ctx.irstft_batch(S, 1024, 256, 3)])
print("irstft_multi", [y.shape for y in aus.irstft_multi([0, 0, 0], S, 1024, 256, 3)])
print("time frames", ctx.time_features_frames(x, [0, 4000], [4000, 5000], 1024, 256, SR).shape)
pcm = (x[:4000].reshape(-1, 2) * 127).astype(np.int8)
print("pcm", ctx.pcm_to_f64(pcm, 2, True, True).shape, ctx.mixdown(np.stack([x[:100], x[100:200], x[200:300]])).shape)
mag, ph = ctx.complex_to_polar(S[0])
print("polar", ctx.polar_to_complex(mag, ph).shape)
aus.multi_shutdown()
print("sanitize smoke (round 2) done")
