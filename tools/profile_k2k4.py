#!/usr/bin/env python
"""Small driver for ncu: one K2 (rstft, spectra out) and one K4+K5 (irstft) launch at fft 2048 / hop 512, 10 min f64."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import aus_rust_b200 as aus
ctx = aus.Context(0); ctx.use_torch_stream()
SR, N, hop = 44100, 2048, 512
n = 600 * SR
F, K = aus.num_frames(n, N, hop), N // 2 + 1
x = torch.empty(n, dtype=torch.float64, device="cuda"); ctx.synth_dev(x, 0, SR, 440.0, 0, 0, "f64")
S = torch.empty((F, K, 2), dtype=torch.float64, device="cuda")
y = torch.empty(N + hop * (F - 1), dtype=torch.float64, device="cuda")
feat = torch.empty((F, 14), dtype=torch.float64, device="cuda")
for _ in range(3):
    ctx.stft_features_batch_dev(x, [0], [n], N, hop, 3, SR, "f64", None, S)
    ctx.irstft_dev(S, F, N, hop, 3, "f64", y, True)
    mag = torch.linalg.vector_norm(S, dim=2).contiguous()
    ctx.analyzer_batch_dev(mag, F, N, SR, "f64", feat)
torch.cuda.synchronize()
print("frames", F, float(y.max()))
