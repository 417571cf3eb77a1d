#!/usr/bin/env python
"""Development: fused feature rate on digital silence (every frame takes the exact / NaN-semantics paths) beside the synthetic signal."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import aus_rust_b200 as aus
SR = 44100
ctx = aus.Context(0); ctx.use_torch_stream()
n = 600 * SR
F = aus.num_frames(n, 2048, 512)
feat = torch.empty((F, 14), dtype=torch.float64, device="cuda")
lens, offs = np.array([n], dtype=np.uint64), np.array([0], dtype=np.uint64)
for name in ("signal", "silence", "half"):
    x = torch.zeros(n, dtype=torch.float64, device="cuda")
    if name != "silence":
        ctx.synth_dev(x, 0, SR, 440.0, 0, 0, "f64")
    if name == "half":
        x[n // 2:] = 0.0
    for prec in ("f64",):
        call = lambda: ctx.stft_features_batch_dev(x, offs, lens, 2048, 512, 2, SR, prec, feat)
        for _ in range(2): call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        f = feat.cpu().numpy()
        print(f"{name:8s} {prec}: {F / ms * 1e3:.3e} frames/s  ({ms:.3f} ms; NaN rows {int(np.isnan(f[:, 0]).sum())} of {F})")
