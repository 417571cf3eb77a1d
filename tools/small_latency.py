#!/usr/bin/env python
"""Development: per-call device time of the fused feature call on small inputs (config 1: one 60 s file), warm
(back to back) and cold (L2 flushed before every call), CUDA events around single calls."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import aus_rust_b200 as aus
SR = 44100
ctx = aus.Context(0); ctx.use_torch_stream()
flushbuf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for secs in (10, 60, 300):
    n = secs * SR
    x = torch.empty(n, dtype=torch.float64, device="cuda"); ctx.synth_dev(x, 0, SR, 440.0, 0, 0, "f64")
    F = aus.num_frames(n, 2048, 512)
    feat = torch.empty((F, 14), dtype=torch.float64, device="cuda")
    lens, offs = np.array([n], dtype=np.uint64), np.array([0], dtype=np.uint64)
    call = lambda: ctx.stft_features_batch_dev(x, offs, lens, 2048, 512, 2, SR, "f64", feat)
    for _ in range(5): call()
    torch.cuda.synchronize()
    res = {}
    for mode in ("warm", "cold"):
        ts = []
        for _ in range(20):
            if mode == "cold": flushbuf.add_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); call(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[mode] = (min(ts), sorted(ts)[len(ts) // 2])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): call()
    e1.record(); torch.cuda.synchronize()
    print(f"{secs:4d} s ({F} frames): warm min/median {res['warm'][0]*1e3:.1f}/{res['warm'][1]*1e3:.1f} us, cold {res['cold'][0]*1e3:.1f}/{res['cold'][1]*1e3:.1f} us, "
          f"50 back-to-back {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per call")
