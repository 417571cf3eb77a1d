#!/usr/bin/env python
"""Development: the host-pointer feature call on a pageable 1-hour input under the three staging policies and several
copy-thread counts (AUS_B200_COPY_THREADS), beside a bare numpy copy into pinned memory."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import aus_rust_b200 as aus
SR = 44100
n = 2 * 3600 * SR
x = np.random.default_rng(0).uniform(-1, 1, n)
F = aus.num_frames(n, 2048, 512)
out = np.empty((F, 14))
pin = torch.empty(n, dtype=torch.float64, pin_memory=True).numpy()
t0 = time.perf_counter(); pin[:] = x; t1 = time.perf_counter()
print(f"single-thread numpy copy pageable -> pinned: {n * 8 / (t1 - t0) / 1e9:.1f} GB/s")
ctx = aus.Context(0)
for mode in (0, 1, 2):
    ctx.set_host_register(mode)
    for rep in range(3):
        t0 = time.perf_counter()
        ctx.stft_features_batch(x, [0], [n], 2048, 512, 2, SR, "f64", out=out)
        dt = time.perf_counter() - t0
    print(f"mode {mode} threads {os.environ.get('AUS_B200_COPY_THREADS', 'default')}: {n * 8 / dt / 1e9:.1f} GB/s  {F / dt:.3e} frames/s")
ctx.set_host_register(2)
t0 = time.perf_counter(); ctx.stft_features_batch(pin, [0], [n], 2048, 512, 2, SR, "f64", out=out); dt = time.perf_counter() - t0
print(f"pinned input: {n * 8 / dt / 1e9:.1f} GB/s")
