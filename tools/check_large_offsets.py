#!/usr/bin/env python
"""One-off robustness check on a B200: a batch whose sample offsets exceed 2^32 (f32: 30 one-hour files = 4.76e9 samples,
19 GB; f64: 16 files = 2.5e9 samples > 2^31).  The features of the last files computed inside the big batch must be
bit-identical to the same files analysed alone at offset 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import aus_rust_b200 as aus

SR, N, HOP = 44100, 2048, 512
n = 3600 * SR
ctx = aus.Context(0)
ctx.use_torch_stream()
for prec, files in (("f32", 30), ("f64", 16)):
    dt = torch.float32 if prec == "f32" else torch.float64
    d = torch.empty(files * n, dtype=dt, device="cuda")
    for i in range(files):
        ctx.synth_dev(d[i * n:(i + 1) * n], 0, SR, 220.0 + 7.0 * i, 0, i, prec)
    offs = np.arange(files, dtype=np.uint64) * np.uint64(n)
    lens = np.full(files, n, dtype=np.uint64)
    F1 = aus.num_frames(n, N, HOP)
    out = torch.empty((files * F1, 14), dtype=dt, device="cuda")
    ctx.stft_features_batch_dev(d, offs, lens, N, HOP, 2, SR, prec, out)
    torch.cuda.synchronize()
    for i in (0, files // 2, files - 1):
        alone = torch.empty((F1, 14), dtype=dt, device="cuda")
        ctx.stft_features_batch_dev(d[i * n:(i + 1) * n], [0], [n], N, HOP, 2, SR, prec, alone)
        torch.cuda.synchronize()
        same = torch.equal(out[i * F1:(i + 1) * F1].nan_to_num(), alone.nan_to_num())
        print(prec, "file", i, "sample offset", int(offs[i]), "identical" if same else "DIFFERENT", flush=True)
        assert same
    # spectra + inverse of the last file through offsets > 2^32 bytes as well
    S = torch.empty((F1, N // 2 + 1, 2), dtype=dt, device="cuda")
    ctx.stft_features_batch_dev(d[(files - 1) * n:], [0], [n], N, HOP, 3, SR, prec, None, S)
    y = torch.empty(N + HOP * (F1 - 1), dtype=dt, device="cuda")
    ctx.irstft_dev(S, F1, N, HOP, 3, prec, y, False)
    torch.cuda.synchronize()
    x = d[(files - 1) * n:]
    body = slice(N, n - 2 * N)
    err = float((y[body] / N - x[body]).abs().max())
    print(prec, "round trip of the last file: max |err|", err, flush=True)
    assert err < (1e-9 if prec == "f64" else 2e-4)
    del d, out, S, y
    torch.cuda.empty_cache()
print("large-offset check ok")
