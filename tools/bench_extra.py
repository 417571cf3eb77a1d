#!/usr/bin/env python
"""Auxiliary device-resident timings for the other rows of SURVEY 8 (not the headline metric): BASELINE configs
1-4 and the K2 / K3 / K4+K5 kernels.  One JSON object per line on stdout.  CUDA-event timing, warm-up 2, best of 5.
Roofline fractions use the algorithmic bytes of SURVEY 8(d) and MEASURED_PEAKS.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch

import __graft_entry__ as g

g.build()
import aus_rust_b200 as aus
import aus_oracle as orc

SR = 44100
dev = torch.device("cuda", 0)
ctx = aus.Context(0)
ctx.use_torch_stream()
peak = 6650.0
pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pp):
    peak = float(json.load(open(pp))["hbm_gbs"])


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def frames_of(n, N, hop):
    return 1 if n <= N else -(-(n - N) // hop) + 1


def synth(n_files, n_samples, prec, base_id=0):
    tdt = torch.float64 if prec == "f64" else torch.float32
    d = torch.empty(n_files * n_samples, dtype=tdt, device=dev)
    for i in range(n_files):
        ctx.synth_dev(d[i * n_samples:(i + 1) * n_samples], 0, SR, orc.file_freq(base_id + i), 0, base_id + i, prec)
    return d


def features_case(name, n_files, n_samples, N, hop, window, prec):
    esz = 8 if prec == "f64" else 4
    d = synth(n_files, n_samples, prec)
    lens = np.full(n_files, n_samples, dtype=np.uint64)
    offs = np.arange(n_files, dtype=np.uint64) * np.uint64(n_samples)
    F = frames_of(n_samples, N, hop) * n_files
    out = torch.empty((F, 14), dtype=d.dtype, device=dev)
    ms = timeit(lambda: ctx.stft_features_batch_dev(d, offs, lens, N, hop, window, SR, prec, out))
    assert torch.isfinite(out).all()
    alg = n_files * n_samples * esz + F * 14 * esz
    print(json.dumps({"case": name, "kernel": "K1 stft_kernel + finalize", "dtype": prec, "fft": N, "hop": hop, "frames": F,
                      "ms": ms, "frames_per_s": F / ms * 1e3, "audio_s_per_s": F / ms * 1e3 * hop / SR,
                      "hbm_frac": alg / (ms * 1e-3) / 1e9 / peak}), flush=True)
    del d, out


for prec in ("f64", "f32"):
    features_case("config1: 60 s mono", 1, 60 * SR, 2048, 512, 2, prec)
    features_case("config3: 1000 x 30 s clips", 1000, 30 * SR, 1024, 256, 2, prec)
    features_case("config4: 1 h, large window", 1, 3600 * SR, 16384, 4096, 2, prec)
    features_case("config5 shard: 4 x 1 h", 4, 3600 * SR, 2048, 512, 2, prec)

# K2 at the headline shape: spectra written out (fft 2048 / hop 512), 1 h of audio
for prec in ("f64", "f32"):
    esz = 8 if prec == "f64" else 4
    N, hop, n = 2048, 512, 3600 * SR
    K, F = N // 2 + 1, frames_of(n, N, hop)
    x = synth(1, n, prec)
    spec = torch.empty((F, K, 2), dtype=x.dtype, device=dev)
    lens, offs = np.array([n], dtype=np.uint64), np.array([0], dtype=np.uint64)
    ms = timeit(lambda: ctx.stft_features_batch_dev(x, offs, lens, N, hop, 2, SR, prec, None, spec))
    print(json.dumps({"case": "rstft only: 1 h mono", "kernel": "K2 stft_kernel (spectra out)", "dtype": prec, "fft": N, "hop": hop,
                      "frames": F, "ms": ms, "frames_per_s": F / ms * 1e3,
                      "hbm_frac": (n * esz + F * K * 2 * esz) / (ms * 1e-3) / 1e9 / peak}), flush=True)
    del x, spec

# config 2: rstft -> irstft round trip, 10 min stereo, fft 4096 hop 1024 (Hamming: finite at the ends)
for prec in ("f64", "f32"):
    esz = 8 if prec == "f64" else 4
    N, hop, n = 4096, 1024, 600 * SR
    K = N // 2 + 1
    F = frames_of(n, N, hop)
    res = {"case": "config2: 10 min stereo round trip", "dtype": prec, "fft": N, "hop": hop, "frames": 2 * F}
    t_f = t_i = t_a = 0.0
    errs = []
    for ch in range(2):
        x = synth(1, n, prec, base_id=100 + ch)
        spec = torch.empty((F, K, 2), dtype=x.dtype, device=dev)
        y = torch.empty(N + hop * (F - 1), dtype=x.dtype, device=dev)
        feat = torch.empty((F, 14), dtype=x.dtype, device=dev)
        lens, offs = np.array([n], dtype=np.uint64), np.array([0], dtype=np.uint64)
        t_f += timeit(lambda: ctx.stft_features_batch_dev(x, offs, lens, N, hop, 3, SR, prec, None, spec))
        t_i += timeit(lambda: ctx.irstft_dev(spec, F, N, hop, 3, prec, y, True))
        mag = torch.linalg.vector_norm(spec, dim=2).contiguous()
        t_a += timeit(lambda: ctx.analyzer_batch_dev(mag, F, N, SR, prec, feat))
        # reconstruction: the reference peak-normalises, so compare up to the (least-squares) gain over the body
        body = slice(N, n - 2 * N)
        gain = float((y[body] * x[body]).sum() / (x[body] * x[body]).sum())
        errs.append(float((y[body] / gain - x[body]).abs().max()))
        del x, spec, y, feat, mag
    res.update({
        "rstft_ms": t_f, "rstft_frames_per_s": 2 * F / t_f * 1e3,
        "rstft_hbm_frac": (2 * n * esz + 2 * F * K * 2 * esz) / (t_f * 1e-3) / 1e9 / peak,
        "irstft_ms": t_i, "irstft_frames_per_s": 2 * F / t_i * 1e3,
        "irstft_hbm_frac": (2 * F * K * 2 * esz + 3 * 2 * (N + hop * (F - 1)) * esz) / (t_i * 1e-3) / 1e9 / peak,
        "analyzer_ms": t_a, "analyzer_frames_per_s": 2 * F / t_a * 1e3,
        "analyzer_hbm_frac": (2 * F * K * esz + 2 * F * 14 * esz) / (t_a * 1e-3) / 1e9 / peak,
        "roundtrip_max_abs_err_body": max(errs)})
    print(json.dumps(res), flush=True)
