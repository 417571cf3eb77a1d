#!/usr/bin/env python
"""Kernel micro-benchmark used while tuning: device-resident timings of K1 (fused features), K2 (spectra out), the mixed
launch, K3 (analyzer) and K4+K5 (irstft) at one fft_size, f64 and f32.  CUDA events, warm-up 2, best of 5.
AUS_B200_LIB selects another build of the library (e.g. a single-size development build).
Usage: kbench.py [--n 2048] [--hop 512] [--seconds 3600] [--kernels k1,k2,mix,k3,k4] [--prec f64,f32] [--check]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import aus_rust_b200 as aus

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2048)
ap.add_argument("--hop", type=int, default=512)
ap.add_argument("--seconds", type=int, default=3600)
ap.add_argument("--kernels", default="k1,k2,mix,k3,k4")
ap.add_argument("--prec", default="f64,f32")
ap.add_argument("--check", action="store_true", help="run tests/soak/quick_check.py (oracle comparison of a short signal) first")
ap.add_argument("--tag", default="")
args = ap.parse_args()

SR = 44100
dev = torch.device("cuda", 0)
ctx = aus.Context(0)
ctx.use_torch_stream()
peak = 6553.6
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


if args.check:          # the oracle comparisons live with the tests (only tests/ may touch oracle/): run them as a child process
    import subprocess
    subprocess.run([sys.executable, os.path.join(ROOT, "tests", "soak", "quick_check.py"), "--n", str(args.n), "--hop", str(args.hop)],
                   check=False, env=dict(os.environ))

N, hop, n = args.n, args.hop, args.seconds * SR
K = N // 2 + 1
F = aus.num_frames(n, N, hop)
lens, offs = np.array([n], dtype=np.uint64), np.array([0], dtype=np.uint64)
kernels = args.kernels.split(",")
for prec in args.prec.split(","):
    tdt = torch.float64 if prec == "f64" else torch.float32
    esz = 8 if prec == "f64" else 4
    x = torch.empty(n, dtype=tdt, device=dev)
    ctx.synth_dev(x, 0, SR, 440.0, 0, 0, prec)
    feat = torch.empty((F, 14), dtype=tdt, device=dev)
    res = {"tag": args.tag, "prec": prec, "N": N, "hop": hop, "frames": F}
    if "k1" in kernels:
        ms = timeit(lambda: ctx.stft_features_batch_dev(x, offs, lens, N, hop, 2, SR, prec, feat))
        res["k1_fps"] = F / ms * 1e3
        res["k1_hbm"] = (n * esz + F * 14 * esz) / (ms * 1e-3) / 1e9 / peak
    need_spec = any(k in kernels for k in ("k2", "mix", "k3", "k4"))
    if need_spec:
        spec = torch.empty((F, K, 2), dtype=tdt, device=dev)
    if "k2" in kernels:
        ms = timeit(lambda: ctx.stft_features_batch_dev(x, offs, lens, N, hop, 2, SR, prec, None, spec))
        res["k2_fps"] = F / ms * 1e3
        res["k2_hbm"] = (n * esz + F * K * 2 * esz) / (ms * 1e-3) / 1e9 / peak
    if "mix" in kernels:
        ms = timeit(lambda: ctx.stft_features_batch_dev(x, offs, lens, N, hop, 2, SR, prec, feat, spec))
        res["mix_fps"] = F / ms * 1e3
        res["mix_hbm"] = (n * esz + F * K * 2 * esz + F * 14 * esz) / (ms * 1e-3) / 1e9 / peak
    if "k3" in kernels or "k4" in kernels:
        ctx.stft_features_batch_dev(x, offs, lens, N, hop, 3, SR, prec, None, spec)
    if "k3" in kernels:
        mag = torch.linalg.vector_norm(spec, dim=2).contiguous()
        ms = timeit(lambda: ctx.analyzer_batch_dev(mag, F, N, SR, prec, feat))
        res["k3_fps"] = F / ms * 1e3
        res["k3_hbm"] = (F * K * esz + F * 14 * esz) / (ms * 1e-3) / 1e9 / peak
        del mag
    if "k4" in kernels:
        y = torch.empty(N + hop * (F - 1), dtype=tdt, device=dev)
        ms = timeit(lambda: ctx.irstft_dev(spec, F, N, hop, 3, prec, y, True))
        res["k4_fps"] = F / ms * 1e3
        res["k4_hbm"] = (F * K * 2 * esz + 3 * (N + hop * (F - 1)) * esz) / (ms * 1e-3) / 1e9 / peak
        del y
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) and v < 10 else (float("%.4g" % v) if isinstance(v, float) else v))
                      for k, v in res.items()}), flush=True)
    del x, feat
    if need_spec:
        del spec
