import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import aus_rust_b200 as aus
SR = 44100
ctx = aus.Context(0); ctx.use_torch_stream()
flushbuf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for n in (60 * SR, 2048 + 512 * 5164, 2048 + 512 * 5164 - 16, 2048 + 512 * 5164 + 1000, 2048 + 512 * 1775, 2048 + 512 * 1775 - 16, 2048+512*3551):
    x = torch.empty(n, dtype=torch.float64, device="cuda"); ctx.synth_dev(x, 0, SR, 440.0, 0, 0, "f64")
    F = aus.num_frames(n, 2048, 512)
    feat = torch.empty((F, 14), dtype=torch.float64, device="cuda")
    lens, offs = np.array([n], dtype=np.uint64), np.array([0], dtype=np.uint64)
    call = lambda: ctx.stft_features_batch_dev(x, offs, lens, 2048, 512, 2, SR, "f64", feat)
    for _ in range(5): call()
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        flushbuf.add_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"n={n} frames={F} last_len={n - 512 * (F - 1)}: cold min/median {min(ts)*1e3:.1f}/{sorted(ts)[10]*1e3:.1f} us", flush=True)
