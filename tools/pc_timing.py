#!/usr/bin/env python
"""Development: per-phase cycle counts of the producer/consumer kernel (needs a build with -DAUS_PC_TIMING=1, loaded
through AUS_B200_LIB).  Phases are clock64() deltas of lane 0 of pair 0 in block 0, averaged per frame."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import aus_rust_b200 as aus
SR = 44100
ctx = aus.Context(0); ctx.use_torch_stream()
n = 3600 * SR
x = torch.empty(n, dtype=torch.float64, device="cuda"); ctx.synth_dev(x, 0, SR, 440.0, 0, 0, "f64")
F = aus.num_frames(n, 2048, 512)
feat = torch.empty((F, 14), dtype=torch.float64, device="cuda")
lens, offs = np.array([n], dtype=np.uint64), np.array([0], dtype=np.uint64)
lib = aus.load_library()
buf = (ctypes.c_ulonglong * 40)()
for rep in range(2):
    ctx.stft_features_batch_dev(x, offs, lens, 2048, 512, 2, SR, "f64", feat)
    torch.cuda.synchronize()
    lib.aus_debug_pc_timing(buf)
v = list(buf)
itp, itc = max(v[32], 1), max(v[33], 1)
names = {0: "C wait full", 1: "C bands", 2: "C allreduce2", 3: "C passB", 4: "C allreduce10+scan", 5: "C totals/ent", 6: "C rolloff ballots",
         7: "C walk+write", 8: "P loop top+window ldg", 9: "P wait TMA", 10: "P fft", 11: "P locate next", 12: "P wait empty", 13: "P tail", 14: "P  pass1 loads+emul", 15: "P  pass1 butterflies", 16: "P  pass1 stores", 17: "P  pass2 loads+cmul",
         18: "P  pass2 butterflies", 19: "P  pass2 stores"}
for i in range(20):
    print(f"{names[i]:24s} {v[i] / (itc if i < 8 else itp):9.0f} cycles/frame")
print("consumer total", sum(v[:8]) / itc, "producer total", sum(v[8:20]) / itp)
