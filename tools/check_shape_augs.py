"""Print the parity of time-stretch / pitch-shift against the torchaudio oracle (float32 and float64) per clip."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import wakeword_trainer_home_b200 as ww  # noqa: E402
from oracle import ta_oracle as tao  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


gen = torch.Generator().manual_seed(42)
B, N = 8, 24000
x = 0.1 * torch.randn(B, N, generator=gen)
t = torch.arange(N) / 16000.0
x[2] = 0.5 * torch.sin(2 * torch.pi * 440.0 * t) + 0.2 * torch.sin(2 * torch.pi * 1230.0 * t)
x[3, N // 2:] = 0.0
x[5] = 0.5 * torch.sin(2 * torch.pi * 440.0 * t)
rates = torch.tensor([0.8013, 1.1987, 0.9371, 1.0629, 1.0, 0.8642, 1.1318, 1.0], dtype=torch.float64)
plan = ww.FeaturePlan(16000, "mel", 40, 40, 400, 160, "cuda")
got = plan.time_stretch(x.cuda(), rates).cpu()
w64, w32 = tao.time_stretch(x.double(), rates), tao.time_stretch(x, rates)
for b in range(B):
    if rates[b] != 1.0:
        print(f"stretch b={b} rate={float(rates[b]):.4f} vs64 {rel(got[b], w64[b]):.2e} vs32 {rel(got[b], w32[b]):.2e} oracle gap {rel(w32[b], w64[b]):.2e}")
steps = torch.tensor([-2, -1, 0, 1, 2, 2, -2, 4], dtype=torch.int32)
got = plan.pitch_shift(x.cuda(), steps).cpu()
w64, w32 = tao.pitch_shift(x.double(), steps), tao.pitch_shift(x, steps)
for b in range(B):
    if steps[b] != 0:
        print(f"pitch b={b} n={int(steps[b])} vs64 {rel(got[b], w64[b]):.2e} vs32 {rel(got[b], w32[b]):.2e} oracle gap {rel(w32[b], w64[b]):.2e}")
