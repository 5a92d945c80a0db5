"""Parity of time-stretch at 'nice' rates whose multiples land exactly on integers (0.9 * 10 = 9, ...)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import wakeword_trainer_home_b200 as ww  # noqa: E402
from oracle import ta_oracle as tao  # noqa: E402

rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
gen = torch.Generator().manual_seed(0)
x = 0.1 * torch.randn(2, 24000, generator=gen)
plan = ww.FeaturePlan(16000, "mel", 40, 40, 400, 160, "cuda")
for r in (0.9, 0.9137, 1.1, 1.25, 0.8, 0.5, 2.0, 1.5):
    rates = torch.tensor([r, r], dtype=torch.float64)
    got = plan.time_stretch(x.cuda(), rates).cpu()
    a32, a64 = tao.time_stretch(x, rates), tao.time_stretch(x.double(), rates)
    ts32 = torch.arange(0, 188, r, dtype=torch.float32)
    mine = (torch.arange(len(ts32), dtype=torch.float64) * r).float()
    print(f"rate {r}: vs64 {rel(got, a64):.2e} vs32 {rel(got, a32):.2e} oracle gap {rel(a32, a64):.2e}  "
          f"floor(arange32) != floor(float32(j*rate)) at {int((ts32.floor() != mine.floor()).sum())} steps")
