"""Minimal driver for ncu: two time-stretch and two pitch-shift calls on the bench batch shape (1024 x 1.5 s)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wakeword_trainer_home_b200 as w

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = 24000
gen = torch.Generator().manual_seed(0)
x = (0.1 * torch.randn(B, N, generator=gen)).cuda()
plan = w.FeaturePlan(16000, "mfcc", 40, 40, 400, 160, "cuda")
rates = (0.8 + 0.4 * torch.rand(B, generator=gen, dtype=torch.float64)).cuda()
steps = torch.randint(-2, 3, (B,), generator=gen, dtype=torch.int32).cuda()
out = torch.empty_like(x)
for _ in range(2):
    plan.time_stretch(x, rates, rate_lo=0.8, out=out)
    plan.pitch_shift(x, steps, step_range=(-2, 2), out=out)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
