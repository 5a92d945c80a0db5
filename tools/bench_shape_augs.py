"""Timing of the waveform-shape augmentations (time-stretch, pitch-shift, resample) on one GPU.
Usage: python tools/bench_shape_augs.py [B] [N]   -> one JSON line per op (ms per batch, clips/s)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import wakeword_trainer_home_b200 as ww


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 24000
    gen = torch.Generator().manual_seed(0)
    x = (0.1 * torch.randn(B, N, generator=gen)).cuda()
    plan = ww.FeaturePlan(16000, "mfcc", 40, 40, 400, 160, "cuda")
    rates = (0.8 + 0.4 * torch.rand(B, generator=gen, dtype=torch.float64)).cuda()
    steps = torch.randint(-2, 3, (B,), generator=gen, dtype=torch.int32).cuda()
    out = torch.empty_like(x)
    res = {}
    res["time_stretch"] = timed(lambda: plan.time_stretch(x, rates, rate_lo=0.8, out=out))
    res["pitch_shift"] = timed(lambda: plan.pitch_shift(x, steps, step_range=(-2, 2), out=out))
    x44 = (0.1 * torch.randn(B, int(N * 44100 / 16000), generator=gen)).cuda()
    res["resample_44100_16000"] = timed(lambda: plan.resample(x44, 44100, 16000))
    x8 = (0.1 * torch.randn(B, N // 2, generator=gen)).cuda()
    res["resample_8000_16000"] = timed(lambda: plan.resample(x8, 8000, 16000))
    for k, ms in res.items():
        print(json.dumps({"op": k, "B": B, "N": N, "ms_per_batch": round(ms, 4), "clips_per_s": round(B / ms * 1e3, 1)}))


if __name__ == "__main__":
    main()
