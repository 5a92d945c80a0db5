"""Minimal driver for ncu: a few steps of the bench workload (BASELINE.json configs[1]) and nothing else."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import wakeword_trainer_home_b200 as w

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--batch", type=int, default=bench.B_PER_GPU)
ap.add_argument("--no-aug", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
plan = w.FeaturePlan(bench.SR, "mfcc", bench.N_MELS, bench.N_MFCC, bench.N_FFT, bench.HOP, dev)
noise, rirs = bench.synth_banks()
plan.register_noise(noise)
plan.register_rirs(rirs)
wav, d = bench.synth(0, a.batch)
wav = wav.to(dev)
aug = None if a.no_aug else w.AugParams(**d).to(dev)
out = None
for i in range(a.steps):
    out = plan.featurize(wav, aug, out=out)
torch.cuda.synchronize()
print("ok", tuple(out.shape), float(out.float().abs().mean()))
