"""Kernel times of the configs[1] step (conv_kernel, feat_kernel) as a function of the batch size:
shows the quantisation of clips over the persistent grid (444 CTA slots for feat, 148 for conv)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import wakeword_trainer_home_b200 as w

dev = torch.device("cuda", 0)
plan = w.FeaturePlan(bench.SR, "mfcc", bench.N_MELS, bench.N_MFCC, bench.N_FFT, bench.HOP, dev)
noise, rirs = bench.synth_banks()
plan.register_noise(noise)
plan.register_rirs(rirs)
for B in [int(a) for a in sys.argv[1:]] or [296, 444, 592, 888, 1024, 1184, 1332, 2048]:
    ring = []
    for i in range(4):
        wav, d = bench.synth(i, B)
        ring.append((wav.to(dev), w.AugParams(**d).to(dev)))
    out = None
    for i in range(5):
        out = plan.featurize(*ring[i % 4], out=out)
    plan.profile(True)
    for i in range(20):
        plan.featurize(*ring[i % 4], out=out)
    c, f, n = plan.profile_read()
    plan.profile(False)
    print(f"B={B:5d} conv {c*1e3:7.1f} us  feat {f*1e3:7.1f} us  per clip: conv {c*1e6/B:6.1f} ns feat {f*1e6/B:6.1f} ns  step {(c+f)*1e6/B:6.1f} ns/clip")
