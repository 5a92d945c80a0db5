"""Per-kernel times of the bench workload (BASELINE.json configs[1]) on both launch shapes, through the library's
CUDA-event hook (wwf_profile_*): where does a step go, and what does the flat path's bookkeeping cost conv_kernel?"""
import json
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
batches = [bench.synth(i, bench.B_PER_GPU) for i in range(4)]
wavs = [b[0].to(dev) for b in batches]
augs = [w.AugParams(**b[1]).to(dev) for b in batches]
out = None
for path in ("flat", "fused", "flat"):
    plan.set_path(path)
    for i in range(5):
        out = plan.featurize(wavs[i % 4], augs[i % 4], out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30):
        plan.featurize(wavs[i % 4], augs[i % 4], out=out)
    e1.record()
    torch.cuda.synchronize()
    plan.profile(True)
    for i in range(30):
        plan.featurize(wavs[i % 4], augs[i % 4], out=out)
    ms, n, nflat = plan.profile_read_kernels()
    plan.profile(False)
    print(json.dumps({"path": path, "step_ms": e0.elapsed_time(e1) / 30, "kernel_ms": {k: round(v, 4) for k, v in ms.items()}}), flush=True)
