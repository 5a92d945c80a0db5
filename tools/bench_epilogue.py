"""Per-kernel times (library CUDA-event hook) of the flat path for the two MFCC shapes with a specialised tensor-core
epilogue: 40 mels x 40 coefficients (BASELINE configs[1], no augmentation here) and the reference's DataConfig defaults
(128 mels x 40 coefficients, n_fft 1024, 2.5 s)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wakeword_trainer_home_b200 as w

for (M, C, nfft, N) in ((40, 40, 400, 24000), (128, 40, 1024, 40000), (64, 13, 512, 24000)):
    plan = w.FeaturePlan(16000, "mfcc", M, C, nfft, 160, "cuda")
    plan.set_path("flat")
    x = (0.1 * torch.randn(1024, N, generator=torch.Generator().manual_seed(0))).cuda()
    out = plan.featurize(x)
    for _ in range(3):
        plan.featurize(x, out=out)
    torch.cuda.synchronize()
    plan.profile(True)
    for _ in range(20):
        plan.featurize(x, out=out)
    ms, n, nflat = plan.profile_read_kernels()
    plan.profile(False)
    print(json.dumps({"n_mels": M, "n_mfcc": C, "n_fft": nfft, "N": N, "kernel_ms": {k: round(v, 4) for k, v in ms.items()}}), flush=True)
