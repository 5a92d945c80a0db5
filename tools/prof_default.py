"""ncu driver for the reference-default feature config (DataConfig: n_fft 1024, hop 160, 128 mels, 2.5 s)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wakeword_trainer_home_b200 as w
plan = w.FeaturePlan(16000, "mel", 128, 40, 1024, 160, "cuda")
x = (0.1 * torch.randn(1024, 40000, generator=torch.Generator().manual_seed(0))).cuda()
out = None
for i in range(3):
    out = plan.featurize(x, out=out)
torch.cuda.synchronize()
print("ok", tuple(out.shape))
