"""Small end-to-end case for compute-sanitizer: every kernel of the library once, small shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wakeword_trainer_home_b200 as w
gen = torch.Generator().manual_seed(0)
for (ftype, n_fft, M, C, N, hop, cm) in (("mfcc", 400, 40, 40, 24000, 160, False), ("mel", 1024, 128, 40, 16000, 160, True),
                                          ("mfcc", 512, 64, 32, 12345, 100, True), ("mel", 2048, 128, 40, 40000, 512, False),
                                          ("mel", 256, 40, 13, 16000, 128, False)):
    B = 5
    x = 0.1 * torch.randn(B, N, generator=gen)
    noise = [0.05 * torch.randn(7001, generator=gen), 0.05 * torch.randn(N + 333, generator=gen)]
    t = torch.arange(3000, dtype=torch.float32)
    rirs = [torch.randn(3000, generator=gen) * torch.exp(-t / 500.0), torch.randn(50, generator=gen)]
    plan = w.FeaturePlan(16000, ftype, M, C, n_fft, hop, "cuda", cmvn=cm, n_freq_masks=2, n_time_masks=2)
    plan.register_noise(noise); plan.register_rirs(rirs)
    T = N // hop + 1
    fs, fl = w.draw_mask_params(gen, B, plan.n_feat, 15, 2); ts, tl = w.draw_mask_params(gen, B, T, 35, 2)
    p = w.AugParams(rir_idx=torch.tensor([0, 1, -1, 0, 1]), noise_idx=torch.tensor([0, 1, 1, -1, 0]),
                    noise_off=torch.tensor([6999, 5, 100, 0, 3000]), snr_db=torch.full((B,), 8.0),
                    fmask_start=fs, fmask_len=fl, tmask_start=ts, tmask_len=tl)
    f = plan.featurize(x.cuda(), p)
    a = plan.augment(x.cuda(), p)
    torch.cuda.synchronize()
    print(ftype, n_fft, tuple(f.shape), bool(torch.isfinite(f).all()), bool(torch.isfinite(a).all()))
sa = w.SpecAugment(15, 35, 2, 2)
print(tuple(sa(torch.randn(2, 1, 64, 50).cuda()).shape))
