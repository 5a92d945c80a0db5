"""Host-side cost of one featurize call (wall clock, small batches): is the Python/ctypes shim the limit?"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wakeword_trainer_home_b200 as w
dev = torch.device("cuda", 0)
gen = torch.Generator().manual_seed(0)
for (B, N, n_fft, M, ftype, aug_on) in ((1, 24000, 400, 40, "mel", False), (32, 40000, 1024, 128, "mel", False),
                                        (128, 40000, 1024, 128, "mel", True), (128, 24000, 400, 40, "mfcc", True)):
    plan = w.FeaturePlan(16000, ftype, M, 40, n_fft, 160, dev, n_freq_masks=2, n_time_masks=2)
    aug = None
    if aug_on:
        plan.register_noise([0.05 * torch.randn(N, generator=gen) for _ in range(8)])
        t = torch.arange(8000, dtype=torch.float32)
        plan.register_rirs([torch.randn(8000, generator=gen) * torch.exp(-t / 1000.0) for _ in range(4)])
        T = N // 160 + 1
        fs, fl = w.draw_mask_params(gen, B, plan.n_feat, 15, 2); ts, tl = w.draw_mask_params(gen, B, T, 35, 2)
        aug = w.AugParams(rir_idx=torch.randint(-1, 4, (B,), generator=gen, dtype=torch.int32),
                          noise_idx=torch.randint(-1, 8, (B,), generator=gen, dtype=torch.int32),
                          noise_off=torch.randint(0, N, (B,), generator=gen), snr_db=5 + 15 * torch.rand(B, generator=gen),
                          fmask_start=fs, fmask_len=fl, tmask_start=ts, tmask_len=tl).to(dev)
    x = (0.1 * torch.randn(B, N, generator=gen)).to(dev)
    out = plan.featurize(x, aug)
    torch.cuda.synchronize()
    n = 300
    t0 = time.perf_counter()
    for _ in range(n):
        plan.featurize(x, aug, out=out)
    t_issue = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    t_total = (time.perf_counter() - t0) / n
    print(json.dumps({"B": B, "N": N, "n_fft": n_fft, "aug": aug_on, "host_issue_us": round(t_issue * 1e6, 1),
                      "wall_us_per_call": round(t_total * 1e6, 1), "clips_per_s_wall": round(B / t_total)}), flush=True)
