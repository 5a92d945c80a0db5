"""Every case of the fuzz test with shared memory and freshly allocated global memory poisoned with NaN before each
call: a kernel that reads memory it (or a predecessor of the same call) did not write shows up as NaN features."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import wakeword_trainer_home_b200 as ww
from wakeword_trainer_home_b200 import _native
lib = _native.load()
rng = np.random.default_rng(20261018)
gen = torch.Generator().manual_seed(20261018)
def poison():
    junk = [torch.full((1 << 24,), float("nan"), device="cuda") for _ in range(8)]
    del junk
    _native.check(lib.wwf_debug_poison_smem(0, 0x7fc00000))
for case in range(28):
    n_fft = int(rng.choice([256, 400, 512, 1024, 2048]))
    hop = int(rng.choice([h for h in (80, 100, 128, 160, 160, 160, 200, 256, 512) if h < n_fft]))
    M = int(rng.choice([m for m in (20, 40, 64, 80, 128) if m <= n_fft // 4]))
    ftype = str(rng.choice(["mel", "mfcc"]))
    C = int(rng.integers(1, min(M, 40) + 1))
    N = int(rng.integers(n_fft // 2 + 1, 3 * n_fft)) if case % 7 == 0 else int(rng.integers(4000, 50000))
    B = int(rng.integers(1, 6))
    use_aug, use_masks, use_cmvn, f16 = (bool(rng.integers(0, 2)) for _ in range(4))
    F, T = (C if ftype == "mfcc" else M), N // hop + 1
    if 4 * M * (T | 1) + (4 * F * (T | 1) if (use_cmvn and ftype == "mfcc") else 0) > 150_000:
        use_cmvn = False
    x = 0.1 * torch.randn(B, N, generator=gen)
    plan = ww.FeaturePlan(16000, ftype, M, C, n_fft, hop, "cuda", cmvn=use_cmvn, out_dtype=torch.float16 if f16 else torch.float32,
                          n_freq_masks=2 if use_masks else 0, n_time_masks=1 if use_masks else 0, mask_value=-3.0)
    ap = ww.AugParams()
    if use_aug:
        noise = [0.05 * torch.randn(int(rng.integers(500, 60000)), generator=gen) for _ in range(3)]
        rirs = [torch.randn(int(rng.integers(1, 9000)), generator=gen) * 0.3 for _ in range(3)]
        plan.register_noise(noise); plan.register_rirs(rirs)
        ap.rir_idx = torch.from_numpy(rng.integers(-1, 3, B).astype(np.int32))
        ap.noise_idx = torch.from_numpy(rng.integers(-1, 3, B).astype(np.int32))
        ap.noise_off = torch.from_numpy(rng.integers(0, 500, B).astype(np.int64))
        ap.snr_db = torch.from_numpy(rng.uniform(0, 25, B).astype(np.float32))
    if use_masks:
        ap.fmask_start, ap.fmask_len = ww.draw_mask_params(gen, B, F, min(15, F), 2)
        ap.tmask_start, ap.tmask_len = ww.draw_mask_params(gen, B, T, min(35, T), 1)
    what = f"case {case}: n_fft={n_fft} hop={hop} M={M} {ftype} C={C} N={N} B={B} aug={use_aug} masks={use_masks} cmvn={use_cmvn} f16={f16}"
    res = []
    for path in ("fused", "flat"):
        try:
            plan.set_path(path)
            xc = x.cuda()
            clean = plan.featurize(xc, ap).float().cpu().numpy()
            plan.release_workspaces()
            poison()
            got = plan.featurize(xc, ap).float().cpu().numpy()
            res.append(f"{path}: nan {int(np.isnan(got).sum())}/{got.size} clean-nan {int(np.isnan(clean).sum())} diff {float(np.nanmax(np.abs(got - clean))):.3g}")
        except ww.WwfError as e:
            res.append(f"{path}: refused")
    print(what, "|", " | ".join(res), flush=True)
