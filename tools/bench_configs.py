"""Device-resident throughput of the other BASELINE.json configs (not the headline bench): ms per batch, clips/s."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wakeword_trainer_home_b200 as w

dev = torch.device("cuda", 0)
gen = torch.Generator().manual_seed(0)

def banks(n_noise, noise_len, n_rir, rir_len):
    noise = [0.05 * torch.randn(noise_len, generator=gen) for _ in range(n_noise)]
    t = torch.arange(rir_len, dtype=torch.float32)
    rirs = [torch.randn(rir_len, generator=gen) * torch.exp(-t / 1000.0) for _ in range(n_rir)]
    return noise, rirs

def run(name, B, N, ftype, n_mels, n_mfcc, n_fft, p_noise, p_rir, masks, out_dtype=torch.float32, iters=30):
    plan = w.FeaturePlan(16000, ftype, n_mels, n_mfcc, n_fft, 160, dev, out_dtype=out_dtype,
                         n_freq_masks=2 if masks else 0, n_time_masks=2 if masks else 0)
    aug = None
    if p_noise > 0 or p_rir > 0 or masks:
        noise, rirs = banks(64, max(N, 24000), 16, 8000)
        plan.register_noise(noise); plan.register_rirs(rirs)
        T = N // 160 + 1
        on_r = torch.rand(B, generator=gen) < p_rir
        on_n = torch.rand(B, generator=gen) < p_noise
        d = dict(rir_idx=torch.where(on_r, torch.randint(0, 16, (B,), generator=gen, dtype=torch.int32), torch.tensor(-1, dtype=torch.int32)),
                 noise_idx=torch.where(on_n, torch.randint(0, 64, (B,), generator=gen, dtype=torch.int32), torch.tensor(-1, dtype=torch.int32)),
                 noise_off=torch.randint(0, N, (B,), generator=gen), snr_db=5 + 15 * torch.rand(B, generator=gen))
        if masks:
            fs, fl = w.draw_mask_params(gen, B, plan.n_feat, 15, 2); ts, tl = w.draw_mask_params(gen, B, T, 35, 2)
            d.update(fmask_start=fs, fmask_len=fl, tmask_start=ts, tmask_len=tl)
        aug = w.AugParams(**d).to(dev)
    ring = [(0.1 * torch.randn(B, N, generator=gen)).to(dev) for _ in range(max(2, min(8, int(300e6 / (B * N * 4)) + 1)))]
    out = None
    for i in range(5):
        out = plan.featurize(ring[i % len(ring)], aug, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        out = plan.featurize(ring[i % len(ring)], aug, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(json.dumps({"config": name, "B": B, "N": N, "ms_per_batch": round(ms, 4), "clips_per_s": round(B / ms * 1e3)}), flush=True)

run("cfg1 log-mel40 n400 B=64", 64, 24000, "mel", 40, 40, 400, 0, 0, False)
run("cfg1 log-mel40 n400 B=1024", 1024, 24000, "mel", 40, 40, 400, 0, 0, False)
run("cfg3 default preset n1024 M128 1.5s B=128", 128, 24000, "mel", 128, 40, 1024, 0.5, 0.25, True)
run("cfg3 default preset n1024 M128 2.5s B=128", 128, 40000, "mel", 128, 40, 1024, 0.5, 0.25, True)
run("cfg3 default preset n1024 M128 1.5s B=1024", 1024, 24000, "mel", 128, 40, 1024, 0.5, 0.25, True)
run("cfg4 edge n400 M64 fp16 2s B=1024", 1024, 32000, "mel", 64, 40, 400, 0.5, 0.3, False, torch.float16)
run("cfg5 sweep n400 M40 2s B=256", 256, 32000, "mel", 40, 40, 400, 0, 0, False)
run("cfg5 sweep n400 M40 2s B=8192", 8192, 32000, "mel", 40, 40, 400, 0, 0, False)
run("ref-default mfcc n1024 M128 C40 2.5s B=1024", 1024, 40000, "mfcc", 128, 40, 1024, 0, 0, False)
run("isolate: mfcc40 n400 no aug B=1024", 1024, 24000, "mfcc", 40, 40, 400, 0, 0, False)
run("isolate: log-mel40 n400 + noise B=1024", 1024, 24000, "mel", 40, 40, 400, 1.0, 0, False)
run("isolate: mfcc40 n400 + noise B=1024", 1024, 24000, "mfcc", 40, 40, 400, 1.0, 0, False)
