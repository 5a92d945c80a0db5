"""Throughput of less common feature configs (other n_fft / hop): looks for anomalies, not a headline."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wakeword_trainer_home_b200 as w
dev = torch.device("cuda", 0)
gen = torch.Generator().manual_seed(0)
x = (0.1 * torch.randn(1024, 24000, generator=gen)).to(dev)
for (n_fft, hop, M, ftype) in ((256, 160, 40, "mel"), (256, 128, 40, "mel"), (512, 160, 64, "mel"), (512, 256, 64, "mfcc"),
                               (1024, 160, 128, "mel"), (1024, 256, 128, "mel"), (2048, 160, 128, "mel"), (2048, 512, 128, "mel"),
                               (400, 160, 40, "mel"), (400, 200, 40, "mel"), (400, 100, 40, "mel")):
    plan = w.FeaturePlan(16000, ftype, M, 32, n_fft, hop, dev)
    out = None
    for i in range(3):
        out = plan.featurize(x, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        out = plan.featurize(x, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    T = 24000 // hop + 1
    print(json.dumps({"n_fft": n_fft, "hop": hop, "n_mels": M, "type": ftype, "ms": round(ms, 4), "clips_per_s": round(1024 / ms * 1e3),
                      "ns_per_frame_fft": round(ms * 1e6 / (1024 * T), 1)}), flush=True)
