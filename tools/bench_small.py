"""Small-batch latency of the feature stage under both forced paths (device time per call, CUDA events)."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CASES = [("mfcc", 128, 40, 1024, 160, 40000), ("mel", 128, 40, 1024, 160, 40000), ("mfcc", 40, 40, 400, 160, 24000), ("mel", 64, 40, 512, 160, 32000)]
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import wakeword_trainer_home_b200 as w
    for (ft, M, C, nfft, hop, N) in CASES:
        plan = w.FeaturePlan(16000, ft, M, C, nfft, hop, "cuda")
        for B in (1, 8, 32, 64, 128):
            x = (0.1 * torch.randn(B, N, generator=torch.Generator().manual_seed(0))).cuda()
            out = plan.featurize(x)
            for _ in range(5):
                plan.featurize(x, out=out)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(50):
                plan.featurize(x, out=out)
            b.record()
            torch.cuda.synchronize()
            print(f"{ft} M={M} n_fft={nfft} N={N} B={B}: {a.elapsed_time(b) / 50 * 1e3:7.1f} us")
else:
    for path in ("fused", "split"):
        env = dict(os.environ, WWF_FEAT_PATH=path)
        print("==", path)
        print(subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True).stdout)
