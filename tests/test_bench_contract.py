"""bench.py's host-side contract, checked without a GPU: the workloads are BASELINE.json's configs, the algorithmic
byte counts are SURVEY.md section 8d's, both arms print the same `config` object, the synthetic draws are deterministic
and the reference arm (the oracle on the host cores) emits a complete JSON line."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_workloads_are_the_baseline_configs():
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert len(base["configs"]) == 5
    w2 = bench.WORKLOADS["cfg2"]                               # configs[1]: the config the metric is quoted on
    assert (w2.feature_type, w2.n_fft, w2.hop, w2.n_mels, w2.n_mfcc, w2.n_samples, w2.batch) == ("mfcc", 400, 160, 40, 40, 24000, 1024)
    assert w2.rir_prob == 1.0 and w2.noise_prob == 1.0 and w2.T == 151
    b = w2.bytes_per_clip()
    assert b["step"] == 248160 and b["conv_kernel"] == 224000 and b["feat_frames_kernel"] == 216160
    w1 = bench.WORKLOADS["cfg1"]
    assert w1.batch == 64 and w1.bytes_per_clip()["step"] == 120160
    assert bench.WORKLOADS["cfg5"].bytes_per_clip()["step"] == 160160
    w4 = bench.WORKLOADS["cfg4"]
    assert w4.f16 and w4.n_mels == 64 and w4.n_samples == 32000 and w4.out_bytes == 2 * 64 * 201
    w3 = bench.cfg3_workload()
    assert abs(w3.bytes_per_clip()["step"] - 229312) < 1                                                  # SURVEY 8d row 3 (1.5 s)
    for wl in list(bench.WORKLOADS.values()) + [w3]:
        assert wl.ring * wl.batch * wl.n_samples * 4 > 126e6                                              # inputs larger than L2


def test_both_arms_print_the_same_config_and_draws_are_deterministic():
    for key, wl in bench.WORKLOADS.items():
        assert bench.config_dict(wl, 4) == bench.config_dict(wl, 4)
        assert bench.config_dict(wl, 1)["batch_per_gpu"] == wl.batch
        wav_a, d_a = bench.synth(5, 16, wl)
        wav_b, d_b = bench.synth(5, 16, wl)
        assert torch.equal(wav_a, wav_b) and d_a.keys() == d_b.keys()
        for k in d_a:
            assert torch.equal(d_a[k], d_b[k])
        if wl.rir_prob > 0:
            assert d_a["rir_idx"].dtype == torch.int32 and int(d_a["rir_idx"].max()) < wl.n_rir
        if 0 < wl.rir_prob < 1:
            assert (bench.synth(6, 512, wl)[1]["rir_idx"] < 0).any()                                       # some clips stay dry
        if wl.mask_prob > 0:
            assert d_a["fmask_start"].shape == (16, 2) and d_a["tmask_len"].shape == (16, 2)


def test_reference_arm_prints_one_complete_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg1", "--steps", "2",
                        "--warmup", "3"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "clips/s" and d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert d["config"] == bench.config_dict(bench.WORKLOADS["cfg1"], 1)
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["value"] > 0 and d["steps"] == 2
