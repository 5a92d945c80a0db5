"""Shared synthetic inputs for the parity tests (same generators as tests/golden/make_golden.py)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden import make_inputs  # noqa: E402,F401

# BASELINE.json north_star tolerance: max-abs 1e-3 dB and relative 1e-4.  Element-wise relative
# error is meaningless near 0 dB crossings (oracle-fp32 vs oracle-fp64 already reaches 5e-2
# there), so it is applied as |a-b| <= ABS + REL*|b| and as a norm-wise bound.
ABS_DB = 1e-3
REL = 1e-4


def assert_features_close(got, ref, what="", abs_tol=ABS_DB, rel_tol=REL):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    assert np.isfinite(got).all(), f"{what}: non-finite output"
    err = np.abs(got - ref)
    bound = abs_tol + rel_tol * np.abs(ref)
    worst = float((err - bound).max())
    assert worst <= 0, f"{what}: max |err| {err.max():.3e} exceeds {abs_tol}+{rel_tol}*|ref| by {worst:.3e}"
    nrm = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)
    assert nrm <= rel_tol, f"{what}: norm-wise relative error {nrm:.3e} > {rel_tol}"
    return float(err.max())


def aug_case_inputs(g):
    """Rebuild the inputs of tests/golden/aug_cfg2.npz from its recorded seed."""
    seed, B, N, L = int(g["seed"]), int(g["B"]), int(g["N"]), int(g["L"])
    n_noise, n_rir = int(g["n_noise"]), int(g["n_rir"])
    gen = torch.Generator().manual_seed(seed)
    x = 0.1 * torch.randn(B, N, generator=gen)
    noise = [0.05 * torch.randn(N + 777 * i, generator=gen) for i in range(n_noise)]
    t = torch.arange(L, dtype=torch.float32)
    rirs = [torch.randn(L - 100 * i, generator=gen) * torch.exp(-t[:L - 100 * i] / 1000.0) for i in range(n_rir)]
    return x, noise, rirs


def synth_banks(seed, n_noise, noise_len, n_rir, rir_len):
    """BASELINE.md config-2 banks: noise 0.05*randn, RIR randn*exp(-t/1000)."""
    gen = torch.Generator().manual_seed(seed)
    noise = [0.05 * torch.randn(noise_len, generator=gen) for _ in range(n_noise)]
    t = torch.arange(rir_len, dtype=torch.float32)
    rirs = [torch.randn(rir_len, generator=gen) * torch.exp(-t / 1000.0) for _ in range(n_rir)]
    return noise, rirs


# ---- host mirror of the on-GPU augmentation sampler (wwf_loader.cuh: draw_aug_kernel) ----------------
def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on numpy uint32 arrays (Salmon et al., SC'11): returns four uint32 arrays."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32).copy() for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0.astype(np.uint64)
        p1 = M1 * c2.astype(np.uint64)
        n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ np.uint32(k0)
        n1 = (p1 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ np.uint32(k1)
        n3 = (p0 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def philox_draws(seed, first_index, B, n_rir, noise_lens, F, T, nF, nT, rir_prob=0.25, noise_prob=0.5,
                 freq_mask_prob=0.5, time_mask_prob=0.5, snr_range=(5.0, 20.0), freq_mask_param=15, time_mask_param=35,
                 stretch_prob=0.0, stretch_range=(0.8, 1.2), pitch_prob=0.0, pitch_range=(-2, 2)):
    """Bit-exact host recomputation of wwf_draw_aug: dict of numpy arrays named like AugParams."""
    idx = np.uint64(first_index) + np.arange(B, dtype=np.uint64)
    c0 = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    c1 = (idx >> np.uint64(32)).astype(np.uint32)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    blk = lambda j, w=0: philox4x32_10(c0, c1, np.full(B, j, np.uint32), np.full(B, w, np.uint32), k0, k1)
    thr = lambda p: 0 if p <= 0 else (1 << 32) if p >= 1 else int(np.floor(p * 4294967296.0))
    u01 = lambda u: (u >> np.uint32(8)).astype(np.float32) * np.float32(5.9604644775390625e-08)
    pick = lambda u, n: ((u.astype(np.uint64) * np.asarray(n, dtype=np.uint64)) >> np.uint64(32)).astype(np.int64)
    a, b = blk(0), blk(1)
    n_noise = len(noise_lens)
    rir = np.where((n_rir > 0) & (a[0].astype(np.uint64) < thr(rir_prob)), pick(a[1], max(n_rir, 1)), -1).astype(np.int32)
    noi = np.where((n_noise > 0) & (a[2].astype(np.uint64) < thr(noise_prob)), pick(a[3], max(n_noise, 1)), -1).astype(np.int32)
    lens = np.asarray(noise_lens, dtype=np.uint64)[np.maximum(noi, 0)] if n_noise else np.ones(B, np.uint64)
    off = np.where(noi >= 0, pick(b[0], lens), 0).astype(np.int64)
    lo, hi = np.float32(snr_range[0]), np.float32(snr_range[1])
    snr = (lo + ((hi - lo) * u01(b[1])).astype(np.float32)).astype(np.float32)
    fon = b[2].astype(np.uint64) < thr(freq_mask_prob)
    ton = b[3].astype(np.uint64) < thr(time_mask_prob)

    def masks(n, param, size, on, j):
        st, ln = np.zeros((B, n), np.int32), np.zeros((B, n), np.int32)
        for q in range(0, n, 2):
            m = blk(j); j += 1
            for h in range(2):
                if q + h < n:
                    value = (u01(m[2 * h]) * np.float32(param)).astype(np.float32)
                    minv = (u01(m[2 * h + 1]) * (np.float32(size) - value).astype(np.float32)).astype(np.float32)
                    ok = on & (param >= 1)
                    st[:, q + h] = np.where(ok, minv.astype(np.int32), 0)
                    ln[:, q + h] = np.where(ok, value.astype(np.int32), 0)
        return st, ln, j

    fs, fl, j = masks(nF, freq_mask_param, F, fon, 2)
    ts, tl, j = masks(nT, time_mask_param, T, ton, j)
    out = dict(rir_idx=rir, noise_idx=noi, noise_off=off, snr_db=snr, fmask_start=fs, fmask_len=fl, tmask_start=ts, tmask_len=tl)
    # time-stretch / pitch-shift: Philox block with counter word 3 = 1, double arithmetic (each op exactly rounded)
    s = blk(0, 1)
    slo, shi = np.float64(stretch_range[0]), np.float64(stretch_range[1])
    rate = slo + (shi - slo) * (s[1].astype(np.float64) * np.float64(2.3283064365386963e-10))
    out["stretch_rate"] = np.where(s[0].astype(np.uint64) < thr(stretch_prob), rate, 1.0).astype(np.float64)
    plo, phi = int(pitch_range[0]), int(pitch_range[1])
    steps = plo + pick(s[3], max(phi - plo + 1, 1))
    out["pitch_steps"] = np.where(s[2].astype(np.uint64) < thr(pitch_prob), steps, 0).astype(np.int32)
    return out
