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
