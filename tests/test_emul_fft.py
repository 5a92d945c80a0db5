"""CPU checks of the CUDA kernels' index math: the __host__ __device__ task functions of
csrc/*.cuh are driven sequentially by tests/emul/emul.cu (built here with nvcc as HOST code)
and compared with numpy.  This exercises radix butterflies, in-place passes, digit-reversed
positions, the two-frames-per-FFT split, and the overlap-save pair pass - not the product path."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "emul.cu")
LIB = os.path.join(HERE, "emul", "libwwf_emul.so")
CSRC = os.path.join(HERE, "..", "wakeword_trainer_home_b200", "csrc")
fp = ctypes.POINTER(ctypes.c_float)


def P(a):
    return a.ctypes.data_as(fp)


@pytest.fixture(scope="module")
def emul():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.run([nvcc, "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets",
                        "-Xcompiler", "-fPIC", "-shared", "-o", LIB, SRC], check=True)
    return ctypes.CDLL(LIB)


@pytest.mark.parametrize("n", [256, 400, 512, 1024, 2048])
def test_complex_fft_and_pair_split(emul, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(2 * n).astype(np.float32)
    out = np.zeros(2 * n, np.float32)
    assert emul.emul_cfft(n, P(x), P(out)) == 0
    ref = np.fft.fft(x[0::2].astype(np.float64) + 1j * x[1::2])
    assert np.abs((out[0::2] + 1j * out[1::2]) - ref).max() <= 4e-7 * np.abs(ref).max()
    fa, fb = rng.standard_normal(n).astype(np.float32), (1e-3 * rng.standard_normal(n)).astype(np.float32)
    pa, pb = np.zeros(n // 2 + 1, np.float32), np.zeros(n // 2 + 1, np.float32)
    assert emul.emul_stft_pair(n, P(fa), P(fb), P(pa), P(pb)) == 0
    ra = np.abs(np.fft.rfft(fa.astype(np.float64))) ** 2
    rb = np.abs(np.fft.rfft(fb.astype(np.float64))) ** 2
    assert np.abs(pa - ra).max() <= 1e-6 * ra.max()
    # the quiet frame shares an FFT with a 60 dB louder one: its error is relative to the LOUD frame
    assert np.abs(pb - rb).max() <= 1e-6 * np.sqrt(ra.max() * rb.max())


def test_pair_tasks_cover_every_bin_once_without_bad_conflicts(emul):
    assert emul.emul_pair_task_coverage() == 0
    assert emul.emul_pair_bank_conflicts() == 1


def test_conv_middle_phases_are_warp_local(emul):
    # conv_kernel separates forward pass 1, the fused run pairs and inverse pass 1 by __syncwarp() only
    assert emul.emul_conv_warp_locality() == 0


@pytest.mark.parametrize("N,L,lmax", [(24000, 8000, 8000), (16000, 500, 8000), (40000, 8000, 8000),
                                      (70000, 16384, 16384), (1000, 3, 3), (24769, 8000, 8000), (24770, 8000, 8000)])
def test_overlap_save_convolution(emul, N, L, lmax):
    rng = np.random.default_rng(N + L)
    x = rng.standard_normal(N).astype(np.float32)
    h = (rng.standard_normal(L) * np.exp(-np.arange(L) / 1000.0)).astype(np.float32)
    y = np.zeros(N, np.float32)
    nb = emul.emul_rir_conv(P(x), N, P(h), L, lmax, P(y))
    assert nb >= 1
    n = N + L - 1
    ref = np.fft.irfft(np.fft.rfft(x.astype(np.float64), n) * np.fft.rfft(h.astype(np.float64), n), n)[:N]
    assert np.abs(y - ref).max() <= 1e-6 * np.abs(ref).max()
    assert (nb == 1) == (N + lmax - 1 <= 32768)


def test_reflect_index_matches_numpy_pad(emul):
    N, pad = 37, 12
    ref = np.pad(np.arange(N), (pad, pad), mode="reflect")
    got = [emul.emul_reflect_index(i, N) for i in range(-pad, N + pad)]
    assert list(ref) == got


def test_c_filterbank_close_to_torchaudio(emul):
    import torchaudio.functional as AF
    for n_freqs, n_mels in ((201, 40), (513, 128)):
        out = np.zeros((n_freqs, n_mels), np.float32)
        emul.emul_mel_fbanks(n_freqs, ctypes.c_float(0.0), ctypes.c_float(8000.0), n_mels, 16000, P(out))
        ref = AF.melscale_fbanks(n_freqs, 0.0, 8000.0, n_mels, 16000).numpy()
        assert np.abs(out - ref).max() <= 2e-5


def test_phase_vocoder_analysis_synthesis_pair(emul):
    """pv_stft_kernel's split and pv_istft_kernel's Hermitian pack + reverse-order inverse passes."""
    rng = np.random.default_rng(7)
    fa, fb = rng.standard_normal(512).astype(np.float32), rng.standard_normal(512).astype(np.float32)
    spa, spb = np.zeros(514, np.float32), np.zeros(514, np.float32)
    ya, yb = np.zeros(512, np.float32), np.zeros(512, np.float32)
    assert emul.emul_pv_roundtrip(P(fa), P(fb), P(spa), P(spb), P(ya), P(yb)) == 0
    for sp, f in ((spa, fa), (spb, fb)):
        ref = np.fft.rfft(f.astype(np.float64))
        assert np.abs((sp[0::2] + 1j * sp[1::2]) - ref).max() <= 5e-7 * np.abs(ref).max()
    assert np.abs(ya - fa).max() <= 2e-6 and np.abs(yb - fb).max() <= 2e-6


@pytest.mark.parametrize("n_fft,n_mels", [(400, 40), (400, 128), (256, 40), (512, 64), (1024, 128), (1024, 80), (2048, 128), (400, 13), (2048, 32),
                                          (2048, 80), (2048, 20), (2048, 40), (1024, 40), (512, 20), (512, 128), (256, 64), (256, 20), (400, 64), (400, 80), (400, 20)])
def test_mel_lane_schedule_equals_dense_filterbank(emul, n_fft, n_mels):
    """The lane schedule of the sparse mel projection (wwf_tables.h: filters cut in halves, sorted into rounds of 32
    lanes, weights interleaved) visits every non-zero of torchaudio's filterbank exactly once: evaluated on the CPU in
    the kernel's order it equals the dense matmul, every filter has one owner lane, split halves sit on adjacent lanes."""
    import torchaudio.functional as AF
    K = n_fft // 2 + 1
    fb = AF.melscale_fbanks(K, 0.0, 8000.0, n_mels, 16000).numpy().astype(np.float32)
    rng = np.random.default_rng(n_fft + n_mels)
    power = (rng.standard_normal(K) ** 2).astype(np.float32)
    out = np.zeros(n_mels, np.float32)
    stats = (ctypes.c_int * 4)()
    rc = emul.emul_mel_schedule(n_fft, K, n_mels, P(np.ascontiguousarray(fb)), P(power), P(out), stats)
    assert rc == 0
    ref = power.astype(np.float64) @ fb.astype(np.float64)
    assert np.abs(out - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())
    rounds, iters, worst, nsplit = list(stats)
    naive = sum(-(-max((int(np.count_nonzero(fb[:, m])) and (np.nonzero(fb[:, m])[0][-1] - np.nonzero(fb[:, m])[0][0] + 1))
                       for m in range(r, min(r + 32, n_mels))) // 2) for r in range(0, n_mels, 32))
    assert iters <= naive + 1, (iters, naive)     # (a round may grow by one iteration to free a bank residue)
    assert worst <= 2, worst          # conflict-free by construction, except first bins too small to be shifted
    print(f"n_fft {n_fft} n_mels {n_mels}: rounds {rounds}, two-tap iterations {iters} (filter-per-lane order: {naive}), "
          f"worst half-warp residue multiplicity {worst}, split filters {nsplit}")
