"""Generate the golden vectors under tests/golden/ by calling torchaudio DIRECTLY.

Run in the build container (torchaudio 2.11.0+cu128, CPU):  python tests/golden/make_golden.py
The reference's own hot-path module (src/data) is absent from its checkout and its tests
hold no vectors (SURVEY.md sections 0 and 4), so these pin the oracle - and through it the CUDA
path - to the dependency that defines the arithmetic.  Inputs are regenerated from the
recorded seeds; only outputs (and the small banks) are stored.
"""
import os
import numpy as np
import torch
import torchaudio
import torchaudio.functional as AF
import torchaudio.transforms as AT

HERE = os.path.dirname(os.path.abspath(__file__))


def make_inputs(seed, B, N):
    g = torch.Generator().manual_seed(seed)
    x = 0.1 * torch.randn(B, N, generator=g)
    x[1] = 0.0                                               # digital silence -> -100 dB everywhere
    t = torch.arange(N) / 16000.0
    x[2] = 0.9 * torch.sin(2 * torch.pi * 440.0 * t)          # full-scale tone -> top_db floor active
    x[3, N // 2:] = 0.0                                      # half silent -> floor inside a clip
    return x


def feature_case(name, seed, B, N, n_fft, hop, n_mels, n_mfcc):
    x = make_inputs(seed, B, N)
    mel = AT.MelSpectrogram(16000, n_fft=n_fft, hop_length=hop, n_mels=n_mels)
    db = AT.AmplitudeToDB("power", top_db=80.0)
    logmel = db(mel(x.unsqueeze(1)))                          # (B,1,M,T) per-clip top_db
    mf = AT.MFCC(16000, n_mfcc=n_mfcc, norm="ortho",
                 melkwargs=dict(n_fft=n_fft, hop_length=hop, n_mels=n_mels))(x.unsqueeze(1))
    # float64 run of the same transforms = "true value"
    mel64 = AT.MelSpectrogram(16000, n_fft=n_fft, hop_length=hop, n_mels=n_mels).double()
    logmel64 = db(mel64(x.double().unsqueeze(1)))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), seed=seed, B=B, N=N, n_fft=n_fft, hop=hop,
                        n_mels=n_mels, n_mfcc=n_mfcc, logmel=logmel.numpy(), mfcc=mf.numpy(),
                        logmel64=logmel64.numpy().astype(np.float64),
                        fb=mel.mel_scale.fb.numpy(), torchaudio=torchaudio.__version__)
    print(name, tuple(logmel.shape), tuple(mf.shape))


def aug_case(name, seed, B, N, L, n_noise, n_rir):
    g = torch.Generator().manual_seed(seed)
    x = 0.1 * torch.randn(B, N, generator=g)
    noise = [0.05 * torch.randn(N + 777 * i, generator=g) for i in range(n_noise)]
    t = torch.arange(L, dtype=torch.float32)
    rirs = [torch.randn(L - 100 * i, generator=g) * torch.exp(-t[:L - 100 * i] / 1000.0) for i in range(n_rir)]
    rir_idx = torch.randint(-1, n_rir, (B,), generator=g)
    noise_idx = torch.randint(-1, n_noise, (B,), generator=g)
    noise_off = torch.randint(0, N, (B,), generator=g)
    snr = 5.0 + 15.0 * torch.rand(B, generator=g)
    y = x.clone()
    for b in range(B):
        if rir_idx[b] >= 0:
            y[b] = AF.fftconvolve(x[b], rirs[rir_idx[b]], "full")[:N]
    rev = y.clone()
    for b in range(B):
        if noise_idx[b] >= 0:
            src = noise[noise_idx[b]]
            nz = src[(noise_off[b] + torch.arange(N)) % src.shape[0]]
            y[b] = AF.add_noise(y[b:b + 1], nz.unsqueeze(0), snr[b:b + 1])[0]
    mf = AT.MFCC(16000, n_mfcc=40, norm="ortho", melkwargs=dict(n_fft=400, hop_length=160, n_mels=40))(y.unsqueeze(1))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), seed=seed, B=B, N=N, L=L, n_noise=n_noise, n_rir=n_rir,
                        rir_idx=rir_idx.numpy(), noise_idx=noise_idx.numpy(), noise_off=noise_off.numpy(),
                        snr=snr.numpy(), reverb=rev.numpy(), mixed=y.numpy(), mfcc=mf.numpy(),
                        torchaudio=torchaudio.__version__)
    print(name, tuple(y.shape), tuple(mf.shape))


def mask_case(name, seed):
    torch.manual_seed(seed)
    spec = torch.randn(4, 1, 64, 50)
    fstart = torch.tensor([[3, 40], [0, 60], [10, 10], [63, 5]], dtype=torch.int32)
    flen = torch.tensor([[5, 14], [0, 4], [3, 7], [1, 0]], dtype=torch.int32)
    tstart = torch.tensor([[0, 20], [49, 3], [7, 30], [15, 15]], dtype=torch.int32)
    tlen = torch.tensor([[10, 34], [1, 0], [2, 20], [5, 9]], dtype=torch.int32)
    out = spec.clone()
    for b in range(4):
        for i in range(2):
            out[b, :, fstart[b, i]:fstart[b, i] + flen[b, i], :] = 0.0
            out[b, :, :, tstart[b, i]:tstart[b, i] + tlen[b, i]] = 0.0
    np.savez_compressed(os.path.join(HERE, name + ".npz"), seed=seed, fstart=fstart.numpy(), flen=flen.numpy(),
                        tstart=tstart.numpy(), tlen=tlen.numpy(), out=out.numpy())
    print(name, tuple(out.shape))


def shape_aug_case(name, seed, B, N):
    """Row A3: torchaudio's own stretch stage (F.functional._stretch_waveform), F.pitch_shift and F.resample,
    float32 and float64.  The float32 stretch is only reproducible to ~1e-3 (float32 phase cumsum), the
    float64 one to ~1e-9: tests compare tightly with the float64 vectors."""
    import math
    g = torch.Generator().manual_seed(seed)
    x = 0.1 * torch.randn(B, N, generator=g)
    t = torch.arange(N) / 16000.0
    x[1] = 0.5 * torch.sin(2 * torch.pi * 440.0 * t)
    steps = [-2, -1, 1, 2][:B]
    rates = [2.0 ** (-float(n) / 12) for n in steps]
    lens = [int(round(N / r)) for r in rates]
    st32 = np.zeros((B, max(lens)), np.float32)
    st64 = np.zeros((B, max(lens)), np.float64)
    ps32 = np.zeros((B, N), np.float32)
    ps64 = np.zeros((B, N), np.float64)
    for b, n in enumerate(steps):
        st32[b, :lens[b]] = AF.functional._stretch_waveform(x[b:b + 1], n)[0].numpy()
        st64[b, :lens[b]] = AF.functional._stretch_waveform(x[b:b + 1].double(), n)[0].numpy()
        ps32[b] = AF.pitch_shift(x[b:b + 1], 16000, n)[0].numpy()
        ps64[b] = AF.pitch_shift(x[b:b + 1].double(), 16000, n)[0].numpy()
    rs = {f"rs_{o}": AF.resample(x[:2], o, 16000).numpy() for o in (44100, 8000, 17959)}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), seed=seed, B=B, N=N, steps=np.array(steps), rates=np.array(rates),
                        lens=np.array(lens), stretch32=st32, stretch64=st64, pitch32=ps32, pitch64=ps64,
                        torchaudio=torchaudio.__version__, **rs)
    print(name, st32.shape, ps32.shape)


if __name__ == "__main__":
    torch.set_num_threads(1)   # deterministic reductions
    feature_case("feat_cfg1", 0, 6, 24000, 400, 160, 40, 40)          # BASELINE.json configs[0] shape
    feature_case("feat_refdefault", 1, 4, 16000, 1024, 160, 128, 40)  # DataConfig defaults, 1.0 s
    feature_case("feat_n512", 2, 4, 19200, 512, 160, 64, 32)          # Edge-like: 64 mels / 32 mfcc
    aug_case("aug_cfg2", 3, 6, 24000, 8000, 3, 3)                     # BASELINE.json configs[1] shape
    shape_aug_case("shape_aug", 5, 4, 6000)                           # stretch / pitch / resample
    mask_case("mask_ref", 4)                                          # tests/test_training_pipeline.py:252-262 shape
