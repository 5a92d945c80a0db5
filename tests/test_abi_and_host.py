"""CPU tests of the drop-in boundary and the host logic: the C-ABI library loads and exports
every symbol include/wwfeat.h declares, validates arguments before touching CUDA, refuses to
run without a GPU (no CPU fallback), and the Python shim's constants / draws / sharding are right."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    import __graft_entry__ as ge
    ge.build()                                   # nvcc cross-compiles for sm_100a without a GPU
    from wakeword_trainer_home_b200 import _native
    _native.load()
    return _native


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "wwfeat.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(wwf_[a-z_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(native):
    lib = C.CDLL(native.LIB_PATH)
    names = _header_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/wwfeat.h but not exported"
    assert set(names) == set(native.SYMBOLS), "ctypes binding and header disagree"
    assert native.load().wwf_version() == 130


def test_library_is_sm100a_native_code(native):
    out = subprocess.run(["cuobjdump", "-lelf", native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def _cfg(native, **kw):
    d = dict(sample_rate=16000, n_fft=400, hop_length=160, n_mels=40, n_mfcc=40, feature_type=0, out_dtype=0,
             cmvn=0, top_db=80.0, f_min=0.0, f_max=0.0, cmvn_eps=1e-5, mask_value=0.0, n_freq_masks=0, n_time_masks=0)
    d.update(kw)
    return native.Config(**d)


@pytest.mark.parametrize("kw,code", [
    (dict(n_fft=300), -2), (dict(n_fft=4096), -2), (dict(hop_length=0), -1), (dict(hop_length=400), -1),
    (dict(n_mels=0), -1), (dict(n_mels=129), -1), (dict(feature_type=1, n_mfcc=41), -1), (dict(feature_type=7), -1),
    (dict(out_dtype=3), -1), (dict(n_freq_masks=9), -1), (dict(sample_rate=0), -1),
])
def test_plan_create_validates_before_cuda(native, kw, code):
    lib = native.load()
    h = C.c_void_p()
    cfg = _cfg(native, **kw)
    assert lib.wwf_plan_create(C.byref(cfg), 0, C.byref(h)) == code
    assert not h.value and len(lib.wwf_last_error()) > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_gpu(native):
    lib = native.load()
    h = C.c_void_p()
    cfg = _cfg(native)
    assert lib.wwf_plan_create(C.byref(cfg), 0, C.byref(h)) == -3          # WWF_ERR_CUDA, loudly
    assert b"cuda" in lib.wwf_last_error().lower()
    import wakeword_trainer_home_b200 as w
    with pytest.raises(Exception):
        w.FeatureExtractor(device="cuda")
    with pytest.raises(w.WwfError):
        w.FeatureExtractor(device="cpu")                                  # the reference's device kwarg, CPU refused


def test_null_and_shape_arguments(native):
    lib = native.load()
    assert lib.wwf_plan_create(None, 0, None) == -1
    assert lib.wwf_plan_info(None, None) == -1
    assert lib.wwf_featurize(None, None, 1, 1000, 1000, None, None, 0, None, 0, None) == -1
    assert lib.wwf_augment(None, None, 1, 1000, 1000, None, None, 0, None, 0, None) == -1
    assert lib.wwf_spec_augment(None, 0, 1, 1, 1, 1, None, None, 0, None, None, 0, 0.0, 0, None) == -1
    assert lib.wwf_workspace_bytes(None, 4, 1000) == 0
    lib.wwf_plan_destroy(None)                                            # must be a no-op
    # time-stretch / pitch-shift / resample entry points: argument errors are reported before any CUDA call
    assert lib.wwf_time_stretch(None, None, 1, 1000, 1000, None, 0.8, None, 1000, None, 0, None) == -1
    assert lib.wwf_pitch_shift(None, None, 1, 1000, 1000, None, -2, 2, None, 1000, None, 0, None) == -1
    assert lib.wwf_pitch_shift(None, None, 1, 1000, 1000, None, -13, 2, None, 1000, None, 0, None) == -1
    assert lib.wwf_resample(None, None, 1, 1000, 1000, 44100, 16000, None, 363, 363, None) == -1
    assert lib.wwf_set_stretch_window(None, None) == -1
    assert lib.wwf_stretch_workspace_bytes(0, 1000, 0.8) == 0 and lib.wwf_stretch_workspace_bytes(4, 1000, 0.01) == 0
    assert lib.wwf_pitch_workspace_bytes(4, 24000, -13, 2) == 0
    a, b = lib.wwf_stretch_workspace_bytes(4, 24000, 0.8), lib.wwf_stretch_workspace_bytes(4, 24000, 0.5)
    assert 0 < a < b and a % 16 == 0                                      # slower rates need more room
    assert lib.wwf_pitch_workspace_bytes(4, 24000, -2, 2) == lib.wwf_stretch_workspace_bytes(4, 24000, 2.0 ** (-2 / 12))


def test_resample_length_matches_torchaudio_rule():
    """wwf_resample_length = ceil(new * n / orig) with the quotient rounded to float32 first, exactly as
    torchaudio computes target_length (TA/functional/functional.py:1427)."""
    import math
    from wakeword_trainer_home_b200 import _native
    lib = _native.load()
    rng = np.random.default_rng(1)
    for orig, new in ((44100, 16000), (8000, 16000), (17959, 16000), (14254, 16000), (48000, 16000), (16000, 16000), (22050, 44100)):
        for n in [1, 2, 999, 24000, 26667] + rng.integers(1, 2_000_000, 20).tolist():
            g = math.gcd(orig, new)
            want = n if orig == new else int(torch.ceil(torch.as_tensor((new // g) * n / (orig // g))).long())
            assert lib.wwf_resample_length(int(n), orig, new) == want, (orig, new, n)
    assert lib.wwf_resample_length(-1, 8000, 16000) == -1 and lib.wwf_resample_length(10, 0, 16000) == -1


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "wakeword_trainer_home_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "torchaudio" not in re.sub(r'""".*?"""|#.*|//.*', "", src, flags=re.S), f"{f} must not call torchaudio"


def test_constants_bit_identical_to_torchaudio():
    import torchaudio.functional as AF
    from wakeword_trainer_home_b200 import constants as K
    for n_fft, n_mels, n_mfcc in ((400, 40, 40), (1024, 128, 40), (512, 64, 32), (256, 40, 13), (2048, 128, 64)):
        assert torch.equal(K.mel_filterbank(n_fft // 2 + 1, 0.0, 8000.0, n_mels, 16000),
                           AF.melscale_fbanks(n_fft // 2 + 1, 0.0, 8000.0, n_mels, 16000))
        assert torch.equal(K.dct_matrix(n_mfcc, n_mels), AF.create_dct(n_mfcc, n_mels, "ortho"))
        assert torch.equal(K.hann_window(n_fft), torch.hann_window(n_fft))


def test_draw_mask_params_and_augparams():
    from wakeword_trainer_home_b200.pipeline import AugParams, draw_mask_params
    from oracle import ta_oracle as tao
    a = draw_mask_params(torch.Generator().manual_seed(7), 32, 151, 35, 2)
    b = tao.draw_mask_params(torch.Generator().manual_seed(7), 32, 151, 35, 2)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])            # host draw logic == oracle's, bit-exact
    s, l = draw_mask_params(torch.Generator().manual_seed(1), 16, 40, 15, 2, p=0.1)
    assert (l < 4).all()                                                  # p caps the width: min(15, int(40*0.1))
    z = draw_mask_params(None, 4, 40, 0, 2)
    assert (z[0] == 0).all() and (z[1] == 0).all()
    p = AugParams(rir_idx=[0, -1], noise_idx=torch.tensor([1, 2]), noise_off=[5, 6], snr_db=[1.0, 2.0]).to("cpu")
    assert p.rir_idx.dtype == torch.int32 and p.noise_off.dtype == torch.int64 and p.snr_db.dtype == torch.float32
    assert p.nbytes() == 2 * 4 + 2 * 4 + 2 * 8 + 2 * 4 and p.fmask_start is None


def test_shard_range_partitions():
    from wakeword_trainer_home_b200.sharding import shard_range, shard_seed
    for n in (0, 1, 7, 1024, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert len({shard_seed(5, r, s) for r in range(8) for s in range(100)}) == 800
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


_GLOO_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from wakeword_trainer_home_b200.sharding import rank_world, shard_range, shard_seed
from wakeword_trainer_home_b200.pipeline import draw_mask_params
rank, local_rank, world = rank_world()
dist.init_process_group("gloo")
n = 1027
a, b = shard_range(n, rank, world)
mine = torch.zeros(n, dtype=torch.int64); mine[a:b] = 1
dist.all_reduce(mine)                       # test-only collective: the feature path itself issues none
assert int(mine.min()) == 1 and int(mine.max()) == 1, "shards must tile the batch exactly once"
g = torch.Generator().manual_seed(shard_seed(11, rank, 3))
s, l = draw_mask_params(g, b - a, 151, 35, 2)
sums = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
dist.all_gather(sums, s.sum().reshape(1).long())
assert len({int(x) for x in sums}) == world, "ranks must draw different augmentations"
t = torch.tensor([float(rank + 1)]); dist.all_reduce(t, op=dist.ReduceOp.MAX)   # bench.py's max-over-ranks timing
assert float(t) == float(world)
dist.destroy_process_group()
print("ok", rank)
'''


def test_two_rank_sharding_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), ROOT],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_philox_mirror_known_answers():
    """The numpy mirror of the on-GPU sampler is the published Philox4x32-10: Random123's known-answer
    vectors (kat_vectors: philox4x32 10 rounds)."""
    from helpers import philox4x32_10
    z = np.zeros(1, np.uint32)
    out = philox4x32_10(z, z, z, z, 0, 0)
    assert [int(v[0]) for v in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = np.full(1, 0xffffffff, np.uint32)
    out = philox4x32_10(f, f, f, f, 0xffffffff, 0xffffffff)
    assert [int(v[0]) for v in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = philox4x32_10(np.array([0x243f6a88], np.uint32), np.array([0x85a308d3], np.uint32), np.array([0x13198a2e], np.uint32),
                        np.array([0x03707344], np.uint32), 0xa4093822, 0x299f31d0)
    assert [int(v[0]) for v in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def _write_wav(path, x, rate, bits=16, fmt_tag=1):
    """x: float (channels, frames) in [-1, 1).  Minimal RIFF writer for the decoder test (all depths, float)."""
    import struct
    ch, frames = x.shape
    inter = x.T.reshape(-1)
    if fmt_tag == 3:
        body = inter.astype("<f4").tobytes()
        bits = 32
    elif bits == 8:
        body = (np.round(inter * 128.0) + 128).clip(0, 255).astype(np.uint8).tobytes()
    elif bits == 16:
        body = np.round(inter * 32768.0).clip(-32768, 32767).astype("<i2").tobytes()
    elif bits == 24:
        v = np.round(inter * 8388608.0).clip(-8388608, 8388607).astype(np.int32)
        body = np.stack([v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF], axis=1).astype(np.uint8).tobytes()
    else:
        body = np.round(inter.astype(np.float64) * 2147483648.0).clip(-2 ** 31, 2 ** 31 - 1).astype("<i4").tobytes()
    fmt = struct.pack("<HHIIHH", fmt_tag, ch, rate, rate * ch * bits // 8, ch * bits // 8, bits)
    junk = b"LIST" + struct.pack("<I", 3) + b"abc" + b"\x00"                  # odd-sized chunk + pad byte
    data = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + junk + b"data" + struct.pack("<I", len(body)) + body
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(data)) + data)


def test_wav_decoder_all_sample_formats(tmp_path):
    """AudioProcessor's host-side container parsing (the numeric part runs on the GPU): every PCM depth and IEEE
    float, stereo de-interleaving, chunks in front of 'data' - and agreement with the stdlib wave module."""
    import wave
    from wakeword_trainer_home_b200.audio_utils import read_wav
    rng = np.random.default_rng(0)
    x = (rng.uniform(-0.9, 0.9, (2, 1000))).astype(np.float32)
    for bits, tag, tol in ((8, 1, 1 / 128), (16, 1, 1 / 32768), (24, 1, 1 / 8388608), (32, 1, 1e-7), (32, 3, 0.0)):
        p = str(tmp_path / f"t{bits}_{tag}.wav")
        _write_wav(p, x, 44100, bits, tag)
        y, rate = read_wav(p)
        assert rate == 44100 and y.shape == x.shape and y.dtype == np.float32
        assert np.abs(y - x).max() <= tol + 1e-7
    p = str(tmp_path / "std.wav")
    pcm = np.round(x.T.reshape(-1) * 32768.0).astype("<i2")
    with wave.open(p, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(8000); w.writeframes(pcm.tobytes())
    y, rate = read_wav(p)
    assert rate == 8000 and np.array_equal(y, pcm.reshape(-1, 2).T.astype(np.float32) / 32768.0)
    with pytest.raises(ValueError):
        (tmp_path / "bad.wav").write_bytes(b"RIFFxxxxWAVEjunk")
        read_wav(str(tmp_path / "bad.wav"))


def test_split_manifest_and_npy_shapes(tmp_path):
    from wakeword_trainer_home_b200 import formats
    p = str(tmp_path / "splits" / "train.json")
    formats.save_split_manifest(p, ["a.wav", "b.wav", "c.wav"], [1, 0, 1], extra={"split": "train"})
    paths, labels, meta = formats.load_split_manifest(p)
    assert paths == ["a.wav", "b.wav", "c.wav"] and labels.tolist() == [1, 0, 1] and meta[0]["path"] == "a.wav"
    for doc, want in (([{"path": "x.wav", "category": "positive"}, {"path": "y.wav", "category": "negative"}], [1, 0]),
                      ({"samples": [{"path": "x.wav", "label": 0}]}, [0]), (["x.wav", "y.wav"], [-1, -1])):
        q = tmp_path / "m.json"
        q.write_text(__import__("json").dumps(doc))
        assert formats.load_split_manifest(str(q))[1].tolist() == want
    with pytest.raises(ValueError):
        (tmp_path / "m2.json").write_text('{"nothing": 1}')
        formats.load_split_manifest(str(tmp_path / "m2.json"))
    assert formats.npy_kind(np.zeros((3, 100))) == "audio" and formats.npy_kind(np.zeros((3, 40, 11))) == "features"
    assert formats.npy_kind(np.zeros((3, 1, 40, 11))) == "features"
    with pytest.raises(ValueError):
        formats.npy_kind(np.zeros(5))


def test_aug_params_all_reverb_hint_is_taken_from_host_draws():
    """AugParams.all_reverb (a scheduling hint for the reverb kernel, never a correctness input): filled in by .to() while
    rir_idx is still a host tensor, kept when given, unknown (None) without rir_idx."""
    import torch
    from wakeword_trainer_home_b200 import AugParams
    assert AugParams(rir_idx=torch.tensor([0, 3, 1], dtype=torch.int32)).to("cpu").all_reverb is True
    assert AugParams(rir_idx=torch.tensor([0, -1, 1], dtype=torch.int32)).to("cpu").all_reverb is False
    assert AugParams(noise_idx=torch.tensor([0, 1], dtype=torch.int32)).to("cpu").all_reverb is None
    assert AugParams(rir_idx=torch.tensor([0, -1], dtype=torch.int32), all_reverb=True).to("cpu").all_reverb is True
    a = AugParams(rir_idx=torch.tensor([2, 2], dtype=torch.int64)).to("cpu")
    assert a.rir_idx.dtype == torch.int32 and a.all_reverb is True and a.to("cpu").all_reverb is True
