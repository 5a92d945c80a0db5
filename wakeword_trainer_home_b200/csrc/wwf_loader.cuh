// wwf_loader.cuh - the device-resident side of the batched loader:
//   * gather_clips_kernel: batch assembly from a clip bank that lives in HBM (float32 or int16 PCM),
//     replacing the per-sample file reads of WakewordDataset.__getitem__ for in-memory datasets;
//   * draw_aug_kernel: the augmentation draws of AudioAugmentation / SpecAugment made ON the GPU with
//     the counter-based Philox4x32-10 generator, so a training step needs no host RNG and no H2D copy.
// The draws are a pure function of (seed, sample index): tests/helpers.py holds a numpy mirror and
// the GPU tests require bit-exact equality, which keeps "all draws explicit" true - the host can
// always recompute what the device drew and hand it to the oracle.
#pragma once
#include <stdint.h>
#include "wwf_fft.cuh"

namespace wwf {

// ---- Philox4x32-10 (Salmon et al., SC'11), the generator behind cuRAND / torch's CUDA RNG -----------
struct Philox4 { uint32_t x, y, z, w; };

WWF_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

// uniform float32 in [0, 1) from the top 24 bits: exact in float32, no rounding anywhere
WWF_HD float u01(uint32_t u) { return (float)(u >> 8) * 5.9604644775390625e-08f; }
// uniform integer in [0, n) without modulo bias beyond 2^-32: high word of u * n
WWF_HD uint32_t pick(uint32_t u, uint32_t n) { return (uint32_t)(((uint64_t)u * n) >> 32); }

struct DrawParams {
  uint64_t seed, first_index;           // sample i of the call is sample number first_index + i of the run
  uint64_t thr_rir, thr_noise, thr_fmask, thr_tmask;   // floor(prob * 2^32): apply when u < thr
  float snr_lo, snr_hi;
  int B, n_rir, n_noise, F, T, fparam, tparam, nF, nT;
  const int64_t* noise_offsets;         // [n_noise + 1] (device) for the offset draw
  int32_t* rir_idx; int32_t* noise_idx; int64_t* noise_off; float* snr_db;
  int32_t* fs; int32_t* fl; int32_t* ts; int32_t* tl;
  // time-stretch / pitch-shift draws (Philox block with counter word 3 = 1)
  uint64_t thr_stretch, thr_pitch;
  double stretch_lo, stretch_hi;
  int pitch_lo, pitch_hi;
  double* stretch_rate; int32_t* pitch_steps;
};

// One (start, len) pair with torchaudio's mask_along_axis arithmetic in float32
// (value = U * param; min_value = U' * (size - value); start = floor(min_value), len = floor(value);
// TA/functional/functional.py:864-870), rounding made explicit so the host mirror is bit-identical.
__device__ __forceinline__ void mask_pair(uint32_t ua, uint32_t ub, int param, int size, bool on, int32_t* start, int32_t* len) {
  const float value = __fmul_rn(u01(ua), (float)param);
  const float minv = __fmul_rn(u01(ub), __fsub_rn((float)size, value));
  *start = on && param >= 1 ? (int32_t)minv : 0;
  *len = on && param >= 1 ? (int32_t)value : 0;
}

// Thread i draws everything for sample first_index + i.  Philox counter = (index lo, index hi, j, 0),
// key = seed; block j = 0: (rir on, rir pick, noise on, noise pick); j = 1: (noise offset, snr,
// freq-mask gate, time-mask gate); j = 2 + q: two freq masks (value, min) x2; then the time masks.
// Counter (index lo, index hi, 0, 1): (stretch gate, stretch u, pitch gate, pitch pick) - rate = lo + (hi - lo) * u / 2^32
// in double (every operation exactly rounded, so the host mirror is bit-identical), semitones = lo + pick(hi - lo + 1).
__global__ void __launch_bounds__(256) draw_aug_kernel(const DrawParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.B) return;
  const uint64_t idx = p.first_index + (uint64_t)i;
  const uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
  const Philox4 a = philox4x32_10(c0, c1, 0u, 0u, k0, k1);
  const Philox4 b = philox4x32_10(c0, c1, 1u, 0u, k0, k1);
  int32_t r = -1, n = -1;
  if (p.n_rir > 0 && (uint64_t)a.x < p.thr_rir) r = (int32_t)pick(a.y, (uint32_t)p.n_rir);
  if (p.n_noise > 0 && (uint64_t)a.z < p.thr_noise) n = (int32_t)pick(a.w, (uint32_t)p.n_noise);
  if (p.rir_idx) p.rir_idx[i] = r;
  if (p.noise_idx) p.noise_idx[i] = n;
  if (p.noise_off) {
    int64_t off = 0;
    if (n >= 0) off = (int64_t)pick(b.x, (uint32_t)(p.noise_offsets[n + 1] - p.noise_offsets[n]));
    p.noise_off[i] = off;
  }
  if (p.snr_db) p.snr_db[i] = __fadd_rn(p.snr_lo, __fmul_rn(__fsub_rn(p.snr_hi, p.snr_lo), u01(b.y)));
  if (p.stretch_rate || p.pitch_steps) {
    const Philox4 s = philox4x32_10(c0, c1, 0u, 1u, k0, k1);
    if (p.stretch_rate) {
      double r = 1.0;
      if ((uint64_t)s.x < p.thr_stretch)
        r = __dadd_rn(p.stretch_lo, __dmul_rn(__dsub_rn(p.stretch_hi, p.stretch_lo), __dmul_rn((double)s.y, 2.3283064365386963e-10)));
      p.stretch_rate[i] = r;
    }
    if (p.pitch_steps) {
      int32_t n = 0;
      if ((uint64_t)s.z < p.thr_pitch && p.pitch_hi >= p.pitch_lo) n = p.pitch_lo + (int32_t)pick(s.w, (uint32_t)(p.pitch_hi - p.pitch_lo + 1));
      p.pitch_steps[i] = n;
    }
  }
  const bool fon = (uint64_t)b.z < p.thr_fmask, ton = (uint64_t)b.w < p.thr_tmask;
  uint32_t j = 2;
  for (int q = 0; q < p.nF; q += 2, ++j) {
    const Philox4 m = philox4x32_10(c0, c1, j, 0u, k0, k1);
    mask_pair(m.x, m.y, p.fparam, p.F, fon, p.fs + (size_t)i * p.nF + q, p.fl + (size_t)i * p.nF + q);
    if (q + 1 < p.nF) mask_pair(m.z, m.w, p.fparam, p.F, fon, p.fs + (size_t)i * p.nF + q + 1, p.fl + (size_t)i * p.nF + q + 1);
  }
  for (int q = 0; q < p.nT; q += 2, ++j) {
    const Philox4 m = philox4x32_10(c0, c1, j, 0u, k0, k1);
    mask_pair(m.x, m.y, p.tparam, p.T, ton, p.ts + (size_t)i * p.nT + q, p.tl + (size_t)i * p.nT + q);
    if (q + 1 < p.nT) mask_pair(m.z, m.w, p.tparam, p.T, ton, p.ts + (size_t)i * p.nT + q + 1, p.tl + (size_t)i * p.nT + q + 1);
  }
}

// out[b][:] = bank[idx[b]][:] (float32) or bank[idx[b]][:] / 32768 (int16 PCM -> float, exact).
// grid = (chunks, B); 16-byte loads and stores when the rows allow it (vec != 0: N and both strides are
// multiples of 8 elements and both base pointers are 16-byte aligned - checked by the host).
template <typename SrcT>
__global__ void __launch_bounds__(256) gather_clips_kernel(const SrcT* __restrict__ bank, int64_t n_clips, int N, int64_t bank_stride,
                                                           const int64_t* __restrict__ idx, float* __restrict__ out, int64_t out_stride, int vec) {
  const int b = blockIdx.y;
  int64_t src = idx[b];
  if (src < 0 || src >= n_clips) src = 0;                    // indices are validated by the host for known ranges
  const SrcT* s = bank + src * bank_stride;
  float* d = out + (size_t)b * out_stride;
  const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  if (vec) {
    if constexpr (sizeof(SrcT) == 2) {
      const uint4* s8 = reinterpret_cast<const uint4*>(s);     // 8 samples per load
      float4* d4 = reinterpret_cast<float4*>(d);
      for (int i = t; i < N / 8; i += nt) {
        const uint4 v = __ldg(s8 + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        float f[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f[2 * j] = (float)(int16_t)(w[j] & 0xffffu) * (1.0f / 32768.0f);
          f[2 * j + 1] = (float)(int16_t)(w[j] >> 16) * (1.0f / 32768.0f);
        }
        d4[2 * i] = make_float4(f[0], f[1], f[2], f[3]);
        d4[2 * i + 1] = make_float4(f[4], f[5], f[6], f[7]);
      }
    } else {
      const float4* s4 = reinterpret_cast<const float4*>(s);
      float4* d4 = reinterpret_cast<float4*>(d);
      for (int i = t; i < N / 4; i += nt) d4[i] = __ldg(s4 + i);
    }
    return;
  }
  for (int i = t; i < N; i += nt) {
    if constexpr (sizeof(SrcT) == 2) d[i] = (float)s[i] * (1.0f / 32768.0f);
    else d[i] = (float)s[i];
  }
}

}  // namespace wwf
