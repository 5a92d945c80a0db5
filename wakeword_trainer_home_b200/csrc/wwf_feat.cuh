// wwf_feat.cuh - the fused feature kernel:  [noise @ SNR mix] -> frame/window -> rFFT -> |.|^2
// -> sparse mel -> dB -> per-clip top_db floor -> [DCT-II] -> [CMVN] -> [SpecAugment] -> store.
// SURVEY.md section 8a rows A1 (mix), A4..A10.
//
// One CTA per clip (the top_db floor needs the clip's maximum before any element can be
// finalised, TA/functional/functional.py:393-402).  Inside the CTA every WARP is autonomous:
// it takes a group of 2*G consecutive frames, packs them two-per-complex-FFT (frame a ->
// real part, frame b -> imaginary part), runs the in-place mixed-radix FFT in its private
// shared-memory scratch with __syncwarp() only, separates the two spectra, applies the
// sparse mel rows and writes dB values into the CTA's [n_mels][T] shared tile.  The only
// CTA-wide barriers are around the per-clip energy reduction (noise mix) and the tile
// maximum.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>
#include "wwf_fft.cuh"

namespace wwf {

constexpr int kMaxMasks = 8;

template <int NFFT> struct StftPlan;
template <> struct StftPlan<256>  { using Rad = Radices<16, 16>;    static constexpr int G = 2; };
template <> struct StftPlan<400>  { using Rad = Radices<16, 25>;    static constexpr int G = 2; };
template <> struct StftPlan<512>  { using Rad = Radices<16, 8, 4>;  static constexpr int G = 1; };
template <> struct StftPlan<1024> { using Rad = Radices<16, 16, 4>; static constexpr int G = 1; };
template <> struct StftPlan<2048> { using Rad = Radices<16, 16, 8>; static constexpr int G = 1; };

struct FeatParams {
  // inputs
  const float* wav;        int64_t wav_stride;   // original clips [B][N]
  const float* rev;        int64_t rev_stride;   // reverberated clips (workspace) or nullptr
  int B, N, T, hop;
  // configuration
  int n_mels, n_mfcc, n_feat, is_mfcc, out_f16, cmvn;
  float top_db, cmvn_eps, mask_value;
  int tile_pitch;                                // odd row pitch of the shared tile (>= T)
  int tile_floats, res_floats;                   // even float counts of the two shared tiles
  // device constants (plan-owned)
  const float* window;                           // [NFFT]
  const float2* tw;                              // concatenated per-pass twiddle tables
  const int* mel_lo;                             // [n_mels] first FFT bin of each filter
  const int* mel_ofs;                            // [n_mels+1] CSR offsets into mel_w
  const float* mel_w;                            // filter weights, bin-contiguous per filter
  const float* dct;                              // [n_mels][n_mfcc]
  // augmentation draws (device, nullable)
  const int32_t* rir_idx; const int32_t* noise_idx; const int64_t* noise_off; const float* snr_db;
  const float* noise_data; const int64_t* noise_offsets; int n_noise;
  const int32_t* fs; const int32_t* fl; const int32_t* ts; const int32_t* tl; int nF, nT;
  // output
  void* out; int64_t out_stride;
};

// ---- small device utilities --------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// torch reflect padding index (edge sample not repeated): i in [-N+1, 2N-2] -> [0, N)
WWF_HD int reflect_index(int i, int N) {
  i = i < 0 ? -i : i;
  return i >= N ? 2 * (N - 1) - i : i;
}

// Two real frames a, b were transformed as one complex signal a + i b.  Given Z[k] and Z[n-k]
// return (|A[k]|^2, |B[k]|^2) with A[k] = (Z[k] + conj Z[n-k])/2, B[k] = -i (Z[k] - conj Z[n-k])/2.
WWF_HD float2 pair_split_power(float2 a, float2 c) {
  const float sx = a.x + c.x, sy = a.y - c.y;   // Z[k] + conj(Z[n-k])
  const float dx = a.x - c.x, dy = a.y + c.y;   // Z[k] - conj(Z[n-k])
  return make_float2(0.25f * fmaf(sx, sx, sy * sy), 0.25f * fmaf(dx, dx, dy * dy));
}

// noise sample for clip position i: bank[(off + i) mod len]
__device__ __forceinline__ float noise_at(const float* nz, int noff, int nlen, int i) {
  int q = noff + i;
  if (q >= nlen) {
    q -= nlen;
    if (q >= nlen) q %= nlen;
  }
  return __ldg(nz + q);
}

// scale of F.add_noise (TA/functional/functional.py:2376-2378), float32 like the oracle
__device__ __forceinline__ float snr_scale(float es, float en, float snr_db) {
  const float snr0 = 10.0f * (log10f(es) - log10f(en));
  return exp10f((snr0 - snr_db) / 20.0f);
}

// Block-wide sum of two values; result broadcast to all threads. red: >= 2*32 floats of smem.
__device__ __forceinline__ void block_sum2(float& a, float& b, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  __syncthreads();
  if (lane == 0) { red[warp] = a; red[32 + warp] = b; }
  __syncthreads();
  float x = lane < nw ? red[lane] : 0.f, y = lane < nw ? red[32 + lane] : 0.f;
  a = warp_sum(x);
  b = warp_sum(y);
}

// Energies of the clip and of its noise segment (two independent accumulators per thread).
__device__ __forceinline__ void clip_energies(const float* x, int N, const float* nz, int noff, int nlen,
                                              float& es, float& en, float* red) {
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
  int i = threadIdx.x;
  for (; i + (int)blockDim.x < N; i += 2 * blockDim.x) {
    const float x0 = __ldg(x + i), x1 = __ldg(x + i + blockDim.x);
    const float n0 = noise_at(nz, noff, nlen, i), n1 = noise_at(nz, noff, nlen, i + blockDim.x);
    a0 = fmaf(x0, x0, a0); a1 = fmaf(x1, x1, a1);
    b0 = fmaf(n0, n0, b0); b1 = fmaf(n1, n1, b1);
  }
  if (i < N) {
    const float x0 = __ldg(x + i), n0 = noise_at(nz, noff, nlen, i);
    a0 = fmaf(x0, x0, a0); b0 = fmaf(n0, n0, b0);
  }
  es = a0 + a1;
  en = b0 + b1;
  block_sum2(es, en, red);
}

template <typename OutT> __device__ __forceinline__ OutT to_out(float v);
template <> __device__ __forceinline__ float to_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __half to_out<__half>(float v) { return __float2half_rn(v); }

// ---- the kernel ------------------------------------------------------------------------
// Dynamic shared memory layout (floats):
//   tile  [n_mels][pitch]            dB mel values of the clip
//   res   [n_feat][pitch]            only if (mfcc && cmvn): DCT output awaiting normalisation
//   zbuf  [nwarps][G][NFFT] float2   per-warp FFT scratch
template <int NFFT, typename OutT>
__global__ void __launch_bounds__(512) feat_kernel(const FeatParams p) {
  using Plan = StftPlan<NFFT>;
  using Rad = typename Plan::Rad;
  constexpr int G = Plan::G;
  constexpr int K = NFFT / 2 + 1;
  static_assert(Rad::n == NFFT, "radix plan");

  extern __shared__ __align__(16) float smem[];
  __shared__ float red[64];
  __shared__ int s_mask[4 * kMaxMasks];

  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int N = p.N, T = p.T, hop = p.hop, M = p.n_mels, pitch = p.tile_pitch;

  float* tile = smem;
  float* res = tile + p.tile_floats;
  float2* zbuf = reinterpret_cast<float2*>(res + p.res_floats);
  float2* z = zbuf + (size_t)warp * G * NFFT;

  // ---- per-clip setup ---------------------------------------------------------------
  const bool has_rev = p.rev != nullptr && p.rir_idx != nullptr && __ldg(p.rir_idx + b) >= 0;
  const float* x = has_rev ? p.rev + (size_t)b * p.rev_stride : p.wav + (size_t)b * p.wav_stride;

  if (tid < 4 * kMaxMasks) {
    const int which = tid / kMaxMasks, i = tid % kMaxMasks;   // 0 fs, 1 fl, 2 ts, 3 tl
    int v = 0;
    if (which < 2) { if (p.fs && i < p.nF) v = __ldg((which == 0 ? p.fs : p.fl) + (size_t)b * p.nF + i); }
    else           { if (p.ts && i < p.nT) v = __ldg((which == 2 ? p.ts : p.tl) + (size_t)b * p.nT + i); }
    s_mask[tid] = v;
  }

  const float* nz = nullptr;
  int noff = 0, nlen = 1;
  float scale = 0.f;
  if (p.noise_idx != nullptr && p.noise_data != nullptr) {
    const int ni = __ldg(p.noise_idx + b);
    if (ni >= 0 && ni < p.n_noise) {
      const int64_t o0 = __ldg(p.noise_offsets + ni), o1 = __ldg(p.noise_offsets + ni + 1);
      nlen = (int)(o1 - o0);
      nz = p.noise_data + o0;
      int64_t off = p.noise_off ? __ldg(p.noise_off + b) : 0;
      off %= nlen; if (off < 0) off += nlen;
      noff = (int)off;
      float es, en;
      clip_energies(x, N, nz, noff, nlen, es, en, red);
      scale = snr_scale(es, en, p.snr_db ? __ldg(p.snr_db + b) : 0.f);
    }
  }
  const bool mix = nz != nullptr;

  // ---- frames: STFT -> power -> mel -> dB into the tile (warp-autonomous) ----------------
  const int ngroups = (T + 2 * G - 1) / (2 * G);
  for (int grp = warp; grp < ngroups; grp += nwarps) {
    const int f0 = grp * 2 * G;
    // 1. load + window: z[g][j] = w[j] * (frame(f0+2g)[j] + i frame(f0+2g+1)[j])
    for (int idx = lane; idx < G * NFFT; idx += 32) {
      const int g = idx / NFFT, j = idx - g * NFFT;
      const int ta = f0 + 2 * g, tb = ta + 1;
      const float w = __ldg(p.window + j);
      float re = 0.f, im = 0.f;
      if (ta < T) {
        const int i = reflect_index(ta * hop - NFFT / 2 + j, N);
        re = __ldg(x + i);
        if (mix) re = fmaf(scale, noise_at(nz, noff, nlen, i), re);
      }
      if (tb < T) {
        const int i = reflect_index(tb * hop - NFFT / 2 + j, N);
        im = __ldg(x + i);
        if (mix) im = fmaf(scale, noise_at(nz, noff, nlen, i), im);
      }
      z[idx] = make_float2(re * w, im * w);
    }
    __syncwarp();
    // 2. forward FFT passes (in place, digit-reversed result)
    static_for<0, Rad::npass>([&](auto I) {
      constexpr int i = decltype(I)::value;
      constexpr int R = Rad::R(i), L = Rad::L(i), tasks = NFFT / R;
      const float2* tw = p.tw + Rad::tw_off(i);
      for (int u = lane; u < G * tasks; u += 32) {
        const int g = u / tasks, uu = u - g * tasks;
        pass_task<R, false>(z + g * NFFT, L, uu, [&](int q) { return __ldg(tw + q); });
      }
      __syncwarp();
    });
    // 3. split the packed pair: A[k] = (Z[k] + conj Z[n-k])/2, B[k] = -i (Z[k] - conj Z[n-k])/2;
    //    store the two power spectra at Z[k]'s slot (only bin k's lane touches it).
    for (int idx = lane; idx < G * K; idx += 32) {
      const int g = idx / K, k = idx - g * K;
      float2* zz = z + g * NFFT;
      const int pk = Rad::pos(k), pm = Rad::pos(k == 0 ? 0 : NFFT - k);
      zz[pk] = pair_split_power(zz[pk], zz[pm]);
    }
    __syncwarp();
    // 4. sparse mel rows + dB
    for (int idx = lane; idx < G * M; idx += 32) {
      const int g = idx / M, m = idx - g * M;
      const float2* zz = z + g * NFFT;
      const int lo = __ldg(p.mel_lo + m), o0 = __ldg(p.mel_ofs + m), o1 = __ldg(p.mel_ofs + m + 1);
      float acc_a = 0.f, acc_b = 0.f;
      for (int o = o0; o < o1; ++o) {
        const float w = __ldg(p.mel_w + o);
        const float2 pw = zz[Rad::pos(lo + (o - o0))];
        acc_a = fmaf(w, pw.x, acc_a);
        acc_b = fmaf(w, pw.y, acc_b);
      }
      const int ta = f0 + 2 * g;
      if (ta < T) tile[m * pitch + ta] = 10.0f * log10f(fmaxf(acc_a, 1e-10f));
      if (ta + 1 < T) tile[m * pitch + ta + 1] = 10.0f * log10f(fmaxf(acc_b, 1e-10f));
    }
    __syncwarp();
  }
  __syncthreads();

  // ---- per-clip top_db floor -------------------------------------------------------------
  float cutoff = -INFINITY;
  if (p.top_db >= 0.f) {
    float mx = -INFINITY;
    for (int m = warp; m < M; m += nwarps)
      for (int t = lane; t < T; t += 32) mx = fmaxf(mx, tile[m * pitch + t]);
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = lane < nwarps ? red[lane] : -INFINITY;
    mx = warp_max(mx);
    cutoff = mx - p.top_db;
  }

  OutT* out = reinterpret_cast<OutT*>(p.out) + (size_t)b * p.out_stride;
  const int F = p.n_feat;
  auto masked = [&](int f, int t) -> bool {
    bool mk = false;
#pragma unroll
    for (int i = 0; i < kMaxMasks; ++i) {
      mk |= (i < p.nF) && (f >= s_mask[i]) && (f < s_mask[i] + s_mask[kMaxMasks + i]);
      mk |= (i < p.nT) && (t >= s_mask[2 * kMaxMasks + i]) && (t < s_mask[2 * kMaxMasks + i] + s_mask[3 * kMaxMasks + i]);
    }
    return mk;
  };

  if (!p.is_mfcc) {
    if (!p.cmvn) {
      for (int m = warp; m < M; m += nwarps)
        for (int t = lane; t < T; t += 32) {
          const float v = fmaxf(tile[m * pitch + t], cutoff);
          out[(size_t)m * T + t] = to_out<OutT>(masked(m, t) ? p.mask_value : v);
        }
      return;
    }
    for (int m = warp; m < M; m += nwarps)
      for (int t = lane; t < T; t += 32) tile[m * pitch + t] = fmaxf(tile[m * pitch + t], cutoff);
    res = tile;
  } else {
    // DCT-II: out[c][t] = sum_m dct[m][c] * max(tile[m][t], cutoff); 8 coefficients per thread
    const int C = F, ncg = (C + 7) / 8;
    for (int idx = tid; idx < ncg * T; idx += blockDim.x) {
      const int cg = idx / T, t = idx - cg * T, c0 = cg * 8;
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      for (int m = 0; m < M; ++m) {
        const float a = fmaxf(tile[m * pitch + t], cutoff);
        const float* d = p.dct + (size_t)m * C + c0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (c0 + i < C) acc[i] = fmaf(a, __ldg(d + i), acc[i]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = c0 + i;
        if (c < C) {
          if (p.cmvn) res[c * pitch + t] = acc[i];
          else out[(size_t)c * T + t] = to_out<OutT>(masked(c, t) ? p.mask_value : acc[i]);
        }
      }
    }
    if (!p.cmvn) return;
  }
  // ---- CMVN epilogue: per row (x - mean) / (population std + eps) ----------------------
  __syncthreads();
  for (int f = warp; f < F; f += nwarps) {
    float s = 0.f;
    for (int t = lane; t < T; t += 32) s += res[f * pitch + t];
    const float mean = warp_sum(s) / (float)T;
    float q = 0.f;
    for (int t = lane; t < T; t += 32) { const float d = res[f * pitch + t] - mean; q = fmaf(d, d, q); }
    const float sd = sqrtf(warp_sum(q) / (float)T);
    const float inv = 1.0f / (sd + p.cmvn_eps);
    for (int t = lane; t < T; t += 32) {
      const float v = (res[f * pitch + t] - mean) * inv;
      out[(size_t)f * T + t] = to_out<OutT>(masked(f, t) ? p.mask_value : v);
    }
  }
}

}  // namespace wwf
