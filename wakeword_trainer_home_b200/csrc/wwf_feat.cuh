// wwf_feat.cuh - the fused feature kernel:  [noise @ SNR mix] -> frame/window -> rFFT -> |.|^2
// -> sparse mel -> dB -> per-clip top_db floor -> [DCT-II] -> [CMVN] -> [SpecAugment] -> store.
// SURVEY.md section 8a rows A1 (mix), A4..A10.
//
// Persistent CTAs, one clip at a time per CTA (the top_db floor needs the clip's maximum
// before any element can be finalised, TA/functional/functional.py:393-402).  All constant
// tables (window, twiddles, sparse mel rows, DCT) are copied to shared memory once per CTA.
// Inside the CTA every WARP is autonomous: it takes a group of 2*G consecutive frames, packs
// them two-per-complex-FFT (frame a -> real part, frame b -> imaginary part), runs the
// in-place mixed-radix FFT in its private shared-memory scratch with __syncwarp() only,
// separates the two spectra, applies the sparse mel rows and writes dB values into the
// CTA's [n_mels][T] shared tile.  The only CTA-wide barriers are around the per-clip
// energy reduction (noise mix), the tile maximum and the clip hand-over.
//
// Frame loading (HOP32 = hop/32 > 0, i.e. hop a multiple of 32 such as the reference's 160):
// the 2G frames of a group overlap, so the warp loads their common sample span ONCE, fully
// coalesced, into registers in lane-cyclic order (sample i of the span lives in lane i%32,
// register i/32).  Because hop is a multiple of 32, sample j of frame f is in the SAME lane
// as sample j of frame f+1 (HOP32 registers further), so both halves of every complex FFT
// input come from the lane's own registers with compile-time indices - no shuffles, no
// per-element index math, and the noise mix happens once per loaded sample.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>
#include "wwf_fft.cuh"

namespace wwf {

constexpr int kMaxMasks = 8;

// Per-n_fft plan: radix list, G = complex FFTs (frame pairs) per warp iteration, the launch bounds
// (max threads per CTA, min CTAs per SM) the kernel is compiled for, and the index map of the warp's
// FFT scratch: the power-of-two sizes need one pad element per 16 (their later passes would
// otherwise be 4- to 16-way bank conflicted); 400 = 16 x 25 has an odd stride and needs none.
template <int NFFT> struct StftPlan;
template <> struct StftPlan<256>  { using Rad = Radices<16, 16>;    using Map = PadMap2;     static constexpr int G = 2, kThreads = 320, kMinCtas = 2; };
template <> struct StftPlan<400>  { using Rad = Radices<16, 25>;    using Map = IdentityMap; static constexpr int G = 2, kThreads = 320, kMinCtas = 2; };
template <> struct StftPlan<512>  { using Rad = Radices<16, 8, 4>;  using Map = PadMap2;     static constexpr int G = 1, kThreads = 512, kMinCtas = 1; };
template <> struct StftPlan<1024> { using Rad = Radices<16, 16, 4>; using Map = PadMap2;     static constexpr int G = 1, kThreads = 512, kMinCtas = 1; };
template <> struct StftPlan<2048> { using Rad = Radices<16, 16, 8>; using Map = PadMap2;     static constexpr int G = 1, kThreads = 512, kMinCtas = 1; };
// complex elements of one FFT's scratch buffer (mapped length, rounded up to an even count)
template <int NFFT> constexpr int stft_zlen() {
  return (typename StftPlan<NFFT>::Map()(NFFT - 1) + 2) & ~1;
}

constexpr int kNoiseBlk = 128;   // granularity of the noise bank's squared-sample prefix sums

// Registered background-noise bank (device view).  sq_prefix holds, per clip, the running sum
// (double) of squared samples at every kNoiseBlk boundary: P[j] = sum_{q < j*128} n[q]^2, with a
// final entry for the whole clip, so the energy of ANY segment costs two table reads plus at
// most 2*127 edge samples instead of a pass over the segment.
struct NoiseBankDev {
  const float* data;              // all clips back to back (borrowed from the caller)
  const int64_t* offsets;         // [count+1] sample offsets
  const double* sq_prefix;        // concatenated per-clip prefix tables
  const int64_t* prefix_offsets;  // [count] start of clip i's table
  int count;
};

struct ClipNoise {
  const float* nz;     // nullptr = this clip has no noise
  const double* P;
  int len, off;
};

struct FeatParams {
  // inputs
  const float* wav;        int64_t wav_stride;   // original clips [B][N]
  const float* rev;        int64_t rev_stride;   // reverberated clips (workspace) or nullptr
  int B, N, T, hop;
  // configuration
  int n_mels, n_mfcc, n_feat, is_mfcc, cmvn;
  float top_db, cmvn_eps, mask_value;
  int tile_pitch;                                // odd row pitch of the shared tile (>= T)
  // dynamic shared-memory layout, offsets in floats from the start (each a multiple of 4):
  //   tile [n_mels][pitch] | res [n_feat][pitch] (mfcc && cmvn only) | window [NFFT] | tw float2[tw_total]
  //   | mel_w | dct [n_mels][c8] | mel_lo int[n_mels] | mel_ofs int[n_mels+1] | rowmask u8[n_feat]
  //   | colmask u8[T] | z float2 [nwarps][G][NFFT]
  int off_res, off_window, off_tw, off_melw, off_dct, off_mello, off_melofs, off_rowmask, off_colmask, off_z;
  int n_melw, c8;                                // mel weight count; n_mfcc rounded up to 8
  // device constants (plan-owned)
  const float* window;                           // [NFFT]
  const float2* tw;                              // concatenated per-pass twiddle tables
  const int* mel_lo;                             // [n_mels] first FFT bin of each filter
  const int* mel_ofs;                            // [n_mels+1] CSR offsets into mel_w
  const float* mel_w;                            // filter weights, bin-contiguous per filter
  const float* dct;                              // [n_mels][n_mfcc]
  // augmentation draws (device, nullable)
  const int32_t* rir_idx; const int32_t* noise_idx; const int64_t* noise_off; const float* snr_db;
  NoiseBankDev noise;
  const float* es_part; int es_nb;               // per-clip energy partials written by conv_kernel [B][es_nb]
  const int32_t* fs; const int32_t* fl; const int32_t* ts; const int32_t* tl; int nF, nT;
  // output
  void* out; int64_t out_stride;
  int* nonfinite_flag;                           // plan-owned device int, set to 1 if any feature is NaN/Inf
  // ---- split path (large batches, see feat_frames_kernel): per-clip intermediates in the caller's workspace ----
  float* tile_g;                                 // dB values [B][T][mp], frame-major
  int mp;                                        // n_mels rounded up to 4
  int* clip_max;                                 // [B] running maximum of the clip's dB values (ordered-int key)
  float* scale_g;                                // [B] noise-mix scale (0 = no noise)
  int ngroups;                                   // frame groups per clip
  // feat_frames_kernel shared memory: window [NFFT] | tw | mel_w | mel_lo | mel_ofs | z float2 [nwarps][G][ZL]
  int f_off_tw, f_off_melw, f_off_mello, f_off_melofs, f_off_z;
  int eb_frames, eb_pitch;                       // feat_epilogue_block_kernel: frames per block, odd smem row pitch
};

// float <-> int key whose signed order equals the float order (for atomicMax on the clip maximum)
__device__ __forceinline__ int float_key(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }
// Identity of the running maximum when it is initialised by a byte-wise memset (no noise: feat_prep_kernel is
// skipped): 0x80808080 orders below the key of every float except -inf / NaN payloads and reads back as -inf.
constexpr int kMaxKeyMemset = (int)0x80808080;
__device__ __forceinline__ float clip_max_value(int k) { return k == kMaxKeyMemset ? -INFINITY : key_float(k); }

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
  const float2 cc = cconj(c);
  const float2 s = cadd(a, cc);                 // Z[k] + conj(Z[n-k])
  const float2 d = csub(a, cc);                 // Z[k] - conj(Z[n-k])
  return make_float2(0.25f * fmaf(s.x, s.x, s.y * s.y), 0.25f * fmaf(d.x, d.x, d.y * d.y));
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

// 10*log10(max(x, 1e-10)) = (10/log2(10)) * log2(.) through MUFU.LG2 (lg2.approx: max abs error
// 2^-22.6 on log2 => < 5e-7 dB; the argument is >= 1e-10, never denormal).
__device__ __forceinline__ float power_to_db(float x) {
  // the clamp value is exact like the oracle's; NaN takes the log branch and stays NaN (torch.clamp keeps NaN)
  return x <= 1e-10f ? -100.0f : 3.01029995663981195f * __log2f(x);
}

// scale of F.add_noise (TA/functional/functional.py:2376-2378), float32 like the oracle
__device__ __forceinline__ float snr_scale(float es, float en, float snr_db) {
  const float snr0 = 10.0f * (log10f(es) - log10f(en));
  return exp10f((snr0 - snr_db) / 20.0f);
}

// Noise clip, wrapped start offset and prefix table of batch item b (nz == nullptr: no noise).
__device__ __forceinline__ ClipNoise resolve_noise(const NoiseBankDev& bank, const int32_t* noise_idx,
                                                   const int64_t* noise_off, int b) {
  ClipNoise c{nullptr, nullptr, 1, 0};
  if (noise_idx == nullptr || bank.data == nullptr) return c;
  const int ni = __ldg(noise_idx + b);
  if (ni < 0 || ni >= bank.count) return c;
  const int64_t o0 = __ldg(bank.offsets + ni), o1 = __ldg(bank.offsets + ni + 1);
  c.len = (int)(o1 - o0);
  c.nz = bank.data + o0;
  c.P = bank.sq_prefix + __ldg(bank.prefix_offsets + ni);
  int64_t off = noise_off ? __ldg(noise_off + b) : 0;
  off %= c.len;
  if (off < 0) off += c.len;
  c.off = (int)off;
  return c;
}

// sum of nz[q]^2 over [a, b), 0 <= a <= b <= len: prefix table for whole 128-blocks, direct sum
// of the (< 128-sample) edges.  Executed by a full warp; every lane returns the result.
static __device__ __noinline__ double warp_seg_energy(const ClipNoise& c, int a, int b) {
  const int lane = threadIdx.x & 31;
  const int lo = (a + kNoiseBlk - 1) / kNoiseBlk, hi = b / kNoiseBlk;
  float e = 0.f;
  double mid = 0.0;
  if (lo <= hi) {
    mid = c.P[hi] - c.P[lo];
    for (int q = a + lane; q < lo * kNoiseBlk; q += 32) { const float v = __ldg(c.nz + q); e = fmaf(v, v, e); }
    for (int q = hi * kNoiseBlk + lane; q < b; q += 32) { const float v = __ldg(c.nz + q); e = fmaf(v, v, e); }
  } else {
    for (int q = a + lane; q < b; q += 32) { const float v = __ldg(c.nz + q); e = fmaf(v, v, e); }
  }
  return mid + (double)warp_sum(e);
}

// Energy of the N-sample noise segment nz[(off + i) mod len], i < N.
__device__ __forceinline__ float warp_noise_energy(const ClipNoise& c, int N) {
  const int first = min(N, c.len - c.off);
  double e = warp_seg_energy(c, c.off, c.off + first);
  int rem = N - first;
  if (rem > 0) {
    const int loops = rem / c.len;
    rem -= loops * c.len;
    if (loops > 0) e += (double)loops * c.P[(c.len + kNoiseBlk - 1) / kNoiseBlk];
    if (rem > 0) e += warp_seg_energy(c, 0, rem);
  }
  return (float)e;
}

// Block-wide sum of x[i]^2, i < N (8 independent loads in flight per thread); result in all
// threads.  red: >= 32 floats of shared memory.  Contains __syncthreads().
static __device__ __noinline__ float block_energy(const float* __restrict__ x, int N, float* red) {
  const int tid = threadIdx.x, nt = blockDim.x;
  float acc[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) acc[u] = 0.f;
  int i = tid;
  for (; i + 7 * nt < N; i += 8 * nt) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(x + i + u * nt);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = fmaf(v[u], v[u], acc[u]);
  }
  for (; i < N; i += nt) { const float v = __ldg(x + i); acc[0] = fmaf(v, v, acc[0]); }
  float s = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
  const int lane = tid & 31, warp = tid >> 5, nw = (nt + 31) >> 5;
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) red[warp] = s;
  __syncthreads();
  s = lane < nw ? red[lane] : 0.f;
  return warp_sum(s);
}

// Mix scale of batch item b (0 if it has no noise): energies of the (possibly reverberated)
// clip and of its noise segment -> F.add_noise's scale.  CTA-uniform control flow.
__device__ __forceinline__ float clip_mix_scale(const ClipNoise& cn, const float* x, int N, bool has_rev,
                                                const float* es_part, int es_nb, int b, const float* snr_db, float* red) {
  if (cn.nz == nullptr) return 0.f;
  float es = 0.f;
  if (has_rev && es_part != nullptr) {
    for (int i = 0; i < es_nb; ++i) es += __ldg(es_part + (size_t)b * es_nb + i);   // fixed order: deterministic
  } else {
    es = block_energy(x, N, red);
  }
  const float en = warp_noise_energy(cn, N);
  return snr_scale(es, en, snr_db ? __ldg(snr_db + b) : 0.f);
}

template <typename OutT> __device__ __forceinline__ OutT to_out(float v);
template <> __device__ __forceinline__ float to_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __half to_out<__half>(float v) { return __float2half_rn(v); }

// ---- the kernel ------------------------------------------------------------------------
template <int NFFT, int HOP32, typename OutT>
__global__ void __launch_bounds__(StftPlan<NFFT>::kThreads, StftPlan<NFFT>::kMinCtas) feat_kernel(const FeatParams p) {
  using Plan = StftPlan<NFFT>;
  using Rad = typename Plan::Rad;
  constexpr int G = Plan::G;
  constexpr int K = NFFT / 2 + 1;
  constexpr int NC = (NFFT + 31) / 32;                       // 32-sample columns per frame
  constexpr bool kNatural = true;                            // power spectra re-stored in bin order (false: left
                                                             // at their digit-reversed slots, mel reads through pos())
  using Map = typename Plan::Map;
  constexpr int ZL = stft_zlen<NFFT>();                      // scratch elements per FFT (with padding)
  const Map zmap;
  constexpr int NI = (G * K + 31) / 32;                      // split items per lane
  static_assert(Rad::n == NFFT, "radix plan");

  extern __shared__ __align__(16) float smem[];
  __shared__ float red[64];
  __shared__ int s_next_group;                                // dynamic frame-group queue of the current clip

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int N = p.N, T = p.T, hop = p.hop, M = p.n_mels, pitch = p.tile_pitch, F = p.n_feat;

  float* tile = smem;
  float* res = smem + p.off_res;
  float* s_window = smem + p.off_window;
  float2* s_tw = reinterpret_cast<float2*>(smem + p.off_tw);
  float* s_melw = smem + p.off_melw;
  float* s_dct = smem + p.off_dct;
  int* s_mello = reinterpret_cast<int*>(smem + p.off_mello);
  int* s_melofs = reinterpret_cast<int*>(smem + p.off_melofs);
  unsigned char* s_rowmask = reinterpret_cast<unsigned char*>(smem + p.off_rowmask);
  unsigned char* s_colmask = reinterpret_cast<unsigned char*>(smem + p.off_colmask);
  float2* z = reinterpret_cast<float2*>(smem + p.off_z) + (size_t)warp * G * ZL;

  // ---- constants -> shared memory, once per (persistent) CTA ---------------------------
  for (int i = tid; i < NFFT; i += blockDim.x) s_window[i] = __ldg(p.window + i);
  for (int i = tid; i < Rad::tw_total; i += blockDim.x) s_tw[i] = __ldg(p.tw + i);
  for (int i = tid; i < p.n_melw; i += blockDim.x) s_melw[i] = __ldg(p.mel_w + i);
  for (int i = tid; i < M; i += blockDim.x) s_mello[i] = __ldg(p.mel_lo + i);
  for (int i = tid; i <= M; i += blockDim.x) s_melofs[i] = __ldg(p.mel_ofs + i);
  if (p.is_mfcc)
    for (int i = tid; i < M * p.c8; i += blockDim.x) {
      const int m = i / p.c8, c = i - m * p.c8;
      s_dct[i] = c < p.n_mfcc ? __ldg(p.dct + (size_t)m * p.n_mfcc + c) : 0.f;
    }
  pdl_wait();   // programmatic dependent launch: the constants above were staged while conv_kernel drained

  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    // ---- per-clip setup -------------------------------------------------------------
    const bool has_rev = p.rev != nullptr && p.rir_idx != nullptr && __ldg(p.rir_idx + b) >= 0;
    const float* x = has_rev ? p.rev + (size_t)b * p.rev_stride : p.wav + (size_t)b * p.wav_stride;

    // SpecAugment flags: one byte per feature row / frame
    for (int i = tid; i < F + T; i += blockDim.x) {
      const bool is_row = i < F;
      const int q = is_row ? i : i - F;
      const int32_t* st = is_row ? p.fs : p.ts;
      const int32_t* ln = is_row ? p.fl : p.tl;
      const int nm = is_row ? p.nF : p.nT;
      bool mk = false;
      if (st != nullptr)
        for (int j = 0; j < nm; ++j) {
          const int s0 = __ldg(st + (size_t)b * nm + j), l = __ldg(ln + (size_t)b * nm + j);
          mk |= (q >= s0) && (q < s0 + l);
        }
      (is_row ? s_rowmask : s_colmask)[q] = mk ? 1 : 0;
    }

    const ClipNoise cn = resolve_noise(p.noise, p.noise_idx, p.noise_off, b);
    const float* nz = cn.nz;
    const int noff = cn.off, nlen = cn.len;
    const float scale = clip_mix_scale(cn, x, N, has_rev, p.es_part, p.es_nb, b, p.snr_db, red);
    const bool mix = nz != nullptr;
    if (tid == 0) s_next_group = 0;
    __syncthreads();   // constants + mask flags + group queue visible

    // ---- frames: STFT -> power -> mel -> dB into the tile (warp-autonomous) --------------
    const int ngroups = (T + 2 * G - 1) / (2 * G);
    float vmax = -INFINITY;                                  // running maximum of the dB values this lane wrote
    // Groups are handed out dynamically: the warp schedulers do not serve the warps of a CTA
    // evenly, and a static split makes everyone wait at the barrier for the slowest warp.
    for (;;) {
      int grp = 0;
      if (lane == 0) grp = atomicAdd(&s_next_group, 1);
      grp = __shfl_sync(0xffffffffu, grp, 0);
      if (grp >= ngroups) break;
      const int f0 = grp * 2 * G;
      // 1. load + window: z[g][j] = w[j] * (frame(f0+2g)[j] + i frame(f0+2g+1)[j])
      bool staged = false;
      if constexpr (HOP32 > 0) {
        constexpr int NR = (2 * G - 1) * HOP32 + NC;         // registers holding the group's sample span
        const int s0 = f0 * hop - NFFT / 2;
        int q0 = 0;                                          // first noise sample of the span (mix only)
        if (mix && s0 >= 0) { q0 = noff + s0; if (q0 >= nlen) q0 %= nlen; }
        // fast path: the whole span is inside the clip (no reflection); the noise segment may wrap
        // around the end of its clip once (needs a noise clip at least as long as the span)
        if (s0 >= 0 && s0 + 32 * NR <= N && f0 + 2 * G <= T && (!mix || (q0 >= 0 && nlen >= 32 * NR))) {
          staged = true;
          // the span is staged in chunks of CH 32-sample columns so that at most (2G-1)*HOP32 + 32
          // registers are live (n_fft 2048 = 64 columns needs two chunks; everything else one)
          constexpr int CH = NC < 32 ? NC : 32, NCHUNK = (NC + CH - 1) / CH, NRC = (2 * G - 1) * HOP32 + CH;
          static_assert(NC % CH == 0, "column chunks");
#pragma unroll
          for (int ch = 0; ch < NCHUNK; ++ch) {
            const int r0 = ch * CH;
            float sreg[NRC];
            const float* xs = x + s0 + 32 * r0 + lane;
#pragma unroll
            for (int r = 0; r < NRC; ++r) sreg[r] = __ldg(xs + 32 * r);
            if (mix) {
              if (q0 + 32 * NR <= nlen) {
                const float* ns = nz + q0 + 32 * r0 + lane;
#pragma unroll
                for (int r = 0; r < NRC; ++r) sreg[r] = fmaf(scale, __ldg(ns + 32 * r), sreg[r]);
              } else {
#pragma unroll
                for (int r = 0; r < NRC; ++r) {
                  int q = q0 + 32 * (r0 + r) + lane;
                  q -= q >= nlen ? nlen : 0;
                  sreg[r] = fmaf(scale, __ldg(nz + q), sreg[r]);
                }
              }
            }
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              const int j = 32 * (r0 + c) + lane;
              if (NFFT % 32 == 0 || j < NFFT) {
                const float w = s_window[j];
#pragma unroll
                for (int g = 0; g < G; ++g)
                  z[g * ZL + zmap(j)] = make_float2(w * sreg[2 * g * HOP32 + c], w * sreg[(2 * g + 1) * HOP32 + c]);
              }
            }
          }
        }
      }
      if (!staged) {
        // boundary groups (reflect padding, frames >= T, wrapping noise) and hops that are not a
        // multiple of 32: plain per-element gather, deliberately not unrolled (cold code)
#pragma unroll 1
        for (int idx = lane; idx < G * NFFT; idx += 32) {
          const int g = idx / NFFT, j = idx - g * NFFT;
          const int ta = f0 + 2 * g, tb = ta + 1;
          const float w = s_window[j];
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
          z[g * ZL + zmap(j)] = make_float2(re * w, im * w);
        }
      }
      __syncwarp();
      // 2. forward FFT passes (in place, digit-reversed result)
      static_for<0, Rad::npass>([&](auto I) {
        constexpr int i = decltype(I)::value;
        constexpr int R = Rad::R(i), L = Rad::L(i), tasks = NFFT / R;
        const float2* tw = s_tw + Rad::tw_off(i);
#pragma unroll 1   // one copy of each radix butterfly: the hot loop has to stay inside the instruction cache
        for (int u = lane; u < G * tasks; u += 32) {
          const int g = u / tasks, uu = u - g * tasks;
          pass_task<R, false, Map>(z + g * ZL, L, uu, [&](int q) { return tw[q]; });
        }
        __syncwarp();
      });
      // 3. split the packed pair into two power spectra (|A[k]|^2, |B[k]|^2)
      if constexpr (kNatural) {
        // read every (Z[k], Z[n-k]) first, then store the powers in plain bin order
        float2 pw[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const int idx = lane + 32 * i;
          if (idx < G * K) {
            const int g = idx / K, k = idx - g * K;
            const float2* zz = z + g * ZL;
            pw[i] = pair_split_power(zz[zmap(Rad::pos(k))], zz[zmap(Rad::pos(k == 0 ? 0 : NFFT - k))]);
          }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const int idx = lane + 32 * i;
          if (idx < G * K) {
            const int g = idx / K, k = idx - g * K;
            z[g * ZL + zmap(k)] = pw[i];
          }
        }
      } else {
        // in place at Z[k]'s slot (only bin k's lane touches it)
        for (int idx = lane; idx < G * K; idx += 32) {
          const int g = idx / K, k = idx - g * K;
          float2* zz = z + g * ZL;
          const int pk = zmap(Rad::pos(k)), pm = zmap(Rad::pos(k == 0 ? 0 : NFFT - k));
          zz[pk] = pair_split_power(zz[pk], zz[pm]);
        }
      }
      __syncwarp();
      // 4. sparse mel rows + dB: one lane per filter, all 2G frames of the group at once
      for (int m = lane; m < M; m += 32) {
        const int lo = s_mello[m], o0 = s_melofs[m], o1 = s_melofs[m + 1];
        // two independent accumulator chains per frame pair (even / odd taps) hide the LDS + FFMA2 latency
        float2 acc[G], acc1[G];
#pragma unroll
        for (int g = 0; g < G; ++g) { acc[g] = make_float2(0.f, 0.f); acc1[g] = make_float2(0.f, 0.f); }
        int o = o0;
        for (; o + 1 < o1; o += 2) {
          const float w0 = s_melw[o], w1 = s_melw[o + 1];
          const int k = lo + (o - o0);
          const int p0 = zmap(kNatural ? k : Rad::pos(k)), p1 = zmap(kNatural ? k + 1 : Rad::pos(k + 1));
#pragma unroll
          for (int g = 0; g < G; ++g) {
            acc[g] = cfma_s(z[g * ZL + p0], w0, acc[g]);
            acc1[g] = cfma_s(z[g * ZL + p1], w1, acc1[g]);
          }
        }
        if (o < o1) {
          const float w0 = s_melw[o];
          const int k = lo + (o - o0);
          const int p0 = zmap(kNatural ? k : Rad::pos(k));
#pragma unroll
          for (int g = 0; g < G; ++g) acc[g] = cfma_s(z[g * ZL + p0], w0, acc[g]);
        }
#pragma unroll
        for (int g = 0; g < G; ++g) acc[g] = cadd(acc[g], acc1[g]);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const int ta = f0 + 2 * g;
          if (ta < T) { const float d = power_to_db(acc[g].x); tile[m * pitch + ta] = d; vmax = fmaxf(vmax, d); }
          if (ta + 1 < T) { const float d = power_to_db(acc[g].y); tile[m * pitch + ta + 1] = d; vmax = fmaxf(vmax, d); }
        }
      }
      __syncwarp();
    }
    // ---- per-clip top_db floor: max over the tile, gathered while it was written ------------
    vmax = warp_max(vmax);
    if (lane == 0) red[warp] = vmax;
    __syncthreads();
    float cutoff = -INFINITY;
    if (p.top_db >= 0.f) {
      float mx = lane < nwarps ? red[lane] : -INFINITY;
      cutoff = warp_max(mx) - p.top_db;
    }

    OutT* out = reinterpret_cast<OutT*>(p.out) + (size_t)b * p.out_stride;
    const OutT mv = to_out<OutT>(p.mask_value);
    bool bad = false;                                        // any non-finite feature of this clip (pre-mask)
    const float* rsrc = tile;   // rows to normalise in the CMVN epilogue
    bool done = false;

    if (!p.is_mfcc) {
      if (!p.cmvn) {
        for (int m = warp; m < M; m += nwarps) {
          const bool rm = s_rowmask[m] != 0;
          for (int t = lane; t < T; t += 32) {
            const float v = fmaxf(tile[m * pitch + t], cutoff);
            bad |= !isfinite(v);
            out[(size_t)m * T + t] = (rm || s_colmask[t]) ? mv : to_out<OutT>(v);
          }
        }
        done = true;
      } else {
        for (int m = warp; m < M; m += nwarps)
          for (int t = lane; t < T; t += 32) tile[m * pitch + t] = fmaxf(tile[m * pitch + t], cutoff);
      }
    } else {
      // DCT-II: out[c][t] = sum_m dct[m][c] * max(tile[m][t], cutoff).  One task = 8 coefficients x 4
      // frames (frames tb, tb+TB, tb+2TB, tb+3TB so that a warp reads consecutive tile columns): each
      // 16-byte broadcast load of DCT coefficients feeds eight FFMA2.
      const int C = F, c8 = p.c8, ncg = c8 / 8, TB = (T + 3) / 4;
      for (int idx = tid; idx < ncg * TB; idx += blockDim.x) {
        const int cg = idx / TB, tb = idx - cg * TB, c0 = cg * 8;
        float2 acc2[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc2[i][j] = make_float2(0.f, 0.f);
        const float* d = s_dct + c0;
#pragma unroll 2
        for (int m = 0; m < M; ++m) {
          const float4 d0 = *reinterpret_cast<const float4*>(d + m * c8);
          const float4 d1 = *reinterpret_cast<const float4*>(d + m * c8 + 4);
          const float* row = tile + m * pitch + tb;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            // frames beyond T read a neighbouring (finite) tile entry and are never stored
            const float a = fmaxf(row[tb + i * TB < T ? i * TB : 0], cutoff);
            acc2[i][0] = cfma_s(make_float2(d0.x, d0.y), a, acc2[i][0]);   // FFMA2: two coefficients per instruction
            acc2[i][1] = cfma_s(make_float2(d0.z, d0.w), a, acc2[i][1]);
            acc2[i][2] = cfma_s(make_float2(d1.x, d1.y), a, acc2[i][2]);
            acc2[i][3] = cfma_s(make_float2(d1.z, d1.w), a, acc2[i][3]);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int t = tb + i * TB;
          if (t < T) {
            const bool cm = s_colmask[t] != 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = c0 + j;
              const float v = (j & 1) ? acc2[i][j >> 1].y : acc2[i][j >> 1].x;
              if (c < C) {
                bad |= !isfinite(v);
                if (p.cmvn) res[c * pitch + t] = v;
                else out[(size_t)c * T + t] = (cm || s_rowmask[c]) ? mv : to_out<OutT>(v);
              }
            }
          }
        }
      }
      rsrc = res;
      done = !p.cmvn;
    }
    if (!done) {
      // ---- CMVN epilogue: per row (x - mean) / (population std + eps) ------------------
      __syncthreads();
      for (int f = warp; f < F; f += nwarps) {
        float s = 0.f;
        for (int t = lane; t < T; t += 32) s += rsrc[f * pitch + t];
        const float mean = warp_sum(s) / (float)T;
        float q = 0.f;
        for (int t = lane; t < T; t += 32) { const float dlt = rsrc[f * pitch + t] - mean; q = fmaf(dlt, dlt, q); }
        const float sd = sqrtf(warp_sum(q) / (float)T);
        const float inv = 1.0f / (sd + p.cmvn_eps);
        const bool rm = s_rowmask[f] != 0;
        for (int t = lane; t < T; t += 32) {
          const float v = (rsrc[f * pitch + t] - mean) * inv;
          bad |= !isfinite(v);
          out[(size_t)f * T + t] = (rm || s_colmask[t]) ? mv : to_out<OutT>(v);
        }
      }
    }
    // the reference's trainer skips batches with non-finite values (src/training/trainer.py:177-179):
    // give the caller a cheap way to know without scanning the features
    if (__any_sync(0xffffffffu, bad) && lane == 0 && p.nonfinite_flag != nullptr) atomicOr(p.nonfinite_flag, 1);
    __syncthreads();   // tile / flags are reused by the next clip
  }
}

// ==========================================================================================================
// Split path for large batches: the fused kernel above ties a CTA to a clip, so B clips over S CTA slots run
// ceil(B / S) rounds and the CTA-wide per-clip phases cost barrier stalls.  Here the frames of ALL clips form
// one flat queue of warp-sized work items (feat_frames_kernel), the dB tiles go through an L2-resident scratch
// and a second, fine-grained kernel finishes them (feat_epilogue_block_kernel).
// ==========================================================================================================
// ---- per-clip preparation ---------------------------------------------------------------------
template <int kUnused = 0>   // a template only so that the header can be included by several translation units
__global__ void __launch_bounds__(256) feat_prep_kernel(const FeatParams p) {
  __shared__ float red[64];
  pdl_wait();   // launched behind conv_kernel with programmatic dependent launch: nothing is read before this
  const int b = blockIdx.x;
  const bool has_rev = p.rev != nullptr && p.rir_idx != nullptr && __ldg(p.rir_idx + b) >= 0;
  const float* x = has_rev ? p.rev + (size_t)b * p.rev_stride : p.wav + (size_t)b * p.wav_stride;
  const ClipNoise cn = resolve_noise(p.noise, p.noise_idx, p.noise_off, b);
  const float scale = clip_mix_scale(cn, x, p.N, has_rev, p.es_part, p.es_nb, b, p.snr_db, red);
  if (threadIdx.x == 0) {
    p.scale_g[b] = scale;
    p.clip_max[b] = float_key(-INFINITY);
  }
}

// ---- frames: STFT -> power -> mel -> dB into the global tile (warp-autonomous, flat over clips) ----
template <int NFFT, int HOP32>
__global__ void __launch_bounds__(StftPlan<NFFT>::kThreads, StftPlan<NFFT>::kMinCtas) feat_frames_kernel(const FeatParams p) {
  using Plan = StftPlan<NFFT>;
  using Rad = typename Plan::Rad;
  constexpr int G = Plan::G;
  constexpr int K = NFFT / 2 + 1;
  constexpr int NC = (NFFT + 31) / 32;                       // 32-sample columns per frame
  using Map = typename Plan::Map;
  constexpr int ZL = stft_zlen<NFFT>();                      // scratch elements per FFT (with padding)
  const Map zmap;
  constexpr int NI = (G * K + 31) / 32;                      // split items per lane
  static_assert(Rad::n == NFFT, "radix plan");

  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int N = p.N, T = p.T, hop = p.hop, M = p.n_mels, mp = p.mp;

  float* s_window = smem;
  float2* s_tw = reinterpret_cast<float2*>(smem + p.f_off_tw);
  float* s_melw = smem + p.f_off_melw;
  int* s_mello = reinterpret_cast<int*>(smem + p.f_off_mello);
  int* s_melofs = reinterpret_cast<int*>(smem + p.f_off_melofs);
  float2* z = reinterpret_cast<float2*>(smem + p.f_off_z) + (size_t)warp * G * ZL;

  // ---- constants -> shared memory, once per (persistent) CTA; the only CTA-wide barrier ----
  for (int i = tid; i < NFFT; i += blockDim.x) s_window[i] = __ldg(p.window + i);
  for (int i = tid; i < Rad::tw_total; i += blockDim.x) s_tw[i] = __ldg(p.tw + i);
  for (int i = tid; i < p.n_melw; i += blockDim.x) s_melw[i] = __ldg(p.mel_w + i);
  for (int i = tid; i < M; i += blockDim.x) s_mello[i] = __ldg(p.mel_lo + i);
  for (int i = tid; i <= M; i += blockDim.x) s_melofs[i] = __ldg(p.mel_ofs + i);
  // programmatic dependent launch: everything above only reads plan constants and overlapped with the tail of
  // the previous kernel; its results (mix scales, reverberated clips) are needed from here on
  pdl_wait();
  __syncthreads();

  const int ngroups = p.ngroups;
  const long long total = (long long)p.B * ngroups;
  // adjacent warps take adjacent groups of the same clip: their sample spans overlap in L1 / L2
  for (long long item = (long long)blockIdx.x * nwarps + warp; item < total; item += (long long)gridDim.x * nwarps) {
    const int b = (int)(item / ngroups), grp = (int)(item - (long long)b * ngroups);
    const bool has_rev = p.rev != nullptr && p.rir_idx != nullptr && __ldg(p.rir_idx + b) >= 0;
    const float* x = has_rev ? p.rev + (size_t)b * p.rev_stride : p.wav + (size_t)b * p.wav_stride;
    const float scale = __ldg(p.scale_g + b);
    const float* nz = nullptr;
    int noff = 0, nlen = 1;
    if (scale != 0.f || (p.noise_idx != nullptr && p.noise.data != nullptr)) {
      const ClipNoise cn = resolve_noise(p.noise, p.noise_idx, p.noise_off, b);
      nz = cn.nz; noff = cn.off; nlen = cn.len;
    }
    const bool mix = nz != nullptr;
    float vmax = -INFINITY;                                  // maximum of the dB values this lane writes
    const int f0 = grp * 2 * G;
    // 1. load + window: z[g][j] = w[j] * (frame(f0+2g)[j] + i frame(f0+2g+1)[j])
    bool staged = false;
    if constexpr (HOP32 > 0) {
      constexpr int NR = (2 * G - 1) * HOP32 + NC;           // registers holding the group's sample span
      const int s0 = f0 * hop - NFFT / 2;
      int q0 = 0;                                            // first noise sample of the span (mix only)
      if (mix && s0 >= 0) { q0 = noff + s0; if (q0 >= nlen) q0 %= nlen; }
      // fast path: the whole span is inside the clip (no reflection); the noise segment may wrap
      // around the end of its clip once (needs a noise clip at least as long as the span)
      if (s0 >= 0 && s0 + 32 * NR <= N && f0 + 2 * G <= T && (!mix || (q0 >= 0 && nlen >= 32 * NR))) {
        staged = true;
        // the span is staged in chunks of CH 32-sample columns so that at most (2G-1)*HOP32 + 32
        // registers are live (n_fft 2048 = 64 columns needs two chunks; everything else one)
        constexpr int CH = NC < 32 ? NC : 32, NCHUNK = (NC + CH - 1) / CH, NRC = (2 * G - 1) * HOP32 + CH;
        static_assert(NC % CH == 0, "column chunks");
#pragma unroll
        for (int ch = 0; ch < NCHUNK; ++ch) {
          const int r0 = ch * CH;
          float sreg[NRC];
          const float* xs = x + s0 + 32 * r0 + lane;
#pragma unroll
          for (int r = 0; r < NRC; ++r) sreg[r] = __ldg(xs + 32 * r);
          if (mix) {
            if (q0 + 32 * NR <= nlen) {
              const float* ns = nz + q0 + 32 * r0 + lane;
#pragma unroll
              for (int r = 0; r < NRC; ++r) sreg[r] = fmaf(scale, __ldg(ns + 32 * r), sreg[r]);
            } else {
#pragma unroll
              for (int r = 0; r < NRC; ++r) {
                int q = q0 + 32 * (r0 + r) + lane;
                q -= q >= nlen ? nlen : 0;
                sreg[r] = fmaf(scale, __ldg(nz + q), sreg[r]);
              }
            }
          }
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const int j = 32 * (r0 + c) + lane;
            if (NFFT % 32 == 0 || j < NFFT) {
              const float w = s_window[j];
#pragma unroll
              for (int g = 0; g < G; ++g)
                z[g * ZL + zmap(j)] = make_float2(w * sreg[2 * g * HOP32 + c], w * sreg[(2 * g + 1) * HOP32 + c]);
            }
          }
        }
      }
    }
    if (!staged) {
      // boundary groups (reflect padding, frames >= T, wrapping noise) and hops that are not a
      // multiple of 32: plain per-element gather, deliberately not unrolled (cold code)
#pragma unroll 1
      for (int idx = lane; idx < G * NFFT; idx += 32) {
        const int g = idx / NFFT, j = idx - g * NFFT;
        const int ta = f0 + 2 * g, tb = ta + 1;
        const float w = s_window[j];
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
        z[g * ZL + zmap(j)] = make_float2(re * w, im * w);
      }
    }
    __syncwarp();
    // 2. forward FFT passes (in place, digit-reversed result)
    static_for<0, Rad::npass>([&](auto I) {
      constexpr int i = decltype(I)::value;
      constexpr int R = Rad::R(i), L = Rad::L(i), tasks = NFFT / R;
      const float2* tw = s_tw + Rad::tw_off(i);
#pragma unroll 1   // one copy of each radix butterfly: the hot loop has to stay inside the instruction cache
      for (int u = lane; u < G * tasks; u += 32) {
        const int g = u / tasks, uu = u - g * tasks;
        pass_task<R, false, Map>(z + g * ZL, L, uu, [&](int q) { return tw[q]; });
      }
      __syncwarp();
    });
    // 3. split the packed pair into two power spectra (|A[k]|^2, |B[k]|^2): read every (Z[k], Z[n-k])
    //    first, then store the powers in plain bin order
    {
      float2 pw[NI];
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int idx = lane + 32 * i;
        if (idx < G * K) {
          const int g = idx / K, k = idx - g * K;
          const float2* zz = z + g * ZL;
          pw[i] = pair_split_power(zz[zmap(Rad::pos(k))], zz[zmap(Rad::pos(k == 0 ? 0 : NFFT - k))]);
        }
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int idx = lane + 32 * i;
        if (idx < G * K) {
          const int g = idx / K, k = idx - g * K;
          z[g * ZL + zmap(k)] = pw[i];
        }
      }
    }
    __syncwarp();
    // 4. sparse mel rows + dB: one lane per filter, all 2G frames of the group at once
    float* tg = p.tile_g + ((size_t)b * T + f0) * mp;
    for (int m = lane; m < M; m += 32) {
      const int lo = s_mello[m], o0 = s_melofs[m], o1 = s_melofs[m + 1];
      // two independent accumulator chains per frame pair (even / odd taps) hide the LDS + FFMA2 latency
      float2 acc[G], acc1[G];
#pragma unroll
      for (int g = 0; g < G; ++g) { acc[g] = make_float2(0.f, 0.f); acc1[g] = make_float2(0.f, 0.f); }
      int o = o0;
      for (; o + 1 < o1; o += 2) {
        const float w0 = s_melw[o], w1 = s_melw[o + 1];
        const int k = lo + (o - o0);
        const int p0 = zmap(k), p1 = zmap(k + 1);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          acc[g] = cfma_s(z[g * ZL + p0], w0, acc[g]);
          acc1[g] = cfma_s(z[g * ZL + p1], w1, acc1[g]);
        }
      }
      if (o < o1) {
        const float w0 = s_melw[o];
        const int p0 = zmap(lo + (o - o0));
#pragma unroll
        for (int g = 0; g < G; ++g) acc[g] = cfma_s(z[g * ZL + p0], w0, acc[g]);
      }
#pragma unroll
      for (int g = 0; g < G; ++g) acc[g] = cadd(acc[g], acc1[g]);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int ta = f0 + 2 * g;
        // frame-major tile: the 32 lanes of a round write 32 consecutive floats of one frame
        if (ta < T) { const float d = power_to_db(acc[g].x); tg[(size_t)(2 * g) * mp + m] = d; vmax = fmaxf(vmax, d); }
        if (ta + 1 < T) { const float d = power_to_db(acc[g].y); tg[(size_t)(2 * g + 1) * mp + m] = d; vmax = fmaxf(vmax, d); }
      }
    }
    vmax = warp_max(vmax);
    if (lane == 0 && vmax > -INFINITY) atomicMax(p.clip_max + b, float_key(vmax));
    __syncwarp();
  }
}

// ---- block epilogue of the split path: top_db floor -> [DCT-II] -> [SpecAugment] -> store ----------------
// One CTA per (clip, block of eb_frames frames).  The block's rows of the frame-major tile are one contiguous
// span of global memory: read with 16-byte coalesced loads, written transposed into shared memory
// [n_mels][pitch], then the same 8-coefficient x 4-frame register tiles as the fused kernel's DCT phase.
// Many small CTAs with short dependent phases: the grid is 2-3 k CTAs, 8+ resident per SM.
// exact i / d for 0 <= i < 2^16, 1 <= d < 2^10 without an integer division
__device__ __forceinline__ int small_div(int i, float inv_d) { return __float2int_rd(((float)i + 0.5f) * inv_d); }

template <typename OutT>
__global__ void __launch_bounds__(256) feat_epilogue_block_kernel(const FeatParams p) {
  extern __shared__ __align__(16) float smem[];               // tile [n_mels][pitch] | dct [n_mels][c8]
  __shared__ unsigned char s_rowmask[128];
  __shared__ unsigned char s_colmask[256];
  const int tid = threadIdx.x;
  const int T = p.T, M = p.n_mels, F = p.n_feat, mp = p.mp, pitch = p.eb_pitch, c8 = p.c8;
  const int nblk = (T + p.eb_frames - 1) / p.eb_frames;
  float* tile = smem;
  float* s_dct = smem + ((M * pitch + 3) & ~3);
  const int mq = mp / 4, cq = c8 / 4;
  const float inv_mq = 1.0f / (float)mq;
  if (p.is_mfcc) {                                            // DCT matrix: once per (persistent) CTA
    const float inv_cq = 1.0f / (float)cq;
    for (int i = tid; i < M * cq; i += blockDim.x) {          // c8 is a multiple of 8: rows in float4 units
      const int m = small_div(i, inv_cq), c = 4 * (i - m * cq);
      const float* src = p.dct + (size_t)m * p.n_mfcc + c;
      reinterpret_cast<float4*>(s_dct)[i] = make_float4(c < p.n_mfcc ? __ldg(src) : 0.f, c + 1 < p.n_mfcc ? __ldg(src + 1) : 0.f,
                                                        c + 2 < p.n_mfcc ? __ldg(src + 2) : 0.f, c + 3 < p.n_mfcc ? __ldg(src + 3) : 0.f);
    }
  }
  pdl_wait();                            // the tiles and maxima of feat_frames_kernel (PDL)
  const OutT mv = to_out<OutT>(p.mask_value);
  float chk = 0.f;                                            // v * 0 accumulates to NaN iff some v is NaN / Inf
  const int total = p.B * nblk;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int b = item / nblk, t0 = (item - b * nblk) * p.eb_frames;
    const int nf = min(p.eb_frames, T - t0);                  // frames in this block
    for (int i = tid; i < F + nf; i += blockDim.x) {          // SpecAugment flags: feature rows, then this block's frames
      const bool is_row = i < F;
      const int q = is_row ? i : t0 + (i - F);
      const int32_t* st = is_row ? p.fs : p.ts;
      const int32_t* ln = is_row ? p.fl : p.tl;
      const int nm = is_row ? p.nF : p.nT;
      bool mk = false;
      if (st != nullptr)
        for (int j = 0; j < nm; ++j) {
          const int s0 = __ldg(st + (size_t)b * nm + j), l = __ldg(ln + (size_t)b * nm + j);
          mk |= (q >= s0) && (q < s0 + l);
        }
      (is_row ? s_rowmask : s_colmask)[is_row ? i : i - F] = mk ? 1 : 0;
    }
    float cutoff = -INFINITY;
    if (p.top_db >= 0.f) cutoff = clip_max_value(p.clip_max[b]) - p.top_db;
    // the block's rows of the frame-major tile are one contiguous span: coalesced 16-byte loads, transposed into smem
    const float4* tg = reinterpret_cast<const float4*>(p.tile_g + ((size_t)b * T + t0) * mp);
    for (int i = tid; i < nf * mq; i += blockDim.x) {
      const int tl = small_div(i, inv_mq), m = 4 * (i - tl * mq);
      const float4 v = tg[i];
      tile[m * pitch + tl] = fmaxf(v.x, cutoff);
      if (m + 1 < M) tile[(m + 1) * pitch + tl] = fmaxf(v.y, cutoff);
      if (m + 2 < M) tile[(m + 2) * pitch + tl] = fmaxf(v.z, cutoff);
      if (m + 3 < M) tile[(m + 3) * pitch + tl] = fmaxf(v.w, cutoff);
    }
    __syncthreads();
    OutT* out = reinterpret_cast<OutT*>(p.out) + (size_t)b * p.out_stride + t0;
    if (!p.is_mfcc) {
      const float inv_nf = 1.0f / (float)nf;
      for (int i = tid; i < M * nf; i += blockDim.x) {
        const int m = small_div(i, inv_nf), tl = i - m * nf;
        const float v = tile[m * pitch + tl];
        chk = fmaf(v, 0.f, chk);
        out[(size_t)m * T + tl] = (s_rowmask[m] || s_colmask[tl]) ? mv : to_out<OutT>(v);
      }
    } else {
      // DCT-II in 8-coefficient x 4-frame register tiles (frames tb, tb+TB, tb+2TB, tb+3TB: a warp reads
      // consecutive tile columns; each 16-byte broadcast load of coefficients feeds eight FFMA2)
      const int C = F, ncg = c8 / 8, TB = (nf + 3) / 4;
      const float inv_tb = 1.0f / (float)TB;
      for (int idx = tid; idx < ncg * TB; idx += blockDim.x) {
        const int cg = small_div(idx, inv_tb), tb = idx - cg * TB, c0 = cg * 8;
        // frames beyond the block read a neighbouring (finite) tile entry and are never stored
        const int o1 = tb + TB < nf ? TB : 0, o2 = tb + 2 * TB < nf ? 2 * TB : 0, o3 = tb + 3 * TB < nf ? 3 * TB : 0;
        float2 acc2[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc2[i][j] = make_float2(0.f, 0.f);
        const float* d = s_dct + c0;
        const float* row = tile + tb;
#pragma unroll 2
        for (int m = 0; m < M; ++m, d += c8, row += pitch) {
          const float4 d0 = *reinterpret_cast<const float4*>(d);
          const float4 d1 = *reinterpret_cast<const float4*>(d + 4);
          const float a[4] = {row[0], row[o1], row[o2], row[o3]};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc2[i][0] = cfma_s(make_float2(d0.x, d0.y), a[i], acc2[i][0]);
            acc2[i][1] = cfma_s(make_float2(d0.z, d0.w), a[i], acc2[i][1]);
            acc2[i][2] = cfma_s(make_float2(d1.x, d1.y), a[i], acc2[i][2]);
            acc2[i][3] = cfma_s(make_float2(d1.z, d1.w), a[i], acc2[i][3]);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int tl = tb + i * TB;
          if (tl < nf) {
            const bool cm = s_colmask[tl] != 0;
            OutT* o = out + (size_t)c0 * T + tl;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float v = (j & 1) ? acc2[i][j >> 1].y : acc2[i][j >> 1].x;
              if (c0 + j < C) {
                chk = fmaf(v, 0.f, chk);
                o[(size_t)j * T] = (cm || s_rowmask[c0 + j]) ? mv : to_out<OutT>(v);
              }
            }
          }
        }
      }
    }
    __syncthreads();                                          // tile and flags are reused by the next block
  }
  if (__any_sync(0xffffffffu, chk != 0.f) && (tid & 31) == 0 && p.nonfinite_flag != nullptr) atomicOr(p.nonfinite_flag, 1);
}

}  // namespace wwf
