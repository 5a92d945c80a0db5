// wwf_conv.cuh - RIR reverb as FFT overlap-save convolution (SURVEY.md section 8a row A2;
// oracle: torchaudio F.fftconvolve(x, h, "full")[..., :N], TA/functional/functional.py:2255-2258).
//
// One CTA convolves one block of P = 2M = 32768 real samples of one clip entirely in shared
// memory:  the P reals are read as M complex numbers z[m] = x[2m] + i x[2m+1] (free: it is
// the same memory), transformed by an M-point in-place DIF FFT (radices 4,16,16,16, result
// digit-reversed), turned into the P-point real spectrum and multiplied by the RIR's
// pre-scaled spectrum pair-by-pair (k, M-k), folded back, and inverse-transformed by the
// adjoint DIT passes, which consume the digit-reversed order - no reordering pass exists.
// A clip with N + Lmax - 1 <= P is a single block (plain zero-padded linear convolution);
// longer clips use overlap-save blocks with H0 = roundup4(Lmax-1) samples of history.
//
// Shared-memory index map pad(i) = i + (i >> 4) makes every pass bank-conflict free for
// 8-byte elements ("K_conv shared-memory layout").
#pragma once
#include <stdint.h>
#include "wwf_fft.cuh"

namespace wwf {

constexpr int kConvLogM = 14;
constexpr int kConvM = 1 << kConvLogM;       // complex FFT length
constexpr int kConvP = 2 * kConvM;           // real block length
constexpr int kConvThreads = 512;
constexpr int kConvPairTasks = kConvM / 2;   // k in [0, M/2) enumerated by pair_task_index; k = M/2 is extra
using ConvRad = Radices<4, 16, 16, 16>;

struct PadMap {
  WWF_HD int operator()(int i) const { return i + (i >> 4); }
};
constexpr int kConvSmemElems = kConvM + (kConvM >> 4);

// Pair-pass task v in [0, M/2] -> frequency k (k <= M/2) and its partner M-k.  Tasks are ordered
// so that the 16 lanes of a half-warp touch 16 distinct 8-byte banks both at pos(k) and at
// pos(M-k): lanes vary the top digit d3 (consecutive positions) over 0..7 for two 16-runs
// that are 8 runs apart.
WWF_HD int pair_task_k(int v) {
  if (v >= kConvPairTasks) return kConvM / 2;
  const int d3 = v & 7;                                   // top digit of k, 0..7  (k < M/2)
  const int a = (v >> 7) * 16 + ((v >> 3) & 1) * 8 + ((v >> 4) & 7);  // run index = pos >> 4
  const int d0 = a >> 8, d1 = (a >> 4) & 15, d2 = a & 15; // pos = d0*4096 + d1*256 + d2*16 + d3
  return d0 + 4 * d1 + 64 * d2 + 1024 * d3;
}

struct ConvParams {
  const float* wav; int64_t wav_stride;      // [B][N]
  float* rev; int64_t rev_stride;            // [B][N] output (workspace)
  float* es_part; int es_nb;                 // [B][es_nb] energy of this block's output samples (for the SNR mix)
  const int32_t* rir_idx;                    // [B]
  int B, N, n_rir;
  int hist;                                  // H0: history samples per block (0 = single block)
  int valid;                                 // V = P - H0 output samples per block
  const float4* spec;                        // [n_rir][M/2+1]: (H''[k], H''[M-k]) in pair-task order
  const float2* tw;                          // pass tables of ConvRad (radix-4 pass: only r = 1)
  const float2* tw_pair;                     // [M/2+1] w_P^k in pair-task order
};

// twiddle table layout for the conv FFT: pass 0 (radix 4, s = 4096) stores only w^j (r = 1);
// w^{2j}, w^{3j} are formed by multiplication.  Passes 1, 2 store the full (R-1)*s tables.
constexpr int kConvTw0 = 0;
constexpr int kConvTw1 = kConvTw0 + ConvRad::S(0);
constexpr int kConvTw2 = kConvTw1 + 15 * ConvRad::S(1);
constexpr int kConvTwTotal = kConvTw2 + 15 * ConvRad::S(2);
// dynamic shared memory: padded data array followed by a copy of the pass twiddle tables
constexpr size_t kConvSmemBytes = (size_t)(kConvSmemElems + kConvTwTotal) * sizeof(float2);

// copy the pass tables into shared memory (once per persistent CTA)
__device__ __forceinline__ void conv_load_tables(float2* s_tw, const float2* __restrict__ tw) {
  for (int i = threadIdx.x; i < kConvTwTotal; i += kConvThreads) s_tw[i] = __ldg(tw + i);
}

template <bool INV>
__device__ __forceinline__ void conv_fft_passes(float2* z, const float2* tw) {   // tw: shared-memory copy
  const int tid = threadIdx.x;
  auto pass4 = [&]() {
    for (int u = tid; u < kConvM / 4; u += kConvThreads) {
      const float2 w1 = tw[kConvTw0 + u];
      const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1);
      pass_task<4, INV, PadMap>(z, ConvRad::L(0), u, [&](int q) {
        const int r = q / ConvRad::S(0);     // q = (r-1)*s + j
        return r == 0 ? w1 : (r == 1 ? w2 : w3);
      });
    }
  };
  auto pass16 = [&](int L, const float2* t) {
    for (int u = tid; u < kConvM / 16; u += kConvThreads)
      pass_task<16, INV, PadMap>(z, L, u, [&](int q) { return t[q]; });
  };
  if constexpr (!INV) {
    pass4();                             __syncthreads();
    pass16(ConvRad::L(1), tw + kConvTw1); __syncthreads();
    pass16(ConvRad::L(2), tw + kConvTw2); __syncthreads();
    pass16(ConvRad::L(3), tw);            __syncthreads();   // s == 1: table unused
  } else {
    pass16(ConvRad::L(3), tw);            __syncthreads();
    pass16(ConvRad::L(2), tw + kConvTw2); __syncthreads();
    pass16(ConvRad::L(1), tw + kConvTw1); __syncthreads();
    pass4();                             __syncthreads();
  }
}

// Load one block of P reals (clip samples [start, start+P), zero outside [0, N)) as M complex.
// All 16 float4 loads of a thread are issued before the first shared-memory store, so the
// HBM latency is paid once per block instead of once per load.
__device__ __forceinline__ void conv_load_block(float2* z, const float* __restrict__ x, int N, int start, bool vec_ok) {
  PadMap pad;
  constexpr int kPer = kConvP / 4 / kConvThreads;     // float4 per thread
  float4 v[kPer];
#pragma unroll
  for (int u = 0; u < kPer; ++u) {
    const int n = start + 4 * (threadIdx.x + u * kConvThreads);
    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n >= 0 && n + 3 < N && vec_ok) {
      v[u] = __ldg(reinterpret_cast<const float4*>(x + n));
    } else if (n + 3 >= 0 && n < N) {
      if (n >= 0) v[u].x = __ldg(x + n);
      if (n + 1 >= 0 && n + 1 < N) v[u].y = __ldg(x + n + 1);
      if (n + 2 >= 0 && n + 2 < N) v[u].z = __ldg(x + n + 2);
      if (n + 3 < N) v[u].w = __ldg(x + n + 3);
    }
  }
#pragma unroll
  for (int u = 0; u < kPer; ++u) {
    const int q = threadIdx.x + u * kConvThreads;
    z[pad(2 * q)] = make_float2(v[u].x, v[u].y);
    z[pad(2 * q + 1)] = make_float2(v[u].z, v[u].w);
  }
}

// One (k, M-k) pair of the real-spectrum step.  A = Z[k], Bm = Z[M-k], w = w_P^k.
// Returns the forward half:  R2k = 2 R[k],  R2m = 2 R[M-k]  (R = P-point real spectrum).
WWF_HD void pair_forward(float2 A, float2 Bm, float2 w, float2& R2k, float2& R2m) {
  const float2 Bc = cconj(Bm);
  const float2 Se = cadd(A, Bc);                 // 2 Xe[k]
  const float2 So = mul_mi<false>(csub(A, Bc));  // 2 Xo[k] = -i (A - conj B)
  const float2 Tt = cmul(w, So);
  R2k = cadd(Se, Tt);
  R2m = cconj(csub(Se, Tt));
}
// Inverse half: from Yk, Ym (spectrum of the result at k and M-k) to Zy[k], Zy[M-k].
WWF_HD void pair_inverse(float2 Yk, float2 Ym, float2 w, float2& Zk, float2& Zm) {
  const float2 Yc = cconj(Ym);
  const float2 Ue = cadd(Yk, Yc);
  const float2 Uo = cmulc(csub(Yk, Yc), w);      // conj(w^k) (Yk - conj Ym)
  Zk = make_float2(Ue.x - Uo.y, Ue.y + Uo.x);    // Ue + i Uo
  Zm = make_float2(Ue.x + Uo.y, Uo.x - Ue.y);    // conj(Ue) + i conj(Uo)
}

// The (k, M-k) pass of one block: 16 pair tasks per thread, global operands (pair twiddle, RIR
// spectrum) fetched four tasks ahead of their use.
__device__ __forceinline__ void conv_pair_pass(float2* zc, const float4* __restrict__ spec, const float2* __restrict__ tw_pair) {
  PadMap pad;
  auto one = [&](int v, float2 w, float4 h) {
    const int k = pair_task_k(v);
    const int pk = pad(ConvRad::pos(k)), pm = pad(ConvRad::pos((kConvM - k) & (kConvM - 1)));
    float2 R2k, R2m, Zk, Zm;
    pair_forward(zc[pk], zc[pm], w, R2k, R2m);
    pair_inverse(cmul(R2k, make_float2(h.x, h.y)), cmul(R2m, make_float2(h.z, h.w)), w, Zk, Zm);
    zc[pm] = Zm;
    zc[pk] = Zk;   // k == 0 and k == M/2 are self-paired: Zk == Zm there
  };
  static_assert(kConvPairTasks % (4 * kConvThreads) == 0, "pair pass unroll");
  for (int v0 = threadIdx.x; v0 < kConvPairTasks; v0 += 4 * kConvThreads) {
    float2 w[4];
    float4 h[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { w[u] = __ldg(tw_pair + v0 + u * kConvThreads); h[u] = __ldg(spec + v0 + u * kConvThreads); }
#pragma unroll
    for (int u = 0; u < 4; ++u) one(v0 + u * kConvThreads, w[u], h[u]);
  }
  if (threadIdx.x == 0) one(kConvPairTasks, __ldg(tw_pair + kConvPairTasks), __ldg(spec + kConvPairTasks));
}

// Persistent: grid = min(#SMs, work items); work item = (clip b, overlap-save block blk).
__global__ void __launch_bounds__(kConvThreads, 1) conv_kernel(const ConvParams p) {
  extern __shared__ __align__(16) float2 zc[];
  __shared__ float red[kConvThreads / 32];
  float2* s_tw = zc + kConvSmemElems;
  conv_load_tables(s_tw, p.tw);
  PadMap pad;
  const int nblk = p.es_nb;
  for (int item = blockIdx.x; item < p.B * nblk; item += gridDim.x) {
    const int b = item / nblk, blk = item - b * nblk;
    const int r = __ldg(p.rir_idx + b);
    if (r < 0 || r >= p.n_rir) continue;                     // dry clip (CTA-uniform)
    const float* x = p.wav + (size_t)b * p.wav_stride;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    __syncthreads();                                          // previous item's stores / table copy done
    conv_load_block(zc, x, p.N, blk * p.valid - p.hist, vec_ok);
    __syncthreads();
    conv_fft_passes<false>(zc, s_tw);
    conv_pair_pass(zc, p.spec + (size_t)r * (kConvPairTasks + 1), p.tw_pair);
    __syncthreads();
    conv_fft_passes<true>(zc, s_tw);

    // store the valid outputs: block sample i in [hist, P) -> clip sample blk*valid + i - hist
    float* y = p.rev + (size_t)b * p.rev_stride;
    const bool st_ok = ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
    float e0 = 0.f, e1 = 0.f;
    for (int q = threadIdx.x; q < kConvP / 4; q += kConvThreads) {
      const int i = 4 * q;
      if (i < p.hist) continue;
      const int n = blk * p.valid + i - p.hist;
      if (n >= p.N) continue;
      const float2 a = zc[pad(2 * q)], c = zc[pad(2 * q + 1)];
      if (n + 3 < p.N && st_ok) {
        *reinterpret_cast<float4*>(y + n) = make_float4(a.x, a.y, c.x, c.y);
        e0 = fmaf(a.x, a.x, fmaf(a.y, a.y, e0));
        e1 = fmaf(c.x, c.x, fmaf(c.y, c.y, e1));
      } else {
        y[n] = a.x; e0 = fmaf(a.x, a.x, e0);
        if (n + 1 < p.N) { y[n + 1] = a.y; e0 = fmaf(a.y, a.y, e0); }
        if (n + 2 < p.N) { y[n + 2] = c.x; e1 = fmaf(c.x, c.x, e1); }
        if (n + 3 < p.N) { y[n + 3] = c.y; e1 = fmaf(c.y, c.y, e1); }
      }
    }
    // energy of this block's output samples, for the SNR mix that follows (fixed reduction order)
    float e = e0 + e1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x < 32) {
      e = threadIdx.x < kConvThreads / 32 ? red[threadIdx.x] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
      if (threadIdx.x == 0 && p.es_part != nullptr) p.es_part[(size_t)b * nblk + blk] = e;
    }
  }
}

// Spectrum of one zero-padded RIR in the layout conv_kernel consumes:
// spec[v] = (R[k], R[M-k]) / (4M)  with k = pair_task_k(v), R = rfft(h, P).
// (R2 = 2R, so the stored value is R2 / (8M); the 1/(4M) folds the two 1/2 of the even/odd
// split and the 1/M of the unnormalised inverse.)
struct SpecParams {
  const float* data; const int64_t* offsets; int n_rir;
  float4* spec; const float2* tw; const float2* tw_pair;
};

__global__ void __launch_bounds__(kConvThreads, 1) rir_spectrum_kernel(const SpecParams p) {
  extern __shared__ __align__(16) float2 zc[];
  const int r = blockIdx.x;
  const int64_t o0 = p.offsets[r], o1 = p.offsets[r + 1];
  const float* h = p.data + o0;
  float2* s_tw = zc + kConvSmemElems;
  conv_load_tables(s_tw, p.tw);
  conv_load_block(zc, h, (int)(o1 - o0), 0, false);
  __syncthreads();
  conv_fft_passes<false>(zc, s_tw);
  PadMap pad;
  const float sc = 1.0f / (8.0f * (float)kConvM);
  float4* spec = p.spec + (size_t)r * (kConvPairTasks + 1);
  for (int v = threadIdx.x; v <= kConvPairTasks; v += kConvThreads) {
    const int k = pair_task_k(v);
    const int pk = pad(ConvRad::pos(k)), pm = pad(ConvRad::pos((kConvM - k) & (kConvM - 1)));
    float2 R2k, R2m;
    pair_forward(zc[pk], zc[pm], __ldg(p.tw_pair + v), R2k, R2m);
    if (k == 0) { R2k.y = 0.f; R2m.y = 0.f; }   // DC and Nyquist of a real signal are real
    spec[v] = make_float4(R2k.x * sc, R2k.y * sc, R2m.x * sc, R2m.y * sc);
  }
}

}  // namespace wwf
