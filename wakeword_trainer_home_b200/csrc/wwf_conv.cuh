// wwf_conv.cuh - RIR reverb as FFT overlap-save convolution (SURVEY.md section 8a row A2;
// oracle: torchaudio F.fftconvolve(x, h, "full")[..., :N], TA/functional/functional.py:2255-2258).
//
// One CTA convolves one block of P = 2M = 32768 real samples of one clip entirely in shared
// memory.  The P reals are read as M = 16384 complex numbers z[m] = x[2m] + i x[2m+1] (free: it
// is the same memory) and transformed by an in-place DIF FFT with radices 32, 32, 16 whose
// output is digit-reversed: frequency k = d0 + 32 d1 + 1024 d2 sits at p = 512 d0 + 16 d1 + d2.
// Only the two radix-32 passes go through shared memory.  The last forward pass (radix 16 on 16
// contiguous elements = one "run"), the real-spectrum unpacking, the multiplication by the
// RIR spectrum, the re-packing and the first inverse pass are FUSED in registers: the partner
// of frequency k = l + 1024 d2 is M - k = (1024 - l) + 1024 (15 - d2), i.e. run(l) pairs with
// run(1024 - l) element d2 <-> 15 - d2, so one thread owns both runs (32 complex values), and
// the inverse consumes the digit-reversed order - no reordering or separate pair pass exists.
// Shared-memory traffic per block: 4 stores + 4 loads of the block (the first forward pass reads global memory,
// the last inverse pass writes it; was 10 round trips with radices 4,16,16,16 and a separate pair pass); two CTA
// barriers per block, one on either side of a middle in which every warp works on its own pair of 512-element
// sub-transforms (WARP-LOCAL MIDDLE below); the next item's inputs arrive by cp.async under the last pass.
//
// A clip with N + Lmax - 1 <= P is a single block (plain zero-padded linear convolution);
// longer clips use overlap-save blocks with H0 = roundup4(Lmax-1) samples of history.
//
// Index map pad(i) = i + (i >> 4) keeps every access pattern conflict-free for 8-byte elements:
// the radix-32 passes touch 16 consecutive elements per half-warp, and the fused tasks are
// ordered (host table) so that the 16 lanes of a half-warp own runs with distinct (run mod 16).
#pragma once
#include <stdint.h>
#include "wwf_fft.cuh"
#include "wwf_mix.cuh"

namespace wwf {

constexpr int kConvLogM = 14;
constexpr int kConvM = 1 << kConvLogM;       // complex FFT length
constexpr int kConvP = 2 * kConvM;           // real block length
constexpr int kConvThreads = 512;
using ConvRad = Radices<32, 32, 16>;
constexpr int kRunLen = 16;                  // last radix: 16 contiguous positions = one run
constexpr int kRuns = kConvM / kRunLen;      // 1024 runs, run(l) for l = k mod 1024
constexpr int kFusedTasks = kRuns / 2;       // 511 run pairs {l, 1024-l} + the self-paired run l = 512 (kFusedSelfTask);
                                             // the other self-paired run, l = 0 (DC / Nyquist), is fused_dc_task
constexpr int kFusedSelfTask = 0;            // the thread that owns l = 512 also does the run l = 0

// WARP-LOCAL MIDDLE.  Position p = 512 d0 + 16 d1 + d2 holds frequency k = d0 + 32 d1 + 1024 d2, so the 512-element
// sub-transform d0 (what pass 1 works on) holds the frequencies k = d0 mod 32, and the partner M - k of any of them
// lies in sub-transform (32 - d0) mod 32.  Everything between the two radix-32 passes over the whole block - forward
// pass 1, the fused run pairs, inverse pass 1 - therefore closes over the sub-transform PAIRS {d0, 32 - d0}
// (d0 = 1..15) and {0, 16} (both self-paired): 16 pairs of 1024 elements = 32 radix-32 tasks and 32 run pairs each,
// exactly one per lane of the 16 warps.  Warp w owns pair w: lanes 0-15 take sub-transform conv_sub_a(w), lanes 16-31
// conv_sub_b(w) in pass 1; lane i owns the runs l = w + 32 i and 1024 - l in the fused phase.  Only __syncwarp()
// separates the three phases (the CTA barriers around pass 0 remain), so the warps drift apart and one warp's
// shared-memory traffic overlaps another's arithmetic.
WWF_HD int conv_sub_a(int w) { return w; }
WWF_HD int conv_sub_b(int w) { return w == 0 ? 16 : 32 - w; }
// l of fused task t = 32 w + lane (host table fused_l; any l in [1, 1023] names the pair {l, 1024 - l})
WWF_HD int conv_fused_l(int t) {
  const int w = t >> 5, i = t & 31;
  if (w > 0) return w + 32 * i;
  if (i >= 16) return 16 + 32 * (i - 16);     // sub-transform 16 pairs with itself: d1 <-> 31 - d1
  return i == 0 ? 512 : 32 * i;               // sub-transform 0: d1 <-> 32 - d1; d1 = 16 is l = 512, d1 = 0 the DC task
}

constexpr int kConvSmemElems = kConvM + (kConvM >> 4);

// Twiddle tables (float2), one copy in shared memory per persistent CTA:
//   pass 0 (radix 32, s = 512): powers w^{j}, w^{2j}, w^{4j}, w^{8j}, w^{16j}  [5][512]; the other 26
//       of the 31 output twiddles w^{jr} are products of at most three of them (formed in registers)
//   pass 1 (radix 32, s = 16): full table [(r-1)*16 + j] = w_512^{jr}
constexpr int kConvTw0 = 0;
constexpr int kConvTw1 = kConvTw0 + 5 * ConvRad::S(0);
constexpr int kConvTwTotal = kConvTw1 + 31 * ConvRad::S(1);
constexpr size_t kConvSmemBytes = (size_t)(kConvSmemElems + kConvTwTotal) * sizeof(float2);

// Per-RIR spectrum layout consumed by the fused task (float4 = (H''[k], H''[M-k]), H'' = rfft(h,P)/(4M)):
//   general  [r*512 + t]      k = l(t) + 1024 r,  t < 512 (task order), r < 16
//   special  [16*512 + i]     i < 9: k = 1024 i   (the run l = 0)
constexpr int kSpecSpecial = 16 * 512;
constexpr int kSpecPerRir = kSpecSpecial + 9;

struct ConvParams {
  const float* wav; int64_t wav_stride;      // [B][N]
  float* rev; int64_t rev_stride;            // [B][N] output (workspace)
  float* es_part; int es_nb;                 // [B][es_nb] energy of each block's output samples (for the SNR mix)
  const int32_t* rir_idx;                    // [B]
  int B, N, n_rir;
  int hist;                                  // H0: history samples per block (0 = single block)
  int valid;                                 // V = P - H0 output samples per block
  const float4* spec;                        // [n_rir][kSpecPerRir]
  const float2* tw;                          // pass tables (kConvTwTotal)
  const uint16_t* fused_l;                   // [512] l of fused task t (bank-conflict-free order)
  const float2* fused_tw;                    // [512] w_P^{l(t)}
  int ordered;                               // 1: work items follow the CTA-built clip order (reverberated clips first)
  int* clip_max; int n_clip_max;             // flat feature path: running clip maxima to reset (or nullptr): saves the
                                             // memset in front of this kernel, so that it can be chained (PDL) too
  // noise-mix records of the flat feature path made HERE (single-block clips, at most kConvMaxOwn items per CTA), or
  // mix_g == nullptr: feat_prep_kernel makes them
  ClipMix* mix_g;                            // [B]
  NoiseBankDev noise; const int32_t* noise_idx; const int64_t* noise_off; const float* snr_db;
};
constexpr int kConvMaxOwn = 16;              // items per CTA whose records fit the shared-memory staging area
// ORDER OF THE WORK ITEMS.  Item i of the persistent grid goes to CTA i mod grid.  With clip order = batch order, a call
// that reverberates only some clips (rir_prob 0.25 - 0.3 in the reference's presets) gives every CTA a random number
// of expensive items: BASELINE configs[3] (1024 clips, 30 % reverberated, 2 blocks each) averaged 4.2 blocks per CTA
// but the slowest had 9 - and the kernel takes as long as the slowest.  Every CTA therefore builds the same
// permutation in shared memory first: the reverberated clips in batch order, then the dry ones (one pass over
// rir_idx + a block-wide prefix sum, ~1 us), and items are (order[i / nblk], i mod nblk): the expensive items are
// dealt round-robin, at most one more on one CTA than on another.  Batches of up to kConvMaxOrder clips (uint16 ids).
constexpr int kConvMaxOrder = 8192;
constexpr size_t conv_smem_bytes(int B, bool ordered) { return kConvSmemBytes + (ordered ? ((size_t)B * 2 + 15) / 16 * 16 : 0); }

// ------------------------------------------------------------------------------------------
// radix-32 pass with derived twiddles (pass 0).  tw5[b*s + j] = w_L^{j 2^b}, b = 0..4.
// ------------------------------------------------------------------------------------------
// register-level core: v[q] = element j + q s of a length-L sub-transform (s = L / 32); forward: DFT_32 then the
// output twiddles w_L^{j r}; inverse: their conjugates, then the inverse DFT_32.
struct NoAfter { WWF_HD void operator()(int) const {} };
// mid(0) runs once every input register has been consumed by the (inverse) twiddle multiplications
template <bool INV, class TwLoad, class Mid = NoAfter>
WWF_HD void pass32_core(float2 (&v)[32], int s, int j, TwLoad tw5, Mid mid = Mid()) {
  constexpr int R = 32;
  const float2 w1 = tw5(j), w2 = tw5(s + j), w4 = tw5(2 * s + j), w8 = tw5(3 * s + j), w16 = tw5(4 * s + j);
  const float2 w24 = cmul(w16, w8);
  auto apply = [&](float2 x, float2 w) { return INV ? cmulc(x, w) : cmul(x, w); };
  if constexpr (!INV) dft<R, false>(v);
  // r = 8 r1 + r2: w^{jr} = low(r2) * high(r1), low in {1, w1, .., w7}, high in {1, w8, w16, w24}
  static_for<0, 8>([&](auto R2) {
    constexpr int r2 = decltype(R2)::value;
    float2 lo = make_float2(1.f, 0.f);
    if constexpr (r2 == 1) lo = w1;
    if constexpr (r2 == 2) lo = w2;
    if constexpr (r2 == 3) lo = cmul(w2, w1);
    if constexpr (r2 == 4) lo = w4;
    if constexpr (r2 == 5) lo = cmul(w4, w1);
    if constexpr (r2 == 6) lo = cmul(w4, w2);
    if constexpr (r2 == 7) lo = cmul(cmul(w4, w2), w1);
    if constexpr (r2 > 0) v[r2] = apply(v[r2], lo);
    v[8 + r2] = apply(v[8 + r2], r2 == 0 ? w8 : cmul(lo, w8));
    v[16 + r2] = apply(v[16 + r2], r2 == 0 ? w16 : cmul(lo, w16));
    v[24 + r2] = apply(v[24 + r2], r2 == 0 ? w24 : cmul(lo, w24));
  });
  mid(0);
  if constexpr (INV) dft<R, true>(v);
}

template <bool INV, class Map, class TwLoad>
WWF_HD void pass32_derived(float2* z, int L, int u, TwLoad tw5, Map map = Map()) {
  constexpr int R = 32;
  const int s = L / R;
  const int blk = u / s, j = u - blk * s;
  const int base = blk * L + j;
  float2 v[R];
#pragma unroll
  for (int q = 0; q < R; ++q) v[q] = z[map(base + q * s)];
  pass32_core<INV>(v, s, j, tw5);
#pragma unroll
  for (int q = 0; q < R; ++q) z[map(base + q * s)] = v[q];
}

// ------------------------------------------------------------------------------------------
// real-spectrum pair algebra
// ------------------------------------------------------------------------------------------
// One (k, M-k) pair.  A = Z[k], Bm = Z[M-k], w = w_P^k.
// Forward half:  R2k = 2 R[k],  R2m = 2 R[M-k]  (R = P-point spectrum of the real block).
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
// Both halves with the RIR spectrum in between: (Z[k], Z[M-k]) -> (Zy[k], Zy[M-k]).
WWF_HD void pair_convolve(float2& zk, float2& zm, float2 w, float4 h) {
  float2 R2k, R2m;
  pair_forward(zk, zm, w, R2k, R2m);
  pair_inverse(cmul(R2k, make_float2(h.x, h.y)), cmul(R2m, make_float2(h.z, h.w)), w, zk, zm);
}

WWF_HD int run_of(int l) { return ConvRad::pos(l) >> 4; }   // run holding frequencies l + 1024 d2

// Fused task: runs of l and 1024 - l, 0 < l <= 512.  wl = w_P^l; spec(r) = (H''[k], H''[M-k]) for k = l + 1024 r.
// l = 512 pairs the run with itself (k = 512 + 1024 r <-> 512 + 1024 (15 - r)): u and v are then two copies of the same
// run, every pair is simply computed from both ends, and both copies are written back to the same place - the same
// straight-line code as every other task, so the warp that owns it does not diverge.
// after(r) runs right after pair r (the kernel re-fills its rotating spectrum registers there).
template <class SpecLoad, class After = NoAfter>
WWF_HD void fused_pair_task(float2* z, int l, float2 wl, SpecLoad spec, After after = After()) {
  const int a = run_of(l), ap = run_of(kRuns - l);
  float2* zu = z + 17 * a;    // pad(16 a + q) = 17 a + q
  float2* zv = z + 17 * ap;
  float2 u[16], v[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) { u[q] = zu[q]; v[q] = zv[q]; }
  dft<16, false>(u);          // last forward pass (s = 1: no twiddles) -> u[r] = Z[l + 1024 r]
  dft<16, false>(v);          //                                        -> v[r] = Z[1024 - l + 1024 r]
  static_for<0, 16>([&](auto Rr) {
    constexpr int r = decltype(Rr)::value;
    const float2 w = cmul_cs<false>(wl, TwC<r, 32>::c, TwC<r, 32>::s);   // w_P^{l + 1024 r} = w_P^l w_32^r
    pair_convolve(u[r], v[15 - r], r == 0 ? wl : w, spec(r));
    after(r);
  });
  dft<16, true>(u);           // first inverse pass
  dft<16, true>(v);
#pragma unroll
  for (int q = 0; q < 16; ++q) { zu[q] = u[q]; zv[q] = v[q]; }
}

// The run l = 0: k = 1024 r pairs with 1024 (16 - r); r = 0 carries DC and Nyquist, r = 8 is its own partner.
// spec(i): k = 1024 i, i < 9.
template <class SpecLoad>
WWF_HD void fused_dc_task(float2* z, SpecLoad spec) {
  float2* zu = z + 17 * run_of(0);
  float2 u[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) u[q] = zu[q];
  dft<16, false>(u);
  {
    float2 a = u[0], b = u[0];
    pair_convolve(a, b, make_float2(1.f, 0.f), spec(0));
    u[0] = a;
  }
  static_for<1, 8>([&](auto Rr) {
    constexpr int r = decltype(Rr)::value;
    pair_convolve(u[r], u[16 - r], make_float2(TwC<r, 32>::c, -TwC<r, 32>::s), spec(r));
  });
  {
    float2 a = u[8], b = u[8];
    pair_convolve(a, b, make_float2(0.f, -1.f), spec(8));   // w_32^8 = -i
    u[8] = a;
  }
  dft<16, true>(u);
#pragma unroll
  for (int q = 0; q < 16; ++q) zu[q] = u[q];
}

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------------
// device side
// ------------------------------------------------------------------------------------------
#ifdef WWF_EMUL_HOST                         // tests/emul compiles these headers for the host with nvcc's default arch
#define WWF_GRID_CONSTANT
#else
#define WWF_GRID_CONSTANT __grid_constant__
#endif
// copy the pass tables into shared memory (once per persistent CTA)
__device__ __forceinline__ void conv_load_tables(float2* s_tw, const float2* __restrict__ tw) {
  for (int i = threadIdx.x; i < kConvTwTotal; i += kConvThreads) s_tw[i] = __ldg(tw + i);
}

// the shared-memory passes: forward = radix-32 (L = M), radix-32 (L = 512); inverse = the reverse
template <bool INV>
__device__ __forceinline__ void conv_smem_passes(float2* z, const float2* tw) {   // tw: shared-memory copy
  static_assert(kConvM / 32 == kConvThreads, "one radix-32 task per thread");
  const int u = threadIdx.x;
  const float2* t0 = tw + kConvTw0;
  const float2* t1 = tw + kConvTw1;
  if constexpr (!INV) {
    pass32_derived<false, PadMap>(z, ConvRad::L(0), u, [&](int q) { return t0[q]; });
    __syncthreads();
    pass_task<32, false, PadMap>(z, ConvRad::L(1), u, [&](int q) { return t1[q]; });
    __syncthreads();
  } else {
    pass_task<32, true, PadMap>(z, ConvRad::L(1), u, [&](int q) { return t1[q]; });
    __syncthreads();
    pass32_derived<true, PadMap>(z, ConvRad::L(0), u, [&](int q) { return t0[q]; });
    __syncthreads();
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

// fused middle of one block: thread t owns run pair l(t) (t = 511: the self-paired run l = 512); thread 511 then
// also does the run l = 0 - the only divergent stretch, half a task long (it used to do BOTH self-paired runs on a
// separate path while its warp waited: that warp spent two task times in this phase and everyone else at the barrier)
// The RIR spectrum comes from L2 (131 KB per RIR, one float4 per pair): its latency was the largest stall of the kernel
// (13 % of the samples on the multiplications that consume it).  The first kSpecPf of a thread's 16 entries are therefore
// requested BEFORE the barrier in front of this phase (conv_fused_prefetch), and each register is re-filled with entry
// r + kSpecPf as soon as pair r has used it: every load has a barrier or kSpecPf pairs of arithmetic to land.
#ifndef WWF_SPEC_PF
#define WWF_SPEC_PF 8
#endif
constexpr int kSpecPf = WWF_SPEC_PF;         // spectrum entries in flight per thread (a power of two <= 16)
__device__ __forceinline__ void conv_fused_prefetch(float4 (&h)[kSpecPf], const float4* __restrict__ spec) {
  const float4* sp = spec + threadIdx.x;
#pragma unroll
  for (int r = 0; r < kSpecPf; ++r) h[r] = __ldg(sp + r * 512);
}
// The run l = 0 (DC / Nyquist), lane-parallel: fused_dc_task on ONE lane was a divergent 384-instruction stretch that
// made warp 0 the last one at the barrier behind the middle.  Here the 16 elements sit on 16 lanes (both half-warps
// compute the same thing; only the first stores): a radix-2 DIF FFT by butterfly shuffles leaves frequency
// bitrev4(q) on lane q, the partner 16 - r comes by one more shuffle, every lane does its own pair (r and 16 - r
// both evaluate the pair (min, max) and keep their half of the result), and the mirrored DIT passes bring the run back
// in natural order - about 90 warp-wide instructions, no divergence.  w16[k] = w_16^k and w32[r] = w_32^r are rows
// of the pass-1 table (t1[(16 - 1) 16 + j] = w_512^{16 j}).
__device__ __forceinline__ float2 shfl_xor2(float2 v, int m) {
  return make_float2(__shfl_xor_sync(0xffffffffu, v.x, m, 16), __shfl_xor_sync(0xffffffffu, v.y, m, 16));
}
__device__ __forceinline__ void conv_fused_dc_warp(float2* zc, const float4* __restrict__ spec_special, const float2* t1) {
  const int q = threadIdx.x & 15;
  const float2* w32 = t1 + 15 * 16;                            // w32[r] = w_32^r, r < 16; w_16^k = w32[2 k]
  float2 x = zc[q];                                            // run_of(0) = 0: element q of the run at pad(q) = q
#pragma unroll
  for (int half = 8; half >= 1; half >>= 1) {                  // forward DIF: (a, b) -> (a + b, (a - b) w^j)
    const float2 y = shfl_xor2(x, half);
    const bool lower = (q & half) != 0;
    float2 s2 = lower ? csub(y, x) : cadd(x, y);
    if (half > 1 && lower) s2 = cmul(s2, w32[2 * (q & (half - 1)) * (8 / half)]);
    x = s2;
  }
  {
    const int r = (int)(__brev((unsigned)q) >> 28);            // this lane's frequency / 1024
    const int rp = (16 - r) & 15, qp = (int)(__brev((unsigned)rp) >> 28);
    const float2 y = make_float2(__shfl_sync(0xffffffffu, x.x, qp, 16), __shfl_sync(0xffffffffu, x.y, qp, 16));
    const bool first = r <= 8;                                 // r = 0 and r = 8 pair with themselves (y == x)
    const int rr = first ? r : 16 - r;
    float2 a = first ? x : y, b = first ? y : x;
    pair_convolve(a, b, rr == 0 ? make_float2(1.f, 0.f) : w32[rr], __ldg(spec_special + rr));
    x = first ? a : b;
  }
#pragma unroll
  for (int half = 1; half <= 8; half <<= 1) {                  // inverse DIT: (A, B) -> (A + B w*^j, A - B w*^j)
    const bool lower = (q & half) != 0;
    if (half > 1 && lower) x = cmulc(x, w32[2 * (q & (half - 1)) * (8 / half)]);
    const float2 y = shfl_xor2(x, half);
    x = lower ? csub(y, x) : cadd(x, y);
  }
  if ((threadIdx.x & 16) == 0) zc[q] = x;
}

__device__ __forceinline__ void conv_fused_middle(float2* zc, const float4* __restrict__ spec, float4 (&h)[kSpecPf],
                                                  const uint16_t* __restrict__ fused_l, const float2* __restrict__ fused_tw,
                                                  const float2* t1) {
  const int t = threadIdx.x;
  const float4* sp = spec + t;
  fused_pair_task(zc, (int)__ldg(fused_l + t), __ldg(fused_tw + t), [&](int r) { return h[r & (kSpecPf - 1)]; },
                  [&](int r) { if (r + kSpecPf < 16) h[r & (kSpecPf - 1)] = __ldg(sp + (r + kSpecPf) * 512); });
#ifdef WWF_DC_ONE_LANE
  if (t == kFusedSelfTask) {
    const float4* sd = spec + kSpecSpecial;
    fused_dc_task(zc, [&](int i) { return __ldg(sd + i); });
  }
#else
  if (t < 32) {                                                // warp 0 (warp-uniform): it owns sub-transform 0, whose first run this is
    __syncwarp();                                              // (its lanes' regular tasks do not touch the run, but keep the order explicit)
    conv_fused_dc_warp(zc, spec + kSpecSpecial, t1);
  }
#endif
}

// ---- the passes of conv_kernel in pointer form ------------------------------------------------------------------
// pad(i) = i + (i >> 4) and every stride of the radix-32 passes is a multiple of 16, so element base + q s sits at
// pad(base) + q (s + s / 16): one address per task, compile-time offsets per element (the generic pass_task recomputed
// the pad map per access: a quarter of the kernel's instructions were LEA / LOP3 / IADD3).
constexpr int kConvStride0 = ConvRad::S(0) + ConvRad::S(0) / 16;   // 544: pass 0, elements u + 512 q
constexpr int kConvStride1 = ConvRad::S(1) + ConvRad::S(1) / 16;   // 17:  pass 1, elements 512 blk + j + 16 q

// second radix-32 pass (sub-transforms of length 512, full twiddle table in shared memory), in place; the warp works
// on its own pair of sub-transforms (see WARP-LOCAL MIDDLE above)
template <bool INV>
__device__ __forceinline__ void conv_pass1(float2* z, const float2* t1) {
  const int u = threadIdx.x, w = u >> 5, j = u & 15;
  const int blk = (u & 16) ? conv_sub_b(w) : conv_sub_a(w);
  float2* zp = z + blk * (512 + 32) + j;                         // pad(512 blk + j), j < 16
  float2 v[32];
#pragma unroll
  for (int q = 0; q < 32; ++q) v[q] = zp[kConvStride1 * q];
  if constexpr (!INV) {
    dft<32, false>(v);
#pragma unroll
    for (int r = 1; r < 32; ++r) v[r] = cmul(v[r], t1[(r - 1) * 16 + j]);
  } else {
#pragma unroll
    for (int r = 1; r < 32; ++r) v[r] = cmulc(v[r], t1[(r - 1) * 16 + j]);
    dft<32, true>(v);
  }
#pragma unroll
  for (int q = 0; q < 32; ++q) zp[kConvStride1 * q] = v[q];
}

// Inputs of the FIRST pass straight from global memory: task u combines complex elements u + 512 q, i.e. the sample
// pairs (x[start + 2 (u + 512 q)], +1) - consecutive threads read consecutive 8-byte pairs, fully coalesced - so the
// block never makes the load -> shared memory -> registers round trip.  Samples outside [0, N) are zero (the
// zero padding of the linear convolution / the history before the clip) and are not read at all.
__device__ __forceinline__ void conv_load_pass0(float2 (&v)[32], const float* __restrict__ x, int N, int start, bool al8) {
  const int n0 = start + 2 * (int)threadIdx.x;
  if (start >= 0 && al8 && !(N & 1)) {
    // the common case (CTA-uniform): the block starts inside the clip, even length, aligned rows - every pair is
    // either wholly inside or wholly outside: one predicated 8-byte load per element, no branches
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      const int n = n0 + 2 * ConvRad::S(0) * q;
      float2 t = make_float2(0.f, 0.f);
      if (n < N) t = __ldg(reinterpret_cast<const float2*>(x + n));
      v[q] = t;
    }
    return;
  }
#pragma unroll
  for (int q = 0; q < 32; ++q) {
    const int n = n0 + 2 * ConvRad::S(0) * q;
    float2 t = make_float2(0.f, 0.f);
    if (n >= 0 && n < N) t.x = __ldg(x + n);
    if (n + 1 >= 0 && n + 1 < N) t.y = __ldg(x + n + 1);
    v[q] = t;
  }
}

// The NEXT item's inputs, staged by the thread that will use them: once a thread's last-pass loads of the current
// item have landed, its 32 shared-memory positions are free, and they are exactly the positions its first pass of the
// next item fills - so it has cp.async bring the next item's sample pairs (the same addresses conv_load_pass0 reads)
// into them and goes on with the inverse butterfly, the global stores and the energy: the L2 latency of those loads
// (about a microsecond per item that every warp of the CTA used to sit out at the same time) runs under arithmetic and
// needs no register.  Only for the common case conv_load_pass0 takes without branches.
__device__ __forceinline__ void conv_stage_pass0(float2* zp0, const float* __restrict__ x, int N, int start) {
  const int n0 = start + 2 * (int)threadIdx.x;
#pragma unroll
  for (int q = 0; q < 32; ++q) {
    const int n = n0 + 2 * ConvRad::S(0) * q;
    if (n < N) {
      const unsigned dst = (unsigned)__cvta_generic_to_shared(zp0 + kConvStride0 * q);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(x + n) : "memory");
    }
  }
}
__device__ __forceinline__ void conv_stage_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void conv_load_staged(float2 (&v)[32], const float2* zp0, int N, int start) {
  const int n0 = start + 2 * (int)threadIdx.x;
#pragma unroll
  for (int q = 0; q < 32; ++q) {
    const int n = n0 + 2 * ConvRad::S(0) * q;
    float2 t = make_float2(0.f, 0.f);
    if (n < N) t = zp0[kConvStride0 * q];
    v[q] = t;
  }
}

// Persistent: grid = min(#SMs, work items); work item = (clip b, overlap-save block blk).
// Shared-memory round trips per block: pass 0 (store only: its inputs come from global memory) | pass 1 | fused middle
// | inverse pass 1 | inverse pass 0 (load only: its results go straight to global memory) = 4 stores + 4 loads of the
// block (was 6 + 6 with a staging copy at either end) and 2 CTA barriers - one on either side of the warp-local middle
// (was 9, then 5).
// The flat feature path's per-clip noise-mix records (ClipMix) are produced here too (MIX): in the prologue every warp
// resolves the noise side of one of the CTA's own clips (bank lookups + segment energy: a chain of dependent loads
// that overlaps the table staging) and parks it in SHARED memory; when a clip's block is done, thread 0 turns the
// block's energy into F.add_noise's scale and writes the record.  Nothing is carried in registers through the hot
// loop - it sits at the 128-register limit and its code generation is touchy: variants that kept anything live across
// it, or merely shared one instantiation with the plain kernel, cost 6-12 us per 1024 clips - so the record code lives
// in two __noinline__ helpers and the kernel is a template.  No kernel of its own (feat_prep_kernel: 16 us of mostly
// launch and load latency per step) sits in front of the frames kernel any more.  Measured (B = 1024, final build):
// plain kernel 106.4 us, with the records 110.6 us - the prologue's pointer chase (three dependent global loads per
// clip), which is serial latency wherever it is put (as its own kernel in front of this one: 8 us) - against 16 us.
static __device__ __noinline__ void conv_build_order(const ConvParams& p, uint16_t* s_order, int* s_scan) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (p.B + kConvThreads - 1) / kConvThreads;
  const int i0 = min(p.B, tid * per), i1 = min(p.B, i0 + per);
  int cnt = 0;
  for (int i = i0; i < i1; ++i) cnt += rir_in_range(__ldg(p.rir_idx + i), p.n_rir) ? 1 : 0;
  int incl = cnt;                                                // inclusive prefix sum over the CTA
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) s_scan[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = lane < kConvThreads / 32 ? s_scan[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    if (lane < kConvThreads / 32) s_scan[lane] = v;             // inclusive warp totals
  }
  __syncthreads();
  const int n_rev = s_scan[kConvThreads / 32 - 1];
  int pr = incl - cnt + (warp > 0 ? s_scan[warp - 1] : 0);       // reverberated clips before i0
  int pd = n_rev + (i0 - pr);                                    // dry clips go behind all reverberated ones
  for (int i = i0; i < i1; ++i) {
    if (rir_in_range(__ldg(p.rir_idx + i), p.n_rir)) s_order[pr++] = (uint16_t)i;
    else s_order[pd++] = (uint16_t)i;
  }
  __syncthreads();
}
// clip of work item `item` (single-block clips: item = position in the order)
template <bool ORD>
__device__ __forceinline__ int conv_item_clip(const uint16_t* s_order, int ci) {
  if constexpr (ORD) return (int)s_order[ci];
  else return ci;
}

template <bool ORD>
static __device__ __noinline__ void conv_mix_prologue(const ConvParams& p, ClipMix* s_mix, const uint16_t* s_order) {
  const int warp = threadIdx.x >> 5;
  for (int k = warp; k < kConvMaxOwn; k += kConvThreads / 32) {
    const int ci = blockIdx.x + k * gridDim.x;                   // (single-block clips: item = position in the order)
    if (ci >= p.B) break;
    const int b = conv_item_clip<ORD>(s_order, ci);
    const ClipNoise cn = resolve_noise(p.noise, p.noise_idx, p.noise_off, b);   // warp-uniform
    ClipMix m{0.f, 0, 0, 1, 0, 0.f, 0.f};
    if (cn.nz != nullptr) {
      m.en = warp_noise_energy(cn, p.N);
      m.snr = p.snr_db ? __ldg(p.snr_db + b) : 0.f;
      m.has_noise = 1; m.noff = cn.off; m.nlen = cn.len; m.nz_off = (long long)(cn.nz - p.noise.data);
    }
    if ((threadIdx.x & 31) == 0) s_mix[k] = m;
  }
}
// record of clip b, the CTA's k-th item, once its energy es is known; one thread
static __device__ __noinline__ void conv_mix_finish(const ConvParams& p, const ClipMix* s_mix, int b, int k, float es) {
  ClipMix m = s_mix[k];
  if (m.has_noise) m.scale = snr_scale(es, m.en, m.snr);
  p.mix_g[b] = m;
}

// The LAST warp (warp 0 already carries the extra DC task of the fused phase), behind a CTA barrier: total energy of
// the item parked in s_pend / s_part (if any) -> es_part, mix record
constexpr int kConvFlushThread = kConvThreads - 32;
template <bool MIX>
__device__ __forceinline__ void conv_flush_energy(const ConvParams& p, const ClipMix* s_mix, const float (*s_part)[kConvThreads / 32],
                                                  int* s_pend) {
  const int b = s_pend[0];
  if (b < 0) return;                                             // warp-uniform
  const int blk = s_pend[1], k = s_pend[3], lane = threadIdx.x & 31;   // k: the CTA's item ordinal
  float e = lane < kConvThreads / 32 ? s_part[s_pend[2]][lane] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
  if (lane == 0) {
    s_pend[0] = -1;
    if (p.es_part != nullptr) p.es_part[(size_t)b * p.es_nb + blk] = e;
    if constexpr (MIX) conv_mix_finish(p, s_mix, b, k, e);      // F.add_noise's scale, now that the clip's energy is known
  }
}

template <bool MIX, bool ORD>
__global__ void __launch_bounds__(kConvThreads, 1) conv_kernel(const WWF_GRID_CONSTANT ConvParams p) {   // (grid constant:
  // the __noinline__ helpers take p by reference; without it the kernel keeps a 200-byte local-memory copy of the
  // parameters and reads N, the strides, ... through LDL inside the hot loop)
  extern __shared__ __align__(16) float2 zc[];
  __shared__ float red[32];
  __shared__ float s_part[2][kConvThreads / 32];                 // per-warp energy partials of the last two items
  __shared__ int s_pend[4];                                      // {clip, block, s_part row, item ordinal} of the item not yet summed
  __shared__ int s_scan[kConvThreads / 32];
  __shared__ ClipMix s_mix[MIX ? kConvMaxOwn : 1];
  if (threadIdx.x == kConvFlushThread) s_pend[0] = -1;
  int par = 0;
  int staged = -1;                                               // item whose first-pass inputs wait in shared memory
  float2* s_tw = zc + kConvSmemElems;
  uint16_t* s_order = reinterpret_cast<uint16_t*>(s_tw + kConvTwTotal);   // [B] when p.ordered
  conv_load_tables(s_tw, p.tw);
  const int u = threadIdx.x;
  const float2* t0 = s_tw + kConvTw0;
  const float2* t1 = s_tw + kConvTw1;
  float2* zp0 = zc + u + (u >> 4);                               // pad(u): pass-0 elements at zp0[544 q]
  const int nblk = p.es_nb;
  pdl_wait();                                                    // the clips may come from a kernel of ours (gather / stretch);
                                                                 // the previous call's epilogue has read its clip maxima
  if (p.clip_max != nullptr)                                     // identity of the running maxima (wwf_feat.cuh: kMaxKeyMemset)
    for (int i = blockIdx.x * kConvThreads + threadIdx.x; i < p.n_clip_max; i += gridDim.x * kConvThreads)
      p.clip_max[i] = (int)0x80808080;
  if constexpr (ORD) conv_build_order(p, s_order, s_scan);
  if constexpr (MIX) conv_mix_prologue<ORD>(p, s_mix, s_order);
  __syncthreads();                                               // twiddle tables and mix records visible (a dry first
                                                                 // item reads its record straight away)
  for (int item = blockIdx.x; item < p.B * nblk; item += gridDim.x) {
    const int ci = item / nblk, blk = item - ci * nblk;
    const int b = conv_item_clip<ORD>(s_order, ci);
    const int r = __ldg(p.rir_idx + b);
    const float* x = p.wav + (size_t)b * p.wav_stride;
    if (!rir_in_range(r, p.n_rir)) {                             // dry clip (CTA-uniform): only its mix record
      if constexpr (MIX) {
        const int k = (item - (int)blockIdx.x) / (int)gridDim.x;
        const float es = s_mix[k].has_noise ? block_energy(x, p.N, red) : 0.f;
        if (threadIdx.x == 0) conv_mix_finish(p, s_mix, b, k, es);
      }
      continue;
    }
    {
      float2 v[32];
      conv_stage_wait();                                         // (nothing pending unless the last item staged one)
      if (staged == item) conv_load_staged(v, zp0, p.N, blk * p.valid - p.hist);
      else conv_load_pass0(v, x, p.N, blk * p.valid - p.hist, (reinterpret_cast<uintptr_t>(x) & 7) == 0);
      pass32_core<false>(v, ConvRad::S(0), u, [&](int q) { return t0[q]; });
      // (no barrier needed here: these are the 32 positions this very thread loaded in the previous item's last pass)
#pragma unroll
      for (int q = 0; q < 32; ++q) zp0[kConvStride0 * q] = v[q];
    }
    __syncthreads();
    if (threadIdx.x >= kConvFlushThread) conv_flush_energy<MIX>(p, s_mix, s_part, s_pend);   // the previous item's energy
    conv_pass1<false>(zc, t1);
    {
      float4 h[kSpecPf];
      conv_fused_prefetch(h, p.spec + (size_t)r * kSpecPerRir);
      __syncwarp();                                              // the warp's own sub-transform pair: no CTA barrier
      conv_fused_middle(zc, p.spec + (size_t)r * kSpecPerRir, h, p.fused_l, p.fused_tw, t1);
    }
    __syncwarp();
    {  // pull the next work item's samples into L2 while this block's inverse passes run (no registers held)
      const int nitem = item + gridDim.x;
      if (nitem < p.B * nblk) {
        const int nci = nitem / nblk, nblk_ = nitem - nci * nblk, nb_ = conv_item_clip<ORD>(s_order, nci);
        const float* nx = p.wav + (size_t)nb_ * p.wav_stride;
        const int nstart = nblk_ * p.valid - p.hist;
        for (int q = threadIdx.x * 32; q < kConvP; q += kConvThreads * 32) {   // one 128-byte line per request
          const int n = nstart + q;
          if (n >= 0 && n < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + n));
        }
      }
    }
    conv_pass1<true>(zc, t1);
    __syncthreads();
    // last inverse pass: results leave the registers for global memory.  Element u + 512 q = block samples
    // i = 2 (u + 512 q), i + 1 -> clip sample blk * valid + i - hist (hist, valid are multiples of 4: n is even and
    // the row base is 16-byte aligned, so the pair is one 8-byte store).  The energy of the stored samples is
    // accumulated on the way for the SNR mix that follows (fixed order: deterministic).
    float e0 = 0.f, e1 = 0.f;
    {
      float2 v[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) v[q] = zp0[kConvStride0 * q];
      pass32_core<true>(v, ConvRad::S(0), u, [&](int q) { return t0[q]; }, [&](int) {
        // every loaded value is in use: the positions are free for the next item's inputs (a dry next item simply
        // leaves them unused; whatever lands is overwritten by the next first pass after its conv_stage_wait())
        // (computed here rather than parked in shared memory by the prefetch block above: that variant spilled)
        const int nitem = item + gridDim.x;
        if (nitem < p.B * nblk) {
          const int nci = nitem / nblk, nstart = (nitem - nci * nblk) * p.valid - p.hist;
          const float* nx = p.wav + (size_t)conv_item_clip<ORD>(s_order, nci) * p.wav_stride;
          if (nstart >= 0 && !(p.N & 1) && (reinterpret_cast<uintptr_t>(nx) & 7) == 0) {
            conv_stage_pass0(zp0, nx, p.N, nstart);
            staged = nitem;
          }
        }
      });
      float* y = p.rev + (size_t)b * p.rev_stride;
      const int nbase = blk * p.valid - p.hist + 2 * u;
      if (p.hist == 0 && !(p.N & 1) && (reinterpret_cast<uintptr_t>(y) & 7) == 0) {
        // single-block clips of even length (CTA-uniform): predicated 8-byte stores, no branches
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          // store and energy under ONE predicate, the values read in place: selecting zeros into copies first cost
          // two moves per element - each a write-after-read wait on a register the store before it still had to
          // read (5 % of the kernel's stall samples) - and 3 M warp-instructions per launch
          if (nbase + 2 * ConvRad::S(0) * q < p.N) {
            *reinterpret_cast<float2*>(y + nbase + 2 * ConvRad::S(0) * q) = v[q];
            if (q & 1) e1 = fmaf(v[q].x, v[q].x, fmaf(v[q].y, v[q].y, e1));
            else e0 = fmaf(v[q].x, v[q].x, fmaf(v[q].y, v[q].y, e0));
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const int i = 2 * (u + ConvRad::S(0) * q), n = nbase + 2 * ConvRad::S(0) * q;
          const bool in0 = i >= p.hist && n < p.N, in1 = i >= p.hist && n + 1 < p.N;
          if (in0) y[n] = v[q].x;
          if (in1) y[n + 1] = v[q].y;
          const float ax = in0 ? v[q].x : 0.f, ay = in1 ? v[q].y : 0.f;
          if (q & 1) e1 = fmaf(ax, ax, fmaf(ay, ay, e1));
          else e0 = fmaf(ax, ax, fmaf(ay, ay, e0));
        }
      }
    }
    // energy of this block's output samples (fixed reduction order): the warps park their partial sums; warp 0 adds
    // them up behind the NEXT barrier every thread passes anyway (conv_flush_energy), so that no barrier ends the item:
    // a thread's last-pass loads and its first-pass stores of the next item touch the same 32 positions, and a warp
    // that is done goes straight on to the next item's global loads while the others still finish this one
    float e = e0 + e1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0) s_part[par][threadIdx.x >> 5] = e;
    if (threadIdx.x == kConvFlushThread) { s_pend[1] = blk; s_pend[2] = par; s_pend[3] = (item - (int)blockIdx.x) / (int)gridDim.x; s_pend[0] = b; }
    par ^= 1;
  }
  __syncthreads();
  if (threadIdx.x >= kConvFlushThread) conv_flush_energy<MIX>(p, s_mix, s_part, s_pend);
}

// Spectrum of one zero-padded RIR in the layout the fused task consumes (registration time).
// Forward FFT including the last radix-16 pass in shared memory, then every (k, M-k) pair.
struct SpecParams {
  const float* data; const int64_t* offsets; int n_rir;
  float4* spec; const float2* tw; const uint16_t* fused_l;
};

__device__ __forceinline__ float2 w_exact(int k) {   // w_P^k, double-precision sincospi
  double s, c;
  sincospi(-2.0 * (double)k / (double)kConvP, &s, &c);
  return make_float2((float)c, (float)s);
}

__global__ void __launch_bounds__(kConvThreads, 1) rir_spectrum_kernel(const SpecParams p) {
  extern __shared__ __align__(16) float2 zc[];
  float2* s_tw = zc + kConvSmemElems;
  conv_load_tables(s_tw, p.tw);
  const int r = blockIdx.x;
  const int64_t o0 = p.offsets[r], o1 = p.offsets[r + 1];
  conv_load_block(zc, p.data + o0, (int)(o1 - o0), 0, false);
  __syncthreads();
  conv_smem_passes<false>(zc, s_tw);
  for (int u = threadIdx.x; u < kRuns; u += kConvThreads)
    pass_task<16, false, PadMap>(zc, ConvRad::L(2), u, [&](int) { return make_float2(1.f, 0.f); });
  __syncthreads();
  PadMap pad;
  const float sc = 1.0f / (8.0f * (float)kConvM);   // R2 = 2R, H'' = R / (4M)
  float4* spec = p.spec + (size_t)r * kSpecPerRir;
  auto entry = [&](int k) {
    float2 R2k, R2m;
    pair_forward(zc[pad(ConvRad::pos(k))], zc[pad(ConvRad::pos((kConvM - k) & (kConvM - 1)))], w_exact(k), R2k, R2m);
    if (k == 0) { R2k.y = 0.f; R2m.y = 0.f; }     // DC and Nyquist of a real signal are real
    return make_float4(R2k.x * sc, R2k.y * sc, R2m.x * sc, R2m.y * sc);
  };
  for (int i = threadIdx.x; i < 16 * kFusedTasks; i += kConvThreads) {
    const int rr = i / kFusedTasks, t = i - rr * kFusedTasks;
    spec[rr * 512 + t] = entry((int)p.fused_l[t] + 1024 * rr);
  }
  if (threadIdx.x < 9) spec[kSpecSpecial + threadIdx.x] = entry(1024 * (int)threadIdx.x);
}
#endif  // __CUDACC__

}  // namespace wwf
