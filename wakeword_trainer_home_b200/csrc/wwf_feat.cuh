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
#include "wwf_mix.cuh"

namespace wwf {

constexpr int kMaxMasks = 8;
constexpr int kMaxMelRounds = 16;              // rounds of the mel lane schedule (128 filters need 4 - 5)

// Per-n_fft plan: radix list, G = complex FFTs (frame pairs) per warp iteration, the launch bounds
// (max threads per CTA, min CTAs per SM) the kernel is compiled for, and the index map of the warp's
// FFT scratch: the power-of-two sizes need one pad element per 16 (their later passes would
// otherwise be 4- to 16-way bank conflicted); 400 = 16 x 25 has an odd stride and needs none.
#ifndef WWF_FLAT400
#define WWF_FLAT400 384
#endif
template <int NFFT> struct StftPlan;
// kFlatThreads: CTA width the flat frames kernel is compiled for (same 2 CTAs per SM): n_fft 400 fits 80 registers
// there (24 resident warps instead of 20: 128.9 -> 125.5 us on the bench workload; 64 registers / 32 warps spill and lose).
template <> struct StftPlan<256>  { using Rad = Radices<16, 16>;    using Map = PadMap2;     static constexpr int G = 2, kThreads = 320, kMinCtas = 2, kFlatThreads = 320; };
template <> struct StftPlan<400>  { using Rad = Radices<16, 25>;    using Map = IdentityMap; static constexpr int G = 2, kThreads = 320, kMinCtas = 2, kFlatThreads = WWF_FLAT400; };
template <> struct StftPlan<512>  { using Rad = Radices<16, 8, 4>;  using Map = PadMap2;     static constexpr int G = 1, kThreads = 512, kMinCtas = 1, kFlatThreads = 512; };
template <> struct StftPlan<1024> { using Rad = Radices<16, 16, 4>; using Map = PadMap2;     static constexpr int G = 1, kThreads = 512, kMinCtas = 1, kFlatThreads = 512; };
template <> struct StftPlan<2048> { using Rad = Radices<16, 16, 8>; using Map = PadMap2;     static constexpr int G = 1, kThreads = 512, kMinCtas = 1, kFlatThreads = 512; };
// complex elements of one FFT's scratch buffer (mapped length, rounded up to an even count)
template <int NFFT> constexpr int stft_zlen() {
  return (typename StftPlan<NFFT>::Map()(NFFT - 1) + 2) & ~1;
}
// Two-pass plans whose last pass fits one warp round for all G transforms (400 = 16 x 25: 2 x 16 tasks; 256 = 16 x 16)
// leave the spectrum in natural order (pass_task_natural); the others in digit-reversed order (Radices::pos).
template <int NFFT> constexpr bool stft_natural_out() {
  using P = StftPlan<NFFT>;
  return P::Rad::npass == 2 && P::G * (NFFT / P::Rad::R(1)) <= 32;
}

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
  //   | mel_w (lane-interleaved) | dct [n_mels][c8] | mel tasks int2[mel_rounds*32] | rowmask u8[n_feat]
  //   | colmask u8[T] | z float2 [nwarps][G][ZL]
  int off_res, off_window, off_tw, off_melw, off_dct, off_meltasks, off_rowmask, off_colmask, off_z;
  int n_melw, mel_rounds, c8;                    // interleaved mel weight count, schedule rounds; n_mfcc rounded up to 8
  unsigned short mel_pairs[kMaxMelRounds];        // two-tap iterations of each round (warp-uniform trip counts)
  // device constants (plan-owned)
  const float* window;                           // [NFFT]
  const float2* tw;                              // concatenated per-pass twiddle tables
  const int2* mel_tasks;                         // [mel_rounds*32] lane schedule of the sparse mel rows (wwf_tables.h)
  const float* mel_w;                            // filter weights, interleaved per round: [(base + tap) * 32 + lane]
  const float* dct;                              // [n_mels][n_mfcc]
  const uint4* dct_frag;                         // the same matrix as split-TF32 mma.sync B fragments (build_dct_fragments)
  const float* dct_colsum;                       // [c8] column sums of dct (float of the double sum), 0 beyond n_mfcc
  // augmentation draws (device, nullable)
  const int32_t* rir_idx; int n_rir;             // a clip is reverberated iff rir_in_range(rir_idx[b], n_rir)
  const int32_t* noise_idx; const int64_t* noise_off; const float* snr_db;
  NoiseBankDev noise;
  const float* es_part; int es_nb;               // per-clip energy partials written by conv_kernel [B][es_nb]
  const int32_t* fs; const int32_t* fl; const int32_t* ts; const int32_t* tl; int nF, nT;
  // output
  void* out; int64_t out_stride;
  int* nonfinite_flag;                           // plan-owned device int, set to 1 if any feature is NaN/Inf
  // ---- flat path (few clips / large batches, see feat_frames_kernel): per-clip intermediates in the caller's workspace ----
  float* tile_g;                                 // dB values [B][T][mp], frame-major (= one flat matrix of B*T rows)
  int mp;                                        // n_mels rounded up to 4
  int* clip_max;                                 // [B] running maximum of the clip's dB values (ordered-int key)
  ClipMix* mix_g;                                // [B] noise-mix records (nullptr: no noise in this call)
  int ngroups;                                   // frame groups per clip
  // feat_frames_kernel shared memory: window [NFFT] | tw | mel_w | mel tasks | z float2 [nwarps][G][ZL]
  int f_off_tw, f_off_melw, f_off_meltasks, f_off_z;
  int eb_frames, eb_pitch;                       // feat_epilogue_block_kernel: frames per block, odd smem row pitch
  int em_k8, em_ap;                              // feat_epilogue_mma_kernel: n_mels rounded up to 8, smem pitch of A
};

// float <-> int key whose signed order equals the float order (for atomicMax on the clip maximum); every NaN maps to
// the largest key, so a NaN dB value makes the clip maximum NaN (torch.amax semantics)
__device__ __forceinline__ int float_key(float f) {
  const int i = __float_as_int(f);
  return f != f ? 0x7fffffff : i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }
// Identity of the running maximum: the maxima are initialised by a byte-wise memset; 0x80808080 orders below the key
// of every float except -inf / NaN payloads and reads back as -inf.
constexpr int kMaxKeyMemset = (int)0x80808080;
__device__ __forceinline__ float clip_max_value(int k) { return k == kMaxKeyMemset ? -INFINITY : key_float(k); }

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

// 10*log10(max(x, 1e-10)) = (10/log2(10)) * log2(.) through MUFU.LG2 (lg2.approx: max abs error
// 2^-22.6 on log2 => < 5e-7 dB; the argument is >= 1e-10, never denormal).
__device__ __forceinline__ float power_to_db(float x) {
  // the clamp value is exact like the oracle's; NaN takes the log branch and stays NaN (torch.clamp keeps NaN).
  // lg2.approx.ftz: the argument is > 1e-10, so flushing denormals changes nothing - but the non-ftz form drags a
  // scale-by-2^24 / subtract-24 sequence for denormal inputs along with every call (4 of ~9 instructions per value)
  float l;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
  return x <= 1e-10f ? -100.0f : 3.01029995663981195f * l;
}

// max that PROPAGATES NaN like torch.max / torch.amax (fmaxf returns the other operand): one FMNMX.NAN.
// A clip with a NaN sample gets a NaN maximum, hence a NaN top_db floor, hence NaN features - exactly what
// AmplitudeToDB does (TA/functional/functional.py:393-402) - and the non-finite flag is raised.
__device__ __forceinline__ float fmax_nan(float a, float b) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 800)
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
#else
  return (a != a || b != b) ? NAN : fmaxf(a, b);
#endif
}
__device__ __forceinline__ float warp_max_nan(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax_nan(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename OutT> __device__ __forceinline__ OutT to_out(float v);
template <> __device__ __forceinline__ float to_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __half to_out<__half>(float v) { return __float2half_rn(v); }

// SpecAugment flag of row / column q of clip b: inside any of the clip's nm explicit (start, length) masks
__device__ __forceinline__ bool in_masks(const int32_t* st, const int32_t* ln, int nm, int b, int q) {
  bool mk = false;
  if (st != nullptr)
    for (int j = 0; j < nm; ++j) {
      const int s0 = __ldg(st + (size_t)b * nm + j), l = __ldg(ln + (size_t)b * nm + j);
      mk |= (q >= s0) && (q < s0 + l);
    }
  return mk;
}

// ==========================================================================================================
// One frame group = 2G consecutive frames of one clip, processed by ONE warp in its private scratch z:
//   load + [noise mix] + window -> FFT (two real frames per complex transform) -> power spectra -> sparse mel -> dB.
// The single body behind both launch shapes (feat_kernel and feat_frames_kernel), so their results are bit-identical.
// sink(m, i, d): dB value d of mel filter m for frame f0 + i (called for frames < T only).  Returns the NaN-propagating
// maximum of the values this lane produced.
// ==========================================================================================================
struct MelTables {          // shared-memory copies
  const float* window;      // [NFFT]
  const float2* tw;         // pass twiddles
  const float* melw;        // lane-interleaved filter weights
  const int2* tasks;        // [rounds*32] mel lane schedule
  int rounds;
  const unsigned short* pairs;   // [rounds] two-tap iterations per round (kernel parameter space: uniform loads)
};

template <int NFFT, int HOP32, class Sink>
__device__ __forceinline__ float frame_group_to_db(const float* __restrict__ x, int N, int T, int hop,
                                                   const float* __restrict__ nz, int noff, int nlen, float scale,
                                                   int f0, float2* z, const MelTables& tb, Sink sink) {
  using Plan = StftPlan<NFFT>;
  using Rad = typename Plan::Rad;
  using Map = typename Plan::Map;
  constexpr int G = Plan::G;
  constexpr int K = NFFT / 2 + 1;
  constexpr int NC = (NFFT + 31) / 32;                       // 32-sample columns per frame
  constexpr int ZL = stft_zlen<NFFT>();                      // scratch elements per FFT (with padding)
  static_assert(Rad::n == NFFT, "radix plan");
  const Map zmap;
  const int lane = threadIdx.x & 31;
  const bool mix = nz != nullptr;
  const float* s_window = tb.window;

  // 1. load + window: z[g][j] = w[j] * (frame(f0+2g)[j] + i frame(f0+2g+1)[j])
  // (kRegFirst: register-staged groups of the 512 / 1024 plans also run the first radix-16 pass here, in registers)
  constexpr bool kRegFirst = HOP32 > 0 && G == 1 && Rad::R(0) == 16 && Rad::S(0) % 32 == 0 && NC <= 32 && NFFT % 32 == 0;
  bool first_done = false;
  bool staged = false;
  if constexpr (HOP32 > 0) {
    constexpr int NR = (2 * G - 1) * HOP32 + NC;             // registers holding the group's sample span
    const int s0 = f0 * hop - NFFT / 2;
    int q0 = 0;                                              // first noise sample of the span (mix only)
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
        if constexpr (kRegFirst) {
          // First pass in registers: its radix-16 tasks combine elements j + S0 q, and S0 is a whole number of
          // 32-sample columns (512: one, 1024: two), so all 16 inputs of task j = 32 c0 + lane are THIS lane's staged
          // columns c0 + CPT q - the windowed samples go through the butterfly and the twiddles before they are stored
          // for the first time: the pass costs no shared-memory loads and the window store is its output store.
          constexpr int S0 = Rad::S(0), CPT = S0 / 32;
          const float2* tw0 = tb.tw + Rad::tw_off(0);
#pragma unroll
          for (int c0 = 0; c0 < CPT; ++c0) {
            const int jt = 32 * c0 + lane;
            float2 v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const int c = c0 + CPT * q;
              const float w = s_window[32 * c + lane];
              v[q] = make_float2(w * sreg[c], w * sreg[HOP32 + c]);
            }
            dft<16, false>(v);
            twiddle16(v, S0, jt, [&](int q) { return tw0[q]; });
#pragma unroll
            for (int r = 0; r < 16; ++r) z[zmap(jt + S0 * r)] = v[r];
          }
          first_done = true;
        } else {
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
  }
  if (!staged) {
    // boundary groups (reflect padding, frames >= T, wrapping noise) and hops that are not a
    // multiple of 32: plain per-element gather, deliberately not unrolled (cold code)
#pragma unroll 1
    for (int idx = lane; idx < G * NFFT; idx += 32) {
      const int g = idx / NFFT, j = idx - g * NFFT;
      const int ta = f0 + 2 * g, tb_ = ta + 1;
      const float w = s_window[j];
      float re = 0.f, im = 0.f;
      if (ta < T) {
        const int i = reflect_index(ta * hop - NFFT / 2 + j, N);
        re = __ldg(x + i);
        if (mix) re = fmaf(scale, noise_at(nz, noff, nlen, i), re);
      }
      if (tb_ < T) {
        const int i = reflect_index(tb_ * hop - NFFT / 2 + j, N);
        im = __ldg(x + i);
        if (mix) im = fmaf(scale, noise_at(nz, noff, nlen, i), im);
      }
      z[g * ZL + zmap(j)] = make_float2(re * w, im * w);
    }
  }
  __syncwarp();
  // 2. forward FFT passes (in place; digit-reversed result, or natural order when the last pass is one warp round)
  constexpr bool kNatural = stft_natural_out<NFFT>();
  // three-pass plans that end with radix 4 (512 = 16.8.4, 1024 = 16.16.4; one transform per warp): the last pass is
  // fused with the spectrum split below (pair units), so the generic loop stops one pass early
  constexpr bool kPairSplit = !kNatural && G == 1 && Rad::npass == 3 && Rad::R(2) == 4 && (NFFT / 8) % 32 == 0;
  static_for<0, Rad::npass - (kPairSplit ? 1 : 0)>([&](auto I) {
    constexpr int i = decltype(I)::value;
    constexpr int R = Rad::R(i), L = Rad::L(i), tasks = NFFT / R;
    const float2* tw = tb.tw + Rad::tw_off(i);
    if constexpr (kRegFirst && i == 0) {
      if (first_done) return;                                  // (warp-uniform) done in registers while loading
    }
    if constexpr (kNatural && i == Rad::npass - 1) {
      // Last pass of a two-pass plan (R0 = 16 tasks per transform, G = 2: lane = 16 g + u) fused with the spectrum
      // split: task u leaves frequencies k = u + 16 r, r < R, in its REGISTERS; the partner Z[n - k] of the split is
      // frequency (16 - u) + 16 (R - 1 - r): lane 16 - u of the same half-warp, register R - 1 - r - a compile-time
      // register index, so one shuffle pair per bin fetches it (u = 0 pairs with itself: 16 (R - r)).  The spectrum
      // never goes back to shared memory: 26 shuffles replace 25 stores + 26 loads of 8 bytes per lane (-13 % of the
      // kernel's shared-memory wavefronts), and the powers land in plain bin order for the mel stage.
      static_assert(!kNatural || (tasks == 16 && G == 2 && Rad::R(0) == 16), "natural-order split: 2 x 16 tasks");
      const int g = lane >> 4, u = lane & 15;
      float2* zz = z + g * ZL;
      float2 v[R];
#pragma unroll
      for (int q = 0; q < R; ++q) v[q] = zz[zmap(u * R + q)];
      dft<R, false>(v);
      __syncwarp();                                            // every lane has read its inputs: zz may be overwritten
      const int src = (lane & 16) | ((16 - u) & 15);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (16 * r <= NFFT / 2) {                              // (compile-time) some lane of the half-warp has k <= n/2
          float2 pz;
          pz.x = __shfl_sync(0xffffffffu, v[R - 1 - r].x, src);
          pz.y = __shfl_sync(0xffffffffu, v[R - 1 - r].y, src);
          if (u == 0) pz = v[r == 0 ? 0 : R - r];
          const int k = u + 16 * r;
          if (k <= NFFT / 2) zz[k] = pair_split_power(v[r], pz);     // plain bin order: what the mel lane schedule is built for
        }
      }
    } else {
      // tasks of the G FFTs share the rounds of 32 lanes, unless one FFT per round costs no extra round (n_fft 400:
      // 25 radix-16 tasks per FFT, 2 rounds either way) - then no half-warp straddles two FFTs (no bank conflicts
      // between their columns or twiddles) and the task index needs no division
      constexpr bool kPerFft = tasks < 32 && G * ((tasks + 31) / 32) == (G * tasks + 31) / 32;
      if constexpr (kPerFft) {
#pragma unroll 1   // one copy of each radix butterfly: the hot loop has to stay inside the instruction cache
        for (int g = 0; g < G; ++g)
          if (lane < tasks) pass_task<R, false, Map>(z + g * ZL, L, lane, [&](int q) { return tw[q]; });
      } else {
#pragma unroll 1
        for (int u = lane; u < G * tasks; u += 32) {
          const int g = u / tasks, uu = u - g * tasks;
          pass_task<R, false, Map>(z + g * ZL, L, uu, [&](int q) { return tw[q]; });
        }
      }
    }
    __syncwarp();
  });
  // 3. split the packed pair into two power spectra (|A[k]|^2, |B[k]|^2): read every (Z[k], Z[n-k])
  //    first, then store the powers in plain bin order.  Half-warps take 16 consecutive bins of ONE transform
  //    (slot = 16-bin block; K rounded up to K16 blocks per transform), so a half-warp's two gathers never straddle
  //    two transforms.  (The natural-order plans have done this inside their last pass.)
  if constexpr (kPairSplit) {
    // Last pass (radix 4 on 4 contiguous elements, no twiddles) + split in registers.  Last-pass task t (= k mod Q,
    // Q = n / 4) produces the bins t + Q r; their split partners n - k = (Q - t) + Q (3 - r) come from task Q - t.  One
    // lane therefore takes the PAIR of tasks (t, Q - t), t = 1 .. Q/2 - 1, and has everything four bins need in its own
    // registers - no digit-reversed gather, no pos() arithmetic per bin, the spectrum is not stored again: 8 loads +
    // 4 stores per unit instead of 8 + 8 (pass) + 8 + 4 (split).  Unit 0 takes the two self-paired tasks 0 and Q/2
    // (bins 0, Q, 2Q = n/2 and Q/2, 3Q/2).  Every unit's results wait in registers until all units of the warp have
    // read their inputs (the bins overwrite the transform in place).
    constexpr int Q = NFFT / 4, NU = Q / 2, ROUNDS = NU / 32;
    float2 pw[ROUNDS][4];
    float2 pw0_extra = make_float2(0.f, 0.f);                  // unit 0 has a fifth bin
#pragma unroll
    for (int rd = 0; rd < ROUNDS; ++rd) {
      const int w = lane + 32 * rd;
      const int ta = w, tb_ = w == 0 ? Q / 2 : Q - w;
      const float2* pa = z + zmap(Rad::pos(ta));               // positions pos(t) + q, q < 4: one group of 16, no pad inside
      const float2* pb = z + zmap(Rad::pos(tb_));
      float2 va[4], vb[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { va[q] = pa[q]; vb[q] = pb[q]; }
      dft<4, false>(va);
      dft<4, false>(vb);
      if (w != 0) {
        pw[rd][0] = pair_split_power(va[0], vb[3]);            // k = t
        pw[rd][1] = pair_split_power(va[1], vb[2]);            // k = t + Q
        pw[rd][2] = pair_split_power(vb[0], va[3]);            // k = Q - t
        pw[rd][3] = pair_split_power(vb[1], va[2]);            // k = 2Q - t
      } else {
        pw[rd][0] = pair_split_power(va[0], va[0]);            // k = 0 (DC)
        pw[rd][1] = pair_split_power(va[1], va[3]);            // k = Q
        pw0_extra = pair_split_power(va[2], va[2]);            // k = 2Q = n / 2 (Nyquist)
        pw[rd][2] = pair_split_power(vb[0], vb[3]);            // k = Q / 2
        pw[rd][3] = pair_split_power(vb[1], vb[2]);            // k = 3Q / 2
      }
    }
    __syncwarp();
#pragma unroll
    for (int rd = 0; rd < ROUNDS; ++rd) {
      const int w = lane + 32 * rd;
      const int ka = w, kb = w == 0 ? Q / 2 : Q - w;
      z[ka] = pw[rd][0];
      z[ka + Q] = pw[rd][1];
      z[kb] = pw[rd][2];
      z[kb + Q] = pw[rd][3];
      if (w == 0) z[2 * Q] = pw0_extra;
    }
  } else if constexpr (!kNatural) {
    constexpr int K16 = (K + 15) / 16;                       // 16-bin blocks per transform
    constexpr int NS = (G * K16 + 1) / 2;                    // split items per lane
    const int hw = lane >> 4, l16 = lane & 15;
    float2 pw[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      const int slot = 2 * i + hw, g = slot / K16, k = 16 * (slot - g * K16) + l16;
      if (g < G && k < K) {
        const float2* zz = z + g * ZL;
        const int kn = k == 0 ? 0 : NFFT - k;
        pw[i] = pair_split_power(zz[zmap(Rad::pos(k))], zz[zmap(Rad::pos(kn))]);
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      const int slot = 2 * i + hw, g = slot / K16, k = 16 * (slot - g * K16) + l16;
      if (g < G && k < K) z[g * ZL + k] = pw[i];           // plain bin order: what the mel lane schedule is built for
    }
  }
  __syncwarp();
  // 4. sparse mel rows + dB through the lane schedule (build_mel_schedule, wwf_tables.h): per round one task per
  //    lane = a run of consecutive bins of one filter, all 2G frames of the group at once
  float vmax = -INFINITY;
  for (int r = 0; r < tb.rounds; ++r) {
    const int2 task = tb.tasks[r * 32 + lane];
    const int k0 = task.x & 0xffff;
    const unsigned flags = (unsigned)task.y >> 24;
    const float* wr = tb.melw + (task.y & 0xffff) * 32 + lane;
    // two independent accumulator chains per frame pair (even / odd taps) hide the LDS + FFMA2 latency.  Every lane
    // runs the round's full (warp-uniform) trip count: the weight rows beyond its own taps are zero, the bins it reads
    // there are this transform's own (finite) scratch.
    float2 acc[G], acc1[G];
#pragma unroll
    for (int g = 0; g < G; ++g) { acc[g] = make_float2(0.f, 0.f); acc1[g] = make_float2(0.f, 0.f); }
    const int np = tb.pairs[r];
    const float2* zk = z + k0;
    for (int it = 0; it < np; ++it) {
      const float w0 = wr[64 * it], w1 = wr[64 * it + 32];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        acc[g] = cfma_s(zk[g * ZL + 2 * it], w0, acc[g]);
        acc1[g] = cfma_s(zk[g * ZL + 2 * it + 1], w1, acc1[g]);
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] = cadd(acc[g], acc1[g]);
    if (__any_sync(0xffffffffu, flags & kMelPartner)) {        // a filter cut in two: its halves sit on adjacent lanes
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float ox = __shfl_xor_sync(0xffffffffu, acc[g].x, 1), oy = __shfl_xor_sync(0xffffffffu, acc[g].y, 1);
        if (flags & kMelPartner) acc[g] = cadd(acc[g], make_float2(ox, oy));
      }
    }
    if (flags & kMelOwner) {
      const int m = (task.y >> 16) & 0xff;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int ta = f0 + 2 * g;
        if (ta < T) { const float d = power_to_db(acc[g].x); sink(m, 2 * g, d); vmax = fmax_nan(vmax, d); }
        if (ta + 1 < T) { const float d = power_to_db(acc[g].y); sink(m, 2 * g + 1, d); vmax = fmax_nan(vmax, d); }
      }
    }
  }
  __syncwarp();
  return vmax;
}

// ---- the single-kernel path ------------------------------------------------------------------------
template <int NFFT, int HOP32, typename OutT>
__global__ void __launch_bounds__(StftPlan<NFFT>::kThreads, StftPlan<NFFT>::kMinCtas) feat_kernel(const FeatParams p) {
  using Plan = StftPlan<NFFT>;
  using Rad = typename Plan::Rad;
  constexpr int G = Plan::G;
  constexpr int ZL = stft_zlen<NFFT>();                      // scratch elements per FFT (with padding)

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
  int2* s_tasks = reinterpret_cast<int2*>(smem + p.off_meltasks);
  unsigned char* s_rowmask = reinterpret_cast<unsigned char*>(smem + p.off_rowmask);
  unsigned char* s_colmask = reinterpret_cast<unsigned char*>(smem + p.off_colmask);
  float2* z = reinterpret_cast<float2*>(smem + p.off_z) + (size_t)warp * G * ZL;
  const MelTables tb{s_window, s_tw, s_melw, s_tasks, p.mel_rounds, p.mel_pairs};

  // ---- constants -> shared memory, once per (persistent) CTA ---------------------------
  for (int i = tid; i < NFFT; i += blockDim.x) s_window[i] = __ldg(p.window + i);
  for (int i = tid; i < Rad::tw_total; i += blockDim.x) s_tw[i] = __ldg(p.tw + i);
  for (int i = tid; i < p.n_melw; i += blockDim.x) s_melw[i] = __ldg(p.mel_w + i);
  for (int i = tid; i < p.mel_rounds * 32; i += blockDim.x) s_tasks[i] = __ldg(p.mel_tasks + i);
  if (p.is_mfcc)
    for (int i = tid; i < M * p.c8; i += blockDim.x) {
      const int m = i / p.c8, c = i - m * p.c8;
      s_dct[i] = c < p.n_mfcc ? __ldg(p.dct + (size_t)m * p.n_mfcc + c) : 0.f;
    }
  pdl_wait();   // programmatic dependent launch: the constants above were staged while conv_kernel drained

  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    // ---- per-clip setup -------------------------------------------------------------
    const bool has_rev = clip_has_rev(p.rev, p.rir_idx, p.n_rir, b);
    const float* x = has_rev ? p.rev + (size_t)b * p.rev_stride : p.wav + (size_t)b * p.wav_stride;

    // SpecAugment flags: one byte per feature row / frame
    for (int i = tid; i < F + T; i += blockDim.x) {
      const bool is_row = i < F;
      const int q = is_row ? i : i - F;
      const bool mk = is_row ? in_masks(p.fs, p.fl, p.nF, b, q) : in_masks(p.ts, p.tl, p.nT, b, q);
      (is_row ? s_rowmask : s_colmask)[q] = mk ? 1 : 0;
    }

    const ClipNoise cn = resolve_noise(p.noise, p.noise_idx, p.noise_off, b);
    const float scale = clip_mix_scale(cn, x, N, has_rev, p.es_part, p.es_nb, b, p.snr_db, red);
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
      float* tcol = tile + f0;
      vmax = fmax_nan(vmax, frame_group_to_db<NFFT, HOP32>(x, N, T, hop, cn.nz, cn.off, cn.len, scale, f0, z, tb,
                                                           [&](int m, int i, float d) { tcol[m * pitch + i] = d; }));
    }
    // ---- per-clip top_db floor: max over the tile, gathered while it was written ------------
    vmax = warp_max_nan(vmax);
    if (lane == 0) red[warp] = vmax;
    __syncthreads();
    float cutoff = -INFINITY;
    if (p.top_db >= 0.f) {
      float mx = lane < nwarps ? red[lane] : -INFINITY;
      cutoff = warp_max_nan(mx) - p.top_db;
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
            const float v = fmax_nan(tile[m * pitch + t], cutoff);
            bad |= !isfinite(v);
            out[(size_t)m * T + t] = (rm || s_colmask[t]) ? mv : to_out<OutT>(v);
          }
        }
        done = true;
      } else {
        for (int m = warp; m < M; m += nwarps)
          for (int t = lane; t < T; t += 32) tile[m * pitch + t] = fmax_nan(tile[m * pitch + t], cutoff);
      }
    } else {
      // DCT-II: out[c][t] = sum_m dct[m][c] * max(tile[m][t], cutoff).  One task = 8 coefficients x 4
      // frames (frames tb, tb+TB, tb+2TB, tb+3TB so that a warp reads consecutive tile columns): each
      // 16-byte broadcast load of DCT coefficients feeds eight FFMA2.
      const int C = F, c8 = p.c8, ncg = c8 / 8, TB = (T + 3) / 4;
      for (int idx = tid; idx < ncg * TB; idx += blockDim.x) {
        const int cg = idx / TB, tb0 = idx - cg * TB, c0 = cg * 8;
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
          const float* row = tile + m * pitch + tb0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            // frames beyond T read a neighbouring (finite) tile entry and are never stored
            const float a = fmax_nan(row[tb0 + i * TB < T ? i * TB : 0], cutoff);
            acc2[i][0] = cfma_s(make_float2(d0.x, d0.y), a, acc2[i][0]);   // FFMA2: two coefficients per instruction
            acc2[i][1] = cfma_s(make_float2(d0.z, d0.w), a, acc2[i][1]);
            acc2[i][2] = cfma_s(make_float2(d1.x, d1.y), a, acc2[i][2]);
            acc2[i][3] = cfma_s(make_float2(d1.z, d1.w), a, acc2[i][3]);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int t = tb0 + i * TB;
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
// Flat path (few clips, large batches): the fused kernel above ties a CTA to a clip, so B clips over S CTA slots run
// ceil(B / S) rounds and the CTA-wide per-clip phases cost barrier stalls.  Here the frame groups of ALL clips form
// one flat queue of warp-sized work items (feat_frames_kernel), the dB tiles go through an L2-resident scratch
// and a second, fine-grained kernel finishes them (feat_epilogue_mma_kernel for MFCC: the DCT on the tensor cores;
// feat_epilogue_block_kernel for log-mel).  The per-clip noise-mix records come from feat_prep_kernel.
// ==========================================================================================================
// ---- per-clip preparation: the noise-mix record (only when the call mixes noise) ----------------------------
// conv_kernel makes these records itself when it can (wwf_conv.cuh); this kernel serves calls without reverb and clips
// of several overlap-save blocks (their energy partials come from several CTAs).  One CTA per clip.  Everything that
// does not depend on the reverb kernel - the bank lookups and the energy of the noise segment, three levels of
// dependent loads - happens BEFORE the programmatic-launch wait, i.e. in the shadow of conv_kernel's last wave; after
// it only the clip's energy is fetched (the per-block partials conv_kernel left, or a pass over a dry clip).
template <int kUnused = 0>   // a template only so that the header can be included by several translation units
__global__ void __launch_bounds__(256) feat_prep_kernel(const FeatParams p) {
  __shared__ float red[64];
  const int b = blockIdx.x;
  const ClipNoise cn = resolve_noise(p.noise, p.noise_idx, p.noise_off, b);        // CTA-uniform
  ClipMix m{0.f, 0, 0, 1, 0, 0.f, 0.f};
  if (cn.nz != nullptr) {
    if (threadIdx.x < 32) m.en = warp_noise_energy(cn, p.N);
    m.snr = p.snr_db ? __ldg(p.snr_db + b) : 0.f;
    m.has_noise = 1; m.noff = cn.off; m.nlen = cn.len; m.nz_off = (long long)(cn.nz - p.noise.data);
  }
  const bool has_rev = clip_has_rev(p.rev, p.rir_idx, p.n_rir, b);
  pdl_wait();
  if (cn.nz != nullptr) {
    float es = 0.f;
    if (has_rev && p.es_part != nullptr) {
      for (int i = 0; i < p.es_nb; ++i) es += __ldcg(p.es_part + (size_t)b * p.es_nb + i);   // fixed order: deterministic
    } else {
      const float* x = has_rev ? p.rev + (size_t)b * p.rev_stride : p.wav + (size_t)b * p.wav_stride;
      es = block_energy(x, p.N, red);
    }
    m.scale = snr_scale(es, m.en, m.snr);
  }
  if (threadIdx.x == 0) p.mix_g[b] = m;
}

// ---- frames: STFT -> power -> mel -> dB into the global tile (warp-autonomous, flat over clips) ----
template <int NFFT, int HOP32>
__global__ void __launch_bounds__(StftPlan<NFFT>::kFlatThreads, StftPlan<NFFT>::kMinCtas) feat_frames_kernel(const FeatParams p) {
  using Plan = StftPlan<NFFT>;
  using Rad = typename Plan::Rad;
  constexpr int G = Plan::G;
  constexpr int ZL = stft_zlen<NFFT>();                      // scratch elements per FFT (with padding)

  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int N = p.N, T = p.T, hop = p.hop, mp = p.mp;

  float* s_window = smem;
  float2* s_tw = reinterpret_cast<float2*>(smem + p.f_off_tw);
  float* s_melw = smem + p.f_off_melw;
  int2* s_tasks = reinterpret_cast<int2*>(smem + p.f_off_meltasks);
  float2* z = reinterpret_cast<float2*>(smem + p.f_off_z) + (size_t)warp * G * ZL;
  const MelTables tb{s_window, s_tw, s_melw, s_tasks, p.mel_rounds, p.mel_pairs};

  // ---- constants -> shared memory, once per (persistent) CTA; the only CTA-wide barrier ----
  for (int i = tid; i < NFFT; i += blockDim.x) s_window[i] = __ldg(p.window + i);
  for (int i = tid; i < Rad::tw_total; i += blockDim.x) s_tw[i] = __ldg(p.tw + i);
  for (int i = tid; i < p.n_melw; i += blockDim.x) s_melw[i] = __ldg(p.mel_w + i);
  for (int i = tid; i < p.mel_rounds * 32; i += blockDim.x) s_tasks[i] = __ldg(p.mel_tasks + i);
  // programmatic dependent launch: everything above only reads plan constants and overlapped with the tail of
  // the previous kernel; its results (mix records, reverberated clips) are needed from here on
  pdl_wait();
  __syncthreads();

  // Flat queue of (clip, group) items, adjacent warps on adjacent groups of the same clip (their sample spans overlap
  // in L1 / L2).  The (clip, group) pair is advanced incrementally: no division in the loop.
  const int ngroups = p.ngroups;
  const int stride = (int)gridDim.x * nwarps;
  const int sb = stride / ngroups, sg = stride - sb * ngroups;
  const int first = (int)blockIdx.x * nwarps + warp;
  int b = first / ngroups, grp = first - b * ngroups;
  while (b < p.B) {
    const bool has_rev = clip_has_rev(p.rev, p.rir_idx, p.n_rir, b);
    const float* x = has_rev ? p.rev + (size_t)b * p.rev_stride : p.wav + (size_t)b * p.wav_stride;
    const float* nz = nullptr;
    int noff = 0, nlen = 1;
    float scale = 0.f;
    if (p.mix_g != nullptr) {
      // both halves of the 32-byte record in one go: fetching the noise offset only once has_noise had arrived put the
      // noise loads of the group one L2 round trip behind its sample loads
      const int4 m0 = __ldg(reinterpret_cast<const int4*>(p.mix_g + b));
      const int4 m1 = __ldg(reinterpret_cast<const int4*>(p.mix_g + b) + 1);
      if (m0.y != 0) {
        nz = p.noise.data + (((long long)m1.y << 32) | (unsigned)m1.x);
        scale = __int_as_float(m0.x); noff = m0.z; nlen = m0.w;
      }
    }
    const int f0 = grp * 2 * G;
    float* tg = p.tile_g + ((size_t)b * T + f0) * mp;
    // frame-major tile: the 32 lanes of a round write 32 consecutive floats of one frame
    float vmax = frame_group_to_db<NFFT, HOP32>(x, N, T, hop, nz, noff, nlen, scale, f0, z, tb,
                                                [&](int m, int i, float d) { tg[(size_t)i * mp + m] = d; });
    vmax = warp_max_nan(vmax);
    if (lane == 0 && !(vmax == -INFINITY)) atomicMax(p.clip_max + b, float_key(vmax));
    grp += sg; b += sb;
    if (grp >= ngroups) { grp -= ngroups; ++b; }
  }
}

// ---- log-mel epilogue of the flat path: top_db floor -> [SpecAugment] -> transposed store ----------------
// One CTA per (clip, block of eb_frames frames).  The block's rows of the frame-major tile are one contiguous
// span of global memory: read with 16-byte coalesced loads, written transposed into shared memory
// [n_mels][pitch], stored row by row.  Many small CTAs with short dependent phases.
// exact i / d for 0 <= i < 2^16, 1 <= d < 2^10 without an integer division
__device__ __forceinline__ int small_div(int i, float inv_d) { return __float2int_rd(((float)i + 0.5f) * inv_d); }

template <typename OutT>
__global__ void __launch_bounds__(256) feat_epilogue_block_kernel(const FeatParams p) {
  extern __shared__ __align__(16) float smem[];               // tile [n_mels][pitch]
  __shared__ unsigned char s_rowmask[128];
  __shared__ unsigned char s_colmask[256];
  const int tid = threadIdx.x;
  const int T = p.T, M = p.n_mels, F = p.n_feat, mp = p.mp, pitch = p.eb_pitch;
  const int nblk = (T + p.eb_frames - 1) / p.eb_frames;
  float* tile = smem;
  const int mq = mp / 4;
  const float inv_mq = 1.0f / (float)mq;
  pdl_wait();                            // the tiles and maxima of feat_frames_kernel (PDL)
  const OutT mv = to_out<OutT>(p.mask_value);
  float chk = 0.f;                                            // v * 0 accumulates to NaN iff some v is NaN / Inf
  const int total = p.B * nblk;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int b = item / nblk, t0 = (item - b * nblk) * p.eb_frames;
    const int nf = min(p.eb_frames, T - t0);                  // frames in this block
    for (int i = tid; i < F + nf; i += blockDim.x) {          // SpecAugment flags: feature rows, then this block's frames
      const bool is_row = i < F;
      const bool mk = is_row ? in_masks(p.fs, p.fl, p.nF, b, i) : in_masks(p.ts, p.tl, p.nT, b, t0 + (i - F));
      (is_row ? s_rowmask : s_colmask)[is_row ? i : i - F] = mk ? 1 : 0;
    }
    float cutoff = -INFINITY;
    if (p.top_db >= 0.f) cutoff = clip_max_value(p.clip_max[b]) - p.top_db;
    // the block's rows of the frame-major tile are one contiguous span: coalesced 16-byte loads, transposed into smem
    const float4* tg = reinterpret_cast<const float4*>(p.tile_g + ((size_t)b * T + t0) * mp);
    for (int i = tid; i < nf * mq; i += blockDim.x) {
      const int tl = small_div(i, inv_mq), m = 4 * (i - tl * mq);
      const float4 v = tg[i];
      tile[m * pitch + tl] = fmax_nan(v.x, cutoff);
      if (m + 1 < M) tile[(m + 1) * pitch + tl] = fmax_nan(v.y, cutoff);
      if (m + 2 < M) tile[(m + 2) * pitch + tl] = fmax_nan(v.z, cutoff);
      if (m + 3 < M) tile[(m + 3) * pitch + tl] = fmax_nan(v.w, cutoff);
    }
    __syncthreads();
    OutT* out = reinterpret_cast<OutT*>(p.out) + (size_t)b * p.out_stride + t0;
    const float inv_nf = 1.0f / (float)nf;
    for (int i = tid; i < M * nf; i += blockDim.x) {
      const int m = small_div(i, inv_nf), tl = i - m * nf;
      const float v = tile[m * pitch + tl];
      chk = fmaf(v, 0.f, chk);
      out[(size_t)m * T + tl] = (s_rowmask[m] || s_colmask[tl]) ? mv : to_out<OutT>(v);
    }
    __syncthreads();                                          // tile and flags are reused by the next block
  }
  if (__any_sync(0xffffffffu, chk != 0.f) && (tid & 31) == 0 && p.nonfinite_flag != nullptr) atomicOr(p.nonfinite_flag, 1);
}

// ---- MFCC epilogue of the flat path on the tensor cores: top_db floor -> DCT-II -> [SpecAugment] -> store ----------
// The frame-major dB tiles of all clips are ONE row-major matrix A of B*T rows x n_mels columns, and
//     mfcc[row][c] = sum_m max(A[row][m], cutoff(clip(row))) * dct[m][c]
// is a [B*T x n_mels] x [n_mels x n_mfcc] GEMM (oracle: torchaudio MFCC's
// matmul(mel_db.T, create_dct(...)), TA/transforms/_transforms.py:672-719).  The scalar version spent 4x the FFMA
// minimum in issue slots (profiles/r01_ncu_split_path_summary.txt); here a CTA takes 128 consecutive rows, clamps
// them into shared memory with 16-byte loads and every warp multiplies a 16-row slab with legacy mma.sync
// m16n8k8 TF32 tiles.  1e-3 dB needs more than TF32's 10 mantissa bits: both operands are split hi + lo
// (x = tf32(x) + tf32(x - tf32(x))) and three products are accumulated in float32 (lo*hi + hi*lo + hi*hi), which
// carries ~21 bits - measured against the float64 oracle in tests/test_gpu_parity.py.  The tensor core's float32
// accumulator rounds toward zero, a bias proportional to the running sum (7e-4 on c0 of a 128-mel clip, measured
// against the FFMA kernel), so the rows are centred first: after the floor every value of a clip lies in
// [max - top_db, max]; with mu = max - top_db / 2 the kernel multiplies (x - mu), |x - mu| <= top_db / 2, and adds
// mu * colsum(dct)[c] in float32 at the end (the DCT is linear) - sums 3-25x smaller, silence exact.  The DCT matrix arrives
// already split and laid out fragment by fragment (one 16-byte shared-memory load per lane, k-step and tile,
// build_dct_fragments in wwf_tables.h).  The D fragment gives every lane (row, coefficient) pairs, so the
// [n_mfcc][T] output is written straight from registers (8 consecutive frames per coefficient = one 32-byte sector
// per lane group).  A rows sit at a pitch = 4 (mod 8) floats: conflict-free fragment loads.
constexpr int kEmRows = 128;          // rows per CTA (8 warps x one 16-row m-tile)
constexpr int kEmThreads = 256;
constexpr int kEmMaxSlots = kEmRows + 1;       // clips a CTA's rows can touch (T >= 1)
constexpr int kEmNT = 5;              // n-tiles (8 coefficients each) accumulated per pass over A

__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r = 0;
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 800)
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
#endif
  return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 800)
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
#endif
}

// KS / NTC > 0: the number of k-steps (n_mels / 8, rounded up) / of 8-coefficient tiles (<= kEmNT) is a compile-time
// constant - the tile predicates fold and the k loop unrolls (40 mels x 40 coefficients: KS = NTC = 5); 0 = run time.
template <typename OutT, int KS = 0, int NTC = 0>
__global__ void __launch_bounds__(kEmThreads, 3) feat_epilogue_mma_kernel(const FeatParams p) {
  static_assert(NTC <= kEmNT, "a specialised tile count must fit one pass");
  extern __shared__ __align__(16) float smem[];               // A [128][ap] | B fragments uint4 [k8/8][c8/8][32]
  __shared__ float s_cut[kEmRows], s_mu[kEmRows];             // per row: top_db floor and centre of its clip's value range
  __shared__ float s_colsum[128];
  __shared__ int s_clip[kEmRows];                             // clip of each row (-1 beyond the last row)
  __shared__ int s_frame[kEmRows];
  __shared__ unsigned char s_colmask[kEmRows];
  __shared__ uint32_t s_rowbits[kEmMaxSlots][4];              // per clip slot: bit c set = feature row c is masked
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = p.T, M = p.n_mels, C = p.n_feat, mp = p.mp, k8 = KS ? KS * 8 : p.em_k8, ap = p.em_ap, c8 = NTC ? NTC * 8 : p.c8;
  const int ntiles = c8 / 8, nfrag = (k8 / 8) * ntiles * 32;
  float* sA = smem;
  uint4* sB = reinterpret_cast<uint4*>(smem + kEmRows * ap);
  // plan constants (staged before the PDL wait): the DCT fragments; the padding columns of A are zeroed once
  for (int i = tid; i < nfrag; i += kEmThreads) sB[i] = __ldg(p.dct_frag + i);
  for (int i = tid; i < c8; i += kEmThreads) s_colsum[i] = __ldg(p.dct_colsum + i);
  for (int i = tid; i < kEmRows * (ap - mp); i += kEmThreads) {
    const int row = i / (ap - mp);
    sA[row * ap + mp + (i - row * (ap - mp))] = 0.f;
  }
  pdl_wait();                                                 // the tiles and maxima of feat_frames_kernel
  const OutT mv = to_out<OutT>(p.mask_value);
  const long long rows = (long long)p.B * T;
  const int mq = mp / 4;
  const float inv_mq = 1.0f / (float)mq;
  const bool has_fmask = p.fs != nullptr && p.nF > 0;         // CTA-uniform
  float chk = 0.f;                                            // v * 0 accumulates to NaN iff some v is NaN / Inf
  constexpr int NL = 5;                                       // 16-byte tile loads in flight per thread
  for (long long r0 = (long long)blockIdx.x * kEmRows; r0 < rows; r0 += (long long)gridDim.x * kEmRows) {
    __syncthreads();                                          // A and the row tables of the previous block are free
    const int b_first = (int)(r0 / T);
    const int nrows = (int)min((long long)kEmRows, rows - r0);
    // A: 128 rows x mp floats are one contiguous span of the flat tile matrix.  All loads of a pass are issued
    // before anything waits on them (the first version waited per load: 21 % of its stall samples).
    const float4* tg = reinterpret_cast<const float4*>(p.tile_g + (size_t)r0 * mp);
    const int nload = nrows * mq;
    for (int base = 0; base < kEmRows * mq; base += kEmThreads * NL) {
      float4 v[NL];
#pragma unroll
      for (int u = 0; u < NL; ++u) {
        const int i = base + u * kEmThreads + tid;
        v[u] = i < nload ? __ldcg(tg + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (base == 0) {
        if (tid < kEmRows) {                                  // per-row clip, frame, floor and column flag
          const long long r = r0 + tid;
          int b = -1, t = 0;
          float cut = -INFINITY, mu = 0.f;
          bool cm = false;
          if (r < rows) {
            b = b_first + (tid + (int)(r0 - (long long)b_first * T)) / T;
            t = (int)(r - (long long)b * T);
            const float mx = clip_max_value(__ldcg(p.clip_max + b));
            if (p.top_db >= 0.f) cut = mx - p.top_db;
            mu = mx - (p.top_db >= 0.f ? 0.5f * p.top_db : 40.0f);
            cm = in_masks(p.ts, p.tl, p.nT, b, t);
          }
          s_clip[tid] = b; s_frame[tid] = t; s_cut[tid] = cut; s_mu[tid] = mu; s_colmask[tid] = cm ? 1 : 0;
        }
        if (has_fmask) {
          const int nslots = min(kEmMaxSlots, min(p.B - b_first, (kEmRows + T - 1) / T + 1));
          for (int i = tid; i < nslots * 4; i += kEmThreads) {
            const int slot = i >> 2, w0 = (i & 3) * 32, b = b_first + slot;
            uint32_t bits = 0;
            for (int j = 0; j < p.nF; ++j) {
              const int s0 = __ldg(p.fs + (size_t)b * p.nF + j), l = __ldg(p.fl + (size_t)b * p.nF + j);
              const int lo = max(s0, w0) - w0, hi = min(s0 + l, w0 + 32) - w0;      // bit range inside this word
              if (hi > lo) bits |= (hi - lo >= 32 ? 0xffffffffu : ((1u << (hi - lo)) - 1u) << lo);
            }
            s_rowbits[slot][i & 3] = bits;
          }
        }
        __syncthreads();
      }
#pragma unroll
      for (int u = 0; u < NL; ++u) {
        const int i = base + u * kEmThreads + tid;
        if (i < kEmRows * mq) {
          const int row = small_div(i, inv_mq), m = 4 * (i - row * mq);
          const float cut = s_cut[row], mu = s_mu[row];        // rows beyond the last: zeros, floor -inf, centre 0
          float4 w = v[u];
          w.x = fmax_nan(w.x, cut) - mu;
          w.y = m + 1 < M ? fmax_nan(w.y, cut) - mu : 0.f;
          w.z = m + 2 < M ? fmax_nan(w.z, cut) - mu : 0.f;
          w.w = m + 3 < M ? fmax_nan(w.w, cut) - mu : 0.f;
          *reinterpret_cast<float4*>(sA + row * ap + m) = w;
        }
      }
    }
    __syncthreads();
    // one 16-row slab per warp; D fragment: c0/c1 = (row g, coefficients 2t, 2t+1), c2/c3 = (row g + 8, same)
    const int g = lane >> 2, t4 = lane & 3;
    const int rowA = warp * 16 + g, rowB = rowA + 8;
    const int bA = s_clip[rowA], bB = s_clip[rowB];
    const float* arow0 = sA + rowA * ap + t4;
    const float* arow1 = sA + rowB * ap + t4;
    // element (row, coefficient c) lives at out[clip][c][frame]: base pointers per row, 32-bit offsets c * T
    OutT* outA = reinterpret_cast<OutT*>(p.out) + (size_t)max(bA, 0) * p.out_stride + s_frame[rowA];
    OutT* outB = reinterpret_cast<OutT*>(p.out) + (size_t)max(bB, 0) * p.out_stride + s_frame[rowB];
    const bool okA = bA >= 0, okB = bB >= 0;
    const bool cmA = s_colmask[rowA] != 0, cmB = s_colmask[rowB] != 0;
    const float muA = s_mu[rowA], muB = s_mu[rowB];
    const uint32_t* bitsA = s_rowbits[min(max(bA - b_first, 0), kEmMaxSlots - 1)];
    const uint32_t* bitsB = s_rowbits[min(max(bB - b_first, 0), kEmMaxSlots - 1)];
    // warp-uniform: every row of the slab exists, nothing of it is masked, no padded coefficient -> plain stores
    const bool plain = !has_fmask && C == c8 && __all_sync(0xffffffffu, okA && okB && !cmA && !cmB);
    for (int n0 = 0; n0 < ntiles; n0 += kEmNT) {
      const int nt = NTC ? NTC : min(kEmNT, ntiles - n0);      // tiles of this pass (warp-uniform)
      float acc[kEmNT][4];
#pragma unroll
      for (int j = 0; j < kEmNT; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; acc[j][3] = 0.f; }
      const uint4* bf = sB + (size_t)n0 * 32 + lane;
#pragma unroll(KS > 0 && KS <= 8 ? KS : 1)
      for (int k0 = 0; k0 < k8; k0 += 8, bf += ntiles * 32) {
        const float av[4] = {arow0[k0], arow1[k0], arow0[k0 + 4], arow1[k0 + 4]};
        uint32_t ah[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { ah[i] = tf32_rna(av[i]); al[i] = tf32_rna(av[i] - __uint_as_float(ah[i])); }
        uint4 b[kEmNT];                                        // (hi k, hi k+4, lo k, lo k+4) of this lane's B fragment
#pragma unroll
        for (int j = 0; j < kEmNT; ++j) b[j] = j < nt ? bf[j * 32] : make_uint4(0u, 0u, 0u, 0u);
        // the three products of a tile form a dependent chain: interleave the tiles, small terms first
#pragma unroll
        for (int j = 0; j < kEmNT; ++j) if (j < nt) mma_tf32_16x8x8(acc[j], al, b[j].x, b[j].y);
#pragma unroll
        for (int j = 0; j < kEmNT; ++j) if (j < nt) mma_tf32_16x8x8(acc[j], ah, b[j].z, b[j].w);
#pragma unroll
        for (int j = 0; j < kEmNT; ++j) if (j < nt) mma_tf32_16x8x8(acc[j], ah, b[j].x, b[j].y);
      }
#pragma unroll
      for (int j = 0; j < kEmNT; ++j) {
        if (j < nt) {
          const int c = (n0 + j) * 8 + 2 * t4;                 // this lane's coefficient pair of the tile: c, c + 1
          const float2 cs = *reinterpret_cast<const float2*>(s_colsum + c);
          acc[j][0] = fmaf(muA, cs.x, acc[j][0]); acc[j][1] = fmaf(muA, cs.y, acc[j][1]);   // un-centre: + mu * colsum[c]
          acc[j][2] = fmaf(muB, cs.x, acc[j][2]); acc[j][3] = fmaf(muB, cs.y, acc[j][3]);
          chk = fmaf(acc[j][0], 0.f, fmaf(acc[j][1], 0.f, fmaf(acc[j][2], 0.f, fmaf(acc[j][3], 0.f, chk))));
          const int o0 = c * T, o1 = o0 + T;
          if (plain) {
            outA[o0] = to_out<OutT>(acc[j][0]); outA[o1] = to_out<OutT>(acc[j][1]);
            outB[o0] = to_out<OutT>(acc[j][2]); outB[o1] = to_out<OutT>(acc[j][3]);
          } else {
            bool m0A = cmA, m1A = cmA, m0B = cmB, m1B = cmB;
            if (has_fmask) {
              const uint32_t wa = bitsA[c >> 5] >> (c & 31), wb = bitsB[c >> 5] >> (c & 31);   // c is even: c + 1 is in the same word
              m0A |= wa & 1u; m1A |= (wa >> 1) & 1u; m0B |= wb & 1u; m1B |= (wb >> 1) & 1u;
            }
            if (okA && c < C) outA[o0] = m0A ? mv : to_out<OutT>(acc[j][0]);
            if (okA && c + 1 < C) outA[o1] = m1A ? mv : to_out<OutT>(acc[j][1]);
            if (okB && c < C) outB[o0] = m0B ? mv : to_out<OutT>(acc[j][2]);
            if (okB && c + 1 < C) outB[o1] = m1B ? mv : to_out<OutT>(acc[j][3]);
          }
        }
      }
    }
  }
  if (__any_sync(0xffffffffu, chk != 0.f) && lane == 0 && p.nonfinite_flag != nullptr) atomicOr(p.nonfinite_flag, 1);
}

// ---- the same epilogue for the unmasked common shapes, warp-autonomous and software-pipelined -------------------
// feat_epilogue_mma_kernel is a sequence of CTA-wide phases per 128-row block (issue the loads | barrier | clamp into
// shared memory | barrier | MMA | stores) and ncu shows it latency-bound, not issue- or bandwidth-bound (a third of the
// issue slots used, 9 M warp-instructions for 49 MB).  Without SpecAugment flags, with n_mels and n_mfcc multiples of 8
// known at compile time and every row complete in the tile matrix (mp == n_mels), a WARP can own a 16-row slab from
// the load to the store: cp.async brings the NEXT slab's 16 x n_mels raw dB values (one contiguous span of the flat
// matrix) into the warp's second buffer while it works on the current one, the clip maxima of the next slab's rows
// are requested one iteration ahead too, the top_db floor and the centring happen on the A fragments as they are
// loaded (every element is loaded by exactly one lane), and nothing but __syncwarp() synchronises.
// Same arithmetic as feat_epilogue_mma_kernel (split TF32, rows centred, un-centred in float32).
template <int KS> struct EmWarp {
  static constexpr int kAp = KS * 8 + 4;                     // pitch = 4 (mod 8) floats: conflict-free fragment loads
  static constexpr int kBufFloats = 16 * kAp;
  static constexpr int kWarps = kEmThreads / 32;
  static constexpr size_t smem_bytes(int ntc) {
    return (size_t)kWarps * 2 * kBufFloats * sizeof(float) + (size_t)KS * ntc * 32 * sizeof(uint4);
  }
};

template <typename OutT, int KS, int NTC>
__global__ void __launch_bounds__(kEmThreads, 3) feat_epilogue_mma_warp_kernel(const FeatParams p) {
  static_assert(NTC >= 1 && NTC <= kEmNT, "one pass over A");
  using W = EmWarp<KS>;
  constexpr int M = KS * 8, kAp = W::kAp, kChunks = 16 * M / 4;      // 16-byte chunks per slab
  extern __shared__ __align__(16) float smem[];               // per warp: 2 x A [16][kAp] | B fragments uint4 [KS][NTC][32]
  __shared__ float s_colsum[NTC * 8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* bufs = smem + (size_t)warp * 2 * W::kBufFloats;
  uint4* sB = reinterpret_cast<uint4*>(smem + (size_t)W::kWarps * 2 * W::kBufFloats);
  for (int i = tid; i < KS * NTC * 32; i += kEmThreads) sB[i] = __ldg(p.dct_frag + i);
  for (int i = tid; i < NTC * 8; i += kEmThreads) s_colsum[i] = __ldg(p.dct_colsum + i);
  __syncthreads();                                            // the only CTA barrier
  pdl_wait();                                                 // the tiles and maxima of feat_frames_kernel
  const int T = p.T;
  const long long rows = (long long)p.B * T;
  const long long nslabs = (rows + 15) / 16;
  const long long stride = (long long)gridDim.x * W::kWarps;
  const int g = lane >> 2, t4 = lane & 3;
  const bool has_floor = p.top_db >= 0.f;
  const float half_range = has_floor ? 0.5f * p.top_db : 40.0f;
  float chk = 0.f;
  // stage slab sl into buffer bi: chunk i = row i / (M/4), columns 4 (i mod M/4) ..; rows beyond the matrix are not read
  auto issue = [&](long long sl, int bi) {
    const float* src = p.tile_g + (size_t)sl * 16 * M;
    const long long left = (rows - sl * 16) * (M / 4);        // chunks that exist
    float* dstb = bufs + bi * W::kBufFloats;
#pragma unroll
    for (int u = 0; u < (kChunks + 31) / 32; ++u) {
      const int i = u * 32 + lane;
      if (i < kChunks && i < left) {
        const int row = i / (M / 4), c4 = i - row * (M / 4);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(dstb + row * kAp + 4 * c4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + 4 * i) : "memory");
      }
    }
  };
  // clip and frame of a row, and its clip's running-maximum key (rows beyond the matrix: clip -1)
  auto row_info = [&](long long r, int& b, int& t, int& key) {
    b = -1; t = 0; key = kMaxKeyMemset;
    if (r < rows) {
      b = (int)(r / T);
      t = (int)(r - (long long)b * T);
      key = __ldcg(p.clip_max + b);
    }
  };
  long long sl = (long long)blockIdx.x * W::kWarps + warp;
  int bA, tA, keyA, bB, tB, keyB;
  if (sl < nslabs) {
    issue(sl, 0);
    row_info(sl * 16 + g, bA, tA, keyA);
    row_info(sl * 16 + g + 8, bB, tB, keyB);
  }
  for (int cur = 0; sl < nslabs; sl += stride, cur ^= 1) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();                                             // every lane's chunks have landed; the other buffer is free
    const long long nsl = sl + stride;
    int nbA = -1, ntA = 0, nkeyA = kMaxKeyMemset, nbB = -1, ntB = 0, nkeyB = kMaxKeyMemset;
    if (nsl < nslabs) {
      issue(nsl, cur ^ 1);
      row_info(nsl * 16 + g, nbA, ntA, nkeyA);
      row_info(nsl * 16 + g + 8, nbB, ntB, nkeyB);
    }
    const float mxA = clip_max_value(keyA), mxB = clip_max_value(keyB);
    const float cutA = has_floor ? mxA - p.top_db : -INFINITY, cutB = has_floor ? mxB - p.top_db : -INFINITY;
    const float muA = bA >= 0 ? mxA - half_range : 0.f, muB = bB >= 0 ? mxB - half_range : 0.f;
    const float* arow0 = bufs + cur * W::kBufFloats + g * kAp + t4;
    const float* arow1 = arow0 + 8 * kAp;
    float acc[NTC][4];
#pragma unroll
    for (int j = 0; j < NTC; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; acc[j][3] = 0.f; }
    const uint4* bf = sB + lane;
#pragma unroll(KS <= 8 ? KS : 2)
    for (int ks = 0; ks < KS; ++ks) {
      // rows beyond the matrix hold stale buffer contents: they become zeros here (their results are not stored)
      const float av[4] = {bA >= 0 ? fmax_nan(arow0[8 * ks], cutA) - muA : 0.f, bB >= 0 ? fmax_nan(arow1[8 * ks], cutB) - muB : 0.f,
                           bA >= 0 ? fmax_nan(arow0[8 * ks + 4], cutA) - muA : 0.f, bB >= 0 ? fmax_nan(arow1[8 * ks + 4], cutB) - muB : 0.f};
      uint32_t ah[4], al[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { ah[i] = tf32_rna(av[i]); al[i] = tf32_rna(av[i] - __uint_as_float(ah[i])); }
      uint4 b[NTC];
#pragma unroll
      for (int j = 0; j < NTC; ++j) b[j] = bf[(ks * NTC + j) * 32];
#pragma unroll
      for (int j = 0; j < NTC; ++j) mma_tf32_16x8x8(acc[j], al, b[j].x, b[j].y);
#pragma unroll
      for (int j = 0; j < NTC; ++j) mma_tf32_16x8x8(acc[j], ah, b[j].z, b[j].w);
#pragma unroll
      for (int j = 0; j < NTC; ++j) mma_tf32_16x8x8(acc[j], ah, b[j].x, b[j].y);
    }
    OutT* outA = reinterpret_cast<OutT*>(p.out) + (size_t)max(bA, 0) * p.out_stride + tA;
    OutT* outB = reinterpret_cast<OutT*>(p.out) + (size_t)max(bB, 0) * p.out_stride + tB;
#pragma unroll
    for (int j = 0; j < NTC; ++j) {
      const int c = j * 8 + 2 * t4;
      const float2 cs = *reinterpret_cast<const float2*>(s_colsum + c);
      acc[j][0] = fmaf(muA, cs.x, acc[j][0]); acc[j][1] = fmaf(muA, cs.y, acc[j][1]);
      acc[j][2] = fmaf(muB, cs.x, acc[j][2]); acc[j][3] = fmaf(muB, cs.y, acc[j][3]);
      const int o0 = c * T, o1 = o0 + T;
      if (bA >= 0) {
        chk = fmaf(acc[j][0], 0.f, fmaf(acc[j][1], 0.f, chk));
        outA[o0] = to_out<OutT>(acc[j][0]); outA[o1] = to_out<OutT>(acc[j][1]);
      }
      if (bB >= 0) {
        chk = fmaf(acc[j][2], 0.f, fmaf(acc[j][3], 0.f, chk));
        outB[o0] = to_out<OutT>(acc[j][2]); outB[o1] = to_out<OutT>(acc[j][3]);
      }
    }
    bA = nbA; tA = ntA; keyA = nkeyA; bB = nbB; tB = ntB; keyB = nkeyB;
  }
  if (__any_sync(0xffffffffu, chk != 0.f) && lane == 0 && p.nonfinite_flag != nullptr) atomicOr(p.nonfinite_flag, 1);
}

}  // namespace wwf
