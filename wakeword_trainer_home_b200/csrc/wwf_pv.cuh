// wwf_pv.cuh - time-stretch, pitch-shift and sinc resampling.
//
// The arithmetic is torchaudio's (the reference's AudioAugmentation module is absent from its checkout;
// its kwargs time_stretch_range / pitch_shift_range are at tests/test_training_pipeline.py:233-234):
//   stretch(rate)  = STFT(n_fft 512, hop 128, periodic Hann, reflect) -> F.phase_vocoder(rate) ->
//                    iSTFT(length = round(N / rate))            TA/functional/functional.py:1644-1693, 732-800
//   pitch_shift(n) = stretch(2^(-n/12)) -> F.resample(int(sr / rate) -> sr) -> crop / zero-pad to N   :1596-1641
//   resample       = polyphase windowed-sinc ("sinc_interp_hann", width 6, rolloff 0.99)              :1305-1497
//
// Data flow per clip (all intermediates in a caller-provided HBM workspace):
//   pv_stft_kernel     wav[N]            -> S[T][272] complex    warp per frame PAIR (two real frames per
//                                                                complex FFT), same radix-16.8.4 plan and
//                                                                padded scratch as feat_kernel<512>
//   pv_vocoder_kernel  S                 -> V[T'][272] complex   thread per (clip, bin): the phase accumulation
//                                                                is a running product of unit phasors over time
//   pv_synth_kernel    V                 -> y[len]               CTA per 17 hop blocks: Hermitian pack, inverse
//                                                                FFTs, overlap-add / window envelope from smem
//   resample_kernel    y (+ coefficient table) -> out[N]
// Accuracy note: torchaudio accumulates the vocoder phase with a float32 cumsum; at ~7e4 rad one float32
// ulp is 8e-3 rad, which is why torchaudio's own float32 and float64 results differ by ~1e-3 relative.
// Here the phase is carried as a unit phasor (see pv_vocoder_kernel), whose error does not grow with the
// size of the angle, so the result sits at the float64 oracle (1e-5) instead of 1e-3 away from it.
#pragma once
#include <stdint.h>
#include "wwf_feat.cuh"

namespace wwf {

constexpr int kPvN = 512, kPvHop = 128, kPvK = 257;
constexpr int kPvPitch = 272;             // float2 per spectrum row (257 bins, rows 128-byte aligned)
constexpr int kPvWarps = 8;               // warps per CTA of the two FFT kernels = frame pairs per CTA
constexpr int kPvMaxSteps = 25;           // pitch range [-12, 12] semitones

using PvPlan = StftPlan<kPvN>;
using PvRad = PvPlan::Rad;
using PvMap = PvPlan::Map;
constexpr int kPvZL = stft_zlen<kPvN>();

// Per-clip rate source: explicit double rates (time-stretch) or integer semitones through a host-built
// table (pitch-shift; the host computes 2^(-n/12) with the C library exactly like the Python reference).
struct PvRate {
  const double* rates;        // [B] or nullptr
  const int32_t* steps;       // [B] or nullptr
  int step_lo, n_steps;       // table covers step_lo .. step_lo + n_steps - 1
  double rate_tab[kPvMaxSteps];
};
__device__ __forceinline__ double pv_clip_rate(const PvRate& r, int b) {
  if (r.rates != nullptr) return r.rates[b];
  const int i = r.steps[b] - r.step_lo;
  return (i < 0 || i >= r.n_steps) ? 1.0 : r.rate_tab[i];
}
// ceil(T / rate) frames (torch.arange(0, T, rate) element count) and round-half-even(N / rate) samples,
// both clamped to the workspace capacity the host sized from its lower bound on the rates
__device__ __forceinline__ int pv_out_frames(int T, double rate, int cap) {
  const double v = ceil((double)T / rate);
  return v > (double)cap ? cap : (int)v;
}
__device__ __forceinline__ int pv_out_len(int N, double rate, int cap) {
  const double v = rint((double)N / rate);
  return v > (double)cap ? cap : (int)v;
}

struct PvParams {
  const float* wav; int64_t wav_stride;
  int B, N, T;                 // T = N / 128 + 1 input frames
  int Tcap, Lcap;              // capacity of V / Y rows and of the stretched waveform per clip
  PvRate rate;
  const float* window;         // [512] periodic Hann, float32
  const float2* tw;            // forward twiddles of Radices<16, 8, 4>
  float2* S; float2* V;        // [B][T][272], [B][Tcap][272]
  float* out; int64_t out_stride; int n_out;   // pv_synth_kernel destination: n_out samples per clip
  int out_pad;                 // 1: zero-fill [len, n_out) and copy untouched clips (time-stretch); 0: write len only
};

// Two real frames a, b transformed as a + i b: given Z[k] and Z[n-k] return the two complex spectra
//   A[k] = (Z[k] + conj Z[n-k]) / 2,   B[k] = -i (Z[k] - conj Z[n-k]) / 2.
WWF_HD void pv_split(float2 zk, float2 zm, float2* A, float2* Bv) {
  *A = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
  *Bv = make_float2(0.5f * (zk.y + zm.y), 0.5f * (zm.x - zk.x));
}
// Inverse of pv_split for two ONESIDED spectra: Z[k] = A[k] + i B[k] and, by Hermitian symmetry of the
// real frames, Z[n-k] = conj A[k] + i conj B[k].
WWF_HD void pv_pack(float2 a, float2 c, float2* zk, float2* zm) {
  *zk = make_float2(a.x - c.y, a.y + c.x);
  *zm = make_float2(a.x + c.y, c.x - a.y);
}

// ---- STFT: clip -> S ----------------------------------------------------------------------------
__global__ void __launch_bounds__(kPvWarps * 32, 4) pv_stft_kernel(const PvParams p) {
  __shared__ __align__(16) float2 s_z[kPvWarps * kPvZL];
  __shared__ float2 s_tw[PvRad::tw_total];
  __shared__ float s_win[kPvN];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.y;
  const double rate = pv_clip_rate(p.rate, b);
  if (rate == 1.0) return;                                   // untouched clip (CTA-uniform)
  for (int i = tid; i < kPvN; i += blockDim.x) s_win[i] = __ldg(p.window + i);
  for (int i = tid; i < PvRad::tw_total; i += blockDim.x) s_tw[i] = __ldg(p.tw + i);
  __syncthreads();
  const int ta = 2 * (blockIdx.x * kPvWarps + warp), tb = ta + 1;
  if (ta >= p.T) return;
  const PvMap zmap;
  float2* z = s_z + warp * kPvZL;
  const float* x = p.wav + (size_t)b * p.wav_stride;
  const int N = p.N;
  const bool inner = ta * kPvHop - kPvN / 2 >= 0 && tb * kPvHop + kPvN / 2 <= N && tb < p.T;
  if (inner) {
    // frame b is frame a shifted by one hop = 4 x 32 samples: both come from one 640-sample span held in
    // registers in lane-cyclic order, all 20 loads in flight before the first use
    constexpr int NC = kPvN / 32, SH = kPvHop / 32;
    float r[NC + SH];
    const float* xs = x + ta * kPvHop - kPvN / 2 + lane;
#pragma unroll
    for (int i = 0; i < NC + SH; ++i) r[i] = __ldg(xs + 32 * i);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int j = lane + 32 * i;
      const float w = s_win[j];
      z[zmap(j)] = make_float2(r[i] * w, r[i + SH] * w);
    }
  } else {
#pragma unroll 1
    for (int j = lane; j < kPvN; j += 32) {                   // boundary frames: reflect padding (cold)
      const float w = s_win[j];
      float re = __ldg(x + reflect_index(ta * kPvHop - kPvN / 2 + j, N)), im = 0.f;
      if (tb < p.T) im = __ldg(x + reflect_index(tb * kPvHop - kPvN / 2 + j, N));
      z[zmap(j)] = make_float2(re * w, im * w);
    }
  }
  __syncwarp();
  static_for<0, PvRad::npass>([&](auto I) {
    constexpr int i = decltype(I)::value;
    constexpr int R = PvRad::R(i), L = PvRad::L(i), tasks = kPvN / R;
    const float2* tw = s_tw + PvRad::tw_off(i);
#pragma unroll 1
    for (int u = lane; u < tasks; u += 32) pass_task<R, false, PvMap>(z, L, u, [&](int q) { return tw[q]; });
    __syncwarp();
  });
  // split the packed pair: A[k] = (Z[k] + conj Z[n-k]) / 2,  B[k] = -i (Z[k] - conj Z[n-k]) / 2
  float2* Sa = p.S + ((size_t)b * p.T + ta) * kPvPitch;
  float2* Sb = Sa + kPvPitch;
  for (int k = lane; k < kPvK; k += 32) {
    float2 A, Bv;
    pv_split(z[zmap(PvRad::pos(k))], z[zmap(PvRad::pos(k == 0 ? 0 : kPvN - k))], &A, &Bv);
    Sa[k] = A;
    if (tb < p.T) Sb[k] = Bv;
  }
}

// ---- phase vocoder: S -> V -----------------------------------------------------------------------
// F.phase_vocoder builds, per bin, phase_acc[j] = angle(S[0]) + sum_{i<j} (pa + wrap(angle(S[i1]) - angle(S[i0]) - pa))
// and emits mag[j] * exp(i phase_acc[j]).  wrap() only removes multiples of 2 pi and the sum is used solely
// through cos / sin, so exp(i phase_acc[j+1]) = exp(i phase_acc[j]) * u(S[i1]) * conj(u(S[i0])) with
// u(z) = z / |z| (u(0) = 1 because torch's angle(0) is 0): the accumulation is a running product of unit
// complex numbers - no atan2, no sincos, no phase_advance table.  Each product rounds the angle by ~6e-8 rad
// (a random walk: ~1e-6 rad after 200 frames, 2e-5 after 2^17), far below torchaudio's own float32 cumsum
// (ulp 8e-3 rad at 7e4 rad); the phasor is renormalised every step.
// One thread per (clip, bin); the frame loads of kVocU consecutive steps are independent of the running
// product and are issued together.
constexpr int kVocU = 8;

// z / |z| and |z| (u = 1, |z| = 0 for z = 0); tiny inputs are rescaled so that x^2 + y^2 cannot underflow
__device__ __forceinline__ float2 pv_unit(float2 z, float* mag) {
  const float ax = fmaxf(fabsf(z.x), fabsf(z.y));
  if (ax == 0.f) { *mag = 0.f; return make_float2(1.f, 0.f); }
  const bool tiny = ax < 1e-18f;
  const float sc = tiny ? 1.8446744e19f : 1.0f;               // 2^64
  const float x = z.x * sc, y = z.y * sc;
  const float m2 = fmaf(x, x, y * y);
  const float inv = rsqrtf(m2);
  const float m = m2 * inv;                                   // sqrt(m2) to ~2 ulp
  *mag = tiny ? m * 5.4210109e-20f : m;                       // 2^-64
  return make_float2(x * inv, y * inv);
}

__global__ void __launch_bounds__(288) pv_vocoder_kernel(const PvParams p) {
  const int b = blockIdx.x, k = threadIdx.x;
  const double rate = pv_clip_rate(p.rate, b);
  if (rate == 1.0 || k >= kPvK) return;
  const int T = p.T;
  const int To = pv_out_frames(T, rate, p.Tcap);
  const float2* S = p.S + (size_t)b * T * kPvPitch + k;
  float2* V = p.V + (size_t)b * p.Tcap * kPvPitch + k;
  float m_;
  float2 E = pv_unit(S[0], &m_);                             // exp(i phase_0), phase_0 = angle of frame 0
  for (int j0 = 0; j0 < To; j0 += kVocU) {
    float2 s0[kVocU], s1[kVocU];
    float al[kVocU];
#pragma unroll
    for (int u = 0; u < kVocU; ++u) {
      const float ts = (float)((double)(j0 + u) * rate);     // arange(0, T, rate) in float32
      const int i0 = (int)ts, i1 = (int)__fadd_rn(ts, 1.0f);
      al[u] = ts - (float)i0;                                // ts % 1.0 (exact for ts >= 0)
      // frames >= T are the two zero frames torchaudio pads with
      s0[u] = (j0 + u < To && i0 < T) ? S[(size_t)i0 * kPvPitch] : make_float2(0.f, 0.f);
      s1[u] = (j0 + u < To && i1 < T) ? S[(size_t)i1 * kPvPitch] : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kVocU; ++u) {
      if (j0 + u < To) {
        float m0, m1;
        const float2 u0 = pv_unit(s0[u], &m0), u1 = pv_unit(s1[u], &m1);
        const float mag = __fadd_rn(__fmul_rn(al[u], m1), __fmul_rn(__fsub_rn(1.0f, al[u]), m0));
        // output j uses the phase accumulated BEFORE this step's increment (torch.cat([phase_0, phase[:-1]]))
        V[(size_t)(j0 + u) * kPvPitch] = make_float2(mag * E.x, mag * E.y);
        const float2 rot = make_float2(fmaf(u1.x, u0.x, u1.y * u0.y), fmaf(u1.y, u0.x, -u1.x * u0.y));   // u1 * conj(u0)
        const float2 e = make_float2(fmaf(E.x, rot.x, -E.y * rot.y), fmaf(E.x, rot.y, E.y * rot.x));
        const float inv = rsqrtf(fmaf(e.x, e.x, e.y * e.y));
        E = make_float2(e.x * inv, e.y * inv);
      }
    }
  }
}

// ---- synthesis: V -> waveform (inverse STFT frames + overlap-add + window envelope, fused) --------------
// A CTA owns kSynBlocks consecutive hop blocks (128 output samples each) of one clip.  Hop block h is the
// sum of frames h-3 .. h, so the CTA inverse-transforms the kSynFrames = kSynBlocks + 3 frames it needs
// (two per warp, packed in one complex FFT), leaves them in its warps' scratch buffers and overlap-adds
// straight out of shared memory: the windowed frames never travel to HBM.
constexpr int kSynWarps = 10, kSynFrames = 2 * kSynWarps, kSynBlocks = kSynFrames - 3;
constexpr int kSynSmemBytes = (kSynWarps * kPvZL + PvRad::tw_total) * (int)sizeof(float2) + kPvN * (int)sizeof(float);

__global__ void __launch_bounds__(kSynWarps * 32) pv_synth_kernel(const PvParams p) {
  extern __shared__ __align__(16) float2 syn_smem[];
  float2* s_z = syn_smem;
  float2* s_tw = s_z + kSynWarps * kPvZL;
  float* s_win = reinterpret_cast<float*>(s_tw + PvRad::tw_total);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.y;
  const double rate = pv_clip_rate(p.rate, b);
  float* out = p.out + (size_t)b * p.out_stride;
  // hop blocks h0 .. h0 + kSynBlocks - 1; output sample s = 128 h + r - 256 (the first two hop blocks are trimmed)
  const int h0 = 2 + blockIdx.x * kSynBlocks;
  const int s_begin = h0 * kPvHop - kPvN / 2, s_end = min(s_begin + kSynBlocks * kPvHop, p.n_out);
  if (s_begin >= p.n_out) return;
  if (rate == 1.0) {
    if (p.out_pad) {
      const float* x = p.wav + (size_t)b * p.wav_stride;
      for (int s = s_begin + tid; s < s_end; s += blockDim.x) out[s] = s < p.N ? __ldg(x + s) : 0.f;
    }
    return;
  }
  const int To = pv_out_frames(p.T, rate, p.Tcap);
  const int len = pv_out_len(p.N, rate, p.Lcap);
  if (s_begin >= len) {                                        // beyond the stretched signal: zero padding only
    if (p.out_pad) for (int s = s_begin + tid; s < s_end; s += blockDim.x) out[s] = 0.f;
    return;
  }
  for (int i = tid; i < kPvN; i += blockDim.x) s_win[i] = __ldg(p.window + i);
  for (int i = tid; i < PvRad::tw_total; i += blockDim.x) s_tw[i] = __ldg(p.tw + i);
  const int jf = h0 - 3;                                       // first frame this CTA needs (may be < 0 ... never: h0 >= 2 -> jf >= -1)
  const int ja = jf + 2 * warp, jb = ja + 1;
  const PvMap zmap;
  float2* z = s_z + warp * kPvZL;
  const bool ha = ja >= 0 && ja < To, hb = jb >= 0 && jb < To;
  __syncthreads();
  if (ha || hb) {
    const float2* Va = p.V + ((size_t)b * p.Tcap + (ha ? ja : jb)) * kPvPitch;
    const float2* Vb = p.V + ((size_t)b * p.Tcap + (hb ? jb : ja)) * kPvPitch;
    // Z = A + i B with Hermitian extension, written where the forward transform would have left bin k
    // (irfft ignores the imaginary parts of the DC and Nyquist bins)
    constexpr int NK = (kPvK + 31) / 32;
    float2 va[NK], vb[NK];
#pragma unroll
    for (int i = 0; i < NK; ++i) {                             // all loads in flight before the first use
      const int k = lane + 32 * i;
      va[i] = (ha && k < kPvK) ? Va[k] : make_float2(0.f, 0.f);
      vb[i] = (hb && k < kPvK) ? Vb[k] : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < NK; ++i) {
      const int k = lane + 32 * i;
      if (k < kPvK) {
        float2 a = va[i], c = vb[i];
        if (k == 0 || k == kPvN / 2) { a.y = 0.f; c.y = 0.f; }
        float2 zk, zm;
        pv_pack(a, c, &zk, &zm);
        z[zmap(PvRad::pos(k))] = zk;
        if (k > 0 && k < kPvN / 2) z[zmap(PvRad::pos(kPvN - k))] = zm;
      }
    }
    __syncwarp();
    static_for<0, PvRad::npass>([&](auto I) {
      constexpr int i = PvRad::npass - 1 - decltype(I)::value;  // inverse (DIT) passes run in reverse order
      constexpr int R = PvRad::R(i), L = PvRad::L(i), tasks = kPvN / R;
      const float2* tw = s_tw + PvRad::tw_off(i);
#pragma unroll 1
      for (int u = lane; u < tasks; u += 32) pass_task<R, true, PvMap>(z, L, u, [&](int q) { return tw[q]; });
      __syncwarp();
    });
  }
  __syncthreads();
  // overlap-add: sample s = 128 h + r - 256 sums frame j = h - d at n = r + 128 d, d = 0..3
  for (int i = tid; i < kSynBlocks * kPvHop; i += blockDim.x) {
    const int s = s_begin + i;
    if (s >= s_end) break;
    if (s >= len) { if (p.out_pad) out[s] = 0.f; continue; }
    const int hl = i >> 7, r = i & (kPvHop - 1);               // local hop block, offset inside it
    float acc = 0.f, env = 0.f;
#pragma unroll
    for (int d = 3; d >= 0; --d) {                             // ascending frame index, like the fold
      const int f = hl + 3 - d;                                // local frame index 0 .. kSynFrames-1
      const int j = jf + f;
      if (j >= 0 && j < To) {
        const int n = r + kPvHop * d;
        const float2 v = s_z[(f >> 1) * kPvZL + zmap(n)];
        const float w = s_win[n];
        acc += (((f & 1) ? v.y : v.x) * (1.0f / kPvN)) * w;
        env += w * w;
      }
    }
    out[s] = env > 0.f ? acc / env : 0.f;                      // env == 0 only if the caller's rate bound was wrong
  }
}

// ---- polyphase windowed-sinc resampling ------------------------------------------------------------------------
// One coefficient table per (orig, new) ratio (after division by their gcd), laid out [tap][phase]:
//   coef[tau][p] = K(p, q) for q = first(p) + tau, first(p) = floor(p * orig / new) - width, tau < ntaps = 2 width + 1,
// where K is torchaudio's kernel evaluated with ITS float32 operation order (it builds the kernel in the
// waveform's dtype; for ratios like 17959:16000 float32 rounding of t moves coefficients by ~1e-3, so the
// order matters for parity).  Taps outside the kernel's extent q in [-width, width + orig) are zero.
struct ResampleDesc {
  const float* coef;          // [ntaps][nw], followed by int32 first[nw] = floor(p * orig / nw) - width
  int orig, nw, width, ntaps; // frequencies divided by their gcd
};

__global__ void __launch_bounds__(256) resample_table_kernel(float* coef, int orig, int nw, int width, int ntaps, float base_freq, float scale) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ntaps * nw) return;
  const int tau = idx / nw, ph = idx - tau * nw;
  const int q = (int)(((int64_t)ph * orig) / nw) - width + tau;       // sample offset relative to the block start
  if (tau == 0) reinterpret_cast<int32_t*>(coef + (size_t)ntaps * nw)[ph] = q;   // first tap of this phase
  float c = 0.f;
  if (q >= -width && q < width + orig) {
    const float idxf = __fdiv_rn((float)q, (float)orig);
    float t = __fadd_rn(__fdiv_rn(-(float)ph, (float)nw), idxf);
    t = __fmul_rn(t, base_freq);
    t = fminf(fmaxf(t, -6.0f), 6.0f);
    const float wa = __fdiv_rn(__fdiv_rn(__fmul_rn(t, 3.14159274101257324f), 6.0f), 2.0f);
    const float cw = cosf(wa);
    const float window = __fmul_rn(cw, cw);
    t = __fmul_rn(t, 3.14159274101257324f);
    const float sinc = t == 0.f ? 1.0f : __fdiv_rn(sinf(t), t);
    c = __fmul_rn(sinc, __fmul_rn(window, scale));
  }
  coef[idx] = c;
}

struct ResampleParams {
  const float* in; int64_t in_stride;
  float* out; int64_t out_stride;
  int B, n_in, n_out;          // n_in: samples per input clip (upper bound when len is per clip); n_out: samples to write
  // fixed-ratio mode: one descriptor for every clip.  pitch mode: per-clip semitone -> descriptor table;
  // in_len / target per step come from the host (functions of N and the step only)
  const int32_t* steps; int step_lo, n_steps;
  const float* wav; int64_t wav_stride;        // pitch mode: source of untouched clips (step 0)
  ResampleDesc desc[kPvMaxSteps];
  int in_len[kPvMaxSteps], target[kPvMaxSteps];
};

__global__ void __launch_bounds__(256) resample_kernel(const ResampleParams p) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n_out) return;
  float* out = p.out + (size_t)b * p.out_stride;
  int d = 0;
  if (p.steps != nullptr) {
    d = p.steps[b] - p.step_lo;
    if (d < 0 || d >= p.n_steps || p.desc[d].coef == nullptr) {      // 0 semitones (or out of range): untouched
      out[i] = __ldg(p.wav + (size_t)b * p.wav_stride + i);
      return;
    }
  }
  const ResampleDesc& rd = p.desc[d];
  if (i >= p.target[d]) { out[i] = 0.f; return; }                       // _fix_waveform_shape zero padding
  const int len = p.in_len[d];
  const float* x = p.in + (size_t)b * p.in_stride;
  const int m = (int)((unsigned)i / (unsigned)rd.nw), ph = i - m * rd.nw;
  const int first = m * rd.orig + __ldg(reinterpret_cast<const int32_t*>(rd.coef + (size_t)rd.ntaps * rd.nw) + ph);
  // taps whose sample lies inside the clip: tau in [t0, t1); same ascending summation order as before
  const int t0 = first < 0 ? -first : 0;
  const int t1 = min(rd.ntaps, len - first);
  const float* c = rd.coef + ph + (size_t)t0 * rd.nw;
  const float* xp = x + first + t0;
  const int nw = rd.nw;
  float acc = 0.f;
  int n = t1 - t0;
  for (; n >= 4; n -= 4, xp += 4, c += 4 * (size_t)nw) {
    const float v0 = __ldg(xp), v1 = __ldg(xp + 1), v2 = __ldg(xp + 2), v3 = __ldg(xp + 3);
    const float c0 = __ldg(c), c1 = __ldg(c + nw), c2 = __ldg(c + 2 * (size_t)nw), c3 = __ldg(c + 3 * (size_t)nw);
    acc = fmaf(v0, c0, acc); acc = fmaf(v1, c1, acc); acc = fmaf(v2, c2, acc); acc = fmaf(v3, c3, acc);
  }
  for (; n > 0; --n, ++xp, c += nw) acc = fmaf(__ldg(xp), __ldg(c), acc);
  out[i] = acc;
}

}  // namespace wwf
