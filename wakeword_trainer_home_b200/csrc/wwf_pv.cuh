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
//   pv_vocoder_kernel  S                 -> V[T'][272] complex   thread per (clip, bin): the phase
//                                                                accumulation is a serial chain over time
//   pv_istft_kernel    V                 -> Y[T'][512] float     warp per frame pair: Hermitian pack,
//                                                                inverse FFT, 1/512 and window
//   pv_ola_kernel      Y                 -> y[len]               4-term overlap-add / window envelope
//   resample_kernel    y (+ coefficient table) -> out[N]
// Accuracy note: torchaudio accumulates the vocoder phase with a float32 cumsum; at ~7e4 rad one float32
// ulp is 8e-3 rad, which is why torchaudio's own float32 and float64 results differ by ~1e-3 relative.
// Here the accumulation runs in double (257 x T' adds per clip) and only the reduced angle is rounded to
// float32, so the result sits at the float64 oracle; everything else follows the float32 operation order.
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
  const float* phase_adv;      // [257] linspace(0, pi * hop, 257), float32
  const float2* tw;            // forward twiddles of Radices<16, 8, 4>
  float2* S; float2* V;        // [B][T][272], [B][Tcap][272]
  float* Y;                    // [B][Tcap][512]
  float* out; int64_t out_stride; int n_out;   // pv_ola_kernel destination: n_out samples per clip
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
__global__ void __launch_bounds__(kPvWarps * 32) pv_stft_kernel(const PvParams p) {
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
#pragma unroll 4
  for (int j = lane; j < kPvN; j += 32) {
    const float w = s_win[j];
    float re, im = 0.f;
    if (inner) {
      re = __ldg(x + ta * kPvHop - kPvN / 2 + j);
      im = __ldg(x + tb * kPvHop - kPvN / 2 + j);
    } else {
      re = __ldg(x + reflect_index(ta * kPvHop - kPvN / 2 + j, N));
      if (tb < p.T) im = __ldg(x + reflect_index(tb * kPvHop - kPvN / 2 + j, N));
    }
    z[zmap(j)] = make_float2(re * w, im * w);
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

// ---- phase vocoder: S -> V ------------------------------------------------------------------------
// One thread per (clip, bin).  Follows F.phase_vocoder line by line: angles, magnitudes and their
// interpolation in float32, the phase increments and their running sum in double.
__global__ void __launch_bounds__(288) pv_vocoder_kernel(const PvParams p) {
  const int b = blockIdx.x, k = threadIdx.x;
  const double rate = pv_clip_rate(p.rate, b);
  if (rate == 1.0 || k >= kPvK) return;
  const int T = p.T;
  const int To = pv_out_frames(T, rate, p.Tcap);
  const float2* S = p.S + (size_t)b * T * kPvPitch + k;
  float2* V = p.V + (size_t)b * p.Tcap * kPvPitch + k;
  const double pa_d = (double)__ldg(p.phase_adv + k);
  const double two_pi_d = 6.283185307179586476925286766559;
  // frames >= T are the two zero frames torchaudio pads with: angle(0) = 0, |0| = 0
  auto load = [&](int t, float& ang, float& nrm) {
    if (t < T) {
      const float2 v = S[(size_t)t * kPvPitch];
      ang = atan2f(v.y, v.x);
      nrm = hypotf(v.x, v.y);
    } else { ang = 0.f; nrm = 0.f; }
  };
  float a_prev0, n_prev0, a_prev1, n_prev1;
  int i_prev0 = 0, i_prev1 = 1;
  load(0, a_prev0, n_prev0);
  load(1, a_prev1, n_prev1);
  double acc = (double)a_prev0;                              // phase_0 = angle of frame 0
  for (int j = 0; j < To; ++j) {
    const float ts = (float)((double)j * rate);              // arange(0, T, rate) in float32
    const int i0 = (int)ts, i1 = (int)__fadd_rn(ts, 1.0f);
    const float alpha = ts - (float)i0;                      // ts % 1.0 (exact for ts >= 0)
    float a0, n0, a1, n1;
    // the frame indices advance monotonically and are uniform across the CTA: reuse what is still valid
    if (i0 == i_prev0) { a0 = a_prev0; n0 = n_prev0; }
    else if (i0 == i_prev1) { a0 = a_prev1; n0 = n_prev1; }
    else load(i0, a0, n0);
    if (i1 == i_prev1) { a1 = a_prev1; n1 = n_prev1; }
    else load(i1, a1, n1);
    i_prev0 = i0; a_prev0 = a0; n_prev0 = n0;
    i_prev1 = i1; a_prev1 = a1; n_prev1 = n1;
    const float mag = __fadd_rn(__fmul_rn(alpha, n1), __fmul_rn(__fsub_rn(1.0f, alpha), n0));
    // output j uses the phase accumulated BEFORE this step's increment (torch.cat([phase_0, phase[:-1]]))
    const double red = acc - two_pi_d * rint(acc / two_pi_d);
    float sn, cs;
    sincosf((float)red, &sn, &cs);
    V[(size_t)j * kPvPitch] = make_float2(mag * cs, mag * sn);
    // phase increment: wrap(angle1 - angle0 - phase_advance) + phase_advance.  In float32 this rounds at
    // the magnitude of phase_advance (up to 402 rad, ulp 3e-5) every step, so it is formed in double too.
    double ph = (double)a1 - (double)a0 - pa_d;
    ph -= two_pi_d * rint(ph / two_pi_d);
    acc += ph + pa_d;
  }
}

// ---- inverse STFT frames: V -> Y ----------------------------------------------------------------------
__global__ void __launch_bounds__(kPvWarps * 32) pv_istft_kernel(const PvParams p) {
  __shared__ __align__(16) float2 s_z[kPvWarps * kPvZL];
  __shared__ float2 s_tw[PvRad::tw_total];
  __shared__ float s_win[kPvN];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.y;
  const double rate = pv_clip_rate(p.rate, b);
  if (rate == 1.0) return;
  const int To = pv_out_frames(p.T, rate, p.Tcap);
  if (2 * blockIdx.x * kPvWarps >= To) return;               // CTA-uniform
  for (int i = tid; i < kPvN; i += blockDim.x) s_win[i] = __ldg(p.window + i);
  for (int i = tid; i < PvRad::tw_total; i += blockDim.x) s_tw[i] = __ldg(p.tw + i);
  __syncthreads();
  const int ja = 2 * (blockIdx.x * kPvWarps + warp), jb = ja + 1;
  if (ja >= To) return;
  const PvMap zmap;
  float2* z = s_z + warp * kPvZL;
  const float2* Va = p.V + ((size_t)b * p.Tcap + ja) * kPvPitch;
  const float2* Vb = Va + kPvPitch;
  const bool hb = jb < To;
  // Z = A + i B with Hermitian extension, written where the forward transform would have left bin k
  // (irfft ignores the imaginary parts of the DC and Nyquist bins)
  for (int k = lane; k < kPvK; k += 32) {
    float2 a = Va[k], c = hb ? Vb[k] : make_float2(0.f, 0.f);
    if (k == 0 || k == kPvN / 2) { a.y = 0.f; c.y = 0.f; }
    float2 zk, zm;
    pv_pack(a, c, &zk, &zm);
    z[zmap(PvRad::pos(k))] = zk;
    if (k > 0 && k < kPvN / 2) z[zmap(PvRad::pos(kPvN - k))] = zm;
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
  float* Ya = p.Y + ((size_t)b * p.Tcap + ja) * kPvN;
  float* Yb = Ya + kPvN;
#pragma unroll 4
  for (int j = lane; j < kPvN; j += 32) {
    const float2 v = z[zmap(j)];
    const float w = s_win[j];
    Ya[j] = (v.x * (1.0f / kPvN)) * w;
    if (hb) Yb[j] = (v.y * (1.0f / kPvN)) * w;
  }
}

// ---- overlap-add + window envelope: Y -> waveform ------------------------------------------------------------
__global__ void __launch_bounds__(256) pv_ola_kernel(const PvParams p) {
  __shared__ float s_w2[kPvN];
  const int b = blockIdx.y;
  const double rate = pv_clip_rate(p.rate, b);
  float* out = p.out + (size_t)b * p.out_stride;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (rate == 1.0) {
    if (p.out_pad && s < p.n_out && s < p.N) out[s] = __ldg(p.wav + (size_t)b * p.wav_stride + s);
    return;
  }
  for (int i = threadIdx.x; i < kPvN; i += blockDim.x) { const float w = __ldg(p.window + i); s_w2[i] = w * w; }
  __syncthreads();
  if (s >= p.n_out) return;
  const int To = pv_out_frames(p.T, rate, p.Tcap);
  const int len = pv_out_len(p.N, rate, p.Lcap);
  if (s >= len) {
    if (p.out_pad) out[s] = 0.f;
    return;
  }
  const int q = s + kPvN / 2;                                // position in the un-trimmed overlap-added signal
  int j_lo = (q - (kPvN - 1) + kPvHop - 1) / kPvHop;
  if (q - (kPvN - 1) < 0) j_lo = 0;
  int j_hi = q / kPvHop;
  if (j_hi > To - 1) j_hi = To - 1;
  const float* Y = p.Y + (size_t)b * p.Tcap * kPvN;
  float acc = 0.f, env = 0.f;
  for (int j = j_lo; j <= j_hi; ++j) {
    const int n = q - j * kPvHop;
    acc += Y[(size_t)j * kPvN + n];
    env += s_w2[n];
  }
  out[s] = env > 0.f ? acc / env : 0.f;                      // env == 0 only if the caller's rate bound was wrong
}

// ---- polyphase windowed-sinc resampling ------------------------------------------------------------------------
// One coefficient table per (orig, new) ratio (after division by their gcd), laid out [tap][phase]:
//   coef[tau][p] = K(p, q) for q = first(p) + tau, first(p) = floor(p * orig / new) - width, tau < ntaps = 2 width + 1,
// where K is torchaudio's kernel evaluated with ITS float32 operation order (it builds the kernel in the
// waveform's dtype; for ratios like 17959:16000 float32 rounding of t moves coefficients by ~1e-3, so the
// order matters for parity).  Taps outside the kernel's extent q in [-width, width + orig) are zero.
struct ResampleDesc {
  const float* coef;          // [ntaps][nw]
  int orig, nw, width, ntaps; // frequencies divided by their gcd
};

__global__ void __launch_bounds__(256) resample_table_kernel(float* coef, int orig, int nw, int width, int ntaps, float base_freq, float scale) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ntaps * nw) return;
  const int tau = idx / nw, ph = idx - tau * nw;
  const int q = (int)(((int64_t)ph * orig) / nw) - width + tau;       // sample offset relative to the block start
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
  const int m = i / rd.nw, ph = i - m * rd.nw;
  const int first = m * rd.orig + (int)(((int64_t)ph * rd.orig) / rd.nw) - rd.width;
  const float* c = rd.coef + ph;
  float acc = 0.f;
  for (int tau = 0; tau < rd.ntaps; ++tau) {
    const int q = first + tau;
    const float v = (q >= 0 && q < len) ? __ldg(x + q) : 0.f;
    acc = fmaf(v, __ldg(c + (size_t)tau * rd.nw), acc);
  }
  out[i] = acc;
}

}  // namespace wwf
