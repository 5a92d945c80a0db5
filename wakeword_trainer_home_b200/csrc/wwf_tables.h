// wwf_tables.h - host-side builders of the constant tables the kernels read (FFT twiddles,
// mel filterbank).  Shared by wwfeat.cu (plan creation) and tests/emul (CPU index-math checks).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>
#include "wwf_conv.cuh"
#include "wwf_fft.cuh"

namespace wwf {

template <class Rad>
inline void build_stft_twiddles(std::vector<float2>& tw) {
  tw.assign(Rad::tw_total > 0 ? Rad::tw_total : 1, make_float2(1.f, 0.f));
  for (int i = 0; i < Rad::npass; ++i) {
    const int R = Rad::R(i), L = Rad::L(i), s = Rad::S(i);
    if (s <= 1) continue;
    for (int r = 1; r < R; ++r)
      for (int j = 0; j < s; ++j) {
        const double a = -2.0 * M_PI * (double)((long long)j * r % L) / (double)L;
        tw[Rad::tw_off(i) + (r - 1) * s + j] = make_float2((float)cos(a), (float)sin(a));
      }
  }
}


// Tables of the overlap-save FFT (wwf_conv.cuh): pass twiddles, fused-task order and w_P^l.
inline void build_conv_tables(std::vector<float2>& tw, std::vector<uint16_t>& fused_l, std::vector<float2>& fused_tw) {
  auto W = [](long long e, long long n) {
    const double a = -2.0 * M_PI * (double)(e % n) / (double)n;
    return make_float2((float)cos(a), (float)sin(a));
  };
  tw.assign(kConvTwTotal, make_float2(1.f, 0.f));
  const int s0 = ConvRad::S(0), s1 = ConvRad::S(1);
  for (int b = 0; b < 5; ++b)
    for (int j = 0; j < s0; ++j) tw[kConvTw0 + b * s0 + j] = W((long long)j << b, ConvRad::L(0));
  for (int r = 1; r < 32; ++r)
    for (int j = 0; j < s1; ++j) tw[kConvTw1 + (r - 1) * s1 + j] = W((long long)j * r, ConvRad::L(1));
  // task t owns runs l and 1024 - l, l = d0 + 32 d1 with d1 = (t+1) & 15, d0 = (t+1) >> 4: the 16 lanes of
  // a half-warp get 16 distinct (run mod 16) = d1 for both runs -> conflict-free 8-byte accesses
  fused_l.resize(kFusedTasks);
  fused_tw.resize(kFusedTasks);
  for (int t = 0; t < kFusedTasks; ++t) {
    const int tt = t + 1, l = (tt >> 4) + 32 * (tt & 15);
    fused_l[t] = (uint16_t)l;
    fused_tw[t] = W(l, kConvP);
  }
}

// torch.linspace(start, end, steps) in float32 (forward from start / backward from end)
inline std::vector<float> linspace32(float start, float end, int steps) {
  std::vector<float> v(steps);
  const float step = (end - start) / (float)(steps - 1);
  const int half = steps / 2;
  for (int i = 0; i < steps; ++i) v[i] = i < half ? start + step * (float)i : end - step * (float)(steps - 1 - i);
  return v;
}

// melscale_fbanks(n_freqs, f_min, f_max, n_mels, sr, norm=None, mel_scale="htk") in float32,
// same operation order as TA/functional/functional.py:490-587.
inline std::vector<float> mel_fbanks32(int n_freqs, float f_min, float f_max, int n_mels, int sr) {
  std::vector<float> all = linspace32(0.f, (float)(sr / 2), n_freqs);
  const double m_min = 2595.0 * log10(1.0 + (double)f_min / 700.0), m_max = 2595.0 * log10(1.0 + (double)f_max / 700.0);
  std::vector<float> m_pts = linspace32((float)m_min, (float)m_max, n_mels + 2), f_pts(n_mels + 2);
  for (int i = 0; i < n_mels + 2; ++i) f_pts[i] = 700.0f * (powf(10.0f, m_pts[i] / 2595.0f) - 1.0f);
  std::vector<float> fb((size_t)n_freqs * n_mels);
  for (int k = 0; k < n_freqs; ++k)
    for (int m = 0; m < n_mels; ++m) {
      const float down = (-1.0f * (f_pts[m] - all[k])) / (f_pts[m + 1] - f_pts[m]);
      const float up = (f_pts[m + 2] - all[k]) / (f_pts[m + 2] - f_pts[m + 1]);
      fb[(size_t)k * n_mels + m] = fmaxf(0.f, fminf(down, up));
    }
  return fb;
}


}  // namespace wwf
