// emul.cu - CPU emulation harness for the index math of the CUDA kernels.
//
// TEST INFRASTRUCTURE ONLY.  The build container has no GPU, so the __host__ __device__ task
// functions of wakeword_trainer_home_b200/csrc (radix butterflies, in-place pass tasks,
// digit-reversed positions, pair-split, overlap-save pair pass) are driven here sequentially
// - one "thread" after the other, one pass after the other - and compared with numpy by
// tests/test_emul_fft.py.  Nothing in the product loads this library.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../wakeword_trainer_home_b200/csrc/wwf_feat.cuh"
#include "../../wakeword_trainer_home_b200/csrc/wwf_conv.cuh"
#include "../../wakeword_trainer_home_b200/csrc/wwf_tables.h"

using namespace wwf;

template <int NFFT>
static void stft_pair(const float* fa, const float* fb, float* pa, float* pb) {
  using Rad = typename StftPlan<NFFT>::Rad;
  std::vector<float2> tw;
  build_stft_twiddles<Rad>(tw);
  std::vector<float2> z(NFFT);
  for (int j = 0; j < NFFT; ++j) z[j] = make_float2(fa[j], fb[j]);
  static_for<0, Rad::npass>([&](auto I) {
    constexpr int i = decltype(I)::value;
    constexpr int R = Rad::R(i), L = Rad::L(i);
    const float2* t = tw.data() + Rad::tw_off(i);
    for (int u = 0; u < NFFT / R; ++u) pass_task<R, false>(z.data(), L, u, [&](int q) { return t[q]; });
  });
  for (int k = 0; k <= NFFT / 2; ++k) {
    const int pk = Rad::pos(k), pm = Rad::pos(k == 0 ? 0 : NFFT - k);
    const float2 pw = pair_split_power(z[pk], z[pm]);
    pa[k] = pw.x;
    pb[k] = pw.y;
  }
}

// forward complex FFT through the same passes; natural-order output
template <int NFFT>
static void cfft(const float* in, float* out) {
  using Rad = typename StftPlan<NFFT>::Rad;
  std::vector<float2> tw;
  build_stft_twiddles<Rad>(tw);
  std::vector<float2> z(NFFT);
  for (int j = 0; j < NFFT; ++j) z[j] = make_float2(in[2 * j], in[2 * j + 1]);
  static_for<0, Rad::npass>([&](auto I) {
    constexpr int i = decltype(I)::value;
    constexpr int R = Rad::R(i), L = Rad::L(i);
    const float2* t = tw.data() + Rad::tw_off(i);
    for (int u = 0; u < NFFT / R; ++u) pass_task<R, false>(z.data(), L, u, [&](int q) { return t[q]; });
  });
  for (int k = 0; k < NFFT; ++k) { out[2 * k] = z[Rad::pos(k)].x; out[2 * k + 1] = z[Rad::pos(k)].y; }
}

template <bool INV>
static void conv_passes(float2* z, const float2* tw) {
  auto pass4 = [&]() {
    for (int u = 0; u < kConvM / 4; ++u) {
      const float2 w1 = tw[kConvTw0 + u];
      const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1);
      pass_task<4, INV, PadMap>(z, ConvRad::L(0), u, [&](int q) { const int r = q / ConvRad::S(0); return r == 0 ? w1 : (r == 1 ? w2 : w3); });
    }
  };
  auto pass16 = [&](int L, const float2* t) {
    for (int u = 0; u < kConvM / 16; ++u) pass_task<16, INV, PadMap>(z, L, u, [&](int q) { return t[q]; });
  };
  if (!INV) { pass4(); pass16(ConvRad::L(1), tw + kConvTw1); pass16(ConvRad::L(2), tw + kConvTw2); pass16(ConvRad::L(3), tw); }
  else { pass16(ConvRad::L(3), tw); pass16(ConvRad::L(2), tw + kConvTw2); pass16(ConvRad::L(1), tw + kConvTw1); pass4(); }
}

static void load_block(float2* z, const float* x, int N, int start) {
  PadMap pad;
  for (int m = 0; m < kConvM; ++m) {
    const int n0 = start + 2 * m, n1 = n0 + 1;
    z[pad(m)] = make_float2((n0 >= 0 && n0 < N) ? x[n0] : 0.f, (n1 >= 0 && n1 < N) ? x[n1] : 0.f);
  }
}

extern "C" {

int emul_stft_pair(int nfft, const float* fa, const float* fb, float* pa, float* pb) {
  switch (nfft) {
    case 256: stft_pair<256>(fa, fb, pa, pb); return 0;
    case 400: stft_pair<400>(fa, fb, pa, pb); return 0;
    case 512: stft_pair<512>(fa, fb, pa, pb); return 0;
    case 1024: stft_pair<1024>(fa, fb, pa, pb); return 0;
    case 2048: stft_pair<2048>(fa, fb, pa, pb); return 0;
  }
  return -1;
}

int emul_cfft(int nfft, const float* in, float* out) {
  switch (nfft) {
    case 256: cfft<256>(in, out); return 0;
    case 400: cfft<400>(in, out); return 0;
    case 512: cfft<512>(in, out); return 0;
    case 1024: cfft<1024>(in, out); return 0;
    case 2048: cfft<2048>(in, out); return 0;
  }
  return -1;
}

// every pair task must cover each k in [0, M/2] exactly once
int emul_pair_task_coverage(void) {
  std::vector<int> seen(kConvM / 2 + 1, 0);
  for (int v = 0; v <= kConvPairTasks; ++v) {
    const int k = pair_task_k(v);
    if (k < 0 || k > kConvM / 2) return -1;
    seen[k]++;
  }
  for (int k = 0; k <= kConvM / 2; ++k) if (seen[k] != 1) return -2 - k;
  return 0;
}

// worst half-warp bank multiplicity (8-byte banks, 16 of them) of the pair pass at pos(k) / pos(M-k)
int emul_pair_bank_conflicts(void) {
  PadMap pad;
  int worst = 1;
  for (int v0 = 0; v0 < kConvPairTasks; v0 += 16) {
    int ca[16] = {0}, cb[16] = {0};
    for (int l = 0; l < 16; ++l) {
      const int k = pair_task_k(v0 + l);
      ca[pad(ConvRad::pos(k)) & 15]++;
      cb[pad(ConvRad::pos((kConvM - k) & (kConvM - 1))) & 15]++;
    }
    for (int i = 0; i < 16; ++i) { if (ca[i] > worst) worst = ca[i]; if (cb[i] > worst) worst = cb[i]; }
  }
  return worst;
}

// y[0..N) = (x * h)[0..N) through the kernel's overlap-save block logic.
int emul_rir_conv(const float* x, int N, const float* h, int L, int lmax, float* y) {
  if (L > kConvP / 2 || lmax < L) return -1;
  std::vector<float2> tw, twp;
  build_conv_twiddles(tw, twp);
  std::vector<float2> z(kConvSmemElems);
  std::vector<float4> spec(kConvPairTasks + 1);
  PadMap pad;
  // spectrum (rir_spectrum_kernel)
  load_block(z.data(), h, L, 0);
  conv_passes<false>(z.data(), tw.data());
  const float sc = 1.0f / (8.0f * (float)kConvM);
  for (int v = 0; v <= kConvPairTasks; ++v) {
    const int k = pair_task_k(v);
    const int pk = pad(ConvRad::pos(k)), pm = pad(ConvRad::pos((kConvM - k) & (kConvM - 1)));
    float2 R2k, R2m;
    pair_forward(z[pk], z[pm], twp[v], R2k, R2m);
    if (k == 0) { R2k.y = 0.f; R2m.y = 0.f; }
    spec[v] = make_float4(R2k.x * sc, R2k.y * sc, R2m.x * sc, R2m.y * sc);
  }
  // blocks (conv_kernel)
  int hist = 0, valid = kConvP, nb = 1;
  if ((int64_t)N + lmax - 1 > kConvP) { hist = (lmax - 1 + 3) & ~3; valid = kConvP - hist; nb = (N + valid - 1) / valid; }
  for (int blk = 0; blk < nb; ++blk) {
    load_block(z.data(), x, N, blk * valid - hist);
    conv_passes<false>(z.data(), tw.data());
    for (int v = 0; v <= kConvPairTasks; ++v) {
      const int k = pair_task_k(v);
      const int pk = pad(ConvRad::pos(k)), pm = pad(ConvRad::pos((kConvM - k) & (kConvM - 1)));
      const float4 hh = spec[v];
      float2 R2k, R2m, Zk, Zm;
      pair_forward(z[pk], z[pm], twp[v], R2k, R2m);
      pair_inverse(cmul(R2k, make_float2(hh.x, hh.y)), cmul(R2m, make_float2(hh.z, hh.w)), twp[v], Zk, Zm);
      z[pm] = Zm;
      z[pk] = Zk;
    }
    conv_passes<true>(z.data(), tw.data());
    for (int m = 0; m < kConvM; ++m) {
      for (int c = 0; c < 2; ++c) {
        const int i = 2 * m + c;
        if (i < hist) continue;
        const int n = blk * valid + i - hist;
        if (n < N) y[n] = c == 0 ? z[pad(m)].x : z[pad(m)].y;
      }
    }
  }
  return nb;
}

int emul_reflect_index(int i, int N) { return reflect_index(i, N); }

int emul_mel_fbanks(int n_freqs, float f_min, float f_max, int n_mels, int sr, float* out) {
  std::vector<float> fb = mel_fbanks32(n_freqs, f_min, f_max, n_mels, sr);
  memcpy(out, fb.data(), fb.size() * sizeof(float));
  return 0;
}

}  // extern "C"
