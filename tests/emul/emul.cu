// emul.cu - CPU emulation harness for the index math of the CUDA kernels.
//
// TEST INFRASTRUCTURE ONLY.  The build container has no GPU, so the __host__ __device__ task
// functions of wakeword_trainer_home_b200/csrc (radix butterflies, in-place pass tasks,
// digit-reversed positions, pair-split, overlap-save pair pass) are driven here sequentially
// - one "thread" after the other, one pass after the other - and compared with numpy by
// tests/test_emul_fft.py.  Nothing in the product loads this library.
#define WWF_EMUL_HOST 1
#include <cstdint>
#include <algorithm>
#include <cstring>
#include <functional>
#include <vector>
#include "../../wakeword_trainer_home_b200/csrc/wwf_feat.cuh"
#include "../../wakeword_trainer_home_b200/csrc/wwf_conv.cuh"
#include "../../wakeword_trainer_home_b200/csrc/wwf_pv.cuh"
#include "../../wakeword_trainer_home_b200/csrc/wwf_tables.h"

using namespace wwf;

template <int NFFT>
static void stft_pair(const float* fa, const float* fb, float* pa, float* pb) {
  using Rad = typename StftPlan<NFFT>::Rad;
  constexpr bool kNatural = stft_natural_out<NFFT>();       // the kernel's choice: last pass stores in natural order
  std::vector<float2> tw;
  build_stft_twiddles<Rad>(tw);
  std::vector<float2> z(NFFT);
  for (int j = 0; j < NFFT; ++j) z[j] = make_float2(fa[j], fb[j]);
  static_for<0, Rad::npass>([&](auto I) {
    constexpr int i = decltype(I)::value;
    constexpr int R = Rad::R(i), L = Rad::L(i);
    const float2* t = tw.data() + Rad::tw_off(i);
    if constexpr (kNatural && i == Rad::npass - 1) {
      const std::vector<float2> snap = z;                    // on the GPU: all lanes load, __syncwarp, all lanes store
      for (int u = 0; u < NFFT / R; ++u) pass_task_natural<R, Rad::R(0)>(snap.data(), z.data(), u, [] {});
    } else {
      for (int u = 0; u < NFFT / R; ++u) pass_task<R, false>(z.data(), L, u, [&](int q) { return t[q]; });
    }
  });
  for (int k = 0; k <= NFFT / 2; ++k) {
    const int kn = k == 0 ? 0 : NFFT - k;
    const int pk = kNatural ? k : Rad::pos(k), pm = kNatural ? kn : Rad::pos(kn);
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

// the two shared-memory radix-32 passes of the conv FFT, one task after the other
template <bool INV>
static void conv_passes(float2* z, const float2* tw) {
  const float2* t0 = tw + kConvTw0;
  const float2* t1 = tw + kConvTw1;
  if (!INV) {
    for (int u = 0; u < kConvM / 32; ++u) pass32_derived<false, PadMap>(z, ConvRad::L(0), u, [&](int q) { return t0[q]; });
    for (int u = 0; u < kConvM / 32; ++u) pass_task<32, false, PadMap>(z, ConvRad::L(1), u, [&](int q) { return t1[q]; });
  } else {
    for (int u = 0; u < kConvM / 32; ++u) pass_task<32, true, PadMap>(z, ConvRad::L(1), u, [&](int q) { return t1[q]; });
    for (int u = 0; u < kConvM / 32; ++u) pass32_derived<true, PadMap>(z, ConvRad::L(0), u, [&](int q) { return t0[q]; });
  }
}

static void load_block(float2* z, const float* x, int N, int start) {
  PadMap pad;
  for (int m = 0; m < kConvM; ++m) {
    const int n0 = start + 2 * m, n1 = n0 + 1;
    z[pad(m)] = make_float2((n0 >= 0 && n0 < N) ? x[n0] : 0.f, (n1 >= 0 && n1 < N) ? x[n1] : 0.f);
  }
}

static float2 w_exact_host(int k) {
  const double a = -2.0 * M_PI * (double)k / (double)kConvP;
  return make_float2((float)cos(a), (float)sin(a));
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

// the 511 run pairs + the self-paired run l = 512 (task kFusedSelfTask) + the run l = 0 must cover every run exactly once
int emul_pair_task_coverage(void) {
  std::vector<float2> tw, ftw;
  std::vector<uint16_t> fl;
  build_conv_tables(tw, fl, ftw);
  std::vector<int> seen(kRuns, 0);
  for (int t = 0; t < kFusedTasks; ++t) {
    const int l = fl[t];
    if (l <= 0 || l >= kRuns || (l == kRuns / 2) != (t == kFusedSelfTask)) return -1;
    seen[run_of(l)]++;
    if (l != kRuns / 2) seen[run_of(kRuns - l)]++;
  }
  seen[run_of(0)]++;
  for (int a = 0; a < kRuns; ++a) if (seen[a] != 1) return -2 - a;
  return 0;
}

// conv_kernel separates forward pass 1, the fused run pairs and inverse pass 1 by __syncwarp() only: every warp must
// touch exactly its own 1024 positions (its pair of 512-element sub-transforms) in all three phases.
// Returns 0, or -(1 + warp) for the first warp whose three position sets differ / overlap another warp's.
int emul_conv_warp_locality(void) {
  std::vector<float2> tw, ftw;
  std::vector<uint16_t> fl;
  build_conv_tables(tw, fl, ftw);
  std::vector<int> owner(kConvM, -1);
  for (int w = 0; w < kConvThreads / 32; ++w) {
    std::vector<char> p1(kConvM, 0), fu(kConvM, 0);
    for (int lane = 0; lane < 32; ++lane) {
      const int blk = (lane & 16) ? conv_sub_b(w) : conv_sub_a(w), j = lane & 15;     // conv_pass1
      for (int q = 0; q < 32; ++q) p1[512 * blk + j + 16 * q] = 1;
      const int t = 32 * w + lane, l = fl[t];                                          // fused_pair_task
      for (int q = 0; q < 16; ++q) { fu[16 * run_of(l) + q] = 1; fu[16 * run_of(kRuns - l) + q] = 1; }
      if (t == kFusedSelfTask) for (int q = 0; q < 16; ++q) fu[16 * run_of(0) + q] = 1;  // fused_dc_task
    }
    int n = 0;
    for (int i = 0; i < kConvM; ++i) {
      if (p1[i] != fu[i]) return -(1 + w);
      if (p1[i]) { if (owner[i] >= 0) return -(1 + w); owner[i] = w; ++n; }
    }
    if (n != 1024) return -(1 + w);
  }
  return 0;
}

// worst half-warp bank multiplicity (8-byte banks, 16 of them) of the fused tasks' run accesses
int emul_pair_bank_conflicts(void) {
  std::vector<float2> tw, ftw;
  std::vector<uint16_t> fl;
  build_conv_tables(tw, fl, ftw);
  int worst = 1;
  for (int t0 = 0; t0 < kFusedTasks; t0 += 16) {
    int ca[16] = {0}, cb[16] = {0};
    for (int i = 0; i < 16 && t0 + i < kFusedTasks; ++i) {
      ca[(17 * run_of(fl[t0 + i])) & 15]++;
      cb[(17 * run_of(kRuns - fl[t0 + i])) & 15]++;
    }
    for (int i = 0; i < 16; ++i) { if (ca[i] > worst) worst = ca[i]; if (cb[i] > worst) worst = cb[i]; }
  }
  return worst;
}

// y[0..N) = (x * h)[0..N) through the kernel's overlap-save block logic.
int emul_rir_conv(const float* x, int N, const float* h, int L, int lmax, float* y) {
  if (L > kConvP / 2 || lmax < L) return -1;
  std::vector<float2> tw, ftw;
  std::vector<uint16_t> fl;
  build_conv_tables(tw, fl, ftw);
  std::vector<float2> z(kConvSmemElems);
  std::vector<float4> spec(kSpecPerRir);
  PadMap pad;
  // spectrum (rir_spectrum_kernel)
  load_block(z.data(), h, L, 0);
  conv_passes<false>(z.data(), tw.data());
  for (int u = 0; u < kRuns; ++u) pass_task<16, false, PadMap>(z.data(), ConvRad::L(2), u, [&](int) { return make_float2(1.f, 0.f); });
  const float sc = 1.0f / (8.0f * (float)kConvM);
  auto entry = [&](int k) {
    float2 R2k, R2m;
    pair_forward(z[pad(ConvRad::pos(k))], z[pad(ConvRad::pos((kConvM - k) & (kConvM - 1)))], w_exact_host(k), R2k, R2m);
    if (k == 0) { R2k.y = 0.f; R2m.y = 0.f; }
    return make_float4(R2k.x * sc, R2k.y * sc, R2m.x * sc, R2m.y * sc);
  };
  for (int rr = 0; rr < 16; ++rr)
    for (int t = 0; t < kFusedTasks; ++t) spec[rr * 512 + t] = entry((int)fl[t] + 1024 * rr);
  for (int i = 0; i < 9; ++i) spec[kSpecSpecial + i] = entry(1024 * i);
  // blocks (conv_kernel)
  int hist = 0, valid = kConvP, nb = 1;
  if ((int64_t)N + lmax - 1 > kConvP) { hist = (lmax - 1 + 3) & ~3; valid = kConvP - hist; nb = (N + valid - 1) / valid; }
  for (int blk = 0; blk < nb; ++blk) {
    load_block(z.data(), x, N, blk * valid - hist);
    const float2* t0 = tw.data() + kConvTw0;
    const float2* t1 = tw.data() + kConvTw1;
    for (int u = 0; u < kConvM / 32; ++u) pass32_derived<false, PadMap>(z.data(), ConvRad::L(0), u, [&](int q) { return t0[q]; });
    // the warp-local middle, one warp COMPLETELY after the other (on the GPU only __syncwarp() separates its phases):
    // a warp that needed another warp's pass-1 results would read stale data here
    for (int w = 0; w < kConvThreads / 32; ++w) {
      auto sub_task = [&](int lane) { return 16 * ((lane & 16) ? conv_sub_b(w) : conv_sub_a(w)) + (lane & 15); };
      for (int lane = 0; lane < 32; ++lane)
        pass_task<32, false, PadMap>(z.data(), ConvRad::L(1), sub_task(lane), [&](int q) { return t1[q]; });
      for (int lane = 0; lane < 32; ++lane) {
        const int t = 32 * w + lane;
        const float4* sp = spec.data() + t;
        fused_pair_task(z.data(), (int)fl[t], ftw[t], [&](int r) { return sp[r * 512]; });
        if (t == kFusedSelfTask) fused_dc_task(z.data(), [&](int i) { return spec[kSpecSpecial + i]; });
      }
      for (int lane = 0; lane < 32; ++lane)
        pass_task<32, true, PadMap>(z.data(), ConvRad::L(1), sub_task(lane), [&](int q) { return t1[q]; });
    }
    for (int u = 0; u < kConvM / 32; ++u) pass32_derived<true, PadMap>(z.data(), ConvRad::L(0), u, [&](int q) { return t0[q]; });
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


// The 512-point analysis / synthesis pair of the phase-vocoder kernels (pv_stft_kernel, pv_istft_kernel):
// two real frames -> packed forward FFT in the padded scratch -> two onesided spectra (spa, spb: 257
// interleaved complex) -> Hermitian pack -> inverse passes in reverse order -> the two frames back.
int emul_pv_roundtrip(const float* fa, const float* fb, float* spa, float* spb, float* ya, float* yb) {
  std::vector<float2> tw;
  build_stft_twiddles<PvRad>(tw);
  std::vector<float2> z(kPvZL);
  const PvMap zmap;
  for (int j = 0; j < kPvN; ++j) z[zmap(j)] = make_float2(fa[j], fb[j]);
  static_for<0, PvRad::npass>([&](auto I) {
    constexpr int i = decltype(I)::value;
    constexpr int R = PvRad::R(i), L = PvRad::L(i);
    const float2* t = tw.data() + PvRad::tw_off(i);
    for (int u = 0; u < kPvN / R; ++u) pass_task<R, false, PvMap>(z.data(), L, u, [&](int q) { return t[q]; });
  });
  std::vector<float2> A(kPvK), B(kPvK);
  for (int k = 0; k < kPvK; ++k) pv_split(z[zmap(PvRad::pos(k))], z[zmap(PvRad::pos(k == 0 ? 0 : kPvN - k))], &A[k], &B[k]);
  for (int k = 0; k < kPvK; ++k) { spa[2 * k] = A[k].x; spa[2 * k + 1] = A[k].y; spb[2 * k] = B[k].x; spb[2 * k + 1] = B[k].y; }
  std::fill(z.begin(), z.end(), make_float2(0.f, 0.f));
  for (int k = 0; k < kPvK; ++k) {
    float2 a = A[k], c = B[k], zk, zm;
    if (k == 0 || k == kPvN / 2) { a.y = 0.f; c.y = 0.f; }
    pv_pack(a, c, &zk, &zm);
    z[zmap(PvRad::pos(k))] = zk;
    if (k > 0 && k < kPvN / 2) z[zmap(PvRad::pos(kPvN - k))] = zm;
  }
  static_for<0, PvRad::npass>([&](auto I) {
    constexpr int i = PvRad::npass - 1 - decltype(I)::value;
    constexpr int R = PvRad::R(i), L = PvRad::L(i);
    const float2* t = tw.data() + PvRad::tw_off(i);
    for (int u = 0; u < kPvN / R; ++u) pass_task<R, true, PvMap>(z.data(), L, u, [&](int q) { return t[q]; });
  });
  for (int j = 0; j < kPvN; ++j) { ya[j] = z[zmap(j)].x * (1.0f / kPvN); yb[j] = z[zmap(j)].y * (1.0f / kPvN); }
  return 0;
}

// Mel lane schedule (build_mel_schedule, wwf_tables.h) evaluated exactly like step 4 of frame_group_to_db: rounds of 32
// lanes, two-tap iterations, partner halves joined by lane ^ 1.  fb = dense [n_freqs][n_mels] filterbank, power =
// [n_freqs]; out_mel = [n_mels].  stats[0] = rounds, [1] = two-tap iterations, [2] = worst number of distinct bank
// residues collisions in a half-warp's first load (1 = conflict-free), [3] = number of split filters.
// Returns 0, or a negative code when the schedule is malformed (a filter without exactly one owner, ...).
int emul_mel_schedule(int n_fft, int n_freqs, int n_mels, const float* fb, const float* power, float* out_mel, int* stats) {
  std::vector<int> lo(n_mels), ofs(n_mels + 1);
  std::vector<float> w;
  for (int m = 0; m < n_mels; ++m) {
    int first = -1, last = -1;
    for (int k = 0; k < n_freqs; ++k)
      if (fb[(size_t)k * n_mels + m] != 0.f) { if (first < 0) first = k; last = k; }
    ofs[m] = (int)w.size();
    lo[m] = first < 0 ? 0 : first;
    if (first >= 0) for (int k = first; k <= last; ++k) w.push_back(fb[(size_t)k * n_mels + m]);
  }
  ofs[n_mels] = (int)w.size();
  if (w.empty()) w.push_back(0.f);
  auto zmap = [](int i) { return i; };                         // the power spectra are stored in plain bin order
  const int extent = n_fft == 400 ? n_fft : n_freqs;          // what wwf_plan_create passes (400: identity-mapped scratch)
  const MelSchedule s = build_mel_schedule(lo, ofs, w, n_fft, extent);
  std::vector<int> owners(n_mels, 0);
  int worst = 1, nsplit = 0;
  for (int r = 0; r < s.rounds; ++r) {
    float acc[32];
    for (int lane = 0; lane < 32; ++lane) {
      const int2 t = s.tasks[(size_t)r * 32 + lane];
      const int k0 = t.x & 0xffff, n = (int)((unsigned)t.x >> 16);
      const float* wr = s.w.data() + (size_t)(t.y & 0xffff) * 32 + lane;
      float a0 = 0.f, a1 = 0.f;
      int i = 0;
      for (; i + 1 < n; i += 2) { a0 = fmaf(power[k0 + i], wr[32 * i], a0); a1 = fmaf(power[k0 + i + 1], wr[32 * i + 32], a1); }
      if (i < n) a0 = fmaf(power[k0 + i], wr[32 * i], a0);
      acc[lane] = a0 + a1;
      if (n > 0 && k0 + n > n_freqs) return -1;
      // the kernel runs every lane for the round's full trip count (zero weights beyond the lane's taps): all of
      // those reads must stay inside the slots this frame group has written
      if (k0 + 2 * s.pairs[r] > extent) return -5;
    }
    for (int h = 0; h < 2; ++h) {
      // distinct ADDRESSES per bank residue: idle lanes (n = 0) read as well, but the same address as another lane is a
      // broadcast, not a conflict
      std::vector<int> seen[16];
      for (int l = 0; l < 16; ++l) {
        const int2 t = s.tasks[(size_t)r * 32 + 16 * h + l];
        const int k0 = zmap(t.x & 0xffff);
        std::vector<int>& v = seen[k0 & 15];
        if (std::find(v.begin(), v.end(), k0) == v.end()) v.push_back(k0);
        worst = std::max(worst, (int)v.size());
      }
    }
    for (int lane = 0; lane < 32; ++lane) {
      const int2 t = s.tasks[(size_t)r * 32 + lane];
      const unsigned flags = (unsigned)t.y >> 24;
      float v = acc[lane];
      if (flags & kMelPartner) {
        const int2 o = s.tasks[(size_t)r * 32 + (lane ^ 1)];
        if (((o.y >> 16) & 0xff) != ((t.y >> 16) & 0xff) || (((unsigned)o.y >> 24) & kMelOwner)) return -2;
        v += acc[lane ^ 1];
        ++nsplit;
      }
      if (flags & kMelOwner) {
        const int m = (t.y >> 16) & 0xff;
        if (m >= n_mels) return -3;
        owners[m]++;
        out_mel[m] = v;
      }
    }
  }
  for (int m = 0; m < n_mels; ++m) if (owners[m] != 1) return -4;
  stats[0] = s.rounds; stats[1] = s.iterations; stats[2] = worst; stats[3] = nsplit;
  return 0;
}

}  // extern "C"
