// wwf_tables.h - host-side builders of the constant tables the kernels read (FFT twiddles,
// mel filterbank).  Shared by wwfeat.cu (plan creation) and tests/emul (CPU index-math checks).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
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
  // task t = 32 w + lane owns the runs l and 1024 - l of warp w's sub-transform pair (conv_fused_l, wwf_conv.cuh): the
  // run index of l = d0 + 32 d1 is 32 d0 + d1, so the 16 lanes of a half-warp get 16 distinct (run mod 16) for both
  // runs -> conflict-free 8-byte accesses
  fused_l.resize(kFusedTasks);
  fused_tw.resize(kFusedTasks);
  for (int t = 0; t < kFusedTasks; ++t) {
    const int l = conv_fused_l(t);
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


// ------------------------------------------------------------------------------------------------------------------
// Lane schedule of the sparse mel projection.
//
// The filterbank has at most two non-zeros per FFT bin, so each filter is a short run of consecutive bins
// (2 .. ~30 taps).  One lane per filter in filter order leaves the warp waiting for its widest filter (21 two-tap
// iterations for 40 mels at n_fft 400, 25 of 32 lanes active on average) and reads bins lo[m] + i whose bank
// residues collide (1.7 - 2.4 wavefronts per load on every configuration, ncu + tools/mel_sched_sim.cu).  The schedule
// built here instead
//   * cuts filters wider than a threshold W into two halves on ADJACENT lanes (one __shfl_xor joins them),
//   * sorts the resulting tasks by length and fills rounds of 32 lanes, so the lanes of a round finish together,
//   * makes every half-warp CONFLICT-FREE by construction: the power spectra are stored in plain bin order (8-byte
//     elements: 16 banks per half-warp) and all lanes advance one bin per tap, so two lanes collide on every tap or
//     on none, depending only on (first bin mod 16).  A task whose residue is taken starts d bins EARLIER with d
//     leading zero weights (first bin k0 - d, n + d taps): tasks are placed longest first, each taking the half-warp
//     and the smallest d that keep the round short,
//   * stores the weights interleaved per round (w[(base + i) * 32 + lane]): conflict-free as well.
// W is chosen by minimising the modelled issue cost (two-tap iterations of each round's longest task + a fixed
// per-round overhead).  A task is an int2: x = first bin | ntaps << 16,  y = weight base (in rows of 32 floats)
// | filter << 16 | flags << 24.
// ------------------------------------------------------------------------------------------------------------------
struct MelSchedule {
  int rounds = 0;
  std::vector<int2> tasks;        // [rounds * 32]
  std::vector<float> w;           // interleaved weights, rows of 32
  std::vector<int> pairs;         // [rounds] two-tap iterations of each round (every lane runs them all: the weight
                                  // rows beyond a lane's own taps are zero)
  int iterations = 0;             // two-tap iterations summed over the rounds (for reports)
  int conflicts = 0;              // half-warp tap rows that still hit a bank twice (0 unless a residue could not be freed)
};

// lo[m] / ofs[m] / w: CSR rows of the filterbank (first bin, offsets into w); n_fft: elements of a transform's scratch
// read_extent: the scratch slots [0, read_extent) are safe for a lane to read and multiply by a zero weight: the
// n_fft / 2 + 1 power values in general; all n_fft slots for the identity-mapped plan (n_fft 400), where this very
// frame group's first FFT pass has written every slot (finite whenever the pair's samples are - and if they are not,
// both frames of the pair are non-finite anyway).
inline MelSchedule build_mel_schedule(const std::vector<int>& lo, const std::vector<int>& ofs, const std::vector<float>& w, int n_fft,
                                      int read_extent = 0) {
  if (read_extent <= 0) read_extent = n_fft / 2 + 1;
  const int M = (int)lo.size();
  struct Task { int m, k0, o, n, flags, d; };                  // d: leading zero-weight taps
  struct Unit { Task a, b; bool pair; int len() const { return pair ? std::max(a.n, b.n) : a.n; } int lanes() const { return pair ? 2 : 1; } };
  struct Placed { int lane; Task t; };
  struct Round { std::vector<Placed> placed; int mx = 0, conflicts = 0; };
  auto layout = [&](int W) {
    std::vector<Unit> u;
    for (int m = 0; m < M; ++m) {
      const int n = ofs[m + 1] - ofs[m];
      if (n > W && n >= 2) {
        const int n1 = (n + 1) / 2;
        u.push_back({{m, lo[m], ofs[m], n1, kMelOwner | kMelPartner, 0}, {m, lo[m] + n1, ofs[m] + n1, n - n1, 0, 0}, true});
      } else {
        u.push_back({{m, lo[m], ofs[m], n, kMelOwner, 0}, {}, false});
      }
    }
    std::stable_sort(u.begin(), u.end(), [](const Unit& x, const Unit& y) { return x.len() > y.len(); });
    std::vector<Round> rounds;
    int used = 32;
    int lanes_used[2] = {0, 0}, pair_cur[2] = {0, 0}, single_cur[2] = {15, 15};
    bool taken[2][16];
    // cheapest shift d <= min(k0, 15) of a task in half h: a shift beyond the round's current length costs issue slots
    // for the whole warp (7 per tap), a residue that is already taken costs a second wavefront on each of the
    // task's taps (3 per tap); `also` = residue of the unit's other half
    auto shift_for = [&](int h, const Task& t, int also, int mx_now, int* cost_out) {
      int best_d = 0, best_c = 1 << 30;
      for (int d = 0; d <= std::min(t.k0, 15); ++d) {
        const int r = (t.k0 - d) & 15;
        const int c = 7 * std::max(0, t.n + d - mx_now) + ((taken[h][r] || r == also) ? 3 * t.n : 0);
        if (c < best_c) { best_c = c; best_d = d; }
      }
      *cost_out = best_c;
      return best_d;
    };
    for (const Unit& x : u) {
      if (used + x.lanes() > 32 || (lanes_used[0] + x.lanes() > 16 && lanes_used[1] + x.lanes() > 16)) {
        rounds.emplace_back();
        used = 0;
        lanes_used[0] = lanes_used[1] = 0; pair_cur[0] = pair_cur[1] = 0; single_cur[0] = single_cur[1] = 15;
        for (auto& tr : taken) for (bool& v : tr) v = false;
      }
      Round& r = rounds.back();
      const int mx_now = std::max(r.mx, x.len());               // (units arrive longest first)
      // the half-warp (with room) where the unit is cheapest; ties: the emptier half
      int best_h = -1, best_cost = 1 << 30, best_da = 0, best_db = 0;
      for (int h = 0; h < 2; ++h) {
        if (lanes_used[h] + x.lanes() > 16) continue;
        int ca = 0, cb = 0, db = 0;
        const int da = shift_for(h, x.a, -1, mx_now, &ca);
        if (x.pair) db = shift_for(h, x.b, (x.a.k0 - da) & 15, mx_now, &cb);
        if (ca + cb < best_cost || (ca + cb == best_cost && lanes_used[h] < lanes_used[best_h])) {
          best_h = h; best_cost = ca + cb; best_da = da; best_db = db;
        }
      }
      const int h = best_h;
      Task ta = x.a, tb = x.b;
      ta.d = best_da; tb.d = best_db;
      if (taken[h][(ta.k0 - ta.d) & 15]) r.conflicts += ta.n;
      taken[h][(ta.k0 - ta.d) & 15] = true;
      r.mx = std::max(r.mx, ta.n + ta.d);
      if (x.pair) {
        if (taken[h][(tb.k0 - tb.d) & 15]) r.conflicts += tb.n;
        taken[h][(tb.k0 - tb.d) & 15] = true;
        r.mx = std::max(r.mx, tb.n + tb.d);
        const int l = 16 * h + pair_cur[h];
        pair_cur[h] += 2;
        r.placed.push_back({l, ta});
        r.placed.push_back({l + 1, tb});
      } else {
        r.placed.push_back({16 * h + single_cur[h]--, ta});
      }
      lanes_used[h] += x.lanes();
      used += x.lanes();
    }
    return rounds;
  };
  auto cost = [](const std::vector<Round>& r) {                  // issue slots: two-tap iterations + per-round overhead
    int c = 0;
    for (const Round& rr : r) c += 14 * ((rr.mx + 1) / 2) + 45 + 8 * rr.conflicts;
    return c;
  };
  int maxw = 1;
  for (int m = 0; m < M; ++m) maxw = std::max(maxw, ofs[m + 1] - ofs[m]);
  int bestW = maxw, bestC = -1;
  for (int W = maxw; W >= 1; --W) {                              // ties: the least splitting
    const int c = cost(layout(W));
    if (bestC < 0 || c < bestC) { bestC = c; bestW = W; }
  }
  const std::vector<Round> rounds = layout(bestW);
  MelSchedule s;
  s.rounds = (int)rounds.size();
  s.tasks.assign((size_t)s.rounds * 32, make_int2(0, 0xff << 16));
  int wbase = 0;
  for (int r = 0; r < s.rounds; ++r) {
    const int rows = 2 * ((rounds[r].mx + 1) / 2);              // every lane runs rows / 2 two-tap iterations
    s.w.resize((size_t)(wbase + rows) * 32, 0.f);
    for (int lane = 0; lane < 32; ++lane) s.tasks[(size_t)r * 32 + lane] = make_int2(0, wbase | (0xff << 16));   // idle: zero weights
    for (const Placed& pl : rounds[r].placed) {
      const Task& t = pl.t;
      // a lane reads bins k0 .. k0 + rows - 1 whatever its own tap count: keep that inside read_extent by starting
      // earlier with more leading zero weights.  (Anything beyond is NOT safe to multiply by a zero weight: the padded
      // index maps of the power-of-two transforms leave slots of the scratch that no pass ever writes, and whatever
      // an earlier kernel left there - a NaN, say - would turn 0 * x into NaN.  tests: poisoned shared memory.)
      int k0 = t.k0 - t.d, d = t.d;
      if (k0 + rows > read_extent) { const int sh = std::min(k0, k0 + rows - read_extent); k0 -= sh; d += sh; }
      for (int i = 0; i < t.n; ++i) s.w[(size_t)(wbase + d + i) * 32 + pl.lane] = w[t.o + i];   // rows < d and >= d + n stay 0
      s.tasks[(size_t)r * 32 + pl.lane] = make_int2(k0 | ((t.n + d) << 16), wbase | (t.m << 16) | (t.flags << 24));
    }
    // idle lanes run the round's taps too (zero weights): they read wherever an ACTIVE lane of their half-warp reads
    // (same addresses = a broadcast), not bin 0 - which shares a bank with every active lane whose first bin is a
    // multiple of 16 and cost that half-warp a second wavefront per tap (11 % of the mel reads of n_fft 400 / 40 mels)
    for (int h = 0; h < 2; ++h) {
      int k0_active = -1;
      for (int l = 0; l < 16 && k0_active < 0; ++l) {
        const int2 t = s.tasks[(size_t)r * 32 + 16 * h + l];
        if (((unsigned)t.x >> 16) > 0) k0_active = t.x & 0xffff;
      }
      if (k0_active < 0) continue;
      for (int l = 0; l < 16; ++l) {
        int2& t = s.tasks[(size_t)r * 32 + 16 * h + l];
        if (((unsigned)t.x >> 16) == 0) t.x = k0_active;         // n stays 0
      }
    }
    wbase += rows;
    s.pairs.push_back(rows / 2);
    s.iterations += rows / 2;
    s.conflicts += rounds[r].conflicts;
  }
  if (s.w.empty()) s.w.push_back(0.f);
  return s;
}

// ------------------------------------------------------------------------------------------------------------------
// DCT matrix as split-TF32 mma.sync.m16n8k8 B fragments (feat_epilogue_mma_kernel, wwf_feat.cuh).
// dct: [n_mels][n_mfcc] float32 (torchaudio's create_dct values).  Output: uint4 per (k-step ks, n-tile j, lane):
//   x = hi(D[8 ks + t][8 j + g]), y = hi(D[8 ks + t + 4][8 j + g]), z / w = the lo halves, g = lane / 4, t = lane % 4,
// zero beyond the matrix.  hi(x) = x rounded to TF32 (round to nearest, ties away: cvt.rna), lo(x) = hi(x - hi(x)).
// ------------------------------------------------------------------------------------------------------------------
inline uint32_t tf32_rna_host(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return u;             // inf / nan pass through
  u += 0x1000u;                                                // half an ulp of the 13 dropped bits, in magnitude
  return u & 0xffffe000u;
}
inline std::vector<uint4> build_dct_fragments(const std::vector<float>& dct, int n_mels, int n_mfcc) {
  const int k8 = (n_mels + 7) & ~7, c8 = (n_mfcc + 7) & ~7, ksteps = k8 / 8, ntiles = c8 / 8;
  std::vector<uint4> out((size_t)ksteps * ntiles * 32);
  auto at = [&](int m, int c) { return (m < n_mels && c < n_mfcc) ? dct[(size_t)m * n_mfcc + c] : 0.f; };
  auto split = [](float v, uint32_t& hi, uint32_t& lo) {
    hi = tf32_rna_host(v);
    float h;
    memcpy(&h, &hi, 4);
    lo = tf32_rna_host(v - h);
  };
  for (int ks = 0; ks < ksteps; ++ks)
    for (int j = 0; j < ntiles; ++j)
      for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, t = lane & 3;
        uint4 f;
        split(at(8 * ks + t, 8 * j + g), f.x, f.z);
        split(at(8 * ks + t + 4, 8 * j + g), f.y, f.w);
        out[((size_t)ks * ntiles + j) * 32 + lane] = f;
      }
  return out;
}

// column sums of the DCT matrix (double accumulation, rounded once), padded with zeros to a multiple of 8 columns:
// the un-centring term of feat_epilogue_mma_kernel
inline std::vector<float> dct_column_sums(const std::vector<float>& dct, int n_mels, int n_mfcc) {
  std::vector<float> out((size_t)((n_mfcc + 7) & ~7), 0.f);
  for (int c = 0; c < n_mfcc; ++c) {
    double s = 0.0;
    for (int m = 0; m < n_mels; ++m) s += (double)dct[(size_t)m * n_mfcc + c];
    out[c] = (float)s;
  }
  return out;
}


}  // namespace wwf
