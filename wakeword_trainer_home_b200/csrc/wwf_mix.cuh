// wwf_mix.cuh - background-noise mixing at a target SNR (oracle:
// torchaudio F.add_noise, TA/functional/functional.py:2317-2382): the device view of the registered noise
// bank, O(1) segment energies from its prefix table, the mix scale, and the per-clip record (ClipMix) that the
// feat_prep_kernel leaves for the flat frames kernel (wwf_feat.cuh).
#pragma once
#include <stdint.h>
#include "wwf_fft.cuh"

namespace wwf {

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

// What the flat frames kernel needs to know about one clip's noise mix; written once per clip by feat_prep_kernel
// (a reverberated clip's scale by conv_kernel, from en / snr, when the records are made before the reverb).
struct alignas(16) ClipMix {
  float scale;          // F.add_noise's scale
  int has_noise;        // 0: the clip is not mixed
  int noff, nlen;       // start offset inside the noise clip (already wrapped), its length
  long long nz_off;     // sample offset of the noise clip inside the bank
  float en, snr;        // energy of the noise segment, target SNR (what the scale is made of besides the clip's energy)
};

// A clip is reverberated iff its RIR index addresses the registered bank; the SAME predicate in the producer
// (conv_kernel) and in every consumer, so an out-of-range index means "dry", never stale workspace rows.
__host__ __device__ __forceinline__ bool rir_in_range(int r, int n_rir) { return r >= 0 && r < n_rir; }

#if defined(__CUDACC__)
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

__device__ __forceinline__ bool clip_has_rev(const float* rev, const int32_t* rir_idx, int n_rir, int b) {
  return rev != nullptr && rir_idx != nullptr && rir_in_range(__ldg(rir_idx + b), n_rir);
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

// Loads of one segment [a, b) of a noise clip, 0 <= a <= b <= len: the prefix-table entries of its whole 128-blocks
// and the (< 128-sample) edges - every edge sample a lane needs (at most 4 at either end, or 8 when the segment lies
// inside two blocks) is requested before the first one is used: as rolled load-then-accumulate loops the edges were
// up to eight SERIAL memory latencies.
struct SegLoads { float v[8]; double p_hi, p_lo; };
__device__ __forceinline__ SegLoads seg_issue(const ClipNoise& c, int a, int b) {
  const int lane = threadIdx.x & 31;
  const int lo = (a + kNoiseBlk - 1) / kNoiseBlk, hi = b / kNoiseBlk;
  SegLoads s;
  s.p_hi = 0.0; s.p_lo = 0.0;
  if (lo <= hi) {
    s.p_hi = c.P[hi]; s.p_lo = c.P[lo];
    const int e0 = lo * kNoiseBlk, t0 = hi * kNoiseBlk;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = a + lane + 32 * i, r = t0 + lane + 32 * i;
      s.v[i] = q < e0 ? __ldg(c.nz + q) : 0.f;
      s.v[4 + i] = r < b ? __ldg(c.nz + r) : 0.f;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int q = a + lane + 32 * i;
      s.v[i] = q < b ? __ldg(c.nz + q) : 0.f;
    }
  }
  return s;
}
// sum of nz[q]^2 over the segment: whole blocks from the table + the edges in a fixed order (missing samples
// contribute fmaf(0, 0, e) = e).  Executed by a full warp; every lane returns the result.
__device__ __forceinline__ double seg_energy(const SegLoads& s) {
  float e = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) e = fmaf(s.v[i], s.v[i], e);
  return (s.p_hi - s.p_lo) + (double)warp_sum(e);
}

// Energy of the N-sample noise segment nz[(off + i) mod len], i < N: the stretch up to the end of the clip, whole
// passes over the clip, and the remainder from its start - the loads of BOTH partial stretches are in flight together
// (a segment that wraps - every one when the noise clip is as long as the speech clip - was two dependent rounds).
static __device__ __noinline__ float warp_noise_energy(const ClipNoise& c, int N) {
  const int first = min(N, c.len - c.off);
  int rem = N - first;
  const int loops = rem > 0 ? rem / c.len : 0;
  rem -= loops * c.len;
  const SegLoads sa = seg_issue(c, c.off, c.off + first);
  const SegLoads sb = seg_issue(c, 0, rem);                    // (rem = 0: no loads, contributes 0)
  const double whole = loops > 0 ? c.P[(c.len + kNoiseBlk - 1) / kNoiseBlk] : 0.0;
  double e = seg_energy(sa);
  if (loops > 0) e += (double)loops * whole;
  if (rem > 0) e += seg_energy(sb);
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

#endif  // __CUDACC__

}  // namespace wwf
