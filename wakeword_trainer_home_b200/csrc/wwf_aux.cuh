// wwf_aux.cuh - the two stand-alone entry points of the reference surface that are not the
// fused feature kernel: AudioAugmentation.__call__'s noise mix writing a waveform (A1) and
// SpecAugment.__call__ on an existing feature tensor (A9).  SURVEY.md section 8a.
#pragma once
#include "wwf_feat.cuh"

namespace wwf {

struct MixParams {
  const float* wav; int64_t wav_stride;
  const float* rev; int64_t rev_stride;     // reverberated clips or nullptr
  const int32_t* rir_idx; const int32_t* noise_idx; const int64_t* noise_off; const float* snr_db;
  NoiseBankDev noise;
  const float* es_part; int es_nb;
  float* out; int64_t out_stride;
  int B, N, n_rir;
};

// One CTA per clip: y = src + scale * noise, scale from the two clip energies
// (F.add_noise, TA/functional/functional.py:2374-2382); plain copy when the clip has no noise.
__global__ void __launch_bounds__(512) mix_kernel(const MixParams p) {
  __shared__ float red[64];
  const int b = blockIdx.x;
  const bool has_rev = clip_has_rev(p.rev, p.rir_idx, p.n_rir, b);
  const float* x = has_rev ? p.rev + (size_t)b * p.rev_stride : p.wav + (size_t)b * p.wav_stride;
  float* y = p.out + (size_t)b * p.out_stride;
  const ClipNoise cn = resolve_noise(p.noise, p.noise_idx, p.noise_off, b);
  const float* nz = cn.nz;
  const int noff = cn.off, nlen = cn.len;
  const float scale = clip_mix_scale(cn, x, p.N, has_rev, p.es_part, p.es_nb, b, p.snr_db, red);
  if (nz != nullptr) {
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) y[i] = fmaf(scale, noise_at(nz, noff, nlen, i), x[i]);
  } else if (x != y) {
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) y[i] = x[i];
  }
}

// Peak normalisation of each clip, y = x / max|x| (unchanged if the clip is all zero): what the reference
// does to a chunk right before FeatureExtractor (src/evaluation/inference.py:189-191).  IEEE division,
// so results are bit-identical to numpy's float32 arithmetic.  One CTA per clip; in place allowed.
__global__ void __launch_bounds__(512) peak_normalize_kernel(const float* __restrict__ wav, int64_t wav_stride, float* out,
                                                             int64_t out_stride, int N) {
  __shared__ float red[32];
  const float* x = wav + (size_t)blockIdx.x * wav_stride;
  float* y = out + (size_t)blockIdx.x * out_stride;
  float m = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) m = fmaxf(m, fabsf(x[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = (threadIdx.x & 31) < (blockDim.x >> 5) ? red[threadIdx.x & 31] : 0.f;
  m = warp_max(m);
  if (m > 0.f) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) y[i] = x[i] / m;
  } else if (x != y) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) y[i] = x[i];
  }
}

// Registration helper: sums[j] = sum of squares of block j (kNoiseBlk samples, last one partial)
// of one noise clip, in double.  One warp per block.
__global__ void __launch_bounds__(256) noise_block_sums_kernel(const float* __restrict__ data, int64_t len, double* __restrict__ sums) {
  const int64_t nblk = (len + kNoiseBlk - 1) / kNoiseBlk;
  const int lane = threadIdx.x & 31;
  for (int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < nblk; j += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    double e = 0.0;
    for (int64_t q = j * kNoiseBlk + lane; q < (j + 1) * kNoiseBlk && q < len; q += 32) { const double v = (double)data[q]; e += v * v; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if (lane == 0) sums[j] = e;
  }
}

// In-place explicit-index SpecAugment: one CTA per (clip, row-chunk); only masked elements are
// written (mask_along_axis fill semantics, TA/functional/functional.py:864-870,939-953).
template <typename T>
__global__ void __launch_bounds__(256) spec_mask_kernel(T* spec, int B, int F, int Tn, int64_t clip_stride,
                                                        const int32_t* fs, const int32_t* fl, int nF,
                                                        const int32_t* ts, const int32_t* tl, int nT, float mask_value) {
  const int b = blockIdx.y;
  T* s = spec + (size_t)b * clip_stride;
  const T mv = to_out<T>(mask_value);
  for (int f = blockIdx.x; f < F; f += gridDim.x) {
    bool row = false;
    for (int i = 0; i < nF; ++i) {
      const int s0 = fs[(size_t)b * nF + i], l = fl[(size_t)b * nF + i];
      row |= (f >= s0) && (f < s0 + l);
    }
    for (int t = threadIdx.x; t < Tn; t += blockDim.x) {
      bool mk = row;
      for (int i = 0; i < nT; ++i) {
        const int s0 = ts[(size_t)b * nT + i], l = tl[(size_t)b * nT + i];
        mk |= (t >= s0) && (t < s0 + l);
      }
      if (mk) s[(size_t)f * Tn + t] = mv;
    }
  }
}

}  // namespace wwf
