// wwfeat.cu - host side of libwwfeat.so: plan construction (device constants), bank
// registration, launch logic and the extern "C" ABI declared in include/wwfeat.h.
// There is deliberately no CPU implementation in this file: every entry point that
// computes needs a CUDA device and fails with WWF_ERR_CUDA otherwise.
#include "../../include/wwfeat.h"

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <functional>
#include <mutex>
#include <vector>

#include "wwf_aux.cuh"
#include "wwf_conv.cuh"
#include "wwf_feat.cuh"
#include "wwf_loader.cuh"
#include "wwf_pv.cuh"
#include "wwf_tables.h"

using namespace wwf;

// feat_kernel is instantiated in wwf_feat_inst.cu, one translation unit per n_fft (parallel build)
namespace wwf {
#define WWF_EXT(N, H)                                                                  \
  extern template __global__ void feat_kernel<N, H, float>(const FeatParams);          \
  extern template __global__ void feat_kernel<N, H, __half>(const FeatParams);         \
  extern template __global__ void feat_frames_kernel<N, H>(const FeatParams);
WWF_EXT(256, 0) WWF_EXT(256, 4) WWF_EXT(256, 5)
WWF_EXT(400, 0) WWF_EXT(400, 4) WWF_EXT(400, 5) WWF_EXT(400, 8)
WWF_EXT(512, 0) WWF_EXT(512, 4) WWF_EXT(512, 5) WWF_EXT(512, 8)
WWF_EXT(1024, 0) WWF_EXT(1024, 4) WWF_EXT(1024, 5) WWF_EXT(1024, 8) WWF_EXT(1024, 16)
WWF_EXT(2048, 0) WWF_EXT(2048, 4) WWF_EXT(2048, 5) WWF_EXT(2048, 8) WWF_EXT(2048, 16)
#undef WWF_EXT
}  // namespace wwf

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define WWF_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return fail(WWF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------
typedef void (*FeatKernel)(const FeatParams);
constexpr int kEpStaticSmem = 6144;   // upper bound of the epilogue kernels' static shared memory (row tables, mask bits)
struct FeatLaunch;
static void free_launch(FeatLaunch* l);

struct wwf_plan {
  wwf_config cfg;
  int device = 0, sm_count = 0, max_smem = 0;
  int K = 0, n_feat = 0, G = 1, zlen = 0, tw_total = 0, n_melw = 0, mel_rounds = 0, max_warps = 16, max_warps_flat = 16;
  unsigned short mel_pairs[kMaxMelRounds] = {};
  FeatKernel kernel = nullptr;          // fused per-clip kernel
  FeatKernel frames = nullptr;          // flat path: flat frames kernel (same n_fft / hop variant)
  FeatKernel epilogue_block = nullptr;  // flat path, log-mel: feat_epilogue_block_kernel<float | __half>
  FeatKernel epilogue_mma = nullptr;    // flat path, MFCC: feat_epilogue_mma_kernel<float | __half>
  FeatKernel epilogue_warp = nullptr;   // flat path, MFCC, calls without SpecAugment flags, common shapes: feat_epilogue_mma_warp_kernel
  size_t epilogue_warp_smem = 0;
  // device constants
  float* d_window = nullptr;
  float2* d_tw = nullptr;
  int2* d_mel_tasks = nullptr;
  int* d_nonfinite = nullptr;
  float* d_mel_w = nullptr;
  float* d_dct = nullptr;
  uint4* d_dct_frag = nullptr;
  float* d_dct_colsum = nullptr;
  // options (environment at plan creation, wwf_plan_set_option afterwards) and the per-(B, N) launch cache
  int opt_path = WWF_PATH_AUTO, opt_pdl = 1, opt_ep_warp = 1, opt_conv_order = 1;
  std::mutex cache_mu;
  std::vector<struct FeatLaunch*> launches;
  // noise bank (borrowed data, owned offsets)
  const float* noise_data = nullptr;
  int64_t* d_noise_offsets = nullptr;
  double* d_noise_prefix = nullptr;
  int64_t* d_noise_prefix_offsets = nullptr;
  int n_noise = 0;
  NoiseBankDev noise_dev() const { return NoiseBankDev{noise_data, d_noise_offsets, d_noise_prefix, d_noise_prefix_offsets, n_noise}; }
  // RIR bank (owned spectra) + conv constants
  float4* d_spec = nullptr;
  float2* d_conv_tw = nullptr;
  uint16_t* d_fused_l = nullptr;
  float2* d_fused_tw = nullptr;
  int n_rir = 0, rir_max_len = 0;
  int feat_warps_override = 0;
  bool generic_load = false;
  // time-stretch / pitch-shift constants and the resampler coefficient tables (built on first use)
  std::mutex lazy_mu;
  float* d_pv_window = nullptr;
  float2* d_pv_tw = nullptr;
  struct Resampler { int orig, nw, width, ntaps; float* coef; };
  std::vector<Resampler> resamplers;
  // optional per-kernel timing (wwf_profile_enable)
  bool prof = false;
  // per profiled call: one event before the first kernel and one after every kernel, plus which kernel each interval
  // timed (slot 0 conv_kernel | 1 feat_prep_kernel | 2 feat_frames_kernel / feat_kernel | 3 epilogue kernel)
  struct ProfCall { std::vector<cudaEvent_t> ev; std::vector<int> slot; bool split; };
  std::vector<ProfCall> prof_calls;
};

template <typename T>
static int upload(T** dst, const std::vector<T>& src) {
  *dst = nullptr;
  if (src.empty()) return WWF_OK;
  WWF_CUDA(cudaMalloc((void**)dst, src.size() * sizeof(T)));
  // pageable host memory: cudaMemcpy may return once the data is staged, before the DMA has landed, and the streams
  // PyTorch hands in are non-blocking (they do not synchronise with the legacy default stream the copy runs on) - so
  // wait for the copy itself; kernels launched afterwards on any stream then see the table
  WWF_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  WWF_CUDA(cudaStreamSynchronize(0));
  return WWF_OK;
}

// Kernel variant: frames loaded through the register-staged path when hop is 128, 160 (the
// reference's hop for every preset), 256 or 512 and the sample span fits in registers; generic
// per-element loads otherwise.
template <int NFFT>
static void select_kernel(wwf_plan* p, std::vector<float2>& tw) {
  using Plan = StftPlan<NFFT>;
  build_stft_twiddles<typename Plan::Rad>(tw);
  p->G = Plan::G;
  p->zlen = stft_zlen<NFFT>();
  p->max_warps = Plan::kThreads / 32;
  p->max_warps_flat = Plan::kFlatThreads / 32;
  p->tw_total = Plan::Rad::tw_total;
  const bool f16 = p->cfg.out_dtype == WWF_OUT_F16;
  p->kernel = f16 ? (FeatKernel)feat_kernel<NFFT, 0, __half> : (FeatKernel)feat_kernel<NFFT, 0, float>;
  p->frames = (FeatKernel)feat_frames_kernel<NFFT, 0>;
  p->epilogue_block = f16 ? (FeatKernel)feat_epilogue_block_kernel<__half> : (FeatKernel)feat_epilogue_block_kernel<float>;
  p->epilogue_mma = f16 ? (FeatKernel)feat_epilogue_mma_kernel<__half> : (FeatKernel)feat_epilogue_mma_kernel<float>;
  {
    // shapes with compile-time k-step / tile counts: 40 mels x 40 coefficients (BASELINE configs[1]) and the
    // reference's DataConfig defaults, 128 mels x 40 coefficients
    const int ks = (p->cfg.n_mels + 7) / 8, ntc = (p->cfg.n_mfcc + 7) / 8;
    if (ks == 5 && ntc == 5) {
      p->epilogue_mma = f16 ? (FeatKernel)feat_epilogue_mma_kernel<__half, 5, 5> : (FeatKernel)feat_epilogue_mma_kernel<float, 5, 5>;
      if (p->cfg.n_mels == 40 && p->cfg.n_mfcc == 40) {        // no padded column or coefficient: the warp-autonomous form
        p->epilogue_warp = f16 ? (FeatKernel)feat_epilogue_mma_warp_kernel<__half, 5, 5> : (FeatKernel)feat_epilogue_mma_warp_kernel<float, 5, 5>;
        p->epilogue_warp_smem = EmWarp<5>::smem_bytes(5);
      }
    }
    else if (ks == 16 && ntc == 5)
      p->epilogue_mma = f16 ? (FeatKernel)feat_epilogue_mma_kernel<__half, 16, 5> : (FeatKernel)feat_epilogue_mma_kernel<float, 16, 5>;
  }
  {
    // register-staged frame loads for hops that are a multiple of 32 and instantiated: 128, 160, 256
    if (!p->generic_load) {
      switch (p->cfg.hop_length) {
        case 128: p->kernel = f16 ? (FeatKernel)feat_kernel<NFFT, 4, __half> : (FeatKernel)feat_kernel<NFFT, 4, float>;
                  p->frames = (FeatKernel)feat_frames_kernel<NFFT, 4>; break;
        case 160: p->kernel = f16 ? (FeatKernel)feat_kernel<NFFT, 5, __half> : (FeatKernel)feat_kernel<NFFT, 5, float>;
                  p->frames = (FeatKernel)feat_frames_kernel<NFFT, 5>; break;
        case 256: if constexpr (NFFT > 256) { p->kernel = f16 ? (FeatKernel)feat_kernel<NFFT, 8, __half> : (FeatKernel)feat_kernel<NFFT, 8, float>;
                                              p->frames = (FeatKernel)feat_frames_kernel<NFFT, 8>; } break;
        case 512: if constexpr (NFFT >= 1024) { p->kernel = f16 ? (FeatKernel)feat_kernel<NFFT, 16, __half> : (FeatKernel)feat_kernel<NFFT, 16, float>;
                                                p->frames = (FeatKernel)feat_frames_kernel<NFFT, 16>; } break;
        default: break;
      }
    }
  }
}

extern "C" int wwf_version(void) { return WWF_VERSION; }
extern "C" const char* wwf_last_error(void) { return g_err; }
extern "C" int64_t wwf_launch_count(void) { return g_launches.load(); }

namespace {
__global__ void poison_smem_kernel(uint32_t word, int n_words) {
  extern __shared__ uint32_t poison_buf[];
  for (int i = threadIdx.x; i < n_words; i += blockDim.x) poison_buf[i] = word;
  __syncthreads();
  if (poison_buf[(threadIdx.x * 97) % n_words] != word) __trap();   // (keeps the stores alive)
}
}  // namespace

extern "C" int wwf_debug_poison_smem(int device, uint32_t word) {
  DeviceGuard guard(device);
  cudaDeviceProp prop;
  WWF_CUDA(cudaGetDeviceProperties(&prop, device));
  const int bytes = (int)prop.sharedMemPerBlockOptin - 1024;
  WWF_CUDA(cudaFuncSetAttribute((const void*)poison_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  // one CTA per SM at a time (it takes nearly all of the SM's shared memory); a few rounds so that every SM gets one
  poison_smem_kernel<<<8 * prop.multiProcessorCount, 1024, bytes>>>(word, bytes / 4);
  WWF_CUDA(cudaGetLastError());
  WWF_CUDA(cudaDeviceSynchronize());
  return WWF_OK;
}

extern "C" void wwf_plan_destroy(wwf_plan* p) {
  if (!p) return;
  DeviceGuard g(p->device);
  cudaFree(p->d_window); cudaFree(p->d_tw); cudaFree(p->d_mel_tasks);
  for (FeatLaunch* l : p->launches) free_launch(l);
  cudaFree(p->d_mel_w); cudaFree(p->d_dct); cudaFree(p->d_dct_frag); cudaFree(p->d_dct_colsum); cudaFree(p->d_nonfinite); cudaFree(p->d_noise_offsets); cudaFree(p->d_spec);
  cudaFree(p->d_noise_prefix); cudaFree(p->d_noise_prefix_offsets);
  for (auto& c : p->prof_calls) for (cudaEvent_t e : c.ev) cudaEventDestroy(e);
  cudaFree(p->d_conv_tw); cudaFree(p->d_fused_l); cudaFree(p->d_fused_tw);
  cudaFree(p->d_pv_window); cudaFree(p->d_pv_tw);
  for (auto& r : p->resamplers) cudaFree(r.coef);
  delete p;
}

extern "C" int wwf_plan_create(const wwf_config* cfg, int device, wwf_plan** out) {
  if (!cfg || !out) return fail(WWF_ERR_INVALID, "wwf_plan_create: null argument");
  *out = nullptr;
  const int n = cfg->n_fft;
  if (!(n == 256 || n == 400 || n == 512 || n == 1024 || n == 2048))
    return fail(WWF_ERR_UNSUPPORTED, "n_fft=%d not supported (256, 400, 512, 1024, 2048)", n);
  if (cfg->hop_length <= 0 || cfg->hop_length >= n) return fail(WWF_ERR_INVALID, "hop_length=%d must be in (0, n_fft)", cfg->hop_length);
  if (cfg->sample_rate <= 0) return fail(WWF_ERR_INVALID, "sample_rate=%d", cfg->sample_rate);
  const int K = n / 2 + 1;
  if (cfg->n_mels < 1 || cfg->n_mels > 128 || cfg->n_mels > K) return fail(WWF_ERR_INVALID, "n_mels=%d must be in [1, min(128, n_fft/2+1)]", cfg->n_mels);
  if (cfg->feature_type != WWF_FEAT_LOGMEL && cfg->feature_type != WWF_FEAT_MFCC) return fail(WWF_ERR_INVALID, "feature_type=%d", cfg->feature_type);
  if (cfg->feature_type == WWF_FEAT_MFCC && (cfg->n_mfcc < 1 || cfg->n_mfcc > cfg->n_mels))
    return fail(WWF_ERR_INVALID, "n_mfcc=%d must be in [1, n_mels=%d]", cfg->n_mfcc, cfg->n_mels);
  if (cfg->out_dtype != WWF_OUT_F32 && cfg->out_dtype != WWF_OUT_F16) return fail(WWF_ERR_INVALID, "out_dtype=%d", cfg->out_dtype);
  if (cfg->n_freq_masks < 0 || cfg->n_freq_masks > kMaxMasks || cfg->n_time_masks < 0 || cfg->n_time_masks > kMaxMasks)
    return fail(WWF_ERR_INVALID, "n_freq_masks/n_time_masks must be in [0, %d]", kMaxMasks);

  int ndev = 0;
  WWF_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(WWF_ERR_CUDA, "CUDA device %d not present (%d devices)", device, ndev);
  DeviceGuard guard(device);
  if (!guard.ok) return fail(WWF_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  cudaDeviceProp prop;
  WWF_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(WWF_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);

  wwf_plan* p = new wwf_plan();
  p->cfg = *cfg;
  p->cfg.window = nullptr; p->cfg.mel_fb = nullptr; p->cfg.dct = nullptr;   // host pointers are not retained
  if (p->cfg.f_max <= 0.f) p->cfg.f_max = (float)(cfg->sample_rate / 2);
  p->device = device;
  p->sm_count = prop.multiProcessorCount;
  p->max_smem = (int)prop.sharedMemPerBlockOptin;
  p->K = K;
  p->n_feat = cfg->feature_type == WWF_FEAT_MFCC ? cfg->n_mfcc : cfg->n_mels;
  // environment overrides are read ONCE, here (experiments and the A/B tools); wwf_plan_set_option changes them later
  if (const char* e = getenv("WWF_FEAT_WARPS")) p->feat_warps_override = atoi(e);
  if (const char* e = getenv("WWF_FEAT_PATH")) p->opt_path = !strcmp(e, "fused") ? WWF_PATH_FUSED : !strcmp(e, "split") ? WWF_PATH_FLAT : WWF_PATH_AUTO;
  p->opt_pdl = getenv("WWF_NO_PDL") ? 0 : 1;
  p->opt_conv_order = getenv("WWF_NO_CONV_ORDER") ? 0 : 1;    // A/B: reverb work items in batch order
  p->opt_ep_warp = getenv("WWF_NO_EP_WARP") ? 0 : 1;          // A/B: the block-wise tensor-core epilogue for every call
  p->generic_load = getenv("WWF_FEAT_GENERIC_LOAD") != nullptr;

  std::vector<float2> tw;
  switch (n) {
    case 256: select_kernel<256>(p, tw); break;
    case 400: select_kernel<400>(p, tw); break;
    case 512: select_kernel<512>(p, tw); break;
    case 1024: select_kernel<1024>(p, tw); break;
    default: select_kernel<2048>(p, tw); break;
  }

  std::vector<float> window(n);
  if (cfg->window) memcpy(window.data(), cfg->window, n * sizeof(float));
  else for (int i = 0; i < n; ++i) window[i] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * i / n));

  const int M = cfg->n_mels;
  std::vector<float> fb;
  if (cfg->mel_fb) fb.assign(cfg->mel_fb, cfg->mel_fb + (size_t)K * M);
  else fb = mel_fbanks32(K, p->cfg.f_min, p->cfg.f_max, M, cfg->sample_rate);
  std::vector<int> lo(M), ofs(M + 1);
  std::vector<float> w;
  for (int m = 0; m < M; ++m) {
    int first = -1, last = -1;
    for (int k = 0; k < K; ++k)
      if (fb[(size_t)k * M + m] != 0.f) { if (first < 0) first = k; last = k; }
    ofs[m] = (int)w.size();
    lo[m] = first < 0 ? 0 : first;
    if (first >= 0) for (int k = first; k <= last; ++k) w.push_back(fb[(size_t)k * M + m]);
  }
  ofs[M] = (int)w.size();
  if (w.empty()) w.push_back(0.f);
  const MelSchedule sched = build_mel_schedule(lo, ofs, w, n, n == 400 ? n : n / 2 + 1);   // (400: the identity-mapped plan)
  if (sched.rounds > kMaxMelRounds) { delete p; return fail(WWF_ERR_UNSUPPORTED, "mel lane schedule needs %d rounds (max %d)", sched.rounds, kMaxMelRounds); }
  for (int r = 0; r < sched.rounds; ++r) p->mel_pairs[r] = (unsigned short)sched.pairs[r];
  p->n_melw = (int)sched.w.size();
  p->mel_rounds = sched.rounds;

  std::vector<float> dct;
  if (cfg->feature_type == WWF_FEAT_MFCC) {
    const int C = cfg->n_mfcc;
    dct.resize((size_t)M * C);
    if (cfg->dct) memcpy(dct.data(), cfg->dct, dct.size() * sizeof(float));
    else
      for (int m = 0; m < M; ++m)
        for (int c = 0; c < C; ++c) {
          double v = cos(M_PI / (double)M * ((double)m + 0.5) * (double)c) * sqrt(2.0 / (double)M);
          if (c == 0) v *= 1.0 / sqrt(2.0);
          dct[(size_t)m * C + c] = (float)v;
        }
  }

  int rc;
  if ((rc = upload(&p->d_window, window)) || (rc = upload(&p->d_tw, tw)) || (rc = upload(&p->d_mel_tasks, sched.tasks)) ||
      (rc = upload(&p->d_mel_w, sched.w)) || (rc = upload(&p->d_dct, dct)) ||
      (rc = upload(&p->d_dct_frag, dct.empty() ? std::vector<uint4>() : build_dct_fragments(dct, M, cfg->n_mfcc))) ||
      (rc = upload(&p->d_dct_colsum, dct.empty() ? std::vector<float>() : dct_column_sums(dct, M, cfg->n_mfcc)))) {
    wwf_plan_destroy(p);
    return rc;
  }
  if (cudaMalloc((void**)&p->d_nonfinite, sizeof(int)) != cudaSuccess || cudaMemset(p->d_nonfinite, 0, sizeof(int)) != cudaSuccess) {
    wwf_plan_destroy(p);
    return fail(WWF_ERR_NOMEM, "cudaMalloc(nonfinite flag) failed");
  }
  cudaError_t e = cudaFuncSetAttribute((const void*)p->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, p->max_smem - 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)p->frames, cudaFuncAttributeMaxDynamicSharedMemorySize, p->max_smem - 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)p->epilogue_block, cudaFuncAttributeMaxDynamicSharedMemorySize, p->max_smem - 2048);
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)p->epilogue_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, p->max_smem - kEpStaticSmem);
  if (e == cudaSuccess && p->epilogue_warp)
    e = cudaFuncSetAttribute((const void*)p->epilogue_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->epilogue_warp_smem);
  if (e != cudaSuccess) {
    wwf_plan_destroy(p);
    return fail(WWF_ERR_CUDA, "cudaFuncSetAttribute(feat_kernel): %s (is libwwfeat.so built for this GPU?)", cudaGetErrorString(e));
  }
  *out = p;
  return WWF_OK;
}

extern "C" int wwf_check_finite(wwf_plan* p, void* stream, int* nonfinite) {
  if (!p || !nonfinite) return fail(WWF_ERR_INVALID, "wwf_check_finite: null argument");
  DeviceGuard guard(p->device);
  cudaStream_t st = (cudaStream_t)stream;
  int v = 0;
  WWF_CUDA(cudaMemcpyAsync(&v, p->d_nonfinite, sizeof(int), cudaMemcpyDeviceToHost, st));
  WWF_CUDA(cudaMemsetAsync(p->d_nonfinite, 0, sizeof(int), st));
  WWF_CUDA(cudaStreamSynchronize(st));
  *nonfinite = v;
  return WWF_OK;
}

extern "C" int wwf_profile_enable(wwf_plan* p, int enable) {
  if (!p) return fail(WWF_ERR_INVALID, "wwf_profile_enable: null plan");
  p->prof = enable != 0;
  return WWF_OK;
}

static int profile_collect(wwf_plan* p, double ms[4], int* n_calls, int* n_split) {
  DeviceGuard guard(p->device);
  const int n = (int)p->prof_calls.size();
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  cudaError_t err = cudaSuccess;
  int ns = 0;
  for (auto& c : p->prof_calls) {
    if (err == cudaSuccess && !c.ev.empty()) err = cudaEventSynchronize(c.ev.back());
    for (size_t k = 0; k + 1 < c.ev.size() && err == cudaSuccess; ++k) {
      float v = 0.f;
      err = cudaEventElapsedTime(&v, c.ev[k], c.ev[k + 1]);
      acc[c.slot[k]] += v;
    }
    ns += c.split ? 1 : 0;
    for (cudaEvent_t e : c.ev) cudaEventDestroy(e);
  }
  p->prof_calls.clear();
  if (err != cudaSuccess) return fail(WWF_ERR_CUDA, "wwf_profile_read: %s", cudaGetErrorString(err));
  for (int k = 0; k < 4; ++k) ms[k] = n ? acc[k] / n : 0.0;
  *n_calls = n;
  if (n_split) *n_split = ns;
  return WWF_OK;
}

extern "C" int wwf_profile_read(wwf_plan* p, double* conv_ms, double* feat_ms, int* n_calls) {
  if (!p || !conv_ms || !feat_ms || !n_calls) return fail(WWF_ERR_INVALID, "wwf_profile_read: null argument");
  double ms[4];
  int rc = profile_collect(p, ms, n_calls, nullptr);
  if (rc) return rc;
  *conv_ms = ms[0]; *feat_ms = ms[1] + ms[2] + ms[3];
  return WWF_OK;
}

extern "C" int wwf_profile_read_kernels(wwf_plan* p, double* kernel_ms, int* n_calls, int* n_split) {
  if (!p || !kernel_ms || !n_calls) return fail(WWF_ERR_INVALID, "wwf_profile_read_kernels: null argument");
  return profile_collect(p, kernel_ms, n_calls, n_split);
}

extern "C" int wwf_plan_info(const wwf_plan* p, wwf_info* out) {
  if (!p || !out) return fail(WWF_ERR_INVALID, "wwf_plan_info: null argument");
  out->n_freq = p->K; out->n_feat = p->n_feat; out->device = p->device; out->sm_count = p->sm_count;
  out->rir_fft_size = p->n_rir > 0 ? kConvP : 0; out->rir_max_len = p->rir_max_len;
  out->n_rir = p->n_rir; out->n_noise = p->n_noise;
  return WWF_OK;
}

extern "C" int wwf_num_frames(const wwf_plan* p, int n_samples) {
  if (!p || n_samples < 0) return fail(WWF_ERR_INVALID, "wwf_num_frames: bad argument");
  return n_samples / p->cfg.hop_length + 1;
}

// ------------------------------------------------------------------------------------------
// banks
// ------------------------------------------------------------------------------------------
static int ensure_conv_constants(wwf_plan* p) {
  if (p->d_conv_tw) return WWF_OK;
  std::vector<float2> tw, ftw;
  std::vector<uint16_t> fl;
  build_conv_tables(tw, fl, ftw);
  int rc;
  if ((rc = upload(&p->d_conv_tw, tw)) || (rc = upload(&p->d_fused_l, fl)) || (rc = upload(&p->d_fused_tw, ftw))) return rc;
  WWF_CUDA(cudaFuncSetAttribute((const void*)conv_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvSmemBytes));
  WWF_CUDA(cudaFuncSetAttribute((const void*)conv_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvSmemBytes));
  WWF_CUDA(cudaFuncSetAttribute((const void*)conv_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv_smem_bytes(kConvMaxOrder, true)));
  WWF_CUDA(cudaFuncSetAttribute((const void*)conv_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv_smem_bytes(kConvMaxOrder, true)));
  WWF_CUDA(cudaFuncSetAttribute((const void*)rir_spectrum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvSmemBytes));
  return WWF_OK;
}

extern "C" int wwf_bank_register(wwf_plan* p, int kind, const float* data, const int64_t* offsets, int count, void* stream) {
  if (!p || !offsets || count < 0 || (count > 0 && !data)) return fail(WWF_ERR_INVALID, "wwf_bank_register: bad argument");
  DeviceGuard guard(p->device);
  if (!guard.ok) return fail(WWF_ERR_CUDA, "cudaSetDevice(%d) failed", p->device);
  cudaStream_t st = (cudaStream_t)stream;
  for (int i = 0; i < count; ++i)
    if (offsets[i + 1] <= offsets[i]) return fail(WWF_ERR_INVALID, "bank clip %d is empty or offsets not increasing", i);
  std::vector<int64_t> ofs(offsets, offsets + count + 1);
  if (kind == WWF_BANK_NOISE) {
    for (int i = 0; i < count; ++i)
      if (offsets[i + 1] - offsets[i] > 0x7fffffffLL) return fail(WWF_ERR_UNSUPPORTED, "noise clip %d longer than 2^31 samples", i);
    cudaFree(p->d_noise_offsets); cudaFree(p->d_noise_prefix); cudaFree(p->d_noise_prefix_offsets);
    p->d_noise_offsets = nullptr; p->d_noise_prefix = nullptr; p->d_noise_prefix_offsets = nullptr;
    p->noise_data = nullptr;
    p->n_noise = 0;
    if (count == 0) return WWF_OK;
    // squared-sample prefix sums at every kNoiseBlk boundary (double): block sums on the GPU,
    // running sum on the host (one-time cost; lets the kernels get any segment's energy in O(1))
    std::vector<int64_t> pofs(count + 1, 0);
    for (int i = 0; i < count; ++i) pofs[i + 1] = pofs[i] + (offsets[i + 1] - offsets[i] + kNoiseBlk - 1) / kNoiseBlk + 1;
    double* d_sums = nullptr;
    WWF_CUDA(cudaMalloc((void**)&d_sums, (size_t)pofs[count] * sizeof(double)));
    for (int i = 0; i < count; ++i) {
      const int64_t len = offsets[i + 1] - offsets[i];
      const int64_t nblk = (len + kNoiseBlk - 1) / kNoiseBlk;
      int grid = (int)((nblk + 7) / 8);
      if (grid > 4096) grid = 4096;
      noise_block_sums_kernel<<<grid, 256, 0, st>>>(data + offsets[i], len, d_sums + pofs[i] + 1);
      g_launches++;
    }
    std::vector<double> pre((size_t)pofs[count]);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(pre.data(), d_sums, pre.size() * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_sums);
    if (e != cudaSuccess) return fail(WWF_ERR_CUDA, "noise_block_sums_kernel: %s", cudaGetErrorString(e));
    for (int i = 0; i < count; ++i) {
      pre[pofs[i]] = 0.0;
      for (int64_t j = pofs[i] + 1; j < pofs[i + 1]; ++j) pre[j] += pre[j - 1];
    }
    pofs.pop_back();
    int rc = upload(&p->d_noise_offsets, ofs);
    if (!rc) rc = upload(&p->d_noise_prefix, pre);
    if (!rc) rc = upload(&p->d_noise_prefix_offsets, pofs);
    if (rc) return rc;
    p->noise_data = data;
    p->n_noise = count;
    return WWF_OK;
  }
  if (kind != WWF_BANK_RIR) return fail(WWF_ERR_INVALID, "unknown bank kind %d", kind);
  int lmax = 0;
  for (int i = 0; i < count; ++i) {
    const int64_t l = offsets[i + 1] - offsets[i];
    if (l > kConvP / 2) return fail(WWF_ERR_UNSUPPORTED, "RIR %d has %lld taps; at most %d supported", i, (long long)l, kConvP / 2);
    if (l > lmax) lmax = (int)l;
  }
  cudaFree(p->d_spec);
  p->d_spec = nullptr;
  p->n_rir = 0;
  p->rir_max_len = 0;
  if (count == 0) return WWF_OK;
  int rc = ensure_conv_constants(p);
  if (rc) return rc;
  int64_t* d_ofs = nullptr;
  if ((rc = upload(&d_ofs, ofs))) return rc;
  WWF_CUDA(cudaMalloc((void**)&p->d_spec, (size_t)count * kSpecPerRir * sizeof(float4)));
  SpecParams sp{data, d_ofs, count, p->d_spec, p->d_conv_tw, p->d_fused_l};
  rir_spectrum_kernel<<<count, kConvThreads, kConvSmemBytes, st>>>(sp);
  g_launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_ofs);
  if (e != cudaSuccess) {
    cudaFree(p->d_spec);
    p->d_spec = nullptr;
    return fail(WWF_ERR_CUDA, "rir_spectrum_kernel: %s", cudaGetErrorString(e));
  }
  p->n_rir = count;
  p->rir_max_len = lmax;
  return WWF_OK;
}

// ------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------
static inline int64_t round_up4(int64_t v) { return (v + 3) & ~(int64_t)3; }

// overlap-save geometry for N-sample clips: history, valid outputs per block, block count
static void conv_geometry(const wwf_plan* p, int N, int* hist, int* valid, int* nb) {
  if ((int64_t)N + p->rir_max_len - 1 <= kConvP) { *hist = 0; *valid = kConvP; *nb = 1; }
  else {
    *hist = (int)round_up4(p->rir_max_len - 1);
    *valid = kConvP - *hist;
    *nb = (N + *valid - 1) / *valid;
  }
}

// workspace = [reverb part: reverberated clips [B][roundup4(N)] + their per-block energies [B][nb]] (RIR bank only)
//           + [flat-path part: dB tiles [B][T][roundup4(n_mels)] | clip maxima int[B4] | ClipMix[B]]
static size_t conv_ws_bytes(const wwf_plan* p, int B, int N) {
  if (p->n_rir == 0) return 0;
  int hist, valid, nb;
  conv_geometry(p, N, &hist, &valid, &nb);
  return ((size_t)B * (size_t)round_up4(N) + (size_t)round_up4((int64_t)B * nb)) * sizeof(float);
}

// Everything about a wwf_featurize call that depends only on (plan, options, B, N): shared-memory layouts, CTA shapes
// (the occupancy queries), grids and workspace sizes.  Built once per shape and cached in the plan, so a call costs
// a mutex, a struct copy and the launches (the 1-clip inference path is host-bound).
struct FeatLaunch {
  int B = 0, N = 0, T = 0;
  FeatParams fp{};               // configuration, plan constants and layout offsets; per-call pointers are filled per call
  // single-kernel path
  bool fused_ok = false;
  int fused_warps = 0, fused_grid = 0;
  size_t fused_smem = 0, fused_fixed = 0;
  // flat path
  bool flat_wanted = false, flat_ok = false;
  size_t flat_bytes = 0, off_clipmax = 0, off_mix = 0;   // (the reverb part of the workspace depends on the RIR bank: per call)
  int frames_warps = 0;
  unsigned frames_grid = 0, ep_grid = 0;
  size_t frames_smem = 0, ep_smem = 0;
  int ep_threads = 0;
  FeatKernel ep_kernel = nullptr;   // feat_epilogue_mma_kernel (MFCC) or feat_epilogue_block_kernel (log-mel)
  unsigned epw_grid = 0;            // feat_epilogue_mma_warp_kernel (0: not available for this plan / shape)
};
static void free_launch(FeatLaunch* l) { delete l; }

// Which launch shape?  Measured with tools/bench_paths.py and tools/bench_small.py (profiles/README.md):
//  * few clips (at most one per two SMs): the fused kernel would put each clip on ONE CTA and leave the rest of the GPU
//    idle - the flat queue spreads a clip's frame groups over every SM (1 clip, n_fft 1024, 2.5 s: 142 -> 37 us);
//  * many clips: the flat queue wins once there are several frame groups per resident warp (no per-clip barriers, no
//    clip-granular grid quantisation): from ~6 for n_fft 256 / 400, whose fused kernel keeps 3 CTAs per SM, from ~2
//    for the larger FFTs, whose fused kernel fits one CTA per SM;
//  * in between the fused kernel (one launch, no intermediate) is faster.  Log-mel has no DCT to amortise the tile's
//    round trip: below n_fft 1024 it only takes the flat queue in the few-clips case.
// WWF_FEAT_PATH=fused|split at plan creation or wwf_plan_set_option(WWF_OPT_FEAT_PATH) forces one (the GPU tests run both).
static bool want_flat(const wwf_plan* p, int B, int N, bool fused_ok) {
  if (p->cfg.cmvn) return false;                               // CMVN needs whole rows of a clip in one CTA
  if (p->opt_path == WWF_PATH_FUSED) return false;
  if (p->opt_path == WWF_PATH_FLAT) return true;
  if (!fused_ok) return true;                                  // long clips: only the flat path has no per-clip tile
  if (2 * B <= p->sm_count) return true;
  const bool mfcc = p->cfg.feature_type == WWF_FEAT_MFCC;
  if (!mfcc && p->cfg.n_fft < 1024) return false;
  const int T = N / p->cfg.hop_length + 1;
  const long long groups = (long long)B * ((T + 2 * p->G - 1) / (2 * p->G));
  const long long per_warp = !mfcc ? 4 : p->cfg.n_fft <= 400 ? 6 : 2;
  return groups >= per_warp * 20 * p->sm_count;
}

static FeatLaunch* build_launch(wwf_plan* p, int B, int N) {
  FeatLaunch* l = new FeatLaunch();
  l->B = B; l->N = N;
  const int T = N / p->cfg.hop_length + 1;
  l->T = T;
  const int F = p->n_feat, M = p->cfg.n_mels;
  const bool mfcc = p->cfg.feature_type == WWF_FEAT_MFCC;
  const int pitch = T | 1;
  const int nfft = p->cfg.n_fft;
  auto al4 = [](long long v) { return (v + 3) & ~3ll; };
  FeatParams& fp = l->fp;
  fp.c8 = mfcc ? ((F + 7) & ~7) : 0;
  fp.n_melw = p->n_melw;
  fp.mel_rounds = p->mel_rounds;
  memcpy(fp.mel_pairs, p->mel_pairs, sizeof(fp.mel_pairs));
  // ---- single-kernel path: shared-memory layout (64-bit: very long clips must not wrap) ----
  long long o = al4((long long)M * pitch);
  const long long off_res = o;      o += (mfcc && p->cfg.cmvn) ? al4((long long)F * pitch) : 0;
  const long long off_window = o;   o += al4(nfft);
  const long long off_tw = o;       o += al4(2 * p->tw_total);
  const long long off_melw = o;     o += al4(p->n_melw);
  const long long off_dct = o;      o += mfcc ? (long long)M * fp.c8 : 0;
  const long long off_meltasks = o; o += 2ll * 32 * p->mel_rounds;
  const long long off_rowmask = o;  o += al4((F + 3) / 4);
  const long long off_colmask = o;  o += al4((T + 3) / 4);
  const long long off_z = o;
  const size_t per_warp = (size_t)p->G * p->zlen * sizeof(float2);
  const size_t budget = (size_t)p->max_smem - 1024;   // static smem of the kernel is 256 B
  l->fused_fixed = (size_t)o * sizeof(float);
  l->fused_ok = l->fused_fixed + per_warp <= budget;           // else: only the flat path (no per-clip tile) can run
  const int cands[] = {16, 14, 13, 12, 11, 10, 8, 6, 4, 2, 1};
  if (l->fused_ok) {
    fp.off_res = (int)off_res; fp.off_window = (int)off_window; fp.off_tw = (int)off_tw; fp.off_melw = (int)off_melw;
    fp.off_dct = (int)off_dct; fp.off_meltasks = (int)off_meltasks; fp.off_rowmask = (int)off_rowmask;
    fp.off_colmask = (int)off_colmask; fp.off_z = (int)off_z;
    // CTA shape: the candidate (warps per CTA) that keeps the most warps resident per SM
    int ctas_per_sm = 1, best = -1;
    for (int c : cands) {
      if (c > p->max_warps || (p->feat_warps_override > 0 && c != p->feat_warps_override)) continue;
      const size_t sm = l->fused_fixed + (size_t)c * per_warp;
      if (sm > budget) continue;
      int nb = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)p->kernel, c * 32, sm) != cudaSuccess || nb < 1) continue;
      // resident warps, plus a bonus for a third CTA: the per-clip CTA-wide phases (max, DCT, store)
      // of one CTA then overlap with the frame phases of two others (measured: 3x6 warps beats 2x10).
      // A small batch cannot fill several CTAs per SM: count only the CTAs it will actually place.
      const int nb_eff = std::min(nb, (B + p->sm_count - 1) / p->sm_count);
      const int score = nb_eff * c + (nb_eff >= 3 ? 3 : 0);
      if (score > best) { best = score; l->fused_warps = c; ctas_per_sm = nb; }
    }
    if (l->fused_warps == 0) l->fused_ok = false;
    l->fused_smem = l->fused_fixed + (size_t)l->fused_warps * per_warp;
    l->fused_grid = std::min(B, p->sm_count * ctas_per_sm);
  }
  fp.B = B; fp.N = N; fp.T = T; fp.hop = p->cfg.hop_length;
  fp.n_mels = M; fp.n_mfcc = p->cfg.n_mfcc; fp.n_feat = F; fp.is_mfcc = mfcc;
  fp.cmvn = p->cfg.cmvn != 0; fp.top_db = p->cfg.top_db; fp.cmvn_eps = p->cfg.cmvn_eps; fp.mask_value = p->cfg.mask_value;
  fp.tile_pitch = pitch;
  fp.window = p->d_window; fp.tw = p->d_tw; fp.mel_tasks = p->d_mel_tasks; fp.mel_w = p->d_mel_w; fp.dct = p->d_dct; fp.dct_frag = p->d_dct_frag; fp.dct_colsum = p->d_dct_colsum;
  fp.nonfinite_flag = p->d_nonfinite;
  // ---- flat path: [prep] -> flat frames -> epilogue ----
  l->flat_wanted = want_flat(p, B, N, l->fused_ok);
  if (l->flat_wanted) {
    fp.mp = (int)round_up4(M);
    fp.ngroups = (T + 2 * p->G - 1) / (2 * p->G);
    const size_t tile_bytes = (size_t)B * T * fp.mp * sizeof(float);
    l->off_clipmax = tile_bytes;
    l->off_mix = l->off_clipmax + (size_t)round_up4(B) * sizeof(int);          // 16-byte aligned
    l->flat_bytes = l->off_mix + (size_t)B * sizeof(ClipMix);
    long long fo = al4(nfft);
    fp.f_off_tw = (int)fo;       fo += al4(2 * p->tw_total);
    fp.f_off_melw = (int)fo;     fo += al4(p->n_melw);
    fp.f_off_meltasks = (int)fo; fo += 2ll * 32 * p->mel_rounds;
    fp.f_off_z = (int)fo;
    const size_t f_fixed = (size_t)fo * sizeof(float);
    int fc = 1, fbest = -1;
    for (int c : cands) {
      if (c > p->max_warps_flat) continue;
      const size_t sm = f_fixed + (size_t)c * per_warp;
      if (sm > budget) continue;
      int nb = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)p->frames, c * 32, sm) != cudaSuccess || nb < 1) continue;
      if (nb * c > fbest) { fbest = nb * c; l->frames_warps = c; fc = nb; }
    }
    l->frames_smem = f_fixed + (size_t)l->frames_warps * per_warp;
    const long long items = (long long)B * fp.ngroups;
    if (l->frames_warps > 0) l->frames_grid = (unsigned)std::min<long long>((long long)p->sm_count * fc, (items + l->frames_warps - 1) / l->frames_warps);
    long long eitems = 0;
    if (mfcc) {
      // tensor-core DCT: 128 rows of the flat [B*T][n_mels] tile matrix per CTA
      fp.em_k8 = (M + 7) & ~7;
      fp.em_ap = fp.em_k8 + 4;
      l->ep_kernel = p->epilogue_mma;
      l->ep_threads = kEmThreads;
      l->ep_smem = (size_t)kEmRows * fp.em_ap * sizeof(float) + (size_t)(fp.em_k8 / 8) * (fp.c8 / 8) * 32 * sizeof(uint4);
      eitems = ((long long)B * T + kEmRows - 1) / kEmRows;
    } else {
      // log-mel: a plain clamp + mask + transpose, 64 frames per CTA
      fp.eb_frames = 64;
      fp.eb_pitch = fp.eb_frames | 1;
      l->ep_kernel = p->epilogue_block;
      l->ep_threads = 256;
      l->ep_smem = (size_t)((M * fp.eb_pitch + 3) & ~3) * sizeof(float);
      eitems = (long long)B * ((T + fp.eb_frames - 1) / fp.eb_frames);
    }
    int eocc = 0;
    if (l->ep_smem <= (size_t)p->max_smem - kEpStaticSmem &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&eocc, (const void*)l->ep_kernel, l->ep_threads, l->ep_smem) == cudaSuccess && eocc >= 1)
    {
      // persistent CTAs with equal shares: 1208 row blocks over 592 CTA slots would be two full rounds plus a third one
      // for 24 CTAs (3 item times for 2.04 items' worth of work); ceil(items / rounds) CTAs do exactly `rounds` each
      const long long slots = (long long)p->sm_count * eocc;
      const long long rounds = (eitems + slots - 1) / slots;
      l->ep_grid = (unsigned)((eitems + rounds - 1) / rounds);
    }
    if (mfcc && p->epilogue_warp && fp.mp == M && p->epilogue_warp_smem <= (size_t)p->max_smem - 1024) {
      int wocc = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wocc, (const void*)p->epilogue_warp, kEmThreads, p->epilogue_warp_smem) == cudaSuccess && wocc >= 1) {
        const long long slabs = ((long long)B * T + 15) / 16, per_cta = kEmThreads / 32;
        l->epw_grid = (unsigned)std::min<long long>((long long)p->sm_count * wocc, (slabs + per_cta - 1) / per_cta);
      }
    }
    l->flat_ok = l->frames_warps > 0 && l->frames_grid > 0 && l->ep_grid > 0 && (long long)B * fp.ngroups < (1ll << 31) - (1 << 20);
  }
  cudaGetLastError();   // a failed occupancy query must not poison the next launch check
  return l;
}

static FeatLaunch* get_launch(wwf_plan* p, int B, int N) {
  std::lock_guard<std::mutex> lk(p->cache_mu);
  for (FeatLaunch* l : p->launches)
    if (l->B == B && l->N == N) return l;
  if (p->launches.size() >= 64) {                              // bounded: drop the oldest shape
    free_launch(p->launches.front());
    p->launches.erase(p->launches.begin());
  }
  p->launches.push_back(build_launch(p, B, N));
  return p->launches.back();
}

extern "C" int wwf_plan_set_option(wwf_plan* p, int option, int value) {
  if (!p) return fail(WWF_ERR_INVALID, "wwf_plan_set_option: null plan");
  std::lock_guard<std::mutex> lk(p->cache_mu);
  if (option == WWF_OPT_CONV_ORDER) {                          // (not part of the cached launch shapes)
    p->opt_conv_order = value != 0;
    return WWF_OK;
  }
  if (option == WWF_OPT_FEAT_PATH) {
    if (value != WWF_PATH_AUTO && value != WWF_PATH_FUSED && value != WWF_PATH_FLAT) return fail(WWF_ERR_INVALID, "wwf_plan_set_option: path %d", value);
    p->opt_path = value;
  } else if (option == WWF_OPT_PDL) {
    p->opt_pdl = value != 0;
  } else if (option == WWF_OPT_EPILOGUE_WARP) {
    p->opt_ep_warp = value != 0;
  } else {
    return fail(WWF_ERR_INVALID, "wwf_plan_set_option: unknown option %d", option);
  }
  for (FeatLaunch* l : p->launches) free_launch(l);            // cached shapes were built for the old options
  p->launches.clear();
  return WWF_OK;
}

extern "C" size_t wwf_workspace_bytes(const wwf_plan* p, int B, int N) {
  if (!p || B <= 0 || N <= 0) return 0;
  DeviceGuard guard(p->device);
  const FeatLaunch* l = get_launch(const_cast<wwf_plan*>(p), B, N);
  return conv_ws_bytes(p, B, N) + (l->flat_ok ? l->flat_bytes : 0);
}

static bool wants_reverb(const wwf_plan* p, const wwf_aug* aug) { return aug && aug->rir_idx && p->n_rir > 0; }

// Overlap-save reverb of the clips whose rir_idx addresses the bank into the workspace: geometry and pointers ...
static int conv_setup(wwf_plan* p, const float* wav, int B, int N, int64_t wav_stride, const wwf_aug* aug,
                      void* workspace, size_t workspace_bytes, ConvParams* cp, bool* active) {
  *active = false;
  *cp = ConvParams{};
  if (!wants_reverb(p, aug)) return WWF_OK;
  const size_t need = conv_ws_bytes(p, B, N);
  if (!workspace || workspace_bytes < need) return fail(WWF_ERR_WORKSPACE, "workspace too small: %zu < %zu bytes", workspace_bytes, need);
  if (reinterpret_cast<uintptr_t>(workspace) & 15) return fail(WWF_ERR_WORKSPACE, "workspace must be 16-byte aligned");
  cp->wav = wav; cp->wav_stride = wav_stride;
  cp->rev = (float*)workspace; cp->rev_stride = round_up4(N);
  cp->rir_idx = aug->rir_idx; cp->B = B; cp->N = N; cp->n_rir = p->n_rir;
  int nb = 1;
  conv_geometry(p, N, &cp->hist, &cp->valid, &nb);
  cp->es_part = cp->rev + (size_t)B * cp->rev_stride;
  cp->es_nb = nb;
  cp->spec = p->d_spec; cp->tw = p->d_conv_tw; cp->fused_l = p->d_fused_l; cp->fused_tw = p->d_fused_tw;
  *active = true;
  return WWF_OK;
}

// ... and the launch (pdl: chained behind one of our own kernels with programmatic dependent launch)
static int conv_launch(wwf_plan* p, ConvParams& cp, cudaStream_t st, bool pdl) {
  const int items = cp.es_nb * cp.B;
  // more items than CTAs: deal the reverberated clips round-robin (the CTAs build the order themselves, wwf_conv.cuh)
  cp.ordered = items > p->sm_count && cp.B <= kConvMaxOrder && p->opt_conv_order ? 1 : 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(items < p->sm_count ? items : p->sm_count); cfg.blockDim = dim3(kConvThreads);
  cfg.dynamicSmemBytes = conv_smem_bytes(cp.B, cp.ordered != 0); cfg.stream = st;
  cudaLaunchAttribute at{};
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &at; cfg.numAttrs = pdl ? 1 : 0;
  if (cp.mix_g != nullptr) {
    if (cp.ordered) WWF_CUDA(cudaLaunchKernelEx(&cfg, conv_kernel<true, true>, cp));
    else WWF_CUDA(cudaLaunchKernelEx(&cfg, conv_kernel<true, false>, cp));
  } else {
    if (cp.ordered) WWF_CUDA(cudaLaunchKernelEx(&cfg, conv_kernel<false, true>, cp));
    else WWF_CUDA(cudaLaunchKernelEx(&cfg, conv_kernel<false, false>, cp));
  }
  g_launches++;
  return WWF_OK;
}

static int check_batch(const wwf_plan* p, const void* wav, int B, int N, int64_t wav_stride, const char* who) {
  if (!p || !wav) return fail(WWF_ERR_INVALID, "%s: null argument", who);
  if (B <= 0) return fail(WWF_ERR_INVALID, "%s: B=%d", who, B);
  if (N <= p->cfg.n_fft / 2) return fail(WWF_ERR_INVALID, "%s: N=%d must exceed n_fft/2=%d (reflect padding)", who, N, p->cfg.n_fft / 2);
  if (N > (1 << 24)) return fail(WWF_ERR_UNSUPPORTED, "%s: N=%d > 2^24 samples", who, N);
  if (wav_stride < N) return fail(WWF_ERR_INVALID, "%s: wav_stride=%lld < N=%d", who, (long long)wav_stride, N);
  return WWF_OK;
}

static cudaError_t launch_feat(FeatKernel k, const FeatParams& fp, unsigned grid, unsigned block, size_t smem, cudaStream_t st, bool pdl) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at{};
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, k, fp);
}

extern "C" int wwf_featurize(wwf_plan* p, const float* wav, int B, int N, int64_t wav_stride, const wwf_aug* aug,
                             void* out, int64_t out_stride, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_batch(p, wav, B, N, wav_stride, "wwf_featurize");
  if (rc) return rc;
  if (!out) return fail(WWF_ERR_INVALID, "wwf_featurize: out is null");
  DeviceGuard guard(p->device);
  if (!guard.ok) return fail(WWF_ERR_CUDA, "cudaSetDevice(%d) failed", p->device);
  const FeatLaunch* l = get_launch(p, B, N);
  const int T = l->T, F = p->n_feat, M = p->cfg.n_mels;
  if (out_stride < (int64_t)F * T) return fail(WWF_ERR_INVALID, "wwf_featurize: out_stride=%lld < n_feat*T=%d", (long long)out_stride, F * T);
  cudaStream_t st = (cudaStream_t)stream;

  FeatParams fp = l->fp;
  fp.wav = wav; fp.wav_stride = wav_stride;
  fp.out = out; fp.out_stride = out_stride;
  const bool reverb = wants_reverb(p, aug);
  if (aug) {
    if (reverb) { fp.rir_idx = aug->rir_idx; fp.n_rir = p->n_rir; }
    if (aug->noise_idx && p->n_noise > 0) {
      fp.noise_idx = aug->noise_idx; fp.noise_off = aug->noise_off; fp.snr_db = aug->snr_db;
      fp.noise = p->noise_dev();
    }
    if (aug->fmask_start && aug->fmask_len && p->cfg.n_freq_masks > 0) { fp.fs = aug->fmask_start; fp.fl = aug->fmask_len; fp.nF = p->cfg.n_freq_masks; }
    if (aug->tmask_start && aug->tmask_len && p->cfg.n_time_masks > 0) { fp.ts = aug->tmask_start; fp.tl = aug->tmask_len; fp.nT = p->cfg.n_time_masks; }
  }
  const size_t conv_bytes = conv_ws_bytes(p, B, N);
  const bool flat = l->flat_ok && workspace && workspace_bytes >= conv_bytes + l->flat_bytes && !(reinterpret_cast<uintptr_t>(workspace) & 15);
  if (flat) {
    char* fw = (char*)workspace + conv_bytes;
    fp.tile_g = (float*)fw;
    fp.clip_max = (int*)(fw + l->off_clipmax);
    fp.mix_g = fp.noise_idx ? (ClipMix*)(fw + l->off_mix) : nullptr;
  }

  // optional per-kernel timing: an event before the first kernel and after every kernel of the call
  wwf_plan::ProfCall pc;
  pc.split = flat;
  auto mark = [&](int slot) -> cudaError_t {                   // slot < 0: the opening event
    if (!p->prof) return cudaSuccess;
    cudaEvent_t e;
    cudaError_t err = cudaEventCreate(&e);
    if (err == cudaSuccess) err = cudaEventRecord(e, st);
    if (err == cudaSuccess) { pc.ev.push_back(e); if (slot >= 0) pc.slot.push_back(slot); }
    return err;
  };
  WWF_CUDA(mark(-1));
  ConvParams cp;
  bool conv_on = false;
  if ((rc = conv_setup(p, wav, B, N, wav_stride, aug, workspace, workspace_bytes, &cp, &conv_on))) return rc;
  fp.rev = conv_on ? cp.rev : nullptr; fp.rev_stride = conv_on ? cp.rev_stride : 0;
  if (!conv_on) { fp.rir_idx = nullptr; fp.n_rir = 0; }
  fp.es_part = conv_on ? cp.es_part : nullptr; fp.es_nb = conv_on ? cp.es_nb : 0;
  // Programmatic dependent launch: every kernel of the call stages its plan constants first and touches inputs /
  // predecessors' results only after cudaGridDependencySynchronize(), so it is safe behind anything on the stream.
  const bool pdl = p->opt_pdl && !p->prof;
  if (flat) {
    // identity of the clip maxima: conv_kernel resets them in its prologue; without reverb a byte-wise memset
    if (conv_on) { cp.clip_max = fp.clip_max; cp.n_clip_max = (int)round_up4(B); }
    else WWF_CUDA(cudaMemsetAsync(fp.clip_max, 0x80, (size_t)round_up4(B) * sizeof(int), st));
  }

  if (flat) {
    // Noise: one mix record per clip.  conv_kernel makes them itself (single-block clips, a bounded number of clips
    // per CTA); otherwise feat_prep_kernel does, after the reverb (one CTA per clip; its noise side runs before the
    // programmatic-launch wait).
    const bool need_mix = fp.mix_g != nullptr;
    const int conv_grid = conv_on ? std::min(cp.es_nb * B, p->sm_count) : 1;
    const bool mix_in_conv = need_mix && conv_on && cp.es_nb == 1 && (B + conv_grid - 1) / conv_grid <= kConvMaxOwn;
    if (conv_on) {
      if (mix_in_conv) {
        cp.mix_g = fp.mix_g; cp.noise = fp.noise; cp.noise_idx = fp.noise_idx; cp.noise_off = fp.noise_off; cp.snr_db = fp.snr_db;
      }
      if ((rc = conv_launch(p, cp, st, pdl))) return rc;
      WWF_CUDA(mark(0));
    }
    if (need_mix && !mix_in_conv) {
      WWF_CUDA(launch_feat((FeatKernel)feat_prep_kernel<0>, fp, (unsigned)B, 256, 0, st, pdl));
      g_launches++;
      WWF_CUDA(mark(1));
    }
    WWF_CUDA(launch_feat(p->frames, fp, l->frames_grid, (unsigned)(l->frames_warps * 32), l->frames_smem, st, pdl));
    WWF_CUDA(mark(2));
    const bool masked = (fp.fs != nullptr && fp.nF > 0) || (fp.ts != nullptr && fp.nT > 0);
    if (l->epw_grid > 0 && !masked && p->opt_ep_warp)
      WWF_CUDA(launch_feat(p->epilogue_warp, fp, l->epw_grid, (unsigned)kEmThreads, p->epilogue_warp_smem, st, pdl));
    else
      WWF_CUDA(launch_feat(l->ep_kernel, fp, l->ep_grid, (unsigned)l->ep_threads, l->ep_smem, st, pdl));
    g_launches += 2;
    WWF_CUDA(cudaGetLastError());
    WWF_CUDA(mark(3));
    if (p->prof) p->prof_calls.push_back(std::move(pc));
    return WWF_OK;
  }
  if (!l->fused_ok)
    return fail(WWF_ERR_UNSUPPORTED, "clip too long for the in-shared-memory tile of the single-kernel path: %zu bytes of tiles/tables "
                "(N=%d, T=%d, n_mels=%d)%s", l->fused_fixed, N, T, M,
                p->cfg.cmvn ? "; CMVN plans have no other path" : "; pass a workspace of wwf_workspace_bytes() to use the flat path");
  if (conv_on) {
    if ((rc = conv_launch(p, cp, st, false))) return rc;
    WWF_CUDA(mark(0));
  }
  WWF_CUDA(launch_feat(p->kernel, fp, (unsigned)l->fused_grid, (unsigned)(l->fused_warps * 32), l->fused_smem, st, pdl));
  g_launches++;
  WWF_CUDA(cudaGetLastError());
  WWF_CUDA(mark(2));
  if (p->prof) p->prof_calls.push_back(std::move(pc));
  return WWF_OK;
}

extern "C" int wwf_augment(wwf_plan* p, const float* wav, int B, int N, int64_t wav_stride, const wwf_aug* aug,
                           float* out_wav, int64_t out_stride, void* workspace, size_t workspace_bytes, void* stream) {
  if (!p || !wav || !out_wav) return fail(WWF_ERR_INVALID, "wwf_augment: null argument");
  if (B <= 0 || N <= 0 || wav_stride < N || out_stride < N) return fail(WWF_ERR_INVALID, "wwf_augment: bad shape B=%d N=%d", B, N);
  DeviceGuard guard(p->device);
  if (!guard.ok) return fail(WWF_ERR_CUDA, "cudaSetDevice(%d) failed", p->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (wants_reverb(p, aug) && out_wav == wav) return fail(WWF_ERR_INVALID, "wwf_augment: in-place is not allowed with reverb");
  ConvParams cp;
  bool conv_on = false;
  int rc = conv_setup(p, wav, B, N, wav_stride, aug, workspace, workspace_bytes, &cp, &conv_on);
  if (rc) return rc;
  if (conv_on && (rc = conv_launch(p, cp, st, false))) return rc;
  float* rev = conv_on ? cp.rev : nullptr;
  const int64_t rev_stride = conv_on ? cp.rev_stride : 0;
  const float* es_part = conv_on ? cp.es_part : nullptr;
  const int es_nb = conv_on ? cp.es_nb : 0;
  MixParams mp{};
  mp.wav = wav; mp.wav_stride = wav_stride; mp.rev = rev; mp.rev_stride = rev_stride;
  if (aug) {
    mp.rir_idx = rev ? aug->rir_idx : nullptr; mp.n_rir = p->n_rir;
    if (aug->noise_idx && p->n_noise > 0) {
      mp.noise_idx = aug->noise_idx; mp.noise_off = aug->noise_off; mp.snr_db = aug->snr_db;
      mp.noise = p->noise_dev();
    }
  }
  mp.es_part = es_part; mp.es_nb = es_nb;
  mp.out = out_wav; mp.out_stride = out_stride; mp.B = B; mp.N = N;
  mix_kernel<<<B, 512, 0, st>>>(mp);
  g_launches++;
  WWF_CUDA(cudaGetLastError());
  return WWF_OK;
}

extern "C" int wwf_gather_clips(const void* bank, int dtype, int64_t n_clips, int N, int64_t bank_stride, const int64_t* idx,
                                int B, float* out, int64_t out_stride, int device, void* stream) {
  if (!bank || !idx || !out || B <= 0 || N <= 0 || n_clips <= 0 || bank_stride < N || out_stride < N)
    return fail(WWF_ERR_INVALID, "wwf_gather_clips: bad argument");
  if (dtype != WWF_BANK_F32 && dtype != WWF_BANK_I16) return fail(WWF_ERR_INVALID, "wwf_gather_clips: dtype=%d", dtype);
  if (B > 65535) return fail(WWF_ERR_UNSUPPORTED, "wwf_gather_clips: B=%d > 65535 clips per call", B);
  DeviceGuard guard(device);
  if (!guard.ok) return fail(WWF_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  cudaStream_t st = (cudaStream_t)stream;
  const int vec = N % 8 == 0 && bank_stride % 8 == 0 && out_stride % 8 == 0 && !(reinterpret_cast<uintptr_t>(bank) & 15) &&
                  !(reinterpret_cast<uintptr_t>(out) & 15);
  // enough CTAs per clip to fill the GPU at small B, at most 8 elements per thread-iteration
  int chunks = (N / 8 + 255) / 256;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms <= 0) sms = 148;
  const int want = (4 * sms + B - 1) / B;
  if (chunks > want) chunks = want;
  if (chunks < 1) chunks = 1;
  const dim3 grid(chunks, B);
  if (dtype == WWF_BANK_F32) gather_clips_kernel<float><<<grid, 256, 0, st>>>((const float*)bank, n_clips, N, bank_stride, idx, out, out_stride, vec);
  else gather_clips_kernel<int16_t><<<grid, 256, 0, st>>>((const int16_t*)bank, n_clips, N, bank_stride, idx, out, out_stride, vec);
  g_launches++;
  WWF_CUDA(cudaGetLastError());
  return WWF_OK;
}

extern "C" int wwf_draw_aug(wwf_plan* p, const wwf_draw_config* dc, uint64_t first_index, int B, int T, const wwf_aug* out, void* stream) {
  if (!p || !dc || !out || B <= 0 || T <= 0) return fail(WWF_ERR_INVALID, "wwf_draw_aug: bad argument");
  auto thr = [](double prob) { return prob <= 0.0 ? (uint64_t)0 : prob >= 1.0 ? (uint64_t)1 << 32 : (uint64_t)floor(prob * 4294967296.0); };
  if (p->cfg.n_freq_masks > 0 && (!out->fmask_start || !out->fmask_len)) return fail(WWF_ERR_INVALID, "wwf_draw_aug: fmask arrays missing");
  if (p->cfg.n_time_masks > 0 && (!out->tmask_start || !out->tmask_len)) return fail(WWF_ERR_INVALID, "wwf_draw_aug: tmask arrays missing");
  if (p->n_noise > 0 && out->noise_idx && !out->noise_off) return fail(WWF_ERR_INVALID, "wwf_draw_aug: noise_off missing");
  DeviceGuard guard(p->device);
  if (!guard.ok) return fail(WWF_ERR_CUDA, "cudaSetDevice(%d) failed", p->device);
  DrawParams dp{};
  dp.seed = dc->seed; dp.first_index = first_index;
  dp.thr_rir = thr(dc->rir_prob); dp.thr_noise = thr(dc->noise_prob); dp.thr_fmask = thr(dc->freq_mask_prob); dp.thr_tmask = thr(dc->time_mask_prob);
  dp.snr_lo = dc->snr_lo; dp.snr_hi = dc->snr_hi;
  dp.B = B; dp.n_rir = p->n_rir; dp.n_noise = p->n_noise; dp.F = p->n_feat; dp.T = T;
  dp.fparam = dc->freq_mask_param; dp.tparam = dc->time_mask_param; dp.nF = p->cfg.n_freq_masks; dp.nT = p->cfg.n_time_masks;
  dp.noise_offsets = p->d_noise_offsets;
  dp.rir_idx = (int32_t*)out->rir_idx; dp.noise_idx = (int32_t*)out->noise_idx; dp.noise_off = (int64_t*)out->noise_off; dp.snr_db = (float*)out->snr_db;
  dp.fs = (int32_t*)out->fmask_start; dp.fl = (int32_t*)out->fmask_len; dp.ts = (int32_t*)out->tmask_start; dp.tl = (int32_t*)out->tmask_len;
  dp.thr_stretch = thr(dc->stretch_prob); dp.thr_pitch = thr(dc->pitch_prob);
  dp.stretch_lo = dc->stretch_lo; dp.stretch_hi = dc->stretch_hi; dp.pitch_lo = dc->pitch_lo; dp.pitch_hi = dc->pitch_hi;
  dp.stretch_rate = (double*)out->stretch_rate; dp.pitch_steps = (int32_t*)out->pitch_steps;
  if (dp.stretch_rate && dc->stretch_prob > 0.0 && !(dc->stretch_lo >= 0.1 && dc->stretch_hi >= dc->stretch_lo))
    return fail(WWF_ERR_INVALID, "wwf_draw_aug: stretch range [%g, %g]", dc->stretch_lo, dc->stretch_hi);
  if (dp.pitch_steps && dc->pitch_prob > 0.0 && (dc->pitch_lo < -12 || dc->pitch_hi > 12 || dc->pitch_hi < dc->pitch_lo))
    return fail(WWF_ERR_INVALID, "wwf_draw_aug: pitch range [%d, %d] must lie in [-12, 12]", dc->pitch_lo, dc->pitch_hi);
  draw_aug_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dp);
  g_launches++;
  WWF_CUDA(cudaGetLastError());
  return WWF_OK;
}

extern "C" int wwf_peak_normalize(const float* wav, int B, int N, int64_t wav_stride, float* out, int64_t out_stride,
                                  int device, void* stream) {
  if (!wav || !out || B <= 0 || N <= 0 || wav_stride < N || out_stride < N) return fail(WWF_ERR_INVALID, "wwf_peak_normalize: bad argument");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(WWF_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  peak_normalize_kernel<<<B, 512, 0, (cudaStream_t)stream>>>(wav, wav_stride, out, out_stride, N);
  g_launches++;
  WWF_CUDA(cudaGetLastError());
  return WWF_OK;
}

extern "C" int wwf_spec_augment(void* spec, int dtype, int B, int F, int T, int64_t clip_stride,
                                const int32_t* fs, const int32_t* fl, int nF, const int32_t* ts, const int32_t* tl, int nT,
                                float mask_value, int device, void* stream) {
  if (!spec || B <= 0 || F <= 0 || T <= 0 || clip_stride < (int64_t)F * T) return fail(WWF_ERR_INVALID, "wwf_spec_augment: bad argument");
  if (nF < 0 || nT < 0 || (nF > 0 && (!fs || !fl)) || (nT > 0 && (!ts || !tl))) return fail(WWF_ERR_INVALID, "wwf_spec_augment: mask arrays missing");
  if (dtype != WWF_OUT_F32 && dtype != WWF_OUT_F16) return fail(WWF_ERR_INVALID, "wwf_spec_augment: dtype=%d", dtype);
  if (nF == 0 && nT == 0) return WWF_OK;
  DeviceGuard guard(device);
  if (!guard.ok) return fail(WWF_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(F < 64 ? F : 64, B);
  if (dtype == WWF_OUT_F32) spec_mask_kernel<float><<<grid, 256, 0, st>>>((float*)spec, B, F, T, clip_stride, fs, fl, nF, ts, tl, nT, mask_value);
  else spec_mask_kernel<__half><<<grid, 256, 0, st>>>((__half*)spec, B, F, T, clip_stride, fs, fl, nF, ts, tl, nT, mask_value);
  g_launches++;
  WWF_CUDA(cudaGetLastError());
  return WWF_OK;
}

// ------------------------------------------------------------------------------------------
// time-stretch / pitch-shift / resample
// ------------------------------------------------------------------------------------------
static int ensure_pv_constants_locked(wwf_plan* p) {
  if (p->d_pv_tw) return WWF_OK;
  std::vector<float> win(kPvN);
  for (int i = 0; i < kPvN; ++i) win[i] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * i / kPvN));
  std::vector<float2> tw;
  build_stft_twiddles<PvRad>(tw);
  int rc;
  if ((rc = upload(&p->d_pv_window, win))) return rc;
  WWF_CUDA(cudaFuncSetAttribute((const void*)pv_synth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSynSmemBytes));
  return upload(&p->d_pv_tw, tw);
}
static int ensure_pv_constants(wwf_plan* p) {
  std::lock_guard<std::mutex> lk(p->lazy_mu);
  return ensure_pv_constants_locked(p);
}

extern "C" int wwf_set_stretch_window(wwf_plan* p, const float* window) {
  if (!p || !window) return fail(WWF_ERR_INVALID, "wwf_set_stretch_window: null argument");
  DeviceGuard guard(p->device);
  if (!guard.ok) return fail(WWF_ERR_CUDA, "cudaSetDevice(%d) failed", p->device);
  std::lock_guard<std::mutex> lk(p->lazy_mu);
  int rc = ensure_pv_constants_locked(p);
  if (rc) return rc;
  // stretch kernels of this plan may still be reading the old window on some (non-blocking) stream: wait for the
  // device, replace the table, wait for the copy to land (a pageable-memory cudaMemcpy may return before it has)
  WWF_CUDA(cudaDeviceSynchronize());
  WWF_CUDA(cudaMemcpy(p->d_pv_window, window, kPvN * sizeof(float), cudaMemcpyHostToDevice));
  WWF_CUDA(cudaStreamSynchronize(0));
  return WWF_OK;
}

struct PvGeom {
  int T, Tcap, Lcap;
  size_t offS, offV, offW, total;
};
static PvGeom pv_geometry(int B, int N, double rate_lo) {
  PvGeom g{};
  g.T = N / kPvHop + 1;
  g.Tcap = (int)ceil((double)g.T / rate_lo);
  g.Lcap = (int)round_up4((int64_t)rint((double)N / rate_lo));
  g.offS = 0;
  g.offV = g.offS + (size_t)B * g.T * kPvPitch * sizeof(float2);
  g.offW = g.offV + (size_t)B * g.Tcap * kPvPitch * sizeof(float2);
  g.total = g.offW + (size_t)B * g.Lcap * sizeof(float);
  return g;
}

extern "C" size_t wwf_stretch_workspace_bytes(int B, int N, double rate_lo) {
  if (B <= 0 || N <= 0 || !(rate_lo >= 0.1)) return 0;
  return pv_geometry(B, N, rate_lo).total;
}

static int check_pv(const wwf_plan* p, const float* wav, int B, int N, int64_t wav_stride, const float* out, int64_t out_stride,
                    const void* ws, size_t ws_bytes, double rate_lo, const char* who) {
  if (!p || !wav || !out) return fail(WWF_ERR_INVALID, "%s: null argument", who);
  if (B <= 0 || wav_stride < N || out_stride < N) return fail(WWF_ERR_INVALID, "%s: bad shape B=%d N=%d", who, B, N);
  if (N <= kPvN / 2) return fail(WWF_ERR_INVALID, "%s: N=%d must exceed %d (reflect padding of the 512-point STFT)", who, N, kPvN / 2);
  if (N > (1 << 24)) return fail(WWF_ERR_UNSUPPORTED, "%s: N=%d > 2^24 samples", who, N);
  if (B > 65535) return fail(WWF_ERR_UNSUPPORTED, "%s: B=%d > 65535 clips per call", who, B);
  if (!(rate_lo >= 0.1)) return fail(WWF_ERR_INVALID, "%s: lower rate bound %g must be >= 0.1", who, rate_lo);
  const size_t need = wwf_stretch_workspace_bytes(B, N, rate_lo);
  if (!ws || ws_bytes < need) return fail(WWF_ERR_WORKSPACE, "%s: workspace too small: %zu < %zu bytes", who, ws_bytes, need);
  if (reinterpret_cast<uintptr_t>(ws) & 15) return fail(WWF_ERR_WORKSPACE, "%s: workspace must be 16-byte aligned", who);
  return WWF_OK;
}

// STFT -> phase vocoder -> inverse STFT frames -> overlap-add; the last kernel writes n_out samples per clip to dst
static int launch_stretch(wwf_plan* p, PvParams& pp, const PvGeom& g, void* ws, cudaStream_t st) {
  int rc = ensure_pv_constants(p);
  if (rc) return rc;
  char* w = (char*)ws;
  pp.T = g.T; pp.Tcap = g.Tcap; pp.Lcap = g.Lcap;
  pp.window = p->d_pv_window; pp.tw = p->d_pv_tw;
  pp.S = (float2*)(w + g.offS); pp.V = (float2*)(w + g.offV);
  const int per_cta = 2 * kPvWarps;
  pv_stft_kernel<<<dim3((g.T + per_cta - 1) / per_cta, pp.B), kPvWarps * 32, 0, st>>>(pp);
  pv_vocoder_kernel<<<pp.B, 288, 0, st>>>(pp);
  const int syn_per_cta = kSynBlocks * kPvHop;
  pv_synth_kernel<<<dim3((pp.n_out + syn_per_cta - 1) / syn_per_cta, pp.B), kSynWarps * 32, kSynSmemBytes, st>>>(pp);
  g_launches += 3;
  WWF_CUDA(cudaGetLastError());
  return WWF_OK;
}

extern "C" int wwf_time_stretch(wwf_plan* p, const float* wav, int B, int N, int64_t wav_stride, const double* rates, double rate_lo,
                                float* out, int64_t out_stride, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_pv(p, wav, B, N, wav_stride, out, out_stride, workspace, workspace_bytes, rate_lo, "wwf_time_stretch");
  if (rc) return rc;
  if (!rates) return fail(WWF_ERR_INVALID, "wwf_time_stretch: rates is null");
  DeviceGuard guard(p->device);
  if (!guard.ok) return fail(WWF_ERR_CUDA, "cudaSetDevice(%d) failed", p->device);
  const PvGeom g = pv_geometry(B, N, rate_lo);
  PvParams pp{};
  pp.wav = wav; pp.wav_stride = wav_stride; pp.B = B; pp.N = N;
  pp.rate.rates = rates;
  pp.out = out; pp.out_stride = out_stride; pp.n_out = N; pp.out_pad = 1;
  return launch_stretch(p, pp, g, workspace, (cudaStream_t)stream);
}

// coefficient table of one (orig, new) ratio, frequencies already divided by their gcd
static int get_resampler(wwf_plan* p, int orig, int nw, cudaStream_t st, ResampleDesc* out) {
  std::lock_guard<std::mutex> lk(p->lazy_mu);
  for (const auto& r : p->resamplers)
    if (r.orig == orig && r.nw == nw) { *out = ResampleDesc{r.coef, r.orig, r.nw, r.width, r.ntaps}; return WWF_OK; }
  // _get_sinc_resample_kernel (TA/functional/functional.py:1343-1350): lowpass_filter_width 6, rolloff 0.99
  const double base_freq = (double)(orig < nw ? orig : nw) * 0.99;
  const int width = (int)ceil(6.0 * orig / base_freq);
  const int ntaps = 2 * width + 1;
  if ((int64_t)ntaps * nw > (int64_t)64 << 20) return fail(WWF_ERR_UNSUPPORTED, "resample %d:%d needs a %lld-entry table", orig, nw, (long long)ntaps * nw);
  float* coef = nullptr;
  WWF_CUDA(cudaMalloc((void**)&coef, ((size_t)ntaps * nw + nw) * sizeof(float)));   // + int32 first[nw]
  resample_table_kernel<<<(ntaps * nw + 255) / 256, 256, 0, st>>>(coef, orig, nw, width, ntaps, (float)base_freq, (float)(base_freq / orig));
  g_launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { cudaFree(coef); return fail(WWF_ERR_CUDA, "resample_table_kernel: %s", cudaGetErrorString(e)); }
  p->resamplers.push_back({orig, nw, width, ntaps, coef});
  *out = ResampleDesc{coef, orig, nw, width, ntaps};
  return WWF_OK;
}

static int gcd_int(int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; }

// ceil(new * n / orig) exactly as torchaudio evaluates it: the quotient passes through float32
// (torch.as_tensor(python float)), TA/functional/functional.py:1427
static int resample_target(int n_in, int orig, int nw) {
  return (int)ceilf((float)((double)((int64_t)nw * n_in) / (double)orig));
}

extern "C" int wwf_resample_length(int n_in, int orig_freq, int new_freq) {
  if (n_in < 0 || orig_freq <= 0 || new_freq <= 0) return fail(WWF_ERR_INVALID, "wwf_resample_length: bad argument");
  if (orig_freq == new_freq) return n_in;
  const int g = gcd_int(orig_freq, new_freq);
  return resample_target(n_in, orig_freq / g, new_freq / g);
}

extern "C" int wwf_resample(wwf_plan* p, const float* in, int B, int n_in, int64_t in_stride, int orig_freq, int new_freq,
                            float* out, int n_out, int64_t out_stride, void* stream) {
  if (!p || !in || !out) return fail(WWF_ERR_INVALID, "wwf_resample: null argument");
  if (B <= 0 || n_in <= 0 || n_out <= 0 || in_stride < n_in || out_stride < n_out || orig_freq <= 0 || new_freq <= 0)
    return fail(WWF_ERR_INVALID, "wwf_resample: bad argument (B=%d n_in=%d n_out=%d %d->%d Hz)", B, n_in, n_out, orig_freq, new_freq);
  if (in == out) return fail(WWF_ERR_INVALID, "wwf_resample: in-place is not allowed");
  if (B > 65535) return fail(WWF_ERR_UNSUPPORTED, "wwf_resample: B=%d > 65535 clips per call", B);
  DeviceGuard guard(p->device);
  if (!guard.ok) return fail(WWF_ERR_CUDA, "cudaSetDevice(%d) failed", p->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (orig_freq == new_freq) {
    WWF_CUDA(cudaMemcpy2DAsync(out, out_stride * sizeof(float), in, in_stride * sizeof(float), (size_t)(n_in < n_out ? n_in : n_out) * sizeof(float), B,
                               cudaMemcpyDeviceToDevice, st));
    return WWF_OK;
  }
  const int g = gcd_int(orig_freq, new_freq);
  ResampleParams rp{};
  int rc = get_resampler(p, orig_freq / g, new_freq / g, st, &rp.desc[0]);
  if (rc) return rc;
  rp.in = in; rp.in_stride = in_stride; rp.out = out; rp.out_stride = out_stride;
  rp.B = B; rp.n_in = n_in; rp.n_out = n_out;
  rp.in_len[0] = n_in;
  rp.target[0] = resample_target(n_in, orig_freq / g, new_freq / g);    // samples beyond it are written as zeros
  resample_kernel<<<dim3((n_out + 255) / 256, B), 256, 0, st>>>(rp);
  g_launches++;
  WWF_CUDA(cudaGetLastError());
  return WWF_OK;
}

static double pitch_rate(int n_steps) { return pow(2.0, -(double)n_steps / 12.0); }   // F.pitch_shift: 2.0 ** (-float(n) / 12)

extern "C" size_t wwf_pitch_workspace_bytes(int B, int N, int step_lo, int step_hi) {
  if (step_hi < step_lo || step_lo < -12 || step_hi > 12) return 0;
  return wwf_stretch_workspace_bytes(B, N, pitch_rate(step_hi > 0 ? step_hi : 0));
}

extern "C" int wwf_pitch_shift(wwf_plan* p, const float* wav, int B, int N, int64_t wav_stride, const int32_t* n_steps, int step_lo, int step_hi,
                               float* out, int64_t out_stride, void* workspace, size_t workspace_bytes, void* stream) {
  if (step_hi < step_lo || step_lo < -12 || step_hi > 12) return fail(WWF_ERR_INVALID, "wwf_pitch_shift: step range [%d, %d] must lie in [-12, 12]", step_lo, step_hi);
  const double rate_lo = pitch_rate(step_hi > 0 ? step_hi : 0);
  int rc = check_pv(p, wav, B, N, wav_stride, out, out_stride, workspace, workspace_bytes, rate_lo, "wwf_pitch_shift");
  if (rc) return rc;
  if (!n_steps) return fail(WWF_ERR_INVALID, "wwf_pitch_shift: n_steps is null");
  DeviceGuard guard(p->device);
  if (!guard.ok) return fail(WWF_ERR_CUDA, "cudaSetDevice(%d) failed", p->device);
  cudaStream_t st = (cudaStream_t)stream;
  const PvGeom g = pv_geometry(B, N, rate_lo);
  const int sr = p->cfg.sample_rate, ns = step_hi - step_lo + 1;
  PvParams pp{};
  ResampleParams rp{};
  pp.rate.steps = n_steps; pp.rate.step_lo = step_lo; pp.rate.n_steps = ns;
  rp.steps = n_steps; rp.step_lo = step_lo; rp.n_steps = ns;
  for (int i = 0; i < ns; ++i) {
    const int n = step_lo + i;
    const double rate = pitch_rate(n);
    pp.rate.rate_tab[i] = n == 0 ? 1.0 : rate;
    if (n == 0) continue;                                      // desc[i].coef stays null: clip passes through
    const int orig = (int)((double)sr / rate);                 // int(sample_rate / rate)
    const int gg = gcd_int(orig, sr);
    if ((rc = get_resampler(p, orig / gg, sr / gg, st, &rp.desc[i]))) return rc;
    rp.in_len[i] = (int)rint((double)N / rate);
    rp.target[i] = resample_target(rp.in_len[i], orig / gg, sr / gg);
  }
  float* W = (float*)((char*)workspace + g.offW);
  pp.wav = wav; pp.wav_stride = wav_stride; pp.B = B; pp.N = N;
  pp.out = W; pp.out_stride = g.Lcap; pp.n_out = g.Lcap; pp.out_pad = 0;
  if ((rc = launch_stretch(p, pp, g, workspace, st))) return rc;
  rp.in = W; rp.in_stride = g.Lcap; rp.out = out; rp.out_stride = out_stride;
  rp.B = B; rp.n_in = g.Lcap; rp.n_out = N;
  rp.wav = wav; rp.wav_stride = wav_stride;
  resample_kernel<<<dim3((N + 255) / 256, B), 256, 0, st>>>(rp);
  g_launches++;
  WWF_CUDA(cudaGetLastError());
  return WWF_OK;
}
