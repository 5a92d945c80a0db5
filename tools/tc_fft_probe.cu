// tc_fft_probe.cu - go / no-go measurement for "put the STFT butterflies on the tensor cores" (VERDICT r1, item 1).
//
// MEASUREMENT TOOL, not product code.  Build + run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo \
//        -I wakeword_trainer_home_b200/csrc -o /tmp/tc_fft_probe tools/tc_fft_probe.cu && /tmp/tc_fft_probe
//
// Workload = the FFT part of feat_frames_kernel<400,5> on bench.py's batch: 1024 clips x 151 frames = 77 312 complex
// 400-point FFTs (two real frames per transform).  Three measurements:
//   1. hmma_rate   raw issue rate of mma.sync.m16n8k8 TF32 (the legacy warp-level tensor path): HMMA per second, all SMs.
//   2. simt_fft    the product's radix 16 x 25 butterflies (wwf_fft.cuh pass_task, packed fp32x2), global -> smem ->
//                  FFT -> global, nothing else.
//   3. tc_fft      the same transform as two small-DFT GEMM stages on mma.sync TF32 with hi/lo split operands
//                  (3 products, float32 accumulation), twiddles between the stages in SIMT:
//                      stage 1  [32 x 32] real form of F16  x  [32 x 56]  (two FFTs = 50 columns of 16 points)
//                      stage 2  [16 x 56] per FFT  x  [56 x 56] real form of F25
//                  = 56 + 98 = 154 HMMA x 3 = 462 HMMA per FFT pair (231 per FFT).
// Reported: microseconds per batch, nanoseconds per clip, max |error| of 10 log10 |X[k]|^2 against a float64 DFT on
// 64 sampled transforms, and the lower bound  231 HMMA x 77 312 / hmma_rate  that NO mma.sync implementation of this
// factorisation can beat.  The tcgen05 bound follows from MEASURED_PEAKS.json (profiles/README.md has the arithmetic).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "wwf_fft.cuh"

using namespace wwf;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int NF = 400, NB = 1024 * 151 / 2;   // 77 312 complex FFTs
using Rad = Radices<16, 25>;

__device__ __forceinline__ uint32_t tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void mma8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---- 1. raw HMMA rate ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hmma_rate_kernel(float* sink, int iters) {
  float acc[8][4];
  for (int j = 0; j < 8; ++j) for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
  uint32_t a[4] = {tf32(1.0f + threadIdx.x), tf32(2.0f), tf32(3.0f), tf32(0.5f)};
  const uint32_t b0 = tf32(0.25f), b1 = tf32(0.125f);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) mma8(acc[j], a, b0, b1);
  }
  float s = 0.f;
  for (int j = 0; j < 8; ++j) for (int e = 0; e < 4; ++e) s += acc[j][e];
  if (s == 12345.678f) sink[0] = s;
}

// ---- 2. SIMT FFT (the product's butterflies) ------------------------------------------------------------------
constexpr int SIMT_WARPS = 10, G = 2;
__global__ void __launch_bounds__(SIMT_WARPS * 32, 2) simt_fft_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                                                                     const float2* __restrict__ tw_g, int npairs) {
  extern __shared__ __align__(16) float2 sm[];
  float2* s_tw = sm;
  float2* z = sm + Rad::tw_total + (threadIdx.x >> 5) * G * NF;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < Rad::tw_total; i += blockDim.x) s_tw[i] = tw_g[i];
  __syncthreads();
  for (int pr = blockIdx.x * SIMT_WARPS + warp; pr < npairs; pr += gridDim.x * SIMT_WARPS) {
    const float2* src = in + (size_t)pr * G * NF;
    for (int i = lane; i < G * NF; i += 32) z[i] = src[i];
    __syncwarp();
    static_for<0, Rad::npass>([&](auto I) {
      constexpr int i = decltype(I)::value;
      constexpr int R = Rad::R(i), L = Rad::L(i), tasks = NF / R;
      const float2* tw = s_tw + Rad::tw_off(i);
      constexpr bool kPerFft = tasks < 32 && G * ((tasks + 31) / 32) == (G * tasks + 31) / 32;
      if constexpr (kPerFft) {
#pragma unroll 1
        for (int g = 0; g < G; ++g)
          if (lane < tasks) pass_task<R, false>(z + g * NF, L, lane, [&](int q) { return tw[q]; });
      } else {
#pragma unroll 1
        for (int u = lane; u < G * tasks; u += 32) {
          const int g = u / tasks, uu = u - g * tasks;
          pass_task<R, false>(z + g * NF, L, uu, [&](int q) { return tw[q]; });
        }
      }
      __syncwarp();
    });
    float2* dst = out + (size_t)pr * G * NF;
    for (int i = lane; i < G * NF; i += 32) dst[i] = z[i];
    __syncwarp();
  }
}

// ---- 3. tensor-core FFT: two small-DFT GEMM stages, split TF32 ------------------------------------------------
// Shared constants per CTA:  A1h/A1l [32][36]   real form of F16:   rows (re k1 | im k1), cols (re n1 | im n1)
//                            B2h/B2l [56][56]   real form of F25:   rows (re n2 | im n2 | pad), cols (re k2 | im k2 | pad)
//                            TW [16][25] float2 w_400^{k1 n2}
// Per warp:  X h/l [32][56]  stage-1 B operand: rows (re n1 | im n1), cols c = 25 f + n2 (f = FFT of the pair), 50..55 zero
//            Y h/l [2][16][60]  stage-2 A operand per FFT: rows k1, cols (re n2 | im n2 | pad)
constexpr int TC_WARPS = 6, PA1 = 36, PX = 56, PY = 60, PB2 = 56;
constexpr int TC_CONST_WORDS = 2 * 32 * PA1 + 2 * 56 * PB2 + 2 * 16 * 25;
constexpr int TC_WARP_WORDS = 2 * 32 * PX + 2 * 2 * 16 * PY;
__global__ void __launch_bounds__(TC_WARPS * 32, 1) tc_fft_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                                                                 const float* __restrict__ consts, int npairs) {
  extern __shared__ __align__(16) float smf[];
  uint32_t* A1h = reinterpret_cast<uint32_t*>(smf);
  uint32_t* A1l = A1h + 32 * PA1;
  uint32_t* B2h = A1l + 32 * PA1;
  uint32_t* B2l = B2h + 56 * PB2;
  float2* TW = reinterpret_cast<float2*>(B2l + 56 * PB2);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  uint32_t* Xh = reinterpret_cast<uint32_t*>(smf) + TC_CONST_WORDS + warp * TC_WARP_WORDS;
  uint32_t* Xl = Xh + 32 * PX;
  uint32_t* Yh = Xl + 32 * PX;
  uint32_t* Yl = Yh + 2 * 16 * PY;
  for (int i = threadIdx.x; i < TC_CONST_WORDS; i += blockDim.x) smf[i] = consts[i];
  for (int i = lane; i < TC_WARP_WORDS; i += 32) Xh[i] = 0;           // pads stay zero
  __syncthreads();
  for (int pr = blockIdx.x * TC_WARPS + warp; pr < npairs; pr += gridDim.x * TC_WARPS) {
    // load + split: x[f][25 n1 + n2] -> X[(re|im) n1][25 f + n2]
    const float2* src = in + (size_t)pr * G * NF;
    for (int i = lane; i < G * NF; i += 32) {
      const int f = i / NF, n = i - f * NF, n1 = n / 25, n2 = n - n1 * 25;
      const float2 v = src[i];
      const uint32_t hr = tf32(v.x), hi = tf32(v.y);
      const int c = 25 * f + n2;
      Xh[n1 * PX + c] = hr; Xl[n1 * PX + c] = tf32(v.x - __uint_as_float(hr));
      Xh[(16 + n1) * PX + c] = hi; Xl[(16 + n1) * PX + c] = tf32(v.y - __uint_as_float(hi));
    }
    __syncwarp();
    // stage 1: D[32 x 56] = A1[32 x 32] X[32 x 56]
    float acc[2][7][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 7; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int r0 = 16 * mt + g, c0 = 8 * ks + t;
        ah[mt][0] = A1h[r0 * PA1 + c0]; ah[mt][1] = A1h[(r0 + 8) * PA1 + c0]; ah[mt][2] = A1h[r0 * PA1 + c0 + 4]; ah[mt][3] = A1h[(r0 + 8) * PA1 + c0 + 4];
        al[mt][0] = A1l[r0 * PA1 + c0]; al[mt][1] = A1l[(r0 + 8) * PA1 + c0]; al[mt][2] = A1l[r0 * PA1 + c0 + 4]; al[mt][3] = A1l[(r0 + 8) * PA1 + c0 + 4];
      }
#pragma unroll
      for (int nt = 0; nt < 7; ++nt) {
        const int r = 8 * ks + t, c = 8 * nt + g;
        const uint32_t bh0 = Xh[r * PX + c], bh1 = Xh[(r + 4) * PX + c], bl0 = Xl[r * PX + c], bl1 = Xl[(r + 4) * PX + c];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma8(acc[mt][nt], al[mt], bh0, bh1);
          mma8(acc[mt][nt], ah[mt], bl0, bl1);
          mma8(acc[mt][nt], ah[mt], bh0, bh1);
        }
      }
    }
    // twiddle + split -> Y[f][k1][(re|im) n2]
#pragma unroll
    for (int nt = 0; nt < 7; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = 8 * nt + 2 * t + (e & 1), k1 = g + (e >= 2 ? 8 : 0);
        if (c < 50) {
          const int f = c >= 25, n2 = c - 25 * f;
          const float2 y = cmul(make_float2(acc[0][nt][e], acc[1][nt][e]), TW[k1 * 25 + n2]);
          const uint32_t hr = tf32(y.x), hi = tf32(y.y);
          uint32_t* yh = Yh + (f * 16 + k1) * PY;
          uint32_t* yl = Yl + (f * 16 + k1) * PY;
          yh[n2] = hr; yl[n2] = tf32(y.x - __uint_as_float(hr));
          yh[25 + n2] = hi; yl[25 + n2] = tf32(y.y - __uint_as_float(hi));
        }
      }
    __syncwarp();
    // stage 2 per FFT: D[16 x 56] = Y[16 x 56] B2[56 x 56]; X[k1 + 16 k2] -> position 25 k1 + k2 (the SIMT layout)
    float2* dst = out + (size_t)pr * G * NF;
#pragma unroll 1
    for (int f = 0; f < 2; ++f) {
      float d[7][4];
#pragma unroll
      for (int nt = 0; nt < 7; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) d[nt][e] = 0.f;
      const uint32_t* yh = Yh + f * 16 * PY;
      const uint32_t* yl = Yl + f * 16 * PY;
#pragma unroll
      for (int ks = 0; ks < 7; ++ks) {
        const int c0 = 8 * ks + t;
        const uint32_t ah[4] = {yh[g * PY + c0], yh[(g + 8) * PY + c0], yh[g * PY + c0 + 4], yh[(g + 8) * PY + c0 + 4]};
        const uint32_t al[4] = {yl[g * PY + c0], yl[(g + 8) * PY + c0], yl[g * PY + c0 + 4], yl[(g + 8) * PY + c0 + 4]};
#pragma unroll
        for (int nt = 0; nt < 7; ++nt) {
          const int r = 8 * ks + t, c = 8 * nt + g;
          const uint32_t bh0 = B2h[r * PB2 + c], bh1 = B2h[(r + 4) * PB2 + c], bl0 = B2l[r * PB2 + c], bl1 = B2l[(r + 4) * PB2 + c];
          mma8(d[nt], al, bh0, bh1);
          mma8(d[nt], ah, bl0, bl1);
          mma8(d[nt], ah, bh0, bh1);
        }
      }
      // D columns: (re k2 | im k2): re and im of one output live in different n-tiles / lanes -> scalar stores
      float* o = reinterpret_cast<float*>(dst + f * NF);
#pragma unroll
      for (int nt = 0; nt < 7; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = 8 * nt + 2 * t + (e & 1), k1 = g + (e >= 2 ? 8 : 0);
          if (c < 50) {
            const int im = c >= 25, k2 = c - 25 * im;
            o[2 * (25 * k1 + k2) + im] = d[nt][e];
          }
        }
    }
    __syncwarp();
  }
}

// ---- host ------------------------------------------------------------------------------------------------------
static float tf32_host(float x) {   // round to nearest, ties away (cvt.rna): add half ulp of the 13 dropped bits, truncate
  uint32_t u; memcpy(&u, &x, 4);
  u += 0x1000u; u &= 0xffffe000u;
  float r; memcpy(&r, &u, 4);
  return r;
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  printf("{\"device\": \"%s\", \"sms\": %d, \"n_fft\": %d, \"complex_ffts\": %d,\n", prop.name, sms, NF, NB);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto time_ms = [&](auto&& launch, int reps) { launch(); launch(); CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms / reps; };

  // 1. raw HMMA rate
  float* sink; CK(cudaMalloc(&sink, 4));
  const int iters = 4096;
  const float ms_h = time_ms([&] { hmma_rate_kernel<<<sms * 8, 256>>>(sink, iters); }, 5);
  const double hmma_per_s = (double)sms * 8 * 8 * iters * 8 / (ms_h * 1e-3);   // warps x 8 per iteration
  printf(" \"hmma_rate\": {\"hmma_per_s\": %.4g, \"tf32_dense_tflops\": %.1f, \"cycles_per_hmma_per_smsp_at_1965MHz\": %.2f},\n",
         hmma_per_s, hmma_per_s * 16 * 8 * 8 * 2 / 1e12, 1.965e9 * sms * 4 / hmma_per_s);

  // data
  std::vector<float2> h_in((size_t)NB * NF);
  srand(1);
  for (size_t i = 0; i < h_in.size(); ++i) {
    const int n = (int)(i % NF);
    const float w = 0.5f - 0.5f * cosf(2.0f * (float)M_PI * n / NF);
    auto rnd = [] { float s = 0.f; for (int k = 0; k < 6; ++k) s += (float)rand() / RAND_MAX - 0.5f; return s; };
    h_in[i] = make_float2(0.1f * w * rnd(), 0.1f * w * rnd());
  }
  float2 *d_in, *d_out1, *d_out2;
  const size_t bytes = h_in.size() * sizeof(float2);
  CK(cudaMalloc(&d_in, bytes)); CK(cudaMalloc(&d_out1, bytes)); CK(cudaMalloc(&d_out2, bytes));
  CK(cudaMemcpy(d_in, h_in.data(), bytes, cudaMemcpyHostToDevice));

  // 2. SIMT
  std::vector<float2> tw(Rad::tw_total);
  for (int i = 0; i < Rad::npass; ++i) {
    const int R = Rad::R(i), L = Rad::L(i), s = Rad::S(i);
    if (s <= 1) continue;
    for (int r = 1; r < R; ++r) for (int j = 0; j < s; ++j) {
      const double a = -2.0 * M_PI * (double)((long long)j * r % L) / (double)L;
      tw[Rad::tw_off(i) + (r - 1) * s + j] = make_float2((float)cos(a), (float)sin(a));
    }
  }
  float2* d_tw; CK(cudaMalloc(&d_tw, tw.size() * sizeof(float2)));
  CK(cudaMemcpy(d_tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
  const size_t sm_simt = (Rad::tw_total + SIMT_WARPS * G * NF) * sizeof(float2);
  CK(cudaFuncSetAttribute(simt_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_simt));
  const float ms_s = time_ms([&] { simt_fft_kernel<<<sms * 2, SIMT_WARPS * 32, sm_simt>>>(d_in, d_out1, d_tw, NB / G); }, 20);
  CK(cudaGetLastError());

  // 3. tensor cores
  std::vector<float> cst(TC_CONST_WORDS, 0.f);
  float* A1h = cst.data(); float* A1l = A1h + 32 * PA1; float* B2h = A1l + 32 * PA1; float* B2l = B2h + 56 * PB2; float* TWc = B2l + 56 * PB2;
  auto put = [&](float* hi, float* lo, int idx, double v) { const float f = (float)v, h = tf32_host(f); hi[idx] = h; lo[idx] = tf32_host(f - h); };
  for (int k = 0; k < 16; ++k) for (int n = 0; n < 16; ++n) {
    const double a = -2.0 * M_PI * ((k * n) % 16) / 16.0, c = cos(a), s = sin(a);      // F = c + i s
    put(A1h, A1l, k * PA1 + n, c); put(A1h, A1l, k * PA1 + 16 + n, -s);                 // re out = c re - s im
    put(A1h, A1l, (16 + k) * PA1 + n, s); put(A1h, A1l, (16 + k) * PA1 + 16 + n, c);     // im out = s re + c im
  }
  for (int n = 0; n < 25; ++n) for (int k = 0; k < 25; ++k) {
    const double a = -2.0 * M_PI * ((n * k) % 25) / 25.0, c = cos(a), s = sin(a);
    put(B2h, B2l, n * PB2 + k, c); put(B2h, B2l, n * PB2 + 25 + k, s);                  // [re n2 row]: re out += c re, im out += s re
    put(B2h, B2l, (25 + n) * PB2 + k, -s); put(B2h, B2l, (25 + n) * PB2 + 25 + k, c);   // [im n2 row]: re out -= s im, im out += c im
  }
  for (int k1 = 0; k1 < 16; ++k1) for (int n2 = 0; n2 < 25; ++n2) {
    const double a = -2.0 * M_PI * (k1 * n2) / 400.0;
    TWc[2 * (k1 * 25 + n2)] = (float)cos(a); TWc[2 * (k1 * 25 + n2) + 1] = (float)sin(a);
  }
  float* d_c; CK(cudaMalloc(&d_c, cst.size() * 4)); CK(cudaMemcpy(d_c, cst.data(), cst.size() * 4, cudaMemcpyHostToDevice));
  const size_t sm_tc = (size_t)(TC_CONST_WORDS + TC_WARPS * TC_WARP_WORDS) * 4;
  CK(cudaFuncSetAttribute(tc_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_tc));
  const float ms_t = time_ms([&] { tc_fft_kernel<<<sms, TC_WARPS * 32, sm_tc>>>(d_in, d_out2, d_c, NB / G); }, 10);
  CK(cudaGetLastError());

  // accuracy on 64 transforms against a float64 DFT
  std::vector<float2> o1(h_in.size()), o2(h_in.size());
  CK(cudaMemcpy(o1.data(), d_out1, bytes, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(o2.data(), d_out2, bytes, cudaMemcpyDeviceToHost));
  double e_simt = 0, e_tc = 0, r_simt = 0, r_tc = 0;
  for (int s = 0; s < 64; ++s) {
    const size_t fidx = (size_t)s * (NB / 64);
    const float2* x = &h_in[fidx * NF];
    double pmax = 0;
    std::vector<double> pr(NF), pi_(NF);
    for (int k = 0; k < NF; ++k) {
      double re = 0, im = 0;
      for (int n = 0; n < NF; ++n) { const double a = -2.0 * M_PI * ((long long)k * n % NF) / NF; re += x[n].x * cos(a) - x[n].y * sin(a); im += x[n].x * sin(a) + x[n].y * cos(a); }
      pr[k] = re; pi_[k] = im; pmax = fmax(pmax, re * re + im * im);
    }
    for (int k = 0; k < NF; ++k) {
      const int p = (k % 16) * 25 + k / 16;
      const double pw = pr[k] * pr[k] + pi_[k] * pi_[k];
      if (pw < pmax * 1e-8) continue;                          // below the 80 dB floor: clamped in the product
      const float2 a = o1[fidx * NF + p], b = o2[fidx * NF + p];
      const double da = fabs(10 * log10(((double)a.x * a.x + (double)a.y * a.y) / pw)), db = fabs(10 * log10(((double)b.x * b.x + (double)b.y * b.y) / pw));
      e_simt = fmax(e_simt, da); e_tc = fmax(e_tc, db);
      r_simt = fmax(r_simt, hypot(a.x - pr[k], a.y - pi_[k]) / sqrt(pmax)); r_tc = fmax(r_tc, hypot(b.x - pr[k], b.y - pi_[k]) / sqrt(pmax));
    }
  }
  const double bound_us = 231.0 * NB / hmma_per_s * 1e6;
  printf(" \"simt_fft\": {\"us_per_batch\": %.1f, \"ns_per_clip\": %.1f, \"max_err_db\": %.3g, \"max_err_rel_to_peak\": %.3g},\n", ms_s * 1e3, ms_s * 1e6 / 1024, e_simt, r_simt);
  printf(" \"tc_fft_mma_sync_3xtf32\": {\"us_per_batch\": %.1f, \"ns_per_clip\": %.1f, \"max_err_db\": %.3g, \"max_err_rel_to_peak\": %.3g, \"hmma_per_fft\": 231},\n", ms_t * 1e3, ms_t * 1e6 / 1024, e_tc, r_tc);
  printf(" \"mma_sync_lower_bound_us\": %.1f, \"io_bytes_per_batch\": %zu}\n", bound_us, 2 * bytes);
  return 0;
}
