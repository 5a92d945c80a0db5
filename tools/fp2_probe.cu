// fp2_probe.cu - issue rate of Blackwell's packed fp32x2 instructions against their scalar forms (sm_100a).
// Answers: is a warp-wide FADD2 / FFMA2 one issue slot at full rate, or does it occupy the FMA pipe for two cycles?
// Each thread runs ILP independent dependency chains of one instruction kind; 148 x 4 CTAs x 256 threads.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp2_probe tools/fp2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND, int ILP>
__global__ void __launch_bounds__(256) probe(float2* out, int iters, float s) {
  float2 a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f + s);
  const float2 c = make_float2(s, 1.0f - s), d = make_float2(1.0f + s, 0.999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (KIND == 0) a[i] = __fadd2_rn(a[i], c);                       // FADD2
      if (KIND == 1) a[i] = __ffma2_rn(a[i], d, c);                    // FFMA2
      if (KIND == 2) { a[i].x = a[i].x + c.x; a[i].y = a[i].y + c.y; } // 2 x FADD
      if (KIND == 3) { a[i].x = fmaf(a[i].x, d.x, c.x); a[i].y = fmaf(a[i].y, d.y, c.y); }  // 2 x FFMA
      if (KIND == 4) a[i] = __fmul2_rn(a[i], d);                       // FMUL2
      if (KIND == 5) { a[i] = __fadd2_rn(a[i], c); a[i].x = fmaf(a[i].x, d.x, c.y); }       // FADD2 + FFMA interleaved
    }
  }
  float2 r = a[0];
#pragma unroll
  for (int i = 1; i < ILP; ++i) { r.x += a[i].x; r.y += a[i].y; }
  if (r.x == 12345.678f) out[threadIdx.x] = r;
}

template <int KIND, int ILP>
static void run(const char* name, int per_iter, int ctas_per_sm, float2* out) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<KIND, ILP><<<sms * ctas_per_sm, 256>>>(out, 100, 0.25f);
  cudaEventRecord(e0);
  probe<KIND, ILP><<<sms * ctas_per_sm, 256>>>(out, iters, 0.25f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double warp_instr = (double)sms * ctas_per_sm * 8 * iters * ILP * per_iter;
  const double cycles = ms * 1e-3 * khz * 1e3;
  printf("{\"kind\": \"%s\", \"ilp\": %d, \"warps_per_sm\": %d, \"ms\": %.4f, \"warp_instr_per_clk_per_smsp\": %.4f}\n", name, ILP,
         ctas_per_sm * 8, ms, warp_instr / cycles / sms / 4.0);
}

int main() {
  float2* out;
  cudaMalloc(&out, 4096);
  for (int occ = 1; occ <= 4; occ *= 2) {
    run<0, 8>("FADD2", 1, occ, out);
    run<1, 8>("FFMA2", 1, occ, out);
    run<4, 8>("FMUL2", 1, occ, out);
    run<2, 8>("FADD x2", 2, occ, out);
    run<3, 8>("FFMA x2", 2, occ, out);
    run<5, 8>("FADD2+FFMA", 2, occ, out);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
