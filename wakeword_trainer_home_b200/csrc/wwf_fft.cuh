// wwf_fft.cuh - register-level DFT butterflies and the in-place mixed-radix pass used by
// both FFT users of the path: the STFT (A4, torch.stft semantics) and the RIR overlap-save
// convolution (A2, torchaudio fftconvolve semantics).  SURVEY.md section 8a.
//
// Everything here is __host__ __device__ so tests/emul can run the exact task functions
// sequentially on the CPU to check index math (no product code path runs on the CPU).
//
// Conventions: forward transform uses w_N = exp(-2*pi*i/N).  A length-n FFT is a list of
// radix passes R_0, R_1, ... with prod R_i = n, executed as in-place decimation in
// frequency: pass i works on sub-transforms of length L_i = n / (R_0 ... R_{i-1}), combines
// elements at stride s_i = L_i / R_i and multiplies output r by w_{L_i}^{j r}.  The result
// is left in digit-reversed order: frequency k = d_0 + R_0 d_1 + R_0 R_1 d_2 + ... sits at
// position p = d_0 s_0 + d_1 s_1 + ... .  The inverse runs the adjoint passes in reverse
// order (decimation in time) and therefore consumes exactly that order - a convolution
// never has to reorder.
#pragma once
#include <cuda_runtime.h>
#include <type_traits>

#define WWF_HD __host__ __device__ __forceinline__

namespace wwf {

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute
// may start while its predecessor drains; it must call pdl_wait() before touching the predecessor's results.
// (Triggering the early start explicitly from inside the predecessor was measured slower: the waiting CTAs take
// registers and shared memory from the kernel that is still running.)  A no-op for a normal launch and for the
// host-side emulation build of the tests.
__device__ __forceinline__ void pdl_wait() {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 900)
  cudaGridDependencySynchronize();
#endif
}


// ----------------------------------------------------------------------------------------
// compile-time helpers
// ----------------------------------------------------------------------------------------
template <int I, int N, class F>
WWF_HD void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

constexpr double kPi = 3.14159265358979323846264338327950288;

// flag bits of a mel lane-schedule task (built on the host in wwf_tables.h, consumed in wwf_feat.cuh)
constexpr int kMelOwner = 1;      // this lane finalises the filter (dB, store)
constexpr int kMelPartner = 2;    // add the partial sum of lane ^ 1 first

// sin/cos by Taylor series around 0 after folding the argument into [-pi/4, pi/4];
// evaluated only at compile time (double precision, then rounded to float once).
constexpr double cx_sin_core(double x) {
  double x2 = x * x, term = x, sum = x;
  for (int i = 1; i < 14; ++i) {
    term *= -x2 / double((2 * i) * (2 * i + 1));
    sum += term;
  }
  return sum;
}
constexpr double cx_cos_core(double x) {
  double x2 = x * x, term = 1.0, sum = 1.0;
  for (int i = 1; i < 14; ++i) {
    term *= -x2 / double((2 * i - 1) * (2 * i));
    sum += term;
  }
  return sum;
}
// cos(2 pi e / n), sin(2 pi e / n) with exact octant folding on the integer ratio.
constexpr double cx_cos2pi(long long e, long long n) {
  e %= n; if (e < 0) e += n;
  if (2 * e > n) e = n - e;                 // cos is even around pi
  if (4 * e > n) return -cx_cos2pi(n - 2 * e, 2 * n);  // cos(pi - t) = -cos t, t = 2pi(n-2e)/(2n)
  if (8 * e > n) return cx_sin_core(2.0 * kPi * double(n - 4 * e) / double(4 * n));  // cos(pi/2 - t) = sin t
  return cx_cos_core(2.0 * kPi * double(e) / double(n));
}
constexpr double cx_sin2pi(long long e, long long n) {
  // sin(a) = cos(a - pi/2) = cos(2 pi (4e - n) / (4n))
  return cx_cos2pi(4 * e - n, 4 * n);
}

// w_N^E = exp(-2 pi i E / N) = (c, -s)
template <int E, int N>
struct TwC {
  static constexpr int e = ((E % N) + N) % N;
  static constexpr float c = float(cx_cos2pi(e, N));
  static constexpr float s = float(cx_sin2pi(e, N));
};

// ----------------------------------------------------------------------------------------
// complex helpers
//
// On sm_100a every complex operation is issued as Blackwell's packed fp32x2 instructions
// (FADD2 / FMUL2 / FFMA2): one instruction works on the (re, im) register pair, a scalar operand
// is broadcast (.F32), and the operand swizzles .LO_HI / .NP give "swap halves" and "negate one
// half" for free - so a complex add, a multiply by +-i folded into an add, and a full complex
// multiply cost 1, 1 and 2 instructions instead of 2, 2 and 4.  Both kernels are
// instruction-issue bound with mostly complex arithmetic, which makes this the main lever.
// The host versions (tests/emul) perform the same roundings in the same order: results are
// bit-identical.
// ----------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
#define WWF_PACKED_F32X2 1
#endif

WWF_HD float2 cadd(float2 a, float2 b) {
#ifdef WWF_PACKED_F32X2
  return __fadd2_rn(a, b);
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
WWF_HD float2 csub(float2 a, float2 b) {
#ifdef WWF_PACKED_F32X2
  return __fadd2_rn(a, make_float2(-b.x, -b.y));
#else
  return make_float2(a.x - b.x, a.y - b.y);
#endif
}
// a * s + c with real scalar s
WWF_HD float2 cfma_s(float2 a, float s, float2 c) {
#ifdef WWF_PACKED_F32X2
  return __ffma2_rn(a, make_float2(s, s), c);
#else
  return make_float2(fmaf(a.x, s, c.x), fmaf(a.y, s, c.y));
#endif
}
WWF_HD float2 cmul_s(float2 a, float s) {
#ifdef WWF_PACKED_F32X2
  return __fmul2_rn(a, make_float2(s, s));
#else
  return make_float2(a.x * s, a.y * s);
#endif
}
WWF_HD float2 cmul(float2 a, float2 b) {   // (a.x b.x - a.y b.y, a.x b.y + a.y b.x)
#ifdef WWF_PACKED_F32X2
  return __ffma2_rn(b, make_float2(a.x, a.x), __fmul2_rn(make_float2(b.y, b.x), make_float2(-a.y, a.y)));
#else
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
#endif
}
WWF_HD float2 cmulc(float2 a, float2 b) {  // a * conj(b) = (a.x b.x + a.y b.y, a.y b.x - a.x b.y)
#ifdef WWF_PACKED_F32X2
  return __ffma2_rn(a, make_float2(b.x, b.x), __fmul2_rn(make_float2(a.y, a.x), make_float2(b.y, -b.y)));
#else
  return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
#endif
}
WWF_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
template <bool INV>
WWF_HD float2 mul_mi(float2 a) {  // forward: * (-i); inverse: * (+i)   (folds into the consumer's swizzle)
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
// (v.x c + v.y s, v.y c - v.x s) = v * (c - i s);  SGN = -1 gives v * (c + i s)
template <bool CONJ>
WWF_HD float2 cmul_cs(float2 v, float c, float s) {
#ifdef WWF_PACKED_F32X2
  return CONJ ? __ffma2_rn(v, make_float2(c, c), __fmul2_rn(make_float2(v.y, v.x), make_float2(-s, s)))
              : __ffma2_rn(v, make_float2(c, c), __fmul2_rn(make_float2(v.y, v.x), make_float2(s, -s)));
#else
  return CONJ ? make_float2(fmaf(v.x, c, -v.y * s), fmaf(v.y, c, v.x * s))
              : make_float2(fmaf(v.x, c, v.y * s), fmaf(v.y, c, -v.x * s));
#endif
}

// v * w_N^E (forward) or v * conj(w_N^E) (inverse), E and N compile-time.
template <int E, int N, bool INV>
WWF_HD float2 twmul(float2 v) {
  constexpr int e = ((E % N) + N) % N;
  constexpr float r = 0.70710678118654752440f;
  if constexpr (e == 0) {
    return v;
  } else if constexpr (4 * e == N) {
    return mul_mi<INV>(v);
  } else if constexpr (2 * e == N) {
    return make_float2(-v.x, -v.y);
  } else if constexpr (4 * e == 3 * N) {
    return mul_mi<!INV>(v);
  } else if constexpr (8 * e == N) {      // fwd (1 - i)/sqrt2: ((x + y) r, (y - x) r)
    return cmul_s(cadd(v, mul_mi<INV>(v)), r);
  } else if constexpr (8 * e == 3 * N) {  // fwd (-1 - i)/sqrt2: ((y - x) r, (-x - y) r)
    return cmul_s(csub(mul_mi<INV>(v), v), r);
  } else if constexpr (8 * e == 5 * N) {  // fwd (-1 + i)/sqrt2: ((-x - y) r, (x - y) r)
    return cmul_s(cadd(v, mul_mi<INV>(v)), -r);
  } else if constexpr (8 * e == 7 * N) {  // fwd (1 + i)/sqrt2: ((x - y) r, (x + y) r)
    return cmul_s(csub(v, mul_mi<INV>(v)), r);
  } else {
    return cmul_cs<INV>(v, TwC<e, N>::c, TwC<e, N>::s);   // forward w = (c, -s)
  }
}

// ----------------------------------------------------------------------------------------
// elementary butterflies on individual registers
// ----------------------------------------------------------------------------------------
template <bool INV>
WWF_HD void dft2(float2& a0, float2& a1) {
  float2 t = a0;
  a0 = cadd(t, a1);
  a1 = csub(t, a1);
}

template <bool INV>
WWF_HD void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_mi<INV>(csub(a1, a3));
  a0 = cadd(t0, t2);
  a2 = csub(t0, t2);
  a1 = cadd(t1, t3);
  a3 = csub(t1, t3);
}

template <bool INV>
WWF_HD void dft5(float2& a0, float2& a1, float2& a2, float2& a3, float2& a4) {
  constexpr float c1 = TwC<1, 5>::c, c2 = TwC<2, 5>::c, s1 = TwC<1, 5>::s, s2 = TwC<2, 5>::s;
  const float2 p1 = cadd(a1, a4), m1 = csub(a1, a4), p2 = cadd(a2, a3), m2 = csub(a2, a3);
  const float2 r1 = cadd(a0, cfma_s(p1, c1, cmul_s(p2, c2)));
  const float2 r2 = cadd(a0, cfma_s(p1, c2, cmul_s(p2, c1)));
  const float2 q1 = cfma_s(m1, s1, cmul_s(m2, s2));
  const float2 q2 = cfma_s(m1, s2, cmul_s(m2, -s1));
  const float2 iq1 = mul_mi<INV>(q1), iq2 = mul_mi<INV>(q2);  // forward: -i q
  a0 = cadd(cadd(a0, p1), p2);
  a1 = cadd(r1, iq1);
  a4 = csub(r1, iq1);
  a2 = cadd(r2, iq2);
  a3 = csub(r2, iq2);
}

template <int R, bool INV>
WWF_HD void dft(float2 (&v)[R]);

// N = A*B point DFT in registers (Cooley-Tukey, natural-order in and out):
//   X[B r1 + r2] = sum_{q1} w_A^{q1 r1} [ w_N^{q1 r2} sum_{q2} x[q1 + A q2] w_B^{q2 r2} ]
// Sub-transforms gather/scatter through small register arrays (free after unrolling), so A and B
// may themselves be composite.
template <int A, int B, bool INV>
WWF_HD void dft_composite(float2 (&v)[A * B]) {
  static_for<0, A>([&](auto Q1) {
    constexpr int q1 = decltype(Q1)::value;
    float2 t[B];
#pragma unroll
    for (int q2 = 0; q2 < B; ++q2) t[q2] = v[q1 + A * q2];
    dft<B, INV>(t);
#pragma unroll
    for (int q2 = 0; q2 < B; ++q2) v[q1 + A * q2] = t[q2];
  });
  static_for<1, A>([&](auto Q1) {
    static_for<1, B>([&](auto R2) {
      constexpr int q1 = decltype(Q1)::value, r2 = decltype(R2)::value;
      v[q1 + A * r2] = twmul<q1 * r2, A * B, INV>(v[q1 + A * r2]);
    });
  });
  float2 o[A * B];
  static_for<0, B>([&](auto R2) {
    constexpr int r2 = decltype(R2)::value;
    float2 t[A];
#pragma unroll
    for (int q1 = 0; q1 < A; ++q1) t[q1] = v[q1 + A * r2];
    dft<A, INV>(t);
#pragma unroll
    for (int r1 = 0; r1 < A; ++r1) o[B * r1 + r2] = t[r1];
  });
#pragma unroll
  for (int i = 0; i < A * B; ++i) v[i] = o[i];
}

template <int R, bool INV>
WWF_HD void dft(float2 (&v)[R]) {
  if constexpr (R == 2) dft2<INV>(v[0], v[1]);
  else if constexpr (R == 4) dft4<INV>(v[0], v[1], v[2], v[3]);
  else if constexpr (R == 5) dft5<INV>(v[0], v[1], v[2], v[3], v[4]);
  else if constexpr (R == 8) dft_composite<2, 4, INV>(v);
  else if constexpr (R == 16) dft_composite<4, 4, INV>(v);
  else if constexpr (R == 25) dft_composite<5, 5, INV>(v);
  else if constexpr (R == 32) dft_composite<4, 8, INV>(v);
  else static_assert(R == 2, "unsupported radix");
}

// ----------------------------------------------------------------------------------------
// One radix-R butterfly task of an in-place pass over one FFT stored at z (through the
// index map `Map`, e.g. shared-memory padding).
//   L  : current sub-transform length, s = L / R the element stride
//   u  : task id in [0, n/R): sub-transform u / s, column j = u % s
//   tw : this pass's table, tw[(r-1)*s + j] = w_L^{j r}, r = 1..R-1 (forward values; the
//        inverse pass multiplies by their conjugates); ignored when s == 1.
// Forward (DIF): v' [r] = w_L^{j r} * DFT_R(v)[r].   Inverse (DIT): v = IDFT_R(conj(tw) * v').
// ----------------------------------------------------------------------------------------
struct IdentityMap {
  WWF_HD constexpr int operator()(int i) const { return i; }
};
// One padding element after every 16: makes the power-of-two strides of radix-2^k passes hit 16
// distinct 8-byte banks per half-warp (stride 16 -> 17, 64 -> 68, 4 -> 4 + carry, ...).
struct PadMap {
  WWF_HD constexpr int operator()(int i) const { return i + (i >> 4); }
};
// Two-level padding: additionally one element per 256.  Keeps the radix passes conflict-free and
// also spreads the digit-reversed gather of the spectrum split (consecutive k -> positions
// 64 apart for 1024 = 16.16.4, 128 apart for 2048, 32 apart for 512) over 16 distinct banks.
struct PadMap2 {
  WWF_HD constexpr int operator()(int i) const { return i + (i >> 4) + (i >> 8); }
};

// Output twiddles of a forward radix-16 task (v[r] *= w_L^{j r}): six loaded powers (w, w^2, w^3 and w^4, w^8, w^12)
// and nine products w^{4a+b} = w^{4a} w^b instead of fifteen loads: the STFT kernels are bound by the shared-memory
// pipe before the FMA pipe.  Every derived twiddle is the product of two correctly rounded table entries (1.5 ulp; a
// chain w^8 w^4 w^2 w would reach 3.5 ulp and pushed the CMVN test, which divides by a row's standard deviation, past
// its bound).  tw[(r-1) s + j] = w_L^{j r}.
template <class TwLoad>
WWF_HD void twiddle16(float2 (&v)[16], int s, int j, TwLoad twload) {
  float2 w[16];
  w[1] = twload(j); w[2] = twload(s + j); w[3] = twload(2 * s + j);
  w[4] = twload(3 * s + j); w[8] = twload(7 * s + j); w[12] = twload(11 * s + j);
#pragma unroll
  for (int a = 1; a < 4; ++a)
#pragma unroll
    for (int b = 1; b < 4; ++b) w[4 * a + b] = cmul(w[4 * a], w[b]);
#pragma unroll
  for (int r = 1; r < 16; ++r) v[r] = cmul(v[r], w[r]);
}

template <int R, bool INV, class Map = IdentityMap, class TwLoad>
WWF_HD void pass_task(float2* z, int L, int u, TwLoad twload, Map map = Map()) {
  const int s = L / R;
  const int blk = u / s;
  const int j = u - blk * s;
  const int base = blk * L + j;
  float2 v[R];
#pragma unroll
  for (int q = 0; q < R; ++q) v[q] = z[map(base + q * s)];
  if constexpr (!INV) {
    dft<R, false>(v);
    if (s > 1) {
      if constexpr (R == 16) {
        twiddle16(v, s, j, twload);
      } else {
#pragma unroll
        for (int r = 1; r < R; ++r) v[r] = cmul(v[r], twload((r - 1) * s + j));
      }
    }
  } else {
    if (s > 1) {
#pragma unroll
      for (int r = 1; r < R; ++r) v[r] = cmulc(v[r], twload((r - 1) * s + j));
    }
    dft<R, true>(v);
  }
#pragma unroll
  for (int q = 0; q < R; ++q) z[map(base + q * s)] = v[q];
}

// Last pass of a TWO-pass plan (R0, R) with the results stored in NATURAL frequency order: task u = d0 transforms the
// R contiguous elements u R .. u R + R - 1 (s = 1: no twiddles) and output r is frequency k = u + R0 r, written to
// zout[map(k)] instead of back in place.  Every task overwrites other tasks' inputs, so this is only valid when ALL
// tasks of the transform run in one converged warp instruction stream (loads of all lanes, sync(), stores of all
// lanes); the host emulation passes a snapshot as zin.  The spectrum split can then read Z[k] and Z[n - k] from
// consecutive addresses (no digit-reversal arithmetic, no bank conflicts).
template <int R, int R0, class Map = IdentityMap, class Sync>
WWF_HD void pass_task_natural(const float2* zin, float2* zout, int u, Sync sync, Map map = Map()) {
  float2 v[R];
#pragma unroll
  for (int q = 0; q < R; ++q) v[q] = zin[map(u * R + q)];
  dft<R, false>(v);
  sync();
#pragma unroll
  for (int r = 0; r < R; ++r) zout[map(u + R0 * r)] = v[r];
}

// ----------------------------------------------------------------------------------------
// Static description of an FFT as a list of radices (up to 4 passes).
// ----------------------------------------------------------------------------------------
template <int R0, int R1 = 1, int R2 = 1, int R3 = 1>
struct Radices {
  static constexpr int n = R0 * R1 * R2 * R3;
  static constexpr int npass = 1 + (R1 > 1) + (R2 > 1) + (R3 > 1);
  static constexpr int R(int i) { return i == 0 ? R0 : i == 1 ? R1 : i == 2 ? R2 : R3; }
  static constexpr int L(int i) {  // sub-transform length of pass i
    int l = n;
    for (int k = 0; k < i; ++k) l /= R(k);
    return l;
  }
  static constexpr int S(int i) { return L(i) / R(i); }
  // twiddle table size of pass i (entries) and offset of pass i in the concatenated table
  static constexpr int tw_size(int i) { return S(i) > 1 ? (R(i) - 1) * S(i) : 0; }
  static constexpr int tw_off(int i) {
    int o = 0;
    for (int k = 0; k < i; ++k) o += tw_size(k);
    return o;
  }
  static constexpr int tw_total = tw_off(npass);
  // position of frequency k after the forward passes
  static WWF_HD int pos(int k) {
    int p = 0;
#pragma unroll
    for (int i = 0; i < npass; ++i) {
      const int d = k % R(i);
      k /= R(i);
      p += d * S(i);
    }
    return p;
  }
};

}  // namespace wwf
