// fft_core.cuh — the in-place shared-memory complex FFT used by k_fft_fast (spectrum.cu).
//
// Decimation in frequency with three stages for M = R1 * 16 * 32 points (R1 in {1,2,4,8,16,32}, so
// M = 512 .. 16384):
//   stage 1  radix R1 over the whole array        (butterfly legs M/R1 apart)
//   stage 2  radix 16 inside blocks of 512        (legs 32 apart)
//   stage 3  radix 32 on 32 contiguous points     (one thread each, no twiddles)
// Every butterfly is done in registers; the array is touched 3 times instead of log4(M) = 7.  The array
// lives at padded positions PAD(p) = p + p/32 + p/512, which keeps all three access patterns and the
// epilogue's consecutive-bin reads at the two-wavefront minimum for 8-byte accesses.  Results stay digit-reversed (fft_position) and are read back only at the
// bins the spectrum keeps.  Twiddles W_N^e come from a two-level table: hi[e >> 6] * lo[e & 63].
//
// All functions are __host__ __device__ so tools/test_fft_core.cu can run them on the CPU, with the
// thread loop emulated, against a double-precision DFT.
#pragma once

#include <cuda_runtime.h>

namespace ssb {

#define SSB_HD __host__ __device__ __forceinline__

// one pad slot per 32 points (stage 2/3 patterns) and one more per 512 (so that X[k], X[k+1], ... — which sit
// 512 points apart after the digit reversal — fall in different banks when the epilogue reads them)
SSB_HD unsigned fft_pad(unsigned p) { return p + (p >> 5) + (p >> 9); }

// Complex add / subtract are the bulk of a butterfly.  On sm_100 both halves go through one packed FP32 instruction
// (FADD2; a - b as FFMA2 b * (-1, -1) + a, exact), halving the issue slots of an issue-bound kernel; the host build
// (tests/test_fft_core_host.py) and SSB_FFT_SCALAR keep the scalar form — same IEEE results either way.
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000 && !defined(SSB_FFT_SCALAR)
SSB_HD float2 c_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
SSB_HD float2 c_sub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }
#else
SSB_HD float2 c_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
SSB_HD float2 c_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
SSB_HD float2 c_mul(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }

struct FftTwiddle {
  const float2* lo;  // W_N^i, i < 64 (or < N when N < 64)
  const float2* hi;  // W_N^(64 i), i < N/64
};

// W_N^e, e < N
SSB_HD float2 fft_twiddle(const FftTwiddle& t, unsigned e) {
  const float2 l = t.lo[e & 63u];
  const unsigned h = e >> 6;
  if (h == 0) return l;
  return c_mul(t.hi[h], l);
}

// cos/sin(2*pi*k/32), k = 0..8, correctly rounded
#define SSB_C32_1 0.98078528040323043f
#define SSB_C32_2 0.92387953251128674f
#define SSB_C32_3 0.83146961230254524f
#define SSB_C32_4 0.70710678118654752f
#define SSB_C32_5 0.55557023301960218f
#define SSB_C32_6 0.38268343236508977f
#define SSB_C32_7 0.19509032201612825f

// W_32^k = exp(-j*2*pi*k/32) for k = 0..15
SSB_HD float2 w32(int k) {
  switch (k) {
    case 0: return make_float2(1.f, 0.f);
    case 1: return make_float2(SSB_C32_1, -SSB_C32_7);
    case 2: return make_float2(SSB_C32_2, -SSB_C32_6);
    case 3: return make_float2(SSB_C32_3, -SSB_C32_5);
    case 4: return make_float2(SSB_C32_4, -SSB_C32_4);
    case 5: return make_float2(SSB_C32_5, -SSB_C32_3);
    case 6: return make_float2(SSB_C32_6, -SSB_C32_2);
    case 7: return make_float2(SSB_C32_7, -SSB_C32_1);
    case 8: return make_float2(0.f, -1.f);
    case 9: return make_float2(-SSB_C32_7, -SSB_C32_1);
    case 10: return make_float2(-SSB_C32_6, -SSB_C32_2);
    case 11: return make_float2(-SSB_C32_5, -SSB_C32_3);
    case 12: return make_float2(-SSB_C32_4, -SSB_C32_4);
    case 13: return make_float2(-SSB_C32_3, -SSB_C32_5);
    case 14: return make_float2(-SSB_C32_2, -SSB_C32_6);
    default: return make_float2(-SSB_C32_1, -SSB_C32_7);
  }
}

template <int R>
SSB_HD constexpr int bitrev(int i) {
  int r = 0;
  for (int b = 1; b < R; b <<= 1) { r = (r << 1) | (i & 1); i >>= 1; }
  return r;
}

// multiply by W_R^k for compile-time k, R: the trivial rotations cost no multiplies
template <int R, int K>
SSB_HD float2 rot(float2 a) {
  constexpr int k32 = K * (32 / R);  // as a power of W_32
  if (k32 == 0) return a;
  if (k32 == 8) return make_float2(a.y, -a.x);  // * (-j)
  if (k32 == 4) return make_float2((a.x + a.y) * SSB_C32_4, (a.y - a.x) * SSB_C32_4);
  if (k32 == 12) return make_float2((a.y - a.x) * SSB_C32_4, -(a.x + a.y) * SSB_C32_4);
  return c_mul(a, w32(k32));
}

// In-register DIF FFT of R points (R = 2..32); x[i] ends up holding X[bitrev<R>(i)].
template <int R, int LEN, int BASE, int J>
struct DifStage {
  // butterflies of the sub-transform of length LEN starting at BASE, leg index J
  SSB_HD static void run(float2 (&x)[R]) {
    if constexpr (J < LEN / 2) {
      const float2 a = x[BASE + J], b = x[BASE + J + LEN / 2];
      x[BASE + J] = c_add(a, b);
      x[BASE + J + LEN / 2] = rot<LEN, J>(c_sub(a, b));
      DifStage<R, LEN, BASE, J + 1>::run(x);
    }
  }
};
template <int R, int LEN, int BASE>
struct DifBlock {
  SSB_HD static void run(float2 (&x)[R]) {
    if constexpr (LEN >= 2) {
      DifStage<R, LEN, BASE, 0>::run(x);
      DifBlock<R, LEN / 2, BASE>::run(x);
      DifBlock<R, LEN / 2, BASE + LEN / 2>::run(x);
    }
  }
};
template <int R>
SSB_HD void fft_regs(float2 (&x)[R]) {
  DifBlock<R, R, 0>::run(x);
}

// w[p] = W^(p) for p = 1..R-1 from the base twiddle W (= w1): the binary powers p = 2, 4, 8, 16 by table
// lookup of W_N^(p*e1) (no error growth), every other power as a product of two of those entries' partial
// products (at most 3 rounded multiplies deep).
template <int R>
SSB_HD void twiddle_powers(const FftTwiddle& tw, unsigned e1, unsigned n_mask, float2 (&w)[R]) {
  // e1 < N / R-ish: p * e1 stays below N for every p < R by construction of the stages
  w[0] = make_float2(1.f, 0.f);
#pragma unroll
  for (int p = 1; p < R; p <<= 1) w[p] = fft_twiddle(tw, (p * e1) & n_mask);
#pragma unroll
  for (int p = 3; p < R; p++) {
    if ((p & (p - 1)) != 0) {          // not a power of two: top bit times the rest
      int top = 1;
      while ((top << 1) <= p) top <<= 1;
      w[p] = c_mul(w[top], w[p - top]);
    }
  }
}

// ---- stage 1: radix R1 over the whole array; butterfly j of M/R1 ----
template <int R1>
SSB_HD void fft_stage1(float2* z, unsigned M, unsigned N, const FftTwiddle& tw, unsigned j) {
  const unsigned Q = M / R1;               // multiple of 32 (M >= 1024 here)
  const unsigned QP = Q + (Q >> 5) + (Q >> 9);  // padded leg stride: PAD(j + q*Q) = PAD(j) + q*QP (Q = 512)
  float2* zj = z + fft_pad(j);
  float2 x[R1];
#pragma unroll
  for (int q = 0; q < R1; q++) x[q] = zj[q * QP];
  fft_regs<R1>(x);
  if (j != 0) {
    float2 w[R1];
    twiddle_powers<R1>(tw, j * (N / M), N - 1, w);  // W_M^(p j) = W_N^(p j N/M)
#pragma unroll
    for (int i = 1; i < R1; i++) x[i] = c_mul(x[i], w[bitrev<R1>(i)]);
  }
#pragma unroll
  for (int i = 0; i < R1; i++) zj[bitrev<R1>(i) * QP] = x[i];
}

// ---- stage 2: radix 16 inside blocks of 512; butterfly t of M/16 ----
SSB_HD void fft_stage2(float2* z, unsigned N, const FftTwiddle& tw, unsigned t) {
  const unsigned b = t >> 5, j = t & 31u;
  float2* zb = z + (b * 529u + j);         // PAD(b*512 + j + q*32) = b*529 + j + q*33
  float2 x[16];
#pragma unroll
  for (int q = 0; q < 16; q++) x[q] = zb[q * 33];
  fft_regs<16>(x);
  if (j != 0) {
    float2 w[16];
    twiddle_powers<16>(tw, j * (N >> 9), N - 1, w);  // W_512^(p j) = W_N^(p j N/512)
#pragma unroll
    for (int i = 1; i < 16; i++) x[i] = c_mul(x[i], w[bitrev<16>(i)]);
  }
#pragma unroll
  for (int i = 0; i < 16; i++) zb[bitrev<16>(i) * 33] = x[i];
}

// ---- stage 3: radix 32 on 32 contiguous points; butterfly t of M/32 ----
SSB_HD void fft_stage3(float2* z, unsigned t) {
  float2* zt = z + (t * 33u + (t >> 4));  // PAD(32 t + q) = 33 t + t/16 + q
  float2 x[32];
#pragma unroll
  for (int q = 0; q < 32; q++) x[q] = zt[q];
  fft_regs<32>(x);
#pragma unroll
  for (int i = 0; i < 32; i++) zt[bitrev<32>(i)] = x[i];
}

// padded position of X[k] after the three stages; r1_shift = log2(R1) = log2(M / 512)
SSB_HD unsigned fft_position(unsigned k, unsigned r1_shift) {
  const unsigned q1 = k & ((1u << r1_shift) - 1u), k1 = k >> r1_shift;
  const unsigned q2 = k1 & 15u, q3 = k1 >> 4;
  return q1 * 529u + q2 * 33u + q3;  // PAD(q1*512 + q2*32 + q3), q3 < 32
}

}  // namespace ssb
