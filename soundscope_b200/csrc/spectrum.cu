// spectrum.cu — Analyzer::get_fft (reference src/analyzer.rs:55-105), get_waveform (:107-137) and
// get_mid_and_side_samples (reference src/audio_player.rs:400-419) as CUDA kernels.
//
// k_fft<LAYOUT>: one CTA per window.  Load stage fuses de-interleave + mid/side (f32, exactly the
// reference's (l+r)/2, (l-r)/2) + Hann multiply (host-built table of spectrum-analyzer's f32
// multipliers); the transform is an in-place shared-memory radix-4 DIF (final radix-2 when log2 is odd)
// whose digit-reversed result is read back only at the kept bins; the store stage fuses the real
// split (two real spectra from one complex transform for mid/side; the packed N/2-point trick for
// mono), |X| -> scale_to_dbfs (analyzer.rs:11-27) and the spectrum-analyzer argument checks.
#include <stdlib.h>

#include "fft_core.cuh"
#include "ssb_internal.cuh"

namespace ssb {

__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
  return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
}

// W_N^k = exp(-j*2*pi*k/N) from the half table T[k], k < N/2
__device__ __forceinline__ float2 tw_n(const float2* __restrict__ T, unsigned k, unsigned half_n) {
  if (k < half_n) return __ldg(&T[k]);
  const float2 t = __ldg(&T[k - half_n]);
  return make_float2(-t.x, -t.y);
}

// position of X[k] after the in-place DIF with radices 4,4,...,(2): digit reversal
__device__ __forceinline__ unsigned dif_position(unsigned k, unsigned M) {
  unsigned p = 0, span = M;
  while (span >= 4) { span >>= 2; p += (k & 3u) * span; k >>= 2; }
  if (span == 2) p += (k & 1u);
  return p;
}

__device__ void fft_inplace(float2* z, unsigned M, unsigned N, const float2* __restrict__ T) {
  const unsigned tid = threadIdx.x, nt = blockDim.x;
  unsigned L = M;
  while (L >= 4) {
    const unsigned Q = L >> 2;
    const unsigned step = N / L;  // W_L^e = W_N^(e*step)
    for (unsigned t = tid; t < (M >> 2); t += nt) {
      const unsigned blk = t / Q, j = t - blk * Q;
      const unsigned base = blk * L + j;
      const float2 a0 = z[base], a1 = z[base + Q], a2 = z[base + 2 * Q], a3 = z[base + 3 * Q];
      const float2 t0 = make_float2(a0.x + a2.x, a0.y + a2.y);
      const float2 t1 = make_float2(a0.x - a2.x, a0.y - a2.y);
      const float2 t2 = make_float2(a1.x + a3.x, a1.y + a3.y);
      const float2 d = make_float2(a1.x - a3.x, a1.y - a3.y);
      const float2 t3 = make_float2(d.y, -d.x);  // -j * (a1 - a3)
      const float2 y0 = make_float2(t0.x + t2.x, t0.y + t2.y);
      const float2 y1 = make_float2(t1.x + t3.x, t1.y + t3.y);
      const float2 y2 = make_float2(t0.x - t2.x, t0.y - t2.y);
      const float2 y3 = make_float2(t1.x - t3.x, t1.y - t3.y);
      z[base] = y0;
      if (j == 0) {
        z[base + Q] = y1; z[base + 2 * Q] = y2; z[base + 3 * Q] = y3;
      } else {
        const unsigned e = j * step;
        z[base + Q] = cmul(y1, tw_n(T, e, N >> 1));
        z[base + 2 * Q] = cmul(y2, tw_n(T, 2 * e, N >> 1));
        z[base + 3 * Q] = cmul(y3, tw_n(T, 3 * e, N >> 1));
      }
    }
    __syncthreads();
    L = Q;
  }
  if (L == 2) {
    for (unsigned t = tid; t < (M >> 1); t += nt) {
      const float2 a = z[2 * t], b = z[2 * t + 1];
      z[2 * t] = make_float2(a.x + b.x, a.y + b.y);
      z[2 * t + 1] = make_float2(a.x - b.x, a.y - b.y);
    }
    __syncthreads();
  }
}

__device__ __forceinline__ float mag_to_dbfs(float re, float im, float nf, int* bad) {
  // spectrum-analyzer complex_to_magnitude (no fused multiply-add) then analyzer.rs:11-27
  const float m = __fsqrt_rn(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)));
  if (m == 0.0f) return -150.0f;
  const float scaled = __fdiv_rn(__fmul_rn(m, 4.0f), nf);
  const float db = 20.0f * log10f(scaled);
  if (!isfinite(db)) *bad = 1;
  return db;
}

// Same formula for the batched kernel: the division by N is an exact power-of-two scaling, and log10 is taken
// as log2 * log10(2) on the SFU (MUFU.LG2, abs error < 2^-22 in log2 units, i.e. < 2e-6 dB: two orders
// below the 1e-4 dB / 1e-5 relative tolerance of the path); sqrt.approx (2 ulp) sits below that too.
__device__ __forceinline__ float mag_to_dbfs_fast(float re, float im, float four_over_n, unsigned* mx) {
  float m;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(m) : "f"(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im))));
  const float db = m == 0.0f ? -150.0f : 6.02059991327962390f * __log2f(m * four_over_n);
  *mx = max(*mx, __float_as_uint(db) & 0x7fffffffu);  // >= 0x7f800000 at the end: a non-finite dB (ScalingError)
  return db;
}

// LAYOUT 0: mono, in[w][N].  LAYOUT 1: stereo in[w][N][2] -> mid & side planes.
// LAYOUT 2 / 3: stereo input, only the mid / only the side plane through the packed-real path
// (used when N complex points do not fit in shared memory).
// YOUT: the kernel writes the reference's y itself — (f64) dB + tilt[bin], the addition analyzer.rs:80-94 does in f64 —
// into y_out[w][plane][bin] instead of the f32 dB values (ssb_fft_batch_device_y).
template <int LAYOUT, bool YOUT = false>
__global__ void __launch_bounds__(512)
k_fft(const float* __restrict__ in, unsigned N, const float* __restrict__ window,
      const float2* __restrict__ T, unsigned k_first, unsigned n_bins, float* __restrict__ db_out,
      unsigned planes_out, unsigned plane_off, int32_t* __restrict__ status,
      const double* __restrict__ tilt = nullptr, double* __restrict__ y_out = nullptr) {
  extern __shared__ float2 z[];
  __shared__ int s_flags[3];  // nan, inf, scaling
  const unsigned w = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const unsigned M = (LAYOUT == 1) ? N : (N >> 1);
  if (tid < 3) s_flags[tid] = 0;
  __syncthreads();
  int f_nan = 0, f_inf = 0;
  const float nf = (float)N;
  if (LAYOUT == 1) {
    const float2* src = reinterpret_cast<const float2*>(in) + (size_t)w * N;
    for (unsigned n = tid; n < N; n += nt) {
      const float2 lr = __ldg(&src[n]);
      const float wn = __ldg(&window[n]);
      const float mid = __fmul_rn(__fadd_rn(lr.x, lr.y), 0.5f);
      const float side = __fmul_rn(__fsub_rn(lr.x, lr.y), 0.5f);
      const float2 v = make_float2(__fmul_rn(wn, mid), __fmul_rn(wn, side));
      f_nan |= (isnan(v.x) ? 1 : 0) | (isnan(v.y) ? 2 : 0);
      f_inf |= (isinf(v.x) ? 1 : 0) | (isinf(v.y) ? 2 : 0);
      z[n] = v;
    }
  } else {
    for (unsigned m = tid; m < M; m += nt) {
      float x0, x1;
      if (LAYOUT == 0) {
        const float2 p = __ldg(reinterpret_cast<const float2*>(in + (size_t)w * N) + m);
        x0 = p.x; x1 = p.y;
      } else {
        const float4 p = __ldg(reinterpret_cast<const float4*>(in + (size_t)w * N * 2) + m);
        if (LAYOUT == 2) { x0 = __fmul_rn(__fadd_rn(p.x, p.y), 0.5f); x1 = __fmul_rn(__fadd_rn(p.z, p.w), 0.5f); }
        else { x0 = __fmul_rn(__fsub_rn(p.x, p.y), 0.5f); x1 = __fmul_rn(__fsub_rn(p.z, p.w), 0.5f); }
      }
      const float2 wn = __ldg(reinterpret_cast<const float2*>(window) + m);
      const float2 v = make_float2(__fmul_rn(wn.x, x0), __fmul_rn(wn.y, x1));
      f_nan |= (isnan(v.x) || isnan(v.y)) ? 1 : 0;
      f_inf |= (isinf(v.x) || isinf(v.y)) ? 1 : 0;
      z[m] = v;
    }
  }
  if (f_nan) atomicOr(&s_flags[0], f_nan);
  if (f_inf) atomicOr(&s_flags[1], f_inf);
  __syncthreads();
  // twiddles of every stage are expressed as powers of W_N (W_L^e = W_N^(e*N/L)), so the one half table
  // serves both the N-point (mid/side) and the packed N/2-point (mono) transforms
  if (M > 1) fft_inplace(z, M, N, T);
  int bad = 0;
  if (LAYOUT == 1) {
    float* o_mid = db_out + ((size_t)w * planes_out + 0) * n_bins;
    float* o_side = db_out + ((size_t)w * planes_out + 1) * n_bins;
    for (unsigned i = tid; i < n_bins; i += nt) {
      const unsigned k = k_first + i;
      const float2 a = z[dif_position(k, M)];
      const float2 b = z[dif_position((N - k) & (N - 1), M)];
      const float mr = 0.5f * (a.x + b.x), mi = 0.5f * (a.y - b.y);
      const float sr = 0.5f * (a.y + b.y), si = -0.5f * (a.x - b.x);
      const float dm = mag_to_dbfs(mr, mi, nf, &bad), ds = mag_to_dbfs(sr, si, nf, &bad);
      if (YOUT) {
        const double tl = __ldg(&tilt[i]);
        y_out[((size_t)w * planes_out + 0) * n_bins + i] = (double)dm + tl;
        y_out[((size_t)w * planes_out + 1) * n_bins + i] = (double)ds + tl;
      } else {
        o_mid[i] = dm;
        o_side[i] = ds;
      }
    }
  } else {
    float* o = db_out + ((size_t)w * planes_out + plane_off) * n_bins;
    for (unsigned i = tid; i < n_bins; i += nt) {
      const unsigned k = k_first + i;
      float xr, xi;
      if (k == M) {
        const float2 z0 = z[0];
        xr = z0.x - z0.y; xi = 0.0f;
      } else {
        const float2 a = z[dif_position(k, M)];
        const float2 b = z[dif_position(M - k, M)];
        const float sr = 0.5f * (a.x + b.x), si = 0.5f * (a.y - b.y);
        const float dr = 0.5f * (a.x - b.x), di = 0.5f * (a.y + b.y);
        const float2 wk = __ldg(&T[k]);  // W_N^k, k < N/2
        xr = sr + (wk.x * di + wk.y * dr);
        xi = si + (wk.y * di - wk.x * dr);
      }
      const float dv = mag_to_dbfs(xr, xi, nf, &bad);
      if (YOUT) y_out[((size_t)w * planes_out + plane_off) * n_bins + i] = (double)dv + __ldg(&tilt[i]);
      else o[i] = dv;
    }
  }
  if (bad) atomicOr(&s_flags[2], 1);
  __syncthreads();
  if (status && tid == 0) {
    if (LAYOUT == 1) {
      for (int pl = 0; pl < 2; pl++) {
        int32_t st = SSB_OK;
        if (s_flags[0] & (1 << pl)) st = SSB_ERR_FFT_NAN;
        else if (s_flags[1] & (1 << pl)) st = SSB_ERR_FFT_INF;
        else if (s_flags[2]) st = SSB_ERR_FFT_SCALING;
        status[(size_t)w * planes_out + pl] = st;
      }
    } else {
      int32_t st = SSB_OK;
      if (s_flags[0]) st = SSB_ERR_FFT_NAN;
      else if (s_flags[1]) st = SSB_ERR_FFT_INF;
      else if (s_flags[2]) st = SSB_ERR_FFT_SCALING;
      status[(size_t)w * planes_out + plane_off] = st;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// k_fft_fast: the same window -> dB pipeline on the three-stage register-butterfly core (fft_core.cuh),
// for transform lengths M = 512 .. 16384.  One CTA (128 threads) per window, up to three CTAs per SM so
// one window's HBM load overlaps another's butterflies.  Twiddle tables are staged into shared memory by
// a TMA bulk copy (cp.async.bulk + mbarrier) issued before the input load, so they land underneath it.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int R1>
__device__ __forceinline__ void run_stage1(float2* z, unsigned M, unsigned N, const FftTwiddle& tw) {
  for (unsigned j = threadIdx.x; j < M / R1; j += blockDim.x) fft_stage1<R1>(z, M, N, tw, j);
}

// NT threads per CTA, MINB CTAs per SM: (192, 3) up to M = 8192 (69 KB of shared memory per window),
// (512, 1) for M = 16384 (135 KB: one window per SM, so the CTA itself has to bring the warps)
template <int LAYOUT, int NT, int MINB, bool YOUT = false>
__global__ void __launch_bounds__(NT, MINB)
k_fft_fast(const float* __restrict__ in, unsigned N, const float* __restrict__ window,
           const float2* __restrict__ g_lo, const float2* __restrict__ g_hi, unsigned k_first, unsigned n_bins,
           float* __restrict__ db_out, unsigned planes_out, unsigned plane_off, int32_t* __restrict__ status,
           const double* __restrict__ tilt = nullptr, double* __restrict__ y_out = nullptr) {
  extern __shared__ __align__(16) unsigned char fsm[];
  const unsigned M = (LAYOUT == 1) ? N : (N >> 1);
  const unsigned n_hi = N >> 6;
  float2* s_lo = reinterpret_cast<float2*>(fsm);                       // [64]
  float2* s_hi = s_lo + 64;                                            // [N/64]
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_hi + n_hi);            // 8-byte aligned: (64 + n_hi) * 8
  int* s_flags = reinterpret_cast<int*>(bar + 1);                      // [4]
  float2* z = reinterpret_cast<float2*>(s_flags + 4);                  // [M + M/32 + M/512]
  // LAYOUT 4: stereo input, one CTA per (window, plane): even CTAs transform mid, odd CTAs side, each through
  // the packed N/2-point path; the pair reads the same 2*N floats back to back, so the second read is an L2 hit
  const unsigned w = (LAYOUT == 4) ? (blockIdx.x >> 1) : blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  if (LAYOUT == 4) plane_off = blockIdx.x & 1u;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sm_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_flags[0] = s_flags[1] = s_flags[2] = 0;
    const unsigned bytes = (64 + n_hi) * (unsigned)sizeof(float2);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sm_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sm_u32(s_lo)), "l"(g_lo), "r"(64u * (unsigned)sizeof(float2)), "r"(sm_u32(bar)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sm_u32(s_hi)), "l"(g_hi), "r"(n_hi * (unsigned)sizeof(float2)), "r"(sm_u32(bar)) : "memory");
  }
  __syncthreads();

  // largest |bit pattern| seen per plane: > 0x7f800000 means a NaN is present, == means an infinity (and no
  // NaN) — exactly the precedence of spectrum-analyzer's checks (NaN first, then infinity)
  unsigned mx0 = 0, mx1 = 0;
  const float four_over_n = 4.0f / (float)N;  // exact: N is a power of two
  // loads are issued in batches of U per thread before any is consumed: the window streams in with
  // U * 128 independent 8/16-byte requests in flight per CTA instead of one per thread
  if (LAYOUT == 1) {
    constexpr int U = 8;
    const float2* src = reinterpret_cast<const float2*>(in) + (size_t)w * N;
    for (unsigned base = 0; base < N; base += U * nt) {
      float2 lr[U];
      float wn[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const unsigned n = base + u * nt + tid;
        lr[u] = n < N ? __ldg(&src[n]) : make_float2(0.f, 0.f);
        wn[u] = n < N ? __ldg(&window[n]) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const float mid = __fmul_rn(__fadd_rn(lr[u].x, lr[u].y), 0.5f);
        const float side = __fmul_rn(__fsub_rn(lr[u].x, lr[u].y), 0.5f);
        const float2 v = make_float2(__fmul_rn(wn[u], mid), __fmul_rn(wn[u], side));
        mx0 = max(mx0, __float_as_uint(v.x) & 0x7fffffffu);
        mx1 = max(mx1, __float_as_uint(v.y) & 0x7fffffffu);
        if (base + u * nt + tid < N) z[fft_pad(base + u * nt + tid)] = v;
      }
    }
  } else {
    constexpr int U = 4;
    for (unsigned base = 0; base < M; base += U * nt) {
      float x0[U], x1[U];
      float2 wn[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const unsigned m = min(base + u * nt + tid, M - 1);  // clamped; out-of-range lanes are not stored
        if (LAYOUT == 0) {
          const float2 p = __ldg(reinterpret_cast<const float2*>(in + (size_t)w * N) + m);
          x0[u] = p.x; x1[u] = p.y;
        } else {
          const float4 p = __ldg(reinterpret_cast<const float4*>(in + (size_t)w * N * 2) + m);
          if (LAYOUT == 2 || (LAYOUT == 4 && plane_off == 0)) { x0[u] = __fmul_rn(__fadd_rn(p.x, p.y), 0.5f); x1[u] = __fmul_rn(__fadd_rn(p.z, p.w), 0.5f); }
          else { x0[u] = __fmul_rn(__fsub_rn(p.x, p.y), 0.5f); x1[u] = __fmul_rn(__fsub_rn(p.z, p.w), 0.5f); }
        }
        wn[u] = __ldg(reinterpret_cast<const float2*>(window) + m);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const float2 v = make_float2(__fmul_rn(wn[u].x, x0[u]), __fmul_rn(wn[u].y, x1[u]));
        mx0 = max(mx0, max(__float_as_uint(v.x) & 0x7fffffffu, __float_as_uint(v.y) & 0x7fffffffu));
        if (base + u * nt + tid < M) z[fft_pad(base + u * nt + tid)] = v;
      }
    }
  }
  if (mx0 >= 0x7f800000u) atomicOr(mx0 > 0x7f800000u ? &s_flags[0] : &s_flags[1], 1);
  if (mx1 >= 0x7f800000u) atomicOr(mx1 > 0x7f800000u ? &s_flags[0] : &s_flags[1], 2);
  // twiddles have landed?
  {
    unsigned ok = 0, spins = 0;
    while (!ok) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(ok) : "r"(sm_u32(bar)) : "memory");
      if (!ok && ++spins > (1u << 26)) __trap();
    }
  }
  __syncthreads();

  const FftTwiddle tw{s_lo, s_hi};
  switch (M >> 9) {
    case 2: run_stage1<2>(z, M, N, tw); break;
    case 4: run_stage1<4>(z, M, N, tw); break;
    case 8: run_stage1<8>(z, M, N, tw); break;
    case 16: run_stage1<16>(z, M, N, tw); break;
    case 32: run_stage1<32>(z, M, N, tw); break;
    default: break;  // M == 512: no first stage
  }
  __syncthreads();
  for (unsigned t = tid; t < (M >> 4); t += nt) fft_stage2(z, N, tw, t);
  __syncthreads();
  for (unsigned t = tid; t < (M >> 5); t += nt) fft_stage3(z, t);
  __syncthreads();

  unsigned mxo = 0;
  const unsigned r1s = 31u - (unsigned)__clz(M >> 9);  // log2(R1)
  if (LAYOUT == 1) {
    float* o_mid = db_out + ((size_t)w * planes_out + 0) * n_bins;
    float* o_side = db_out + ((size_t)w * planes_out + 1) * n_bins;
    for (unsigned i = tid; i < n_bins; i += nt) {
      const unsigned k = k_first + i;
      const float2 a = z[fft_position(k, r1s)];
      const float2 b = z[fft_position((N - k) & (N - 1), r1s)];
      const float mr = 0.5f * (a.x + b.x), mi = 0.5f * (a.y - b.y);
      const float sr = 0.5f * (a.y + b.y), si = -0.5f * (a.x - b.x);
      const float dm = mag_to_dbfs_fast(mr, mi, four_over_n, &mxo), ds = mag_to_dbfs_fast(sr, si, four_over_n, &mxo);
      if (YOUT) {
        const double tl = __ldg(&tilt[i]);
        y_out[((size_t)w * planes_out + 0) * n_bins + i] = (double)dm + tl;
        y_out[((size_t)w * planes_out + 1) * n_bins + i] = (double)ds + tl;
      } else {
        o_mid[i] = dm;
        o_side[i] = ds;
      }
    }
  } else {
    float* o = db_out + ((size_t)w * planes_out + plane_off) * n_bins;
    for (unsigned i = tid; i < n_bins; i += nt) {
      const unsigned k = k_first + i;
      float xr, xi;
      if (k == M) {
        const float2 z0 = z[fft_position(0, r1s)];
        xr = z0.x - z0.y; xi = 0.0f;
      } else {
        const float2 a = z[fft_position(k, r1s)];
        const float2 b = z[fft_position(M - k, r1s)];
        const float sr = 0.5f * (a.x + b.x), si = 0.5f * (a.y - b.y);
        const float dr = 0.5f * (a.x - b.x), di = 0.5f * (a.y + b.y);
        const float2 wk = fft_twiddle(tw, k);  // W_N^k
        xr = sr + (wk.x * di + wk.y * dr);
        xi = si + (wk.y * di - wk.x * dr);
      }
      const float dv = mag_to_dbfs_fast(xr, xi, four_over_n, &mxo);
      if (YOUT) y_out[((size_t)w * planes_out + plane_off) * n_bins + i] = (double)dv + __ldg(&tilt[i]);
      else o[i] = dv;
    }
  }
  if (mxo >= 0x7f800000u) atomicOr(&s_flags[2], 1);
  __syncthreads();
  if (status && tid == 0) {
    if (LAYOUT == 1) {
      for (int pl = 0; pl < 2; pl++) {
        int32_t st = SSB_OK;
        if (s_flags[0] & (1 << pl)) st = SSB_ERR_FFT_NAN;
        else if (s_flags[1] & (1 << pl)) st = SSB_ERR_FFT_INF;
        else if (s_flags[2]) st = SSB_ERR_FFT_SCALING;
        status[(size_t)w * planes_out + pl] = st;
      }
    } else {
      int32_t st = SSB_OK;
      if (s_flags[0]) st = SSB_ERR_FFT_NAN;
      else if (s_flags[1]) st = SSB_ERR_FFT_INF;
      else if (s_flags[2]) st = SSB_ERR_FFT_SCALING;
      status[(size_t)w * planes_out + plane_off] = st;
    }
  }
}

template <int LAYOUT, bool YOUT = false>
static cudaError_t launch_fft_layout(const FftPlan& plan, const float* d_in, size_t n_windows, float* d_db,
                                     unsigned planes_out, unsigned plane_off, int32_t* d_status, cudaStream_t s,
                                     double* d_y = nullptr) {
  const unsigned N = (unsigned)plan.n;
  const unsigned M = (LAYOUT == 1) ? N : (N >> 1);
  if (M >= 512 && plan.d_tw_lo) {
    const size_t fsmem = (size_t)(64 + (N >> 6)) * sizeof(float2) + 8 + 16 + (size_t)(M + (M >> 5) + (M >> 9) + 1) * sizeof(float2);
    if (M > 8192) {
      cudaError_t fe = cudaFuncSetAttribute(k_fft_fast<LAYOUT, 512, 1, YOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
      if (fe) return fe;
      k_fft_fast<LAYOUT, 512, 1, YOUT><<<(unsigned)(LAYOUT == 4 ? 2 * n_windows : n_windows), 512, fsmem, s>>>(d_in, N, plan.d_window, plan.d_tw_lo, plan.d_tw_hi,
                                                                        (unsigned)plan.k_first, (unsigned)plan.n_bins, d_db,
                                                                        planes_out, plane_off, d_status, plan.d_tilt, d_y);
    } else {
      // CTA shape for M <= 8192: 256 x 3 (measured 32.5 % of the HBM peak at N = 8192; SSB_FFT_CFG=0 selects
      // 192 x 3: 29.8 %, SSB_FFT_CFG=2 selects 256 x 2: 28.7 %) — a tuning knob, all exact
      static const int cfg = [] { const char* e = getenv("SSB_FFT_CFG"); return e ? atoi(e) : 1; }();   // read once (thread-safe)
#define SSB_LAUNCH_FFT(NT, MB)                                                                                        \
  do {                                                                                                                 \
    cudaError_t fe = cudaFuncSetAttribute(k_fft_fast<LAYOUT, NT, MB, YOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)fsmem);                                                                 \
    if (fe) return fe;                                                                                                 \
    k_fft_fast<LAYOUT, NT, MB, YOUT><<<(unsigned)(LAYOUT == 4 ? 2 * n_windows : n_windows), NT, fsmem, s>>>(d_in, N, plan.d_window, plan.d_tw_lo,      \
                                                                     plan.d_tw_hi, (unsigned)plan.k_first,             \
                                                                     (unsigned)plan.n_bins, d_db, planes_out,          \
                                                                     plane_off, d_status, plan.d_tilt, d_y);           \
  } while (0)
      if (cfg == 0) SSB_LAUNCH_FFT(192, 3);
      else if (cfg == 2) SSB_LAUNCH_FFT(256, 2);
      else SSB_LAUNCH_FFT(256, 3);
#undef SSB_LAUNCH_FFT
    }
    return cudaGetLastError();
  }
  const size_t smem = (size_t)(M ? M : 1) * sizeof(float2);
  cudaError_t e = cudaFuncSetAttribute(k_fft<LAYOUT, YOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e) return e;
  unsigned threads = M / 4 < 64 ? 64 : (M / 4 > 512 ? 512 : M / 4);
  k_fft<LAYOUT, YOUT><<<(unsigned)n_windows, threads, smem, s>>>(d_in, N, plan.d_window, plan.d_twiddle,
                                                                (unsigned)plan.k_first, (unsigned)plan.n_bins, d_db,
                                                                planes_out, plane_off, d_status, plan.d_tilt, d_y);
  return cudaGetLastError();
}

cudaError_t launch_fft(const FftPlan& plan, const float* d_in, int layout, size_t n_windows, float* d_db_out,
                       int32_t* d_status, cudaStream_t s, uint64_t* launches) {
  if (!n_windows) return cudaSuccess;
  cudaError_t e;
  if (layout == SSB_FFT_MONO) {
    e = launch_fft_layout<0>(plan, d_in, n_windows, d_db_out, 1, 0, d_status, s);
    if (launches) ++*launches;
    return e;
  }
  if (plan.n <= 8192) {
    e = launch_fft_layout<1>(plan, d_in, n_windows, d_db_out, 2, 0, d_status, s);
    if (launches) ++*launches;
    return e;
  }
  // N = 16384 / 32768: one N-point complex transform needs 135+ KB of shared memory (one CTA per SM); two packed
  // N/2-point transforms per window (mid, side) in neighbouring CTAs keep three CTAs per SM at N = 16384
  e = launch_fft_layout<4>(plan, d_in, n_windows, d_db_out, 2, 0, d_status, s);
  if (launches) ++*launches;
  return e;
}

// The same transforms with the reference's y = (f64) dB + tilt as the output (d_y_out[w][plane][bin], f64).
cudaError_t launch_fft_y(const FftPlan& plan, const float* d_in, int layout, size_t n_windows, double* d_y_out,
                         int32_t* d_status, cudaStream_t s, uint64_t* launches) {
  if (!n_windows) return cudaSuccess;
  if (!plan.d_tilt) return cudaErrorInvalidValue;
  cudaError_t e;
  if (layout == SSB_FFT_MONO) e = launch_fft_layout<0, true>(plan, d_in, n_windows, nullptr, 1, 0, d_status, s, d_y_out);
  else if (plan.n <= 8192) e = launch_fft_layout<1, true>(plan, d_in, n_windows, nullptr, 2, 0, d_status, s, d_y_out);
  else e = launch_fft_layout<4, true>(plan, d_in, n_windows, nullptr, 2, 0, d_status, s, d_y_out);
  if (launches) ++*launches;
  return e;
}

// ------------------------------------------------------------------------------------------------
// get_waveform: min/max decimation.  Column bounds are the reference's f64 expressions, evaluated
// with IEEE mul/div/ceil on the device (bit-exact with analyzer.rs:118-120).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_waveform(const float* __restrict__ x, unsigned long long len, double spp, unsigned long long columns,
           float* __restrict__ out) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned long long warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned long long n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  for (unsigned long long i = warp; i < columns; i += n_warps) {
    const unsigned long long start = (unsigned long long)__dmul_rn((double)i, spp);
    unsigned long long end = (unsigned long long)ceil(__dmul_rn((double)(i + 1), spp));
    if (end > len) end = len;
    float mn, mx;
    if (end > start) {
      mn = mx = x[start];  // Rust reduce(): first element seeds, NaN-ignoring f32::min / f32::max
      for (unsigned long long j = start + lane; j < end; j += 32) {
        const float v = x[j];
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      }
    } else {
      mn = mx = 0.0f;
    }
    if (lane == 0) { out[2 * i] = mn; out[2 * i + 1] = mx; }
  }
}

// Short columns (the usual case: 48-100 samples per column): one thread per column.  A lane walks its own
// column front to back — scalar until 16-byte aligned, then float4, then the scalar tail — and neighbouring
// lanes own neighbouring columns, so the warp as a whole touches a contiguous span whose 128-byte lines stay
// in L1 until every lane has used them.  (Measured alternatives: one warp per ~100-sample column 30 % of the
// HBM peak, this kernel with scalar loads 38 %, a shared-memory staged version 14 %.)
__global__ void __launch_bounds__(256)
k_waveform_thread(const float* __restrict__ x, unsigned long long len, double spp, unsigned long long columns,
                  float* __restrict__ out) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < columns;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long start = (unsigned long long)__dmul_rn((double)i, spp);
    unsigned long long end = (unsigned long long)ceil(__dmul_rn((double)(i + 1), spp));
    if (end > len) end = len;
    float mn = 0.0f, mx = 0.0f;
    if (end > start) {
      mn = mx = x[start];
      unsigned long long j = start + 1;
      const bool base_aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
      for (; j < end && (!base_aligned || (j & 3)); j++) {
        const float v = x[j];
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
      }
      for (; j + 4 <= end; j += 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(x + j));
        mn = fminf(fminf(mn, q.x), fminf(q.y, fminf(q.z, q.w)));
        mx = fmaxf(fmaxf(mx, q.x), fmaxf(q.y, fmaxf(q.z, q.w)));
      }
      for (; j < end; j++) {
        const float v = x[j];
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
      }
    }
    out[2 * i] = mn;
    out[2 * i + 1] = mx;
  }
}

cudaError_t launch_waveform(const float* d_samples, size_t len, size_t window, float* d_minmax, size_t columns,
                            cudaStream_t s, uint64_t* launches) {
  if (!columns) return cudaSuccess;
  const double spp = (double)len / (double)window;
  const unsigned tpb = 256;
  if (spp <= 1024.0) {
    size_t blocks = (columns + tpb - 1) / tpb;
    if (blocks > 148 * 32) blocks = 148 * 32;
    k_waveform_thread<<<(unsigned)blocks, tpb, 0, s>>>(d_samples, (unsigned long long)len, spp,
                                                       (unsigned long long)columns, d_minmax);
  } else {
    size_t blocks = (columns * 32 + tpb - 1) / tpb;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_waveform<<<(unsigned)blocks, tpb, 0, s>>>(d_samples, (unsigned long long)len, spp, (unsigned long long)columns,
                                                d_minmax);
  }
  if (launches) ++*launches;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// get_mid_and_side_samples
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_mid_side(const float2* __restrict__ in, size_t frames, float* __restrict__ mid, float* __restrict__ side) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < frames; i += (size_t)gridDim.x * blockDim.x) {
    const float2 lr = __ldg(&in[i]);
    mid[i] = __fmul_rn(__fadd_rn(lr.x, lr.y), 0.5f);
    side[i] = __fmul_rn(__fsub_rn(lr.x, lr.y), 0.5f);
  }
}

cudaError_t launch_mid_side(const float* d_in, size_t frames, float* d_mid, float* d_side, cudaStream_t s,
                            uint64_t* launches) {
  if (!frames) return cudaSuccess;
  size_t blocks = (frames + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_mid_side<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const float2*>(d_in), frames, d_mid, d_side);
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace ssb
