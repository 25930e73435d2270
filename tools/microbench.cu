// microbench.cu — measures the few SM constants the loudness kernel design depends on (B200, sm_100a):
// DFMA throughput / dependent latency, F2F.F64.F32 throughput, and an integer f32->f64 widening.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double* out, double a, double b, int iters) {
  double v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) v[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = fma(v[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// three distinct register operands per DFMA (coefficients made thread-dependent so they live in R registers)
template <int ILP>
__global__ void k_dfma3(double* out, double a, double b, int iters) {
  double v[ILP];
  double ca[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { v[i] = threadIdx.x * 1e-9 + i; ca[i] = a + threadIdx.x * 1e-12 * (i + 1); }
  double acc = b + threadIdx.x * 1e-13;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = fma(v[i], ca[i], acc);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA (mma.sync m8n8k4 f64) alone / mixed with DFMA: is the FP64 tensor path a separate pipe on sm_100a?
template <int N_MMA, int N_FMA>
__global__ void k_dmma_mix(double* out, double a, double b, int iters) {
  double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
  double v[8];
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = threadIdx.x * 1e-9 + i;
  const double ax = a + threadIdx.x * 1e-12, bx = b + threadIdx.x * 1e-12;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int m = 0; m < N_MMA; m++) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[m & 3]), "+d"(c1[m & 3]) : "d"(ax), "d"(bx));
    }
#pragma unroll
    for (int f = 0; f < N_FMA; f++) v[f & 7] = fma(v[f & 7], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += v[i];
#pragma unroll
  for (int i = 0; i < 4; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_cvt_hw(double* out, const float* in, int iters) {
  float x = in[threadIdx.x & 31];
  double acc0 = 0, acc1 = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      x = __int_as_float(__float_as_int(x) ^ (it + i));  // keep the conversion from being hoisted
      double d = (double)x;
      // fold with integer ops only so the FP64 pipe sees just the conversions
      acc0 = __longlong_as_double(__double_as_longlong(acc0) ^ __double_as_longlong(d));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1;
}

__device__ __forceinline__ double widen_int(float x) {
  const unsigned u = __float_as_uint(x);
  const unsigned hi = (u & 0x80000000u) | (((u & 0x7fffffffu) >> 3) + 0x38000000u);
  const unsigned lo = u << 29;
  return __hiloint2double((int)hi, (int)lo);
}

__global__ void k_cvt_int(double* out, const float* in, int iters) {
  float x = in[threadIdx.x & 31];
  double acc0 = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      x = __int_as_float(__float_as_int(x) ^ (it + i));
      double d = widen_int(x);
      acc0 = __longlong_as_double(__double_as_longlong(acc0) ^ __double_as_longlong(d));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("device %s sms %d clock %d kHz\n", p.name, p.multiProcessorCount, clk_khz);
  const int sms = p.multiProcessorCount;
  double* out; float* in;
  cudaMalloc(&out, sizeof(double) * sms * 32 * 1024);
  cudaMalloc(&in, 128);
  cudaMemset(in, 0x3f, 128);
  const int iters = 1 << 14;
  // throughput: many warps, ILP 8
  for (int warps : {1, 2, 4, 8, 16, 32}) {
    float ms = time_ms([&] { k_dfma<8><<<sms, warps * 32>>>(out, 1.0000001, 1e-9, iters); });
    double fma = (double)sms * warps * 32 * 8 * iters;
    printf("dfma ilp8 warps/SM %2d: %.3f ms  %.2f TFLOP/s  %.1f DFMA/clk/SM @%d MHz\n", warps, ms, 2 * fma / ms / 1e9,
           fma / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1000);
  }
  for (int warps : {4, 8, 16}) {
    float ms = time_ms([&] { k_dfma3<8><<<sms, warps * 32>>>(out, 1.0000001, 1e-9, iters); });
    double fma = (double)sms * warps * 32 * 8 * iters;
    printf("dfma 3-register-operand ilp8 warps/SM %2d: %.1f DFMA/clk/SM\n", warps, fma / (ms * 1e-3) / sms / (clk_khz * 1e3));
  }
  {
    const int it2 = 1 << 12;
    float t_mma = time_ms([&] { k_dmma_mix<8, 0><<<sms, 512>>>(out, 1.0000001, 1e-9, it2); });
    float t_fma = time_ms([&] { k_dmma_mix<0, 32><<<sms, 512>>>(out, 1.0000001, 1e-9, it2); });
    float t_mix = time_ms([&] { k_dmma_mix<8, 32><<<sms, 512>>>(out, 1.0000001, 1e-9, it2); });
    double mma_flops = 2.0 * 256 * 8 * it2 * 16.0 * sms;  // m8n8k4 = 256 FMA per warp instruction
    printf("dmma m8n8k4: %.3f ms (%.2f TFLOP/s) | dfma x32: %.3f ms | both: %.3f ms  (sum %.3f, max %.3f)\n", t_mma,
           mma_flops / t_mma / 1e9, t_fma, t_mix, t_mma + t_fma, t_mma > t_fma ? t_mma : t_fma);
  }
  // latency: one warp per SM, ILP 1
  {
    float ms = time_ms([&] { k_dfma<1><<<sms, 32>>>(out, 1.0000001, 1e-9, iters * 8); });
    printf("dfma dependent latency: %.2f cycles @%d MHz nominal\n", ms * 1e-3 * clk_khz * 1e3 / (iters * 8.0), clk_khz / 1000);
  }
  for (int ilp_warps : {1, 4}) {
    float ms = time_ms([&] { k_dfma<4><<<sms, 32 * ilp_warps>>>(out, 1.0000001, 1e-9, iters * 2); });
    double fma = (double)sms * ilp_warps * 32 * 4 * iters * 2;
    printf("dfma ilp4 warps/SM %d: %.1f DFMA/clk/SM\n", ilp_warps, fma / (ms * 1e-3) / sms / (clk_khz * 1e3));
  }
  for (int warps : {4, 16, 32}) {
    float ms = time_ms([&] { k_cvt_hw<<<sms, warps * 32>>>(out, in, iters); });
    double n = (double)sms * warps * 32 * 8 * iters;
    printf("F2F.F64.F32 warps/SM %2d: %.1f cvt/clk/SM\n", warps, n / (ms * 1e-3) / sms / (clk_khz * 1e3));
    ms = time_ms([&] { k_cvt_int<<<sms, warps * 32>>>(out, in, iters); });
    printf("int widen     warps/SM %2d: %.1f cvt/clk/SM\n", warps, n / (ms * 1e-3) / sms / (clk_khz * 1e3));
  }
  return 0;
}
