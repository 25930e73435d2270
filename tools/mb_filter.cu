// mb_filter.cu — how close can W warps per SM sub-partition get to the FP64 pipe with the K-weighting loop?
// Each variant runs the real per-sample instruction mix (LDS.128 + select + F2F + DFMA chains) on data in
// shared memory, with no global traffic, and reports SM cycles per sample-step and the implied FP64 pipe
// utilisation (a warp DFMA occupies the 16-lane pipe of its sub-partition for 2 cycles).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mb_filter tools/mb_filter.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

struct Coef {
  double na[5], b[5], cy[5];
};

__device__ __forceinline__ float pick2(const float4& q, int f, int c) {
  return f == 0 ? (c ? q.y : q.x) : (c ? q.w : q.z);
}

// MODE 0: pass 2, reference form (4 + 5 + 1 = 10 DFMA, 1 F2F)
// MODE 1: pass 2, scaled output y' = x + sum c_i v_i (4 + 4 + 1 = 9 DFMA, 1 F2F)
// MODE 2: MODE 0 + fused zero-state pass of another tile (14 DFMA, 2 F2F)   [round-1 kernel]
// MODE 3: MODE 1 + fused zero-state pass (13 DFMA, 2 F2F)
// MODE 4: MODE 3, two independent (pass 2 + pass 1) pairs per lane (26 DFMA per step, 4 F2F): ILP 4
// MODE 5: MODE 3 with the samples already f64 in registers (no F2F at all: upper bound for "convert once")
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(const __grid_constant__ Coef a, double* out, int iters, long long* cyc) {
  extern __shared__ unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // fill 64 KB of shared memory with small values
  float* sf = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) sf[i] = 1e-3f * (float)((i * 2654435761u) >> 20) - 2.0f;
  __syncthreads();
  const int c = lane & 1;
  const unsigned char* base = smem + warp * 2048 % 16384 + (lane >> 1) * 128;
  double v1 = 0, v2 = 0, v3 = 0, v4 = 0, acc = 0;
  double z1 = 0, z2 = 0, z3 = 0, z4 = 0;
  double w1 = 0, w2 = 0, w3 = 0, w4 = 0, acc2 = 0;
  double u1 = 0, u2 = 0, u3 = 0, u4 = 0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    const unsigned char* line = base + (it & 7) * 16;
    const unsigned char* line2 = base + 16384 + (it & 7) * 16;
#pragma unroll
    for (int qi = 0; qi < 8; qi++) {
      const float4 q = *reinterpret_cast<const float4*>(line + ((qi ^ (lane >> 2 & 7)) << 4));
      float4 qn = q, q3 = q, q4 = q;
      if (MODE >= 2) qn = *reinterpret_cast<const float4*>(line2 + ((qi ^ (lane >> 2 & 7)) << 4));
      if (MODE == 4) {
        q3 = *reinterpret_cast<const float4*>(line + 8192 + ((qi ^ (lane >> 2 & 7)) << 4));
        q4 = *reinterpret_cast<const float4*>(line2 + 8192 + ((qi ^ (lane >> 2 & 7)) << 4));
      }
#pragma unroll
      for (int f = 0; f < 2; f++) {
        const float xf = pick2(q, f, c);
        const double x = MODE == 5 ? __hiloint2double(__float_as_int(xf) >> 3, 0) * 1e-300 : (double)xf;
        double t = fma(a.na[4], v4, x);
        t = fma(a.na[3], v3, t);
        t = fma(a.na[2], v2, t);
        const double v0 = fma(a.na[1], v1, t);
        double y;
        if (MODE == 0 || MODE == 2) {
          y = a.b[4] * v4;
          y = fma(a.b[3], v3, y);
          y = fma(a.b[2], v2, y);
          y = fma(a.b[1], v1, y);
          y = fma(a.b[0], v0, y);
        } else {
          y = fma(a.cy[4], v4, x);
          y = fma(a.cy[3], v3, y);
          y = fma(a.cy[2], v2, y);
          y = fma(a.cy[1], v1, y);
        }
        v4 = v3; v3 = v2; v2 = v1; v1 = v0;
        acc = fma(y, y, acc);
        if (MODE >= 2) {
          const float xnf = pick2(qn, f, c);
          const double xn = MODE == 5 ? __hiloint2double(__float_as_int(xnf) >> 3, 0) * 1e-300 : (double)xnf;
          double tn = fma(a.na[4], z4, xn);
          tn = fma(a.na[3], z3, tn);
          tn = fma(a.na[2], z2, tn);
          const double z0 = fma(a.na[1], z1, tn);
          z4 = z3; z3 = z2; z2 = z1; z1 = z0;
        }
        if (MODE == 4) {
          const double x3 = (double)pick2(q3, f, c);
          double t3 = fma(a.na[4], w4, x3);
          t3 = fma(a.na[3], w3, t3);
          t3 = fma(a.na[2], w2, t3);
          const double w0 = fma(a.na[1], w1, t3);
          double y3 = fma(a.cy[4], w4, x3);
          y3 = fma(a.cy[3], w3, y3);
          y3 = fma(a.cy[2], w2, y3);
          y3 = fma(a.cy[1], w1, y3);
          w4 = w3; w3 = w2; w2 = w1; w1 = w0;
          acc2 = fma(y3, y3, acc2);
          const double x4 = (double)pick2(q4, f, c);
          double t4 = fma(a.na[4], u4, x4);
          t4 = fma(a.na[3], u3, t4);
          t4 = fma(a.na[2], u2, t4);
          const double u0 = fma(a.na[1], u1, t4);
          u4 = u3; u3 = u2; u2 = u1; u1 = u0;
        }
      }
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + acc2 + v1 + z1 + w1 + u1 + v4 + z4 + w4 + u4;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(int warps_per_smsp, const Coef& c, double* d_out, long long* d_cyc, int sms) {
  const int threads = warps_per_smsp * 4 * 32;
  const int iters = 4000;
  auto kern = k<MODE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  kern<<<sms, threads, 65536>>>(c, d_out, 200, d_cyc);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<<<sms, threads, 65536>>>(c, d_out, iters, d_cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[256];
  cudaMemcpy(h, d_cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < sms; i++) avg += (double)h[i];
  avg /= sms;
  const double steps = (double)iters * 16;  // sample-steps per warp
  const int dfma[6] = {10, 9, 14, 13, 26, 13};
  const double cyc_per_step = avg / steps;
  const double util = dfma[MODE] * 2.0 * warps_per_smsp / cyc_per_step;
  printf("mode %d warps/smsp %d: %.2f cycles/step/warp, fp64 pipe %.1f %%, %.3f ms, clk %.0f MHz, %.2f cycles per DFMA-lane-sample\n",
         MODE, warps_per_smsp, cyc_per_step, 100.0 * util, ms, avg / (ms * 1e3),
         cyc_per_step / warps_per_smsp / (MODE == 4 ? 2 : 1));
  cudaError_t e = cudaGetLastError();
  if (e) printf("  error: %s\n", cudaGetErrorString(e));
}

int main() {
  Coef c;
  const double a[5] = {1.0, -3.68070674801639, 5.08704520879759, -3.13154635528588, 0.72520807726273};
  const double b[5] = {1.53512485958697, -5.76194590858032, 8.11691004925258, -5.08848181111208, 1.19839281085285};
  for (int i = 0; i < 5; i++) { c.na[i] = -a[i]; c.b[i] = b[i]; c.cy[i] = b[i] / b[0] - a[i]; }
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* d_out;
  long long* d_cyc;
  cudaMalloc(&d_out, 148 * 512 * sizeof(double) * 2);
  cudaMalloc(&d_cyc, 256 * sizeof(long long));
  for (int w = 1; w <= 4; w++) {
    run<0>(w, c, d_out, d_cyc, sms);
    run<1>(w, c, d_out, d_cyc, sms);
    run<2>(w, c, d_out, d_cyc, sms);
    run<3>(w, c, d_out, d_cyc, sms);
    run<4>(w, c, d_out, d_cyc, sms);
    run<5>(w, c, d_out, d_cyc, sms);
  }
  return 0;
}
