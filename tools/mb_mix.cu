// mb_mix.cu — FP64 pipe sharing between heterogeneous warps on one SM sub-partition: how fast does a "full filter" warp
// (9 DFMA / sample) step when it shares its sub-partition with another one and with 0 / 1 / 2 "zero-state" warps
// (4 DFMA / sample, one dependent chain)?  All warps loop until warp 0 has done its steps; each reports steps / cycles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mb_mix tools/mb_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

struct Coef { double na[5], cy[5]; };

__device__ __forceinline__ float pick2(const float4& q, int f, int c) { return f == 0 ? (c ? q.y : q.x) : (c ? q.w : q.z); }

// roles: 0 = P2 (recursion + scaled tap + square: 9 DFMA), 1 = P1 (recursion only: 4 DFMA), 2 = fused (13 DFMA)
__global__ void __launch_bounds__(512, 1) k(const __grid_constant__ Coef a, int n_p2, int n_p1, int n_fused, int steps16,
                                            double* out, long long* cyc, int* done_steps) {
  extern __shared__ unsigned char smem[];
  __shared__ volatile int stop;
  float* sf = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) sf[i] = 1e-3f * (float)((i * 2654435761u) >> 20) - 2.0f;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = warp >> 2;  // index of this warp inside its sub-partition
  const int role = per < n_p2 ? 0 : (per < n_p2 + n_p1 ? 1 : 2);
  const int c = lane & 1;
  const unsigned char* base = smem + (warp * 2048) % 16384 + (lane >> 1) * 128;
  double v1 = 0, v2 = 0, v3 = 0, v4 = 0, acc = 0, z1 = 0, z2 = 0, z3 = 0, z4 = 0;
  const long long t0 = clock64();
  int it = 0;
  for (; it < (1 << 28); it++) {
    if (warp == 0 ? it >= steps16 : stop) break;
    const unsigned char* line = base + (it & 7) * 16;
    const unsigned char* line2 = base + 16384 + (it & 7) * 16;
#pragma unroll
    for (int qi = 0; qi < 8; qi++) {
      const float4 q = *reinterpret_cast<const float4*>(line + ((qi ^ (lane >> 2 & 7)) << 4));
      float4 qn = q;
      if (role == 2) qn = *reinterpret_cast<const float4*>(line2 + ((qi ^ (lane >> 2 & 7)) << 4));
#pragma unroll
      for (int f = 0; f < 2; f++) {
        const double x = (double)pick2(q, f, c);
        if (role != 1) {
          double t = fma(a.na[4], v4, x);
          t = fma(a.na[3], v3, t);
          t = fma(a.na[2], v2, t);
          const double v0 = fma(a.na[1], v1, t);
          double y = fma(a.cy[4], v4, x);
          y = fma(a.cy[3], v3, y);
          y = fma(a.cy[2], v2, y);
          y = fma(a.cy[1], v1, y);
          v4 = v3; v3 = v2; v2 = v1; v1 = v0;
          acc = fma(y, y, acc);
        }
        if (role != 0) {
          const double xn = role == 2 ? (double)pick2(qn, f, c) : x;
          double tn = fma(a.na[4], z4, xn);
          tn = fma(a.na[3], z3, tn);
          tn = fma(a.na[2], z2, tn);
          const double z0 = fma(a.na[1], z1, tn);
          z4 = z3; z3 = z2; z2 = z1; z1 = z0;
        }
      }
    }
  }
  const long long t1 = clock64();
  if (warp == 0 && lane == 0) stop = 1;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + v1 + z1 + v4 + z4;
  if (lane == 0) {
    cyc[blockIdx.x * 16 + warp] = t1 - t0;
    done_steps[blockIdx.x * 16 + warp] = it * 16;
  }
}

int main() {
  Coef c;
  const double a[5] = {1.0, -3.68070674801639, 5.08704520879759, -3.13154635528588, 0.72520807726273};
  const double b[5] = {1.53512485958697, -5.76194590858032, 8.11691004925258, -5.08848181111208, 1.19839281085285};
  for (int i = 0; i < 5; i++) { c.na[i] = -a[i]; c.cy[i] = b[i] / b[0] - a[i]; }
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* d_out; long long* d_cyc; int* d_steps;
  cudaMalloc(&d_out, 148 * 512 * sizeof(double));
  cudaMalloc(&d_cyc, 148 * 16 * sizeof(long long));
  cudaMalloc(&d_steps, 148 * 16 * sizeof(int));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const int mixes[][3] = {{1, 0, 0}, {2, 0, 0}, {3, 0, 0}, {4, 0, 0}, {2, 1, 0}, {2, 2, 0}, {1, 1, 0}, {1, 2, 0}, {0, 2, 0}, {0, 4, 0},
                          {0, 0, 2}, {0, 0, 3}, {0, 0, 4}, {1, 0, 1}, {2, 0, 2}};
  for (auto& m : mixes) {
    const int warps = (m[0] + m[1] + m[2]) * 4;
    k<<<sms, warps * 32, 65536>>>(c, m[0], m[1], m[2], 100, d_out, d_cyc, d_steps);
    k<<<sms, warps * 32, 65536>>>(c, m[0], m[1], m[2], 3000, d_out, d_cyc, d_steps);
    cudaDeviceSynchronize();
    long long hc[16]; int hs[16];
    cudaMemcpy(hc, d_cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    cudaMemcpy(hs, d_steps, sizeof(hs), cudaMemcpyDeviceToHost);
    printf("P2 x%d  P1 x%d  fused x%d per sub-partition:", m[0], m[1], m[2]);
    double pipe = 0;
    for (int w = 0; w < warps; w += 4) {
      const int per = w >> 2;
      const int role = per < m[0] ? 0 : (per < m[0] + m[1] ? 1 : 2);
      const double cps = (double)hc[w] / hs[w];
      const int dfma = role == 0 ? 9 : (role == 1 ? 4 : 13);
      pipe += 2.0 * dfma / cps;
      printf("  %s %.1f cyc/step", role == 0 ? "P2" : (role == 1 ? "P1" : "F"), cps);
    }
    printf("  | fp64 pipe %.0f %%\n", 100 * pipe);
    cudaError_t e = cudaGetLastError();
    if (e) printf("  error: %s\n", cudaGetErrorString(e));
  }
  return 0;
}
