// mb_fma.cu — FP32 FMA-pipe issue rates on one SM sub-partition (B200): warp-instructions per cycle per SMSP for
//   FFMA reg*reg+reg, FFMA reg*const-bank+reg, FFMA reg*imm+reg, FFMA2 reg*reg+reg, FFMA2 reg*uniform-scalar+reg,
// with 1..8 warps per sub-partition and 8 independent accumulator chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mb_fma tools/mb_fma.cu
#include <cstdio>
#include <cuda_runtime.h>

struct K { float c[8]; };
constexpr int ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(const __grid_constant__ K kc, float* out, long long* cyc, float seed) {
  float a[8], x[8];
  float2 a2[8], x2[8];
  for (int i = 0; i < 8; i++) {
    a[i] = seed * (threadIdx.x + i);
    x[i] = seed + i;
    a2[i] = make_float2(a[i], a[i] + 1.f);
    x2[i] = make_float2(x[i], x[i] + 2.f);
  }
  float r0 = seed * 3.f, r1 = seed * 5.f;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) a[i] = fmaf(x[i], (i & 1) ? r0 : r1, a[i]);          // reg * reg + reg
      if (MODE == 1) a[i] = fmaf(x[i], kc.c[i], a[i]);                   // reg * const bank + reg
      if (MODE == 2) a[i] = fmaf(x[i], 0.99991f + 0.00001f * i, a[i]);   // reg * imm + reg
      if (MODE == 3) a2[i] = __ffma2_rn(x2[i], make_float2(r0, r1), a2[i]);                 // packed, reg operands
      if (MODE == 4) a2[i] = __ffma2_rn(x2[i], make_float2(kc.c[i], kc.c[i]), a2[i]);       // packed, uniform scalar broadcast
      if (MODE == 5) a2[i] = __ffma2_rn(x2[i], make_float2(kc.c[i], kc.c[(i + 1) & 7]), a2[i]);   // packed, uniform pair
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 8; i++) s += a[i] + a2[i].x + a2[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, const K& kc, float* d_out, long long* d_cyc) {
  printf("%-44s", name);
  for (int warps_per_smsp = 1; warps_per_smsp <= 8; warps_per_smsp *= 2) {
    const int threads = warps_per_smsp * 4 * 32;
    k<MODE><<<148, threads>>>(kc, d_out, d_cyc, 1e-6f);
    k<MODE><<<148, threads>>>(kc, d_out, d_cyc, 1e-6f);
    cudaDeviceSynchronize();
    long long hc[148];
    cudaMemcpy(hc, d_cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; i++) avg += (double)hc[i];
    avg /= 148;
    const double instr_per_smsp = (double)ITERS * 8 * warps_per_smsp;
    printf("  %dw: %.2f cyc/instr", warps_per_smsp, avg / instr_per_smsp);
  }
  cudaError_t e = cudaGetLastError();
  printf("%s\n", e ? cudaGetErrorString(e) : "");
}

int main() {
  K kc;
  for (int i = 0; i < 8; i++) kc.c[i] = 0.9999f + 1e-5f * i;
  float* d_out; long long* d_cyc;
  cudaMalloc(&d_out, 148 * 1024 * sizeof(float));
  cudaMalloc(&d_cyc, 148 * sizeof(long long));
  printf("cycles per warp-instruction per SM sub-partition (lower = faster; 8 independent chains per thread)\n");
  run<0>("FFMA  reg * reg + reg", kc, d_out, d_cyc);
  run<1>("FFMA  reg * const-bank + reg", kc, d_out, d_cyc);
  run<2>("FFMA  reg * immediate + reg", kc, d_out, d_cyc);
  run<3>("FFMA2 reg * reg + reg", kc, d_out, d_cyc);
  run<4>("FFMA2 reg * uniform scalar (broadcast) + reg", kc, d_out, d_cyc);
  run<5>("FFMA2 reg * uniform pair + reg", kc, d_out, d_cyc);
  return 0;
}
