// Host emulation of the shared-memory FFT core (soundscope_b200/csrc/fft_core.cuh): the same
// __host__ __device__ stage functions, with the CTA's thread loop run sequentially.  Built by
// tests/test_fft_core_host.py into a throw-away .so; lets the CPU test suite check the device algorithm's
// index arithmetic, twiddles and digit reversal against numpy without a GPU.
#include <math.h>
#include <vector>

#include "../soundscope_b200/csrc/fft_core.cuh"

using namespace ssb;

extern "C" int fft_core_host(const float* in_re_im, unsigned M, unsigned N, float* out_re_im) {
  if (M < 512 || M > 16384 || (M & (M - 1)) || N < M || (N & (N - 1))) return 1;
  std::vector<float2> lo(64), hi(N / 64 ? N / 64 : 1);
  for (unsigned i = 0; i < 64; i++) {
    const double a = 2.0 * M_PI * (double)i / (double)N;
    lo[i] = make_float2((float)cos(a), (float)-sin(a));
  }
  for (unsigned i = 0; i < N / 64; i++) {
    const double a = 2.0 * M_PI * (double)(64 * i) / (double)N;
    hi[i] = make_float2((float)cos(a), (float)-sin(a));
  }
  FftTwiddle tw{lo.data(), hi.data()};
  std::vector<float2> z(M + M / 32 + M / 512 + 1);
  for (unsigned p = 0; p < M; p++) z[fft_pad(p)] = make_float2(in_re_im[2 * p], in_re_im[2 * p + 1]);
  const unsigned R1 = M / 512;
  for (unsigned j = 0; j < M / (R1 ? R1 : 1) && R1 > 1; j++) {
    switch (R1) {
      case 2: fft_stage1<2>(z.data(), M, N, tw, j); break;
      case 4: fft_stage1<4>(z.data(), M, N, tw, j); break;
      case 8: fft_stage1<8>(z.data(), M, N, tw, j); break;
      case 16: fft_stage1<16>(z.data(), M, N, tw, j); break;
      case 32: fft_stage1<32>(z.data(), M, N, tw, j); break;
    }
  }
  for (unsigned t = 0; t < M / 16; t++) fft_stage2(z.data(), N, tw, t);
  for (unsigned t = 0; t < M / 32; t++) fft_stage3(z.data(), t);
  unsigned sh = 0;
  while ((512u << sh) < M) sh++;
  for (unsigned k = 0; k < M; k++) {
    const float2 v = z[fft_position(k, sh)];
    out_re_im[2 * k] = v.x;
    out_re_im[2 * k + 1] = v.y;
  }
  return 0;
}
