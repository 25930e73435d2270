// bench_cabi.cu — the cfg2-style loudness step driven through the C ABI from plain C++/CUDA (no Python, no torch):
// what a native host (the reference's Rust shim) would do, and a fast A/B driver for kernel work
// (starts in milliseconds, so a GPU-box minute holds dozens of runs).
//
//   nvcc -O2 -std=c++17 -I include tools/bench_cabi.cu -o tools/bench_cabi -L soundscope_b200 -lsoundscope_b200 \
//        -Xlinker -rpath -Xlinker $PWD/soundscope_b200
//   tools/bench_cabi [n_streams=4096] [frames=19200] [channels=2] [rate=48000] [mode=loudness|all] [steps=20] [warmup=3]
//
// Inputs are generated on the device (two alternating buffers, each larger than L2 at the default shape); every step is
// ssb_add_frames_f32_device + ssb_results_device on one stream, timed with CUDA events; the filter kernel's own time
// comes from ssb_profile_read.  Prints one JSON line.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "soundscope_b200.h"

#define CHECK_CUDA(x)                                                                    \
  do {                                                                                   \
    cudaError_t e_ = (x);                                                                \
    if (e_ != cudaSuccess) {                                                             \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      return 2;                                                                          \
    }                                                                                    \
  } while (0)
#define CHECK_SSB(h, x)                                                                        \
  do {                                                                                         \
    int32_t rc_ = (x);                                                                         \
    if (rc_ != SSB_OK) {                                                                       \
      fprintf(stderr, "%s:%d %s -> %d (%s)\n", __FILE__, __LINE__, #x, rc_, ssb_last_error(h)); \
      return 3;                                                                                \
    }                                                                                          \
  } while (0)

// per-stream tone 100 * 2^((s % 64) / 8) Hz at 0.25 plus 0.05 of hashed noise (the shape of bench.py's generator)
__global__ void k_fill(float* x, size_t n_streams, size_t frames, int channels, float rate, unsigned seed) {
  const size_t total = n_streams * frames * (size_t)channels;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t s = i / (frames * channels);
    const size_t f = (i / channels) % frames;
    const int c = (int)(i % channels);
    const float freq = 100.0f * exp2f((float)(s % 64) / 8.0f);
    unsigned h = (unsigned)(i * 2654435761u) ^ seed;
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    const float noise = (float)(h >> 8) * (1.0f / 8388608.0f) - 1.0f;
    x[i] = 0.25f * sinf(6.2831853f * freq * (float)f / rate + 0.7f * (float)c + (float)(s % 17)) + 0.05f * noise;
  }
}

int main(int argc, char** argv) {
  const size_t n_streams = argc > 1 ? strtoull(argv[1], nullptr, 10) : 4096;
  const size_t frames = argc > 2 ? strtoull(argv[2], nullptr, 10) : 19200;
  const int channels = argc > 3 ? atoi(argv[3]) : 2;
  const unsigned rate = argc > 4 ? (unsigned)atoi(argv[4]) : 48000;
  const bool all = argc > 5 && strcmp(argv[5], "all") == 0;
  const int steps = argc > 6 ? atoi(argv[6]) : 20;
  const int warmup = argc > 7 ? atoi(argv[7]) : 3;
  const int32_t mode = all ? SSB_MODE_ALL : (SSB_MODE_I | SSB_MODE_LRA | SSB_MODE_HISTOGRAM);

  cudaStream_t stream;
  CHECK_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  const size_t n = n_streams * frames * (size_t)channels;
  float* x[2];
  for (int i = 0; i < 2; i++) {
    CHECK_CUDA(cudaMalloc(&x[i], n * sizeof(float)));
    k_fill<<<148 * 8, 256, 0, stream>>>(x[i], n_streams, frames, channels, (float)rate, 1234u + 17u * (unsigned)i);
  }
  CHECK_CUDA(cudaGetLastError());

  ssb_analyzer* h = nullptr;
  int32_t rc = ssb_analyzer_create(&h, (uint32_t)channels, rate, mode, n_streams, -1, 0);
  if (rc != SSB_OK) {
    fprintf(stderr, "ssb_analyzer_create -> %d\n", rc);
    return 3;
  }
  CHECK_SSB(h, ssb_set_stream(h, stream));
  double* d_res = nullptr;
  CHECK_CUDA(cudaMalloc(&d_res, n_streams * ssb_result_stride(h) * sizeof(double)));

  for (int i = 0; i < warmup; i++) {
    CHECK_SSB(h, ssb_add_frames_f32_device(h, x[i & 1], frames));
    CHECK_SSB(h, ssb_results_device(h, d_res));
  }
  CHECK_CUDA(cudaStreamSynchronize(stream));
  CHECK_SSB(h, ssb_profile_enable(h, 1));
  const uint64_t l0 = ssb_launch_count(h);
  cudaEvent_t e0, e1;
  CHECK_CUDA(cudaEventCreate(&e0));
  CHECK_CUDA(cudaEventCreate(&e1));
  CHECK_CUDA(cudaEventRecord(e0, stream));
  for (int i = 0; i < steps; i++) {
    CHECK_SSB(h, ssb_add_frames_f32_device(h, x[i & 1], frames));
    CHECK_SSB(h, ssb_results_device(h, d_res));
  }
  CHECK_CUDA(cudaEventRecord(e1, stream));
  CHECK_CUDA(cudaStreamSynchronize(stream));
  float ms = 0.f;
  CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  double filt_ms = 0.0;
  uint64_t filt_n = 0;
  CHECK_SSB(h, ssb_profile_read(h, &filt_ms, &filt_n));
  const uint64_t launches = ssb_launch_count(h) - l0;

  // one result row back on the host as a sanity check that the step produced numbers
  double row[4 + 2 * 64];
  CHECK_CUDA(cudaMemcpy(row, d_res, ssb_result_stride(h) * sizeof(double), cudaMemcpyDeviceToHost));
  const double samples = (double)n * steps;
  const double kern_ms = filt_n ? filt_ms / (double)filt_n : 0.0;
  printf("{\"driver\": \"bench_cabi\", \"n_streams\": %zu, \"frames\": %zu, \"channels\": %d, \"rate\": %u, \"mode\": \"%s\", "
         "\"steps\": %d, \"ms_per_step\": %.6f, \"samples_per_s\": %.6e, \"filter_kernel_ms\": %.6f, "
         "\"filter_algorithmic_gbs\": %.1f, \"gpu_launches\": %llu, \"stream0_momentary_lufs\": %.6f, \"stream0_integrated_lufs\": %.6f}\n",
         n_streams, frames, channels, rate, all ? "all" : "loudness", steps, ms / steps, samples / (ms * 1e-3), kern_ms,
         kern_ms > 0 ? (double)n * 4.0 / (kern_ms * 1e-3) / 1e9 : 0.0, (unsigned long long)launches, row[0], row[2]);
  ssb_analyzer_destroy(h);
  cudaFree(d_res);
  cudaFree(x[0]);
  cudaFree(x[1]);
  return 0;
}
