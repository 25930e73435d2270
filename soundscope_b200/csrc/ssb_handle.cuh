// ssb_handle.cuh — the handle behind the C ABI and the helpers its translation units share (capi.cu,
// capture.cu).  Not part of the public ABI.
#pragma once

#include <stdarg.h>
#include <stdio.h>

#include "ssb_internal.cuh"

struct ssb_analyzer {
  using LoudParams = ssb::LoudParams;
  using GateParams = ssb::GateParams;
  using LoudState = ssb::LoudState;
  using FftPlan = ssb::FftPlan;
  int device = 0;
  int sm_count = 0;           // multiprocessors of `device` (queried at create; grids are sized from it)
  cudaStream_t own_stream = nullptr, stream = nullptr;
  uint32_t channels = 0, rate = 0;
  int32_t mode = 0;
  size_t n_streams = 0;
  uint32_t flags = 0;

  LoudParams lp{};
  GateParams gp{};
  LoudState st{};
  uint64_t total_frames = 0;  // frames fed per stream since the last reset
  uint64_t gated_upto = 0;    // buckets [0, gated_upto) have been entered into the histograms
  size_t ring_pos = 0;

  double* d_hist_tables = nullptr;  // energies[1000] | boundaries[1001]
  double* d_scan_powers = nullptr;  // Pt^m table of the scan kernel (depends on the rate)
  ssb_analyzer* oneshot = nullptr;  // cached Mode::all() meter of calculate_integrated_lufs
  double* d_results = nullptr;
  double* h_results = nullptr;  // pinned
  bool results_valid = false;
  bool lra_cache_valid = false; // st.cache[].lra is every stream's LRA for the current short-term histograms
  bool icache_valid = false;    // st.cache[]'s gating sums match the block histograms (StreamCache)
  bool lean_enabled = true;     // SSB_RESULTS_LEAN=0 at create: always the full histogram scan (A/B, tests)
  bool dres_valid = false;      // d_results already holds the rows for the current feed position (fused epilogue)
  bool meter_ok = false;        // false while (re)initialisation failed half-way: every meter call then fails loudly

  float* d_stage[2] = {nullptr, nullptr};
  size_t stage_cap = 0;  // floats per staging buffer
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
  int stage_idx = 0;
  float* d_scratch = nullptr;  // outputs of single-shot host calls
  size_t scratch_cap = 0;      // bytes
  void* h_scratch = nullptr;   // pinned mirror
  void* d_pcm = nullptr;       // raw PCM bytes of ssb_add_frames_pcm (host form)
  size_t pcm_cap = 0;
  size_t h_scratch_cap = 0;

  std::map<std::pair<size_t, uint32_t>, FftPlan> plans;
  std::map<std::pair<size_t, uint32_t>, std::pair<std::vector<double>, std::vector<double>>> axes;

  struct Gather {
    void* base = nullptr;                        // [2][world][n][stride] f64 rows | flags[world] u64 (at +rows_bytes) | counter (at +rows_bytes+128)
    void* peer_base[ssb::kMaxGatherRanks] = {};  // every rank's allocation as mapped here (own = base)
    size_t rows_bytes = 0;
    int world = 0, rank = 0, parity = 0;
    bool open = false;
    unsigned long long epoch = 0;                // ssb_gather_wait calls so far
  } gather;

  uint64_t launches = 0;
  int force_kernel = 0;  // tests: 0 auto, 1 generic kernel, 2 serial rows kernel, 3 round-1 tile kernel, 4 scan kernel,
                         // 5 k_loudness_wtile (mixed T4/T5 warps), 6 k_loudness_wtile (uniform T4 warps)
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  size_t prof_used = 0;
  int32_t tick_fft_status[2] = {0, 0};  // mid / side get_fft statuses of the last process_tick / mic_tick
  char err[256] = {0};
};

namespace ssb {

inline int32_t fail(ssb_analyzer* h, int32_t code, const char* fmt, ...) {
  if (h) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(h->err, sizeof(h->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

inline int32_t cuda_fail(ssb_analyzer* h, cudaError_t e, const char* what) {
  return fail(h, SSB_ERR_CUDA + (int32_t)e, "%s: %s", what, cudaGetErrorString(e));
}

#define CK(call)                                                  \
  do {                                                            \
    cudaError_t e__ = (call);                                     \
    if (e__ != cudaSuccess) return cuda_fail(h, e__, #call);      \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// gather.cu: GatherArgs of this handle's results launches (world == 0 when no gather is open)
GatherArgs peek_gather_args(ssb_analyzer* h);

// capi.cu
int32_t ensure_stage(ssb_analyzer* h, size_t floats);
int32_t ensure_scratch(ssb_analyzer* h, size_t bytes);          // device scratch + pinned host mirror of the same size
int32_t ensure_scratch_device(ssb_analyzer* h, size_t bytes);   // device scratch only
// feed `frames` frames per stream from device memory laid out [stream][in_stride_frames][C]
// d_results (nullable): where the caller wants the result rows for the position after this feed; *written says
// whether the filter launch produced them itself (fused epilogue) — otherwise the caller launches k_results
int32_t feed_device(ssb_analyzer* h, const float* d_in, size_t frames, size_t in_stride_frames,
                    double* d_results = nullptr, bool* written = nullptr);
int32_t get_plan(ssb_analyzer* h, size_t n, uint32_t rate, FftPlan** out);
int32_t fft_shape_check(size_t n, uint32_t rate);
// launches k_results for stream rows (gating pending buckets first) and leaves them in h->d_results; no copy
int32_t launch_results_now(ssb_analyzer* h);
// cached (x, tilt) of get_fft for (n, rate)
const std::pair<std::vector<double>, std::vector<double>>& fft_axis_cached(ssb_analyzer* h, size_t n, uint32_t rate);
// get_waveform's column count for (waveform_window, len): `(w * 1000.) as usize` columns, cut at the first column
// whose start is past the end (analyzer.rs:113-124); *window_out = the `window` the reference divides by
size_t waveform_window_columns(double waveform_window, size_t len, size_t* window_out);

}  // namespace ssb
