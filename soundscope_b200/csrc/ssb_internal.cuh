// ssb_internal.cuh — shared declarations of libsoundscope_b200.so (not part of the public ABI).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/soundscope_b200.h"

namespace ssb {

constexpr int kMaxChannels = 64;
constexpr int kNB = 64;              // per-100 ms energy buckets kept per (stream, channel)
constexpr int kMaxBucketsPerLaunch = kNB - 31;  // so a 3 s window ending at any bucket of the launch is intact
constexpr int kHistBins = 1000;
constexpr int kTpHist = 24;          // input samples of true-peak history kept per (stream, channel)

// Parameters of the K-weighting + peak kernels; passed by value (lives in the constant bank).
struct LoudParams {
  double b[5];
  double a[5];          // a[0] == 1, unused
  double na[5];         // -a[i]: constant-bank DFMA operands (a negated register operand costs a third RF read)
  float tp4[3][12];     // factor-4 interpolator, phases 1..3, tap t multiplies x[n-t]
  float tp2[24];        // factor-2 interpolator, phase 1
  int tp_factor;        // 0 (none), 2 or 4 — ebur128's rate rule
  int channels;
  uint32_t s100;        // samples_in_100ms
  int do_filter, do_sample_peak, do_true_peak;
  uint64_t active_mask; // bit c set: channel c is not Channel::Unused
  double handoff[16];   // D A^64 D: segment hand-off matrix of the time-segmented kernels (tile_handoff_matrix)
  double handoff80[16]; // D A^80 D: the 80-frame segments of k_loudness_wtile's type-A warps
};

struct GateParams {
  int channels;
  uint32_t s100;
  int do_i, do_lra;
  float weight[kMaxChannels];  // 1.0 / 1.41 / 0.0 (unused)
};

// --- host-side tables (host_tables.cu) ---------------------------------------------------------
void kweight_coeffs(uint32_t rate, double b[5], double a[5]);
void default_channel_weights(uint32_t channels, float w[kMaxChannels], uint64_t* active_mask);
int truepeak_taps(uint32_t rate, float tp4[3][12], float tp2[24], int force_factor = 0);   // returns factor 0/2/4
void histogram_tables(double energies[1000], double boundaries[1001]);
double histogram_bound0();                                           // boundaries[0]
float libm_cosf(float x);                                            // libm 0.2.16 (musl) cosf
void hann_multipliers(size_t n, std::vector<float>& w);             // spectrum-analyzer hann_window
size_t fft_bin_range(size_t n, uint32_t rate, size_t* k_first);
void fft_axis(size_t n, uint32_t rate, std::vector<double>& x, std::vector<double>& tilt, size_t* k_first);

// Per-stream running sums of the histogram gating (one 64-byte line per stream).  ebur128's gated_loudness walks the
// 1000-bin block histogram twice per query; between two queries only the newly gated blocks change it, so the lean
// results path (loudness_results.cuh: results_lean) keeps what the walk computes and patches it: the count and energy
// sum of every block above the absolute gate, the relative gate's start bin for those sums, and the count and energy
// sum of the bins from `start` up.  `lra` is the loudness range as of the last short-term histogram change.  A full
// scan (results_for_stream) rewrites all of it; the host tracks validity per handle (ssb_analyzer::icache_valid /
// lra_cache_valid).  All zero is the correct cache of an empty meter.
struct alignas(64) StreamCache {
  unsigned long long n_all;
  double sum_all;
  unsigned long long n_above;
  double sum_above;
  long long start;
  double lra;
  unsigned long long n_st;   // short-term (3 s) entries above the absolute gate, and the sum of their bin energies:
  double sum_st;             // the first pass of ebur128's loudness_range (valid with `lra`)
};

// --- kernel launchers (loudness.cu) ------------------------------------------------------------
struct LoudState {
  size_t n_streams;
  double* filt;       // [n][C][4]
  double* bucket;     // [n][C][kNB]
  uint32_t* block_hist;  // [n][1000]
  uint32_t* st_hist;     // [n][1000]
  float* speak;       // [n][C]
  float* tpeak;       // [n][C]
  float* tphist;      // [n][C][kTpHist]
  double* ring;       // [n][ring_frames][C] or nullptr
  size_t ring_frames;
  double* ring_e;     // [n][2] scratch: momentary / short-term energies from the ring
  StreamCache* cache; // [n] running gating sums + loudness range (see StreamCache)
  const double* hist_energies;    // [1000]
  const double* hist_boundaries;  // [1001]
};

// Arguments of the per-stream result / gating code (loudness_results.cuh): k_results and the fused epilogue of
// k_loudness_wtile take the same struct.
// Publication of the result rows to every rank's gather buffer (gather.cu); world == 0: none.  The stores are
// fire-and-forget; arrival is established once per ssb_gather_wait (a system-scope fence + flag exchange in its own
// small kernel after the last results launch in stream order), not per launch.
constexpr int kMaxGatherRanks = 8;
struct GatherArgs {
  int world, rank;
  double* rows[kMaxGatherRanks];                // per destination rank: block `rank` of the selected parity
};

struct ResultsArgs {
  const double* bucket;          // [n][C][kNB]
  const uint32_t* block_hist;    // [n][1000] (read through L2 after the gating atomics)
  const uint32_t* st_hist;
  uint32_t* block_hist_rw;
  uint32_t* st_hist_rw;
  const float* speak;
  const float* tpeak;
  const double* ring;            // or nullptr
  size_t ring_frames, ring_pos;
  const double* ring_e;          // [n][2] from k_ring_energy, or nullptr
  const double* energies;        // [1000] bin-centre energies
  const double* bounds;          // [1001] bin boundaries
  uint64_t buckets_done;
  int aligned, mode;
  double* out;                   // [n][4 + 2C]
  uint64_t gate_first, gate_last;  // buckets to enter into the histograms first (none when gate_last < gate_first)
  GatherArgs ga;                   // zero unless a gather is open on the handle
  // The short-term histogram only changes when a 3 s entry is gated (once per second of audio), so the loudness range
  // is kept per stream: lra_from_cache != 0 -> read it instead of loading and scanning the 1000 bins again.  The caller
  // sets it only when no pending bucket of this call is a short-term entry and the cache is current.
  StreamCache* cache;
  int lra_from_cache;
  // lean != 0: the block-histogram sums in `cache` are current and at most kLeanPending buckets are pending, so the
  // integrated loudness is patched instead of re-scanned (results_lean: 16 lanes per stream).
  int lean;
  // lra_fast != 0 (lean launches that gate a 3 s entry while the cache is current): the lean code gates the entry
  // (st_back = how many buckets before the last completed one it ends) and lra_scan_fast finishes from the cached sums
  int lra_fast, st_back;
  double bound0;                   // boundaries[0]: the absolute gate (-70 LUFS) as an energy
};
constexpr int kLeanPending = 10;   // pending buckets a lean call can gate (one lane each; at most one 3 s entry among them)
constexpr int kLeanSlots = 40;     // buckets per channel a lean call looks back over (3 s window of the oldest pending entry)
constexpr int kLeanMaxChannels = 6;   // 4 + 2 C result values fit the 16 lanes of a stream

inline ResultsArgs make_results_args(const LoudState& st, uint64_t buckets_done, int aligned, size_t ring_pos, int mode,
                                     double* d_out, uint64_t gate_first, uint64_t gate_last, const double* ring_e) {
  ResultsArgs ra{};
  ra.bucket = st.bucket;
  ra.block_hist = st.block_hist;
  ra.st_hist = st.st_hist;
  ra.block_hist_rw = st.block_hist;
  ra.st_hist_rw = st.st_hist;
  ra.speak = st.speak;
  ra.tpeak = st.tpeak;
  ra.ring = st.ring;
  ra.ring_frames = st.ring_frames;
  ra.ring_pos = ring_pos;
  ra.ring_e = ring_e;
  ra.energies = st.hist_energies;
  ra.bounds = st.hist_boundaries;
  ra.buckets_done = buckets_done;
  ra.aligned = aligned;
  ra.mode = mode;
  ra.out = d_out;
  ra.gate_first = gate_first;
  ra.gate_last = gate_last;
  ra.cache = st.cache;
  ra.lra_from_cache = 0;
  ra.lean = 0;
  ra.lra_fast = 0;
  ra.st_back = -1;
  ra.bound0 = histogram_bound0();
  return ra;
}


// Second-generation batch kernel (loudness_wtile.cu): mono / stereo, no ring; variant 0 = mixed T4/T5 warps on 320-frame
// tiles, variant 1 = uniform T4 warps on 256-frame tiles.  Consumes the leading whole tiles; with `ra` and a chunk of
// whole tiles it also gates the completed buckets and writes the result rows (*results_written).
int wtile_frames(int variant);
bool wtile_path_usable(const LoudParams& p, const LoudState& st, const float* d_in, size_t frames,
                       size_t in_stride_frames, int variant);
cudaError_t launch_loudness_wtile(const LoudParams& p, const LoudState& st, const GateParams& gp, const float* d_in,
                                  size_t frames, size_t in_stride_frames, uint32_t pos0, uint64_t bucket0, int variant,
                                  const ResultsArgs* ra, int sm_count, int device, cudaStream_t s, uint64_t* launches,
                                  size_t* consumed, bool* results_written);
// stream count from which the serial many-streams kernel replaces the time-segmented ones (SSB_SERIAL_MIN overrides)
size_t serial_min_streams();

// Filters `frames` frames per stream starting `pos0` frames into 100 ms bucket number `bucket0`.
// At most kMaxBucketsPerLaunch buckets may complete inside one call.
cudaError_t launch_loudness_generic(const LoudParams& p, const LoudState& st, const float* d_in,
                                    size_t frames, size_t in_stride_frames, uint32_t pos0,
                                    uint64_t bucket0, size_t ring_pos, cudaStream_t s, uint64_t* launches);
// TMA-tiled, time-segmented fast path (loudness_tile.cu): mono/stereo, no ring, no true peak.  Consumes the
// leading whole tiles of the chunk; the caller sends the remainder through the generic kernel.
bool tile_path_usable(const LoudParams& p, const LoudState& st, const float* d_in, size_t frames,
                      size_t in_stride_frames);
cudaError_t launch_loudness_tile(const LoudParams& p, const LoudState& st, const float* d_in, size_t frames,
                                 size_t in_stride_frames, uint32_t pos0, uint64_t bucket0, cudaStream_t s,
                                 uint64_t* launches, size_t* consumed, int force_kernel, int sm_count);
// D A^64 D for the K-weighting denominator a[] (double-double on the host, rounded once); cached in LoudParams
void tile_handoff_matrix(const double a[5], double P[16]);
// D A^n D for any n (loudness_tile.cu); the scan kernel's table holds n = 64 m, m = 0..32
void tile_handoff_power(const double a[5], int n, double P[16]);
// Few-streams / long-audio kernel (loudness_scan.cu): one CTA per stream, time-parallel prefix scan of the
// filter state; mono/stereo, any chunking, optional ring.
void scan_power_table(const double a[5], double* host_table);
int scan_power_table_doubles();
bool scan_path_usable(const LoudParams& p, const LoudState& st, size_t frames);
cudaError_t launch_loudness_scan(const LoudParams& p, const LoudState& st, const double* d_powers, const float* d_in,
                                 size_t frames, size_t in_stride_frames, uint32_t pos0, uint64_t bucket0,
                                 size_t ring_pos, cudaStream_t s, uint64_t* launches);
// Whole-file one-shot: time-chunked scan of ONE stream from the reset state into d_file_buckets[C][bucket_stride]
// (energy sum of every complete 100 ms bucket by global index), then gating of all its blocks into the histograms.
cudaError_t launch_loudness_scan_file(const LoudParams& p, const LoudState& st, const double* d_powers, const float* d_in,
                                      size_t frames, double* d_file_buckets, size_t bucket_stride, size_t chunk_buckets,
                                      cudaStream_t s, uint64_t* launches);
cudaError_t launch_file_gating(const GateParams& g, const LoudState& st, const double* d_file_buckets, size_t bucket_stride,
                               uint64_t n_buckets, cudaStream_t s, uint64_t* launches);
// Gating for buckets [j_first, j_last] completed by the preceding filter launch.
cudaError_t launch_gating(const GateParams& g, const LoudState& st, uint64_t j_first, uint64_t j_last,
                          cudaStream_t s, uint64_t* launches);
// Per-stream scalars -> d_out[n][4+2C]; aligned != 0: the feed position is on the 100 ms grid.  Buckets
// [gate_first, gate_last] (none when gate_last < gate_first) are gated inside the same launch first.
cudaError_t launch_results(const GateParams& g, const LoudState& st, uint64_t buckets_done, int aligned,
                           size_t ring_pos, int mode, double* d_out, cudaStream_t s, uint64_t* launches,
                           uint64_t gate_first, uint64_t gate_last, const GatherArgs* ga = nullptr, int lra_from_cache = 0,
                           int lean = 0, int st_back = -1);
cudaError_t launch_reset(const LoudState& st, int channels, cudaStream_t s, uint64_t* launches);
cudaError_t launch_histogram_index(const LoudState& st, const double* d_e, size_t n, int32_t* d_out, cudaStream_t s);

// --- kernel launchers (spectrum.cu) ------------------------------------------------------------
struct FftPlan {
  size_t n = 0;
  uint32_t rate = 0;
  size_t k_first = 0, n_bins = 0;
  float* d_window = nullptr;   // [n] Hann multipliers
  float2* d_twiddle = nullptr; // [n/2] exp(-j*2*pi*k/n)
  float2* d_tw_lo = nullptr;   // [64]   W_n^i           (two-level table of the fast kernel)
  float2* d_tw_hi = nullptr;   // [n/64] W_n^(64 i)
  double* d_tilt = nullptr;    // [n_bins] the f64 tilt of analyzer.rs:80-94 per kept bin (launch_fft_y)
};
cudaError_t launch_fft(const FftPlan& plan, const float* d_in, int layout, size_t n_windows,
                       float* d_db_out, int32_t* d_status, cudaStream_t s, uint64_t* launches);
cudaError_t launch_fft_y(const FftPlan& plan, const float* d_in, int layout, size_t n_windows,
                         double* d_y_out, int32_t* d_status, cudaStream_t s, uint64_t* launches);
cudaError_t launch_waveform(const float* d_samples, size_t len, size_t window, float* d_minmax,
                            size_t columns, cudaStream_t s, uint64_t* launches);
cudaError_t launch_mid_side(const float* d_in, size_t frames, float* d_mid, float* d_side,
                            cudaStream_t s, uint64_t* launches);

}  // namespace ssb
