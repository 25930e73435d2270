// capi.cu — the C ABI declared in include/soundscope_b200.h: handle, argument checks in the
// reference's error vocabulary, staging, launch scheduling.  No compute happens on the host:
// every entry point either launches the CUDA kernels or fails.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <memory>
#include <new>

#include "ssb_handle.cuh"

using namespace ssb;


namespace ssb {

namespace {

void free_meter(ssb_analyzer* h) {
  cudaFree(h->st.filt); cudaFree(h->st.bucket); cudaFree(h->st.block_hist); cudaFree(h->st.st_hist);
  cudaFree(h->st.speak); cudaFree(h->st.tpeak); cudaFree(h->st.tphist); cudaFree(h->st.ring); cudaFree(h->st.ring_e);
  cudaFree(h->st.cache);
  cudaFree(h->d_results);
  cudaFree(h->d_scan_powers);
  h->d_scan_powers = nullptr;
  if (h->h_results) cudaFreeHost(h->h_results);
  h->st = LoudState{};
  h->d_results = nullptr;
  h->h_results = nullptr;
}

// (re)build everything that depends on (channels, rate): EbuR128::new
int32_t init_meter(ssb_analyzer* h, uint32_t channels, uint32_t rate) {
  if (channels == 0 || channels > (uint32_t)kMaxChannels)
    return fail(h, SSB_ERR_NOMEM, "EbuR128::new: channels %u outside 1..=64 (Error::NoMem)", channels);
  if (rate < 16 || rate > 2822400)
    return fail(h, SSB_ERR_NOMEM, "EbuR128::new: rate %u outside 16..=2822400 (Error::NoMem)", rate);
  h->meter_ok = false;
  free_meter(h);
  h->channels = channels;
  h->rate = rate;
  LoudParams& lp = h->lp;
  memset(&lp, 0, sizeof(lp));
  kweight_coeffs(rate, lp.b, lp.a);
  for (int i = 0; i < 5; i++) lp.na[i] = -lp.a[i];
  tile_handoff_matrix(lp.a, lp.handoff);
  tile_handoff_power(lp.a, 80, lp.handoff80);
  if (channels <= 2 && h->n_streams <= 64) {
    std::vector<double> tab(scan_power_table_doubles());
    scan_power_table(lp.a, tab.data());
    CK(cudaMalloc(&h->d_scan_powers, tab.size() * sizeof(double)));
    CK(cudaMemcpy(h->d_scan_powers, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  lp.channels = (int)channels;
  lp.s100 = (rate + 5) / 10;
  lp.do_filter = 1;  // every mode contains M
  lp.do_sample_peak = (h->mode & SSB_MODE_SAMPLE_PEAK) == SSB_MODE_SAMPLE_PEAK;
  lp.do_true_peak = (h->mode & SSB_MODE_TRUE_PEAK) == SSB_MODE_TRUE_PEAK;
  lp.tp_factor = lp.do_true_peak ? truepeak_taps(rate, lp.tp4, lp.tp2) : 0;
  GateParams& gp = h->gp;
  memset(&gp, 0, sizeof(gp));
  gp.channels = (int)channels;
  gp.s100 = lp.s100;
  gp.do_i = (h->mode & SSB_MODE_I) == SSB_MODE_I;
  gp.do_lra = (h->mode & SSB_MODE_LRA) == SSB_MODE_LRA;
  default_channel_weights(channels, gp.weight, &lp.active_mask);

  const size_t n = h->n_streams, chains = n * channels;
  LoudState& st = h->st;
  st.n_streams = n;
  CK(cudaMalloc(&st.filt, chains * 4 * sizeof(double)));
  CK(cudaMalloc(&st.bucket, chains * kNB * sizeof(double)));
  CK(cudaMalloc(&st.block_hist, n * kHistBins * sizeof(uint32_t)));
  CK(cudaMalloc(&st.st_hist, n * kHistBins * sizeof(uint32_t)));
  CK(cudaMalloc(&st.speak, chains * sizeof(float)));
  CK(cudaMalloc(&st.tpeak, chains * sizeof(float)));
  CK(cudaMalloc(&st.tphist, chains * kTpHist * sizeof(float)));
  CK(cudaMalloc(&st.cache, n * sizeof(StreamCache)));
  st.ring = nullptr;
  st.ring_e = nullptr;
  st.ring_frames = 0;
  if (h->flags & SSB_FLAG_RING) {
    // ebur128: 3 s (mode S) or 400 ms of filtered samples, rounded up to a multiple of samples_in_100ms
    const size_t window_ms = ((h->mode & SSB_MODE_S) == SSB_MODE_S) ? 3000 : 400;
    size_t rf = (size_t)rate * window_ms / 1000;
    if (rf % lp.s100) rf = rf + lp.s100 - (rf % lp.s100);
    st.ring_frames = rf;
    CK(cudaMalloc(&st.ring, n * rf * channels * sizeof(double)));
    CK(cudaMalloc(&st.ring_e, n * 2 * sizeof(double)));
  }
  st.hist_energies = h->d_hist_tables;
  st.hist_boundaries = h->d_hist_tables + 1000;
  const size_t stride = 4 + 2 * (size_t)channels;
  CK(cudaMalloc(&h->d_results, n * stride * sizeof(double)));
  CK(cudaMallocHost(&h->h_results, n * stride * sizeof(double)));
  h->total_frames = 0;
  h->gated_upto = 0;
  h->ring_pos = 0;
  h->results_valid = false;
  h->dres_valid = false;
  CK(launch_reset(st, (int)channels, h->stream, &h->launches));
  h->lra_cache_valid = h->icache_valid = true;   // zeroed by the reset: the cache of an empty meter
  {
    const char* e = getenv("SSB_RESULTS_LEAN");   // A/B switch: 0 = always the full histogram scan
    h->lean_enabled = !(e && atoi(e) == 0);
  }
  h->meter_ok = true;
  return SSB_OK;
}

}  // namespace

int32_t ensure_stage(ssb_analyzer* h, size_t floats) {
  if (floats <= h->stage_cap) return SSB_OK;
  CK(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < 2; i++) {
    cudaFree(h->d_stage[i]);
    h->d_stage[i] = nullptr;
    CK(cudaMalloc(&h->d_stage[i], floats * sizeof(float)));
  }
  h->stage_cap = floats;
  return SSB_OK;
}

int32_t ensure_scratch_device(ssb_analyzer* h, size_t bytes) {
  if (bytes > h->scratch_cap) {
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_scratch);
    h->d_scratch = nullptr;
    h->scratch_cap = 0;
    CK(cudaMalloc(&h->d_scratch, bytes));
    h->scratch_cap = bytes;
  }
  return SSB_OK;
}

int32_t ensure_scratch(ssb_analyzer* h, size_t bytes) {
  const int32_t drc = ensure_scratch_device(h, bytes);
  if (drc) return drc;
  if (bytes > h->h_scratch_cap) {
    if (h->h_scratch) cudaFreeHost(h->h_scratch);
    h->h_scratch = nullptr;
    CK(cudaMallocHost(&h->h_scratch, bytes));
    h->h_scratch_cap = bytes;
  }
  return SSB_OK;
}

// enter every completed-but-ungated bucket into the histograms (separate launch)
static int32_t flush_gating(ssb_analyzer* h) {
  const uint64_t done = h->total_frames / h->lp.s100;
  if (done > h->gated_upto) {
    CK(launch_gating(h->gp, h->st, h->gated_upto, done - 1, h->stream, &h->launches));
    h->gated_upto = done;
    h->lra_cache_valid = h->icache_valid = false;   // k_gating enters blocks / short-term energies behind the cache's back
  }
  return SSB_OK;
}

// What the results code of one launch may take from the per-stream cache (StreamCache), and the cache's state after it.
//  * loudness range: the short-term histogram only changes when a 3 s entry is gated (bucket j with j >= 29 and
//    (j - 29) % 10 == 0, once per second of audio), so a launch scans it only then, or when the cache is stale;
//  * integrated loudness: the lean path patches the cached gating sums with the pending blocks; it needs a current
//    cache, the bucket-based momentary / short-term windows (no ring), few pending buckets and few channels.  Every
//    full scan rebuilds the cache, so one non-lean launch makes the next ones lean again.
struct ResultsMode {
  int lra_from_cache, lean;
  int st_back;   // >= 0: the lean code gates the launch's 3 s entry itself and the loudness range comes from lra_scan_fast
  bool lra_valid_before, i_valid_before;
};
static ResultsMode results_mode_for_launch(ssb_analyzer* h, uint64_t gate_first, uint64_t gate_last) {
  ResultsMode m;
  m.lra_valid_before = h->lra_cache_valid;
  m.i_valid_before = h->icache_valid;
  bool st_entry = false;
  uint64_t j_st = 0;
  for (uint64_t j = gate_first; j <= gate_last && gate_last >= gate_first; j++)
    if (j >= 29 && (j - 29) % 10 == 0) { st_entry = true; j_st = j; break; }
  const uint64_t n_pending = gate_last >= gate_first ? gate_last - gate_first + 1 : 0;
  const bool want_i = (h->mode & SSB_MODE_I) == SSB_MODE_I, want_lra = (h->mode & SSB_MODE_LRA) == SSB_MODE_LRA;
  m.lra_from_cache = (h->lra_cache_valid && !st_entry) ? 1 : 0;
  m.lean = (h->lean_enabled && h->icache_valid && want_i && !h->st.ring && h->channels <= (uint32_t)kLeanMaxChannels &&
            n_pending <= (uint64_t)kLeanPending) ? 1 : 0;
  m.st_back = (m.lean && st_entry && want_lra && h->lra_cache_valid) ? (int)(gate_last - j_st) : -1;
  if (want_lra) h->lra_cache_valid = true;
  if (want_i) h->icache_valid = true;
  return m;
}
static void restore_results_mode(ssb_analyzer* h, const ResultsMode& m) {
  h->lra_cache_valid = m.lra_valid_before;
  h->icache_valid = m.i_valid_before;
}

// feed `frames` frames per stream from device memory laid out [stream][in_stride_frames][C]
int32_t feed_device(ssb_analyzer* h, const float* d_in, size_t frames, size_t in_stride_frames, double* d_results,
                    bool* written) {
  if (written) *written = false;
  if (!h->meter_ok) return fail(h, SSB_ERR_NOMEM, "the loudness meter is not initialised (a previous create/reinit failed)");
  const uint32_t s100 = h->lp.s100;
  const size_t C = h->channels;
  size_t done = 0;
  h->results_valid = false;
  h->dres_valid = false;
  while (done < frames) {
    const uint32_t pos = (uint32_t)(h->total_frames % s100);
    const uint64_t bucket0 = h->total_frames / s100;
    const size_t max_frames = (size_t)kMaxBucketsPerLaunch * s100 - pos;
    const size_t n = frames - done < max_frames ? frames - done : max_frames;
    // Gating is lazy: completed buckets wait for the next query (k_results gates them itself) unless this launch
    // would overwrite a ring slot that a pending bucket's 3 s window still needs (64 slots, windows reach back 29)
    if ((h->total_frames + n) / s100 > h->gated_upto + (uint64_t)(kNB - 30)) {
      const int32_t frc = flush_gating(h);
      if (frc) return frc;
    }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (h->profiling) {
      if (h->prof_used == h->prof_events.size()) {
        CK(cudaEventCreate(&ev0));
        CK(cudaEventCreate(&ev1));
        h->prof_events.emplace_back(ev0, ev1);
      }
      ev0 = h->prof_events[h->prof_used].first;
      ev1 = h->prof_events[h->prof_used].second;
      h->prof_used++;
      CK(cudaEventRecord(ev0, h->stream));
    }
    size_t tiled = 0;
    const int fk = h->force_kernel;
    const bool last_chunk = done + n == frames;
    // the second-generation batch kernel: mono / stereo, below the stream count where the serial kernel takes over
    const int wvariant = fk == 6 ? 1 : 0;
    const bool want_wtile = (fk == 5 || fk == 6 || (fk == 0 && h->st.n_streams < serial_min_streams()));
    if (h->d_scan_powers && (fk == 0 || fk == 4) && scan_path_usable(h->lp, h->st, n)) {
      CK(launch_loudness_scan(h->lp, h->st, h->d_scan_powers, d_in + done * C, n, in_stride_frames, pos, bucket0,
                              h->ring_pos, h->stream, &h->launches));
      tiled = n;
    } else if (want_wtile && wtile_path_usable(h->lp, h->st, d_in + done * C, n, in_stride_frames, wvariant)) {
      // results fused into the filter launch when the caller wants them and this launch ends the feed on whole tiles
      ResultsArgs ra{};
      const ResultsArgs* rap = nullptr;
      ResultsMode rm{};
      const uint64_t frames_after = h->total_frames + n;
      const uint64_t done_after = frames_after / s100;
      if (d_results && last_chunk) {
        const bool pending = done_after > h->gated_upto;
        ra = make_results_args(h->st, done_after, (frames_after % s100) == 0, h->ring_pos, h->mode, d_results,
                               pending ? h->gated_upto : 1, pending ? done_after - 1 : 0, nullptr);
        ra.ga = peek_gather_args(h);   // rows also go to every rank's gather buffer when one is open
        rap = &ra;
        rm = results_mode_for_launch(h, pending ? h->gated_upto : 1, pending ? done_after - 1 : 0);
        ra.lra_from_cache = rm.lra_from_cache;
        ra.lean = rm.lean;
        ra.st_back = rm.st_back;
        ra.lra_fast = rm.st_back >= 0 ? 1 : 0;
      }
      bool wrote = false;
      // (if the launch turns out not to write the rows, the cache flag goes back to what it was)
      CK(launch_loudness_wtile(h->lp, h->st, h->gp, d_in + done * C, n, in_stride_frames, pos, bucket0, wvariant, rap,
                               h->sm_count, h->device, h->stream, &h->launches, &tiled, &wrote));
      if (wrote) {
        if (done_after > h->gated_upto) h->gated_upto = done_after;
        if (written) *written = true;
      } else if (rap) {
        restore_results_mode(h, rm);
      }
    } else if (fk != 1 && fk != 4 && fk != 5 && fk != 6 && tile_path_usable(h->lp, h->st, d_in + done * C, n, in_stride_frames))
      CK(launch_loudness_tile(h->lp, h->st, d_in + done * C, n, in_stride_frames, pos, bucket0, h->stream,
                              &h->launches, &tiled, fk, h->sm_count));
    if (tiled < n) {
      const uint64_t t2 = h->total_frames + tiled;
      CK(launch_loudness_generic(h->lp, h->st, d_in + (done + tiled) * C, n - tiled, in_stride_frames,
                                 (uint32_t)(t2 % s100), t2 / s100, h->ring_pos, h->stream, &h->launches));
    }
    if (ev1) CK(cudaEventRecord(ev1, h->stream));
    h->total_frames += n;
    if (h->st.ring_frames) h->ring_pos = (h->ring_pos + n) % h->st.ring_frames;
    done += n;
  }
  return SSB_OK;
}

int32_t launch_results_now(ssb_analyzer* h) {
  const int aligned = (h->total_frames % h->lp.s100) == 0;
  const uint64_t done = h->total_frames / h->lp.s100;
  const GatherArgs ga = peek_gather_args(h);
  const uint64_t gf = done > h->gated_upto ? h->gated_upto : 1, gl = done > h->gated_upto ? done - 1 : 0;
  const ResultsMode rm = results_mode_for_launch(h, gf, gl);
  CK(launch_results(h->gp, h->st, done, aligned, h->ring_pos, h->mode, h->d_results, h->stream, &h->launches, gf, gl,
                    ga.world ? &ga : nullptr, rm.lra_from_cache, rm.lean, rm.st_back));
  if (done > h->gated_upto) h->gated_upto = done;
  return SSB_OK;
}

const std::pair<std::vector<double>, std::vector<double>>& fft_axis_cached(ssb_analyzer* h, size_t n, uint32_t rate) {
  auto key = std::make_pair(n, rate);
  auto it = h->axes.find(key);
  if (it == h->axes.end()) {
    std::vector<double> x, t;
    fft_axis(n, rate, x, t, nullptr);
    it = h->axes.emplace(key, std::make_pair(std::move(x), std::move(t))).first;
  }
  return it->second;
}

static int32_t refresh_results(ssb_analyzer* h) {
  if (!h->meter_ok) return fail(h, SSB_ERR_NOMEM, "the loudness meter is not initialised (a previous create/reinit failed)");
  if (h->results_valid) return SSB_OK;
  if (!h->dres_valid) {
    const int32_t rrc = launch_results_now(h);
    if (rrc) return rrc;
    h->dres_valid = true;
  }
  const size_t stride = 4 + 2 * (size_t)h->channels;
  CK(cudaMemcpyAsync(h->h_results, h->d_results, h->n_streams * stride * sizeof(double), cudaMemcpyDeviceToHost,
                     h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->results_valid = true;
  return SSB_OK;
}

int32_t get_plan(ssb_analyzer* h, size_t n, uint32_t rate, FftPlan** out) {
  auto key = std::make_pair(n, rate);
  auto it = h->plans.find(key);
  if (it == h->plans.end()) {
    FftPlan p;
    p.n = n;
    p.rate = rate;
    p.n_bins = fft_bin_range(n, rate, &p.k_first);
    std::vector<float> w;
    hann_multipliers(n, w);
    std::vector<float2> tw(n / 2 ? n / 2 : 1);
    for (size_t k = 0; k < n / 2; k++) {
      // exp(-j*2*pi*k/n), exact on the axes, rounded once from double elsewhere
      if (k == 0) tw[k] = make_float2(1.f, 0.f);
      else if (4 * k == n) tw[k] = make_float2(0.f, -1.f);
      else {
        const double ang = 2.0 * M_PI * (double)k / (double)n;
        tw[k] = make_float2((float)cos(ang), (float)-sin(ang));
      }
    }
    CK(cudaMalloc(&p.d_window, n * sizeof(float)));
    CK(cudaMalloc(&p.d_twiddle, tw.size() * sizeof(float2)));
    CK(cudaMemcpyAsync(p.d_window, w.data(), n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(p.d_twiddle, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
    std::vector<float2> lo(64), hi(n >= 64 ? n / 64 : 0);
    auto wn = [&](size_t e) {
      const size_t m = e % n;
      if (m == 0) return make_float2(1.f, 0.f);
      if (4 * m == n) return make_float2(0.f, -1.f);
      if (2 * m == n) return make_float2(-1.f, 0.f);
      if (4 * m == 3 * n) return make_float2(0.f, 1.f);
      const double ang = 2.0 * M_PI * (double)m / (double)n;
      return make_float2((float)cos(ang), (float)-sin(ang));
    };
    if (n >= 1024) {  // the fast kernel serves transform lengths >= 512
      for (size_t i = 0; i < 64; i++) lo[i] = wn(i);
      for (size_t i = 0; i < hi.size(); i++) hi[i] = wn(64 * i);
      CK(cudaMalloc(&p.d_tw_lo, lo.size() * sizeof(float2)));
      CK(cudaMalloc(&p.d_tw_hi, hi.size() * sizeof(float2)));
      CK(cudaMemcpyAsync(p.d_tw_lo, lo.data(), lo.size() * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
      CK(cudaMemcpyAsync(p.d_tw_hi, hi.data(), hi.size() * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
    }
    {
      const std::vector<double>& tilt = fft_axis_cached(h, n, rate).second;   // per kept bin, the host's f64 values
      if (!tilt.empty()) {
        CK(cudaMalloc(&p.d_tilt, tilt.size() * sizeof(double)));
        CK(cudaMemcpyAsync(p.d_tilt, tilt.data(), tilt.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      }
    }
    CK(cudaStreamSynchronize(h->stream));  // w / tw / lo / hi are locals
    it = h->plans.emplace(key, p).first;
  }
  *out = &it->second;
  return SSB_OK;
}

// spectrum-analyzer's argument checks that do not need the data
int32_t fft_shape_check(size_t n, uint32_t rate) {
  if (n < 2) return SSB_ERR_FFT_TOO_FEW_SAMPLES;
  if ((n & (n - 1)) || n > 32768) return SSB_ERR_FFT_NOT_POW2;
  if (20000.0f > (float)rate / 2.0f) return SSB_ERR_FFT_BAD_LIMIT;
  return SSB_OK;
}

size_t waveform_window_columns(double waveform_window, size_t len, size_t* window_out) {
  const double w = waveform_window * 1000.;
  size_t window = 0;  // Rust `as usize` saturates; NaN -> 0
  if (w > 0.0) window = w >= 18446744073709551615.0 ? SIZE_MAX : (size_t)w;
  *window_out = window;
  if (!window || !len) return 0;
  const double spp = (double)len / (double)window;
  // the loop breaks at the first column whose start is past the end (analyzer.rs:122-124); starts are monotone
  size_t cols = window;
  while (cols > 0 && (size_t)((double)(cols - 1) * spp) >= len) cols--;
  return cols;
}

}  // namespace ssb

extern "C" {

uint32_t ssb_abi_version(void) { return SSB_ABI_VERSION; }

int32_t ssb_analyzer_create(ssb_analyzer** out, uint32_t channels, uint32_t rate, int32_t mode, size_t n_streams,
                            int32_t device, uint32_t flags) {
  if (!out) return SSB_ERR_INVALID_ARG;
  *out = nullptr;
  if (n_streams == 0) return SSB_ERR_INVALID_ARG;
  if ((mode & SSB_MODE_M) == 0 || (mode & ~SSB_MODE_ALL)) return SSB_ERR_INVALID_MODE;
  // integrated loudness / LRA are implemented in ebur128's histogram mode, the one Mode::all() selects
  if (((mode & SSB_MODE_I) == SSB_MODE_I || (mode & SSB_MODE_LRA) == SSB_MODE_LRA) && !(mode & SSB_MODE_HISTOGRAM))
    return SSB_ERR_INVALID_MODE;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return SSB_ERR_NO_DEVICE;
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) return SSB_ERR_NO_DEVICE;
  }
  if (device >= count) return SSB_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SSB_ERR_NO_DEVICE;
  if (prop.major != 10) return SSB_ERR_NO_DEVICE;  // kernels are built for sm_100a only
  ssb_analyzer* h = new (std::nothrow) ssb_analyzer();
  if (!h) return SSB_ERR_NOMEM;
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->mode = mode;
  h->n_streams = n_streams;
  h->flags = flags;
  DeviceGuard g(device);
  int32_t rc = SSB_OK;
  do {
    if ((e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking))) { rc = SSB_ERR_CUDA + e; break; }
    h->stream = h->own_stream;
    for (int i = 0; i < 2; i++) {
      if ((e = cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming))) { rc = SSB_ERR_CUDA + e; break; }
      if ((e = cudaEventCreateWithFlags(&h->ev_consumed[i], cudaEventDisableTiming))) { rc = SSB_ERR_CUDA + e; break; }
    }
    if (rc) break;
    double tables[2001];
    histogram_tables(tables, tables + 1000);
    if ((e = cudaMalloc(&h->d_hist_tables, sizeof(tables)))) { rc = SSB_ERR_CUDA + e; break; }
    if ((e = cudaMemcpy(h->d_hist_tables, tables, sizeof(tables), cudaMemcpyHostToDevice))) { rc = SSB_ERR_CUDA + e; break; }
    rc = init_meter(h, channels, rate);
  } while (0);
  if (rc != SSB_OK) {
    ssb_analyzer_destroy(h);
    return rc;
  }
  *out = h;
  return SSB_OK;
}

void ssb_analyzer_destroy(ssb_analyzer* h) {
  if (!h) return;
  DeviceGuard g(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->oneshot) ssb_analyzer_destroy(h->oneshot);
  h->oneshot = nullptr;
  ssb_gather_destroy(h);
  free_meter(h);
  cudaFree(h->d_hist_tables);
  for (int i = 0; i < 2; i++) {
    cudaFree(h->d_stage[i]);
    if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]);
    if (h->ev_consumed[i]) cudaEventDestroy(h->ev_consumed[i]);
  }
  cudaFree(h->d_scratch);
  cudaFree(h->d_pcm);
  if (h->h_scratch) cudaFreeHost(h->h_scratch);
  for (auto& ev : h->prof_events) {
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  for (auto& kv : h->plans) {
    cudaFree(kv.second.d_window);
    cudaFree(kv.second.d_twiddle);
    cudaFree(kv.second.d_tw_lo);
    cudaFree(kv.second.d_tw_hi);
    cudaFree(kv.second.d_tilt);
  }
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
}

int32_t ssb_create_loudness_meter(ssb_analyzer* h, uint32_t channels, uint32_t rate) {
  if (!h) return SSB_ERR_INVALID_ARG;
  DeviceGuard g(h->device);
  // the reference stores the rate before EbuR128::new can fail (analyzer.rs:50-51)
  const uint32_t old_rate = h->rate;
  cudaStreamSynchronize(h->stream);
  int32_t rc = init_meter(h, channels, rate);
  if (rc == SSB_ERR_NOMEM) { h->rate = rate; (void)old_rate; }
  return rc;
}

uint32_t ssb_sample_rate(const ssb_analyzer* h) { return h ? h->rate : 0; }
uint32_t ssb_channels(const ssb_analyzer* h) { return h ? h->channels : 0; }
size_t ssb_n_streams(const ssb_analyzer* h) { return h ? h->n_streams : 0; }
const char* ssb_last_error(const ssb_analyzer* h) { return h ? h->err : "null handle"; }
uint64_t ssb_launch_count(const ssb_analyzer* h) { return h ? h->launches : 0; }

int32_t ssb_set_stream(ssb_analyzer* h, void* cuda_stream) {
  if (!h) return SSB_ERR_INVALID_ARG;
  DeviceGuard g(h->device);
  CK(cudaStreamSynchronize(h->stream));
  h->stream = (cudaStream_t)cuda_stream;
  return SSB_OK;
}

int32_t ssb_use_own_stream(ssb_analyzer* h) {
  if (!h) return SSB_ERR_INVALID_ARG;
  DeviceGuard g(h->device);
  CK(cudaStreamSynchronize(h->stream));
  h->stream = h->own_stream;
  return SSB_OK;
}

int32_t ssb_sync(ssb_analyzer* h) {
  if (!h) return SSB_ERR_INVALID_ARG;
  DeviceGuard g(h->device);
  CK(cudaStreamSynchronize(h->stream));
  return SSB_OK;
}

int32_t ssb_add_frames_f32_device(ssb_analyzer* h, const float* d_interleaved, size_t frames_per_stream) {
  if (!h) return SSB_ERR_INVALID_ARG;
  if (frames_per_stream == 0) return SSB_OK;
  if (!d_interleaved) return fail(h, SSB_ERR_INVALID_ARG, "null input");
  DeviceGuard g(h->device);
  return feed_device(h, d_interleaved, frames_per_stream, frames_per_stream);
}

int32_t ssb_add_frames_f32_device_results(ssb_analyzer* h, const float* d_interleaved, size_t frames_per_stream,
                                          double* d_out) {
  if (!h || !d_out) return SSB_ERR_INVALID_ARG;
  if (frames_per_stream && !d_interleaved) return fail(h, SSB_ERR_INVALID_ARG, "null input");
  DeviceGuard g(h->device);
  bool wrote = false;
  if (frames_per_stream) {
    const int32_t rc = feed_device(h, d_interleaved, frames_per_stream, frames_per_stream, d_out, &wrote);
    if (rc) return rc;
  }
  return wrote ? SSB_OK : ssb_results_device(h, d_out);
}

int32_t ssb_add_frames_f32(ssb_analyzer* h, const float* interleaved, size_t frames_per_stream) {
  if (!h) return SSB_ERR_INVALID_ARG;
  if (frames_per_stream == 0) return SSB_OK;
  if (!interleaved) return fail(h, SSB_ERR_INVALID_ARG, "null input");
  DeviceGuard g(h->device);
  const size_t floats = h->n_streams * frames_per_stream * h->channels;
  int32_t rc = ensure_stage(h, floats);
  if (rc) return rc;
  const int i = h->stage_idx;
  h->stage_idx ^= 1;
  // the staging buffer may still be read by the kernels of two calls ago
  CK(cudaEventSynchronize(h->ev_consumed[i]));
  CK(cudaMemcpyAsync(h->d_stage[i], interleaved, floats * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  CK(cudaEventRecord(h->ev_copied[i], h->stream));
  {
    // the rows land in the handle's own result buffer when the filter launch can write them itself, so the query that
    // usually follows (get_shortterm_lufs after add_samples, tui.rs:1539-1543) is a copy, not a launch
    bool wrote = false;
    rc = feed_device(h, h->d_stage[i], frames_per_stream, frames_per_stream, h->d_results, &wrote);
    if (rc) return rc;
    h->dres_valid = wrote;
  }
  CK(cudaEventRecord(h->ev_consumed[i], h->stream));
  CK(cudaEventSynchronize(h->ev_copied[i]));  // caller may reuse its buffer; kernels keep running
  return SSB_OK;
}

int32_t ssb_add_samples(ssb_analyzer* h, const float* interleaved, size_t len) {
  if (!h) return SSB_ERR_INVALID_ARG;
  if (len % h->channels != 0)
    return fail(h, SSB_ERR_NOMEM, "add_frames_f32: %zu samples is not a multiple of %u channels (Error::NoMem)", len,
                h->channels);
  return ssb_add_frames_f32(h, interleaved, len / h->channels);
}

int32_t ssb_reset(ssb_analyzer* h) {
  if (!h) return SSB_ERR_INVALID_ARG;
  if (!h->meter_ok) return fail(h, SSB_ERR_NOMEM, "the loudness meter is not initialised");
  DeviceGuard g(h->device);
  CK(launch_reset(h->st, (int)h->channels, h->stream, &h->launches));
  h->total_frames = 0;
  h->gated_upto = 0;
  h->ring_pos = 0;
  h->results_valid = false;
  h->dres_valid = false;
  h->lra_cache_valid = h->icache_valid = true;   // zeroed by launch_reset
  return SSB_OK;
}

static int32_t copy_column(ssb_analyzer* h, int col, double* out) {
  const size_t stride = 4 + 2 * (size_t)h->channels;
  for (size_t s = 0; s < h->n_streams; s++) out[s] = h->h_results[s * stride + col];
  return SSB_OK;
}

int32_t ssb_loudness_momentary(ssb_analyzer* h, double* out) {
  if (!h || !out) return SSB_ERR_INVALID_ARG;
  DeviceGuard g(h->device);
  if (!h->st.ring && (h->total_frames % h->lp.s100) != 0)
    return fail(h, SSB_ERR_UNALIGNED_QUERY, "momentary query %llu frames into a 100 ms block needs SSB_FLAG_RING",
                (unsigned long long)(h->total_frames % h->lp.s100));
  int32_t rc = refresh_results(h);
  return rc ? rc : copy_column(h, 0, out);
}

int32_t ssb_loudness_shortterm(ssb_analyzer* h, double* out) {
  if (!h || !out) return SSB_ERR_INVALID_ARG;
  if ((h->mode & SSB_MODE_S) != SSB_MODE_S) return fail(h, SSB_ERR_INVALID_MODE, "mode lacks S (Error::InvalidMode)");
  DeviceGuard g(h->device);
  if (!h->st.ring && (h->total_frames % h->lp.s100) != 0)
    return fail(h, SSB_ERR_UNALIGNED_QUERY, "short-term query %llu frames into a 100 ms block needs SSB_FLAG_RING",
                (unsigned long long)(h->total_frames % h->lp.s100));
  int32_t rc = refresh_results(h);
  return rc ? rc : copy_column(h, 1, out);
}

int32_t ssb_loudness_global(ssb_analyzer* h, double* out) {
  if (!h || !out) return SSB_ERR_INVALID_ARG;
  if ((h->mode & SSB_MODE_I) != SSB_MODE_I) return fail(h, SSB_ERR_INVALID_MODE, "mode lacks I (Error::InvalidMode)");
  DeviceGuard g(h->device);
  int32_t rc = refresh_results(h);
  return rc ? rc : copy_column(h, 2, out);
}

int32_t ssb_loudness_range(ssb_analyzer* h, double* out) {
  if (!h || !out) return SSB_ERR_INVALID_ARG;
  if ((h->mode & SSB_MODE_LRA) != SSB_MODE_LRA) return fail(h, SSB_ERR_INVALID_MODE, "mode lacks LRA (Error::InvalidMode)");
  DeviceGuard g(h->device);
  int32_t rc = refresh_results(h);
  return rc ? rc : copy_column(h, 3, out);
}

int32_t ssb_true_peak(ssb_analyzer* h, double* out) {
  if (!h || !out) return SSB_ERR_INVALID_ARG;
  if ((h->mode & SSB_MODE_TRUE_PEAK) != SSB_MODE_TRUE_PEAK)
    return fail(h, SSB_ERR_INVALID_MODE, "mode lacks TRUE_PEAK (Error::InvalidMode)");
  DeviceGuard g(h->device);
  int32_t rc = refresh_results(h);
  if (rc) return rc;
  const size_t C = h->channels, stride = 4 + 2 * C;
  for (size_t s = 0; s < h->n_streams; s++)
    for (size_t c = 0; c < C; c++) out[s * C + c] = h->h_results[s * stride + 4 + c];
  return SSB_OK;
}

int32_t ssb_sample_peak(ssb_analyzer* h, double* out) {
  if (!h || !out) return SSB_ERR_INVALID_ARG;
  if ((h->mode & SSB_MODE_SAMPLE_PEAK) != SSB_MODE_SAMPLE_PEAK)
    return fail(h, SSB_ERR_INVALID_MODE, "mode lacks SAMPLE_PEAK (Error::InvalidMode)");
  DeviceGuard g(h->device);
  int32_t rc = refresh_results(h);
  if (rc) return rc;
  const size_t C = h->channels, stride = 4 + 2 * C;
  for (size_t s = 0; s < h->n_streams; s++)
    for (size_t c = 0; c < C; c++) out[s * C + c] = h->h_results[s * stride + 4 + C + c];
  return SSB_OK;
}

int32_t ssb_get_true_peak(ssb_analyzer* h, double* left, double* right) {
  if (!h || !left || !right) return SSB_ERR_INVALID_ARG;
  if ((h->mode & SSB_MODE_TRUE_PEAK) != SSB_MODE_TRUE_PEAK)
    return fail(h, SSB_ERR_INVALID_MODE, "mode lacks TRUE_PEAK (Error::InvalidMode)");
  if (h->channels < 2) return fail(h, SSB_ERR_INVALID_CHANNEL_INDEX, "true_peak(1) on a mono meter (Error::InvalidChannelIndex)");
  DeviceGuard g(h->device);
  int32_t rc = refresh_results(h);
  if (rc) return rc;
  *left = h->h_results[4];
  *right = h->h_results[5];
  return SSB_OK;
}

size_t ssb_result_stride(const ssb_analyzer* h) { return h ? 4 + 2 * (size_t)h->channels : 0; }

int32_t ssb_results_device(ssb_analyzer* h, double* d_out) {
  if (!h || !d_out) return SSB_ERR_INVALID_ARG;
  if (!h->meter_ok) return fail(h, SSB_ERR_NOMEM, "the loudness meter is not initialised");
  DeviceGuard g(h->device);
  const int aligned = (h->total_frames % h->lp.s100) == 0;
  {
    const uint64_t done = h->total_frames / h->lp.s100;
    const GatherArgs ga = peek_gather_args(h);
    const uint64_t gf = done > h->gated_upto ? h->gated_upto : 1, gl = done > h->gated_upto ? done - 1 : 0;
    const ResultsMode rm = results_mode_for_launch(h, gf, gl);
    CK(launch_results(h->gp, h->st, done, aligned, h->ring_pos, h->mode, d_out, h->stream, &h->launches, gf, gl,
                      ga.world ? &ga : nullptr, rm.lra_from_cache, rm.lean, rm.st_back));
      if (done > h->gated_upto) h->gated_upto = done;
  }
  return SSB_OK;
}

// `d_resident`: the same samples already on the device (complete, any stream), or nullptr to copy them from `samples`
static int32_t one_shot_integrated(ssb_analyzer* h, uint32_t channels, const float* samples, const float* d_resident,
                                   size_t len, double* out, int32_t* is_some) {
  if (!h || !out || !is_some || (!samples && len)) return SSB_ERR_INVALID_ARG;
  *is_some = 0;
  DeviceGuard g(h->device);
  // The reference builds a fresh EbuR128 per call; a cached meter that is reset gives the same result without
  // paying device allocations every time.
  ssb_analyzer* tmp = h->oneshot;
  int32_t rc = SSB_OK;
  if (tmp && (tmp->channels != channels || tmp->rate != h->rate)) {
    ssb_analyzer_destroy(tmp);
    tmp = h->oneshot = nullptr;
  }
  if (!tmp) {
    // the reference builds the meter with Mode::all() (analyzer.rs:171) but reads only loudness_global(): the peak
    // detectors' output is dropped with the meter, so this one computes K-weighting + gating only — same value
    rc = ssb_analyzer_create(&tmp, channels, h->rate, SSB_MODE_I | SSB_MODE_HISTOGRAM, 1, h->device, 0);
    if (rc == SSB_ERR_NOMEM) return SSB_OK;  // EbuR128::new failed -> None
    if (rc) return fail(h, rc, "calculate_integrated_lufs: cannot create meter");
    h->oneshot = tmp;
  } else {
    rc = ssb_reset(tmp);
    if (rc) return fail(h, rc, "calculate_integrated_lufs: reset failed");
  }
  const uint64_t launches0 = tmp->launches;
  // analyzer.rs:175: chunks(sample_rate * 2); a chunk that is not whole frames makes add_frames_f32 fail -> None
  const size_t chunk = (size_t)h->rate * 2;
  bool ok = true;
  for (size_t off = 0; off < len && ok; off += chunk) {
    const size_t n = len - off < chunk ? len - off : chunk;
    if (n % channels != 0) ok = false;
  }
  if (ok && len) {
    const size_t frames = len / channels;
    const uint32_t s100 = tmp->lp.s100;
    const uint64_t n_buckets = frames / s100;             // complete 100 ms buckets
    // mono / stereo files of at least a second: time-chunked scan, every chunk on its own SM (loudness_scan.cu)
    const bool file_path = tmp->d_scan_powers && tmp->force_kernel == 0 && h->force_kernel == 0 && n_buckets >= 10 &&
                           scan_path_usable(tmp->lp, tmp->st, frames);
    const size_t in_bytes = (len * sizeof(float) + 255) & ~(size_t)255;
    const size_t stride = (size_t)n_buckets + 1;
    const size_t own_in = d_resident ? 0 : in_bytes;
    rc = ensure_scratch_device(tmp, own_in + (file_path ? (size_t)channels * stride * sizeof(double) : 0) + 256);
    if (rc) return fail(h, rc, "calculate_integrated_lufs: %s", tmp->err);
    const float* d = d_resident;
    cudaError_t e = cudaSuccess;
    if (!d_resident) {
      e = cudaMemcpyAsync(tmp->d_scratch, samples, len * sizeof(float), cudaMemcpyHostToDevice, tmp->stream);
      d = tmp->d_scratch;
    }
    if (e) return cuda_fail(h, e, "calculate_integrated_lufs");
    if (file_path) {
      double* d_fb = reinterpret_cast<double*>(reinterpret_cast<char*>(tmp->d_scratch) + own_in);
      // 1 s chunks (+0.4 s run-in) until the file outgrows four chunks per SM, then longer ones
      size_t chunk_buckets = 10;
      const size_t max_chunks = 4 * (size_t)(tmp->sm_count > 0 ? tmp->sm_count : 148);
      if ((n_buckets + chunk_buckets - 1) / chunk_buckets > max_chunks) chunk_buckets = (n_buckets + max_chunks - 1) / max_chunks;
      e = launch_loudness_scan_file(tmp->lp, tmp->st, tmp->d_scan_powers, d, frames, d_fb, stride, chunk_buckets,
                                    tmp->stream, &tmp->launches);
      if (!e) e = launch_file_gating(tmp->gp, tmp->st, d_fb, stride, n_buckets, tmp->stream, &tmp->launches);
      if (e) return cuda_fail(h, e, "calculate_integrated_lufs (file path)");
      tmp->results_valid = false;   // total_frames stays 0: nothing is pending for the streaming gating
      tmp->lra_cache_valid = tmp->icache_valid = false;
    } else {
      for (size_t off = 0; off < len && !rc; off += chunk) {
        const size_t n = len - off < chunk ? len - off : chunk;
        rc = feed_device(tmp, d + off, n / channels, n / channels);
      }
    }
  }
  if (!rc && ok) {
    double v = 0;
    rc = ssb_loudness_global(tmp, &v);
    if (!rc) { *out = v; *is_some = 1; }
  }
  h->launches += tmp->launches - launches0;
  return rc;
}

int32_t ssb_calculate_integrated_lufs(ssb_analyzer* h, uint32_t channels, const float* samples, size_t len,
                                      double* out, int32_t* is_some) {
  return one_shot_integrated(h, channels, samples, nullptr, len, out, is_some);
}

int32_t ssb_fft_bins(size_t n, uint32_t rate, size_t* k_first, size_t* n_bins) {
  if (!n_bins) return SSB_ERR_INVALID_ARG;
  int32_t rc = fft_shape_check(n, rate);
  if (rc) return rc;
  size_t k0 = 0;
  *n_bins = fft_bin_range(n, rate, &k0);
  if (k_first) *k_first = k0;
  return SSB_OK;
}

int32_t ssb_fft_axis(size_t n, uint32_t rate, double* x_out, double* tilt_out, size_t cap, size_t* n_bins) {
  if (!n_bins) return SSB_ERR_INVALID_ARG;
  int32_t rc = fft_shape_check(n, rate);
  if (rc) return rc;
  std::vector<double> x, t;
  fft_axis(n, rate, x, t, nullptr);
  *n_bins = x.size();
  if (x.size() > cap) return SSB_ERR_CAPACITY;
  if (x_out) memcpy(x_out, x.data(), x.size() * sizeof(double));
  if (tilt_out) memcpy(tilt_out, t.data(), t.size() * sizeof(double));
  return SSB_OK;
}

int32_t ssb_fft_batch_device(ssb_analyzer* h, const float* d_in, int32_t layout, size_t n, size_t n_windows,
                             float* d_db_out, int32_t* d_status) {
  if (!h || !d_in || !d_db_out) return SSB_ERR_INVALID_ARG;
  if (layout != SSB_FFT_MONO && layout != SSB_FFT_MID_SIDE) return fail(h, SSB_ERR_INVALID_ARG, "bad layout");
  // the kernels read the windows with 8- and 16-byte vector loads: a misaligned device pointer (an odd-offset view of a
  // larger buffer) would fault and poison the context, so it is refused here
  if (((uintptr_t)d_in & 15) != 0 || ((uintptr_t)d_db_out & 3) != 0)
    return fail(h, SSB_ERR_INVALID_ARG, "ssb_fft_batch_device: d_in must be 16-byte aligned (got %p)", (const void*)d_in);
  int32_t rc = fft_shape_check(n, h->rate);
  if (rc) return fail(h, rc, "get_fft: invalid length %zu / rate %u", n, h->rate);
  DeviceGuard g(h->device);
  FftPlan* plan = nullptr;
  rc = get_plan(h, n, h->rate, &plan);
  if (rc) return rc;
  CK(launch_fft(*plan, d_in, layout, n_windows, d_db_out, d_status, h->stream, &h->launches));
  return SSB_OK;
}

int32_t ssb_fft_batch_device_y(ssb_analyzer* h, const float* d_in, int32_t layout, size_t n, size_t n_windows,
                               double* d_y_out, int32_t* d_status) {
  if (!h || !d_in || !d_y_out) return SSB_ERR_INVALID_ARG;
  if (layout != SSB_FFT_MONO && layout != SSB_FFT_MID_SIDE) return fail(h, SSB_ERR_INVALID_ARG, "bad layout");
  if (((uintptr_t)d_in & 15) != 0 || ((uintptr_t)d_y_out & 7) != 0)
    return fail(h, SSB_ERR_INVALID_ARG, "ssb_fft_batch_device_y: d_in must be 16-byte aligned (got %p)", (const void*)d_in);
  int32_t rc = fft_shape_check(n, h->rate);
  if (rc) return fail(h, rc, "get_fft: invalid length %zu / rate %u", n, h->rate);
  DeviceGuard g(h->device);
  FftPlan* plan = nullptr;
  rc = get_plan(h, n, h->rate, &plan);
  if (rc) return rc;
  CK(launch_fft_y(*plan, d_in, layout, n_windows, d_y_out, d_status, h->stream, &h->launches));
  return SSB_OK;
}

int32_t ssb_get_fft(ssb_analyzer* h, const float* samples, size_t n, double* xy_out, size_t cap, size_t* n_points) {
  if (!h || !n_points || (!samples && n)) return SSB_ERR_INVALID_ARG;
  *n_points = 0;
  int32_t shape = fft_shape_check(n, h->rate);
  if (shape == SSB_ERR_FFT_TOO_FEW_SAMPLES) return fail(h, shape, "get_fft: too few samples (%zu)", n);
  if (shape) {
    // the crate checks NaN / infinity of the WINDOWED samples before the length and the limit; the Hann
    // multiplier is finite and only multiplier[0] is zero (0 * inf = NaN), so inspect the input directly
    bool any_nan = false, any_inf = false;
    for (size_t i = 0; i < n; i++) {
      if (isnan(samples[i]) || (i == 0 && isinf(samples[i]))) any_nan = true;
      else if (isinf(samples[i])) any_inf = true;
    }
    if (any_nan) return fail(h, SSB_ERR_FFT_NAN, "get_fft: NaN values not supported");
    if (any_inf) return fail(h, SSB_ERR_FFT_INF, "get_fft: infinity values not supported");
    return fail(h, shape, shape == SSB_ERR_FFT_NOT_POW2 ? "get_fft: length %zu is not a supported power of two"
                                                          : "get_fft: 20 kHz limit above Nyquist (n=%zu)", n);
  }
  DeviceGuard g(h->device);
  FftPlan* plan = nullptr;
  int32_t rc = get_plan(h, n, h->rate, &plan);
  if (rc) return rc;
  const size_t nb = plan->n_bins;
  *n_points = nb;
  if (nb > cap || !xy_out) return fail(h, SSB_ERR_CAPACITY, "get_fft: need room for %zu points", nb);
  const size_t in_bytes = n * sizeof(float), out_bytes = nb * sizeof(float) + sizeof(int32_t);
  const size_t out_off = (in_bytes + 255) & ~(size_t)255;
  rc = ensure_scratch(h, out_off + out_bytes + 256);
  if (rc) return rc;
  float* d_in = h->d_scratch;
  float* d_db = reinterpret_cast<float*>(reinterpret_cast<char*>(h->d_scratch) + out_off);
  int32_t* d_status = reinterpret_cast<int32_t*>(d_db + nb);
  CK(cudaMemcpyAsync(d_in, samples, in_bytes, cudaMemcpyHostToDevice, h->stream));
  CK(launch_fft(*plan, d_in, SSB_FFT_MONO, 1, d_db, d_status, h->stream, &h->launches));
  CK(cudaMemcpyAsync(h->h_scratch, d_db, out_bytes, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const float* db = static_cast<const float*>(h->h_scratch);
  const int32_t status = *reinterpret_cast<const int32_t*>(db + nb);
  if (status) return fail(h, status, "get_fft: window rejected (%d)", status);
  const auto& axes = fft_axis_cached(h, n, h->rate);
  const std::vector<double>& ax = axes.first;
  const std::vector<double>& tilt = axes.second;
  for (size_t i = 0; i < nb; i++) {
    xy_out[2 * i] = ax[i];
    xy_out[2 * i + 1] = (double)db[i] + tilt[i];  // analyzer.rs:82-84: val as f64 + compensation
  }
  return SSB_OK;
}

int32_t ssb_process_tick(ssb_analyzer* h, const float* tail, size_t n_fft, size_t lufs_samples, double* xy_mid,
                         double* xy_side, size_t cap, size_t* n_points, double* shortterm_lufs, int32_t* fft_status,
                         int32_t* lufs_status) {
  if (!h || !tail || !n_points || !shortterm_lufs || !fft_status || !lufs_status) return SSB_ERR_INVALID_ARG;
  if (h->channels != 2 || h->n_streams != 1) return fail(h, SSB_ERR_INVALID_ARG, "process_tick needs one stereo stream");
  if (!h->meter_ok) return fail(h, SSB_ERR_NOMEM, "the loudness meter is not initialised (a previous create/reinit failed)");
  if (lufs_samples > 2 * n_fft) return fail(h, SSB_ERR_INVALID_ARG, "lufs_samples exceeds the tail");
  *n_points = 0;
  *fft_status = fft_shape_check(n_fft, h->rate);
  h->tick_fft_status[0] = h->tick_fft_status[1] = *fft_status;
  *lufs_status = (lufs_samples % 2 != 0) ? SSB_ERR_NOMEM : SSB_OK;  // add_frames_f32: ragged -> Error::NoMem
  DeviceGuard g(h->device);
  FftPlan* plan = nullptr;
  size_t nb = 0;
  if (*fft_status == SSB_OK) {
    int32_t rc = get_plan(h, n_fft, h->rate, &plan);
    if (rc) return rc;
    nb = plan->n_bins;
    *n_points = nb;
    if (nb > cap || !xy_mid || !xy_side) return fail(h, SSB_ERR_CAPACITY, "process_tick: need room for %zu points", nb);
  }
  const size_t in_bytes = 2 * n_fft * sizeof(float);
  const size_t out_off = (in_bytes + 255) & ~(size_t)255;
  const size_t db_bytes = 2 * nb * sizeof(float) + 2 * sizeof(int32_t);
  const size_t stride = 4 + 2 * (size_t)h->channels;
  int32_t rc = ensure_scratch(h, out_off + db_bytes + 256);
  if (rc) return rc;
  float* d_in = h->d_scratch;
  float* d_db = reinterpret_cast<float*>(reinterpret_cast<char*>(h->d_scratch) + out_off);
  int32_t* d_status = reinterpret_cast<int32_t*>(d_db + 2 * nb);
  CK(cudaMemcpyAsync(d_in, tail, in_bytes, cudaMemcpyHostToDevice, h->stream));
  if (plan) {
    CK(launch_fft(*plan, d_in, SSB_FFT_MID_SIDE, 1, d_db, d_status, h->stream, &h->launches));
    CK(cudaMemcpyAsync(h->h_scratch, d_db, db_bytes, cudaMemcpyDeviceToHost, h->stream));
  }
  if (*lufs_status == SSB_OK && lufs_samples) {
    rc = feed_device(h, d_in + (2 * n_fft - lufs_samples), lufs_samples / 2, lufs_samples / 2);
    if (rc) return rc;
  }
  const int aligned = (h->total_frames % h->lp.s100) == 0;
  if (!h->st.ring && !aligned) *lufs_status = *lufs_status ? *lufs_status : SSB_ERR_UNALIGNED_QUERY;
  {
    const int32_t rrc = launch_results_now(h);
    if (rrc) return rrc;
  }
  CK(cudaMemcpyAsync(h->h_results, h->d_results, stride * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->results_valid = true;
  *shortterm_lufs = h->h_results[1];
  if ((h->mode & SSB_MODE_S) != SSB_MODE_S && *lufs_status == SSB_OK) *lufs_status = SSB_ERR_INVALID_MODE;
  if (plan) {
    const float* db = static_cast<const float*>(h->h_scratch);
    const int32_t* st = reinterpret_cast<const int32_t*>(db + 2 * nb);
    *fft_status = st[0] ? st[0] : st[1];
    h->tick_fft_status[0] = st[0];   // the reference handles the two get_fft results independently (tui.rs:1505-1523)
    h->tick_fft_status[1] = st[1];
    const auto& axes = fft_axis_cached(h, n_fft, h->rate);
    const std::vector<double>& ax = axes.first;
    const std::vector<double>& tilt = axes.second;
    for (size_t i = 0; i < nb; i++) {
      xy_mid[2 * i] = ax[i];
      xy_mid[2 * i + 1] = (double)db[i] + tilt[i];
      xy_side[2 * i] = ax[i];
      xy_side[2 * i + 1] = (double)db[nb + i] + tilt[i];
    }
  }
  return SSB_OK;
}

int32_t ssb_waveform_device(ssb_analyzer* h, const float* d_samples, size_t len, double waveform_window,
                            float* d_minmax_out, size_t cap_columns, size_t* n_columns) {
  if (!h || !n_columns) return SSB_ERR_INVALID_ARG;
  size_t window = 0;
  const size_t cols = waveform_window_columns(waveform_window, len, &window);
  *n_columns = cols;
  if (cols > cap_columns) return fail(h, SSB_ERR_CAPACITY, "get_waveform: need room for %zu columns", cols);
  if (!cols) return SSB_OK;
  if (!d_samples || !d_minmax_out) return SSB_ERR_INVALID_ARG;
  DeviceGuard g(h->device);
  CK(launch_waveform(d_samples, len, window, d_minmax_out, cols, h->stream, &h->launches));
  return SSB_OK;
}

int32_t ssb_get_waveform(ssb_analyzer* h, const float* samples, size_t len, double waveform_window, double* xy_out,
                         size_t cap, size_t* n_points) {
  if (!h || !n_points || (!samples && len)) return SSB_ERR_INVALID_ARG;
  size_t window = 0;
  const size_t cols = waveform_window_columns(waveform_window, len, &window);
  *n_points = 2 * cols;
  if (2 * cols > cap || (cols && !xy_out)) return fail(h, SSB_ERR_CAPACITY, "get_waveform: need room for %zu points", 2 * cols);
  if (!cols) return SSB_OK;
  DeviceGuard g(h->device);
  const size_t in_bytes = len * sizeof(float), out_bytes = cols * 2 * sizeof(float);
  const size_t out_off = (in_bytes + 255) & ~(size_t)255;
  int32_t rc = ensure_scratch(h, out_off + out_bytes);
  if (rc) return rc;
  float* d_in = h->d_scratch;
  float* d_out = reinterpret_cast<float*>(reinterpret_cast<char*>(h->d_scratch) + out_off);
  CK(cudaMemcpyAsync(d_in, samples, in_bytes, cudaMemcpyHostToDevice, h->stream));
  CK(launch_waveform(d_in, len, window, d_out, cols, h->stream, &h->launches));
  CK(cudaMemcpyAsync(h->h_scratch, d_out, out_bytes, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const float* mm = static_cast<const float*>(h->h_scratch);
  for (size_t i = 0; i < cols; i++) {
    xy_out[4 * i + 0] = (double)i;
    xy_out[4 * i + 1] = (double)mm[2 * i];
    xy_out[4 * i + 2] = (double)i;
    xy_out[4 * i + 3] = (double)mm[2 * i + 1];
  }
  return SSB_OK;
}

int32_t ssb_preanalyze_file(ssb_analyzer* h, const float* samples, size_t len, uint32_t rate, double duration_s,
                            double* xy_out, size_t cap, size_t* n_points, double* integrated, int32_t* is_some) {
  if (!h || !n_points || !integrated || !is_some || (!samples && len)) return SSB_ERR_INVALID_ARG;
  *is_some = 0;
  // tui.rs:1213: waveform over the whole interleaved file
  int32_t rc = ssb_get_waveform(h, samples, len, duration_s, xy_out, cap, n_points);
  if (rc) return rc;
  // tui.rs:1218: the reference keeps going when the meter cannot be created (it shows an error popup)
  const int32_t meter_rc = ssb_create_loudness_meter(h, 2, rate);
  (void)meter_rc;
  // tui.rs:1229: integrated loudness of the file, channels hard-coded to 2.  The samples are already in the
  // handle's device scratch (ssb_get_waveform put them there and synchronised): the one-shot meter reads that copy.
  const bool resident = n_points && *n_points > 0 && h->d_scratch && h->scratch_cap >= len * sizeof(float);
  return one_shot_integrated(h, 2, samples, resident ? h->d_scratch : nullptr, len, integrated, is_some);
}

int32_t ssb_mid_side_device(ssb_analyzer* h, const float* d_interleaved, size_t len, float* d_mid, float* d_side) {
  if (!h) return SSB_ERR_INVALID_ARG;
  const size_t frames = len / 2;
  if (!frames) return SSB_OK;
  if (!d_interleaved || !d_mid || !d_side) return SSB_ERR_INVALID_ARG;
  if (((uintptr_t)d_interleaved & 7) != 0)   // read as (l, r) pairs with 8-byte loads
    return fail(h, SSB_ERR_INVALID_ARG, "ssb_mid_side_device: d_interleaved must be 8-byte aligned (got %p)", (const void*)d_interleaved);
  DeviceGuard g(h->device);
  CK(launch_mid_side(d_interleaved, frames, d_mid, d_side, h->stream, &h->launches));
  return SSB_OK;
}

int32_t ssb_mid_side(ssb_analyzer* h, const float* interleaved, size_t len, float* mid, float* side, size_t* frames_out) {
  if (!h || !frames_out || (!interleaved && len)) return SSB_ERR_INVALID_ARG;
  const size_t frames = len / 2;  // zip(): an odd trailing sample is dropped
  *frames_out = frames;
  if (!frames) return SSB_OK;
  if (!mid || !side) return SSB_ERR_INVALID_ARG;
  DeviceGuard g(h->device);
  const size_t in_bytes = frames * 2 * sizeof(float), half = frames * sizeof(float);
  const size_t out_off = (in_bytes + 255) & ~(size_t)255;
  int32_t rc = ensure_scratch(h, out_off + 2 * half);
  if (rc) return rc;
  float* d_in = h->d_scratch;
  float* d_mid = reinterpret_cast<float*>(reinterpret_cast<char*>(h->d_scratch) + out_off);
  float* d_side = d_mid + frames;
  CK(cudaMemcpyAsync(d_in, interleaved, in_bytes, cudaMemcpyHostToDevice, h->stream));
  CK(launch_mid_side(d_in, frames, d_mid, d_side, h->stream, &h->launches));
  CK(cudaMemcpyAsync(mid, d_mid, half, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(side, d_side, half, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return SSB_OK;
}

int32_t ssb_tick_fft_status(const ssb_analyzer* h, int32_t mid_side[2]) {
  if (!h || !mid_side) return SSB_ERR_INVALID_ARG;
  mid_side[0] = h->tick_fft_status[0];
  mid_side[1] = h->tick_fft_status[1];
  return SSB_OK;
}

int32_t ssb_debug_force_generic(ssb_analyzer* h, int32_t on) {
  if (!h) return SSB_ERR_INVALID_ARG;
  h->force_kernel = on;
  return SSB_OK;
}

int32_t ssb_true_peak_factor(const ssb_analyzer* h) { return h ? h->lp.tp_factor : 0; }

int32_t ssb_debug_force_true_peak_factor(ssb_analyzer* h, int32_t factor) {
  if (!h || (factor != 2 && factor != 4)) return SSB_ERR_INVALID_ARG;
  if (!h->lp.do_true_peak) return fail(h, SSB_ERR_INVALID_MODE, "mode lacks TRUE_PEAK (Error::InvalidMode)");
  DeviceGuard g(h->device);
  CK(cudaStreamSynchronize(h->stream));
  h->lp.tp_factor = truepeak_taps(h->rate, h->lp.tp4, h->lp.tp2, factor);
  return SSB_OK;
}

int32_t ssb_profile_enable(ssb_analyzer* h, int32_t on) {
  if (!h) return SSB_ERR_INVALID_ARG;
  DeviceGuard g(h->device);
  CK(cudaStreamSynchronize(h->stream));
  h->profiling = on != 0;
  h->prof_used = 0;
  return SSB_OK;
}

int32_t ssb_profile_read(ssb_analyzer* h, double* filter_ms, uint64_t* filter_launches) {
  if (!h || !filter_ms || !filter_launches) return SSB_ERR_INVALID_ARG;
  DeviceGuard g(h->device);
  CK(cudaStreamSynchronize(h->stream));
  double total = 0.0;
  for (size_t i = 0; i < h->prof_used; i++) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->prof_events[i].first, h->prof_events[i].second));
    total += ms;
  }
  *filter_ms = total;
  *filter_launches = h->prof_used;
  h->prof_used = 0;
  return SSB_OK;
}

int32_t ssb_filter_coeffs(const ssb_analyzer* h, double b[5], double a[5]) {
  if (!h || !b || !a) return SSB_ERR_INVALID_ARG;
  memcpy(b, h->lp.b, sizeof(h->lp.b));
  memcpy(a, h->lp.a, sizeof(h->lp.a));
  return SSB_OK;
}

int32_t ssb_debug_histogram_index(ssb_analyzer* h, const double* energies, size_t n, int32_t* idx_out) {
  if (!h || (n && (!energies || !idx_out))) return SSB_ERR_INVALID_ARG;
  if (!h->meter_ok) return fail(h, SSB_ERR_NOMEM, "the loudness meter is not initialised");
  if (!n) return SSB_OK;
  DeviceGuard g(h->device);
  const size_t in_bytes = (n * sizeof(double) + 255) & ~(size_t)255;
  int32_t rc = ensure_scratch_device(h, in_bytes + n * sizeof(int32_t));
  if (rc) return rc;
  double* d_e = reinterpret_cast<double*>(h->d_scratch);
  int32_t* d_out = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(h->d_scratch) + in_bytes);
  CK(cudaMemcpyAsync(d_e, energies, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(launch_histogram_index(h->st, d_e, n, d_out, h->stream));
  CK(cudaMemcpyAsync(idx_out, d_out, n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return SSB_OK;
}

int32_t ssb_histograms(ssb_analyzer* h, size_t stream, uint64_t block[1000], uint64_t shortterm[1000]) {
  if (!h || stream >= h->n_streams || !block || !shortterm) return SSB_ERR_INVALID_ARG;
  DeviceGuard g(h->device);
  {
    int32_t rc = flush_gating(h);
    if (rc) return rc;
  }
  std::vector<uint32_t> tmp(2 * kHistBins);
  CK(cudaMemcpyAsync(tmp.data(), h->st.block_hist + stream * kHistBins, kHistBins * sizeof(uint32_t),
                     cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(tmp.data() + kHistBins, h->st.st_hist + stream * kHistBins, kHistBins * sizeof(uint32_t),
                     cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < kHistBins; i++) { block[i] = tmp[i]; shortterm[i] = tmp[kHistBins + i]; }
  return SSB_OK;
}

}  // extern "C"
