/*
 * soundscope_b200.h — C ABI of the B200-native analyzer hot path (libsoundscope_b200.so).
 *
 * The reference (bananaofhappiness/soundscope v1.9.0) has no FFI for this path: the boundary is
 * the inherent-method surface of `analyzer::Analyzer` (reference src/analyzer.rs:47-183) plus the
 * free function `get_mid_and_side_samples` (reference src/audio_player.rs:400-419).  Every entry
 * point below names the reference item it replaces.  A Rust shim that keeps `analyzer.rs`'s
 * signatures and forwards to these symbols is given in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns an int32 status (SSB_OK == 0);
 *   - a handle owns n_streams independent meters ("streams"); n_streams == 1 is exactly one
 *     reference `Analyzer`.  Batched inputs are stream-major:  in[s][frame][channel] (each
 *     stream's slice is what the reference would pass to `add_samples`);
 *   - `*_device` variants take DEVICE pointers and are asynchronous on the handle's stream;
 *     host variants copy through pinned staging and are complete (results valid, caller's
 *     buffer reusable) at return;
 *   - a handle is not thread-safe (the reference's `&mut self`); distinct handles may be used
 *     concurrently;
 *   - there is no CPU fallback: if no sm_100-class device is usable, create fails with
 *     SSB_ERR_NO_DEVICE / SSB_ERR_CUDA.
 */
#ifndef SOUNDSCOPE_B200_H
#define SOUNDSCOPE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSB_ABI_VERSION 2

/* ---- status codes ------------------------------------------------------------------------ */
enum {
  SSB_OK = 0,
  /* ebur128::Error (returned by add_samples / get_*_lufs / get_true_peak, analyzer.rs:139-164) */
  SSB_ERR_NOMEM = 1,                 /* Error::NoMem: channels 0 or >64, rate <16 or >2822400, ragged input */
  SSB_ERR_INVALID_MODE = 2,          /* Error::InvalidMode */
  SSB_ERR_INVALID_CHANNEL_INDEX = 3, /* Error::InvalidChannelIndex (get_true_peak on a mono meter) */
  /* spectrum_analyzer::SpectrumAnalyzerError surfaced by get_fft (analyzer.rs:60-65) */
  SSB_ERR_FFT_TOO_FEW_SAMPLES = 4,
  SSB_ERR_FFT_NAN = 5,
  SSB_ERR_FFT_INF = 6,
  SSB_ERR_FFT_NOT_POW2 = 7,          /* not a power of two, or above the crate's largest size (32768) */
  SSB_ERR_FFT_BAD_LIMIT = 8,         /* 20 kHz upper limit above Nyquist */
  SSB_ERR_FFT_SCALING = 9,
  /* this library's own */
  SSB_ERR_INVALID_ARG = 10,
  SSB_ERR_CAPACITY = 11,             /* caller's output buffer too small; required size written to *n_out */
  SSB_ERR_UNALIGNED_QUERY = 12,      /* momentary/short-term query off the 100 ms grid on a handle built without SSB_FLAG_RING */
  SSB_ERR_NO_DEVICE = 13,
  SSB_ERR_BUSY = 14,                 /* a capture-ring snapshot kept colliding with the producer; nothing was wrong, try again */
  SSB_ERR_CUDA = 100                 /* 100 + cudaError_t */
};

/* ---- ebur128::Mode bits (the reference always passes Mode::all(), analyzer.rs:36,51,171) --- */
enum {
  SSB_MODE_M = 1 << 0,
  SSB_MODE_S = (1 << 1) | SSB_MODE_M,
  SSB_MODE_I = (1 << 2) | SSB_MODE_M,
  SSB_MODE_LRA = (1 << 3) | SSB_MODE_S,
  SSB_MODE_SAMPLE_PEAK = (1 << 4) | SSB_MODE_M,
  SSB_MODE_TRUE_PEAK = (1 << 5) | SSB_MODE_M | SSB_MODE_SAMPLE_PEAK,
  SSB_MODE_HISTOGRAM = 1 << 6,
  SSB_MODE_ALL = 0x7f
};

/* ---- create flags ------------------------------------------------------------------------ */
enum {
  /* keep the 3 s ring of K-weighted samples per stream (what ebur128 keeps), so momentary and
   * short-term queries are exact at ANY feed position, as the reference's tick loop needs
   * (tui.rs:1528-1543 feeds 8192 frames per tick).  Costs 8 B/sample of extra HBM writes and
   * 24*rate*channels bytes per stream; meant for few-stream, reference-shaped use.  Without it
   * the handle keeps only per-100 ms energy sums (O(1) state per stream, the batch mode) and
   * M/S queries are exact on the 100 ms grid and refused off it. */
  SSB_FLAG_RING = 1 << 0
};

typedef struct ssb_analyzer ssb_analyzer;

/* ---- lifetime ---------------------------------------------------------------------------- */
/* Analyzer::default() (analyzer.rs:34-45) == ssb_analyzer_create(&h, 2, 44100, SSB_MODE_ALL, 1, dev, SSB_FLAG_RING).
 * `device` < 0 selects the current CUDA device. */
int32_t ssb_analyzer_create(ssb_analyzer** out, uint32_t channels, uint32_t rate, int32_t mode,
                            size_t n_streams, int32_t device, uint32_t flags);
void ssb_analyzer_destroy(ssb_analyzer* h);
/* Analyzer::create_loudness_meter (analyzer.rs:49-53): replace the meter(s), keep n_streams/flags/mode. */
int32_t ssb_create_loudness_meter(ssb_analyzer* h, uint32_t channels, uint32_t rate);
/* Analyzer::sample_rate (analyzer.rs:166-168) */
uint32_t ssb_sample_rate(const ssb_analyzer* h);
uint32_t ssb_channels(const ssb_analyzer* h);
size_t ssb_n_streams(const ssb_analyzer* h);
/* last error text of this handle (never NULL) */
const char* ssb_last_error(const ssb_analyzer* h);
/* run this handle's work on a caller-owned CUDA stream (cudaStream_t; NULL is the CUDA default stream) */
int32_t ssb_set_stream(ssb_analyzer* h, void* cuda_stream);
/* go back to the handle's own non-blocking stream (the state after create) */
int32_t ssb_use_own_stream(ssb_analyzer* h);
int32_t ssb_sync(ssb_analyzer* h);
/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
uint64_t ssb_launch_count(const ssb_analyzer* h);

/* ---- loudness: Analyzer::add_samples / reset / get_* (analyzer.rs:139-164) ---------------- */
/* add_samples(&[f32]) -> EbuR128::add_frames_f32.  `interleaved` holds n_streams slices of
 * frames_per_stream*channels f32 (HOST memory, pageable or pinned). */
int32_t ssb_add_frames_f32(ssb_analyzer* h, const float* interleaved, size_t frames_per_stream);
/* same, `interleaved` is a DEVICE pointer; asynchronous on the handle's stream */
int32_t ssb_add_frames_f32_device(ssb_analyzer* h, const float* d_interleaved, size_t frames_per_stream);
/* the reference's slice-length form: len = number of f32 for ONE stream; len % channels != 0 -> SSB_ERR_NOMEM */
int32_t ssb_add_samples(ssb_analyzer* h, const float* interleaved, size_t len);
int32_t ssb_reset(ssb_analyzer* h);
/* out[n_streams]; -inf (not an error) where the reference returns -inf */
int32_t ssb_loudness_momentary(ssb_analyzer* h, double* out);
int32_t ssb_loudness_shortterm(ssb_analyzer* h, double* out); /* get_shortterm_lufs, analyzer.rs:147 */
int32_t ssb_loudness_global(ssb_analyzer* h, double* out);    /* get_integrated_lufs, analyzer.rs:151 */
int32_t ssb_loudness_range(ssb_analyzer* h, double* out);     /* get_loudness_range,  analyzer.rs:155 */
/* out[n_streams*channels], linear amplitude = max(true peak, sample peak) as EbuR128::true_peak */
int32_t ssb_true_peak(ssb_analyzer* h, double* out);
int32_t ssb_sample_peak(ssb_analyzer* h, double* out);
/* get_true_peak (analyzer.rs:159-164): channels 0 and 1 of stream 0; mono -> SSB_ERR_INVALID_CHANNEL_INDEX */
int32_t ssb_get_true_peak(ssb_analyzer* h, double* left, double* right);
/* all scalars in one pass, written to DEVICE memory as rows of ssb_result_stride(h) doubles:
 * [momentary, shortterm, global, range, true_peak[channels], sample_peak[channels]] — the buffer
 * the multi-GPU gather moves.  Asynchronous. */
size_t ssb_result_stride(const ssb_analyzer* h);
int32_t ssb_results_device(ssb_analyzer* h, double* d_out);
/* add_samples followed by the four queries (the per-tick pair analyzer.rs:139-141 + :147-164, tui.rs:1539-1543), as
 * one call on DEVICE memory: feeds frames_per_stream frames per stream and writes the result rows for the new
 * position to d_out.  One kernel launch when the batch kernel applies (mono / stereo, whole 320-frame tiles: gating
 * and the result rows run in the filter kernel's epilogue); otherwise the same as ssb_add_frames_f32_device +
 * ssb_results_device.  Asynchronous. */
int32_t ssb_add_frames_f32_device_results(ssb_analyzer* h, const float* d_interleaved, size_t frames_per_stream,
                                          double* d_out);
/* ---- multi-GPU: the gather of the per-stream result rows -------------------------------
 * Streams are independent (every meter's state is private to its EbuR128, analyzer.rs:29-32), so ranks own disjoint
 * blocks of streams and the only exchange is the gather of the result rows.  Here the kernel that computes a rank's
 * rows also stores them into block `rank` of EVERY rank's gather buffer (peer memory over NVLink, mapped with CUDA
 * IPC) — no collective kernel.  All ranks use the same n_streams / channels.  Protocol:
 *   ssb_gather_create(h, world, rank, handle64)   allocate; returns this rank's 64-byte IPC handle
 *   (exchange the handles through your process group)
 *   ssb_gather_open(h, handles)                   handles = world x 64 bytes in rank order (own slot ignored)
 *   ... every ssb_results_device / ssb_add_frames_f32_device_results now also publishes into parity `p` ...
 *   ssb_gather_wait(h)                            enqueue: fence, tell every rank, wait until every rank has told us
 *                                                 (collective: every rank calls it the same number of times)
 *   ssb_gather_rows(h, p)                         DEVICE pointer to [world * n_streams][ssb_result_stride] f64
 *   ssb_gather_select(h, p ^ 1)                   next publishes go to the other half while `p` is read */
int32_t ssb_gather_create(ssb_analyzer* h, uint32_t world, uint32_t rank, void* ipc_handle_out);
int32_t ssb_gather_open(ssb_analyzer* h, const void* ipc_handles);
int32_t ssb_gather_select(ssb_analyzer* h, int32_t parity);
double* ssb_gather_rows(ssb_analyzer* h, int32_t parity);
uint64_t ssb_gather_epoch(const ssb_analyzer* h);
int32_t ssb_gather_wait(ssb_analyzer* h);
int32_t ssb_gather_destroy(ssb_analyzer* h);

/* Analyzer::calculate_integrated_lufs (analyzer.rs:170-182): a fresh meter at the handle's rate over the whole
 * interleaved file; a `sample_rate*2`-sample chunk that is not whole frames (or an invalid channel count) gives
 * *is_some = 0, the reference's `None`.  The reference builds the meter with Mode::all() but reads nothing except
 * loudness_global(), so only K-weighting and gating run here; mono / stereo files of a second or more are cut into
 * time chunks that run on different SMs (each from a 0.4 s zero-state run-in: the inherited state has decayed by e^-95,
 * block energies agree with the serial pass to the recursion's rounding noise, < 1e-8 LU). */
int32_t ssb_calculate_integrated_lufs(ssb_analyzer* h, uint32_t channels, const float* samples,
                                      size_t len, double* out, int32_t* is_some);

/* ---- spectrum: Analyzer::get_fft (analyzer.rs:55-105) -------------------------------------- */
/* one mono window of n samples (HOST) -> (x, dB) pairs, x = log-frequency position 0..100.
 * xy_out holds cap pairs (2*cap doubles); *n_points = pairs produced (or required on SSB_ERR_CAPACITY). */
int32_t ssb_get_fft(ssb_analyzer* h, const float* samples, size_t n, double* xy_out, size_t cap,
                    size_t* n_points);
/* number of bins get_fft keeps for (n, rate) and the first kept bin index */
int32_t ssb_fft_bins(size_t n, uint32_t rate, size_t* k_first, size_t* n_bins);
/* the (n, rate)-only part of get_fft's output: x[k] and the pink tilt 10*log10(f/1000) per kept bin */
int32_t ssb_fft_axis(size_t n, uint32_t rate, double* x_out, double* tilt_out, size_t cap, size_t* n_bins);
enum { SSB_FFT_MONO = 0, SSB_FFT_MID_SIDE = 1 };
/* batched: n_windows windows of n frames (DEVICE).  SSB_FFT_MONO: in[w][n] f32.  SSB_FFT_MID_SIDE:
 * in[w][n][2] interleaved stereo; mid/side are formed as get_mid_and_side_samples does
 * (audio_player.rs:400-419) and both spectra are produced.  d_db_out[w][planes][n_bins] f32 receives
 * scale_to_dbfs (analyzer.rs:11-27) BEFORE the f64 tilt (planes = 1 or 2); add ssb_fft_axis's tilt to
 * get the reference's y.  d_status[w] (optional) receives per-window SSB_ERR_FFT_* codes. 
 * ALIGNMENT: d_in must be 16-byte aligned (windows are read with vector loads); a misaligned pointer returns
 * SSB_ERR_INVALID_ARG instead of faulting. */
int32_t ssb_fft_batch_device(ssb_analyzer* h, const float* d_in, int32_t layout, size_t n,
                             size_t n_windows, float* d_db_out, int32_t* d_status);
/* The same with the reference's y as the output: d_y_out[w][planes][n_bins] f64 = (f64) scale_to_dbfs + tilt, the
 * addition analyzer.rs:80-94 performs in f64, done in the kernel's store (for SSB_FFT_MONO bit-identical to the y of
 * ssb_get_fft on the same window).  x comes from ssb_fft_axis.  d_y_out must be 8-byte aligned. */
int32_t ssb_fft_batch_device_y(ssb_analyzer* h, const float* d_in, int32_t layout, size_t n,
                               size_t n_windows, double* d_y_out, int32_t* d_status);

/* ---- one player tick in one call (north_star's `Analyzer::process`) ------- */
/* What tui.rs:1482-1552 (analyze_audio_file_samples) does per playback-position message, fused:
 *   mid/side of the last n_fft stereo frames -> get_fft(mid), get_fft(side)   (tui.rs:1488-1515)
 *   add_samples(last lufs_samples interleaved samples) -> get_shortterm_lufs  (tui.rs:1528-1543)
 * `tail` holds the last n_fft stereo frames (2*n_fft interleaved f32, HOST); the meter is fed the final
 * lufs_samples of it (the reference passes 16384).  One H2D copy, three kernels, one D2H copy.
 * xy_mid / xy_side receive (x, dB) pairs like ssb_get_fft (cap pairs each).  fft_status / lufs_status return
 * the per-part status so the caller can reproduce the reference's independent error handling
 * (`vec![(0., 0.)]` on an FFT error, error popup on a meter error).  Needs a stereo handle. */
int32_t ssb_process_tick(ssb_analyzer* h, const float* tail, size_t n_fft, size_t lufs_samples,
                         double* xy_mid, double* xy_side, size_t cap, size_t* n_points,
                         double* shortterm_lufs, int32_t* fft_status, int32_t* lufs_status);

/* ---- whole-file pre-analysis in one call -------------------------------- */
/* What tui.rs:1207-1241 (receive_audio_file) does when a file is selected, with one H2D copy:
 *   Analyzer::get_waveform(samples, duration_s)          -> xy_out / n_points     (tui.rs:1213-1216)
 *   create_loudness_meter(2, rate) on this handle                                  (tui.rs:1218-1222)
 *   calculate_integrated_lufs(2, samples)                 -> integrated / is_some  (tui.rs:1229-1233)
 * `samples` is the whole interleaved file (HOST).  The reference hard-codes 2 channels for the meter. */
int32_t ssb_preanalyze_file(ssb_analyzer* h, const float* samples, size_t len, uint32_t rate, double duration_s,
                            double* xy_out, size_t cap, size_t* n_points, double* integrated, int32_t* is_some);

/* ---- decoded-PCM input ---------------------------------------------------- */
/* What AudioFile::decode_file (audio_player.rs:169-267) hands the analyzer for WAV / AIFF files: symphonia's
 * PCM decoder reads the container's interleaved integer/float samples and
 * `SampleBuffer::<f32>::copy_interleaved_ref` (audio_player.rs:248) converts each to f32 with
 * symphonia-core 0.5.5's `FromSample` rules (Cargo.lock:2085-2087; un-vendored, restated in
 * oracle/capture_ref.py):  u8: (s-128)/128;  s8: s/128;  s16: s/32768;  s24: s/8388608;
 * s32: ((s as f64)/2147483648) as f32;  f32: identity;  f64: s as f32.  Every rule is exact or a single
 * round-to-nearest-even, so the device conversion is bit-exact.  Compressed codecs (mp3, aac, flac, vorbis,
 * alac, adpcm) are serial bitstreams and stay on the CPU. */
enum {
  SSB_PCM_U8 = 0, SSB_PCM_S8 = 1,
  SSB_PCM_S16LE = 2, SSB_PCM_S16BE = 3,
  SSB_PCM_S24LE = 4, SSB_PCM_S24BE = 5,   /* 3 bytes per sample, packed */
  SSB_PCM_S32LE = 6, SSB_PCM_S32BE = 7,
  SSB_PCM_F32LE = 8, SSB_PCM_F32BE = 9,
  SSB_PCM_F64LE = 10, SSB_PCM_F64BE = 11
};
/* bytes per sample of a format; 0 for an unknown one */
size_t ssb_pcm_bytes_per_sample(int32_t format);
/* n_samples interleaved PCM samples (HOST) -> n_samples f32 (HOST), the Vec<f32> decode_file returns */
int32_t ssb_pcm_to_f32(ssb_analyzer* h, const void* pcm, size_t n_samples, int32_t format, float* out);
/* same on DEVICE pointers, asynchronous on the handle's stream */
int32_t ssb_pcm_to_f32_device(ssb_analyzer* h, const void* d_pcm, size_t n_samples, int32_t format, float* d_out);
/* add_samples on raw PCM: n_streams slices of frames_per_stream*channels samples.  HOST form: the raw bytes
 * cross PCIe (2 B/sample for 16-bit audio instead of 4), are converted on the device and fed to the meter. */
int32_t ssb_add_frames_pcm(ssb_analyzer* h, const void* pcm, int32_t format, size_t frames_per_stream);
int32_t ssb_add_frames_pcm_device(ssb_analyzer* h, const void* d_pcm, int32_t format, size_t frames_per_stream);

/* ---- capture ring + one microphone tick (SURVEY.md §8(f)-4 and the live half of §8(f)-1) --------- */
/* `RBuffer` (tui.rs:37): AllocRingBuffer<f32> of 30*rate VALUES, zero-filled (main.rs:63-65,
 * tui.rs:1783-1786), written by the cpal input callback (audio_capture.rs:40-52) and read whole by the TUI
 * thread every tick (tui.rs:1428, 1458).  Here the ring is pinned host memory the capture thread writes with
 * plain stores (no CUDA call, no lock: single producer, release-published write counter) plus a device mirror
 * that each tick tops up with only the values written since the previous tick. */
typedef struct ssb_capture_ring ssb_capture_ring;
int32_t ssb_capture_ring_create(ssb_capture_ring** out, size_t capacity_values, int32_t device);
void ssb_capture_ring_destroy(ssb_capture_ring* r);
size_t ssb_capture_ring_capacity(const ssb_capture_ring* r);
/* total values pushed since create (monotonic) */
uint64_t ssb_capture_ring_written(const ssb_capture_ring* r);
/* the input callback (audio_capture.rs:40-52): is_mono == 0: `extend(data)`; is_mono != 0: the reference's
 * up-mix `[x0, 0, x1, 0, x2, ...]` (first sample alone, every later one preceded by 0.0: 2n-1 values, so each
 * callback flips the left/right parity of what follows — reproduced as is).  Producer thread only. */
int32_t ssb_capture_ring_push(ssb_capture_ring* r, const float* data, size_t n, int32_t is_mono);
/* `to_vec()`: the capacity values oldest -> newest (HOST out, cap >= capacity).  Consumer thread. */
int32_t ssb_capture_ring_to_vec(ssb_capture_ring* r, float* out, size_t cap);
/* analyze_microphone_input (tui.rs:1427-1480) in one call, on one snapshot of the ring:
 *   (mid, side) = get_mid_and_side_samples(ring.to_vec())                              (tui.rs:1428-1429)
 *   get_fft(mid[15*rate - n_fft .. 15*rate]), get_fft(side[..same..])  -> xy_mid, xy_side (tui.rs:1431-1452)
 *   Analyzer::get_waveform(mid, waveform_window)                       -> xy_wave          (tui.rs:1456)
 *   add_samples(ring.to_vec()[30*rate - lufs_samples .. 30*rate]); get_shortterm_lufs()    (tui.rs:1458-1479)
 * with rate = the handle's sample rate; the reference passes n_fft = lufs_samples = 16384 and 15.0.  The meter
 * may have any channel count (a mono device gets a mono meter fed the up-mixed values, as the reference does).
 * Ranges the reference's slices would panic on (ring shorter than 30*rate values, n_fft > 15*rate, ...) return
 * SSB_ERR_INVALID_ARG.  xy_wave holds wave_cap pairs; statuses as in ssb_process_tick. */
int32_t ssb_mic_tick(ssb_analyzer* h, ssb_capture_ring* ring, size_t n_fft, size_t lufs_samples,
                     double waveform_window, double* xy_mid, double* xy_side, size_t cap, size_t* n_points,
                     double* xy_wave, size_t wave_cap, size_t* n_wave_points, double* shortterm_lufs,
                     int32_t* fft_status, int32_t* lufs_status);

/* ---- waveform + mid/side (stateless) ------------------------------------------------------ */
/* Analyzer::get_waveform (analyzer.rs:107-137): HOST samples -> (i, min), (i, max) pairs. */
int32_t ssb_get_waveform(ssb_analyzer* h, const float* samples, size_t len, double waveform_window,
                         double* xy_out, size_t cap, size_t* n_points);
/* device form: d_minmax_out[2*columns] f32 (min, max per column); *n_columns = columns produced */
int32_t ssb_waveform_device(ssb_analyzer* h, const float* d_samples, size_t len, double waveform_window,
                            float* d_minmax_out, size_t cap_columns, size_t* n_columns);
/* get_mid_and_side_samples (audio_player.rs:400-419): HOST in, HOST out; *frames = len/2 */
int32_t ssb_mid_side(ssb_analyzer* h, const float* interleaved, size_t len, float* mid, float* side,
                     size_t* frames);
/* device form.  ALIGNMENT: d_interleaved must be 8-byte aligned (SSB_ERR_INVALID_ARG otherwise). */
int32_t ssb_mid_side_device(ssb_analyzer* h, const float* d_interleaved, size_t len, float* d_mid,
                            float* d_side);

/* The two get_fft statuses of the last ssb_process_tick / ssb_mic_tick separately: mid_side[0] for the mid spectrum,
 * mid_side[1] for the side spectrum (the reference handles the two results independently, tui.rs:1505-1523; the tick
 * calls' own fft_status is the first non-zero of the two). */
int32_t ssb_tick_fft_status(const ssb_analyzer* h, int32_t mid_side[2]);

/* ---- kernel timing (bench.py's roofline leg) ------------------------------------------------ */
/* when enabled, every K-weighting (filter) launch is bracketed by CUDA events on the handle's stream */
int32_t ssb_profile_enable(ssb_analyzer* h, int32_t on);
/* synchronises, then returns and clears the accumulated filter-kernel time and launch count */
int32_t ssb_profile_read(ssb_analyzer* h, double* filter_ms, uint64_t* filter_launches);

/* tests only: pick the filter kernel — 0 automatic, 1 generic (thread per channel), 2 serial many-streams
 * kernel, 3 round-1 time-segmented tile kernel, 4 few-streams scan kernel (generic when it does not apply),
 * 5 / 6 the warp-pipelined batch kernel with mixed T4/T5 warps / uniform T4 warps */
int32_t ssb_debug_force_generic(ssb_analyzer* h, int32_t on);

/* The true-peak oversampling factor in use: ebur128's rate rule (4 below 96 kHz, 2 below 192 kHz, 0 above).
 * ssb_debug_force_true_peak_factor overrides it (2 or 4) — benchmarks only: BASELINE config 5 asks for "4x" at 96 kHz,
 * where the reference itself oversamples 2x, so results with a forced factor are NOT the reference's. */
int32_t ssb_true_peak_factor(const ssb_analyzer* h);
int32_t ssb_debug_force_true_peak_factor(ssb_analyzer* h, int32_t factor);

/* ---- introspection used by the tests ------------------------------------------------------ */
/* ebur128's find_histogram_index as the gating kernels evaluate it (closed-form guess corrected against the boundary
 * table), for n block energies in HOST memory: idx_out[i] = bin 0..999, or -1 below the absolute gate (-70 LUFS). */
int32_t ssb_debug_histogram_index(ssb_analyzer* h, const double* energies, size_t n, int32_t* idx_out);
int32_t ssb_filter_coeffs(const ssb_analyzer* h, double b[5], double a[5]);
/* copy the two 1000-bin histograms of stream s to HOST */
int32_t ssb_histograms(ssb_analyzer* h, size_t stream, uint64_t block[1000], uint64_t shortterm[1000]);
uint32_t ssb_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SOUNDSCOPE_B200_H */
