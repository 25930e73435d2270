/*
 * oracle.h — CPU restatement of soundscope's analyzer hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under oracle/ is part of the product.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load liboracle.so, and only
 * as the checker / the timed CPU baseline.  The product (soundscope_b200/) never links,
 * imports or calls it.
 *
 * PARITY STATUS
 *   first-party arithmetic (reference src/analyzer.rs, src/audio_player.rs): restated
 *   line by line, pinned by the reference's own unit-test inputs (tests/golden/).
 *   third-party arithmetic (crates ebur128 0.1.10, spectrum-analyzer 1.7.0,
 *   microfft 0.6.0, libm 0.2.16 — pinned in reference Cargo.lock:566-569,1941-1944,
 *   1074-1077,968-971 but NOT vendored under /root/reference, and no Rust toolchain
 *   exists in this image): restated from the published algorithms (libebur128, which
 *   the ebur128 crate ports; ITU-R BS.1770-4; EBU Tech 3341/3342; musl libm, which the
 *   libm crate ports).  For that part: PARITY UNPINNED against the reference binary;
 *   it is pinned instead against standards-based known answers (BS.1770 48 kHz
 *   coefficient table, EBU 3341/3342 synthetic cases, f64 numpy FFT).
 */
#ifndef SOUNDSCOPE_ORACLE_H
#define SOUNDSCOPE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ebur128::Mode bitflags (ebur128 0.1.10, mirrors libebur128's enum mode). */
enum {
  ORC_MODE_M = 1 << 0,
  ORC_MODE_S = (1 << 1) | ORC_MODE_M,
  ORC_MODE_I = (1 << 2) | ORC_MODE_M,
  ORC_MODE_LRA = (1 << 3) | ORC_MODE_S,
  ORC_MODE_SAMPLE_PEAK = (1 << 4) | ORC_MODE_M,
  ORC_MODE_TRUE_PEAK = (1 << 5) | ORC_MODE_M | ORC_MODE_SAMPLE_PEAK,
  ORC_MODE_HISTOGRAM = 1 << 6,
  ORC_MODE_ALL = 0x7f
};

/* ebur128::Error */
enum { ORC_OK = 0, ORC_ERR_NOMEM = 1, ORC_ERR_INVALID_MODE = 2, ORC_ERR_INVALID_CHANNEL_INDEX = 3 };

/* spectrum-analyzer error classes as surfaced through Analyzer::get_fft */
enum {
  ORC_FFT_OK = 0,
  ORC_FFT_TOO_FEW_SAMPLES = 4,
  ORC_FFT_NAN = 5,
  ORC_FFT_INF = 6,
  ORC_FFT_NOT_POW2 = 7,
  ORC_FFT_BAD_LIMIT = 8,
  ORC_FFT_SCALING = 9
};

typedef struct orc_ebur128 orc_ebur128;

/* EbuR128::new(channels, rate, mode)  — analyzer.rs:36,51,171 */
orc_ebur128* orc_ebur128_new(uint32_t channels, uint32_t rate, int mode, int* err);
void orc_ebur128_free(orc_ebur128* st);
/* EbuR128::add_frames_f32(interleaved) — analyzer.rs:140,176.  n_samples = frames*channels */
int orc_ebur128_add_frames_f32(orc_ebur128* st, const float* src, size_t n_samples);
void orc_ebur128_reset(orc_ebur128* st);                              /* analyzer.rs:144 */
int orc_ebur128_loudness_momentary(orc_ebur128* st, double* out);
int orc_ebur128_loudness_shortterm(orc_ebur128* st, double* out);    /* analyzer.rs:148 */
int orc_ebur128_loudness_global(orc_ebur128* st, double* out);       /* analyzer.rs:152,181 */
int orc_ebur128_loudness_range(orc_ebur128* st, double* out);        /* analyzer.rs:156 */
int orc_ebur128_true_peak(orc_ebur128* st, uint32_t ch, double* out);/* analyzer.rs:160-161 */
int orc_ebur128_sample_peak(orc_ebur128* st, uint32_t ch, double* out);
/* introspection for tests */
void orc_ebur128_coeffs(const orc_ebur128* st, double b[5], double a[5]);
void orc_ebur128_histograms(const orc_ebur128* st, uint64_t block[1000], uint64_t shortterm[1000]);
size_t orc_interp_taps(uint32_t rate, unsigned* factor, unsigned counts[4]);
double orc_histogram_energy(unsigned i);
double orc_histogram_boundary(unsigned i);
size_t orc_find_histogram_index(double energy);   /* ebur128 find_histogram_index (bisection), energy >= boundary[0] */

/* Analyzer::calculate_integrated_lufs — analyzer.rs:170-182.  returns 1 = Some, 0 = None */
int orc_calculate_integrated_lufs(uint32_t channels, uint32_t sample_rate, const float* samples,
                                  size_t len, double* out);

/* Batch driver used only as the timed CPU baseline: n_streams independent meters,
 * stream s reads in[s*frames*channels ..]; OpenMP over streams when threads > 1. */
typedef struct orc_batch orc_batch;
orc_batch* orc_batch_new(size_t n_streams, uint32_t channels, uint32_t rate, int mode);
orc_batch* orc_batch_new_mt(size_t n_streams, uint32_t channels, uint32_t rate, int mode, int threads);
void orc_batch_free(orc_batch* b);
int orc_batch_add_frames(orc_batch* b, const float* in, size_t frames, int threads);
void orc_batch_histograms(orc_batch* b, size_t s, uint64_t block[1000], uint64_t shortterm[1000]);
void orc_batch_query(orc_batch* b, double* momentary, double* shortterm, double* global,
                     double* range, double* true_peak /* n*channels */, int threads);

/* spectrum path — analyzer.rs:55-105 + spectrum-analyzer/microfft restatement */
float orc_cosf(float x);                                  /* libm 0.2.16 cosf (musl port) */
void orc_hann_window(const float* in, size_t n, float* out);
int orc_rfft_mag(const float* windowed, size_t n, float* mag /* n/2+1 */);
float orc_scale_to_dbfs(float val, float n);              /* analyzer.rs:11-27 */
/* Analyzer::get_fft: writes (x, dB) pairs; *n_out = number of pairs. */
int orc_get_fft(const float* samples, size_t n, uint32_t sample_rate, double* xy_out, size_t cap,
                size_t* n_out);
size_t orc_fft_bin_range(size_t n, uint32_t sample_rate, size_t* k_first);

/* Analyzer::get_waveform — analyzer.rs:107-137.  writes (x, value) pairs */
size_t orc_get_waveform(const float* samples, size_t len, double waveform_window, double* xy_out,
                        size_t cap);
/* get_mid_and_side_samples — audio_player.rs:400-419.  returns frames written */
size_t orc_mid_side(const float* interleaved, size_t len, float* mid, float* side);

#ifdef __cplusplus
}
#endif
#endif
