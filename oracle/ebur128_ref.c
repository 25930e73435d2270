/*
 * ebur128_ref.c — CPU restatement of the loudness meter the reference drives through
 * `ebur128::EbuR128` (crate ebur128 0.1.10, reference Cargo.lock:566-569; call sites
 * reference src/analyzer.rs:36,51,140,144,148,152,156,160-161,171-181).
 *
 * TEST INFRASTRUCTURE — see oracle.h.  The crate is a Rust port of libebur128 and is not
 * vendored in /root/reference; this file restates libebur128's published algorithm
 * (ITU-R BS.1770-4 K-weighting as one 4th-order DF-II filter in f64, 400 ms gating blocks
 * every 100 ms, 1000-bin 0.1 LU histograms for integrated loudness and LRA, 49-tap
 * Hann-windowed-sinc polyphase true-peak interpolator with the rate -> factor rule).
 * PARITY UNPINNED against the crate's binary; pinned by tests/test_oracle_kat.py against the
 * BS.1770 coefficient table and EBU Tech 3341/3342 style known answers.
 *
 * Build with -ffp-contract=off: Rust never contracts a*b+c into an FMA, so neither do we.
 */
#include "oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_CHANNELS 64
#define ALMOST_ZERO 0.000001
#define INTERP_TAPS 49
#define INTERP_MAX_DELAY 25

enum channel_kind { CH_UNUSED = 0, CH_LEFT, CH_RIGHT, CH_CENTER, CH_LEFT_SURROUND, CH_RIGHT_SURROUND };

/* ---- polyphase interpolator (libebur128 interp_create / interp_process) ------------------ */
typedef struct {
  unsigned factor, delay;
  unsigned count[4];
  unsigned index[4][INTERP_MAX_DELAY];
  float coeff[4][INTERP_MAX_DELAY]; /* ebur128 0.1.x keeps taps and accumulates in f32 (dasp f32 frames) */
  float z[ORC_MAX_CHANNELS][INTERP_MAX_DELAY];
  unsigned zi;
} interp_t;

static void interp_init(interp_t* it, unsigned factor) {
  memset(it, 0, sizeof(*it));
  it->factor = factor;
  it->delay = (INTERP_TAPS + factor - 1) / factor;
  for (unsigned j = 0; j < INTERP_TAPS; j++) {
    double m = (double)j - (double)(INTERP_TAPS - 1) / 2.0;
    double c = 1.0;
    if (fabs(m) > ALMOST_ZERO) c = sin(m * M_PI / factor) / (m * M_PI / factor);
    c *= 0.5 * (1 - cos(2 * M_PI * j / (INTERP_TAPS - 1)));
    if (fabs(c) > ALMOST_ZERO) {
      unsigned f = j % factor;
      unsigned t = it->count[f]++;
      it->coeff[f][t] = (float)c;
      it->index[f][t] = j / factor;
    }
  }
}

size_t orc_interp_taps(uint32_t rate, unsigned* factor, unsigned counts[4]) {
  interp_t it;
  unsigned f = rate < 96000 ? 4 : (rate < 192000 ? 2 : 0);
  *factor = f;
  memset(counts, 0, 4 * sizeof(unsigned));
  if (!f) return 0;
  interp_init(&it, f);
  size_t total = 0;
  for (unsigned i = 0; i < f; i++) { counts[i] = it.count[i]; total += it.count[i]; }
  return total;
}

/* ---- meter state ------------------------------------------------------------------------- */
struct orc_ebur128 {
  uint32_t channels, rate;
  int mode;
  int channel_map[ORC_MAX_CHANNELS];
  size_t s100;              /* samples_in_100ms = (rate + 5) / 10 */
  double* audio_data;       /* ring of K-weighted samples, interleaved */
  size_t audio_data_frames; /* ring length in frames (3 s rounded up to a multiple of s100) */
  size_t audio_data_index;  /* write position in SAMPLES (frames*channels) */
  size_t needed_frames;
  size_t short_term_frame_counter;
  double b[5], a[5];
  double v[ORC_MAX_CHANNELS][5];
  uint64_t block_hist[1000];
  uint64_t st_hist[1000];
  double sample_peak[ORC_MAX_CHANNELS], prev_sample_peak[ORC_MAX_CHANNELS];
  double true_peak[ORC_MAX_CHANNELS], prev_true_peak[ORC_MAX_CHANNELS];
  int has_interp;
  interp_t interp;
};

static double g_hist_energies[1000];
static double g_hist_boundaries[1001];
static int g_hist_ready = 0;

static void hist_init(void) {
  if (g_hist_ready) return;
  g_hist_boundaries[0] = pow(10.0, (-70.0 + 0.691) / 10.0);
  for (int i = 0; i < 1000; i++) g_hist_energies[i] = pow(10.0, ((double)i / 10.0 - 69.95 + 0.691) / 10.0);
  for (int i = 1; i < 1001; i++) g_hist_boundaries[i] = pow(10.0, ((double)i / 10.0 - 70.0 + 0.691) / 10.0);
  g_hist_ready = 1;
}
double orc_histogram_energy(unsigned i) { hist_init(); return g_hist_energies[i]; }
double orc_histogram_boundary(unsigned i) { hist_init(); return g_hist_boundaries[i]; }

static size_t find_histogram_index(double energy) {
  size_t lo = 0, hi = 1000, mid;
  do {
    mid = (lo + hi) / 2;
    if (energy >= g_hist_boundaries[mid]) lo = mid; else hi = mid;
  } while (hi - lo != 1);
  return lo;
}

/* exported for the bin-edge tests (golden set G5): the crate's bisection on the 1001 boundaries */
size_t orc_find_histogram_index(double energy) { hist_init(); return find_histogram_index(energy); }

static double energy_to_loudness(double e) { return 10.0 * log10(e) - 0.691; }

/* libebur128 ebur128_init_filter: BS.1770 pre-filter (shelf) x RLB high-pass, convolved. */
static void init_filter(orc_ebur128* st) {
  double f0 = 1681.974450955533, G = 3.999843853973347, Q = 0.7071752369554196;
  double K = tan(M_PI * f0 / (double)st->rate);
  double Vh = pow(10.0, G / 20.0);
  double Vb = pow(Vh, 0.4996667741545416);
  double pb[3], pa[3] = {1.0, 0.0, 0.0}, rb[3] = {1.0, -2.0, 1.0}, ra[3] = {1.0, 0.0, 0.0};
  double a0 = 1.0 + K / Q + K * K;
  pb[0] = (Vh + Vb * K / Q + K * K) / a0;
  pb[1] = 2.0 * (K * K - Vh) / a0;
  pb[2] = (Vh - Vb * K / Q + K * K) / a0;
  pa[1] = 2.0 * (K * K - 1.0) / a0;
  pa[2] = (1.0 - K / Q + K * K) / a0;
  f0 = 38.13547087602444;
  Q = 0.5003270373238773;
  K = tan(M_PI * f0 / (double)st->rate);
  ra[1] = 2.0 * (K * K - 1.0) / (1.0 + K / Q + K * K);
  ra[2] = (1.0 - K / Q + K * K) / (1.0 + K / Q + K * K);
  st->b[0] = pb[0] * rb[0];
  st->b[1] = pb[0] * rb[1] + pb[1] * rb[0];
  st->b[2] = pb[0] * rb[2] + pb[1] * rb[1] + pb[2] * rb[0];
  st->b[3] = pb[1] * rb[2] + pb[2] * rb[1];
  st->b[4] = pb[2] * rb[2];
  st->a[0] = pa[0] * ra[0];
  st->a[1] = pa[0] * ra[1] + pa[1] * ra[0];
  st->a[2] = pa[0] * ra[2] + pa[1] * ra[1] + pa[2] * ra[0];
  st->a[3] = pa[1] * ra[2] + pa[2] * ra[1];
  st->a[4] = pa[2] * ra[2];
}

void orc_ebur128_coeffs(const orc_ebur128* st, double b[5], double a[5]) {
  memcpy(b, st->b, sizeof(st->b));
  memcpy(a, st->a, sizeof(st->a));
}

static void default_channel_map(orc_ebur128* st) {
  uint32_t n = st->channels;
  if (n == 4) {
    st->channel_map[0] = CH_LEFT; st->channel_map[1] = CH_RIGHT;
    st->channel_map[2] = CH_LEFT_SURROUND; st->channel_map[3] = CH_RIGHT_SURROUND;
  } else if (n == 5) {
    st->channel_map[0] = CH_LEFT; st->channel_map[1] = CH_RIGHT; st->channel_map[2] = CH_CENTER;
    st->channel_map[3] = CH_LEFT_SURROUND; st->channel_map[4] = CH_RIGHT_SURROUND;
  } else {
    for (uint32_t i = 0; i < n; i++) {
      switch (i) {
        case 0: st->channel_map[i] = CH_LEFT; break;
        case 1: st->channel_map[i] = CH_RIGHT; break;
        case 2: st->channel_map[i] = CH_CENTER; break;
        case 3: st->channel_map[i] = CH_UNUSED; break;
        case 4: st->channel_map[i] = CH_LEFT_SURROUND; break;
        case 5: st->channel_map[i] = CH_RIGHT_SURROUND; break;
        default: st->channel_map[i] = CH_UNUSED; break;
      }
    }
  }
}

static void meter_clear(orc_ebur128* st) {
  memset(st->audio_data, 0, st->audio_data_frames * st->channels * sizeof(double));
  st->audio_data_index = 0;
  st->needed_frames = st->s100 * 4;
  st->short_term_frame_counter = 0;
  memset(st->v, 0, sizeof(st->v));
  memset(st->block_hist, 0, sizeof(st->block_hist));
  memset(st->st_hist, 0, sizeof(st->st_hist));
  memset(st->sample_peak, 0, sizeof(st->sample_peak));
  memset(st->prev_sample_peak, 0, sizeof(st->prev_sample_peak));
  memset(st->true_peak, 0, sizeof(st->true_peak));
  memset(st->prev_true_peak, 0, sizeof(st->prev_true_peak));
  memset(st->interp.z, 0, sizeof(st->interp.z));
  st->interp.zi = 0;
}

static int meter_init(orc_ebur128* st, uint32_t channels, uint32_t rate, int mode) {
  if (channels == 0 || channels > ORC_MAX_CHANNELS) return ORC_ERR_NOMEM;
  if (rate < 16 || rate > 2822400) return ORC_ERR_NOMEM;
  hist_init();
  memset(st, 0, sizeof(*st));
  st->channels = channels;
  st->rate = rate;
  st->mode = mode;
  default_channel_map(st);
  st->s100 = (rate + 5) / 10;
  size_t window_ms;
  if ((mode & ORC_MODE_S) == ORC_MODE_S) window_ms = 3000;
  else if ((mode & ORC_MODE_M) == ORC_MODE_M) window_ms = 400;
  else return ORC_ERR_NOMEM;
  st->audio_data_frames = (size_t)rate * window_ms / 1000;
  if (st->audio_data_frames % st->s100)
    st->audio_data_frames = st->audio_data_frames + st->s100 - (st->audio_data_frames % st->s100);
  st->audio_data = (double*)malloc(st->audio_data_frames * channels * sizeof(double));
  if (!st->audio_data) return ORC_ERR_NOMEM;
  init_filter(st);
  st->has_interp = 0;
  if ((mode & ORC_MODE_TRUE_PEAK) == ORC_MODE_TRUE_PEAK) {
    if (rate < 96000) { interp_init(&st->interp, 4); st->has_interp = 1; }
    else if (rate < 192000) { interp_init(&st->interp, 2); st->has_interp = 1; }
  }
  meter_clear(st);
  return ORC_OK;
}

orc_ebur128* orc_ebur128_new(uint32_t channels, uint32_t rate, int mode, int* err) {
  orc_ebur128* st = (orc_ebur128*)malloc(sizeof(orc_ebur128));
  int rc = st ? meter_init(st, channels, rate, mode) : ORC_ERR_NOMEM;
  if (err) *err = rc;
  if (rc != ORC_OK) { free(st); return NULL; }
  return st;
}

void orc_ebur128_free(orc_ebur128* st) {
  if (!st) return;
  free(st->audio_data);
  free(st);
}

void orc_ebur128_reset(orc_ebur128* st) { meter_clear(st); }

/* libebur128 ebur128_calc_gating_block: channel-weighted mean square over the last
 * frames_per_block frames of the ring, summed in ring order. */
static double calc_gating_block(const orc_ebur128* st, size_t frames_per_block) {
  double sum = 0.0;
  const size_t C = st->channels;
  const size_t idx_frames = st->audio_data_index / C;
  for (size_t c = 0; c < C; c++) {
    if (st->channel_map[c] == CH_UNUSED) continue;
    double channel_sum = 0.0;
    if (st->audio_data_index < frames_per_block * C) {
      for (size_t i = 0; i < idx_frames; i++) {
        double y = st->audio_data[i * C + c];
        channel_sum += y * y;
      }
      for (size_t i = st->audio_data_frames - (frames_per_block - idx_frames); i < st->audio_data_frames; i++) {
        double y = st->audio_data[i * C + c];
        channel_sum += y * y;
      }
    } else {
      for (size_t i = idx_frames - frames_per_block; i < idx_frames; i++) {
        double y = st->audio_data[i * C + c];
        channel_sum += y * y;
      }
    }
    if (st->channel_map[c] == CH_LEFT_SURROUND || st->channel_map[c] == CH_RIGHT_SURROUND) channel_sum *= 1.41;
    sum += channel_sum;
  }
  return sum / (double)frames_per_block;
}

/* libebur128 EBUR128_FILTER macro body for f32 input (scaling factor 1.0). */
static void filter_f32(orc_ebur128* st, const float* src, size_t frames) {
  const size_t C = st->channels;
  double* audio_data = st->audio_data + st->audio_data_index;

  if ((st->mode & ORC_MODE_SAMPLE_PEAK) == ORC_MODE_SAMPLE_PEAK) {
    for (size_t c = 0; c < C; c++) {
      double max = 0.0;
      for (size_t i = 0; i < frames; i++) {
        double cur = fabs((double)src[i * C + c]);
        if (cur > max) max = cur;
      }
      if (max > st->prev_sample_peak[c]) st->prev_sample_peak[c] = max;
    }
  }
  if ((st->mode & ORC_MODE_TRUE_PEAK) == ORC_MODE_TRUE_PEAK && st->has_interp) {
    interp_t* it = &st->interp;
    for (size_t i = 0; i < frames; i++) {
      for (size_t c = 0; c < C; c++) {
        it->z[c][it->zi] = src[i * C + c];
        for (unsigned f = 0; f < it->factor; f++) {
          float acc = 0.0f;
          for (unsigned t = 0; t < it->count[f]; t++) {
            int k = (int)it->zi - (int)it->index[f][t];
            if (k < 0) k += (int)it->delay;
            acc += it->z[c][k] * it->coeff[f][t];
          }
          double val = fabs((double)acc);
          if (val > st->prev_true_peak[c]) st->prev_true_peak[c] = val;
        }
      }
      it->zi++;
      if (it->zi == it->delay) it->zi = 0;
    }
  }
  for (size_t c = 0; c < C; c++) {
    if (st->channel_map[c] == CH_UNUSED) continue;
    double* v = st->v[c];
    const double* a = st->a;
    const double* b = st->b;
    for (size_t i = 0; i < frames; i++) {
      v[0] = (double)src[i * C + c] - a[1] * v[1] - a[2] * v[2] - a[3] * v[3] - a[4] * v[4];
      audio_data[i * C + c] = b[0] * v[0] + b[1] * v[1] + b[2] * v[2] + b[3] * v[3] + b[4] * v[4];
      v[4] = v[3];
      v[3] = v[2];
      v[2] = v[1];
      v[1] = v[0];
    }
    for (int j = 1; j <= 4; j++) v[j] = fabs(v[j]) < DBL_MIN ? 0.0 : v[j];
  }
}

int orc_ebur128_add_frames_f32(orc_ebur128* st, const float* src, size_t n_samples) {
  const size_t C = st->channels;
  if (n_samples % C != 0) return ORC_ERR_NOMEM; /* ebur128 0.1.10 Interleaved::new -> Error::NoMem */
  size_t frames = n_samples / C;
  size_t src_index = 0;
  for (size_t c = 0; c < C; c++) { st->prev_sample_peak[c] = 0.0; st->prev_true_peak[c] = 0.0; }
  while (frames > 0) {
    if (frames >= st->needed_frames) {
      filter_f32(st, src + src_index, st->needed_frames);
      src_index += st->needed_frames * C;
      frames -= st->needed_frames;
      st->audio_data_index += st->needed_frames * C;
      if ((st->mode & ORC_MODE_I) == ORC_MODE_I) {
        double e = calc_gating_block(st, st->s100 * 4);
        if (e >= g_hist_boundaries[0]) ++st->block_hist[find_histogram_index(e)];
      }
      if ((st->mode & ORC_MODE_LRA) == ORC_MODE_LRA) {
        st->short_term_frame_counter += st->needed_frames;
        if (st->short_term_frame_counter == st->s100 * 30) {
          double e = calc_gating_block(st, st->s100 * 30);
          if (e >= g_hist_boundaries[0]) ++st->st_hist[find_histogram_index(e)];
          st->short_term_frame_counter = st->s100 * 20;
        }
      }
      st->needed_frames = st->s100;
      if (st->audio_data_index == st->audio_data_frames * C) st->audio_data_index = 0;
    } else {
      filter_f32(st, src + src_index, frames);
      st->audio_data_index += frames * C;
      if ((st->mode & ORC_MODE_LRA) == ORC_MODE_LRA) st->short_term_frame_counter += frames;
      st->needed_frames -= frames;
      frames = 0;
    }
  }
  for (size_t c = 0; c < C; c++) {
    if (st->prev_sample_peak[c] > st->sample_peak[c]) st->sample_peak[c] = st->prev_sample_peak[c];
    if (st->prev_true_peak[c] > st->true_peak[c]) st->true_peak[c] = st->prev_true_peak[c];
  }
  return ORC_OK;
}

static int energy_in_interval(const orc_ebur128* st, size_t interval_frames, double* out) {
  if (interval_frames > st->audio_data_frames) return ORC_ERR_INVALID_MODE;
  *out = calc_gating_block(st, interval_frames);
  return ORC_OK;
}

int orc_ebur128_loudness_momentary(orc_ebur128* st, double* out) {
  double e;
  int rc = energy_in_interval(st, st->s100 * 4, &e);
  if (rc) return rc;
  *out = e <= 0.0 ? -HUGE_VAL : energy_to_loudness(e);
  return ORC_OK;
}

int orc_ebur128_loudness_shortterm(orc_ebur128* st, double* out) {
  double e;
  if ((st->mode & ORC_MODE_S) != ORC_MODE_S) return ORC_ERR_INVALID_MODE;
  int rc = energy_in_interval(st, st->s100 * 30, &e);
  if (rc) return rc;
  *out = e <= 0.0 ? -HUGE_VAL : energy_to_loudness(e);
  return ORC_OK;
}

/* libebur128 ebur128_gated_loudness, histogram branch (Mode::all() contains HISTOGRAM). */
int orc_ebur128_loudness_global(orc_ebur128* st, double* out) {
  if ((st->mode & ORC_MODE_I) != ORC_MODE_I) return ORC_ERR_INVALID_MODE;
  double relative_threshold = 0.0, gated = 0.0;
  uint64_t above = 0;
  for (int i = 0; i < 1000; i++) {
    relative_threshold += (double)st->block_hist[i] * g_hist_energies[i];
    above += st->block_hist[i];
  }
  if (!above) { *out = -HUGE_VAL; return ORC_OK; }
  relative_threshold /= (double)above;
  relative_threshold *= pow(10.0, -10.0 / 10.0);
  size_t start_index;
  if (relative_threshold < g_hist_boundaries[0]) start_index = 0;
  else {
    start_index = find_histogram_index(relative_threshold);
    if (relative_threshold > g_hist_energies[start_index]) ++start_index;
  }
  above = 0;
  for (size_t j = start_index; j < 1000; j++) {
    gated += (double)st->block_hist[j] * g_hist_energies[j];
    above += st->block_hist[j];
  }
  if (!above) { *out = -HUGE_VAL; return ORC_OK; }
  gated /= (double)above;
  *out = energy_to_loudness(gated);
  return ORC_OK;
}

/* libebur128 ebur128_loudness_range_multiple, histogram branch (EBU Tech 3342). */
int orc_ebur128_loudness_range(orc_ebur128* st, double* out) {
  if ((st->mode & ORC_MODE_LRA) != ORC_MODE_LRA) return ORC_ERR_INVALID_MODE;
  uint64_t stl_size = 0;
  double stl_power = 0.0;
  for (int j = 0; j < 1000; j++) {
    stl_size += st->st_hist[j];
    stl_power += (double)st->st_hist[j] * g_hist_energies[j];
  }
  if (!stl_size) { *out = 0.0; return ORC_OK; }
  stl_power /= (double)stl_size;
  double stl_integrated = pow(10.0, -20.0 / 10.0) * stl_power;
  size_t index;
  if (stl_integrated < g_hist_boundaries[0]) index = 0;
  else {
    index = find_histogram_index(stl_integrated);
    if (stl_integrated > g_hist_energies[index]) ++index;
  }
  stl_size = 0;
  for (size_t j = index; j < 1000; j++) stl_size += st->st_hist[j];
  if (!stl_size) { *out = 0.0; return ORC_OK; }
  uint64_t percentile_low = (uint64_t)((double)(stl_size - 1) * 0.1 + 0.5);
  uint64_t percentile_high = (uint64_t)((double)(stl_size - 1) * 0.95 + 0.5);
  stl_size = 0;
  size_t j = index;
  while (stl_size <= percentile_low) stl_size += st->st_hist[j++];
  double l_en = g_hist_energies[j - 1];
  while (stl_size <= percentile_high) stl_size += st->st_hist[j++];
  double h_en = g_hist_energies[j - 1];
  *out = energy_to_loudness(h_en) - energy_to_loudness(l_en);
  return ORC_OK;
}

int orc_ebur128_true_peak(orc_ebur128* st, uint32_t ch, double* out) {
  if ((st->mode & ORC_MODE_TRUE_PEAK) != ORC_MODE_TRUE_PEAK) return ORC_ERR_INVALID_MODE;
  if (ch >= st->channels) return ORC_ERR_INVALID_CHANNEL_INDEX;
  *out = st->true_peak[ch] > st->sample_peak[ch] ? st->true_peak[ch] : st->sample_peak[ch];
  return ORC_OK;
}

int orc_ebur128_sample_peak(orc_ebur128* st, uint32_t ch, double* out) {
  if ((st->mode & ORC_MODE_SAMPLE_PEAK) != ORC_MODE_SAMPLE_PEAK) return ORC_ERR_INVALID_MODE;
  if (ch >= st->channels) return ORC_ERR_INVALID_CHANNEL_INDEX;
  *out = st->sample_peak[ch];
  return ORC_OK;
}

void orc_ebur128_histograms(const orc_ebur128* st, uint64_t block[1000], uint64_t shortterm[1000]) {
  memcpy(block, st->block_hist, sizeof(st->block_hist));
  memcpy(shortterm, st->st_hist, sizeof(st->st_hist));
}

/* Analyzer::calculate_integrated_lufs (reference src/analyzer.rs:170-182): a fresh meter with
 * Mode::all(), fed in chunks of sample_rate*2 samples; any error -> None. */
int orc_calculate_integrated_lufs(uint32_t channels, uint32_t sample_rate, const float* samples,
                                  size_t len, double* out) {
  int err;
  orc_ebur128* st = orc_ebur128_new(channels, sample_rate, ORC_MODE_ALL, &err);
  if (!st) return 0;
  size_t chunk = (size_t)sample_rate * 2;
  for (size_t off = 0; off < len; off += chunk) {
    size_t n = len - off < chunk ? len - off : chunk;
    if (orc_ebur128_add_frames_f32(st, samples + off, n) != ORC_OK) { orc_ebur128_free(st); return 0; }
  }
  int rc = orc_ebur128_loudness_global(st, out);
  orc_ebur128_free(st);
  return rc == ORC_OK;
}

/* ---- batch driver: the timed CPU baseline ------------------------------------------------ */
struct orc_batch {
  size_t n;
  uint32_t channels;
  orc_ebur128** st;
};

orc_batch* orc_batch_new(size_t n_streams, uint32_t channels, uint32_t rate, int mode) {
  return orc_batch_new_mt(n_streams, channels, rate, mode, 1);
}

/* `threads` > 1: meters are created (first-touched) by the thread that will later process them,
 * with the same static schedule, so each meter's ring lives on its worker's NUMA node. */
orc_batch* orc_batch_new_mt(size_t n_streams, uint32_t channels, uint32_t rate, int mode, int threads) {
  orc_batch* b = (orc_batch*)calloc(1, sizeof(*b));
  if (!b) return NULL;
  b->n = n_streams;
  b->channels = channels;
  b->st = (orc_ebur128**)calloc(n_streams, sizeof(orc_ebur128*));
  int bad = 0;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1) reduction(| : bad)
  for (long i = 0; i < (long)n_streams; i++) {
    int err;
    b->st[i] = orc_ebur128_new(channels, rate, mode, &err);
    if (!b->st[i]) bad |= 1;
  }
  if (bad) { orc_batch_free(b); return NULL; }
  return b;
}

void orc_batch_free(orc_batch* b) {
  if (!b) return;
  if (b->st) for (size_t i = 0; i < b->n; i++) orc_ebur128_free(b->st[i]);
  free(b->st);
  free(b);
}

int orc_batch_add_frames(orc_batch* b, const float* in, size_t frames, int threads) {
  const size_t per = frames * b->channels;
  int rc_all = 0;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1) reduction(| : rc_all)
  for (long i = 0; i < (long)b->n; i++) rc_all |= orc_ebur128_add_frames_f32(b->st[i], in + (size_t)i * per, per);
  return rc_all;
}

void orc_batch_histograms(orc_batch* b, size_t s, uint64_t block[1000], uint64_t shortterm[1000]) {
  orc_ebur128_histograms(b->st[s], block, shortterm);
}

void orc_batch_query(orc_batch* b, double* momentary, double* shortterm, double* global,
                     double* range, double* true_peak, int threads) {
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
  for (long i = 0; i < (long)b->n; i++) {
    orc_ebur128* st = b->st[i];
    if (momentary) orc_ebur128_loudness_momentary(st, &momentary[i]);
    if (shortterm && orc_ebur128_loudness_shortterm(st, &shortterm[i])) shortterm[i] = NAN;
    if (global && orc_ebur128_loudness_global(st, &global[i])) global[i] = NAN;
    if (range && orc_ebur128_loudness_range(st, &range[i])) range[i] = NAN;
    if (true_peak)
      for (uint32_t c = 0; c < b->channels; c++)
        if (orc_ebur128_true_peak(st, c, &true_peak[(size_t)i * b->channels + c])) true_peak[(size_t)i * b->channels + c] = NAN;
  }
}
