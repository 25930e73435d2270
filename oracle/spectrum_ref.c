/*
 * spectrum_ref.c — CPU restatement of Analyzer::get_fft, get_waveform and
 * get_mid_and_side_samples.  TEST INFRASTRUCTURE — see oracle.h.
 *
 * First-party arithmetic (pinned by the reference's own unit-test inputs):
 *   scale_to_dbfs                reference src/analyzer.rs:11-27
 *   pink tilt + log-frequency x  reference src/analyzer.rs:67-102
 *   get_waveform                 reference src/analyzer.rs:107-137
 *   get_mid_and_side_samples     reference src/audio_player.rs:400-419
 * Third-party arithmetic (un-vendored crates; PARITY UNPINNED, restated from the published
 * algorithms and checked against an f64 FFT in tests/):
 *   hann_window, samples_fft_to_spectrum   spectrum-analyzer 1.7.0 (Cargo.lock:1941-1944)
 *   rfft (radix-2 DIT, N/2-point complex + real recombination)  microfft 0.6.0 (Cargo.lock:1074-1077)
 *   cosf, sqrtf                            libm 0.2.16 = musl port (Cargo.lock:968-971)
 *
 * Build with -ffp-contract=off (Rust never fuses a*b+c).
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- libm 0.2.16 cosf (port of musl src/math/cosf.c, __cosdf.c, __sindf.c) --------------- */
static float k_cosdf(double x) {
  static const double C0 = -0x1ffffffd0c5e81.0p-54, C1 = 0x155553e1053a42.0p-57,
                      C2 = -0x16c087e80f1e27.0p-62, C3 = 0x199342e0ee5069.0p-68;
  double z = x * x;
  double w = z * z;
  double r = C2 + z * C3;
  return (float)(((1.0 + z * C0) + w * C1) + (w * z) * r);
}

static float k_sindf(double x) {
  static const double S1 = -0x15555554cbac77.0p-55, S2 = 0x111110896efbb2.0p-59,
                      S3 = -0x1a00f9e2cae774.0p-65, S4 = 0x16cd878c3b46a7.0p-71;
  double z = x * x;
  double w = z * z;
  double r = S3 + z * S4;
  double s = z * x;
  return (float)((x + s * (S1 + z * S2)) + s * w * r);
}

float orc_cosf(float x) {
  static const double c1pio2 = 1 * M_PI_2, c2pio2 = 2 * M_PI_2, c3pio2 = 3 * M_PI_2, c4pio2 = 4 * M_PI_2;
  uint32_t ix;
  memcpy(&ix, &x, 4);
  unsigned sign = ix >> 31;
  ix &= 0x7fffffff;
  if (ix <= 0x3f490fda) { /* |x| ~<= pi/4 */
    if (ix < 0x39800000) return 1.0f; /* |x| < 2**-12 */
    return k_cosdf(x);
  }
  if (ix <= 0x407b53d1) { /* |x| ~<= 5*pi/4 */
    if (ix > 0x4016cbe3) return -k_cosdf(sign ? x + c2pio2 : x - c2pio2);
    return sign ? k_sindf(x + c1pio2) : k_sindf(c1pio2 - x);
  }
  if (ix <= 0x40e231d5) { /* |x| ~<= 9*pi/4 */
    if (ix > 0x40afeddf) return k_cosdf(sign ? x + c4pio2 : x - c4pio2);
    return sign ? k_sindf(-x - c3pio2) : k_sindf(x - c3pio2);
  }
  /* the Hann window never leaves [0, 2*pi); larger arguments are outside the restated path */
  return (float)cos((double)x);
}

/* spectrum-analyzer 1.7.0 windows::hann_window: periodic Hann, all f32. */
void orc_hann_window(const float* in, size_t n, float* out) {
  const float pi = 3.14159274101257324f; /* core::f32::consts::PI */
  const float n_f32 = (float)n;
  for (size_t i = 0; i < n; i++) {
    float two_pi_i = 2.0f * pi * (float)i;
    float c = orc_cosf(two_pi_i / n_f32);
    float multiplier = 0.5f * (1.0f - c);
    out[i] = multiplier * in[i];
  }
}

/* ---- microfft 0.6.0 style real FFT: N/2-point complex radix-2 DIT + recombination -------- */
typedef struct { float re, im; } cpx;

static void twiddle(size_t k, size_t n, float* c, float* s) {
  /* exp(-j*2*pi*k/n) from a correctly rounded f32 table; exact at the quadrant points */
  if (k == 0) { *c = 1.0f; *s = 0.0f; return; }
  if (4 * k == n) { *c = 0.0f; *s = -1.0f; return; }
  if (2 * k == n) { *c = -1.0f; *s = 0.0f; return; }
  double ang = 2.0 * M_PI * (double)k / (double)n;
  *c = (float)cos(ang);
  *s = (float)-sin(ang);
}

static void cfft_radix2(cpx* z, size_t m) {
  /* bit reversal */
  for (size_t i = 1, j = 0; i < m; i++) {
    size_t bit = m >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { cpx t = z[i]; z[i] = z[j]; z[j] = t; }
  }
  for (size_t len = 2; len <= m; len <<= 1) {
    size_t half = len >> 1;
    for (size_t j = 0; j < half; j++) {
      float wr, wi;
      twiddle(j, len, &wr, &wi);
      for (size_t i = j; i < m; i += len) {
        cpx a = z[i], b = z[i + half];
        cpx t;
        if (j == 0) t = b;
        else if (4 * j == len) { t.re = b.im; t.im = -b.re; }
        else { t.re = b.re * wr - b.im * wi; t.im = b.re * wi + b.im * wr; }
        z[i].re = a.re + t.re; z[i].im = a.im + t.im;
        z[i + half].re = a.re - t.re; z[i + half].im = a.im - t.im;
      }
    }
  }
}

/* |X[k]| for k = 0..n/2 of the real FFT of `windowed` (length n, power of two >= 2). */
int orc_rfft_mag(const float* windowed, size_t n, float* mag) {
  if (n < 2 || (n & (n - 1))) return ORC_FFT_NOT_POW2;
  size_t m = n / 2;
  cpx* z = (cpx*)malloc(sizeof(cpx) * (m + 1));
  if (!z) return ORC_ERR_NOMEM;
  for (size_t i = 0; i < m; i++) { z[i].re = windowed[2 * i]; z[i].im = windowed[2 * i + 1]; }
  if (m > 1) cfft_radix2(z, m);
  /* recombine: X[k] = (Z[k]+conj(Z[m-k]))/2 - j*w^k*(Z[k]-conj(Z[m-k]))/2, w = exp(-j*2*pi/n) */
  float dc = z[0].re + z[0].im, ny = z[0].re - z[0].im;
  mag[0] = sqrtf(dc * dc + 0.0f * 0.0f);
  mag[m] = sqrtf(ny * ny + 0.0f * 0.0f);
  for (size_t k = 1; k < m; k++) {
    cpx a = z[k], b = z[m - k];
    float sr = 0.5f * (a.re + b.re), si = 0.5f * (a.im - b.im);   /* (Z[k]+conj(Z[m-k]))/2 */
    float dr = 0.5f * (a.re - b.re), di = 0.5f * (a.im + b.im);   /* (Z[k]-conj(Z[m-k]))/2 */
    float wr, wi;
    twiddle(k, n, &wr, &wi);
    /* -j*w*d = -j*(wr+j*wi)*(dr+j*di) = (wr*di + wi*dr) + j*(wi*di - wr*dr) */
    float xr = sr + (wr * di + wi * dr);
    float xi = si + (wi * di - wr * dr);
    mag[k] = sqrtf(xr * xr + xi * xi); /* spectrum-analyzer complex_to_magnitude, libm sqrtf */
  }
  free(z);
  return ORC_FFT_OK;
}

/* reference src/analyzer.rs:11-27 */
float orc_scale_to_dbfs(float val, float n) {
  if (val == 0.0f) return -150.0f;
  float scaled = val * 4.0f / n;
  return 20.0f * log10f(scaled / 1.0f);
}

size_t orc_fft_bin_range(size_t n, uint32_t sample_rate, size_t* k_first) {
  float res = (float)sample_rate / (float)(uint32_t)n; /* fft_calc_frequency_resolution */
  size_t count = 0, first = 0;
  for (size_t k = 0; k <= n / 2; k++) {
    float fr = (float)k * res;
    if (fr >= 20.0f && fr <= 20000.0f) { if (!count) first = k; count++; }
  }
  if (k_first) *k_first = first;
  return count;
}

/* Analyzer::get_fft — reference src/analyzer.rs:55-105 */
int orc_get_fft(const float* samples, size_t n, uint32_t sample_rate, double* xy_out, size_t cap,
                size_t* n_out) {
  *n_out = 0;
  float* w = (float*)malloc(sizeof(float) * (n ? n : 1));
  if (!w) return ORC_ERR_NOMEM;
  orc_hann_window(samples, n, w);                       /* analyzer.rs:57 */
  /* samples_fft_to_spectrum's argument checks, in its order */
  int rc = ORC_FFT_OK;
  if (n < 2) rc = ORC_FFT_TOO_FEW_SAMPLES;
  if (!rc) for (size_t i = 0; i < n; i++) if (isnan(w[i])) { rc = ORC_FFT_NAN; break; }
  if (!rc) for (size_t i = 0; i < n; i++) if (isinf(w[i])) { rc = ORC_FFT_INF; break; }
  if (!rc && (n & (n - 1))) rc = ORC_FFT_NOT_POW2;
  if (!rc && n > 32768) rc = ORC_FFT_NOT_POW2;          /* crate's largest microfft size */
  if (!rc && 20000.0f > (float)sample_rate / 2.0f) rc = ORC_FFT_BAD_LIMIT; /* Range(20,20000).verify */
  if (rc) { free(w); return rc; }
  float* mag = (float*)malloc(sizeof(float) * (n / 2 + 1));
  if (!mag) { free(w); return ORC_ERR_NOMEM; }
  rc = orc_rfft_mag(w, n, mag);
  free(w);
  if (rc) { free(mag); return rc; }
  const float res = (float)sample_rate / (float)(uint32_t)n;
  const double min_freq_log = log10(20.0), max_freq_log = log10(20000.0);
  const double log_range = max_freq_log - min_freq_log;
  size_t cnt = 0;
  for (size_t k = 0; k <= n / 2; k++) {
    float fr = (float)k * res;
    if (!(fr >= 20.0f && fr <= 20000.0f)) continue;
    float db = orc_scale_to_dbfs(mag[k], (float)(uint32_t)n);
    if (!isfinite(db)) { free(mag); return ORC_FFT_SCALING; }
    double freq = (double)fr, val = (double)db;
    double compensation = 10.0 * log10(freq / 1000.0);   /* analyzer.rs:82 */
    double y = val + compensation;
    double x = (log10(freq) - min_freq_log) / log_range * 100.0; /* analyzer.rs:96-98 */
    if (cnt < cap) { xy_out[2 * cnt] = x; xy_out[2 * cnt + 1] = y; }
    cnt++;
  }
  free(mag);
  *n_out = cnt;
  return ORC_FFT_OK;
}

/* Analyzer::get_waveform — reference src/analyzer.rs:107-137.  f32::min/max ignore NaN. */
size_t orc_get_waveform(const float* samples, size_t len, double waveform_window, double* xy_out,
                        size_t cap) {
  double wtmp = waveform_window * 1000.;
  size_t window; /* Rust `as usize`: saturating, NaN -> 0 */
  if (!(wtmp > 0.0)) window = 0;
  else if (wtmp >= 18446744073709551615.0) window = SIZE_MAX;
  else window = (size_t)wtmp;
  double spp = (double)len / (double)window;
  size_t cnt = 0;
  for (size_t i = 0; i < window; i++) {
    double s = (double)i * spp;
    size_t start = !(s > 0.0) ? 0 : (s >= 18446744073709551615.0 ? SIZE_MAX : (size_t)s);
    double e = ceil((double)(i + 1) * spp);
    size_t end = !(e > 0.0) ? 0 : (e >= 18446744073709551615.0 ? SIZE_MAX : (size_t)e);
    if (end > len) end = len;
    if (start >= len) break;
    float mn = 0.0f, mx = 0.0f;
    if (end > start) {
      mn = samples[start]; mx = samples[start];
      for (size_t j = start + 1; j < end; j++) { mn = fminf(mn, samples[j]); mx = fmaxf(mx, samples[j]); }
    }
    if (cnt + 2 <= cap) {
      xy_out[2 * cnt] = (double)i; xy_out[2 * cnt + 1] = (double)mn;
      xy_out[2 * cnt + 2] = (double)i; xy_out[2 * cnt + 3] = (double)mx;
    }
    cnt += 2;
  }
  return cnt;
}

/* get_mid_and_side_samples — reference src/audio_player.rs:400-419 */
size_t orc_mid_side(const float* s, size_t len, float* mid, float* side) {
  size_t frames = len / 2; /* zip(left, right): odd tail sample dropped */
  for (size_t i = 0; i < frames; i++) {
    float l = s[2 * i], r = s[2 * i + 1];
    mid[i] = (l + r) / 2.f;
    side[i] = (l - r) / 2.f;
  }
  return frames;
}
