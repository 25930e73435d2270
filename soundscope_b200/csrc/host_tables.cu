// host_tables.cu — constants the kernels consume, computed once on the host in the arithmetic the
// reference's crates use.  Product code: does not include or call anything under oracle/.
//
//   kweight_coeffs           ebur128 0.1.10 filter init (= libebur128 ebur128_init_filter; BS.1770-4)
//   default_channel_weights  ebur128 default channel map + BS.1770 channel weights
//   truepeak_taps            ebur128 interp: 49-tap Hann-windowed sinc, |c| <= 1e-6 pruned, rate->factor rule
//   histogram_tables         ebur128 histogram energies / boundaries (0.1 LU bins from -70 LUFS)
//   hann_multipliers         spectrum-analyzer 1.7.0 windows::hann_window with libm 0.2.16 cosf (musl)
//   fft_bin_range, fft_axis  spectrum-analyzer frequency filter + reference src/analyzer.rs:67-102
#include <math.h>
#include <string.h>

#include "ssb_internal.cuh"

namespace ssb {

void kweight_coeffs(uint32_t rate, double b[5], double a[5]) {
  double f0 = 1681.974450955533, G = 3.999843853973347, Q = 0.7071752369554196;
  double K = tan(M_PI * f0 / (double)rate);
  double Vh = pow(10.0, G / 20.0);
  double Vb = pow(Vh, 0.4996667741545416);
  double a0 = 1.0 + K / Q + K * K;
  const double pb[3] = {(Vh + Vb * K / Q + K * K) / a0, 2.0 * (K * K - Vh) / a0, (Vh - Vb * K / Q + K * K) / a0};
  const double pa[3] = {1.0, 2.0 * (K * K - 1.0) / a0, (1.0 - K / Q + K * K) / a0};
  f0 = 38.13547087602444;
  Q = 0.5003270373238773;
  K = tan(M_PI * f0 / (double)rate);
  const double rb[3] = {1.0, -2.0, 1.0};
  const double ra[3] = {1.0, 2.0 * (K * K - 1.0) / (1.0 + K / Q + K * K), (1.0 - K / Q + K * K) / (1.0 + K / Q + K * K)};
  b[0] = pb[0] * rb[0];
  b[1] = pb[0] * rb[1] + pb[1] * rb[0];
  b[2] = pb[0] * rb[2] + pb[1] * rb[1] + pb[2] * rb[0];
  b[3] = pb[1] * rb[2] + pb[2] * rb[1];
  b[4] = pb[2] * rb[2];
  a[0] = pa[0] * ra[0];
  a[1] = pa[0] * ra[1] + pa[1] * ra[0];
  a[2] = pa[0] * ra[2] + pa[1] * ra[1] + pa[2] * ra[0];
  a[3] = pa[1] * ra[2] + pa[2] * ra[1];
  a[4] = pa[2] * ra[2];
}

void default_channel_weights(uint32_t channels, float w[kMaxChannels], uint64_t* active_mask) {
  // kinds: 0 unused, 1 L/R/C (weight 1.0), 2 surround (weight 1.41)
  int kind[kMaxChannels] = {0};
  if (channels == 4) {
    kind[0] = 1; kind[1] = 1; kind[2] = 2; kind[3] = 2;
  } else if (channels == 5) {
    kind[0] = 1; kind[1] = 1; kind[2] = 1; kind[3] = 2; kind[4] = 2;
  } else {
    for (uint32_t i = 0; i < channels; i++) kind[i] = (i <= 2) ? 1 : ((i == 4 || i == 5) ? 2 : 0);
  }
  uint64_t mask = 0;
  for (uint32_t i = 0; i < kMaxChannels; i++) {
    w[i] = 0.0f;
    if (i < channels && kind[i]) { w[i] = kind[i] == 2 ? 1.41f : 1.0f; mask |= (1ull << i); }
  }
  *active_mask = mask;
}

int truepeak_taps(uint32_t rate, float tp4[3][12], float tp2[24], int force_factor) {
  const int taps = 49;
  // ebur128's rate rule; force_factor (2 or 4) overrides it for the explicitly non-parity benchmark variant
  const int factor = force_factor ? force_factor : (rate < 96000 ? 4 : (rate < 192000 ? 2 : 0));
  memset(tp4, 0, sizeof(float) * 36);
  memset(tp2, 0, sizeof(float) * 24);
  if (!factor) return 0;
  for (int j = 0; j < taps; j++) {
    double m = (double)j - (double)(taps - 1) / 2.0;
    double c = 1.0;
    if (fabs(m) > 0.000001) c = sin(m * M_PI / factor) / (m * M_PI / factor);
    c *= 0.5 * (1 - cos(2 * M_PI * j / (taps - 1)));
    if (fabs(c) > 0.000001) {
      int f = j % factor, t = j / factor;
      // phase 0 keeps a single unit tap (a pure delay); its |output| is a sample magnitude, which
      // EbuR128::true_peak folds in through max(true_peak, sample_peak) — not stored here.
      if (f == 0) continue;
      if (factor == 4) tp4[f - 1][t] = (float)c; else tp2[t] = (float)c;
    }
  }
  return factor;
}

double histogram_bound0() {
  static const double b0 = pow(10.0, (-70.0 + 0.691) / 10.0);
  return b0;
}

void histogram_tables(double energies[1000], double boundaries[1001]) {
  boundaries[0] = histogram_bound0();
  for (int i = 0; i < 1000; i++) energies[i] = pow(10.0, ((double)i / 10.0 - 69.95 + 0.691) / 10.0);
  for (int i = 1; i < 1001; i++) boundaries[i] = pow(10.0, ((double)i / 10.0 - 70.0 + 0.691) / 10.0);
}

// ---- libm 0.2.16 cosf: the kernels of musl's cosf evaluated in double, rounded once -----------
static float cosdf(double x) {
  const double C0 = -0x1ffffffd0c5e81.0p-54, C1 = 0x155553e1053a42.0p-57, C2 = -0x16c087e80f1e27.0p-62,
               C3 = 0x199342e0ee5069.0p-68;
  double z = x * x, w = z * z, r = C2 + z * C3;
  return (float)(((1.0 + z * C0) + w * C1) + (w * z) * r);
}
static float sindf(double x) {
  const double S1 = -0x15555554cbac77.0p-55, S2 = 0x111110896efbb2.0p-59, S3 = -0x1a00f9e2cae774.0p-65,
               S4 = 0x16cd878c3b46a7.0p-71;
  double z = x * x, w = z * z, r = S3 + z * S4, s = z * x;
  return (float)((x + s * (S1 + z * S2)) + s * w * r);
}
float libm_cosf(float x) {
  const double pio2 = M_PI_2;
  uint32_t ix;
  memcpy(&ix, &x, 4);
  const bool neg = ix >> 31;
  ix &= 0x7fffffff;
  if (ix <= 0x3f490fda) return ix < 0x39800000 ? 1.0f : cosdf(x);
  if (ix <= 0x407b53d1) {
    if (ix > 0x4016cbe3) return -cosdf(neg ? x + 2 * pio2 : x - 2 * pio2);
    return neg ? sindf(x + pio2) : sindf(pio2 - x);
  }
  if (ix <= 0x40e231d5) {
    if (ix > 0x40afeddf) return cosdf(neg ? x + 4 * pio2 : x - 4 * pio2);
    return neg ? sindf(-x - 3 * pio2) : sindf(x - 3 * pio2);
  }
  return (float)cos((double)x);  // not reached by the window (argument stays in [0, 2*pi))
}

void hann_multipliers(size_t n, std::vector<float>& w) {
  w.resize(n);
  const float pi = 3.14159274101257324f;
  const float nf = (float)n;
  for (size_t i = 0; i < n; i++) {
    volatile float two_pi_i = 2.0f * pi * (float)i;   // volatile: one f32 rounding per step, no contraction
    volatile float arg = two_pi_i / nf;
    volatile float c = libm_cosf(arg);
    volatile float one_minus = 1.0f - c;
    w[i] = 0.5f * one_minus;
  }
}

size_t fft_bin_range(size_t n, uint32_t rate, size_t* k_first) {
  const float res = (float)rate / (float)(uint32_t)n;
  size_t cnt = 0, first = 0;
  for (size_t k = 0; k <= n / 2; k++) {
    volatile float fr = (float)k * res;
    if (fr >= 20.0f && fr <= 20000.0f) { if (!cnt) first = k; cnt++; }
  }
  if (k_first) *k_first = first;
  return cnt;
}

void fft_axis(size_t n, uint32_t rate, std::vector<double>& x, std::vector<double>& tilt, size_t* k_first) {
  size_t k0 = 0;
  const size_t cnt = fft_bin_range(n, rate, &k0);
  const float res = (float)rate / (float)(uint32_t)n;
  const double min_freq_log = log10(20.0), max_freq_log = log10(20000.0);
  const double log_range = max_freq_log - min_freq_log;
  x.resize(cnt);
  tilt.resize(cnt);
  for (size_t i = 0; i < cnt; i++) {
    volatile float fr = (float)(k0 + i) * res;
    const double freq = (double)fr;
    tilt[i] = 10.0 * log10(freq / 1000.0);
    x[i] = (log10(freq) - min_freq_log) / log_range * 100.0;
  }
  if (k_first) *k_first = k0;
}

}  // namespace ssb
