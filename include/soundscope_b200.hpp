// soundscope_b200.hpp — header-only C++ mirror of the reference's `analyzer::Analyzer`
// (reference src/analyzer.rs:29-183) over the C ABI in soundscope_b200.h.  Same method names, argument
// meaning and error behaviour; `Result<_, E>` becomes an exception carrying the reference's error name.
// (The reference is Rust; no Rust toolchain exists in this image, so the compiled-language host side is C++.
// The Rust shim is in INTEGRATION.md.)
#pragma once

#include <cmath>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "soundscope_b200.h"

namespace soundscope {

struct Error : std::runtime_error {
  int32_t code;
  Error(int32_t c, const std::string& what) : std::runtime_error(name(c) + ": " + what), code(c) {}
  static std::string name(int32_t c) {
    switch (c) {
      case SSB_ERR_NOMEM: return "NoMem";
      case SSB_ERR_INVALID_MODE: return "InvalidMode";
      case SSB_ERR_INVALID_CHANNEL_INDEX: return "InvalidChannelIndex";
      case SSB_ERR_FFT_TOO_FEW_SAMPLES: return "TooFewSamples";
      case SSB_ERR_FFT_NAN: return "NaNValuesNotSupported";
      case SSB_ERR_FFT_INF: return "InfinityValuesNotSupported";
      case SSB_ERR_FFT_NOT_POW2: return "SamplesLengthNotAPowerOfTwo";
      case SSB_ERR_FFT_BAD_LIMIT: return "InvalidFrequencyLimit";
      case SSB_ERR_FFT_SCALING: return "ScalingError";
      case SSB_ERR_UNALIGNED_QUERY: return "UnalignedQuery";
      case SSB_ERR_NO_DEVICE: return "NoDevice";
      default: return "status " + std::to_string(c);
    }
  }
};

class Analyzer {
 public:
  // Default (analyzer.rs:34-45): EbuR128::new(2, 44100, Mode::all()); the reference panics on failure
  explicit Analyzer(int device = -1) {
    const int32_t rc = ssb_analyzer_create(&h_, 2, 44100, SSB_MODE_ALL, 1, device, SSB_FLAG_RING);
    if (rc) throw Error(rc, "Failed to create loudness meter");
  }
  ~Analyzer() { ssb_analyzer_destroy(h_); }
  Analyzer(const Analyzer&) = delete;
  Analyzer& operator=(const Analyzer&) = delete;

  void create_loudness_meter(uint32_t channels, uint32_t rate) { check(ssb_create_loudness_meter(h_, channels, rate)); }  // :49-53

  std::vector<std::pair<double, double>> get_fft(const std::vector<float>& samples) const {  // :55-105
    std::vector<std::pair<double, double>> out(samples.size() / 2 + 1);
    size_t n = 0;
    check(ssb_get_fft(h_, samples.data(), samples.size(), reinterpret_cast<double*>(out.data()), out.size(), &n));
    out.resize(n);
    return out;
  }

  std::vector<std::pair<double, double>> get_waveform(const std::vector<float>& samples, double waveform_window) const {  // :107-137
    size_t n = 0;
    int32_t rc = ssb_get_waveform(h_, samples.data(), samples.size(), waveform_window, nullptr, 0, &n);
    if (rc != SSB_OK && rc != SSB_ERR_CAPACITY) check(rc);
    std::vector<std::pair<double, double>> out(n);
    check(ssb_get_waveform(h_, samples.data(), samples.size(), waveform_window, reinterpret_cast<double*>(out.data()), n, &n));
    return out;
  }

  void add_samples(const std::vector<float>& samples) { check(ssb_add_samples(h_, samples.data(), samples.size())); }  // :139-141
  void reset() { check(ssb_reset(h_)); }                                                                                // :143-145
  double get_shortterm_lufs() { double v; check(ssb_loudness_shortterm(h_, &v)); return v; }                            // :147-149
  double get_integrated_lufs() { double v; check(ssb_loudness_global(h_, &v)); return v; }                              // :151-153
  double get_loudness_range() { double v; check(ssb_loudness_range(h_, &v)); return v; }                                // :155-157
  std::pair<double, double> get_true_peak() { double l, r; check(ssb_get_true_peak(h_, &l, &r)); return {l, r}; }       // :159-164
  uint32_t sample_rate() const { return ssb_sample_rate(h_); }                                                          // :166-168

  std::optional<double> calculate_integrated_lufs(uint32_t channels, const std::vector<float>& samples) {               // :170-182
    double v = 0;
    int32_t some = 0;
    if (ssb_calculate_integrated_lufs(h_, channels, samples.data(), samples.size(), &v, &some) != SSB_OK || !some) return std::nullopt;
    return v;
  }

  // get_mid_and_side_samples (audio_player.rs:400-419)
  std::pair<std::vector<float>, std::vector<float>> get_mid_and_side_samples(const std::vector<float>& samples) const {
    std::vector<float> mid(samples.size() / 2), side(samples.size() / 2);
    size_t frames = 0;
    check(ssb_mid_side(h_, samples.data(), samples.size(), mid.data(), side.data(), &frames));
    return {std::move(mid), std::move(side)};
  }

  // One player tick, tui.rs:1482-1552 (analyze_audio_file_samples): `tail` = the last n_fft stereo frames
  struct Tick {
    std::vector<std::pair<double, double>> mid_fft, side_fft;   // empty when get_fft would return Err
    std::vector<std::pair<double, double>> waveform;            // microphone tick only
    double shortterm_lufs = 0;
    int32_t fft_status = 0, lufs_status = 0;
  };
  Tick process_tick(const std::vector<float>& tail, size_t lufs_samples = 16384) {
    Tick t;
    const size_t n_fft = tail.size() / 2, cap = n_fft / 2 + 1;
    t.mid_fft.resize(cap);
    t.side_fft.resize(cap);
    size_t n = 0;
    check(ssb_process_tick(h_, tail.data(), n_fft, lufs_samples, reinterpret_cast<double*>(t.mid_fft.data()),
                           reinterpret_cast<double*>(t.side_fft.data()), cap, &n, &t.shortterm_lufs, &t.fft_status,
                           &t.lufs_status));
    t.mid_fft.resize(t.fft_status ? 0 : n);
    t.side_fft.resize(t.fft_status ? 0 : n);
    return t;
  }

  // One microphone tick, tui.rs:1427-1480 (analyze_microphone_input), on one snapshot of the capture ring
  Tick analyze_microphone_input(ssb_capture_ring* ring, size_t n_fft = 16384, size_t lufs_samples = 16384,
                                double waveform_window = 15.0) {
    Tick t;
    const size_t cap = n_fft / 2 + 1;
    const double w = waveform_window * 1000.0;
    const size_t wave_cap = 2 * (w > 0 ? (size_t)w : 0) + 2;
    t.mid_fft.resize(cap);
    t.side_fft.resize(cap);
    t.waveform.resize(wave_cap);
    size_t n = 0, nw = 0;
    check(ssb_mic_tick(h_, ring, n_fft, lufs_samples, waveform_window, reinterpret_cast<double*>(t.mid_fft.data()),
                       reinterpret_cast<double*>(t.side_fft.data()), cap, &n, reinterpret_cast<double*>(t.waveform.data()),
                       wave_cap, &nw, &t.shortterm_lufs, &t.fft_status, &t.lufs_status));
    t.mid_fft.resize(t.fft_status ? 0 : n);
    t.side_fft.resize(t.fft_status ? 0 : n);
    t.waveform.resize(nw);
    return t;
  }

  // AudioFile::decode_file's sample conversion for WAV / AIFF PCM (audio_player.rs:248): raw interleaved PCM -> f32
  std::vector<float> pcm_to_f32(const void* pcm, size_t n_samples, int32_t format) const {
    std::vector<float> out(n_samples);
    check(ssb_pcm_to_f32(h_, pcm, n_samples, format, out.data()));
    return out;
  }

  ssb_analyzer* handle() const { return h_; }

 private:
  void check(int32_t rc) const { if (rc) throw Error(rc, ssb_last_error(h_)); }
  ssb_analyzer* h_ = nullptr;
};

// `RBuffer` (tui.rs:37): AllocRingBuffer<f32>::new(capacity) + fill(0.0), written by the cpal callback
// (audio_capture.rs:40-52).  push() is the callback body; the consumer is Analyzer::analyze_microphone_input.
class CaptureRing {
 public:
  explicit CaptureRing(size_t capacity, int device = -1) {
    const int32_t rc = ssb_capture_ring_create(&r_, capacity, device);
    if (rc) throw Error(rc, "ssb_capture_ring_create");
  }
  ~CaptureRing() { ssb_capture_ring_destroy(r_); }
  CaptureRing(const CaptureRing&) = delete;
  CaptureRing& operator=(const CaptureRing&) = delete;
  void push(const float* data, size_t n, bool is_mono) {
    const int32_t rc = ssb_capture_ring_push(r_, data, n, is_mono ? 1 : 0);
    if (rc) throw Error(rc, "ssb_capture_ring_push");
  }
  std::vector<float> to_vec() {
    std::vector<float> out(ssb_capture_ring_capacity(r_));
    const int32_t rc = ssb_capture_ring_to_vec(r_, out.data(), out.size());
    if (rc) throw Error(rc, "ssb_capture_ring_to_vec");
    return out;
  }
  ssb_capture_ring* handle() const { return r_; }

 private:
  ssb_capture_ring* r_ = nullptr;
};

}  // namespace soundscope
