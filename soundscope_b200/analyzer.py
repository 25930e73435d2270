"""Host-side mirror of the reference's `analyzer::Analyzer` (reference src/analyzer.rs:29-183).

Same method names, argument meaning and error behaviour; every method is one call through the C ABI
(include/soundscope_b200.h) into the CUDA kernels.  Inputs and outputs are host numpy arrays, as the
reference's are host slices / Vecs.
"""
import ctypes as C

import numpy as np

from ._lib import FLAG_RING, MODE_ALL, SsbError, check, lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data


class Analyzer:
    """One reference `Analyzer`: a Mode::all() loudness meter plus the stateless spectrum/waveform ops."""

    def __init__(self, device=-1):
        # Default (analyzer.rs:34-45): EbuR128::new(2, 44100, Mode::all()); panics on failure
        self._h = C.c_void_p()
        self._device = device
        rc = lib().ssb_analyzer_create(C.byref(self._h), 2, 44100, MODE_ALL, 1, device, FLAG_RING)
        if rc:
            raise SsbError(rc, "Failed to create loudness meter")

    def close(self):
        if getattr(self, "_h", None):
            lib().ssb_analyzer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # analyzer.rs:49-53
    def create_loudness_meter(self, channels, rate):
        check(self._h, lib().ssb_create_loudness_meter(self._h, channels, rate))

    # analyzer.rs:55-105 -> ndarray [n_points, 2] of (chart_x, dB)
    def get_fft(self, samples):
        a, p = _f32(samples)
        cap = a.size // 2 + 1
        out = np.empty((max(cap, 1), 2), dtype=np.float64)
        n = C.c_size_t(0)
        check(self._h, lib().ssb_get_fft(self._h, p, a.size, out.ctypes.data, cap, C.byref(n)))
        return out[: n.value]

    # analyzer.rs:107-137 (associated fn in the reference; needs a device here, so any instance serves)
    def get_waveform(self, samples, waveform_window):
        a, p = _f32(samples)
        n = C.c_size_t(0)
        rc = lib().ssb_get_waveform(self._h, p, a.size, float(waveform_window), None, 0, C.byref(n))
        if rc not in (0, 11):  # 11 = SSB_ERR_CAPACITY: the sizing call
            check(self._h, rc)
        out = np.empty((max(n.value, 1), 2), dtype=np.float64)
        check(self._h, lib().ssb_get_waveform(self._h, p, a.size, float(waveform_window), out.ctypes.data,
                                               n.value, C.byref(n)))
        return out[: n.value]

    # analyzer.rs:139-141
    def add_samples(self, samples):
        a, p = _f32(samples)
        check(self._h, lib().ssb_add_samples(self._h, p, a.size))

    # analyzer.rs:143-145
    def reset(self):
        check(self._h, lib().ssb_reset(self._h))

    def _scalar(self, fn):
        v = C.c_double(0)
        check(self._h, fn(self._h, C.addressof(v)))
        return v.value

    def get_momentary_lufs(self):
        return self._scalar(lib().ssb_loudness_momentary)

    # analyzer.rs:147-149
    def get_shortterm_lufs(self):
        return self._scalar(lib().ssb_loudness_shortterm)

    # analyzer.rs:151-153
    def get_integrated_lufs(self):
        return self._scalar(lib().ssb_loudness_global)

    # analyzer.rs:155-157
    def get_loudness_range(self):
        return self._scalar(lib().ssb_loudness_range)

    # analyzer.rs:159-164
    def get_true_peak(self):
        l, r = C.c_double(0), C.c_double(0)
        check(self._h, lib().ssb_get_true_peak(self._h, C.byref(l), C.byref(r)))
        return l.value, r.value

    # analyzer.rs:166-168
    def sample_rate(self):
        return lib().ssb_sample_rate(self._h)

    # analyzer.rs:170-182 -> float or None
    def calculate_integrated_lufs(self, channels, samples):
        a, p = _f32(samples)
        out, some = C.c_double(0), C.c_int32(0)
        check(self._h, lib().ssb_calculate_integrated_lufs(self._h, channels, p, a.size, C.byref(out), C.byref(some)))
        return out.value if some.value else None

    def process_tick(self, tail, lufs_samples=16384):
        """One player tick (reference src/tui.rs:1482-1552) in one call: `tail` = the last n_fft stereo frames,
        interleaved.  Returns (mid_fft, side_fft, shortterm_lufs, fft_status, lufs_status); the FFT arrays are
        None when the reference's get_fft would return Err (it then shows `vec![(0., 0.)]`)."""
        a, p = _f32(tail)
        n_fft = a.size // 2
        cap = n_fft // 2 + 1
        mid = np.empty((max(cap, 1), 2), dtype=np.float64)
        side = np.empty((max(cap, 1), 2), dtype=np.float64)
        n, st = C.c_size_t(0), C.c_double(0)
        fs, ls = C.c_int32(0), C.c_int32(0)
        check(self._h, lib().ssb_process_tick(self._h, p, n_fft, lufs_samples, mid.ctypes.data, side.ctypes.data, cap,
                                               C.byref(n), C.byref(st), C.byref(fs), C.byref(ls)))
        ok = fs.value == 0
        return (mid[: n.value] if ok else None, side[: n.value] if ok else None, st.value, fs.value, ls.value)

    def analyze_microphone_input(self, ring, n_fft=16384, lufs_samples=16384, waveform_window=15.0):
        """One microphone tick (reference src/tui.rs:1427-1480) on one snapshot of `ring` (a CaptureRing):
        mid/side spectra of mid[15*rate - n_fft .. 15*rate], the 15 s waveform of mid, add_samples of the ring's
        last `lufs_samples` values and the short-term loudness.  Returns (mid_fft, side_fft, waveform,
        shortterm_lufs, fft_status, lufs_status); the FFT arrays are None when get_fft would return Err."""
        cap = n_fft // 2 + 1
        mid = np.empty((max(cap, 1), 2), dtype=np.float64)
        side = np.empty((max(cap, 1), 2), dtype=np.float64)
        w = waveform_window * 1000.0
        wave_cap = 2 * (int(w) if w > 0 else 0) + 2
        wave = np.empty((wave_cap, 2), dtype=np.float64)
        n, nw, st = C.c_size_t(0), C.c_size_t(0), C.c_double(0)
        fs, ls = C.c_int32(0), C.c_int32(0)
        check(self._h, lib().ssb_mic_tick(self._h, ring._r, n_fft, lufs_samples, float(waveform_window),
                                          mid.ctypes.data, side.ctypes.data, cap, C.byref(n), wave.ctypes.data, wave_cap,
                                          C.byref(nw), C.byref(st), C.byref(fs), C.byref(ls)))
        ok = fs.value == 0
        return (mid[: n.value] if ok else None, side[: n.value] if ok else None, wave[: nw.value], st.value,
                fs.value, ls.value)

    def add_pcm(self, raw, fmt):
        """add_samples on raw interleaved PCM bytes (the decode_file conversion runs on the device)."""
        from .capture import pcm_format
        code = pcm_format(fmt)
        b = np.frombuffer(raw, dtype=np.uint8) if not isinstance(raw, np.ndarray) else np.ascontiguousarray(raw).view(np.uint8).ravel()
        bps = lib().ssb_pcm_bytes_per_sample(code)
        ch = lib().ssb_channels(self._h)
        n = b.size // bps if bps else 0
        if not bps or n % ch:
            raise SsbError(1 if bps else 10, "ragged PCM input")
        check(self._h, lib().ssb_add_frames_pcm(self._h, b.ctypes.data, code, n // ch))

    def preanalyze_file(self, samples, rate, duration_s):
        """File-selected pre-analysis (reference src/tui.rs:1207-1241) in one call: returns
        (waveform [n, 2], integrated LUFS or None); the handle's meter becomes (2, rate)."""
        a, p = _f32(samples)
        w = duration_s * 1000.0
        cap = 2 * (int(w) if w > 0 else 0) + 2
        out = np.empty((cap, 2), dtype=np.float64)
        n, v, some = C.c_size_t(0), C.c_double(0), C.c_int32(0)
        check(self._h, lib().ssb_preanalyze_file(self._h, p, a.size, rate, float(duration_s), out.ctypes.data, cap,
                                                  C.byref(n), C.byref(v), C.byref(some)))
        return out[: n.value], (v.value if some.value else None)

    # introspection (tests)
    def filter_coeffs(self):
        b, a = np.zeros(5), np.zeros(5)
        check(self._h, lib().ssb_filter_coeffs(self._h, b.ctypes.data, a.ctypes.data))
        return b, a

    def histograms(self, stream=0):
        blk, st = np.zeros(1000, dtype=np.uint64), np.zeros(1000, dtype=np.uint64)
        check(self._h, lib().ssb_histograms(self._h, stream, blk.ctypes.data, st.ctypes.data))
        return blk, st

    def get_mid_and_side_samples(self, samples):
        return get_mid_and_side_samples(samples, self)


_default = None


def get_mid_and_side_samples(samples, analyzer=None):
    """reference src/audio_player.rs:400-419 — (mid, side) f32 arrays of len(samples)//2 frames."""
    global _default
    if analyzer is None:
        if _default is None:
            _default = Analyzer()
        analyzer = _default
    a, p = _f32(samples)
    frames = a.size // 2
    mid = np.empty(frames, dtype=np.float32)
    side = np.empty(frames, dtype=np.float32)
    n = C.c_size_t(0)
    check(analyzer._h, lib().ssb_mid_side(analyzer._h, p, a.size, mid.ctypes.data, side.ctypes.data, C.byref(n)))
    return mid, side
