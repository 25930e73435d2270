"""ctypes binding of liboracle.so (oracle/ebur128_ref.c, oracle/spectrum_ref.c).

TEST INFRASTRUCTURE ONLY — see oracle/oracle.h.  Mirrors the reference's `Analyzer` method
surface (reference src/analyzer.rs:47-183) on the CPU so parity tests read like the
reference's own tests.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

MODE_M = 1
MODE_S = 2 | MODE_M
MODE_I = 4 | MODE_M
MODE_LRA = 8 | MODE_S
MODE_SAMPLE_PEAK = 16 | MODE_M
MODE_TRUE_PEAK = 32 | MODE_M | MODE_SAMPLE_PEAK
MODE_HISTOGRAM = 64
MODE_ALL = 0x7F
MODE_LOUDNESS = MODE_I | MODE_LRA | MODE_HISTOGRAM  # M|S|I|LRA|HISTOGRAM, no peaks

OK, ERR_NOMEM, ERR_INVALID_MODE, ERR_INVALID_CHANNEL_INDEX = 0, 1, 2, 3


def _host_tag():
    """What -march=native means on this machine: the CPU flag set (the .so travels between boxes with the tree)."""
    import hashlib
    try:
        with open("/proc/cpuinfo") as f:
            flags = next((ln for ln in f if ln.startswith("flags")), "")
    except OSError:
        flags = ""
    return hashlib.sha1(flags.encode()).hexdigest()[:16]


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc only, no reference sources).  Built -march=native, so it
    is rebuilt when the host CPU differs from the one that built the copy in the tree."""
    srcs = [os.path.join(_HERE, f) for f in ("ebur128_ref.c", "spectrum_ref.c", "oracle.h", "Makefile")]
    stamp = _SO + ".host"
    tag = _host_tag()
    try:
        same_host = open(stamp).read().strip() == tag
    except OSError:
        same_host = False
    if force or not same_host or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
        with open(stamp, "w") as f:
            f.write(tag)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        f32p, f64p, u64p = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_uint64)
        L.orc_ebur128_new.restype = C.c_void_p
        L.orc_ebur128_new.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_int)]
        L.orc_ebur128_free.argtypes = [C.c_void_p]
        L.orc_ebur128_add_frames_f32.argtypes = [C.c_void_p, f32p, C.c_size_t]
        L.orc_ebur128_reset.argtypes = [C.c_void_p]
        for n in ("momentary", "shortterm", "global", "range"):
            getattr(L, f"orc_ebur128_loudness_{n}").argtypes = [C.c_void_p, f64p]
        L.orc_ebur128_true_peak.argtypes = [C.c_void_p, C.c_uint32, f64p]
        L.orc_ebur128_sample_peak.argtypes = [C.c_void_p, C.c_uint32, f64p]
        L.orc_ebur128_coeffs.argtypes = [C.c_void_p, f64p, f64p]
        L.orc_ebur128_histograms.argtypes = [C.c_void_p, u64p, u64p]
        L.orc_interp_taps.restype = C.c_size_t
        L.orc_interp_taps.argtypes = [C.c_uint32, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
        L.orc_histogram_energy.restype = C.c_double
        L.orc_histogram_energy.argtypes = [C.c_uint]
        L.orc_histogram_boundary.restype = C.c_double
        L.orc_histogram_boundary.argtypes = [C.c_uint]
        L.orc_find_histogram_index.restype = C.c_size_t
        L.orc_find_histogram_index.argtypes = [C.c_double]
        L.orc_calculate_integrated_lufs.argtypes = [C.c_uint32, C.c_uint32, f32p, C.c_size_t, f64p]
        L.orc_batch_new.restype = C.c_void_p
        L.orc_batch_new.argtypes = [C.c_size_t, C.c_uint32, C.c_uint32, C.c_int]
        L.orc_batch_new_mt.restype = C.c_void_p
        L.orc_batch_new_mt.argtypes = [C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_int]
        L.orc_batch_free.argtypes = [C.c_void_p]
        L.orc_batch_add_frames.argtypes = [C.c_void_p, f32p, C.c_size_t, C.c_int]
        L.orc_batch_query.argtypes = [C.c_void_p, f64p, f64p, f64p, f64p, f64p, C.c_int]
        L.orc_batch_histograms.argtypes = [C.c_void_p, C.c_size_t, u64p, u64p]
        L.orc_cosf.restype = C.c_float
        L.orc_cosf.argtypes = [C.c_float]
        L.orc_hann_window.argtypes = [f32p, C.c_size_t, f32p]
        L.orc_rfft_mag.argtypes = [f32p, C.c_size_t, f32p]
        L.orc_scale_to_dbfs.restype = C.c_float
        L.orc_scale_to_dbfs.argtypes = [C.c_float, C.c_float]
        L.orc_get_fft.argtypes = [f32p, C.c_size_t, C.c_uint32, f64p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.orc_fft_bin_range.restype = C.c_size_t
        L.orc_fft_bin_range.argtypes = [C.c_size_t, C.c_uint32, C.POINTER(C.c_size_t)]
        L.orc_get_waveform.restype = C.c_size_t
        L.orc_get_waveform.argtypes = [f32p, C.c_size_t, C.c_double, f64p, C.c_size_t]
        L.orc_mid_side.restype = C.c_size_t
        L.orc_mid_side.argtypes = [f32p, C.c_size_t, f32p, f32p]
        _lib = L
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _f64ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class OracleError(Exception):
    def __init__(self, code, what=""):
        super().__init__(f"oracle error {code} {what}")
        self.code = code


class EbuR128:
    """ebur128::EbuR128 restatement (single stream)."""

    def __init__(self, channels, rate, mode=MODE_ALL):
        err = C.c_int(0)
        self._h = lib().orc_ebur128_new(channels, rate, mode, C.byref(err))
        if not self._h:
            raise OracleError(err.value, "EbuR128::new")
        self.channels, self.rate, self.mode = channels, rate, mode

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_ebur128_free(self._h)
            self._h = None

    def add_frames_f32(self, samples):
        a, p = _f32(samples)
        rc = lib().orc_ebur128_add_frames_f32(self._h, p, a.size)
        if rc:
            raise OracleError(rc, "add_frames_f32")

    def reset(self):
        lib().orc_ebur128_reset(self._h)

    def _q(self, name):
        out = C.c_double(0)
        rc = getattr(lib(), f"orc_ebur128_loudness_{name}")(self._h, C.byref(out))
        if rc:
            raise OracleError(rc, name)
        return out.value

    def loudness_momentary(self):
        return self._q("momentary")

    def loudness_shortterm(self):
        return self._q("shortterm")

    def loudness_global(self):
        return self._q("global")

    def loudness_range(self):
        return self._q("range")

    def true_peak(self, ch):
        out = C.c_double(0)
        rc = lib().orc_ebur128_true_peak(self._h, ch, C.byref(out))
        if rc:
            raise OracleError(rc, "true_peak")
        return out.value

    def sample_peak(self, ch):
        out = C.c_double(0)
        rc = lib().orc_ebur128_sample_peak(self._h, ch, C.byref(out))
        if rc:
            raise OracleError(rc, "sample_peak")
        return out.value

    def coeffs(self):
        b = np.zeros(5)
        a = np.zeros(5)
        lib().orc_ebur128_coeffs(self._h, _f64ptr(b), _f64ptr(a))
        return b, a

    def histograms(self):
        blk = np.zeros(1000, dtype=np.uint64)
        st = np.zeros(1000, dtype=np.uint64)
        p = C.POINTER(C.c_uint64)
        lib().orc_ebur128_histograms(self._h, blk.ctypes.data_as(p), st.ctypes.data_as(p))
        return blk, st


class Analyzer:
    """CPU restatement of reference `analyzer::Analyzer` (src/analyzer.rs:29-183)."""

    def __init__(self):  # Default: analyzer.rs:34-45
        self._meter = EbuR128(2, 44100, MODE_ALL)
        self._rate = 44100

    def create_loudness_meter(self, channels, rate):  # analyzer.rs:49-53
        self._rate = rate
        self._meter = EbuR128(channels, rate, MODE_ALL)

    def get_fft(self, samples):  # analyzer.rs:55-105
        return get_fft(samples, self._rate)

    @staticmethod
    def get_waveform(samples, waveform_window):  # analyzer.rs:107-137
        return get_waveform(samples, waveform_window)

    def add_samples(self, samples):  # analyzer.rs:139-141
        self._meter.add_frames_f32(samples)

    def reset(self):  # analyzer.rs:143-145
        self._meter.reset()

    def get_shortterm_lufs(self):
        return self._meter.loudness_shortterm()

    def get_integrated_lufs(self):
        return self._meter.loudness_global()

    def get_loudness_range(self):
        return self._meter.loudness_range()

    def get_true_peak(self):  # analyzer.rs:159-164
        return self._meter.true_peak(0), self._meter.true_peak(1)

    def sample_rate(self):
        return self._rate

    def calculate_integrated_lufs(self, channels, samples):  # analyzer.rs:170-182
        return calculate_integrated_lufs(channels, self._rate, samples)


def calculate_integrated_lufs(channels, rate, samples):
    a, p = _f32(samples)
    out = C.c_double(0)
    ok = lib().orc_calculate_integrated_lufs(channels, rate, p, a.size, C.byref(out))
    return out.value if ok else None


def get_fft(samples, rate):
    a, p = _f32(samples)
    cap = a.size // 2 + 1
    out = np.zeros((max(cap, 1), 2))
    n = C.c_size_t(0)
    rc = lib().orc_get_fft(p, a.size, rate, _f64ptr(out), cap, C.byref(n))
    if rc:
        raise OracleError(rc, "get_fft")
    return out[: n.value].copy()


def fft_bin_range(n, rate):
    k0 = C.c_size_t(0)
    cnt = lib().orc_fft_bin_range(n, rate, C.byref(k0))
    return k0.value, cnt


def hann_window(samples):
    a, p = _f32(samples)
    out = np.empty_like(a)
    lib().orc_hann_window(p, a.size, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def rfft_mag(windowed):
    a, p = _f32(windowed)
    out = np.empty(a.size // 2 + 1, dtype=np.float32)
    rc = lib().orc_rfft_mag(p, a.size, out.ctypes.data_as(C.POINTER(C.c_float)))
    if rc:
        raise OracleError(rc, "rfft_mag")
    return out


def cosf(x):
    return lib().orc_cosf(C.c_float(x))


def scale_to_dbfs(val, n):
    return lib().orc_scale_to_dbfs(C.c_float(val), C.c_float(n))


def get_waveform(samples, waveform_window):
    a, p = _f32(samples)
    w = waveform_window * 1000.0
    window = 0 if not (w > 0) else int(min(w, 2**40))
    cap = (2 * window if a.size else 0) + 2
    out = np.zeros((cap, 2))
    n = lib().orc_get_waveform(p, a.size, float(waveform_window), _f64ptr(out), cap)
    assert n <= cap
    return out[:n].copy()


def mid_side(samples):
    a, p = _f32(samples)
    frames = a.size // 2
    mid = np.empty(frames, dtype=np.float32)
    side = np.empty(frames, dtype=np.float32)
    fp = C.POINTER(C.c_float)
    lib().orc_mid_side(p, a.size, mid.ctypes.data_as(fp), side.ctypes.data_as(fp))
    return mid, side


def interp_taps(rate):
    factor = C.c_uint(0)
    counts = (C.c_uint * 4)()
    total = lib().orc_interp_taps(rate, C.byref(factor), counts)
    return factor.value, list(counts), total


def histogram_energy(i):
    return lib().orc_histogram_energy(i)


def histogram_boundary(i):
    return lib().orc_histogram_boundary(i)


def find_histogram_index(energy):
    """ebur128 find_histogram_index (the crate's bisection over the 1001 boundaries); energy >= boundary[0]."""
    return int(lib().orc_find_histogram_index(float(energy)))


class Batch:
    """n independent meters — the timed CPU baseline (bench.py) and batch parity checker."""

    def __init__(self, n_streams, channels, rate, mode=MODE_ALL, threads=1):
        self._h = lib().orc_batch_new_mt(n_streams, channels, rate, mode, threads)
        if not self._h:
            raise OracleError(ERR_NOMEM, "batch_new")
        self.n, self.channels, self.rate, self.mode = n_streams, channels, rate, mode

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_batch_free(self._h)
            self._h = None

    def add_frames(self, x, threads=1):
        """x: [n_streams, frames, channels] float32."""
        a, p = _f32(x)
        frames = a.size // (self.n * self.channels)
        assert frames * self.n * self.channels == a.size
        rc = lib().orc_batch_add_frames(self._h, p, frames, threads)
        if rc:
            raise OracleError(rc, "batch_add_frames")

    def _per_stream_hist(self, s):
        blk = np.zeros(1000, dtype=np.uint64)
        st = np.zeros(1000, dtype=np.uint64)
        p = C.POINTER(C.c_uint64)
        lib().orc_batch_histograms(self._h, s, blk.ctypes.data_as(p), st.ctypes.data_as(p))
        return blk, st

    def query(self, threads=1):
        n, ch = self.n, self.channels
        m, s, g, r = (np.zeros(n) for _ in range(4))
        tp = np.zeros((n, ch))
        has_tp = (self.mode & MODE_TRUE_PEAK) == MODE_TRUE_PEAK
        lib().orc_batch_query(self._h, _f64ptr(m), _f64ptr(s), _f64ptr(g), _f64ptr(r),
                              _f64ptr(tp) if has_tp else None, threads)
        return {"momentary": m, "shortterm": s, "global": g, "range": r, "true_peak": tp if has_tp else None}
