"""ctypes loader for libsoundscope_b200.so (the C ABI of include/soundscope_b200.h).

Loading never falls back to a CPU implementation: a missing or unloadable library raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("SSB_LIB") or os.path.join(HERE, "libsoundscope_b200.so")  # SSB_LIB: tuning experiments only

MODE_M = 1
MODE_S = 2 | MODE_M
MODE_I = 4 | MODE_M
MODE_LRA = 8 | MODE_S
MODE_SAMPLE_PEAK = 16 | MODE_M
MODE_TRUE_PEAK = 32 | MODE_M | MODE_SAMPLE_PEAK
MODE_HISTOGRAM = 64
MODE_ALL = 0x7F
MODE_LOUDNESS = MODE_I | MODE_LRA | MODE_HISTOGRAM  # M|S|I|LRA|HISTOGRAM, no peak detectors
FLAG_RING = 1
FFT_MONO, FFT_MID_SIDE = 0, 1
# ssb_pcm_* sample formats (include/soundscope_b200.h)
PCM_U8, PCM_S8, PCM_S16LE, PCM_S16BE, PCM_S24LE, PCM_S24BE = 0, 1, 2, 3, 4, 5
PCM_S32LE, PCM_S32BE, PCM_F32LE, PCM_F32BE, PCM_F64LE, PCM_F64BE = 6, 7, 8, 9, 10, 11
PCM_FORMATS = {
    "u8": PCM_U8, "s8": PCM_S8, "s16le": PCM_S16LE, "s16be": PCM_S16BE, "s24le": PCM_S24LE, "s24be": PCM_S24BE,
    "s32le": PCM_S32LE, "s32be": PCM_S32BE, "f32le": PCM_F32LE, "f32be": PCM_F32BE, "f64le": PCM_F64LE, "f64be": PCM_F64BE,
}

OK = 0
ERR_NAMES = {
    1: "NoMem", 2: "InvalidMode", 3: "InvalidChannelIndex", 4: "TooFewSamples", 5: "NaNValuesNotSupported",
    6: "InfinityValuesNotSupported", 7: "SamplesLengthNotAPowerOfTwo", 8: "InvalidFrequencyLimit",
    9: "ScalingError", 10: "InvalidArgument", 11: "Capacity", 12: "UnalignedQuery", 13: "NoDevice", 14: "Busy",
}


class SsbError(RuntimeError):
    """Non-zero status from the C ABI; `.code` is the SSB_ERR_* value, `.name` the reference's error name."""

    def __init__(self, code, message=""):
        self.code = code
        self.name = ERR_NAMES.get(code, f"Cuda({code - 100})" if code >= 100 else str(code))
        super().__init__(f"{self.name} ({code}): {message}")


# every symbol include/soundscope_b200.h declares; tests/test_abi.py checks the .so exports each one
SYMBOLS = [
    "ssb_abi_version", "ssb_analyzer_create", "ssb_analyzer_destroy", "ssb_create_loudness_meter",
    "ssb_sample_rate", "ssb_channels", "ssb_n_streams", "ssb_last_error", "ssb_set_stream", "ssb_use_own_stream", "ssb_sync",
    "ssb_launch_count", "ssb_add_frames_f32", "ssb_add_frames_f32_device", "ssb_add_frames_f32_device_results", "ssb_add_samples", "ssb_reset",
    "ssb_loudness_momentary", "ssb_loudness_shortterm", "ssb_loudness_global", "ssb_loudness_range",
    "ssb_true_peak", "ssb_sample_peak", "ssb_get_true_peak", "ssb_result_stride", "ssb_results_device",
    "ssb_calculate_integrated_lufs", "ssb_get_fft", "ssb_fft_bins", "ssb_fft_axis", "ssb_fft_batch_device", "ssb_fft_batch_device_y",
    "ssb_process_tick", "ssb_preanalyze_file", "ssb_get_waveform", "ssb_waveform_device", "ssb_mid_side", "ssb_mid_side_device", "ssb_filter_coeffs",
    "ssb_histograms", "ssb_profile_enable", "ssb_profile_read", "ssb_debug_force_generic", "ssb_gather_create", "ssb_gather_open", "ssb_gather_select", "ssb_gather_rows", "ssb_gather_epoch", "ssb_gather_wait", "ssb_gather_destroy",
    "ssb_tick_fft_status", "ssb_true_peak_factor", "ssb_debug_force_true_peak_factor", "ssb_debug_histogram_index",
    "ssb_pcm_bytes_per_sample", "ssb_pcm_to_f32", "ssb_pcm_to_f32_device", "ssb_add_frames_pcm", "ssb_add_frames_pcm_device",
    "ssb_capture_ring_create", "ssb_capture_ring_destroy", "ssb_capture_ring_capacity", "ssb_capture_ring_written",
    "ssb_capture_ring_push", "ssb_capture_ring_to_vec", "ssb_mic_tick",
]

_lib = None


def library_path():
    return _SO


def lib():
    """Load (once) and type the C ABI.  Raises if the shared library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise ImportError(
            f"{_SO} is missing: build it with `python -m soundscope_b200.build` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    L = C.CDLL(_SO)
    vp, f32p, f64p, szp, i32p = C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_int32)
    sig = {
        "ssb_abi_version": (C.c_uint32, []),
        "ssb_analyzer_create": (C.c_int32, [C.POINTER(vp), C.c_uint32, C.c_uint32, C.c_int32, C.c_size_t, C.c_int32, C.c_uint32]),
        "ssb_analyzer_destroy": (None, [vp]),
        "ssb_create_loudness_meter": (C.c_int32, [vp, C.c_uint32, C.c_uint32]),
        "ssb_sample_rate": (C.c_uint32, [vp]),
        "ssb_channels": (C.c_uint32, [vp]),
        "ssb_n_streams": (C.c_size_t, [vp]),
        "ssb_last_error": (C.c_char_p, [vp]),
        "ssb_set_stream": (C.c_int32, [vp, vp]),
        "ssb_use_own_stream": (C.c_int32, [vp]),
        "ssb_sync": (C.c_int32, [vp]),
        "ssb_launch_count": (C.c_uint64, [vp]),
        "ssb_add_frames_f32": (C.c_int32, [vp, f32p, C.c_size_t]),
        "ssb_add_frames_f32_device": (C.c_int32, [vp, f32p, C.c_size_t]),
        "ssb_add_frames_f32_device_results": (C.c_int32, [vp, f32p, C.c_size_t, f64p]),
        "ssb_add_samples": (C.c_int32, [vp, f32p, C.c_size_t]),
        "ssb_reset": (C.c_int32, [vp]),
        "ssb_loudness_momentary": (C.c_int32, [vp, f64p]),
        "ssb_loudness_shortterm": (C.c_int32, [vp, f64p]),
        "ssb_loudness_global": (C.c_int32, [vp, f64p]),
        "ssb_loudness_range": (C.c_int32, [vp, f64p]),
        "ssb_true_peak": (C.c_int32, [vp, f64p]),
        "ssb_sample_peak": (C.c_int32, [vp, f64p]),
        "ssb_get_true_peak": (C.c_int32, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "ssb_result_stride": (C.c_size_t, [vp]),
        "ssb_results_device": (C.c_int32, [vp, f64p]),
        "ssb_debug_histogram_index": (C.c_int32, [vp, f64p, C.c_size_t, vp]),
        "ssb_gather_create": (C.c_int32, [vp, C.c_uint32, C.c_uint32, vp]),
        "ssb_gather_open": (C.c_int32, [vp, vp]),
        "ssb_gather_select": (C.c_int32, [vp, C.c_int32]),
        "ssb_gather_rows": (C.c_void_p, [vp, C.c_int32]),
        "ssb_gather_epoch": (C.c_uint64, [vp]),
        "ssb_gather_wait": (C.c_int32, [vp]),
        "ssb_gather_destroy": (C.c_int32, [vp]),
        "ssb_tick_fft_status": (C.c_int32, [vp, vp]),
        "ssb_true_peak_factor": (C.c_int32, [vp]),
        "ssb_debug_force_true_peak_factor": (C.c_int32, [vp, C.c_int32]),
        "ssb_calculate_integrated_lufs": (C.c_int32, [vp, C.c_uint32, f32p, C.c_size_t, C.POINTER(C.c_double), i32p]),
        "ssb_get_fft": (C.c_int32, [vp, f32p, C.c_size_t, f64p, C.c_size_t, szp]),
        "ssb_fft_bins": (C.c_int32, [C.c_size_t, C.c_uint32, szp, szp]),
        "ssb_fft_axis": (C.c_int32, [C.c_size_t, C.c_uint32, f64p, f64p, C.c_size_t, szp]),
        "ssb_fft_batch_device": (C.c_int32, [vp, f32p, C.c_int32, C.c_size_t, C.c_size_t, f32p, vp]),
        "ssb_fft_batch_device_y": (C.c_int32, [vp, f32p, C.c_int32, C.c_size_t, C.c_size_t, f64p, vp]),
        "ssb_process_tick": (C.c_int32, [vp, f32p, C.c_size_t, C.c_size_t, f64p, f64p, C.c_size_t, szp,
                                         C.POINTER(C.c_double), i32p, i32p]),
        "ssb_preanalyze_file": (C.c_int32, [vp, f32p, C.c_size_t, C.c_uint32, C.c_double, f64p, C.c_size_t, szp,
                                            C.POINTER(C.c_double), i32p]),
        "ssb_get_waveform": (C.c_int32, [vp, f32p, C.c_size_t, C.c_double, f64p, C.c_size_t, szp]),
        "ssb_waveform_device": (C.c_int32, [vp, f32p, C.c_size_t, C.c_double, f32p, C.c_size_t, szp]),
        "ssb_mid_side": (C.c_int32, [vp, f32p, C.c_size_t, f32p, f32p, szp]),
        "ssb_mid_side_device": (C.c_int32, [vp, f32p, C.c_size_t, f32p, f32p]),
        "ssb_filter_coeffs": (C.c_int32, [vp, f64p, f64p]),
        "ssb_histograms": (C.c_int32, [vp, C.c_size_t, vp, vp]),
        "ssb_profile_enable": (C.c_int32, [vp, C.c_int32]),
        "ssb_debug_force_generic": (C.c_int32, [vp, C.c_int32]),
        "ssb_profile_read": (C.c_int32, [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
        "ssb_pcm_bytes_per_sample": (C.c_size_t, [C.c_int32]),
        "ssb_pcm_to_f32": (C.c_int32, [vp, vp, C.c_size_t, C.c_int32, f32p]),
        "ssb_pcm_to_f32_device": (C.c_int32, [vp, vp, C.c_size_t, C.c_int32, f32p]),
        "ssb_add_frames_pcm": (C.c_int32, [vp, vp, C.c_int32, C.c_size_t]),
        "ssb_add_frames_pcm_device": (C.c_int32, [vp, vp, C.c_int32, C.c_size_t]),
        "ssb_capture_ring_create": (C.c_int32, [C.POINTER(vp), C.c_size_t, C.c_int32]),
        "ssb_capture_ring_destroy": (None, [vp]),
        "ssb_capture_ring_capacity": (C.c_size_t, [vp]),
        "ssb_capture_ring_written": (C.c_uint64, [vp]),
        "ssb_capture_ring_push": (C.c_int32, [vp, f32p, C.c_size_t, C.c_int32]),
        "ssb_capture_ring_to_vec": (C.c_int32, [vp, f32p, C.c_size_t]),
        "ssb_mic_tick": (C.c_int32, [vp, vp, C.c_size_t, C.c_size_t, C.c_double, f64p, f64p, C.c_size_t, szp,
                                     f64p, C.c_size_t, szp, C.POINTER(C.c_double), i32p, i32p]),
    }
    assert set(sig) == set(SYMBOLS)
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(handle, rc):
    if rc != OK:
        msg = lib().ssb_last_error(handle).decode() if handle else ""
        raise SsbError(rc, msg)
