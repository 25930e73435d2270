"""Host-side mirror of the formats either side of the analyzer path.

* `pcm_to_f32` — what the reference's `AudioFile::decode_file` (src/audio_player.rs:169-267) yields for
  WAV / AIFF PCM: symphonia's decoded samples converted to interleaved f32 by
  `SampleBuffer::<f32>::copy_interleaved_ref` (audio_player.rs:248).
* `CaptureRing` — the reference's `RBuffer` (src/tui.rs:37): AllocRingBuffer<f32> of 30*rate values, zero-filled
  (main.rs:63-65, tui.rs:1783-1786), fed by the cpal callback (audio_capture.rs:40-52, mono up-mix quirk included).

Every method is one call through the C ABI (include/soundscope_b200.h); nothing is computed in Python.
"""
import ctypes as C

import numpy as np

from ._lib import PCM_FORMATS, SsbError, check, lib


def pcm_format(fmt):
    """'s16le' / 's24be' / ... or the numeric SSB_PCM_* code -> code"""
    if isinstance(fmt, str):
        return PCM_FORMATS[fmt.lower()]
    return int(fmt)


def pcm_bytes_per_sample(fmt):
    return lib().ssb_pcm_bytes_per_sample(pcm_format(fmt))


def pcm_to_f32(analyzer, raw, fmt):
    """raw: bytes-like / uint8 array of interleaved PCM -> float32 array of len(raw) // bytes_per_sample samples."""
    code = pcm_format(fmt)
    b = np.frombuffer(raw, dtype=np.uint8) if not isinstance(raw, np.ndarray) else np.ascontiguousarray(raw).view(np.uint8).ravel()
    bps = lib().ssb_pcm_bytes_per_sample(code)
    if not bps:
        raise SsbError(10, f"unknown PCM format {fmt!r}")
    n = b.size // bps
    out = np.empty(n, dtype=np.float32)
    check(analyzer._h, lib().ssb_pcm_to_f32(analyzer._h, b.ctypes.data, n, code, out.ctypes.data))
    return out


class CaptureRing:
    """`Arc<Mutex<AllocRingBuffer<f32>>>` with `capacity` values, pre-filled with 0.0."""

    def __init__(self, capacity, device=-1):
        self._r = C.c_void_p()
        rc = lib().ssb_capture_ring_create(C.byref(self._r), int(capacity), device)
        if rc:
            raise SsbError(rc, "ssb_capture_ring_create")

    def close(self):
        if getattr(self, "_r", None):
            lib().ssb_capture_ring_destroy(self._r)
            self._r = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def capacity(self):
        return lib().ssb_capture_ring_capacity(self._r)

    @property
    def written(self):
        return lib().ssb_capture_ring_written(self._r)

    def push(self, data, is_mono=False):
        """One cpal callback (audio_capture.rs:40-52): `extend(data)`, or the mono up-mix `[x0, 0, x1, 0, x2, ...]`."""
        a = np.ascontiguousarray(data, dtype=np.float32)
        rc = lib().ssb_capture_ring_push(self._r, a.ctypes.data, a.size, 1 if is_mono else 0)
        if rc:
            raise SsbError(rc, "ssb_capture_ring_push")

    def to_vec(self):
        out = np.empty(self.capacity, dtype=np.float32)
        rc = lib().ssb_capture_ring_to_vec(self._r, out.ctypes.data, out.size)
        if rc:
            raise SsbError(rc, "ssb_capture_ring_to_vec")
        return out
