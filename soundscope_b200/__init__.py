"""soundscope_b200 — B200-native (sm_100a) implementation of soundscope's analyzer hot path.

Host-side mirror of the reference's `analyzer::Analyzer` (reference src/analyzer.rs:29-183) over the
C ABI in include/soundscope_b200.h.  All arithmetic runs in hand-written CUDA kernels inside
libsoundscope_b200.so; there is no CPU fallback — loading fails loudly if the library is missing and
creating an analyzer fails if no sm_100 device is usable.
"""
from ._lib import (  # noqa: F401
    SsbError, lib, library_path, MODE_ALL, MODE_I, MODE_LRA, MODE_M, MODE_S, MODE_SAMPLE_PEAK,
    MODE_TRUE_PEAK, MODE_HISTOGRAM, MODE_LOUDNESS, FLAG_RING, FFT_MONO, FFT_MID_SIDE, PCM_FORMATS,
)
from .analyzer import Analyzer, get_mid_and_side_samples  # noqa: F401
from .batch import BatchAnalyzer  # noqa: F401
from .capture import CaptureRing, pcm_bytes_per_sample, pcm_to_f32  # noqa: F401

__all__ = ["Analyzer", "BatchAnalyzer", "get_mid_and_side_samples", "CaptureRing", "pcm_to_f32",
           "pcm_bytes_per_sample", "SsbError", "lib", "library_path"]
