"""The synthetic signals live in soundscope_b200/synth.py (bench.py and smoke() use them too); re-exported for the tests."""
from soundscope_b200.synth import (  # noqa: F401
    ref_mic_test_assertions, ref_mic_test_ring_fill, ref_sine_f32, stream_batch, sweep_stereo,
)
