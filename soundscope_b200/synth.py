"""Deterministic synthetic inputs of the BASELINE configs (cfg1 sweep, cfg2-style stream batches, the reference tests' own
generators), shared by bench.py, smoke() and the parity tests (no RNG state leaks).  Host-side numpy only."""
import numpy as np


def ref_sine_f32(freq, n=16384, rate=44100, amp=1.0):
    """The reference tests' generator (analyzer.rs:199-204): f32 arithmetic throughout."""
    i = np.arange(n, dtype=np.float32)
    t = i / np.float32(rate)
    ph = np.float32(2.0) * np.float32(np.pi) * np.float32(freq) * t
    return (np.float32(amp) * np.sin(ph.astype(np.float32))).astype(np.float32)


def sweep_stereo(seconds=10.0, rate=48000, amp=0.5, side_gain=0.5):
    """cfg1: log sine sweep 20 Hz -> 20 kHz, closed-form phase, R = side_gain * L; interleaved f32."""
    n = int(seconds * rate)
    t = np.arange(n, dtype=np.float64) / rate
    f0, f1 = 20.0, 20000.0
    k = np.log(f1 / f0) / seconds
    phase = 2 * np.pi * f0 * (np.exp(k * t) - 1.0) / k
    left = amp * np.sin(phase)
    x = np.empty(2 * n, dtype=np.float32)
    x[0::2] = left
    x[1::2] = side_gain * left
    return x


def stream_batch(n_streams, frames, channels, seed=0x5EED, rate=48000, t0=0):
    """cfg2-style batch: per-stream tone 100*2^((s%64)/8) Hz at 0.25 + 0.05 uniform noise -> [n, frames, C] f32."""
    rng = np.random.default_rng(seed)
    t = (np.arange(frames, dtype=np.float64) + t0) / rate
    s = np.arange(n_streams)
    f = 100.0 * 2.0 ** ((s % 64) / 8.0)
    ph = rng.uniform(0, 2 * np.pi, size=(n_streams, channels))
    x = 0.25 * np.sin(2 * np.pi * f[:, None, None] * t[None, :, None] + ph[:, None, :])
    x += 0.05 * rng.uniform(-1, 1, size=(n_streams, frames, channels))
    return x.astype(np.float32)


def ref_mic_test_ring_fill(sr):
    """The reference's microphone-tick tests (tui.rs:2271-2368) fill the ring with
    `(i as f32 * 500.0 * 2.0 * PI / sr as f32).sin()` for i in 0..sr*30 — f32 arithmetic throughout."""
    i = np.arange(sr * 30, dtype=np.float32)
    ph = (i * np.float32(500.0) * np.float32(2.0) * np.float32(np.pi) / np.float32(sr)).astype(np.float32)
    return np.sin(ph).astype(np.float32)


def ref_mic_test_assertions(sr, mid_fft):
    """tui.rs:2289-2302 (and the 48000 / 96000 copies): the spectrum is non-empty and the point at
    round(500 / (sr / 2) * len) reads below -20 dB."""
    assert mid_fft is not None and len(mid_fft) > 0
    freq_bin = np.float32(500.0) / (np.float32(sr) / np.float32(2.0)) * np.float32(len(mid_fft))
    bin_idx = int(np.round(freq_bin))
    assert bin_idx < len(mid_fft), f"Bin index out of range: {bin_idx}"
    amp = mid_fft[bin_idx][1]
    assert amp < -20.0, f"Expected strong signal at ~500Hz, got: {amp}"
    return bin_idx, amp
