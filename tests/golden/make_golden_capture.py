"""Generates tests/golden/golden_capture_v1.npz: fixed vectors for the formats either side of the path
 — PCM -> f32 for every sample format, a capture-ring push sequence (mono up-mix
quirk, wrap-around) and one microphone tick.

Like golden_v1.npz these come from the CPU oracle (oracle/capture_ref.py + oracle/), not from the reference
binary (Rust, not buildable here): they pin the oracle against regressions and give the GPU tests fixed targets.

    python tests/golden/make_golden_capture.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

FORMATS = ["u8", "s8", "s16le", "s16be", "s24le", "s24be", "s32le", "s32be", "f32le", "f32be", "f64le", "f64be"]
RING_CAP = 4096
RING_PUSHES = [(100, 0), (33, 1), (1, 1), (2000, 0), (1500, 1), (5000, 0), (7, 1), (64, 0)]   # (values, is_mono)


def pcm_raw(fmt):
    rng = np.random.default_rng(1000 + FORMATS.index(fmt))
    bps = O.capture_ref.pcm_bytes_per_sample(fmt)
    if fmt.startswith("f64"):
        v = np.concatenate([[0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, 1e-45, 3.4e38, 3.5e38, 1e-300, 0.1, 1 + 2.0 ** -24],
                            rng.standard_normal(244), rng.standard_normal(256) * 1e-40])
        return np.frombuffer(v.astype(">f8" if fmt.endswith("be") else "<f8").tobytes(), dtype=np.uint8).copy()
    return rng.integers(0, 256, 509 * bps, dtype=np.uint8)


def ring_pushes():
    rng = np.random.default_rng(77)
    return [(rng.standard_normal(n).astype(np.float32), mono) for n, mono in RING_PUSHES]


def mic_signal(rate=44100):
    rng = np.random.default_rng(78)
    n = 15 * rate
    t = np.arange(n) / rate
    x = np.empty(2 * n, dtype=np.float32)
    x[0::2] = 0.4 * np.sin(2 * np.pi * 440 * t) + 0.02 * rng.standard_normal(n)
    x[1::2] = 0.2 * np.sin(2 * np.pi * 440 * t + 1.0) + 0.02 * rng.standard_normal(n)
    return x


def main():
    g = {}
    for fmt in FORMATS:
        raw = pcm_raw(fmt)
        g[f"pcm_{fmt}_raw"] = raw
        g[f"pcm_{fmt}_f32bits"] = O.capture_ref.pcm_to_f32(raw.tobytes(), fmt).view(np.uint32)
    r = O.capture_ref.RingRef(RING_CAP)
    for d, mono in ring_pushes():
        r.callback(d, bool(mono))
    g["ring_to_vec"] = r.to_vec()
    rate = 44100
    ring = O.capture_ref.RingRef(30 * rate)
    x = mic_signal(rate)
    ring.callback(x[: x.size // 2 + 1], False)      # odd split: exercises a pair straddling two callbacks
    ring.callback(x[x.size // 2 + 1:], False)
    a = O.Analyzer()
    mid, side, wave, st, err = O.capture_ref.mic_tick(ring.to_vec(), a)
    assert err is None
    g["mic_mid_db"], g["mic_side_db"] = mid[:, 1], side[:, 1]
    g["mic_wave"] = wave[:, 1].astype(np.float32)
    g["mic_shortterm"] = np.array([st])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_capture_v1.npz")
    np.savez_compressed(out, **g)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
