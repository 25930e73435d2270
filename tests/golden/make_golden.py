"""Generates tests/golden/golden_v1.npz.

The reference is Rust and cannot be built or imported in this image (no cargo/rustc, crates not
vendored), so these vectors are produced by the CPU oracle (oracle/, a restatement of the reference's
algorithm) on the reference's own unit-test inputs (reference src/analyzer.rs:191-385) and on the
BASELINE cfg1 sweep.  They pin the oracle against regressions and give the GPU tests fixed targets;
they are NOT outputs of the reference binary ("parity unpinned").

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from tests.signals import ref_sine_f32, sweep_stereo  # noqa: E402


def main():
    g = {}
    rate = 44100
    res = np.float32(rate) / np.float32(16384.0)
    bin_1k = int(np.round(np.float32(1000.0) / res))
    bin_125 = int(np.round(np.float32(125.0) / res))
    f_1k = float(np.float32(bin_1k) * res)
    f_125 = float(np.float32(bin_125) * res)
    g["fft_freqs"] = np.array([440.0, f_1k, f_125])
    for name, f in (("440", 440.0), ("1k", f_1k), ("125", f_125)):
        g[f"fft_{name}"] = O.get_fft(ref_sine_f32(f), rate)
    # test_get_waveform input (analyzer.rs:327)
    wsamples = np.sin(np.arange(44100, dtype=np.float32) / np.float32(44100.0)).astype(np.float32)
    g["waveform_15s"] = O.get_waveform(wsamples, 15.0)
    # test_loudness_measurements input (analyzer.rs:366-368)
    i = np.arange(88200, dtype=np.float32)
    loud = (np.float32(0.1) * np.sin(np.float32(440.0) * np.float32(2.0) * np.float32(np.pi) * (i / np.float32(44100.0)))).astype(np.float32)
    a = O.Analyzer()
    a.add_samples(loud)
    g["loudness_test"] = np.array([a.get_integrated_lufs(), a.get_shortterm_lufs(), a.get_loudness_range(), *a.get_true_peak()])
    # cfg1: 10 s stereo sweep at 48 kHz, fed the way calculate_integrated_lufs and the player tick do
    sw = sweep_stereo(10.0, 48000)
    a = O.Analyzer()
    a.create_loudness_meter(2, 48000)
    g["sweep_integrated_oneshot"] = np.array([a.calculate_integrated_lufs(2, sw)])
    m = O.EbuR128(2, 48000)
    for off in range(0, sw.size, 9600):
        m.add_frames_f32(sw[off:off + 9600])
    g["sweep_scalars"] = np.array([m.loudness_momentary(), m.loudness_shortterm(), m.loudness_global(),
                                   m.loudness_range(), m.true_peak(0), m.true_peak(1)])
    mid, side = O.mid_side(sw)
    for pos in (16384, 200000, 440000):
        g[f"sweep_mid_{pos}"] = O.get_fft(mid[pos - 16384:pos], 48000)[:, 1]
        g[f"sweep_side_{pos}"] = O.get_fft(side[pos - 16384:pos], 48000)[:, 1]
    g["sweep_waveform_10s"] = O.get_waveform(sw, 10.0)[:, 1].astype(np.float32)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(out, **g)
    print(out, os.path.getsize(out), "bytes")
    for k, v in g.items():
        print(k, v.shape, v.ravel()[:3])


if __name__ == "__main__":
    main()
