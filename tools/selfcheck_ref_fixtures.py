"""Harness self-check for tests/test_ref_fixtures.py: writes a file with the SAME keys and generators as the Rust fixture
generator (oracle/ref_rust/src/main.rs), but computed by the CPU oracle, so the comparison code can be exercised before a
Rust toolchain exists.  NOT a reference fixture — never commit its output as tests/golden/ref_v1.npz.

    python tools/selfcheck_ref_fixtures.py /tmp/ref_selfcheck.npz
    SSB_REF_FIXTURES=/tmp/ref_selfcheck.npz python -m pytest tests/test_ref_fixtures.py -q -m "not gpu"
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
from soundscope_b200.synth import ref_sine_f32, sweep_stereo
from tests.test_ref_fixtures import G3_CASES, noise, row, tone_segments


def main():
    out = {}
    for name, f in (("g1_fft_440", 440.0), ("g1_fft_bin372", np.float32(372.0) * np.float32(44100.0) / np.float32(16384.0)),
                    ("g1_fft_125", np.float32(46.0) * np.float32(44100.0) / np.float32(16384.0))):
        x = ref_sine_f32(f)
        out[name + "_in"] = x
        out[name] = np.asarray(O.Analyzer().get_fft(x))
    s = np.sin(np.arange(44100, dtype=np.float32) / np.float32(44100.0)).astype(np.float32)
    out["g1_waveform_in"] = s
    out["g1_waveform"] = np.asarray(O.Analyzer().get_waveform(s, 15.0))
    i = np.arange(88200, dtype=np.float32)
    x = (np.float32(0.1) * np.sin(np.float32(440.0) * np.float32(2.0) * np.float32(np.pi) * (i / np.float32(44100.0)))).astype(np.float32)
    a = O.Analyzer()
    a.add_samples(x)
    out["g1_loudness_in"], out["g1_loudness"] = x, row(a)
    for tag, gain in (("g2_sweep", 0.5), ("g2_sweep_anti", -1.0)):
        x = sweep_stereo(10.0, 48000, 0.5, gain)
        mid, side = (x[0::2] + x[1::2]) / np.float32(2), (x[0::2] - x[1::2]) / np.float32(2)
        a = O.Analyzer()
        a.create_loudness_meter(2, 48000)
        rows, pos, hop = [], 16384 + 2048, 0
        while pos <= x.size:
            a.add_samples(x[pos - 16384:pos])
            rows.append(row(a))
            if hop % 32 == 0 and pos // 2 >= 16384:
                p = pos // 2
                out[f"{tag}_mid_fft_{hop}"] = np.asarray(a.get_fft(mid[p - 16384:p]))
                out[f"{tag}_side_fft_{hop}"] = np.asarray(a.get_fft(side[p - 16384:p]))
            pos += 2048
            hop += 1
        out[tag + "_ticks"] = np.array(rows)
        out[tag + "_oneshot"] = np.array([a.calculate_integrated_lufs(2, x)])
        out[tag + "_waveform"] = np.asarray(a.get_waveform(x, 10.0))
    for name, (rate, segs) in G3_CASES.items():
        a = O.Analyzer()
        a.create_loudness_meter(2, rate)
        x = tone_segments(rate, segs)
        for off in range(0, x.size, rate * 2):
            a.add_samples(x[off:off + rate * 2])
        out[name] = row(a)
    for name, rate, ph in (("g3_tp_fs4_0", 48000, 0.0), ("g3_tp_fs4_45", 48000, 45.0), ("g3_tp_fs4_45_96k", 96000, 45.0),
                           ("g3_tp_fs4_45_192k", 192000, 45.0)):
        v = (0.5 * np.sin(2 * np.pi * 0.25 * np.arange(rate) + ph * np.pi / 180.0)).astype(np.float32)
        x = np.empty(2 * rate, dtype=np.float32)
        x[0::2], x[1::2] = v, (0.5 * v.astype(np.float64)).astype(np.float32)
        a = O.Analyzer()
        a.create_loudness_meter(2, rate)
        a.add_samples(x)
        out[name] = row(a)
    for lg in (1, 4, 9, 12, 13, 14, 15):
        n = 1 << lg
        try:
            out[f"g4_fft_noise_{n}"] = np.asarray(O.Analyzer().get_fft(noise(100 + lg, n)))
        except Exception:
            out[f"g4_fft_noise_{n}_err"] = np.array([1.0])
    for ch, rate in ((1, 48000), (2, 44100), (4, 48000), (5, 48000), (6, 96000), (8, 48000)):
        a = O.Analyzer()
        a.create_loudness_meter(ch, rate)
        base = noise(7 * ch + rate, rate * 5 * ch)
        x = (np.float32(0.3) * base * (np.float32(1.0) - np.float32(0.1) * (np.arange(base.size) % ch).astype(np.float32))).astype(np.float32)
        for off in range(0, x.size, rate * ch):
            a.add_samples(x[off:off + rate * ch])
        tp = a.get_true_peak() if ch >= 2 else (np.nan, np.nan)
        out[f"g4_meter_{ch}ch_{rate}"] = np.array([a.get_shortterm_lufs(), a.get_integrated_lufs(), a.get_loudness_range(), tp[0], tp[1]])
    levels, res = [], []
    for step in range(400):
        db = -23.2 + 0.001 * step
        v = (10.0 ** (db / 20.0) * np.sin(2 * np.pi * 1000.0 * np.arange(48000 * 3) / 48000.0)).astype(np.float32)
        a = O.Analyzer()
        a.create_loudness_meter(2, 48000)
        a.add_samples(np.repeat(v, 2))
        levels.append(db)
        res.append(a.get_integrated_lufs())
    out["g5_level_db"], out["g5_integrated"] = np.array(levels), np.array(res)
    np.savez_compressed(sys.argv[1], **out)
    print(len(out), "arrays ->", sys.argv[1])


if __name__ == "__main__":
    main()
