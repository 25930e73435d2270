"""Parity against fixtures produced by the REFERENCE ITSELF (oracle/ref_rust: the reference's src/analyzer.rs compiled
against ebur128 0.1.10 / spectrum-analyzer 1.7.0), golden sets G1-G5 of SURVEY.md section 8c.

tests/golden/ref_v1.npz does not exist until someone runs the generator on a machine with a Rust toolchain (there is
none in this image or on the GPU boxes: profiles/r2_probe.txt).  While it is absent every test here emits a
"PARITY UNPINNED" warning and skips; when it is present the CPU oracle (and, under -m gpu, the CUDA path through the
C ABI) is compared with the reference's numbers at the tolerances BASELINE.json states: bit-exact waveform, 1e-4 LU,
1e-5 * max|X| FFT magnitude, 2e-6 relative true peak.
"""
import os
import warnings

import numpy as np
import pytest

from tests.signals import sweep_stereo

REF = os.environ.get("SSB_REF_FIXTURES") or os.path.join(os.path.dirname(__file__), "golden", "ref_v1.npz")
UNPINNED = ("PARITY UNPINNED: tests/golden/ref_v1.npz is absent (no Rust toolchain here; see oracle/ref_rust/README.md) — "
            "third-party arithmetic is checked against published algorithms and standards only")
LU_TOL, TP_RTOL, FFT_REL = 1e-4, 2e-6, 1e-5


def ref():
    if not os.path.exists(REF):
        warnings.warn(UNPINNED)
        pytest.skip(UNPINNED)
    return np.load(REF)


def noise(seed, n):
    """splitmix64 -> uniform [-1, 1) f32, bit for bit the generator of oracle/ref_rust/src/main.rs"""
    with np.errstate(over="ignore"):
        s = (np.uint64(seed) + np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
        z = s
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(40)).astype(np.float64) / float(1 << 24) * 2.0 - 1.0).astype(np.float32)


def tone_segments(rate, segs):
    out, n0 = [], 0
    for db, secs in segs:
        amp = 10.0 ** (db / 20.0)
        n = int(secs * rate)
        v = (amp * np.sin(2 * np.pi * 1000.0 * (n0 + np.arange(n)) / rate)).astype(np.float32)
        out.append(np.repeat(v, 2))
        n0 += n
    return np.concatenate(out)


G3_CASES = {
    "g3_3341_1": (48000, [(-23.0, 20.0)]), "g3_3341_2": (48000, [(-33.0, 20.0)]),
    "g3_3341_3": (48000, [(-36.0, 10.0), (-23.0, 60.0), (-36.0, 10.0)]),
    "g3_3341_4": (48000, [(-72.0, 10.0), (-36.0, 10.0), (-23.0, 60.0), (-36.0, 10.0), (-72.0, 10.0)]),
    "g3_3342_1": (48000, [(-20.0, 20.0), (-30.0, 20.0)]), "g3_3342_2": (48000, [(-20.0, 20.0), (-15.0, 20.0)]),
    "g3_3342_3": (44100, [(-40.0, 20.0), (-20.0, 20.0)]),
    "g3_3342_4": (96000, [(-50.0, 20.0), (-35.0, 20.0), (-20.0, 20.0), (-35.0, 20.0), (-50.0, 20.0)]),
}


def row(a):
    l, r = a.get_true_peak()
    return np.array([a.get_shortterm_lufs(), a.get_integrated_lufs(), a.get_loudness_range(), l, r])


def check_row(got, want):
    for i in range(3):
        assert (np.isneginf(got[i]) and np.isneginf(want[i])) or abs(got[i] - want[i]) <= LU_TOL, (i, got, want)
    assert np.all(np.abs(got[3:] - want[3:]) <= TP_RTOL * np.maximum(want[3:], 1e-30)), (got, want)


def check_fft(got, want):
    got, want = np.asarray(got, dtype=np.float64).reshape(-1, 2), np.asarray(want, dtype=np.float64).reshape(-1, 2)
    assert got.shape == want.shape
    if want.size == 0:
        return
    assert np.array_equal(got[:, 0], want[:, 0])          # chart x: first-party f64 arithmetic, exact
    mg, mw = 10.0 ** (got[:, 1] / 20.0), 10.0 ** (want[:, 1] / 20.0)
    assert np.max(np.abs(mg - mw)) <= FFT_REL * mw.max()


def run_all(make_analyzer, R):
    """Feeds every fixture's input to an Analyzer-shaped object (oracle.Analyzer or soundscope_b200.Analyzer)."""
    # G1: the reference's own unit-test inputs
    for name in ("g1_fft_440", "g1_fft_bin372", "g1_fft_125"):
        check_fft(np.asarray(make_analyzer().get_fft(R[name + "_in"])), R[name])
    a = make_analyzer()
    assert np.array_equal(np.asarray(a.get_waveform(R["g1_waveform_in"], 15.0)), R["g1_waveform"])
    a.add_samples(R["g1_loudness_in"])
    check_row(row(a), R["g1_loudness"])
    # G2: the cfg1 sweep in the player's overlapping feed
    for tag, gain in (("g2_sweep", 0.5), ("g2_sweep_anti", -1.0)):
        x = sweep_stereo(10.0, 48000, 0.5, gain)
        a = make_analyzer()
        a.create_loudness_meter(2, 48000)
        ticks, pos, hop = R[tag + "_ticks"], 16384 + 2048, 0
        while pos <= x.size:
            a.add_samples(x[pos - 16384:pos])
            if hop % 8 == 0:
                check_row(row(a), ticks[hop])
            pos += 2048
            hop += 1
        assert hop == len(ticks)
        got = a.calculate_integrated_lufs(2, x)
        assert abs(got - R[tag + "_oneshot"][0]) <= LU_TOL
        assert np.array_equal(np.asarray(a.get_waveform(x, 10.0)), R[tag + "_waveform"])
        mid, side = (x[0::2] + x[1::2]) / np.float32(2), (x[0::2] - x[1::2]) / np.float32(2)
        for h in range(0, hop, 32):
            p = (16384 + 2048 + 2048 * h) // 2
            if p < 16384:
                continue
            check_fft(np.asarray(a.get_fft(mid[p - 16384:p])), R[f"{tag}_mid_fft_{h}"])
            check_fft(np.asarray(a.get_fft(side[p - 16384:p])), R[f"{tag}_side_fft_{h}"])
    # G3: EBU-style tone sequences and true-peak phase cases
    for name, (rate, segs) in G3_CASES.items():
        a = make_analyzer()
        a.create_loudness_meter(2, rate)
        x = tone_segments(rate, segs)
        for off in range(0, x.size, rate * 2):
            a.add_samples(x[off:off + rate * 2])
        check_row(row(a), R[name])
    for name, rate, ph in (("g3_tp_fs4_0", 48000, 0.0), ("g3_tp_fs4_45", 48000, 45.0), ("g3_tp_fs4_45_96k", 96000, 45.0),
                           ("g3_tp_fs4_45_192k", 192000, 45.0)):
        v = (0.5 * np.sin(2 * np.pi * 0.25 * np.arange(rate) + ph * np.pi / 180.0)).astype(np.float32)
        x = np.empty(2 * rate, dtype=np.float32)
        x[0::2], x[1::2] = v, (0.5 * v.astype(np.float64)).astype(np.float32)
        a = make_analyzer()
        a.create_loudness_meter(2, rate)
        a.add_samples(x)
        check_row(row(a), R[name])
    # G4: noise spectra (error cases included) and multichannel meters
    for lg in (1, 4, 9, 12, 13, 14, 15):
        n = 1 << lg
        a = make_analyzer()
        if f"g4_fft_noise_{n}_err" in R.files:
            with pytest.raises(Exception):
                a.get_fft(noise(100 + lg, n))
        else:
            check_fft(np.asarray(a.get_fft(noise(100 + lg, n))), R[f"g4_fft_noise_{n}"])
    for ch, rate in ((1, 48000), (2, 44100), (4, 48000), (5, 48000), (6, 96000), (8, 48000)):
        a = make_analyzer()
        a.create_loudness_meter(ch, rate)
        base = noise(7 * ch + rate, rate * 5 * ch)
        x = (np.float32(0.3) * base * (np.float32(1.0) - np.float32(0.1) * (np.arange(base.size) % ch).astype(np.float32))).astype(np.float32)
        for off in range(0, x.size, rate * ch):
            a.add_samples(x[off:off + rate * ch])
        want = R[f"g4_meter_{ch}ch_{rate}"]
        got = np.array([a.get_shortterm_lufs(), a.get_integrated_lufs(), a.get_loudness_range()])
        assert np.all(np.abs(got - want[:3]) <= LU_TOL), (ch, rate, got, want)
        if ch >= 2:
            l, r = a.get_true_peak()
            assert abs(l - want[3]) <= TP_RTOL * want[3] and abs(r - want[4]) <= TP_RTOL * want[4]
    # G5: tone levels stepped across histogram bin edges — the 0.1 LU jumps must fall on the same steps
    for db, want in zip(R["g5_level_db"], R["g5_integrated"]):
        amp = 10.0 ** (db / 20.0)
        v = (amp * np.sin(2 * np.pi * 1000.0 * np.arange(48000 * 3) / 48000.0)).astype(np.float32)
        a = make_analyzer()
        a.create_loudness_meter(2, 48000)
        a.add_samples(np.repeat(v, 2))
        assert abs(a.get_integrated_lufs() - want) <= LU_TOL, (db, a.get_integrated_lufs(), want)


def test_generator_inputs_are_reproducible():
    """The Python mirrors of the Rust generators (no fixture needed): splitmix64 noise known answers."""
    x = noise(0, 4)
    # first outputs of splitmix64(seed 0): 0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, ... -> top 24 bits -> [-1, 1)
    want = np.array([0xE220A8 / (1 << 24) * 2 - 1, 0x6E789E / (1 << 24) * 2 - 1], dtype=np.float32)
    assert np.array_equal(x[:2], want)


def test_oracle_matches_reference_fixtures(oracle):
    R = ref()
    run_all(lambda: oracle.Analyzer(), R)


@pytest.mark.gpu
def test_gpu_matches_reference_fixtures(ssb, cuda):
    R = ref()
    run_all(lambda: ssb.Analyzer(), R)
