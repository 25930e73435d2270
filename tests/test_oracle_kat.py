"""Pins the CPU oracle (oracle/) — the parity checker — against everything available without the
reference binary: the reference's own unit tests restated verbatim (reference src/analyzer.rs:185-399),
the ITU-R BS.1770-4 48 kHz coefficient table, EBU Tech 3341 / 3342 synthetic known answers, an f64 FFT,
an independent numpy restatement of the first-party index arithmetic, and the committed golden vectors.
"""
import os

import numpy as np
import pytest

from tests.signals import ref_sine_f32, sweep_stereo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")


def tone_stereo(levels_db, durations_s, rate=48000, freq=1000.0, channels=2):
    parts = []
    t0 = 0
    for lv, d in zip(levels_db, durations_s):
        n = int(round(d * rate))
        t = (np.arange(n) + t0) / rate
        parts.append((10 ** (lv / 20.0)) * np.sin(2 * np.pi * freq * t))
        t0 += n
    mono = np.concatenate(parts).astype(np.float32)
    return np.repeat(mono[:, None], channels, axis=1).ravel()


# ---- the reference's own tests, restated against the oracle -------------------------------------
def test_ref_get_fft_nonempty(oracle):
    a = oracle.Analyzer()
    r = a.get_fft(ref_sine_f32(440.0))
    assert len(r) == 7423


def test_ref_dbfs_calibration(oracle):
    res = np.float32(44100) / np.float32(16384.0)
    target_bin = int(np.round(np.float32(1000.0) / res))
    assert target_bin == 372
    r = oracle.Analyzer().get_fft(ref_sine_f32(float(np.float32(target_bin) * res)))
    assert -1.0 <= r[:, 1].max() <= 1.0


def test_ref_pink_noise_compensation(oracle):
    res = np.float32(44100) / np.float32(16384.0)
    a = oracle.Analyzer()
    mx = []
    for f in (1000.0, 125.0):
        b = int(np.round(np.float32(f) / res))
        mx.append(a.get_fft(ref_sine_f32(float(np.float32(b) * res)))[:, 1].max())
    assert -10.5 <= mx[1] - mx[0] <= -8.0


def test_ref_get_waveform(oracle):
    s = np.sin(np.arange(44100, dtype=np.float32) / np.float32(44100.0)).astype(np.float32)
    w = oracle.Analyzer.get_waveform(s, 15.0)
    assert len(w) == 30000
    i = np.arange(15000)
    assert np.array_equal(w[0::2, 0], i) and np.array_equal(w[1::2, 0], i)
    assert np.all(w[0::2, 1] <= w[1::2, 1])


def test_ref_loudness_measurements(oracle):
    i = np.arange(88200, dtype=np.float32)
    s = (np.float32(0.1) * np.sin(np.float32(440.0) * np.float32(2.0) * np.float32(np.pi) * (i / np.float32(44100.0)))).astype(np.float32)
    a = oracle.Analyzer()
    a.add_samples(s)
    lufs = a.get_integrated_lufs()
    assert -100.0 < lufs < 0.0
    l, r = a.get_true_peak()
    assert 0.0 <= l <= 1.0 and 0.0 <= r <= 1.0


def test_ref_analyzer_reinit(oracle):
    a = oracle.Analyzer()
    a.create_loudness_meter(1, 48000)
    a.create_loudness_meter(6, 96000)
    with pytest.raises(oracle.OracleError):
        a.create_loudness_meter(0, 48000)
    with pytest.raises(oracle.OracleError):
        a.create_loudness_meter(65, 48000)
    with pytest.raises(oracle.OracleError):
        a.create_loudness_meter(2, 15)
    # mono meter: get_true_peak asks for channel 1 -> InvalidChannelIndex (tui.rs:950-956 shows the error)
    a.create_loudness_meter(1, 48000)
    with pytest.raises(oracle.OracleError) as e:
        a.get_true_peak()
    assert e.value.code == oracle.ERR_INVALID_CHANNEL_INDEX


# ---- standards-based known answers ---------------------------------------------------------------
def test_bs1770_coefficient_table(oracle):
    b, a = oracle.EbuR128(2, 48000).coeffs()
    pb = np.array([1.53512485958697, -2.69169618940638, 1.19839281085285])
    pa = np.array([1.0, -1.69065929318241, 0.73248077421585])
    rb = np.array([1.0, -2.0, 1.0])
    ra = np.array([1.0, -1.99004745483398, 0.99007225036621])
    assert np.allclose(b, np.convolve(pb, rb), rtol=0, atol=2e-14)
    assert np.allclose(a, np.convolve(pa, ra), rtol=0, atol=2e-14)


def test_interpolator_taps(oracle):
    assert oracle.interp_taps(44100) == (4, [1, 12, 12, 12], 37)
    assert oracle.interp_taps(48000) == (4, [1, 12, 12, 12], 37)
    assert oracle.interp_taps(96000) == (2, [1, 24, 0, 0], 25)
    assert oracle.interp_taps(191999)[0] == 2
    assert oracle.interp_taps(192000) == (0, [0, 0, 0, 0], 0)


def test_histogram_tables(oracle):
    assert oracle.histogram_boundary(0) == pytest.approx(10 ** ((-70 + 0.691) / 10), rel=1e-15)
    assert oracle.histogram_energy(0) == pytest.approx(10 ** ((-69.95 + 0.691) / 10), rel=1e-15)
    assert oracle.histogram_boundary(1000) == pytest.approx(10 ** ((30 + 0.691) / 10), rel=1e-15)


@pytest.mark.parametrize("level", [-23.0, -33.0])
def test_ebu3341_case1_2(oracle, level):
    m = oracle.EbuR128(2, 48000)
    m.add_frames_f32(tone_stereo([level], [20.0]))
    for v in (m.loudness_momentary(), m.loudness_shortterm(), m.loudness_global()):
        assert abs(v - level) <= 0.1


def test_ebu3341_case3_4_5(oracle):
    cases = [
        ([-36, -23, -36], [10, 60, 10]),
        ([-72, -36, -23, -36, -72], [10, 10, 60, 10, 10]),
        ([-26, -20, -26], [20, 20.1, 20]),
    ]
    for lv, du in cases:
        m = oracle.EbuR128(2, 48000)
        m.add_frames_f32(tone_stereo(lv, du))
        assert abs(m.loudness_global() - (-23.0)) <= 0.1, (lv, m.loudness_global())


def test_ebu3341_case6_surround(oracle):
    # case 6 (5.0): L/R -28, C -24, Ls/Rs -30 dBFS -> -23.0 LUFS; fed as 6 channels with a silent LFE at index 3
    rate, n = 48000, 48000 * 20
    t = np.arange(n) / rate
    s = np.sin(2 * np.pi * 1000 * t)
    x = np.zeros((n, 6), dtype=np.float32)
    for c, lv in ((0, -28), (1, -28), (2, -24), (4, -30), (5, -30)):
        x[:, c] = 10 ** (lv / 20) * s
    m = oracle.EbuR128(6, rate)
    m.add_frames_f32(x.ravel())
    assert abs(m.loudness_global() - (-23.0)) <= 0.1


def test_ebu3342_lra(oracle):
    cases = [([-20, -30], [20, 20], 10.0), ([-20, -15], [20, 20], 5.0), ([-40, -20], [20, 20], 20.0),
             ([-50, -35, -20, -35, -50], [20, 20, 20, 20, 20], 15.0)]
    for lv, du, want in cases:
        m = oracle.EbuR128(2, 48000)
        m.add_frames_f32(tone_stereo(lv, du))
        assert abs(m.loudness_range() - want) <= 1.0, (lv, m.loudness_range())


def test_ebu3341_true_peak(oracle):
    rate = 48000
    n = rate * 2
    i = np.arange(n)
    for phase_deg, want_db in ((0.0, -6.0), (45.0, -6.0)):
        x = 0.5 * np.sin(2 * np.pi * (rate / 4) * i / rate + np.deg2rad(phase_deg))
        m = oracle.EbuR128(2, rate)
        m.add_frames_f32(np.repeat(x.astype(np.float32)[:, None], 2, 1).ravel())
        tp_db = 20 * np.log10(m.true_peak(0))
        assert want_db - 0.4 <= tp_db <= want_db + 0.2, (phase_deg, tp_db)


def test_chunking_invariance(oracle):
    x = sweep_stereo(4.0, 48000)
    a = oracle.EbuR128(2, 48000)
    a.add_frames_f32(x)
    b = oracle.EbuR128(2, 48000)
    rng = np.random.default_rng(3)
    off = 0
    while off < x.size:
        n = 2 * int(rng.integers(1, 20000))
        b.add_frames_f32(x[off:off + n])
        off += n
    assert a.loudness_global() == b.loudness_global()
    assert a.loudness_shortterm() == b.loudness_shortterm()
    assert a.true_peak(0) == b.true_peak(0)


def test_ragged_and_empty(oracle):
    m = oracle.EbuR128(2, 48000)
    m.add_frames_f32(np.zeros(0, dtype=np.float32))
    assert m.loudness_global() == -np.inf
    assert m.loudness_shortterm() == -np.inf
    assert m.loudness_range() == 0.0
    with pytest.raises(oracle.OracleError):
        m.add_frames_f32(np.zeros(3, dtype=np.float32))
    a = oracle.Analyzer()
    assert a.calculate_integrated_lufs(2, np.zeros(44100 * 2 + 1, dtype=np.float32)) is None  # ragged tail chunk
    assert a.calculate_integrated_lufs(0, np.zeros(10, dtype=np.float32)) is None
    assert a.calculate_integrated_lufs(2, np.zeros(0, dtype=np.float32)) == -np.inf


# ---- spectrum path -------------------------------------------------------------------------------
def test_rfft_against_f64(oracle):
    rng = np.random.default_rng(1)
    for n in (2, 4, 64, 1024, 8192, 16384, 32768):
        x = rng.standard_normal(n).astype(np.float32)
        ref = np.abs(np.fft.rfft(x.astype(np.float64)))
        got = oracle.rfft_mag(x)
        assert np.max(np.abs(got - ref)) <= 1e-6 * ref.max()


def test_hann_is_periodic_f32(oracle):
    w = oracle.hann_window(np.ones(16384, dtype=np.float32))
    ref = 0.5 * (1 - np.cos(2 * np.pi * np.arange(16384) / 16384))
    assert w[0] == 0.0 and np.max(np.abs(w - ref)) < 3e-7
    assert w[8192] == 1.0


def test_get_fft_errors(oracle):
    a = oracle.Analyzer()
    codes = {}
    for name, arr in (("short", np.zeros(1)), ("npow2", np.zeros(1000)), ("nan", np.full(1024, np.nan)),
                      ("inf", np.r_[0.0, np.full(1023, np.inf)]), ("toolong", np.zeros(65536))):
        with pytest.raises(oracle.OracleError) as e:
            a.get_fft(arr.astype(np.float32))
        codes[name] = e.value.code
    assert codes == {"short": 4, "npow2": 7, "nan": 5, "inf": 6, "toolong": 7}
    a.create_loudness_meter(2, 22050)
    with pytest.raises(oracle.OracleError) as e:
        a.get_fft(np.zeros(1024, dtype=np.float32))
    assert e.value.code == 8
    a.create_loudness_meter(2, 48000)
    z = a.get_fft(np.zeros(1024, dtype=np.float32))
    fr = (np.arange(1, 427, dtype=np.float32) * np.float32(48000 / 1024)).astype(np.float64)
    assert z.shape == (426, 2) and np.allclose(z[:, 1], -150.0 + 10 * np.log10(fr / 1000.0), rtol=0, atol=1e-12)


def test_scale_to_dbfs(oracle):
    assert oracle.scale_to_dbfs(0.0, 16384.0) == -150.0
    assert oracle.scale_to_dbfs(4096.0, 16384.0) == 0.0
    assert abs(oracle.scale_to_dbfs(409.6, 16384.0) + 20.0) < 1e-5


# ---- first-party index arithmetic: independent numpy restatement ---------------------------------
def np_waveform(samples, window_s):
    window = int(window_s * 1000.0)
    spp = np.float64(len(samples)) / np.float64(window)
    pts = []
    for i in range(window):
        start = int(np.float64(i) * spp)
        end = min(int(np.ceil(np.float64(i + 1) * spp)), len(samples))
        if start >= len(samples):
            break
        ch = samples[start:end]
        pts.append((i, ch.min() if len(ch) else 0.0))
        pts.append((i, ch.max() if len(ch) else 0.0))
    return np.array(pts, dtype=np.float64).reshape(-1, 2)


@pytest.mark.parametrize("n,win", [(44100, 15.0), (960000, 10.0), (1000, 15.0), (7, 0.003), (12345, 1.2345), (100, 0.0)])
def test_waveform_index_math(oracle, n, win):
    rng = np.random.default_rng(n)
    s = rng.uniform(-1, 1, n).astype(np.float32)
    got = oracle.get_waveform(s, win)
    ref = np_waveform(s, win)
    assert got.shape == ref.shape and np.array_equal(got, ref)


def test_waveform_nan_and_empty(oracle):
    s = np.array([0.5, np.nan, -0.25, np.nan], dtype=np.float32)
    w = oracle.get_waveform(s, 0.001)
    assert np.array_equal(w, [[0, -0.25], [0, 0.5]])   # f32::min/max ignore NaN
    assert oracle.get_waveform(np.zeros(0, dtype=np.float32), 1.0).shape == (0, 2)


def test_mid_side(oracle):
    rng = np.random.default_rng(5)
    s = rng.uniform(-1, 1, 2001).astype(np.float32)
    mid, side = oracle.mid_side(s)
    l, r = s[0:2000:2], s[1:2000:2]
    assert len(mid) == 1000
    assert np.array_equal(mid, (l + r) / np.float32(2)) and np.array_equal(side, (l - r) / np.float32(2))


# ---- golden vectors ------------------------------------------------------------------------------
def test_golden_regression(oracle):
    g = np.load(GOLD)
    f = g["fft_freqs"]
    for name, fr in zip(("440", "1k", "125"), f):
        assert np.array_equal(oracle.get_fft(ref_sine_f32(float(fr)), 44100), g[f"fft_{name}"])
    s = np.sin(np.arange(44100, dtype=np.float32) / np.float32(44100.0)).astype(np.float32)
    assert np.array_equal(oracle.get_waveform(s, 15.0), g["waveform_15s"])
    sw = sweep_stereo(10.0, 48000)
    a = oracle.Analyzer()
    a.create_loudness_meter(2, 48000)
    assert a.calculate_integrated_lufs(2, sw) == g["sweep_integrated_oneshot"][0]
    m = oracle.EbuR128(2, 48000)
    for off in range(0, sw.size, 9600):
        m.add_frames_f32(sw[off:off + 9600])
    got = [m.loudness_momentary(), m.loudness_shortterm(), m.loudness_global(), m.loudness_range(), m.true_peak(0), m.true_peak(1)]
    assert np.array_equal(got, g["sweep_scalars"])


# ---- formats either side of the path: oracle/capture_ref.py -------------------------

def test_pcm_rules_known_values(oracle):
    """symphonia-core `FromSample<S> for f32`: full-scale and mid-scale values of every integer format."""
    R = oracle.capture_ref
    assert list(R.pcm_to_f32(bytes([0, 128, 255]), "u8")) == [-1.0, 0.0, 0.9921875]
    assert list(R.pcm_to_f32(bytes([0x80, 0x00, 0x7f]), "s8")) == [-1.0, 0.0, 0.9921875]
    assert list(R.pcm_to_f32(np.array([-32768, 0, 32767, 1], "<i2").tobytes(), "s16le")) == [-1.0, 0.0, 0.999969482421875, 2.0 ** -15]
    assert list(R.pcm_to_f32(np.array([-32768, 32767], ">i2").tobytes(), "s16be")) == [-1.0, 0.999969482421875]
    assert list(R.pcm_to_f32(bytes([0x00, 0x00, 0x80, 0xff, 0xff, 0x7f, 0x01, 0x00, 0x00]), "s24le")) == [-1.0, 1 - 2.0 ** -23, 2.0 ** -23]
    assert list(R.pcm_to_f32(bytes([0x80, 0x00, 0x00, 0x7f, 0xff, 0xff]), "s24be")) == [-1.0, 1 - 2.0 ** -23]
    # 2^31 - 1 is not representable in f32: the single rounding goes UP to 1.0; 2^31 - 129 is the tie that stays below
    got = R.pcm_to_f32(np.array([-2 ** 31, 2 ** 31 - 1, 2 ** 31 - 128, 2 ** 31 - 129, 1], "<i4").tobytes(), "s32le")
    assert list(got) == [-1.0, 1.0, 1.0 - 2.0 ** -24, 1.0 - 2.0 ** -24, 2.0 ** -31]
    bits = np.array([0x7fc00001, 0xff800000, 0x00000001, 0x3f800000], "<u4")
    assert np.array_equal(R.pcm_to_f32(bits.tobytes(), "f32le").view(np.uint32), bits)       # identity, NaN payload kept
    assert np.array_equal(R.pcm_to_f32(bits.astype(">u4").tobytes(), "f32be").view(np.uint32), bits)
    assert list(R.pcm_to_f32(np.array([0.1, 1e39, -1e-46], "<f8").tobytes(), "f64le")) == [np.float32(0.1), np.inf, -0.0]


def test_pcm_rules_insensitive_to_spelling(oracle):
    """The crate is un-vendored; both plausible spellings of each rule give the same bits (exhaustive for 8/16-bit)."""
    R = oracle.capture_ref
    rng = np.random.default_rng(0)
    cases = {"u8": np.arange(256, dtype=np.uint8).tobytes(), "s16le": np.arange(-32768, 32768).astype("<i2").tobytes(),
             "s16be": np.arange(-32768, 32768).astype(">i2").tobytes(),
             "s24le": rng.integers(0, 256, 3 * 200000, dtype=np.uint8).tobytes(),
             "s24be": rng.integers(0, 256, 3 * 200000, dtype=np.uint8).tobytes()}
    for fmt, raw in cases.items():
        assert np.array_equal(R.pcm_to_f32(raw, fmt).view(np.uint32), R.pcm_to_f32_alt(raw, fmt).view(np.uint32)), fmt
    # s32: `(s as f64 / 2^31) as f32` vs `s as f32 / 2^31` — one rounding either way
    raw = np.concatenate([rng.integers(-2 ** 31, 2 ** 31, 500000), [2 ** 31 - 1, -2 ** 31, 2 ** 24 + 1, 2 ** 25 + 3]]).astype("<i4").tobytes()
    assert np.array_equal(R.pcm_to_f32(raw, "s32le").view(np.uint32), R.pcm_to_f32_alt(raw, "s32le").view(np.uint32))


def test_capture_ring_mono_quirk(oracle):
    """audio_capture.rs:43-48: `[x0, 0, x1, 0, x2]` — 2n-1 values per callback, so the L/R parity flips."""
    r = oracle.capture_ref.RingRef(10)
    r.callback([1, 2, 3], True)
    assert list(r.to_vec()) == [0, 0, 0, 0, 0, 1, 0, 2, 0, 3]
    r.callback([4, 5], True)
    assert list(r.to_vec()) == [0, 0, 1, 0, 2, 0, 3, 4, 0, 5]      # 3 and 4 now share a stereo frame
    r.callback(np.arange(10, 23), False)                             # oversize push: the last 10 survive
    assert list(r.to_vec()) == list(range(13, 23))
    r.callback([], True)
    assert list(r.to_vec()) == list(range(13, 23))


def test_mic_tick_oracle_composition(oracle):
    """tui.rs:1427-1480 on the oracle: slices at 15*rate / 30*rate, waveform of mid over the whole ring."""
    rate = 44100
    ring = oracle.capture_ref.RingRef(30 * rate)
    t = np.arange(15 * rate) / rate
    x = np.empty(30 * rate, dtype=np.float32)
    x[0::2] = 0.5 * np.sin(2 * np.pi * 1000 * t)
    x[1::2] = 0.5 * np.sin(2 * np.pi * 1000 * t)
    ring.callback(x, False)
    a = oracle.Analyzer()
    mid_fft, side_fft, wave, st, err = oracle.capture_ref.mic_tick(ring.to_vec(), a)
    assert err is None and mid_fft.shape == (7423, 2) and wave.shape == (30000, 2)
    assert side_fft[:, 1].max() == -150.0 + 10 * np.log10(20000.0 / 1000.0) or side_fft[:, 1].max() < -100    # L == R: silent side
    assert abs(mid_fft[np.argmax(mid_fft[:, 1]), 1] - (20 * np.log10(0.5 * 4 / 2 / 2) + 0.0)) < 2.0   # Hann coherent gain 0.5
    # 8192 frames of a -6 dBFS-peak stereo 1 kHz tone in a 3 s window: 10*log10(2 * 0.125 * 8192 / 132300) = -18.10
    assert abs(st - (-18.10)) < 0.05


def test_golden_capture_regression(oracle):
    """oracle/capture_ref.py against the committed vectors (tests/golden/make_golden_capture.py)."""
    from tests.golden import make_golden_capture as M
    g = np.load(os.path.join(os.path.dirname(GOLD), "golden_capture_v1.npz"))
    for fmt in M.FORMATS:
        raw = M.pcm_raw(fmt)
        assert np.array_equal(raw, g[f"pcm_{fmt}_raw"])
        assert np.array_equal(oracle.capture_ref.pcm_to_f32(raw.tobytes(), fmt).view(np.uint32), g[f"pcm_{fmt}_f32bits"]), fmt
    r = oracle.capture_ref.RingRef(M.RING_CAP)
    for d, mono in M.ring_pushes():
        r.callback(d, bool(mono))
    assert np.array_equal(r.to_vec().view(np.uint32), g["ring_to_vec"].view(np.uint32))


# ---- independent implementations (SURVEY §8(c) "secondary cross-checks") -----------------------------------
def _program_like(rate, seconds, seed):
    rng = np.random.default_rng(seed)
    n = int(rate * seconds)
    t = np.arange(n) / rate
    left = 0.3 * np.sin(2 * np.pi * 440 * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 0.2 * t)) + 0.02 * rng.standard_normal(n)
    right = 0.2 * np.sin(2 * np.pi * 660 * t) + 0.02 * rng.standard_normal(n)
    return np.stack([left, right], 1).astype(np.float32)


@pytest.mark.parametrize("rate", [44100, 48000, 96000])
def test_oracle_against_scipy_lfilter(oracle, rate):
    """The oracle's DF-II recursion + ring sums against scipy.signal.lfilter (an independent transposed-DF-II in f64)
    and plain mean squares: momentary and short-term loudness agree to 1e-9 LU."""
    from scipy import signal
    x = _program_like(rate, 5.0, rate)
    m = oracle.EbuR128(2, rate)
    m.add_frames_f32(x.ravel())
    b, a = m.coeffs()
    y = signal.lfilter(b, a, x.astype(np.float64), axis=0)
    s100 = (rate + 5) // 10
    for frames, got in ((4 * s100, m.loudness_momentary()), (30 * s100, m.loudness_shortterm())):
        want = -0.691 + 10 * np.log10(np.mean(y[-frames:] ** 2, axis=0).sum())
        assert abs(got - want) < 1e-9


def test_oracle_against_torchaudio_loudness(oracle):
    """torchaudio.functional.loudness is an independent BS.1770 integrated-loudness implementation (exact block list
    instead of ebur128's 0.1 LU histogram, its own K-weighting design): agreement within the histogram's quantisation."""
    import torch
    import torchaudio
    for seed, rate in ((1, 48000), (2, 44100)):
        x = _program_like(rate, 10.0, seed)
        m = oracle.EbuR128(2, rate)
        m.add_frames_f32(x.ravel())
        ta = torchaudio.functional.loudness(torch.from_numpy(np.ascontiguousarray(x.T)), rate).item()
        assert abs(m.loudness_global() - ta) < 0.1, (m.loudness_global(), ta)


def test_oracle_true_peak_against_polyphase_resampler(oracle):
    """ebur128's 49-tap interpolator against scipy's (longer) polyphase resampler: inter-sample peaks within 2 %, and a
    fs/4 tone sampled 45 degrees off its crests reads +3 dB over the sample peak (the EBU 3341 true-peak construction)."""
    from scipy import signal
    x = _program_like(48000, 4.0, 3)
    m = oracle.EbuR128(2, 48000)
    m.add_frames_f32(x.ravel())
    up = np.abs(signal.resample_poly(x.astype(np.float64), 4, 1, axis=0)).max(axis=0)
    for c in range(2):
        assert abs(m.true_peak(c) / up[c] - 1.0) < 0.02
    n = np.arange(48000)
    tone = (0.5 * np.sin(2 * np.pi * 0.25 * n + np.pi / 4)).astype(np.float32)       # samples at +-0.3536, crests at 0.5
    m = oracle.EbuR128(1, 48000)
    m.add_frames_f32(tone)
    assert abs(m.sample_peak(0) - 0.5 / np.sqrt(2)) < 1e-6
    assert -0.4 <= 20 * np.log10(m.true_peak(0) / 0.5) <= 0.2          # EBU Tech 3341's true-peak tolerance


@pytest.mark.parametrize("sr", [44100, 48000, 96000])
def test_ref_analyze_microphone_input(oracle, sr):
    """The reference's own microphone-tick tests (tui.rs:2271-2368) on the oracle: `create_test_app` gives a ring of
    44100 * 30 values (tui.rs:2200) and a default (2 ch, 44100 Hz) device analyzer; the test enqueues sr * 30 samples of
    a sine and calls analyze_microphone_input()."""
    from tests.signals import ref_mic_test_assertions, ref_mic_test_ring_fill
    ring = oracle.capture_ref.RingRef(44100 * 30)
    ring.callback(ref_mic_test_ring_fill(sr), False)          # buffer.clear(); buffer.enqueue(sample) x sr*30: ends full
    mid, side, wave, st, err = oracle.capture_ref.mic_tick(ring.to_vec(), oracle.Analyzer())
    ref_mic_test_assertions(sr, mid)
    assert err is None and mid.shape == (7423, 2) and wave.shape == (30000, 2) and np.isfinite(st)
