"""GPU parity for the formats either side of the path through the C ABI:

  decoded PCM -> f32           bit-exact for every format (uint32 views compared, NaN payloads of f32 input included)
  capture ring                  bit-exact to_vec() after any push sequence (mono up-mix quirk, wrap, oversize pushes)
  microphone tick               waveform bit-exact; spectra within the FFT tolerance of test_gpu_spectrum.py;
                                short-term LUFS within 1e-9 LU of the oracle fed the same ring snapshots
"""
import numpy as np
import pytest

from tests.test_gpu_spectrum import assert_db_close

pytestmark = pytest.mark.gpu

FORMATS = ["u8", "s8", "s16le", "s16be", "s24le", "s24be", "s32le", "s32be", "f32le", "f32be", "f64le", "f64be"]


def _raw(fmt, n, seed):
    """n samples of raw PCM: random bytes for integer / f32 formats (every bit pattern is legal), finite + inf
    doubles for f64 (a NaN's f64 -> f32 payload is hardware-defined), extremes first."""
    rng = np.random.default_rng(seed)
    from oracle.capture_ref import pcm_bytes_per_sample
    bps = pcm_bytes_per_sample(fmt)
    if fmt.startswith("f64"):
        v = np.concatenate([
            np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, 1e-45, 3.4e38, 3.5e38, 1e-300, 0.1, 1 + 2 ** -24, 1 + 2 ** -23 + 2 ** -24]),
            rng.standard_normal(max(n, 16)), rng.standard_normal(max(n, 16)) * 1e-40])[:n] if n else np.zeros(0)
        return v.astype(">f8" if fmt.endswith("be") else "<f8").tobytes()
    b = rng.integers(0, 256, size=n * bps, dtype=np.uint8)
    ext = {1: [0x00, 0x7f, 0x80, 0xff], 2: [0x00, 0x80, 0xff, 0x7f, 0x00, 0x00, 0x80, 0x00],
           3: [0x00, 0x00, 0x80, 0xff, 0xff, 0x7f, 0x80, 0x00, 0x00, 0x7f, 0xff, 0xff],
           4: [0x00, 0x00, 0x00, 0x80, 0xff, 0xff, 0xff, 0x7f, 0x80, 0x00, 0x00, 0x00, 0x7f, 0xff, 0xff, 0xff,
               0x01, 0x00, 0x00, 0x01, 0xff, 0xff, 0xff, 0xfe]}[bps]
    k = min(len(ext), b.size) // bps * bps
    b[:k] = ext[:k]
    return b.tobytes()


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("fmt", FORMATS)
def test_pcm_to_f32_bit_exact(ssb, oracle, cuda, fmt):
    a = ssb.Analyzer()
    for n in (0, 1, 3, 4, 5, 1023, 4096 + 3, (1 << 20) + 1):
        raw = _raw(fmt, n, seed=n + 1)
        got = ssb.pcm_to_f32(a, raw, fmt)
        want = oracle.capture_ref.pcm_to_f32(raw, fmt)
        assert got.size == want.size == n
        assert np.array_equal(_bits(got), _bits(want)), f"{fmt} n={n}"


@pytest.mark.parametrize("fmt", FORMATS)
def test_pcm_device_unaligned_and_vector_paths_agree(ssb, oracle, cuda, fmt):
    """the device entry point with aligned pointers (vector loads) and with a 1-byte / 1-float offset (byte loads)"""
    torch = cuda
    n = 50001
    raw = _raw(fmt, n, seed=7)
    want = oracle.capture_ref.pcm_to_f32(raw, fmt)
    b = ssb.BatchAnalyzer(1, 2, 48000)
    buf = torch.zeros(len(raw) + 64, dtype=torch.uint8, device="cuda")
    for off in (0, 1, 16):
        buf[off:off + len(raw)] = torch.frombuffer(bytearray(raw), dtype=torch.uint8).cuda()
        out = torch.full((n + 8,), 7.0, dtype=torch.float32, device="cuda")
        for ooff in (0, 1):
            got = b.pcm_to_f32_device(buf[off:off + len(raw)], fmt, out=out[ooff:ooff + n])
            torch.cuda.synchronize()
            assert np.array_equal(_bits(got.cpu().numpy()), _bits(want)), f"{fmt} in+{off} out+{ooff}"
            assert float(out[ooff + n]) == 7.0  # nothing written past the end


@pytest.mark.parametrize("fmt,channels,rate", [("s16le", 2, 48000), ("s24le", 2, 44100), ("s32be", 1, 48000), ("u8", 6, 96000)])
def test_add_pcm_equals_add_samples_of_converted(ssb, oracle, cuda, fmt, channels, rate):
    """decode_file -> add_samples on the reference == raw PCM straight into the meter here"""
    rng = np.random.default_rng(5)
    frames = rate * 4 + 123
    t = np.arange(frames) / rate
    x = 0.4 * np.sin(2 * np.pi * 440 * t)[:, None] * np.linspace(1.0, 0.3, channels)[None, :] + 0.02 * rng.standard_normal((frames, channels))
    bps = oracle.capture_ref.pcm_bytes_per_sample(fmt)
    if fmt == "u8":
        raw = np.clip(np.round(x * 127 + 128), 0, 255).astype(np.uint8).tobytes()
    elif bps == 2:
        raw = np.clip(np.round(x * 32767), -32768, 32767).astype("<i2").tobytes()
    elif bps == 3:
        v = np.clip(np.round(x * 8388607), -8388608, 8388607).astype(np.int32).ravel()
        raw = np.stack([v & 0xff, (v >> 8) & 0xff, (v >> 16) & 0xff], axis=1).astype(np.uint8).tobytes()
    else:
        raw = np.clip(np.round(x * 2147483647), -2147483648, 2147483647).astype(">i4").tobytes()
    f = oracle.capture_ref.pcm_to_f32(raw, fmt)
    a, o = ssb.Analyzer(), oracle.Analyzer()
    a.create_loudness_meter(channels, rate)
    o.create_loudness_meter(channels, rate)
    chunk = rate * channels  # 1 s per call
    for i in range(0, f.size, chunk):
        a.add_pcm(raw[i * bps:(i + chunk) * bps], fmt)
        o.add_samples(f[i:i + chunk])
    assert abs(a.get_shortterm_lufs() - o.get_shortterm_lufs()) <= 1e-9
    assert abs(a.get_integrated_lufs() - o.get_integrated_lufs()) <= 1e-4
    # and the same bits as feeding the converted f32 through the f32 entry point
    a2 = ssb.Analyzer()
    a2.create_loudness_meter(channels, rate)
    for i in range(0, f.size, chunk):
        a2.add_samples(f[i:i + chunk])
    assert a.get_integrated_lufs() == a2.get_integrated_lufs() and a.get_shortterm_lufs() == a2.get_shortterm_lufs()
    if channels > 1:
        with pytest.raises(ssb.SsbError) as e:          # ragged: not a whole number of frames -> Error::NoMem
            a.add_pcm(raw[: bps * (channels + 1)], fmt)
        assert e.value.code == 1
    with pytest.raises(ssb.SsbError) as e:
        a.add_pcm(raw[: bps * channels], 99)             # unknown format code
    assert e.value.code == 10


def test_pcm_batch_device_matches_f32_batch(ssb, oracle, cuda):
    """many streams: raw s16 on the device -> meter == converted f32 -> meter (identical results rows)"""
    torch = cuda
    from tests.signals import stream_batch
    n, frames = 96, 19200
    x = stream_batch(n, frames, 2, seed=11)
    raw = np.clip(np.round(x * 32767), -32768, 32767).astype("<i2")
    f = oracle.capture_ref.pcm_to_f32(raw.tobytes(), "s16le").reshape(n, frames, 2)
    b1, b2 = ssb.BatchAnalyzer(n, 2, 48000), ssb.BatchAnalyzer(n, 2, 48000)
    b1.add_frames_pcm_device(torch.from_numpy(raw.view(np.uint8).reshape(-1)).cuda(), "s16le")
    b2.add_frames_device(torch.from_numpy(f).cuda())
    r1, r2 = b1.results_device().cpu().numpy(), b2.results_device().cpu().numpy()
    assert np.array_equal(r1, r2)
    b3 = ssb.BatchAnalyzer(n, 2, 48000)
    b3.add_frames_pcm_host(raw.view(np.uint8).reshape(-1), "s16le")
    assert np.array_equal(b3.results_device().cpu().numpy(), r2)


def test_capture_ring_matches_reference_semantics(ssb, oracle, cuda):
    cap = 1000
    r, o = ssb.CaptureRing(cap), oracle.capture_ref.RingRef(cap)
    assert r.capacity == cap and r.written == 0
    assert np.array_equal(r.to_vec(), o.to_vec())  # zero-filled
    rng = np.random.default_rng(3)
    total = 0
    for n, mono in [(10, False), (7, True), (1, True), (0, True), (0, False), (333, False), (400, True), (999, False),
                    (1000, False), (1001, False), (2500, False), (501, True), (700, True), (3, False), (1, False)]:
        d = rng.standard_normal(n).astype(np.float32)
        r.push(d, mono)
        o.callback(d, mono)
        total += (2 * n - 1 if n else 0) if mono else n
        assert r.written == total
        assert np.array_equal(_bits(r.to_vec()), _bits(o.to_vec())), (n, mono)


def _tick_inputs(rate, seconds, seed):
    rng = np.random.default_rng(seed)
    n = int(rate * seconds)
    t = np.arange(n) / rate
    l = 0.5 * np.sin(2 * np.pi * 700 * t) + 0.05 * rng.standard_normal(n)
    r = 0.3 * np.sin(2 * np.pi * 700 * t + 0.4) + 0.05 * rng.standard_normal(n)
    x = np.empty(2 * n, dtype=np.float32)
    x[0::2], x[1::2] = l, r
    return x


def _check_tick(got, want, st_tol=1e-9):
    mid, side, wave, st, fs, ls = got
    wm, ws, ww, wst, werr = want
    assert (mid is None) == (wm is None) and (side is None) == (ws is None)
    if wm is not None:
        assert fs == 0
        assert np.array_equal(mid[:, 0], wm[:, 0]) and np.array_equal(side[:, 0], ws[:, 0])
        assert_db_close(mid[:, 1], wm[:, 1])
        assert_db_close(side[:, 1], ws[:, 1])
    assert np.array_equal(wave, ww)
    assert (ls != 0) == (werr is not None)
    if np.isfinite(wst):
        assert abs(st - wst) <= st_tol
    else:
        assert st == wst


def test_mic_tick_stereo_device(ssb, oracle, cuda):
    """analyze_microphone_input on a stereo 48 kHz device: ticks interleaved with capture callbacks of odd sizes"""
    rate = 48000
    x = _tick_inputs(rate, 20.0, 1)
    a, o = ssb.Analyzer(), oracle.Analyzer()
    a.create_loudness_meter(2, rate)     # tui.rs:1800-1802 (device selected)
    o.create_loudness_meter(2, rate)
    r, oref = ssb.CaptureRing(30 * rate), oracle.capture_ref.RingRef(30 * rate)
    _check_tick(a.analyze_microphone_input(r), oracle.capture_ref.mic_tick(oref.to_vec(), o))   # silence: -inf
    pos = 0
    for step, n in enumerate([2 * 480, 2 * 4800, 2 * 333 + 1, 2 * 24000 + 1, 2 * 96000, 2 * 700000, 2 * 8192]):
        n = min(n, x.size - pos)
        r.push(x[pos:pos + n])
        oref.callback(x[pos:pos + n], False)
        pos += n
        _check_tick(a.analyze_microphone_input(r), oracle.capture_ref.mic_tick(oref.to_vec(), o))
    # a second tick with nothing new pushed re-feeds the same 16384 values, as the reference does
    _check_tick(a.analyze_microphone_input(r), oracle.capture_ref.mic_tick(oref.to_vec(), o))


def test_mic_tick_mono_device_quirk(ssb, oracle, cuda):
    """mono device at 44.1 kHz: mono meter, up-mixed ring with the left/right parity flipping every callback"""
    rate = 44100
    rng = np.random.default_rng(2)
    a, o = ssb.Analyzer(), oracle.Analyzer()
    a.create_loudness_meter(1, rate)
    o.create_loudness_meter(1, rate)
    r, oref = ssb.CaptureRing(30 * rate), oracle.capture_ref.RingRef(30 * rate)
    t0 = 0
    for n in (441, 4410, 1, 44100, 300000, 512):
        t = (np.arange(n) + t0) / rate
        d = (0.4 * np.sin(2 * np.pi * 300 * t) + 0.01 * rng.standard_normal(n)).astype(np.float32)
        t0 += n
        r.push(d, True)
        oref.callback(d, True)
        _check_tick(a.analyze_microphone_input(r), oracle.capture_ref.mic_tick(oref.to_vec(), o))


def test_mic_tick_errors(ssb, oracle, cuda):
    a = ssb.Analyzer()                       # default meter: 2 ch, 44100
    small = ssb.CaptureRing(30 * 22050)      # shorter than 30 * rate: the reference's slice would panic
    with pytest.raises(ssb.SsbError) as e:
        a.analyze_microphone_input(small)
    assert e.value.code == 10
    a.create_loudness_meter(2, 22050)        # 20 kHz above Nyquist: get_fft fails, waveform and meter still run
    o = oracle.Analyzer()
    o.create_loudness_meter(2, 22050)
    x = _tick_inputs(22050, 3.0, 4)
    oref = oracle.capture_ref.RingRef(30 * 22050)
    small.push(x)
    oref.callback(x, False)
    got = a.analyze_microphone_input(small)
    assert got[0] is None and got[4] == 8
    _check_tick(got, oracle.capture_ref.mic_tick(oref.to_vec(), o))
    a.create_loudness_meter(3, 22050)        # 16384 values are not a whole number of 3-channel frames -> NoMem
    got = a.analyze_microphone_input(small)
    assert got[5] == 1


def test_golden_capture_vectors(ssb, cuda):
    """the committed vectors (tests/golden/golden_capture_v1.npz): PCM bits, ring contents, one microphone tick"""
    import os
    from tests.golden import make_golden_capture as M
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_capture_v1.npz"))
    a = ssb.Analyzer()
    for fmt in M.FORMATS:
        got = ssb.pcm_to_f32(a, g[f"pcm_{fmt}_raw"].tobytes(), fmt)
        assert np.array_equal(_bits(got), g[f"pcm_{fmt}_f32bits"]), fmt
    r = ssb.CaptureRing(M.RING_CAP)
    for d, mono in M.ring_pushes():
        r.push(d, bool(mono))
    assert np.array_equal(_bits(r.to_vec()), _bits(g["ring_to_vec"]))
    rate = 44100
    ring = ssb.CaptureRing(30 * rate)
    x = M.mic_signal(rate)
    ring.push(x[: x.size // 2 + 1])
    ring.push(x[x.size // 2 + 1:])
    mid, side, wave, st, fs, ls = a.analyze_microphone_input(ring)     # default meter: 2 ch, 44100
    assert fs == 0 and ls == 0
    assert_db_close(mid[:, 1], g["mic_mid_db"])
    assert_db_close(side[:, 1], g["mic_side_db"])
    assert np.array_equal(wave[:, 1].astype(np.float32), g["mic_wave"])
    assert abs(st - g["mic_shortterm"][0]) <= 1e-9


def test_capture_ring_concurrent_producer_snapshots_are_consistent(ssb, cuda):
    """single producer (the capture callback's thread) pushing while the consumer snapshots: every snapshot is a
    contiguous window of the pushed sequence (ctypes releases the GIL, so the two really overlap)"""
    import threading
    cap = 1 << 14
    r = ssb.CaptureRing(cap)
    stop = threading.Event()

    def producer():
        k = 1
        while not stop.is_set() and k < (1 << 24) - 200:
            n = 1 + (k % 97)
            r.push(np.arange(k, k + n, dtype=np.float32))    # consecutive integers, exact in f32
            k += n

    t = threading.Thread(target=producer)
    t.start()
    good = 0
    try:
        for _ in range(400):
            try:
                v = r.to_vec()
            except ssb.SsbError:
                continue          # lapped four times in a row: refused, never a wrong answer
            nz = np.flatnonzero(v)
            if nz.size:
                assert nz[-1] == cap - 1 and np.all(np.diff(nz) == 1)          # zeros (initial fill) come first
                assert np.all(np.diff(v[nz[0]:]) == 1.0), "torn snapshot"
            good += 1
    finally:
        stop.set()
        t.join()
    assert good > 0


@pytest.mark.parametrize("sr", [44100, 48000, 96000])
def test_reference_mic_tests_on_gpu(ssb, oracle, cuda, sr):
    """The reference's own microphone-tick tests (tui.rs:2271-2368) through the C ABI: a 44100 * 30 ring (tui.rs:2200),
    the default analyzer, sr * 30 enqueued sine samples; the reference's assertions, then parity with the oracle."""
    from tests.signals import ref_mic_test_assertions, ref_mic_test_ring_fill
    s = ref_mic_test_ring_fill(sr)
    ring, oref = ssb.CaptureRing(44100 * 30), oracle.capture_ref.RingRef(44100 * 30)
    ring.push(s)
    oref.callback(s, False)
    a, o = ssb.Analyzer(), oracle.Analyzer()
    got = a.analyze_microphone_input(ring)
    ref_mic_test_assertions(sr, got[0])
    _check_tick(got, oracle.capture_ref.mic_tick(oref.to_vec(), o))
