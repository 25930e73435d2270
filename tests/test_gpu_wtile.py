"""GPU parity of the second-generation batch kernel (csrc/loudness_wtile.cu: warp-private TMA pipelines, mixed T4/T5
warps, fused gating + results epilogue) against the thread-per-channel kernel and the CPU oracle, through the C ABI.

Reference path: Analyzer::add_samples + get_shortterm_lufs / get_integrated_lufs / get_loudness_range / get_true_peak
(src/analyzer.rs:139-164).  Tolerances as in test_gpu_loudness.py: 1e-4 LU stated by BASELINE.json, asserted tighter:
identical histograms, 2e-9 LU on momentary / short-term (time-segmented recursion), bit-exact sample peak, 2e-6 true peak.
"""
import numpy as np
import pytest

from tests.signals import stream_batch

pytestmark = pytest.mark.gpu

LU_TOL = 1e-4
TP_RTOL = 2e-6
SEG_LU_TOL = 2e-9


def close_lu(a, b, tol=LU_TOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    both_inf = np.isneginf(a) & np.isneginf(b)
    with np.errstate(invalid="ignore"):
        return np.all(both_inf | (np.abs(a - b) <= tol))


@pytest.mark.parametrize("variant", [5, 6])
@pytest.mark.parametrize("n,channels,frames,mode_name,rate", [
    (301, 2, 19200, "MODE_LOUDNESS", 48000),     # 2-3 streams per CTA: type-A warps only
    (4096, 2, 9600, "MODE_LOUDNESS", 48000),     # cfg2's stream count: 27-28 streams per CTA, every warp live
    (4096, 2, 9600, "MODE_ALL", 48000),
    (9000, 2, 4800, "MODE_LOUDNESS", 48000),     # 60-61 streams per CTA: three passes per warp through one TMA ring
    (1000, 1, 19200, "MODE_ALL", 48000),         # mono: 8 / 6 streams per warp
    (5000, 1, 4800 + 640, "MODE_LOUDNESS", 48000),
    (2000, 2, 8820, "MODE_ALL", 44100),          # bucket boundaries inside tiles, ragged tail through the generic kernel
    (2100, 2, 19200, "MODE_ALL", 96000),         # 2x interpolator
    (37, 2, 19200 + 77, "MODE_ALL", 48000),      # fewer streams than SMs
])
def test_wtile_matches_generic_and_oracle(ssb, oracle, cuda, variant, n, channels, frames, mode_name, rate):
    torch = cuda
    mode = getattr(ssb, mode_name) | ssb.MODE_SAMPLE_PEAK
    reps = 3
    x = stream_batch(n, frames * reps, channels, seed=frames + channels + n, rate=rate)
    xd = torch.from_numpy(x).cuda()
    fast = ssb.BatchAnalyzer(n, channels, rate, mode)
    fast.force_kernel(variant)
    slow = ssb.BatchAnalyzer(n, channels, rate, mode)
    slow.force_generic(True)
    sub = np.unique(np.concatenate([np.arange(0, n, max(1, n // 97)), [n - 1]]))
    ob = oracle.Batch(len(sub), channels, rate, getattr(oracle, mode_name) | oracle.MODE_SAMPLE_PEAK)
    for k in range(reps):
        sl = xd[:, k * frames:(k + 1) * frames, :].contiguous()
        fast.add_frames_device(sl)
        slow.add_frames_device(sl)
        ob.add_frames(np.ascontiguousarray(x[sub, k * frames:(k + 1) * frames, :]))
    want = ob.query()
    assert close_lu(fast.loudness_global()[sub], want["global"])
    assert close_lu(fast.loudness_range()[sub], want["range"])
    assert close_lu(fast.loudness_global(), slow.loudness_global(), 1e-9)
    assert np.array_equal(fast.sample_peak(), np.abs(x).max(axis=1).astype(np.float64))
    if mode_name == "MODE_ALL":
        assert np.all(np.abs(fast.true_peak()[sub] - want["true_peak"]) <= TP_RTOL * want["true_peak"])
        assert np.array_equal(fast.true_peak(), slow.true_peak())   # same taps, same f32 FMA order in every kernel
    if (frames * reps) % ((rate + 5) // 10) == 0:
        d = np.abs(fast.loudness_momentary()[sub] - want["momentary"])
        print("wtile max |dLUFS| vs oracle:", d[np.isfinite(d)].max() if np.isfinite(d).any() else 0.0)
        assert close_lu(fast.loudness_momentary()[sub], want["momentary"], SEG_LU_TOL)
        assert close_lu(fast.loudness_shortterm()[sub], want["shortterm"], SEG_LU_TOL)
        assert close_lu(fast.loudness_momentary(), slow.loudness_momentary(), SEG_LU_TOL)
    for i, s in enumerate(sub[:: max(1, len(sub) // 7)]):
        hb, hs = fast.histograms(int(s))
        gb, gs = slow.histograms(int(s))
        assert np.array_equal(hb, gb) and np.array_equal(hs, gs)
        j = int(np.searchsorted(sub, s))
        assert np.array_equal(hb, ob._per_stream_hist(j)[0])


@pytest.mark.parametrize("n,channels,mode_name", [(4096, 2, "MODE_LOUDNESS"), (777, 2, "MODE_ALL"), (1500, 1, "MODE_ALL")])
def test_fused_results_equal_separate_query(ssb, cuda, n, channels, mode_name):
    """ssb_add_frames_f32_device_results (gating + result rows in the filter kernel's epilogue) writes the rows
    ssb_add_frames_f32_device + ssb_results_device write, launch after launch, with one kernel launch per call."""
    torch = cuda
    rate, frames = 48000, 19200
    mode = getattr(ssb, mode_name)
    a = ssb.BatchAnalyzer(n, channels, rate, mode)
    b = ssb.BatchAnalyzer(n, channels, rate, mode)
    out = torch.empty((n, a.stride), dtype=torch.float64, device="cuda")
    for k in range(12):     # 12 x 400 ms: past the first 3 s window, several LRA entries
        xd = torch.from_numpy(stream_batch(n, frames, channels, seed=40 + k, t0=k * frames)).cuda()
        l0 = a.launches
        a.add_frames_results_device(xd, out)
        assert a.launches - l0 == 1
        b.add_frames_device(xd)
        want = b.results_device()
        assert torch.equal(out.cpu().nan_to_num(nan=-777.0), want.cpu().nan_to_num(nan=-777.0))
    for s in (0, n // 2, n - 1):
        ha, hb = a.histograms(s), b.histograms(s)
        assert np.array_equal(ha[0], hb[0]) and np.array_equal(ha[1], hb[1])
    # a ragged feed falls back to filter + k_results and still returns the right rows
    xd = torch.from_numpy(stream_batch(n, 5000, channels, seed=99)).cuda()
    a.add_frames_results_device(xd, out)
    b.add_frames_device(xd)
    assert torch.equal(out.cpu().nan_to_num(nan=-777.0), b.results_device().cpu().nan_to_num(nan=-777.0))


def test_host_feed_uses_fused_rows(ssb, oracle, cuda):
    """The host-facing pair add_frames_f32 + loudness_* (the e2e path of bench.py) reads the rows the fused epilogue
    left in the handle: same values as the oracle, one launch per feed."""
    n, ch, rate, frames = 512, 2, 48000, 19200
    a = ssb.BatchAnalyzer(n, ch, rate, ssb.MODE_ALL)
    ob = oracle.Batch(n, ch, rate, oracle.MODE_ALL)
    for k in range(9):
        x = stream_batch(n, frames, ch, seed=7 + k, t0=k * frames)
        l0 = a.launches
        a.add_frames_host(x)
        got_i, got_s = a.loudness_global(), a.loudness_shortterm()
        assert a.launches - l0 == 1
        ob.add_frames(x)
    want = ob.query()
    assert close_lu(got_i, want["global"]) and close_lu(got_s, want["shortterm"], SEG_LU_TOL)
    assert np.all(np.abs(a.true_peak() - want["true_peak"]) <= TP_RTOL * want["true_peak"])
