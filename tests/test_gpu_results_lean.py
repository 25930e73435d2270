"""GPU parity of the lean results path (csrc/loudness_results.cuh: results_lean — per-stream running sums of the
histogram gating, 16 lanes per stream) against the full histogram scan of the same library (SSB_RESULTS_LEAN=0 at
create) and against the CPU oracle, through the C ABI.

Reference path: Analyzer::add_samples followed by get_shortterm_lufs / get_integrated_lufs / get_loudness_range /
get_true_peak after every tick (src/analyzer.rs:139-164).  What the lean path must reproduce is ebur128's
gated_loudness over the block histogram: the signals below sweep each stream's level over 30-50 dB, with exact silence
in between, so the relative gate's start bin moves up and down across many bins between queries and the "bins between
the old and the new gate change sides" bookkeeping is exercised in both directions.

Tolerances: momentary / short-term bit-identical to the full scan (same additions in the same order); histograms
identical; integrated within 1e-9 LU of the full scan (a different summation order of the same bin energies) and within
BASELINE.json's 1e-4 LU of the oracle; loudness range identical (it is the same scan); peaks identical.
"""
import os

import numpy as np
import pytest

from tests.signals import stream_batch

pytestmark = pytest.mark.gpu


def swept_batch(torch, n, frames, channels, seed, rate, t0):
    """stream_batch (128 distinct streams, repeated) shaped on the GPU by a per-stream level sweep: -45 .. -5 dB over
    5-9 s periods, exact zeros for part of every cycle.  Returns a cuda f32 tensor [n, frames, C]; every consumer
    (lean, full scan, oracle) is fed these same values."""
    base = torch.from_numpy(stream_batch(min(n, 128), frames, channels, seed=seed, rate=rate, t0=t0)).cuda()
    s = torch.arange(n, device="cuda", dtype=torch.float64)
    t = (torch.arange(frames, device="cuda", dtype=torch.float64) + t0) / rate
    period = 5.0 + (s % 5)
    ph = 2 * np.pi * (t[None, :] / period[:, None] + 0.37 * s[:, None])
    gain = 10.0 ** ((-25.0 + 20.0 * torch.sin(ph)) / 20.0) * 4.0
    gain = torch.where(torch.cos(0.61 * ph + 0.2 * s[:, None]) > 0.93, torch.zeros_like(gain), gain)   # silence
    x = base[torch.arange(n, device="cuda") % base.shape[0]].to(torch.float64) * gain[:, :, None]
    return x.to(torch.float32).contiguous()


def make(ssb, n, ch, rate, mode, lean):
    old = os.environ.get("SSB_RESULTS_LEAN")
    os.environ["SSB_RESULTS_LEAN"] = "1" if lean else "0"
    try:
        return ssb.BatchAnalyzer(n, ch, rate, mode)
    finally:
        if old is None:
            del os.environ["SSB_RESULTS_LEAN"]
        else:
            os.environ["SSB_RESULTS_LEAN"] = old


def rows_close(got, want, ch):
    g, w = got.cpu().numpy(), want.cpu().numpy()
    # momentary, short-term, LRA, peaks: identical (NaN == NaN, -inf == -inf)
    for col in [0, 1, 3] + list(range(4, 4 + 2 * ch)):
        assert np.array_equal(g[:, col], w[:, col], equal_nan=True), f"column {col}"
    gi, wi = g[:, 2], w[:, 2]
    same_inf = np.isneginf(gi) & np.isneginf(wi)
    with np.errstate(invalid="ignore"):
        ok = same_inf | (np.abs(gi - wi) <= 1e-9)
    assert np.all(ok), f"integrated: max diff {np.nanmax(np.abs(np.where(same_inf, 0, gi - wi)))}"


@pytest.mark.parametrize("n,ch,rate,frames,mode_name,fused", [
    (4200, 2, 48000, 19200, "MODE_LOUDNESS", True),   # cfg2's cadence: 4 pending buckets per query, fused epilogue, every pair live
    (9000, 2, 48000, 9600, "MODE_ALL", True),         # three passes per pair
    (1000, 1, 48000, 4800, "MODE_ALL", True),         # one pending bucket per query, mono (8 / 6 streams per set)
    (300, 2, 48000, 48000, "MODE_LOUDNESS", True),    # 10 pending buckets: every lean launch also gates a 3 s entry
    (300, 2, 48000, 52800, "MODE_LOUDNESS", False),   # 11 pending: over the lean limit, full scans rebuild the cache
    (257, 2, 44100, 4410 * 3 + 100, "MODE_ALL", False),   # feed positions off the 100 ms grid: M / S are NaN
    (130, 6, 96000, 19200, "MODE_ALL", False),        # 5.1: k_loudness_rows_any + k_results_lean
    (50, 2, 48000, 19200, "MODE_LOUDNESS", True),     # fewer streams than SMs
])
def test_lean_results_match_full_scan_and_oracle(ssb, oracle, cuda, n, ch, rate, frames, mode_name, fused):
    torch = cuda
    mode = getattr(ssb, mode_name)
    lean = make(ssb, n, ch, rate, mode, True)
    full = make(ssb, n, ch, rate, mode, False)
    sub = np.unique(np.concatenate([np.arange(0, n, max(1, n // 23)), [n - 1]]))
    ob = oracle.Batch(len(sub), ch, rate, getattr(oracle, mode_name))
    out = torch.empty((n, lean.stride), dtype=torch.float64, device="cuda")
    n_calls = max(12, int(np.ceil(14.0 * rate / frames)))   # >= 14 s: several level cycles, several LRA entries
    n_calls = min(n_calls, 60)
    for k in range(n_calls):
        xd = swept_batch(torch, n, frames, ch, seed=11 + k, rate=rate, t0=k * frames)
        if fused:
            lean.add_frames_results_device(xd, out)
            got = out
        else:
            lean.add_frames_device(xd)
            got = lean.results_device()
        full.add_frames_device(xd)
        want = full.results_device()
        rows_close(got, want, ch)
        ob.add_frames(np.ascontiguousarray(xd[torch.from_numpy(sub).cuda()].cpu().numpy()))
        if k == 3:
            # a query with nothing pending in between (the cache must not be applied twice)
            rows_close(lean.results_device(), full.results_device(), ch)
    ref = ob.query()
    gi = lean.loudness_global()[sub]
    both_inf = np.isneginf(gi) & np.isneginf(ref["global"])
    with np.errstate(invalid="ignore"):
        assert np.all(both_inf | (np.abs(gi - ref["global"]) <= 1e-4))
        lr = lean.loudness_range()[sub]
        assert np.all(np.abs(lr - ref["range"]) <= 1e-4)
    for s in (0, n // 3, n - 1):
        ha, hb = lean.histograms(s), full.histograms(s)
        assert np.array_equal(ha[0], hb[0]) and np.array_equal(ha[1], hb[1])
        assert ha[0].sum() > 0


def test_lean_survives_reset_and_long_unqueried_feed(ssb, cuda):
    """reset() zeroes the cache (valid for an empty meter); a feed long enough for the lazy k_gating flush invalidates
    it, the next query is a full scan that rebuilds it, and the queries after that are lean again."""
    torch = cuda
    n, ch, rate, frames = 400, 2, 48000, 19200
    lean = make(ssb, n, ch, rate, ssb.MODE_ALL, True)
    full = make(ssb, n, ch, rate, ssb.MODE_ALL, False)
    out = torch.empty((n, lean.stride), dtype=torch.float64, device="cuda")
    t0 = 0
    for phase in range(2):
        for k in range(5):
            x = swept_batch(torch, n, frames, ch, 5 + k, rate, t0)
            t0 += frames
            lean.add_frames_results_device(x, out)
            full.add_frames_device(x)
            rows_close(out, full.results_device(), ch)
        # 12 s without a query: more than the 34 pending buckets the lazy gating allows -> k_gating runs
        for k in range(3):
            x = swept_batch(torch, n, 4 * 48000, ch, 50 + k, rate, t0)
            t0 += 4 * 48000
            lean.add_frames_device(x)
            full.add_frames_device(x)
        rows_close(lean.results_device(), full.results_device(), ch)
        for k in range(4):
            x = swept_batch(torch, n, frames, ch, 70 + k, rate, t0)
            t0 += frames
            lean.add_frames_results_device(x, out)
            full.add_frames_device(x)
            rows_close(out, full.results_device(), ch)
        if phase == 0:
            lean.reset()
            full.reset()
            t0 = 0
