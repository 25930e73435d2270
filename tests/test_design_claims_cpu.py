"""CPU checks of two numerical claims DESIGN.md builds kernels on (no GPU, no product code: scipy's f64 filter and
the oracle's K-weighting coefficients).

1. File mode (loudness_scan.cu): a chunk of a file that starts 0.4 s early from ZERO filter state reproduces, after the
   run-in, the serial recursion over the whole file down to that recursion's own rounding noise.  The inherited state
   has decayed by e^-95 (nothing); what remains is that two f64 runs of a direct-form filter with poles at 0.995 never
   re-synchronise their roundings: ~2e-11 of full scale at 48 kHz (3e-9 at 192 kHz), the same floor any re-association of the recursion has (the
   time-segmented batch kernel's hand-off sits there too), i.e. < 1e-9 LU on a block energy.
2. Power-of-two input scaling commutes exactly with the recursion and the energy sums (what lets raw integer PCM run
   through the filter unscaled).
"""
import numpy as np
import pytest
from scipy import signal


def _signal(rate, seconds, seed):
    rng = np.random.default_rng(seed)
    n = int(rate * seconds)
    t = np.arange(n) / rate
    x = 0.4 * np.sin(2 * np.pi * 55.0 * t) + 0.1 * rng.standard_normal(n)
    x[n // 3: n // 3 + rate // 2] += 0.5          # half a second of DC: the worst case for the 38 Hz high-pass tail
    return x.astype(np.float32).astype(np.float64)


@pytest.mark.parametrize("rate", [8000, 22050, 48000, 96000, 192000])
def test_zero_state_run_in_reproduces_the_serial_recursion(oracle, rate):
    b, a = oracle.EbuR128(1, rate).coeffs()
    x = _signal(rate, 6.0, rate)
    y = signal.lfilter(b, a, x)
    s100 = (rate + 5) // 10
    warm, chunk = 4 * s100, 10 * s100
    for c0 in range(chunk, x.size - chunk, chunk):
        yc = signal.lfilter(b, a, x[c0 - warm: c0 + chunk])[warm:]
        ys = y[c0: c0 + chunk]
        # the floor grows with the square of the sample rate (poles move towards z = 1): 1e-9 of full scale at 48 kHz is
        # two orders above what is measured there (2e-11)
        tol = 1e-9 * max(1.0, (rate / 48000.0) ** 2)
        assert np.max(np.abs(yc - ys)) <= tol * np.max(np.abs(y)), f"chunk at {c0}: max |d| = {np.max(np.abs(yc - ys))}"
        # per 100 ms bucket energy: what the gating sees (1e-9 relative = 4e-9 LU)
        ec, es = np.add.reduceat(yc * yc, np.arange(0, chunk, s100)), np.add.reduceat(ys * ys, np.arange(0, chunk, s100))
        assert np.max(np.abs(ec - es) / es) <= tol
    # without the run-in the chunk is simply wrong: the high-pass has not settled
    y0 = signal.lfilter(b, a, x[chunk: 2 * chunk])
    assert np.max(np.abs(y0 - y[chunk: 2 * chunk])) > 1e-3
    # and the bound behind it: the largest pole radius to the power of the run-in length
    r = np.max(np.abs(np.roots(a)))
    assert r ** warm < 1e-38


def test_power_of_two_scaling_commutes_with_the_filter_and_the_energy_sums(oracle):
    b, a = oracle.EbuR128(1, 48000).coeffs()
    rng = np.random.default_rng(1)
    q = rng.integers(-32768, 32768, 48000).astype(np.float64)            # raw s16 samples
    y_int = signal.lfilter(b, a, q)                                      # recursion on integer-valued input
    y_flt = signal.lfilter(b, a, q / 32768.0)                            # symphonia's f32 conversion first (exact in f64 too)
    assert np.array_equal(y_int / 32768.0, y_flt)
    e_int = np.add.reduceat(y_int * y_int, np.arange(0, q.size, 4800))
    e_flt = np.add.reduceat(y_flt * y_flt, np.arange(0, q.size, 4800))
    assert np.array_equal(e_int / 2.0 ** 30, e_flt)
