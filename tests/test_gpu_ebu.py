"""Standards-based known answers straight through the CUDA path (not via the oracle): EBU Tech 3341 loudness
and true-peak cases, EBU Tech 3342 loudness-range cases, synthesised (the reference ships no audio)."""
import numpy as np
import pytest

from tests.test_oracle_kat import tone_stereo

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("level", [-23.0, -33.0])
def test_ebu3341_case1_2(ssb, cuda, level):
    a = ssb.Analyzer()
    a.create_loudness_meter(2, 48000)
    a.add_samples(tone_stereo([level], [20.0]))
    for v in (a.get_momentary_lufs(), a.get_shortterm_lufs(), a.get_integrated_lufs()):
        assert abs(v - level) <= 0.1


def test_ebu3341_gating_cases(ssb, cuda):
    for lv, du in (([-36, -23, -36], [10, 60, 10]), ([-72, -36, -23, -36, -72], [10, 10, 60, 10, 10]),
                   ([-26, -20, -26], [20, 20.1, 20])):
        a = ssb.Analyzer()
        a.create_loudness_meter(2, 48000)
        x = tone_stereo(lv, du)
        for off in range(0, x.size, 96000):          # 1 s chunks, like calculate_integrated_lufs
            a.add_samples(x[off:off + 96000])
        assert abs(a.get_integrated_lufs() - (-23.0)) <= 0.1, (lv, a.get_integrated_lufs())
        assert abs(a.calculate_integrated_lufs(2, x) - (-23.0)) <= 0.1


def test_ebu3341_case6_surround_batch(ssb, cuda):
    """5.1 (LFE silent) through the multichannel kernel, as a batch of identical streams."""
    torch = cuda
    rate, n = 48000, 48000 * 20
    t = np.arange(n) / rate
    s = np.sin(2 * np.pi * 1000 * t)
    x = np.zeros((n, 6), dtype=np.float32)
    for c, lv in ((0, -28), (1, -28), (2, -24), (4, -30), (5, -30)):
        x[:, c] = 10 ** (lv / 20) * s
    b = ssb.BatchAnalyzer(7, 6, rate, ssb.MODE_ALL)
    b.add_frames_device(torch.from_numpy(np.broadcast_to(x, (7, n, 6)).copy()).cuda())
    assert np.all(np.abs(b.loudness_global() - (-23.0)) <= 0.1)


def test_ebu3342_lra(ssb, cuda):
    for lv, du, want in (([-20, -30], [20, 20], 10.0), ([-20, -15], [20, 20], 5.0), ([-40, -20], [20, 20], 20.0),
                         ([-50, -35, -20, -35, -50], [20, 20, 20, 20, 20], 15.0)):
        a = ssb.Analyzer()
        a.create_loudness_meter(2, 48000)
        a.add_samples(tone_stereo(lv, du))
        assert abs(a.get_loudness_range() - want) <= 1.0, (lv, a.get_loudness_range())


def test_ebu3341_true_peak(ssb, cuda):
    rate = 48000
    i = np.arange(rate * 2)
    for phase_deg in (0.0, 45.0):
        x = 0.5 * np.sin(2 * np.pi * (rate / 4) * i / rate + np.deg2rad(phase_deg))
        a = ssb.Analyzer()
        a.create_loudness_meter(2, rate)
        a.add_samples(np.repeat(x.astype(np.float32)[:, None], 2, 1).ravel())
        tp_db = 20 * np.log10(a.get_true_peak()[0])
        assert -6.4 <= tp_db <= -5.8, (phase_deg, tp_db)


def test_histogram_bin_edges_g5(ssb, oracle, cuda):
    """SURVEY section 8c, golden set G5: block energies placed on, and 1-2 ulp either side of, every one of the 1001
    histogram boundaries (and on the bin centres, and log-uniform random ones) land in the bin the crate's bisection
    picks.  The gating kernels reach the index from a closed-form guess corrected against the boundary table
    (csrc/loudness_results.cuh find_histogram_index); one wrong edge decision moves a stream's integrated loudness by
    ~0.1 LU / n_blocks, so this is asserted exactly."""
    b = ssb.BatchAnalyzer(1, 2, 48000, ssb.MODE_ALL)
    bounds = np.array([oracle.histogram_boundary(i) for i in range(1001)])
    centres = np.array([oracle.histogram_energy(i) for i in range(1000)])
    cases = [bounds, centres]
    for k in (1, 2, 3):
        up, dn = bounds.copy(), bounds.copy()
        for _ in range(k):
            up, dn = np.nextafter(up, np.inf), np.nextafter(dn, -np.inf)
        cases += [up, dn]
    rng = np.random.default_rng(5)
    cases.append(10.0 ** rng.uniform(-8.0, 4.0, 200000))           # -80 .. +40 dB re full scale: below the gate and past the top bin
    cases.append(np.array([bounds[0] * 0.5, bounds[1000], bounds[1000] * 4.0, 1e300, 5e-324, 0.0]))
    e = np.concatenate(cases)
    got = b.histogram_index(e)
    want = np.array([oracle.find_histogram_index(x) if x >= bounds[0] else -1 for x in e], dtype=np.int32)
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, (e[bad[:5]], got[bad[:5]], want[bad[:5]])
