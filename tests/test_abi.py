"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and refuses to run without a device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "soundscope_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ssb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(ssb):
    L = ctypes.CDLL(ssb.library_path())
    names = header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/soundscope_b200.h but not exported"
    from soundscope_b200._lib import SYMBOLS
    assert sorted(SYMBOLS) == names


def test_abi_version(ssb):
    assert ssb.lib().ssb_abi_version() == 2


def test_no_cpu_fallback(ssb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(ssb.SsbError) as e:
        ssb.Analyzer()
    assert e.value.code == 13  # SSB_ERR_NO_DEVICE


def test_product_does_not_reference_oracle():
    pkg = os.path.join(ROOT, "soundscope_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
                if f.endswith((".cu", ".cuh", ".cpp")):
                    assert '#include "oracle' not in src and "orc_" not in src, f


def test_stateless_abi_calls(ssb):
    L = ssb.lib()
    k0, nb = ctypes.c_size_t(0), ctypes.c_size_t(0)
    assert L.ssb_fft_bins(16384, 44100, ctypes.byref(k0), ctypes.byref(nb)) == 0
    assert (k0.value, nb.value) == (8, 7423)
    assert L.ssb_fft_bins(16384, 48000, ctypes.byref(k0), ctypes.byref(nb)) == 0
    assert (k0.value, nb.value) == (7, 6820)
    assert L.ssb_fft_bins(8192, 48000, ctypes.byref(k0), ctypes.byref(nb)) == 0
    assert (k0.value, nb.value) == (4, 3410)
    assert L.ssb_fft_bins(1000, 48000, ctypes.byref(k0), ctypes.byref(nb)) == 7   # not a power of two
    assert L.ssb_fft_bins(65536, 48000, ctypes.byref(k0), ctypes.byref(nb)) == 7  # above the crate's largest size
    assert L.ssb_fft_bins(1, 48000, ctypes.byref(k0), ctypes.byref(nb)) == 4
    assert L.ssb_fft_bins(4096, 22050, ctypes.byref(k0), ctypes.byref(nb)) == 8   # 20 kHz above Nyquist
    assert L.ssb_fft_bins(4096, 40000, ctypes.byref(k0), ctypes.byref(nb)) == 0   # 20000 <= 20000


def test_fft_axis_matches_oracle(ssb, oracle):
    import numpy as np
    from tests.signals import ref_sine_f32
    L = ssb.lib()
    for n, rate in ((16384, 44100), (16384, 48000), (8192, 48000), (4096, 96000)):
        k0, nb = oracle.fft_bin_range(n, rate)
        x, t = np.empty(nb), np.empty(nb)
        m = ctypes.c_size_t(0)
        assert L.ssb_fft_axis(n, rate, x.ctypes.data, t.ctypes.data, nb, ctypes.byref(m)) == 0
        assert m.value == nb
        ref = oracle.get_fft(ref_sine_f32(1000.0, n, rate), rate)
        assert np.array_equal(x, ref[:, 0])  # chart x depends only on (n, rate): bit-exact


def test_capture_entry_points_without_device(ssb):
    """stateless parts of the capture ABI work anywhere; anything that needs memory fails loudly without a device"""
    import torch
    L = ssb.lib()
    sizes = {"u8": 1, "s8": 1, "s16le": 2, "s16be": 2, "s24le": 3, "s24be": 3, "s32le": 4, "s32be": 4,
             "f32le": 4, "f32be": 4, "f64le": 8, "f64be": 8}
    for name, code in ssb.PCM_FORMATS.items():
        assert L.ssb_pcm_bytes_per_sample(code) == sizes[name]
    assert L.ssb_pcm_bytes_per_sample(99) == 0 and L.ssb_pcm_bytes_per_sample(-1) == 0
    if not torch.cuda.is_available():
        with pytest.raises(ssb.SsbError) as e:
            ssb.CaptureRing(1000)
        assert e.value.code == 13  # SSB_ERR_NO_DEVICE: the ring is pinned + device memory, there is no host-only ring


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU oracle on the host cores) prints one JSON line with the contract's keys;
    it needs no GPU, so the driver-facing format is checked here."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "samples/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == "samples/s" and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert "cfg2" in line["config"]["workload"] and line["vs_baseline"] is None and line["gpu_launches"] == 0
    # Mode::all() (what the reference's Analyzer always builds, analyzer.rs:36,51,171) is reported beside the loudness modes
    assert line["value_all"] > 0 and line["cpu_baseline_all"]["kind"] == "port" and line["e2e_all"]["value"] == line["value_all"]
    # the config object is the one the GPU arm prints for the same --gpus (the driver compares them)
    import bench
    assert line["config"] == bench.bench_config(1)


def test_header_is_plain_c_and_links_from_c(ssb):
    """the boundary is a C ABI: include/soundscope_b200.h compiles with gcc -std=c11 -pedantic and a C program links
    against the shared library (create fails loudly with SSB_ERR_NO_DEVICE when there is no GPU)"""
    import subprocess
    import tempfile
    import torch
    prog = r'''
#include "soundscope_b200.h"
#include <stdio.h>
int main(void) {
  ssb_analyzer* h = NULL;
  ssb_capture_ring* r = NULL;
  int32_t rc = ssb_analyzer_create(&h, 2, 44100, SSB_MODE_ALL, 1, -1, SSB_FLAG_RING);
  int32_t rr = ssb_capture_ring_create(&r, 1000, -1);
  printf("%d %d %u %zu\n", rc, rr, ssb_abi_version(), ssb_pcm_bytes_per_sample(SSB_PCM_S24LE));
  if (rc == SSB_OK) ssb_analyzer_destroy(h);
  if (rr == SSB_OK) ssb_capture_ring_destroy(r);
  return rc * 100 + rr;
}
'''
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(src, "w").write(prog)
        libdir = os.path.dirname(ssb.library_path())
        subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                               src, "-o", exe, "-L", libdir, "-lsoundscope_b200", f"-Wl,-rpath,{libdir}"])
        r = subprocess.run([exe], capture_output=True, text=True)
        want = 0 if torch.cuda.is_available() else 13 * 100 + 13
        assert r.returncode == want % 256, r.stdout + r.stderr


def test_bench_clock_sampler_parsing():
    """bench.py's nvidia-smi sampler: rows inside the timed region are told apart by nvidia-smi's own timestamps"""
    import datetime
    import sys
    sys.path.insert(0, ROOT)
    import bench
    c = bench.ClockSampler(0)
    c.rows = ["2026/10/17 01:00:00.000, 0, 120, 1965, 140.2, 0x0000000000000001, Not Active, Not Active, Not Active, Not Active",
              "2026/10/17 01:00:01.020, 0, 1965, 1965, 640.0, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active",
              "2026/10/17 01:00:01.040, 0, 1950, 1965, 905.5, 0x0000000000000004, Not Active, Not Active, Not Active, Active",
              "2026/10/17 01:00:01.060, 0, 1965, 1965, 700.0, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active",
              "garbage", "2026/10/17 01:00:02.000, 0, 120, 1965, 141.0, 0x0000000000000001, Not Active, Not Active, Not Active, Not Active"]
    c.t_begin = datetime.datetime(2026, 10, 17, 1, 0, 1, 0)
    c.t_end = datetime.datetime(2026, 10, 17, 1, 0, 1, 100000)
    out = c.summarise()
    assert out == {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"], "samples": 3,
                   "samples_in_timed_region": 3, "samples_total": 5}
    c.t_begin = c.t_end = None          # no marks: every row counts
    assert c.summarise()["samples"] == 5 and c.summarise()["sm_mhz"] == 1950.0
    assert bench.ClockSampler.parse_row("x, 0, n/a, 1965, 1, 0, a, b, c, d") is None
