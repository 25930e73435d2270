"""The C++ host-side mirror (include/soundscope_b200.hpp) compiles against the C ABI header and links with
the shared library (no GPU needed: compile + link only; running it needs a device and is covered by the
ctypes path, which makes the same calls)."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROG = r'''
#include "soundscope_b200.hpp"
#include <cstdio>
int main() {
  try {
    soundscope::Analyzer a;
    a.create_loudness_meter(2, 48000);
    std::vector<float> x(16384, 0.0f);
    a.add_samples(x);
    auto fft = a.get_fft(x);
    auto wf = a.get_waveform(x, 1.0);
    auto ms = a.get_mid_and_side_samples(x);
    soundscope::CaptureRing ring(30 * 48000);
    ring.push(x.data(), x.size(), false);
    ring.push(x.data(), 100, true);
    auto mic = a.analyze_microphone_input(ring.handle());
    auto tick = a.process_tick(std::vector<float>(32768, 0.0f));
    const short pcm[4] = {0, 16384, -32768, 32767};
    auto f = a.pcm_to_f32(pcm, 4, SSB_PCM_S16LE);
    if (f[1] != 0.5f || f[2] != -1.0f || mic.waveform.size() != 30000 || tick.mid_fft.size() != 6820) return 2;
    std::printf("%f %zu %zu %zu\n", a.get_integrated_lufs(), fft.size(), wf.size(), ms.first.size());
  } catch (const soundscope::Error& e) {
    std::printf("error %d %s\n", e.code, e.what());
    return e.code == SSB_ERR_NO_DEVICE ? 42 : 1;
  }
  return 0;
}
'''


import pytest


@pytest.mark.gpu
def test_cpp_mirror_runs_on_gpu(ssb, cuda):
    """the same program on the GPU box: every mirror method returns reference-shaped results"""
    test_cpp_mirror_compiles_links_and_fails_loudly_without_gpu(ssb)


def test_cpp_mirror_compiles_links_and_fails_loudly_without_gpu(ssb):
    import torch
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.cpp")
        exe = os.path.join(d, "t")
        open(src, "w").write(PROG)
        libdir = os.path.dirname(ssb.library_path())
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                               "-L", libdir, "-lsoundscope_b200", f"-Wl,-rpath,{libdir}"])
        r = subprocess.run([exe], capture_output=True, text=True)
        if torch.cuda.is_available():
            assert r.returncode == 0, r.stdout + r.stderr
        else:
            assert r.returncode == 42, r.stdout + r.stderr   # SSB_ERR_NO_DEVICE: no CPU fallback


def test_native_bench_driver_compiles_and_links(ssb):
    """tools/bench_cabi.cu: the cfg2 step through the C ABI from plain C++/CUDA (no Python, no torch).  Compile + link
    here; without a device it exits non-zero at its first CUDA call (no CPU fallback anywhere)."""
    import torch
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "bench_cabi")
        libdir = os.path.dirname(ssb.library_path())
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                               "-ccbin", "/usr/bin/g++", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "bench_cabi.cu"),
                               "-o", exe, "-L", libdir, "-lsoundscope_b200", "-Xlinker", "-rpath", "-Xlinker", libdir])
        r = subprocess.run([exe, "64", "9600", "2", "48000", "loudness", "2", "1"], capture_output=True, text=True)
        if torch.cuda.is_available():
            assert r.returncode == 0 and '"driver": "bench_cabi"' in r.stdout, r.stdout + r.stderr
        else:
            assert r.returncode != 0
