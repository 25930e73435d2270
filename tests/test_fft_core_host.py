"""CPU check of the device FFT algorithm: soundscope_b200/csrc/fft_core.cuh is __host__ __device__, so the
stage functions the CUDA kernel runs are executed here on the host (thread loop emulated) and compared with
numpy's double-precision FFT — index arithmetic, padding, twiddle tables and digit reversal included."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tools", "fft_core_host.cu")
SO = os.path.join(ROOT, "tools", "_fft_core_host.so")


@pytest.fixture(scope="module")
def core():
    deps = [SRC, os.path.join(ROOT, "soundscope_b200", "csrc", "fft_core.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-O2", "-std=c++17", "-ccbin", "/usr/bin/g++", "-shared",
                               "-Xcompiler", "-fPIC", "-o", SO, SRC])
    L = ctypes.CDLL(SO)
    L.fft_core_host.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_uint, ctypes.c_void_p]
    return L


@pytest.mark.parametrize("M,N", [(512, 512), (512, 1024), (1024, 1024), (2048, 4096), (4096, 8192), (8192, 8192),
                                 (8192, 16384), (16384, 16384), (16384, 32768)])
def test_core_matches_numpy(core, M, N):
    rng = np.random.default_rng(M + N)
    x = (rng.standard_normal(M) + 1j * rng.standard_normal(M)).astype(np.complex64)
    out = np.empty(M, dtype=np.complex64)
    assert core.fft_core_host(x.ctypes.data, M, N, out.ctypes.data) == 0
    ref = np.fft.fft(x.astype(np.complex128))
    err = np.max(np.abs(out - ref)) / np.max(np.abs(ref))
    assert err < 5e-7, err
