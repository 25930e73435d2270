"""Batched, device-resident form of the analyzer path: n_streams independent meters per handle and
batched spectra, driven with torch CUDA tensors (torch is plumbing: device memory, streams, NCCL).

Layouts follow the C ABI (include/soundscope_b200.h): loudness input [n_streams, frames, channels] f32;
FFT input [n_windows, n] (mono) or [n_windows, n, 2] (stereo -> mid/side).
"""
import ctypes as C

import numpy as np

from ._lib import FFT_MID_SIDE, FFT_MONO, MODE_ALL, SsbError, check, lib


class BatchAnalyzer:
    def __init__(self, n_streams, channels=2, rate=48000, mode=MODE_ALL, device=-1, flags=0, use_torch_stream=True):
        self._h = C.c_void_p()
        rc = lib().ssb_analyzer_create(C.byref(self._h), channels, rate, mode, n_streams, device, flags)
        if rc:
            raise SsbError(rc, "ssb_analyzer_create")
        self.n_streams, self.channels, self.rate, self.mode = n_streams, channels, rate, mode
        self.stride = lib().ssb_result_stride(self._h)
        if use_torch_stream:
            import torch
            check(self._h, lib().ssb_set_stream(self._h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def close(self):
        if getattr(self, "_h", None):
            lib().ssb_analyzer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return lib().ssb_launch_count(self._h)

    def force_generic(self, on=True):
        check(self._h, lib().ssb_debug_force_generic(self._h, 1 if on else 0))

    def force_kernel(self, which):
        """tests: 0 automatic, 1 generic, 2 serial many-streams kernel, 3 round-1 tile kernel, 4 scan kernel,
        5 / 6 warp-pipelined batch kernel (mixed T4/T5 warps / uniform T4 warps)"""
        check(self._h, lib().ssb_debug_force_generic(self._h, int(which)))

    def true_peak_factor(self):
        return lib().ssb_true_peak_factor(self._h)

    def force_true_peak_factor(self, factor):
        """benchmarks only: 2 or 4 regardless of ebur128's rate rule (results are then NOT the reference's)"""
        check(self._h, lib().ssb_debug_force_true_peak_factor(self._h, int(factor)))

    def profile(self, on=True):
        check(self._h, lib().ssb_profile_enable(self._h, 1 if on else 0))

    def profile_read(self):
        """(total filter-kernel ms, launches) since the last read; CUDA-event time on the handle's stream."""
        ms, n = C.c_double(0), C.c_uint64(0)
        check(self._h, lib().ssb_profile_read(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def sync(self):
        check(self._h, lib().ssb_sync(self._h))

    def reset(self):
        check(self._h, lib().ssb_reset(self._h))

    def add_frames_device(self, x):
        """x: torch.float32 CUDA tensor [n_streams, frames, channels], contiguous."""
        assert x.is_cuda and x.is_contiguous() and x.dtype.is_floating_point and x.element_size() == 4
        assert x.shape[0] == self.n_streams and x.shape[2] == self.channels
        check(self._h, lib().ssb_add_frames_f32_device(self._h, C.c_void_p(x.data_ptr()), x.shape[1]))

    def add_frames_results_device(self, x, out=None):
        """add_frames_device + results_device as one C-ABI call (one kernel launch when the batch kernel applies).
        Returns the [n_streams, 4+2C] f64 CUDA tensor of result rows for the position after the feed."""
        import torch
        assert x.is_cuda and x.is_contiguous() and x.dtype.is_floating_point and x.element_size() == 4
        assert x.shape[0] == self.n_streams and x.shape[2] == self.channels
        if out is None:
            out = torch.empty((self.n_streams, self.stride), dtype=torch.float64, device=x.device)
        check(self._h, lib().ssb_add_frames_f32_device_results(self._h, C.c_void_p(x.data_ptr()), x.shape[1],
                                                               C.c_void_p(out.data_ptr())))
        return out

    def add_frames_host(self, x):
        """x: host float32 array / pinned tensor [n_streams, frames, channels]; H2D copy is inside the call."""
        if hasattr(x, "data_ptr"):
            assert x.is_contiguous() and not x.is_cuda
            ptr, frames = x.data_ptr(), x.shape[1]
        else:
            x = np.ascontiguousarray(x, dtype=np.float32)
            ptr, frames = x.ctypes.data, x.shape[1]
        check(self._h, lib().ssb_add_frames_f32(self._h, C.c_void_p(ptr), frames))

    def add_frames_pcm_device(self, raw, fmt):
        """raw: uint8 CUDA tensor holding [n_streams, frames, channels] interleaved PCM samples of format `fmt`."""
        from .capture import pcm_format
        code = pcm_format(fmt)
        bps = lib().ssb_pcm_bytes_per_sample(code)
        assert raw.is_cuda and raw.is_contiguous() and raw.element_size() == 1
        frames = raw.numel() // (bps * self.n_streams * self.channels)
        assert frames * bps * self.n_streams * self.channels == raw.numel()
        check(self._h, lib().ssb_add_frames_pcm_device(self._h, C.c_void_p(raw.data_ptr()), code, frames))

    def add_frames_pcm_host(self, raw, fmt):
        """raw: host uint8 array / pinned tensor of interleaved PCM; the raw bytes cross PCIe, conversion is on the device."""
        from .capture import pcm_format
        code = pcm_format(fmt)
        bps = lib().ssb_pcm_bytes_per_sample(code)
        if hasattr(raw, "data_ptr"):
            ptr, nbytes = raw.data_ptr(), raw.numel() * raw.element_size()
        else:
            raw = np.ascontiguousarray(raw).view(np.uint8).ravel()
            ptr, nbytes = raw.ctypes.data, raw.size
        frames = nbytes // (bps * self.n_streams * self.channels)
        assert frames * bps * self.n_streams * self.channels == nbytes
        check(self._h, lib().ssb_add_frames_pcm(self._h, C.c_void_p(ptr), code, frames))

    def pcm_to_f32_device(self, raw, fmt, out=None):
        """raw: uint8 CUDA tensor of interleaved PCM -> float32 CUDA tensor (same sample order)."""
        import torch
        from .capture import pcm_format
        code = pcm_format(fmt)
        bps = lib().ssb_pcm_bytes_per_sample(code)
        assert raw.is_cuda and raw.is_contiguous() and raw.element_size() == 1
        n = raw.numel() // bps
        if out is None:
            out = torch.empty(n, dtype=torch.float32, device=raw.device)
        check(self._h, lib().ssb_pcm_to_f32_device(self._h, C.c_void_p(raw.data_ptr()), n, code, C.c_void_p(out.data_ptr())))
        return out

    # ---- multi-GPU gather of the result rows (include/soundscope_b200.h, "multi-GPU") ----------------
    def gather_create(self, world, rank):
        """-> this rank's 64-byte CUDA IPC handle of its gather buffer"""
        buf = C.create_string_buffer(64)
        check(self._h, lib().ssb_gather_create(self._h, world, rank, buf))
        return bytes(buf.raw)

    def gather_open(self, handles):
        """handles: world x 64 bytes in rank order (own slot ignored)"""
        check(self._h, lib().ssb_gather_open(self._h, C.c_char_p(handles)))

    def gather_select(self, parity):
        check(self._h, lib().ssb_gather_select(self._h, int(parity)))

    def gather_wait(self):
        check(self._h, lib().ssb_gather_wait(self._h))

    def gather_epoch(self):
        return lib().ssb_gather_epoch(self._h)

    def gather_rows(self, parity, world):
        """[world * n_streams, stride] f64 CUDA tensor aliasing half `parity` of this rank's gather buffer"""
        import torch
        ptr = lib().ssb_gather_rows(self._h, int(parity))
        if not ptr:
            raise SsbError(3, "no gather buffer")
        n = world * self.n_streams * self.stride

        class _Mem:   # __cuda_array_interface__ view of library-owned memory (the handle outlives the tensor's use)
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_Mem(), device="cuda").view(world * self.n_streams, self.stride)

    def gather_destroy(self):
        check(self._h, lib().ssb_gather_destroy(self._h))

    def results_device(self, out=None):
        """[n_streams, 4+2C] f64 CUDA tensor: momentary, shortterm, integrated, LRA, true_peak[C], sample_peak[C]."""
        import torch
        if out is None:
            out = torch.empty((self.n_streams, self.stride), dtype=torch.float64, device="cuda")
        check(self._h, lib().ssb_results_device(self._h, C.c_void_p(out.data_ptr())))
        return out

    def _col(self, fn, width=None):
        out = np.empty(self.n_streams * (width or 1), dtype=np.float64)
        check(self._h, fn(self._h, out.ctypes.data))
        return out if width is None else out.reshape(self.n_streams, width)

    def loudness_momentary(self):
        return self._col(lib().ssb_loudness_momentary)

    def loudness_shortterm(self):
        return self._col(lib().ssb_loudness_shortterm)

    def loudness_global(self):
        return self._col(lib().ssb_loudness_global)

    def loudness_range(self):
        return self._col(lib().ssb_loudness_range)

    def true_peak(self):
        return self._col(lib().ssb_true_peak, self.channels)

    def sample_peak(self):
        return self._col(lib().ssb_sample_peak, self.channels)

    def histogram_index(self, energies):
        """find_histogram_index as the gating kernels evaluate it: bin per block energy, -1 below the absolute gate."""
        e = np.ascontiguousarray(energies, dtype=np.float64)
        out = np.empty(e.size, dtype=np.int32)
        check(self._h, lib().ssb_debug_histogram_index(self._h, e.ctypes.data, e.size, out.ctypes.data))
        return out

    def histograms(self, stream):
        blk, st = np.zeros(1000, dtype=np.uint64), np.zeros(1000, dtype=np.uint64)
        check(self._h, lib().ssb_histograms(self._h, stream, blk.ctypes.data, st.ctypes.data))
        return blk, st

    # ---- spectrum -------------------------------------------------------------------------------
    def fft_bins(self, n):
        k0, nb = C.c_size_t(0), C.c_size_t(0)
        check(self._h, lib().ssb_fft_bins(n, self.rate, C.byref(k0), C.byref(nb)))
        return k0.value, nb.value

    def fft_axis(self, n):
        _, nb = self.fft_bins(n)
        x, tilt = np.empty(nb), np.empty(nb)
        m = C.c_size_t(0)
        check(self._h, lib().ssb_fft_axis(n, self.rate, x.ctypes.data, tilt.ctypes.data, nb, C.byref(m)))
        return x, tilt

    def fft_batch_device(self, x, out=None, status=None):
        """x: [W, n] (mono) or [W, n, 2] (stereo -> mid, side) f32 CUDA tensor -> dB [W, planes, n_bins] f32."""
        import torch
        assert x.is_cuda and x.is_contiguous()
        assert x.data_ptr() % 16 == 0, "fft_batch_device: the input view must start on a 16-byte boundary"
        layout = FFT_MID_SIDE if x.dim() == 3 else FFT_MONO
        w, n = x.shape[0], x.shape[1]
        _, nb = self.fft_bins(n)
        planes = 2 if layout == FFT_MID_SIDE else 1
        if out is None:
            out = torch.empty((w, planes, nb), dtype=torch.float32, device=x.device)
        sp = C.c_void_p(status.data_ptr()) if status is not None else None
        check(self._h, lib().ssb_fft_batch_device(self._h, C.c_void_p(x.data_ptr()), layout, n, w,
                                                  C.c_void_p(out.data_ptr()), sp))
        return out

    def fft_batch_y_device(self, x, out=None, status=None):
        """Like fft_batch_device, but the kernel writes the reference's y = (f64) dB + tilt (analyzer.rs:80-94):
        [W, planes, n_bins] f64; pair it with fft_axis(n)[0] for the chart points."""
        import torch
        assert x.is_cuda and x.is_contiguous()
        assert x.data_ptr() % 16 == 0, "fft_batch_y_device: the input view must start on a 16-byte boundary"
        layout = FFT_MID_SIDE if x.dim() == 3 else FFT_MONO
        w, n = x.shape[0], x.shape[1]
        _, nb = self.fft_bins(n)
        planes = 2 if layout == FFT_MID_SIDE else 1
        if out is None:
            out = torch.empty((w, planes, nb), dtype=torch.float64, device=x.device)
        sp = C.c_void_p(status.data_ptr()) if status is not None else None
        check(self._h, lib().ssb_fft_batch_device_y(self._h, C.c_void_p(x.data_ptr()), layout, n, w,
                                                    C.c_void_p(out.data_ptr()), sp))
        return out

    def waveform_device(self, x, waveform_window):
        """x: 1-D f32 CUDA tensor -> [columns, 2] (min, max) f32 CUDA tensor."""
        import torch
        assert x.is_cuda and x.is_contiguous()
        n = C.c_size_t(0)
        lib().ssb_waveform_device(self._h, None, x.numel(), float(waveform_window), None, 0, C.byref(n))
        out = torch.empty((n.value, 2), dtype=torch.float32, device=x.device)
        check(self._h, lib().ssb_waveform_device(self._h, C.c_void_p(x.data_ptr()), x.numel(), float(waveform_window),
                                                 C.c_void_p(out.data_ptr()), n.value, C.byref(n)))
        return out

    def mid_side_device(self, x):
        import torch
        assert x.is_cuda and x.is_contiguous()
        assert x.data_ptr() % 8 == 0, "mid_side_device: the input view must start on an 8-byte boundary"
        frames = x.numel() // 2
        mid = torch.empty(frames, dtype=torch.float32, device=x.device)
        side = torch.empty(frames, dtype=torch.float32, device=x.device)
        check(self._h, lib().ssb_mid_side_device(self._h, C.c_void_p(x.data_ptr()), x.numel(),
                                                 C.c_void_p(mid.data_ptr()), C.c_void_p(side.data_ptr())))
        return mid, side
