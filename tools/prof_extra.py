"""ncu driver for the kernels outside the cfg2 step: 6-channel 96 kHz streams (k_loudness_rows_any, Mode::all),
PCM conversion (k_pcm_to_f32), one microphone tick (k_ring_tick + k_fft_fast + k_loudness_scan + k_results)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import soundscope_b200 as S

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(1)
n6, f6 = 16384, 9600
x6 = (torch.rand((n6, f6, 6), generator=g, device=dev) - 0.5).contiguous()
for md in (S.MODE_LOUDNESS, S.MODE_ALL):
    an = S.BatchAnalyzer(n6, 6, 96000, md, device=0)
    for i in range(2):
        an.add_frames_device(x6)
    torch.cuda.synchronize()
    del an
del x6
an = S.BatchAnalyzer(1, 2, 48000, S.MODE_LOUDNESS, device=0)
for fmt, bps in (("s16le", 2), ("s24le", 3)):
    raw = torch.randint(0, 256, ((1 << 27) * bps,), generator=g, device=dev, dtype=torch.uint8)
    out = torch.empty(1 << 27, dtype=torch.float32, device=dev)
    for i in range(2):
        an.pcm_to_f32_device(raw, fmt, out=out)
    torch.cuda.synchronize()
    del raw, out
single = S.Analyzer(device=0)
single.create_loudness_meter(2, 48000)
ring = S.CaptureRing(30 * 48000, device=0)
ring.push(np.random.default_rng(6).uniform(-0.5, 0.5, 30 * 48000).astype(np.float32))
for i in range(3):
    ring.push(np.random.default_rng(7 + i).uniform(-0.5, 0.5, 768).astype(np.float32))
    single.analyze_microphone_input(ring)
print("done")
# whole-file one-shot (file-mode k_loudness_scan + k_file_gating) on a 60 s stereo file
from soundscope_b200.synth import sweep_stereo
whole = np.tile(sweep_stereo(10.0, 48000), 6)
for i in range(2):
    single.calculate_integrated_lufs(2, whole)
print("done one-shot")
