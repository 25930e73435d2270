"""Timing of the batched mid/side FFT (BASELINE config 3 shape): prints windows/s and algorithmic GB/s."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import soundscope_b200 as S

n = int(os.environ.get("NFFT", 8192))
w = int(os.environ.get("NWIN", 16384))
torch.cuda.set_device(0)
an = S.BatchAnalyzer(1, 2, 48000, S.MODE_LOUDNESS, device=0)
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = (torch.rand((w, n, 2), generator=g, device="cuda") * 2 - 1).contiguous()
out = an.fft_batch_device(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 5
e0.record()
for _ in range(K):
    an.fft_batch_device(x, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
nb = out.shape[2]
byts = w * (n * 2 * 4 + 2 * nb * 4)
print(f"fft n={n} windows={w}: {ms*1e3:.1f} us -> {w/(ms*1e-3):.3e} stereo windows/s, {byts/(ms*1e-3)/1e9:.0f} GB/s ({byts/(ms*1e-3)/1e9/6514.8*100:.1f}% of 6514.8)")
