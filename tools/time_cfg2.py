"""Quick timing of the cfg2 loudness step (device-resident), for kernel tuning: prints ms/step and GB/s."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import soundscope_b200 as S
import bench
from bench import make_input_device_chunked, N_STREAMS, FRAMES
bench.CHANNELS = CHANNELS = int(os.environ.get("CHANNELS", 2))
bench.RATE = RATE = int(os.environ.get("RATE", 48000))
FRAMES = int(os.environ.get("FRAMES", FRAMES))

n = int(os.environ.get("N_STREAMS", N_STREAMS))
mode = S.MODE_ALL if "--all" in sys.argv else S.MODE_LOUDNESS
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
an = S.BatchAnalyzer(n, CHANNELS, RATE, mode, device=0)
xs = [make_input_device_chunked(torch, n, FRAMES, 1234 + i, dev, chunk=2048) for i in range(1 if n * FRAMES * CHANNELS * 4 > 8e9 else 2)]
xs = xs * 2
for i in range(4):
    an.add_frames_device(xs[i & 1])
torch.cuda.synchronize()
an.profile(True)
K = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(K):
    an.add_frames_device(xs[i & 1])
e1.record()
torch.cuda.synchronize()
ms, cnt = an.profile_read()
step = e0.elapsed_time(e1) / K
b = n * FRAMES * CHANNELS * 4
print(f"ch={CHANNELS} rate={RATE} frames={FRAMES} n={n} mode={'all' if mode==S.MODE_ALL else 'loudness'}: step {step*1e3:.1f} us, filter kernel {ms/cnt*1e3:.1f} us "
      f"-> {b/(ms/cnt*1e-3)/1e9:.0f} GB/s ({b/(ms/cnt*1e-3)/1e9/6514.8*100:.1f}% of 6514.8), {n*FRAMES*CHANNELS/(step*1e-3):.3e} samples/s")
