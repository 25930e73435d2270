"""Small driver for ncu: a few cfg2 steps through the C ABI (device-resident inputs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import soundscope_b200 as S
from bench import make_input_device, N_STREAMS, FRAMES, CHANNELS, RATE

mode = S.MODE_ALL if "--all" in sys.argv else S.MODE_LOUDNESS
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
an = S.BatchAnalyzer(N_STREAMS, CHANNELS, RATE, mode, device=0)
an.force_kernel(int(os.environ.get('FORCE', 0)))
xs = [make_input_device(torch, N_STREAMS, FRAMES, 1234 + i, dev) for i in range(2)]
for i in range(int(os.environ.get('N_LAUNCH', 6))):
    if os.environ.get('NOFUSE'):
        an.add_frames_device(xs[i & 1])
    else:
        res = an.add_frames_results_device(xs[i & 1])
res = an.results_device()
torch.cuda.synchronize()
if "--fft" in sys.argv:
    xf = make_input_device(torch, 4096, 8192, 99, dev)
    for i in range(3):
        out = an.fft_batch_device(xf)
    torch.cuda.synchronize()
print("done", an.launches)
