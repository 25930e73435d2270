"""ncu driver: cfg2 steps exactly as bench.py runs them (filter launch + results query every step)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import soundscope_b200 as S
from bench import make_input_device, N_STREAMS, FRAMES, CHANNELS, RATE
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
an = S.BatchAnalyzer(N_STREAMS, CHANNELS, RATE, S.MODE_LOUDNESS, device=0)
xs = [make_input_device(torch, N_STREAMS, FRAMES, 1234 + i, dev) for i in range(2)]
res = torch.empty((N_STREAMS, an.stride), dtype=torch.float64, device=dev)
for i in range(8):
    an.add_frames_device(xs[i & 1])
    an.results_device(res)
torch.cuda.synchronize()
print("done", an.launches)
