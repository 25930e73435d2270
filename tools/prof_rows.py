"""ncu driver: the serial many-streams kernel (32768 streams x 400 ms)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import soundscope_b200 as S
from bench import make_input_device, FRAMES, CHANNELS, RATE
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
an = S.BatchAnalyzer(32768, CHANNELS, RATE, S.MODE_LOUDNESS, device=0)
x = make_input_device(torch, 32768, FRAMES, 7, dev)
for i in range(3):
    an.add_frames_device(x)
torch.cuda.synchronize()
print("done")
