"""A/B timing of the cfg2 filter kernels: round-1 tile kernel (3), wtile mixed (5), wtile uniform (6), automatic (0);
MODE_LOUDNESS and MODE_ALL; plus the fused add+results call.  Prints one line per variant."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import soundscope_b200 as S
from bench import make_input_device_chunked

n = int(os.environ.get("N_STREAMS", 4096)); CH = int(os.environ.get("CHANNELS", 2)); RATE = int(os.environ.get("RATE", 48000))
FRAMES = int(os.environ.get("FRAMES", 19200))
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
xs = [make_input_device_chunked(torch, n, FRAMES, 1234 + i, dev, chunk=2048, channels=CH) for i in range(2)]
peak = 6514.8
MODES = os.environ.get("MODES", "loudness,all").split(",")
FORCES = [int(v) for v in os.environ.get("FORCES", "3,5,6,0").split(",")]
for mode_name, mode in (("loudness", S.MODE_LOUDNESS), ("all", S.MODE_ALL)):
    if mode_name not in MODES:
        continue
    for fk in FORCES:
        an = S.BatchAnalyzer(n, CH, RATE, mode, device=0)
        an.force_kernel(fk)
        res = torch.empty((n, an.stride), dtype=torch.float64, device=dev)
        for i in range(4):
            an.add_frames_device(xs[i & 1])
        torch.cuda.synchronize()
        an.profile(True)
        K = 25
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            an.add_frames_device(xs[i & 1])
        e1.record()
        torch.cuda.synchronize()
        ms, cnt = an.profile_read()
        an.profile(False)
        # fused feed + results, and the two-call form
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(K):
            an.add_frames_results_device(xs[i & 1], res)
        f1.record()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for i in range(K):
            an.add_frames_device(xs[i & 1])
            an.results_device(res)
        g1.record()
        torch.cuda.synchronize()
        b = n * FRAMES * CH * 4
        k_us = ms / cnt * 1e3
        print(f"n={n} ch={CH} rate={RATE} frames={FRAMES} mode={mode_name} force={fk}: filter kernel {k_us:.1f} us = "
              f"{b / (k_us * 1e-6) / 1e9:.0f} GB/s ({b / (k_us * 1e-6) / 1e9 / peak * 100:.1f}% of {peak}); "
              f"step(add only) {e0.elapsed_time(e1) / K * 1e3:.1f} us; fused add+results {f0.elapsed_time(f1) / K * 1e3:.1f} us; "
              f"add + results (2 calls) {g0.elapsed_time(g1) / K * 1e3:.1f} us", flush=True)
        del an
