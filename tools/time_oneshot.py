"""Latency of the whole-file one-shot (Analyzer::calculate_integrated_lufs) for a 10 s and a 5 min 48 kHz stereo file."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import soundscope_b200 as S
from soundscope_b200.synth import sweep_stereo

a = S.Analyzer()
a.create_loudness_meter(2, 48000)
x = sweep_stereo(10.0, 48000)
for name, y, reps in (("10 s", x, 20), ("5 min", np.tile(x, 30), 5)):
    for i in range(3):
        v = a.calculate_integrated_lufs(2, y)
    t0 = time.perf_counter()
    for i in range(reps):
        v = a.calculate_integrated_lufs(2, y)
    print(f"one-shot {name}: {(time.perf_counter() - t0) / reps * 1e3:.3f} ms  ({v:.6f} LUFS, {y.size * 4 / 1e6:.1f} MB pageable host input)")
