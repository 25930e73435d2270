#!/usr/bin/env python
"""bench.py — the analyzer hot path on BASELINE.json's metric: audio samples/s (48 kHz stereo f32).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg4|cfg5]

Workload (BASELINE configs[1], named in config.workload): 4096 independent 48 kHz stereo streams
per GPU; ONE STEP = 25 consecutive 400 ms launches (10 s of audio per stream, meter state carried: K-weighting,
100 ms energy buckets, block gating histograms), every launch followed by the query of the per-stream scalars
(momentary / short-term / integrated / LRA [/ true peak / sample peak]) — the reference's per-tick pair
`add_samples` + `get_*_lufs` (src/analyzer.rs:139-164).  Under torchrun every rank owns its own 4096 streams
(weak scaling, no data-path collective); the result rows of every launch are gathered to all ranks.

Two mode legs are measured with the same step:
  value / roofline / e2e              ebur128 modes M|S|I|LRA|HISTOGRAM — BASELINE's "fused K-weight+RMS kernel"
  value_all / roofline_all / e2e_all  Mode::all() — what the reference's Analyzer always builds (analyzer.rs:36,51,171):
                                      adds sample peak and the 4x oversampled true peak
  value*     device-resident inputs, whole-job samples/s over all ranks (max-over-ranks device time)
  e2e*       the same through the host-facing C-ABI call (pinned host buffers; H2D of every launch's input and D2H of
             its result rows inside the timed region)
  roofline*  the dominant kernel (the K-weighting filter kernel): algorithmic 4 B/sample over its average CUDA-event
             duration, against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline / --impl reference   the CPU oracle (oracle/: a C restatement of the reference's algorithm; the Rust
             reference cannot be built in this image: no cargo/rustc, crates not vendored) on the box's host cores;
             one reference step is a bounded sample of the same workload (one 400 ms launch of the 4096 streams + query)

--config cfg4 / cfg5 run BASELINE configs[3] / [4] per-GPU shards under torchrun (extra lines for profiles/; the
driver's contract line is the default cfg2).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_STREAMS = 4096
CHANNELS = 2
RATE = 48000
FRAMES = 19200            # 400 ms
LAUNCHES_PER_STEP = 25    # 10 s of audio per stream per step
WORKLOAD = ("cfg2: 4096 streams/GPU x 48 kHz stereo f32, one step = 25 launches of 400 ms (19200 frames) each followed by the "
            "per-stream query, K-weighting + gated RMS (M|S|I|LRA|HISTOGRAM); Mode::all() reported beside it as *_all")
METRIC = "audio samples/sec (48 kHz stereo f32) through FFT+LUFS"
UNIT = "samples/s"


def bench_config(n_gpus):
    """The `config` object of the JSON line — identical in both arms for the same --gpus."""
    return {"workload": WORKLOAD, "streams_per_gpu": N_STREAMS, "frames_per_launch": FRAMES,
            "launches_per_step": LAUNCHES_PER_STEP, "channels": CHANNELS, "rate": RATE,
            "parallelism": f"streams sharded x{n_gpus}, result rows of every launch written to all ranks over NVLink" if n_gpus > 1 else "single GPU",
            "l2": "two alternating 629 MB inputs per GPU (> 126 MB L2), no flush needed"}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def make_input_np(n_streams, frames, seed):
    from soundscope_b200.synth import stream_batch
    return stream_batch(n_streams, frames, CHANNELS, seed=seed, rate=RATE)


def make_input_device(torch, n_streams, frames, seed, device, channels=None, rate=None):
    """cfg2 generator on the device: per-stream tone 100*2^((s%64)/8) Hz at 0.25 + 0.05 uniform noise."""
    ch = channels or CHANNELS
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    t = torch.arange(frames, device=device, dtype=torch.float64) / (rate or RATE)
    s = torch.arange(n_streams, device=device, dtype=torch.float64)
    f = 100.0 * torch.pow(torch.tensor(2.0, device=device, dtype=torch.float64), (s % 64) / 8.0)
    ph = torch.rand((n_streams, 1, ch), generator=g, device=device, dtype=torch.float64) * 6.283185307179586
    x = 0.25 * torch.sin(6.283185307179586 * f[:, None, None] * t[None, :, None] + ph)
    x = x + 0.05 * (torch.rand((n_streams, frames, ch), generator=g, device=device, dtype=torch.float64) * 2 - 1)
    return x.to(torch.float32).contiguous()


def make_input_device_chunked(torch, n_streams, frames, seed, device, chunk=4096, channels=None, rate=None):
    """The same generator filled in slices of `chunk` streams, so its f64 temporaries stay small next to a large batch."""
    ch = channels or CHANNELS
    x = torch.empty((n_streams, frames, ch), dtype=torch.float32, device=device)
    for s0 in range(0, n_streams, chunk):
        n = min(chunk, n_streams - s0)
        x[s0:s0 + n] = make_input_device(torch, n, frames, seed + s0, device, channels=ch, rate=rate)
    return x


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed regions (B200_PROFILING.md recipe).  Rows carry nvidia-smi's
    own timestamp, so the ones that fall inside a marked window are told apart from the idle ones around it whatever the
    pipe buffering; the median is taken over the in-region rows when there are any."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.windows = []
        self.t_begin = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def mark_begin(self):
        import datetime
        self.t_begin = datetime.datetime.now()

    def mark_end(self):
        import datetime
        self.t_end = datetime.datetime.now()
        self.windows.append((self.t_begin, self.t_end))

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    @staticmethod
    def parse_row(r):
        """-> (timestamp or None, sm_mhz, sm_max_mhz, [reason names]) or None for a malformed row"""
        import datetime
        p = [c.strip() for c in r.split(",")]
        if len(p) < 10:
            return None
        try:
            sm, mx = float(p[2]), float(p[3])
        except ValueError:
            return None
        try:
            ts = datetime.datetime.strptime(p[0], "%Y/%m/%d %H:%M:%S.%f")
        except ValueError:
            ts = None
        reasons = [nm for nm, v in zip(ClockSampler.NAMES, p[6:10]) if v.lower().startswith("active")]
        return ts, sm, mx, reasons

    def summarise(self):
        parsed = [q for q in (self.parse_row(r) for r in self.rows) if q is not None]
        wins = self.windows or ([(self.t_begin, self.t_end)] if self.t_begin and self.t_end else [])
        inside = [q for q in parsed if q[0] is not None and any(b <= q[0] <= e for b, e in wins)]
        use = inside if inside else parsed
        sm = sorted(q[1] for q in use)
        reasons = sorted({nm for q in use for nm in q[3]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": use[-1][2] if use else None, "reasons": reasons,
                "samples": len(use), "samples_in_timed_region": len(inside), "samples_total": len(parsed)}

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        return self.summarise()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (bounded sample of the same workload)
# ------------------------------------------------------------------------------------------------
def cpu_oracle_run(mode_name, steps, warmup, threads, n_streams=N_STREAMS, budget_s=60.0):
    """One CPU step = one 400 ms launch of the n_streams streams (add_frames + query of every stream's scalars).
    Returns (samples_per_s, ms_per_step, steps_done, sample_description).  Stops adding steps past budget_s."""
    import oracle as O
    mode = O.MODE_ALL if mode_name == "all" else O.MODE_LOUDNESS
    x = [make_input_np(n_streams, FRAMES, seed=77 + i) for i in range(2)]
    b = O.Batch(n_streams, CHANNELS, RATE, mode, threads=threads)
    for i in range(warmup):
        b.add_frames(x[i & 1], threads=threads)
        b.query(threads=threads)
    t0 = time.perf_counter()
    done = 0
    for i in range(steps):
        b.add_frames(x[i & 1], threads=threads)
        b.query(threads=threads)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    samples = done * n_streams * FRAMES * CHANNELS
    modes = "Mode::all()" if mode_name == "all" else "modes M|S|I|LRA|HISTOGRAM"
    return samples / dt, dt / done * 1e3, done, (f"{done} step(s), each ONE 400 ms launch of {n_streams} streams x {FRAMES} frames x {CHANNELS} ch + query "
                                                f"(1/{LAUNCHES_PER_STEP} of the GPU arm's step; oracle port, {modes}, {threads} threads)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    W = max(1, min(args.warmup, 5))
    v, ms, done, sample = cpu_oracle_run("loudness", max(1, args.steps), W, threads)
    va, msa, donea, samplea = cpu_oracle_run("all", max(1, args.steps), W, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": bench_config(args.gpus),
        "note": "CPU restatement of the reference algorithm (oracle/), all host threads; the Rust reference cannot be built here (no cargo/rustc on this image or the GPU box, crates not vendored)",
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "steps": done},
        "value_all": va, "ms_per_step_all": msa,
        "cpu_baseline_all": {"value": va, "unit": UNIT, "cores": threads, "kind": "port", "sample": samplea},
        "e2e_all": {"value": va, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "steps": donea},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Env:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def timed_device_leg(env, S, an, xs, gather, K, W, launches_per_step, clocks=None):
    """W warm-up steps, then K timed steps (barrier + synchronize on both sides, CUDA events on the launching stream,
    max over ranks).  One step = launches_per_step x (feed one 400 ms launch + result rows [+ gather to all ranks])."""
    torch = env.torch
    res = gather.local_rows() if gather is not None else torch.empty((an.n_streams, an.stride), dtype=torch.float64, device=env.dev)

    def step(i):
        for l in range(launches_per_step):
            an.add_frames_results_device(xs[(i * launches_per_step + l) & 1], res)
            if gather is not None:
                gather.publish()
        if gather is not None:
            gather.wait()

    for i in range(W):
        step(i)
    env.barrier()
    an.profile(True)
    l0 = an.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    if clocks:
        clocks.mark_begin()
    e0.record()
    for i in range(K):
        step(i)
    e1.record()
    env.barrier()
    if clocks:
        clocks.mark_end()
    ms_total = env.max_over_ranks(e0.elapsed_time(e1))
    launches = an.launches - l0
    filt_ms, filt_n = an.profile_read()
    # the filter phase alone: the same kernel launched without its gating / results epilogue (feed only), same inputs
    for i in range(3):
        an.add_frames_device(xs[i & 1])
    env.torch.cuda.synchronize()
    an.profile(True)
    for i in range(launches_per_step * 2):
        an.add_frames_device(xs[i & 1])
    fo_ms, fo_n = an.profile_read()
    an.profile(False)
    return {"ms_total": ms_total, "launches": int(launches), "kernel_ms": filt_ms / max(filt_n, 1), "kernel_launches": int(filt_n),
            "filter_only_ms": fo_ms / max(fo_n, 1)}


def timed_e2e_leg(env, an, hx, K, W, launches_per_step, pcm_fmt=None):
    """Host-facing calls: every launch copies its input from pinned host memory and reads its result rows back."""
    torch = env.torch

    def step(i):
        for l in range(launches_per_step):
            x = hx[(i * launches_per_step + l) & 1]
            if pcm_fmt:
                an.add_frames_pcm_host(x, pcm_fmt)
            else:
                an.add_frames_host(x)
            an.loudness_global()   # D2H read of every stream's result row

    for i in range(W):
        step(i)
    env.barrier()
    t0 = time.perf_counter()
    for i in range(K):
        step(i)
    torch.cuda.synchronize()
    return env.max_over_ranks(time.perf_counter() - t0)


BINDING_LOUDNESS = ("FP64 pipe + issue + shared-memory wavefronts all ~60-67 % busy (13 DFMA per sample with the time-segmentation "
                    "pass), not HBM: DESIGN.md section 3.1")
BINDING_ALL = ("FMA pipe + issue: the 4x true-peak interpolator is 36 f32 FMA per sample on top of the filter (FFMA 1.5, FFMA2 2.0-2.3 "
               "cycles per warp instruction measured, profiles/r2_mb_fma.txt), not HBM: DESIGN.md section 3.1")


def roofline_block(kernel_ms, bytes_per_launch, peak, peak_kind, traffic, kernel_name, filter_only_ms=None,
                   binding=BINDING_LOUDNESS):
    """`achieved` / `frac` are for the kernel of the timed region — the fused launch: K-weighting filter + gating + per-stream
    result rows.  `filter_phase_*` is the same kernel launched without that epilogue (what round 1's separate filter kernel
    was), measured right after the timed region on the same inputs."""
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9 if kernel_ms else None
    out = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
           "frac": (achieved / peak) if achieved else None, "traffic": traffic,
           "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs, burst copy)", "kernel_ms": kernel_ms,
           "algorithmic_bytes_per_launch": bytes_per_launch,
           "binding_resource": binding}
    if filter_only_ms:
        fo = bytes_per_launch / (filter_only_ms * 1e-3) / 1e9
        out.update({"filter_phase_ms": filter_only_ms, "filter_phase_achieved": fo, "filter_phase_frac": fo / peak})
    return out


def traffic_of(key):
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(tpath)).get(key)
    except Exception:
        return None


def run_ours(args):
    env = Env()
    torch = env.torch
    import soundscope_b200 as S
    from soundscope_b200.sharding import PeerGather

    K, W = args.steps, args.warmup
    LPS = LAUNCHES_PER_STEP
    n_gpus = env.world
    peaks, peak_kind = measured_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    samples_per_launch = N_STREAMS * FRAMES * CHANNELS
    samples_per_step = samples_per_launch * LPS

    # two distinct 629 MB inputs, alternated: every launch streams data that is not in the 126 MB L2
    xs = [make_input_device(torch, N_STREAMS, FRAMES, 1234 + 17 * env.rank + i, env.dev) for i in range(2)]
    clocks = ClockSampler(env.local)
    if env.rank == 0:
        clocks.start()

    legs = {}
    for name, mode in (("loudness", S.MODE_LOUDNESS), ("all", S.MODE_ALL)):
        an = S.BatchAnalyzer(N_STREAMS, CHANNELS, RATE, mode, device=env.local)
        gather = PeerGather(an, env) if env.world > 1 else None
        legs[name] = timed_device_leg(env, S, an, xs, gather, K, W, LPS, clocks if env.rank == 0 else None)
        legs[name]["gather"] = gather.kind if gather is not None else None
        if gather is not None:
            gather.close()
        del an
    clk = clocks.stop() if env.rank == 0 else None

    # ---- e2e: host-facing calls, pinned host inputs, H2D + result read-back inside the timed region ----
    hx = [torch.empty((N_STREAMS, FRAMES, CHANNELS), dtype=torch.float32).pin_memory() for _ in range(2)]
    for i in range(2):
        hx[i].copy_(xs[i].cpu())
    e2e = {}
    e2e_W = max(1, min(W, 2))
    for name, mode in (("loudness", S.MODE_LOUDNESS), ("all", S.MODE_ALL)):
        an2 = S.BatchAnalyzer(N_STREAMS, CHANNELS, RATE, mode, device=env.local)
        dt = timed_e2e_leg(env, an2, hx, K, e2e_W, LPS)
        e2e[name] = {"value": n_gpus * samples_per_step * K / dt, "unit": UNIT, "h2d_bytes_per_step": samples_per_step * 4,
                     "d2h_bytes_per_step": LPS * N_STREAMS * an2.stride * 8, "steps": K}
        del an2
    # the same with 16-bit PCM on the wire (what a WAV file hands the analyzer), 2 B/sample over PCIe
    raw16 = [torch.empty((N_STREAMS, FRAMES, CHANNELS), dtype=torch.int16).pin_memory() for _ in range(2)]
    for i in range(2):
        raw16[i].copy_((xs[i].cpu() * 32767.0).round().to(torch.int16))
    an7 = S.BatchAnalyzer(N_STREAMS, CHANNELS, RATE, S.MODE_LOUDNESS, device=env.local)
    k16 = max(1, K // 4)
    dt16 = timed_e2e_leg(env, an7, [r.view(torch.uint8) for r in raw16], k16, 1, LPS, pcm_fmt="s16le")
    e2e_s16 = {"value": n_gpus * samples_per_step * k16 / dt16, "unit": UNIT, "h2d_bytes_per_step": samples_per_step * 2,
               "d2h_bytes_per_step": LPS * N_STREAMS * an7.stride * 8, "steps": k16, "wire_format": "s16le"}
    del an7, raw16, hx

    extras = {}
    if env.rank == 0 and env.world == 1 and not args.no_extras:   # single-GPU runs only: the scaling runs stay lean
        try:
            extras = run_extras(env, S, xs, peak)
        except Exception as ex:  # extras never invalidate the headline
            extras = {"error": repr(ex)}

    env.barrier()
    if env.rank != 0:
        env.close()
        return

    # CPU baseline leg: bounded sample of the same workload on the host cores (rank 0, N=1 only)
    cpu = cpu_all = None
    if env.world == 1 and not args.no_cpu:
        threads = host_threads()
        v, _, _, sample = cpu_oracle_run("loudness", 3, 1, threads, budget_s=12.0)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
        v, _, _, sample = cpu_oracle_run("all", 3, 1, threads, budget_s=15.0)
        cpu_all = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}

    L, A = legs["loudness"], legs["all"]
    value = n_gpus * samples_per_step * K / (L["ms_total"] * 1e-3)
    value_all = n_gpus * samples_per_step * K / (A["ms_total"] * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": W,
        "ms_per_step": L["ms_total"] / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": bench_config(n_gpus),
        "clocks": clk,
        "e2e": e2e["loudness"],
        "gpu_launches": L["launches"],
        "roofline": roofline_block(L["kernel_ms"], samples_per_launch * 4, peak, peak_kind,
                                   traffic_of("filter_kernel_dram_bytes_per_launch"), "k_loudness_wtile<2,0> (K-weighting + 100 ms energy buckets + lean gating/results epilogue)",
                                   L["filter_only_ms"]),
        "cpu_baseline": cpu,
        "ms_per_launch": L["ms_total"] / K / LPS, "launch_overhead_us": (L["ms_total"] / K / LPS - L["kernel_ms"]) * 1e3,
        "gather": L["gather"],
        # ---- Mode::all(): what the reference's Analyzer always runs (analyzer.rs:36,51,171) ----
        "value_all": value_all, "ms_per_step_all": A["ms_total"] / K, "gpu_launches_all": A["launches"],
        "roofline_all": roofline_block(A["kernel_ms"], samples_per_launch * 4, peak, peak_kind,
                                       traffic_of("filter_kernel_all_dram_bytes_per_launch"), "k_loudness_wtile<2,4> (adds sample peak + 4x true-peak FIR, 36 f32 FMA per sample)",
                                       A["filter_only_ms"], BINDING_ALL),
        "e2e_all": e2e["all"], "cpu_baseline_all": cpu_all,
        "e2e_s16": e2e_s16,
        "extras": extras,
        "streams_realtime_equiv": value / (RATE * CHANNELS),
    }
    print(json.dumps(line))
    env.close()


def run_extras(env, S, xs, hbm):
    """Single-GPU extras: the other kernels of the path, each with its algorithmic bytes and fraction of the HBM peak."""
    import numpy as np
    torch, dev, local = env.torch, env.dev, env.local
    extras = {}
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    an = S.BatchAnalyzer(N_STREAMS, CHANNELS, RATE, S.MODE_LOUDNESS, device=local)
    g = torch.Generator(device=dev)
    g.manual_seed(99)

    def time_it(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        a0.record()
        for _ in range(reps):
            fn()
        a1.record()
        torch.cuda.synchronize()
        return a0.elapsed_time(a1) / reps

    # BASELINE config 3: real FFT, mid/side, batch = 65536 windows resident on the device; 8192 points as BASELINE names it and
    # 16384 points, the only size the reference uses (tui.rs:1431,1488) with half the windows (the same 4 GiB of input)
    for nfft, nwin in ((8192, 65536), (16384, 32768)):
        xf = (torch.rand((nwin, nfft, 2), generator=g, device=dev) * 2 - 1).contiguous()
        out = an.fft_batch_device(xf)
        fms = time_it(lambda: an.fft_batch_device(xf, out=out))
        nb = out.shape[2]
        fbytes = nwin * (nfft * 2 * 4 + 2 * nb * 4)
        extras[f"fft{nfft}_midside"] = {"stereo_windows_per_s": nwin / (fms * 1e-3), "samples_per_s": nwin * nfft * 2 / (fms * 1e-3),
                                        "algorithmic_gbs": fbytes / (fms * 1e-3) / 1e9, "frac_of_hbm_peak": fbytes / (fms * 1e-3) / 1e9 / hbm,
                                        "windows": nwin, "kernel_ms": fms, "algorithmic_bytes_per_window": fbytes // nwin}
        del xf, out
    # stateless rows of SURVEY 8a: min-max waveform (a11) and mid/side (a12) over 2^27 stereo-interleaved samples
    big = (torch.rand(1 << 27, generator=g, device=dev) * 2 - 1).contiguous()
    for name, fn, byts in (("waveform", lambda: an.waveform_device(big, 1342.17728), big.numel() * 4 + 2 * 1342177 * 4),
                           ("mid_side", lambda: an.mid_side_device(big), big.numel() * 8)):
        t = time_it(fn)
        extras[name] = {"samples_per_s": big.numel() / (t * 1e-3), "algorithmic_gbs": byts / (t * 1e-3) / 1e9,
                        "frac_of_hbm_peak": byts / (t * 1e-3) / 1e9 / hbm}
    del big
    # one reference-shaped player tick (tui.rs:1482-1552) through the host-facing call
    tail = (np.random.default_rng(5).uniform(-0.5, 0.5, 32768)).astype(np.float32)
    single = S.Analyzer(device=local)
    single.create_loudness_meter(2, RATE)
    for i in range(5):
        single.process_tick(tail, 16384)
    t0 = time.perf_counter()
    for i in range(50):
        single.process_tick(tail, 16384)
    extras["process_tick_us"] = (time.perf_counter() - t0) / 50 * 1e6
    # whole-file pre-analysis (tui.rs:1229-1233): calculate_integrated_lufs on a 10 s / 5 min 48 kHz stereo file
    from soundscope_b200.synth import sweep_stereo
    whole = sweep_stereo(10.0, RATE)
    for i in range(3):
        single.calculate_integrated_lufs(2, whole)
    t0 = time.perf_counter()
    for i in range(10):
        single.calculate_integrated_lufs(2, whole)
    extras["integrated_lufs_10s_file_ms"] = (time.perf_counter() - t0) / 10 * 1e3
    whole5 = np.tile(whole, 30)
    for i in range(2):
        single.calculate_integrated_lufs(2, whole5)
    t0 = time.perf_counter()
    for i in range(3):
        single.calculate_integrated_lufs(2, whole5)
    extras["integrated_lufs_5min_file_ms"] = (time.perf_counter() - t0) / 3 * 1e3
    del whole5
    # BASELINE configs 4 and 5, one GPU's shard (the N-GPU lines are produced by --config cfg4 / cfg5 under torchrun)
    for cfg in ("cfg4", "cfg5", "cfg5_x4"):
        r = run_shard_config(env, S, cfg, steps=3, warmup=1, hbm=hbm)
        extras[cfg + "_shard"] = r
    # decoded PCM -> f32 (2 or 3 B read + 4 B written per sample), 2^28 samples
    for fmt, bps in (("s16le", 2), ("s24le", 3)):
        nraw = 1 << 28
        raw = torch.randint(0, 256, (nraw * bps,), generator=g, device=dev, dtype=torch.uint8)
        outp = torch.empty(nraw, dtype=torch.float32, device=dev)
        tp = time_it(lambda: an.pcm_to_f32_device(raw, fmt, out=outp))
        extras["pcm_" + fmt] = {"samples_per_s": nraw / (tp * 1e-3), "algorithmic_gbs": nraw * (bps + 4) / (tp * 1e-3) / 1e9,
                                "frac_of_hbm_peak": nraw * (bps + 4) / (tp * 1e-3) / 1e9 / hbm, "kernel_ms": tp}
        del raw, outp
    # one microphone tick (tui.rs:1427-1480) through the host-facing call
    ring = S.CaptureRing(30 * RATE, device=local)
    ring.push(np.random.default_rng(6).uniform(-0.5, 0.5, 30 * RATE).astype(np.float32))
    chunk = np.random.default_rng(7).uniform(-0.5, 0.5, 2 * 384).astype(np.float32)
    for i in range(5):
        ring.push(chunk)
        single.analyze_microphone_input(ring)
    t0 = time.perf_counter()
    for i in range(50):
        ring.push(chunk)
        single.analyze_microphone_input(ring)
    extras["mic_tick_us"] = (time.perf_counter() - t0) / 50 * 1e6
    return extras


SHARD_CONFIGS = {
    # BASELINE configs[3]: 10^6 stereo 48 kHz streams over 8 GPUs = 125 000 per GPU, 400 ms per launch, Mode::all scalars gathered
    "cfg4": dict(n=125000, ch=2, rate=48000, frames=19200, mode="all", tp=None,
                 what="cfg4 shard: 125000 stereo 48 kHz streams/GPU x 400 ms per launch, Mode::all() (4x true peak) + gather of LUFS/true-peak rows"),
    # BASELINE configs[4]: 5.1 (6-channel) 96 kHz streams, 16384 per GPU, 200 ms per launch; ebur128's rate rule gives 2x at 96 kHz
    "cfg5": dict(n=16384, ch=6, rate=96000, frames=19200, mode="all", tp=None,
                 what="cfg5 shard: 16384 six-channel 96 kHz streams/GPU x 200 ms per launch, Mode::all() with the reference's 2x true-peak oversampling at 96 kHz (parity configuration)"),
    # the same with BASELINE's literal "4x" at 96 kHz: NOT what ebur128 computes at this rate — labelled non-parity
    "cfg5_x4": dict(n=16384, ch=6, rate=96000, frames=19200, mode="all", tp=4,
                    what="cfg5 shard, NON-PARITY variant: 4x true-peak oversampling forced at 96 kHz (BASELINE wording; ebur128 itself uses 2x at this rate)"),
}


def run_shard_config(env, S, cfg, steps, warmup, hbm, gather_factory=None):
    """One GPU's shard of a BASELINE config: `steps` timed launches (filter + results [+ gather]); returns a dict."""
    torch, dev = env.torch, env.dev
    c = SHARD_CONFIGS[cfg]
    n, ch, rate, frames = c["n"], c["ch"], c["rate"], c["frames"]
    x = make_input_device_chunked(torch, n, frames, 777 + env.rank, dev, chunk=4096 if ch <= 2 else 1024, channels=ch, rate=rate)
    mode = S.MODE_ALL if c["mode"] == "all" else S.MODE_LOUDNESS
    an = S.BatchAnalyzer(n, ch, rate, mode, device=env.local)
    if c["tp"]:
        an.force_true_peak_factor(c["tp"])
    gather = gather_factory(an) if gather_factory else None
    res = gather.local_rows() if gather is not None else torch.empty((n, an.stride), dtype=torch.float64, device=dev)

    def step():
        an.add_frames_results_device(x, res)
        if gather is not None:
            gather.publish()
            gather.wait()

    for i in range(warmup):
        step()
    env.barrier()
    an.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step()
    e1.record()
    env.barrier()
    ms = env.max_over_ranks(e0.elapsed_time(e1)) / steps
    kms, kn = an.profile_read()
    an.profile(False)
    b = n * frames * ch * 4
    out = {"what": c["what"], "samples_per_s_per_gpu": n * frames * ch / (ms * 1e-3), "step_ms": ms, "kernel_ms": kms / max(kn, 1),
           "algorithmic_gbs": b / (kms / max(kn, 1) * 1e-3) / 1e9, "frac_of_hbm_peak": b / (kms / max(kn, 1) * 1e-3) / 1e9 / hbm,
           "step_frac_of_hbm_peak": b / (ms * 1e-3) / 1e9 / hbm, "realtime_streams_equiv_per_gpu": n * frames / rate / (ms * 1e-3),
           "true_peak_factor": an.true_peak_factor(), "l2": f"{b / 1e9:.1f} GB resident input per GPU (> 126 MB L2)"}
    if gather is not None:
        gather.close()
    del an, x
    return out


def run_shard_main(args):
    """--config cfg4|cfg5|cfg5_x4 under torchrun: per-GPU shard + gather of the result rows; rank 0 prints one JSON line."""
    env = Env()
    import soundscope_b200 as S
    from soundscope_b200.sharding import PeerGather
    peaks, _ = measured_peaks()
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    clocks = ClockSampler(env.local)
    if env.rank == 0:
        clocks.start()
        clocks.mark_begin()
    r = run_shard_config(env, S, args.config, steps=args.steps, warmup=max(3, args.warmup), hbm=hbm,
                         gather_factory=(lambda an: PeerGather(an, env)) if env.world > 1 else None)
    if env.rank == 0:
        clocks.mark_end()
        c = SHARD_CONFIGS[args.config]
        line = {"metric": METRIC, "value": r["samples_per_s_per_gpu"] * env.world, "unit": UNIT, "n_gpus": env.world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": r["step_ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": c["what"], "streams_per_gpu": c["n"], "frames_per_launch": c["frames"], "channels": c["ch"], "rate": c["rate"],
                           "total_streams": c["n"] * env.world, "l2": r["l2"]},
                "clocks": clocks.stop(), "detail": r}
        print(json.dumps(line))
    env.barrier()
    env.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)   # 20 steps x 25 launches x ~0.16 ms = ~80 ms timed region per mode leg
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg2"] + sorted(SHARD_CONFIGS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the single-GPU extras (FFT, waveform, cfg4/cfg5 shards, ticks)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.config != "cfg2":
        run_shard_main(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
