"""CPU restatement (numpy) of the data formats either side of the analyzer path — TEST INFRASTRUCTURE ONLY
(see oracle/oracle.h): only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.

* `pcm_to_f32`     — reference src/audio_player.rs:169-267 (`decode_file`): for WAV / AIFF input symphonia's PCM
                     decoder yields the container's samples and `SampleBuffer::<f32>::copy_interleaved_ref`
                     (audio_player.rs:248) converts them to interleaved f32.
* `RingRef`        — `RBuffer` (src/tui.rs:37), created and zero-filled at main.rs:63-65 / tui.rs:1783-1786, written by
                     the cpal callback (src/audio_capture.rs:40-52).
* `mic_tick`       — `analyze_microphone_input` (src/tui.rs:1427-1480) on one ring snapshot.

The conversion rules live in symphonia-core 0.5.5 (`conv.rs`, `impl FromSample<S> for f32`; Cargo.lock:2085-2087), the
ring in ringbuffer's `AllocRingBuffer` — both un-vendored third-party crates that cannot be built here (no Rust
toolchain): **parity unpinned against the reference binary**.  The rules are restated from the published sources
([UPSTREAM-RECALL]): every integer format is normalised by 2^(bits-1) after re-centring unsigned input, the 32-bit
format goes through f64.  tests/test_oracle_kat.py checks the two plausible spellings of each rule
(`(s - 128) / 128` vs `s / 128 - 1`, ...) agree bit for bit, so the recall cannot be wrong in a way that matters.
"""
from collections import deque

import numpy as np

# format name -> (numpy dtype of one sample as stored, bytes per sample)
_LAYOUT = {
    "u8": ("u1", 1), "s8": ("i1", 1), "s16le": ("<i2", 2), "s16be": (">i2", 2), "s24le": (None, 3), "s24be": (None, 3),
    "s32le": ("<i4", 4), "s32be": (">i4", 4), "f32le": ("<u4", 4), "f32be": (">u4", 4), "f64le": ("<f8", 8), "f64be": (">f8", 8),
}


def pcm_bytes_per_sample(fmt):
    return _LAYOUT[fmt][1]


def _s24(raw, big):
    b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
    v = (b[:, 0] << 16 | b[:, 1] << 8 | b[:, 2]) if big else (b[:, 2] << 16 | b[:, 1] << 8 | b[:, 0])
    return np.where(v & 0x800000, v - (1 << 24), v).astype(np.int32)  # i24: sign-extend (read_i24 / read_be_i24)


def pcm_to_f32(raw, fmt):
    """symphonia-core conv.rs, `FromSample<S> for f32`, applied to every stored sample in order."""
    raw = bytes(raw)
    dt, bps = _LAYOUT[fmt]
    raw = raw[: len(raw) // bps * bps]
    f32 = np.float32
    if fmt == "u8":      # u8 -> i8 by re-centring, then i8 rule
        s = np.frombuffer(raw, dtype=dt).astype(np.int32) - 128
        return s.astype(f32) / f32(128.0)
    if fmt == "s8":      # f32::from(s) / 128.0
        return np.frombuffer(raw, dtype=dt).astype(f32) / f32(128.0)
    if fmt in ("s16le", "s16be"):   # f32::from(s) / 32_768.0
        return np.frombuffer(raw, dtype=dt).astype(f32) / f32(32768.0)
    if fmt in ("s24le", "s24be"):   # (s.clamped().inner() as f32) / 8_388_608.0
        return _s24(raw, fmt.endswith("be")).astype(f32) / f32(8388608.0)
    if fmt in ("s32le", "s32be"):   # (f64::from(s) / 2_147_483_648.0) as f32
        return (np.frombuffer(raw, dtype=dt).astype(np.float64) / 2147483648.0).astype(f32)
    if fmt in ("f32le", "f32be"):   # identity (bit pattern kept)
        return np.frombuffer(raw, dtype=dt).astype(np.uint32).view(f32)
    if fmt in ("f64le", "f64be"):   # s as f32
        with np.errstate(over="ignore"):
            return np.frombuffer(raw, dtype=dt).astype(np.float64).astype(f32)
    raise KeyError(fmt)


def pcm_to_f32_alt(raw, fmt):
    """The other plausible spelling of each unsigned / integer rule (`s / 2^(b-1) - 1`, multiply by the reciprocal):
    used only to show the recalled rules are insensitive to the spelling."""
    raw = bytes(raw)
    f32 = np.float32
    if fmt == "u8":
        return np.frombuffer(raw, dtype="u1").astype(f32) / f32(128.0) - f32(1.0)
    if fmt in ("s16le", "s16be"):
        return np.frombuffer(raw, dtype=_LAYOUT[fmt][0]).astype(f32) * f32(1.0 / 32768.0)
    if fmt in ("s24le", "s24be"):
        return (_s24(raw, fmt.endswith("be")).astype(np.float64) / 8388608.0).astype(f32)
    if fmt in ("s32le", "s32be"):
        return np.frombuffer(raw, dtype=_LAYOUT[fmt][0]).astype(f32) / f32(2147483648.0)
    return pcm_to_f32(raw, fmt)


class RingRef:
    """AllocRingBuffer::<f32>::new(capacity) + fill(0.0): a FIFO of the last `capacity` values."""

    def __init__(self, capacity):
        self.buf = deque([np.float32(0.0)] * capacity, maxlen=capacity)  # buf.fill(0.0)

    def callback(self, data, is_mono):
        """audio_capture.rs:40-52, line by line."""
        data = np.asarray(data, dtype=np.float32)
        if is_mono:
            out = []
            for i, x in enumerate(data):          # .enumerate().flat_map(|(i, &x)| if i == 0 { vec![x] } else { vec![0., x] })
                out.extend([x] if i == 0 else [np.float32(0.0), x])
            self.buf.extend(out)                  # audio_buf.extend(data)
        else:
            self.buf.extend(data.tolist())        # audio_buf.extend(data.iter().copied())

    def to_vec(self):
        return np.array(self.buf, dtype=np.float32)


def mic_tick(samples, analyzer, n_fft=2 ** 14, lufs_samples=2 ** 14, waveform_window=15.0):
    """tui.rs:1427-1480 with `samples` = latest_captured_samples.to_vec() and `analyzer` an oracle `Analyzer`
    (oracle/binding.py).  Returns (mid_fft | None, side_fft | None, waveform, shortterm, lufs_error | None)."""
    from . import binding as B
    mid, side = B.mid_side(samples)                               # :1429
    sr = analyzer.sample_rate()                                   # :1430
    left = 15 * sr - n_fft                                        # :1431
    try:
        mid_fft = analyzer.get_fft(mid[left:15 * sr])             # :1434-1443
    except B.OracleError:
        mid_fft = None
    try:
        side_fft = analyzer.get_fft(side[left:15 * sr])           # :1444-1453
    except B.OracleError:
        side_fft = None
    wave = B.get_waveform(mid, waveform_window)                   # :1456
    lb = 30 * sr - lufs_samples                                   # :1465
    err = None
    try:
        analyzer.add_samples(samples[lb:30 * sr])                 # :1466-1471
    except B.OracleError as e:
        err = e
    st = analyzer.get_shortterm_lufs()                            # :1472-1478
    return mid_fft, side_fft, wave, st, err
