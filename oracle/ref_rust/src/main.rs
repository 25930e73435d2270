//! gen_fixtures: golden vectors G1-G5 of SURVEY.md section 8c, produced by the reference's own `Analyzer`
//! (src/analyzer.rs, compiled from where it lies) on top of ebur128 0.1.10 / spectrum-analyzer 1.7.0.
//!
//! Output: <out>/ref_v1.bin (little-endian arrays back to back) + <out>/ref_v1.json (name -> dtype, shape, offset).
//! Inputs are generated here with closed forms (no RNG crates): the same generators exist in
//! soundscope_b200/synth.py and tests/test_ref_fixtures.py regenerates them to feed the oracle and the GPU path.
#![allow(dead_code)]

#[path = "/root/reference/src/analyzer.rs"]
mod analyzer;

use analyzer::Analyzer;
use std::f64::consts::PI;
use std::io::Write;

struct Out {
    bin: Vec<u8>,
    index: Vec<String>,
}

impl Out {
    fn f64s(&mut self, name: &str, shape: &[usize], v: &[f64]) {
        let off = self.bin.len();
        for x in v {
            self.bin.extend_from_slice(&x.to_le_bytes());
        }
        self.index.push(format!(
            "\"{}\": {{\"dtype\": \"f8\", \"shape\": {:?}, \"offset\": {}}}",
            name, shape, off
        ));
    }
    fn f32s(&mut self, name: &str, v: &[f32]) {
        let off = self.bin.len();
        for x in v {
            self.bin.extend_from_slice(&x.to_le_bytes());
        }
        self.index.push(format!(
            "\"{}\": {{\"dtype\": \"f4\", \"shape\": [{}], \"offset\": {}}}",
            name, v.len(), off
        ));
    }
    fn pairs(&mut self, name: &str, v: &[(f64, f64)]) {
        let flat: Vec<f64> = v.iter().flat_map(|p| [p.0, p.1]).collect();
        self.f64s(name, &[v.len(), 2], &flat);
    }
}

/// splitmix64 -> uniform in [-1, 1): the noise of G4 (mirrored bit for bit in tests/test_ref_fixtures.py)
fn noise(seed: u64, n: usize) -> Vec<f32> {
    let mut s = seed;
    (0..n)
        .map(|_| {
            s = s.wrapping_add(0x9E3779B97F4A7C15);
            let mut z = s;
            z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
            z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
            z ^= z >> 31;
            ((z >> 40) as f64 / (1u64 << 24) as f64 * 2.0 - 1.0) as f32
        })
        .collect()
}

/// the reference tests' sine (analyzer.rs:199-204): f32 arithmetic throughout
fn ref_sine(freq: f32, n: usize, rate: f32, amp: f32) -> Vec<f32> {
    (0..n)
        .map(|i| amp * (2.0 * std::f32::consts::PI * freq * (i as f32 / rate)).sin())
        .collect()
}

/// cfg1: log sweep 20 Hz -> 20 kHz, closed-form phase in f64, R = 0.5 L, interleaved
fn sweep_stereo(seconds: f64, rate: u32, amp: f64, side_gain: f64) -> Vec<f32> {
    let n = (seconds * rate as f64) as usize;
    let (f0, f1) = (20.0f64, 20000.0f64);
    let k = (f1 / f0).ln() / seconds;
    let mut x = Vec::with_capacity(2 * n);
    for i in 0..n {
        let t = i as f64 / rate as f64;
        let phase = 2.0 * PI * f0 * ((k * t).exp() - 1.0) / k;
        let l = amp * phase.sin();
        x.push(l as f32);
        x.push((side_gain * l) as f32);
    }
    x
}

/// the free function of src/audio_player.rs:400-419 (that file pulls in the playback crates, so it is restated here)
fn mid_side(samples: &[f32]) -> (Vec<f32>, Vec<f32>) {
    let left = samples.iter().step_by(2);
    let right = samples.iter().skip(1).step_by(2);
    let mid = left.clone().zip(right.clone()).map(|(l, r)| (l + r) / 2.).collect();
    let side = left.zip(right).map(|(l, r)| (l - r) / 2.).collect();
    (mid, side)
}

fn loudness_row(a: &mut Analyzer) -> Vec<f64> {
    let s = a.get_shortterm_lufs().unwrap_or(f64::NAN);
    let i = a.get_integrated_lufs().unwrap_or(f64::NAN);
    let r = a.get_loudness_range().unwrap_or(f64::NAN);
    let (l, rr) = a.get_true_peak().unwrap_or((f64::NAN, f64::NAN));
    vec![s, i, r, l, rr]
}

fn main() {
    let dir = std::env::args().nth(1).unwrap_or_else(|| "../_ref".to_string());
    std::fs::create_dir_all(&dir).unwrap();
    let mut o = Out { bin: Vec::new(), index: Vec::new() };

    // ---- G1: the six analyzer tests' own inputs (analyzer.rs:191-385) ----
    for (name, f) in [("g1_fft_440", 440.0f32), ("g1_fft_bin372", 372.0 * 44100.0 / 16384.0), ("g1_fft_125", 46.0 * 44100.0 / 16384.0)] {
        let a = Analyzer::default();
        let x = ref_sine(f, 16384, 44100.0, 1.0);
        o.f32s(&format!("{name}_in"), &x); // f32 sin is the platform libm's: the input travels with the output
        o.pairs(name, &a.get_fft(&x).unwrap());
    }
    {
        let s: Vec<f32> = (0..44100).map(|i| (i as f32 / 44100.0).sin()).collect();
        o.f32s("g1_waveform_in", &s);
        o.pairs("g1_waveform", &Analyzer::get_waveform(&s, 15.0));
        let mut a = Analyzer::default();
        let x: Vec<f32> = (0..88200).map(|i| 0.1 * (440.0 * 2.0 * std::f32::consts::PI * (i as f32 / 44100.0)).sin()).collect();
        a.add_samples(&x).unwrap();
        o.f32s("g1_loudness_in", &x);
        o.f64s("g1_loudness", &[5], &loudness_row(&mut a));
    }

    // ---- G2: cfg1 sweep, L = R / 2 and L = -R: per-hop spectra (every 2048 samples, N = 16384), ticks, one-shot ----
    for (tag, side_gain) in [("g2_sweep", 0.5f64), ("g2_sweep_anti", -1.0f64)] {
        let x = sweep_stereo(10.0, 48000, 0.5, side_gain);
        let (mid, side) = mid_side(&x);
        let mut a = Analyzer::default();
        a.create_loudness_meter(2, 48000).unwrap();
        let mut rows = Vec::new();
        let mut pos = 16384 + 2048;
        let mut hop = 0;
        while pos <= x.len() {
            a.add_samples(&x[pos - 16384..pos]).unwrap(); // tui.rs:1528-1543: overlapping windows, as the player feeds them
            rows.extend(loudness_row(&mut a));
            if hop % 32 == 0 && pos / 2 >= 16384 {
                let p = pos / 2;
                o.pairs(&format!("{tag}_mid_fft_{hop}"), &a.get_fft(&mid[p - 16384..p]).unwrap());
                o.pairs(&format!("{tag}_side_fft_{hop}"), &a.get_fft(&side[p - 16384..p]).unwrap());
            }
            pos += 2048;
            hop += 1;
        }
        o.f64s(&format!("{tag}_ticks"), &[rows.len() / 5, 5], &rows);
        o.f64s(&format!("{tag}_oneshot"), &[1], &[a.calculate_integrated_lufs(2, &x).unwrap_or(f64::NAN)]);
        o.pairs(&format!("{tag}_waveform"), &Analyzer::get_waveform(&x, 10.0));
    }

    // ---- G3: EBU Tech 3341 / 3342 style synthetic cases: tones at -23 / -33 dBFS, level steps, true-peak phase cases ----
    for (tag, rate, segs) in [
        ("g3_3341_1", 48000u32, vec![(-23.0f64, 20.0f64)]),
        ("g3_3341_2", 48000, vec![(-33.0, 20.0)]),
        ("g3_3341_3", 48000, vec![(-36.0, 10.0), (-23.0, 60.0), (-36.0, 10.0)]),
        ("g3_3341_4", 48000, vec![(-72.0, 10.0), (-36.0, 10.0), (-23.0, 60.0), (-36.0, 10.0), (-72.0, 10.0)]),
        ("g3_3342_1", 48000, vec![(-20.0, 20.0), (-30.0, 20.0)]),
        ("g3_3342_2", 48000, vec![(-20.0, 20.0), (-15.0, 20.0)]),
        ("g3_3342_3", 44100, vec![(-40.0, 20.0), (-20.0, 20.0)]),
        ("g3_3342_4", 96000, vec![(-50.0, 20.0), (-35.0, 20.0), (-20.0, 20.0), (-35.0, 20.0), (-50.0, 20.0)]),
    ] {
        let mut a = Analyzer::default();
        a.create_loudness_meter(2, rate).unwrap();
        let mut n0 = 0usize;
        for (db, secs) in segs {
            let amp = 10f64.powf(db / 20.0);
            let n = (secs * rate as f64) as usize;
            let x: Vec<f32> = (0..n)
                .flat_map(|i| {
                    let v = (amp * (2.0 * PI * 1000.0 * (n0 + i) as f64 / rate as f64).sin()) as f32;
                    [v, v]
                })
                .collect();
            for chunk in x.chunks(rate as usize * 2) {
                a.add_samples(chunk).unwrap();
            }
            n0 += n;
        }
        o.f64s(tag, &[5], &loudness_row(&mut a));
    }
    for (tag, rate, phase_deg) in [("g3_tp_fs4_0", 48000u32, 0.0f64), ("g3_tp_fs4_45", 48000, 45.0), ("g3_tp_fs4_45_96k", 96000, 45.0), ("g3_tp_fs4_45_192k", 192000, 45.0)] {
        let mut a = Analyzer::default();
        a.create_loudness_meter(2, rate).unwrap();
        let x: Vec<f32> = (0..rate as usize)
            .flat_map(|i| {
                let v = (0.5 * (2.0 * PI * 0.25 * i as f64 + phase_deg * PI / 180.0).sin()) as f32;
                [v, (0.5 * v as f64) as f32]
            })
            .collect();
        a.add_samples(&x).unwrap();
        o.f64s(tag, &[5], &loudness_row(&mut a));
    }

    // ---- G4: noise for the FFT error statistics, N = 2 .. 32768, and multichannel meters (channel map, weights) ----
    for lg in [1usize, 4, 9, 12, 13, 14, 15] {
        let n = 1usize << lg;
        let a = Analyzer::default();
        match a.get_fft(&noise(100 + lg as u64, n)) {
            Ok(v) => o.pairs(&format!("g4_fft_noise_{n}"), &v),
            Err(_) => o.f64s(&format!("g4_fft_noise_{n}_err"), &[1], &[1.0]),
        }
    }
    for (ch, rate) in [(1u32, 48000u32), (2, 44100), (4, 48000), (5, 48000), (6, 96000), (8, 48000)] {
        let mut a = Analyzer::default();
        a.create_loudness_meter(ch, rate).unwrap();
        let frames = rate as usize * 5;
        let base = noise(7 * ch as u64 + rate as u64, frames * ch as usize);
        let x: Vec<f32> = base.iter().enumerate().map(|(i, v)| 0.3 * v * (1.0 - 0.1 * (i % ch as usize) as f32)).collect();
        for chunk in x.chunks(rate as usize * ch as usize) {
            a.add_samples(chunk).unwrap();
        }
        let s = a.get_shortterm_lufs().unwrap_or(f64::NAN);
        let i = a.get_integrated_lufs().unwrap_or(f64::NAN);
        let r = a.get_loudness_range().unwrap_or(f64::NAN);
        let tp = a.get_true_peak().map(|p| vec![p.0, p.1]).unwrap_or(vec![f64::NAN, f64::NAN]);
        o.f64s(&format!("g4_meter_{ch}ch_{rate}"), &[5], &[s, i, r, tp[0], tp[1]]);
    }

    // ---- G5: block energies on the histogram bin edges.  find_histogram_index is private to the crate, so the edges are
    //      approached from outside: constant-level 1 kHz tones whose 400 ms block energy is stepped in 0.001 LU steps across
    //      bin boundaries; the integrated loudness jumps by one bin (0.1 LU) exactly where the crate's index flips. ----
    {
        let rate = 48000u32;
        let mut levels = Vec::new();
        let mut results = Vec::new();
        for step in 0..400 {
            let db = -23.2 + 0.001 * step as f64;
            let amp = 10f64.powf(db / 20.0);
            let mut a = Analyzer::default();
            a.create_loudness_meter(2, rate).unwrap();
            let x: Vec<f32> = (0..rate as usize * 3)
                .flat_map(|i| {
                    let v = (amp * (2.0 * PI * 1000.0 * i as f64 / rate as f64).sin()) as f32;
                    [v, v]
                })
                .collect();
            a.add_samples(&x).unwrap();
            levels.push(db);
            results.push(a.get_integrated_lufs().unwrap_or(f64::NAN));
        }
        o.f64s("g5_level_db", &[levels.len()], &levels);
        o.f64s("g5_integrated", &[results.len()], &results);
    }

    std::fs::File::create(format!("{dir}/ref_v1.bin")).unwrap().write_all(&o.bin).unwrap();
    let json = format!("{{\n{}\n}}\n", o.index.join(",\n"));
    std::fs::write(format!("{dir}/ref_v1.json"), json).unwrap();
    println!("wrote {} arrays, {} bytes to {dir}/ref_v1.bin", o.index.len(), o.bin.len());
}
