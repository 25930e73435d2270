// loudness_results.cuh — gating and per-stream result scalars as device functions, shared by k_results (loudness.cu)
// and the fused epilogue of the batch filter kernel (loudness_wtile.cu).  One warp per stream.
// Reference: src/analyzer.rs:147-164 -> ebur128 loudness_momentary / loudness_shortterm / loudness_global /
// loudness_range / true_peak (crate ebur128 0.1.10 = libebur128, histogram mode).
#pragma once

#include "ssb_internal.cuh"

namespace ssb {

// ebur128 find_histogram_index: the bin i with boundaries[i] <= energy < boundaries[i+1] (clamped to 0..999).
// The crate bisects the 1001 boundaries; the same index is reached here from a closed-form guess
// (bins are 0.1 LU wide from -70 LUFS) corrected against the table, which replaces ten dependent loads by two.
__device__ __forceinline__ int find_histogram_index(const double* __restrict__ bounds, double energy) {
  int idx = (int)floor((10.0 * log10(energy) - 0.691 + 70.0) * 10.0);
  idx = idx < 0 ? 0 : (idx > 999 ? 999 : idx);
  while (idx > 0 && energy < bounds[idx]) --idx;
  while (idx < 999 && energy >= bounds[idx + 1]) ++idx;
  return idx;
}

// channel-weighted sum of NB buckets ending at bucket j (ebur128 calc_gating_block: per channel sum oldest to
// newest, surround channels x1.41, summed over channels in channel order).  All NB loads are issued before the
// first add: one memory latency per channel instead of NB.
template <int NB>
__device__ __forceinline__ double window_energy_t(const double* __restrict__ bk, const GateParams& g, uint64_t j) {
  double sum = 0.0;
  for (int c = 0; c < g.channels; c++) {
    const float w = g.weight[c];
    if (w == 0.0f) continue;
    double v[NB];
#pragma unroll
    for (int k = 0; k < NB; k++) v[k] = __ldcg(&bk[c * kNB + (int)((j - (uint64_t)(NB - 1 - k)) % kNB)]);
    double ch = 0.0;
#pragma unroll
    for (int k = 0; k < NB; k++) ch += v[k];
    if (w != 1.0f) ch *= 1.41;
    sum += ch;
  }
  return sum / (double)((uint64_t)NB * g.s100);
}
__device__ __forceinline__ double window_energy(const double* __restrict__ bk, const GateParams& g, uint64_t j, int nb) {
  return nb == 4 ? window_energy_t<4>(bk, g, j) : window_energy_t<30>(bk, g, j);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double energy_to_loudness(double e) { return 10.0 * log10(e) - 0.691; }

// mean square of the last `win_frames` frames of the ring ending at ring_pos (ebur128 calc_gating_block
// over the ring; the ring starts zeroed, so an under-filled window reads zeros exactly as the crate does)
__device__ __forceinline__ double ring_energy(const double* __restrict__ rg, const GateParams& g, size_t ring_frames,
                              size_t ring_pos, size_t win_frames, int lane) {
  double sum = 0.0;
  for (int c = 0; c < g.channels; c++) {
    const float w = g.weight[c];
    if (w == 0.0f) continue;
    double part = 0.0;
    for (size_t i = lane; i < win_frames; i += 32) {
      size_t idx = ring_pos + ring_frames - win_frames + i;
      if (idx >= ring_frames) idx -= ring_frames;
      const double y = rg[idx * g.channels + c];
      part = fma(y, y, part);
    }
    double ch = warp_sum(part);
    if (w != 1.0f) ch *= 1.41;
    sum += ch;
  }
  return sum / (double)win_frames;
}

// All 32 lanes of one warp call this for stream s: gates the pending buckets, then writes the stream's result row
// [momentary, shortterm, integrated, LRA, true_peak[C], sample_peak[C]].
//
// Latency is what this costs (every step is a dependent global round trip), so the loads that do not depend on
// anything are issued first: both 1000-bin histograms go to registers lane-major (lane l owns bins [32 l, 32 l + 32):
// eight 16-byte loads each), then the pending buckets are gated (one lane per bucket).  A gated block goes to the
// histogram in memory with a fire-and-forget atomic and to this call's sums directly (the lane that gated it adds its
// bin's energy to its partial sums; a new short-term entry is patched into the register copy), so nothing is read back
// and no fence is needed.
// `energies` / `bounds`: the 1000 bin-centre energies and 1001 bin boundaries — ra.energies / ra.bounds, or a copy of
// them in shared memory (the fused epilogue of k_loudness_wtile: with 225 KB of the SM's 228 KB carved out as shared
// memory there is no L1 left, and every table lookup of the gating / percentile code would be an L2 round trip).
//
// PARTS selects what the call does (R_ALL: everything).  The lean path (results_lean below) handles momentary /
// short-term / integrated / peaks itself and calls this with R_LRA | R_GATHER only on the launches that gate a 3 s
// entry: that gates the entry, scans the short-term histogram, writes o[3] and publishes the finished row.
constexpr int R_MS = 1, R_I = 2, R_LRA = 4, R_PEAKS = 8, R_GATHER = 16, R_ALL = 31;

template <int PARTS>
__device__ __forceinline__ void results_for_stream(const GateParams& g, const ResultsArgs& ra, const double* __restrict__ energies,
                                                   const double* __restrict__ bounds, const size_t s, const int lane) {
  const int C = g.channels;
  const size_t stride = 4 + 2 * (size_t)C;
  double* o = ra.out + s * stride;
  const double NaN = __longlong_as_double(0x7ff8000000000000ll);
  const double NEG_INF = __longlong_as_double(0xfff0000000000000ll);
  const bool want_i = (PARTS & R_I) && (ra.mode & SSB_MODE_I) == SSB_MODE_I;
  const bool want_lra = (PARTS & R_LRA) && (ra.mode & SSB_MODE_LRA) == SSB_MODE_LRA;
  const bool gate_i = (PARTS & R_I) && g.do_i, gate_lra = (PARTS & R_LRA) && g.do_lra;
  const bool scan_lra = want_lra && !ra.lra_from_cache;   // else: the range as of the last short-term entry is still valid
  const int bin0 = lane * 32;
  const bool pending = ra.gate_last >= ra.gate_first;
  // more pending buckets than lanes (only after a long unqueried feed): gate them the slow way first — atomics, fence,
  // then the histogram loads see them
  const bool one_round = !pending || ra.gate_last - ra.gate_first < 32;
  const double* bkp = ra.bucket + s * (size_t)C * kNB;
  if (!one_round) {
    for (uint64_t j = ra.gate_first + lane; j <= ra.gate_last; j += 32) {
      if (gate_i && j >= 3) {
        const double e = window_energy(bkp, g, j, 4);
        if (e >= bounds[0]) atomicAdd(&ra.block_hist_rw[s * kHistBins + find_histogram_index(bounds, e)], 1u);
      }
      if (gate_lra && j >= 29 && (j - 29) % 10 == 0) {
        const double e = window_energy(bkp, g, j, 30);
        if (e >= bounds[0]) atomicAdd(&ra.st_hist_rw[s * kHistBins + find_histogram_index(bounds, e)], 1u);
      }
    }
    __threadfence();
    __syncwarp();
  }
  // --- both histograms into registers (1000 = 31 * 32 + 8: whole quads only) ---
  uint32_t hi_[32], hl_[32];
  {
    const uint4* hb4 = reinterpret_cast<const uint4*>(ra.block_hist + s * kHistBins) + lane * 8;
    const uint4* hs4 = reinterpret_cast<const uint4*>(ra.st_hist + s * kHistBins) + lane * 8;
#pragma unroll
    for (int q = 0; q < 8; q++) {
      uint4 v = make_uint4(0, 0, 0, 0), w = make_uint4(0, 0, 0, 0);
      if (want_i && bin0 + 4 * q < kHistBins) v = __ldcg(hb4 + q);
      if (scan_lra && bin0 + 4 * q < kHistBins) w = __ldcg(hs4 + q);
      hi_[4 * q] = v.x; hi_[4 * q + 1] = v.y; hi_[4 * q + 2] = v.z; hi_[4 * q + 3] = v.w;
      hl_[4 * q] = w.x; hl_[4 * q + 1] = w.y; hl_[4 * q + 2] = w.z; hl_[4 * q + 3] = w.w;
    }
  }
  // --- gating of the pending buckets, one lane each: my_b / my_s = the bin this lane's bucket entered (-1: none) ---
  int my_b = -1, my_s = -1;
  if (pending && one_round) {
    const uint64_t j = ra.gate_first + lane;
    if (j <= ra.gate_last) {
      if (gate_i && j >= 3) {
        const double e = window_energy(bkp, g, j, 4);
        if (e >= bounds[0]) {
          my_b = find_histogram_index(bounds, e);
          atomicAdd(&ra.block_hist_rw[s * kHistBins + my_b], 1u);
        }
      }
      if (gate_lra && j >= 29 && (j - 29) % 10 == 0) {
        const double e = window_energy(bkp, g, j, 30);
        if (e >= bounds[0]) {
          my_s = find_histogram_index(bounds, e);
          atomicAdd(&ra.st_hist_rw[s * kHistBins + my_s], 1u);
        }
      }
    }
  }

  // --- momentary / short-term ---
  double e_m = NaN, e_s = NaN;
  if (!(PARTS & R_MS)) {
  } else if (ra.ring_e) {
    e_m = ra.ring_e[s * 2];
    if ((ra.mode & SSB_MODE_S) == SSB_MODE_S) e_s = ra.ring_e[s * 2 + 1];
  } else if (ra.ring) {
    const double* rg = ra.ring + s * ra.ring_frames * C;
    e_m = ring_energy(rg, g, ra.ring_frames, ra.ring_pos, (size_t)g.s100 * 4, lane);
    if ((ra.mode & SSB_MODE_S) == SSB_MODE_S) e_s = ring_energy(rg, g, ra.ring_frames, ra.ring_pos, (size_t)g.s100 * 30, lane);
  } else if (ra.aligned) {
    // buckets not yet produced since the last reset hold zeros, like the crate's zeroed ring
    const uint64_t j = ra.buckets_done + kNB - 1;  // last completed bucket, biased to stay non-negative mod kNB
    e_m = window_energy(bkp, g, j, 4);
    if ((ra.mode & SSB_MODE_S) == SSB_MODE_S) e_s = window_energy(bkp, g, j, 30);
  }
  if ((PARTS & R_MS) && lane == 0) {
    o[0] = (e_m == e_m) ? (e_m <= 0.0 ? NEG_INF : energy_to_loudness(e_m)) : NaN;
    o[1] = (e_s == e_s) ? (e_s <= 0.0 ? NEG_INF : energy_to_loudness(e_s)) : NaN;
  }

  // --- integrated: ebur128 gated_loudness, histogram branch ---
  double integrated = NaN;
  if (want_i) {
    double pw = 0.0;
    unsigned long long cnt = 0;
#pragma unroll
    for (int t = 0; t < 32; t++) {
      if (hi_[t]) pw = fma((double)hi_[t], energies[bin0 + t], pw);
      cnt += hi_[t];
    }
    const double my_e = my_b >= 0 ? energies[my_b] : 0.0;   // the block this lane has just gated
    if (my_b >= 0) { pw += my_e; cnt += 1; }
    pw = warp_sum(pw);
    cnt = warp_sum_u64(cnt);
    // the running sums the lean path continues from (StreamCache): rebuilt by every full scan
    int start = 0;
    double gp_all = 0.0;
    unsigned long long gc_all = 0;
    if (!cnt) integrated = NEG_INF;
    else {
      double rel = pw / (double)cnt;
      rel *= 0.1;  // 10^(-10/10)
      if (rel < bounds[0]) start = 0;
      else {
        start = find_histogram_index(bounds, rel);
        if (rel > energies[start]) ++start;
      }
      double gp = 0.0;
      unsigned long long gc = 0;
#pragma unroll
      for (int t = 0; t < 32; t++) {
        if (hi_[t] && bin0 + t >= start) {
          gp = fma((double)hi_[t], energies[bin0 + t], gp);
          gc += hi_[t];
        }
      }
      if (my_b >= start) { gp += my_e; gc += 1; }
      gp = warp_sum(gp);
      gc = warp_sum_u64(gc);
      integrated = gc ? energy_to_loudness(gp / (double)gc) : NEG_INF;
      gp_all = gp;
      gc_all = gc;
    }
    if (lane == 0) {
      StreamCache& sc = ra.cache[s];
      sc.n_all = cnt; sc.sum_all = pw; sc.n_above = gc_all; sc.sum_above = gp_all; sc.start = start;
    }
  }
  // --- loudness range: ebur128 loudness_range, histogram branch (EBU Tech 3342) ---
  double lra = NaN;
  if (want_lra && !scan_lra) lra = __ldcg(&ra.cache[s].lra);
  if (scan_lra) {
    // new short-term entries (at most one per second of audio) are patched into the register copy of their bin
    unsigned news = __ballot_sync(0xffffffffu, my_s >= 0);
    while (news) {
      const int src = __ffs(news) - 1;
      news &= news - 1;
      const int idx = __shfl_sync(0xffffffffu, my_s, src);
#pragma unroll
      for (int t = 0; t < 32; t++) hl_[t] += (bin0 + t == idx) ? 1u : 0u;
    }
    double pw = 0.0;
    unsigned long long cnt = 0;
#pragma unroll
    for (int t = 0; t < 32; t++) {
      if (hl_[t]) pw = fma((double)hl_[t], energies[bin0 + t], pw);
      cnt += hl_[t];
    }
    pw = warp_sum(pw);
    cnt = warp_sum_u64(cnt);
    if (lane == 0) {   // the running sums lra_scan_fast continues from
      ra.cache[s].n_st = cnt;
      ra.cache[s].sum_st = pw;
    }
    if (!cnt) lra = 0.0;
    else {
      const double stl_integrated = 0.01 * (pw / (double)cnt);  // 10^(-20/10)
      int index;
      if (stl_integrated < bounds[0]) index = 0;
      else {
        index = find_histogram_index(bounds, stl_integrated);
        if (stl_integrated > energies[index]) ++index;
      }
      // lane totals above the relative gate, their exclusive prefix, and the grand total
      unsigned long long mine = 0;
#pragma unroll
      for (int t = 0; t < 32; t++) if (bin0 + t >= index) mine += hl_[t];
      unsigned long long incl = mine;
#pragma unroll
      for (int o2 = 1; o2 < 32; o2 <<= 1) {
        const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o2);
        if (lane >= o2) incl += up;
      }
      const unsigned long long above = __shfl_sync(0xffffffffu, incl, 31);
      if (!above) lra = 0.0;
      else {
        const unsigned long long excl = incl - mine;
        const unsigned long long lo = (unsigned long long)((double)(above - 1) * 0.1 + 0.5);
        const unsigned long long hi = (unsigned long long)((double)(above - 1) * 0.95 + 0.5);
        // ebur128 walks `while (size <= p) size += hist[j++]` and takes bin j-1: the first bin whose running
        // count exceeds p.  The lane whose range (excl, incl] contains p+1 finds it in its registers.
        int lo_bin = -1, hi_bin = -1;
        unsigned long long run = excl;
#pragma unroll
        for (int t = 0; t < 32; t++) {
          if (bin0 + t >= index) {
            run += hl_[t];
            if (lo_bin < 0 && run > lo && excl <= lo) lo_bin = bin0 + t;
            if (hi_bin < 0 && run > hi && excl <= hi) hi_bin = bin0 + t;
          }
        }
        // exactly one lane found each (its excl <= p < incl); max-reduce the -1s away
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) {
          lo_bin = max(lo_bin, __shfl_xor_sync(0xffffffffu, lo_bin, o2));
          hi_bin = max(hi_bin, __shfl_xor_sync(0xffffffffu, hi_bin, o2));
        }
        lra = energy_to_loudness(energies[hi_bin]) - energy_to_loudness(energies[lo_bin]);
      }
    }
  }
  if (lane == 0) {
    if (PARTS & R_I) o[2] = integrated;
    if (PARTS & R_LRA) o[3] = lra;
    if (scan_lra) ra.cache[s].lra = lra;
  }
  // --- peaks: EbuR128::true_peak = max(true_peak, sample_peak) ---
  for (int c = lane; (PARTS & R_PEAKS) && c < C; c += 32) {
    const float spv = __ldcg(&ra.speak[s * C + c]), tpv = __ldcg(&ra.tpeak[s * C + c]);
    o[4 + c] = (double)fmaxf(spv, tpv);
    o[4 + C + c] = (double)spv;
  }
  // --- multi-GPU: the same row into block `rank` of every rank's gather buffer (peer stores over NVLink) ---
  if ((PARTS & R_GATHER) && ra.ga.world > 0) {
    __syncwarp();
    for (size_t i = lane; i < stride; i += 32) {
      const double v = __ldcg(&o[i]);
      for (int p = 0; p < ra.ga.world; p++) {
        double* dst = ra.ga.rows[p] + s * stride + i;
        if (dst != &o[i]) *dst = v;
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// The lean path: what a query costs when the meters are read after every few buckets (the analyzer's per-tick
// pattern, src/analyzer.rs:139-164; BASELINE config 2: 4 new buckets per launch).
//
// results_for_stream above is a chain of dependent global round trips on one warp per stream, two 4 KB histogram loads
// and ~8000 instructions of mostly cold code: 38 us per 4096-stream query on an otherwise idle GPU.  Here a stream gets
// 16 lanes (two streams per warp) and the query is restated so that nothing in it grows with the histogram:
//   1. one round of independent loads: the last 40 bucket sums of every channel (5 per lane in stereo, staged through
//      shared memory so that any lane can sum any window in the reference's order), the stream's StreamCache line, peaks;
//   2. lane roles: lane 0 the 3 s window, lane 1 the 400 ms window, lanes 2..11 one pending 400 ms block each —
//      oldest bucket first, channel by channel, exactly window_energy_t's additions;
//   3. one log10 per lane (loudness for lanes 0 / 1, the bin guess for the block lanes), one table round trip for
//      the bin's boundaries and energy;
//   4. integrated loudness from the cached sums: totals += new blocks; relative gate -> start bin; when the start bin
//      moved, the histogram bins between the old and the new one (typically none or one) are loaded and moved across;
//      new blocks at or above the gate are added; energy_to_loudness(sum_above / n_above);
//   5. the row is assembled across the 16 lanes and stored with one coalesced store (and, with a gather open, one per
//      rank); the block atomics go out last, after every load of this call has returned (__syncwarp orders them).
// Launches that gate a 3 s entry (once per second of audio) finish with results_for_stream<R_LRA | R_GATHER>.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double group_sum16(double v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long group_sum16_u64(unsigned long long v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// bin of `energy` starting from the guess `idx` with its table entries already loaded (the common case: no correction)
__device__ __forceinline__ int fix_histogram_index(const double* __restrict__ bounds, double energy, int idx, double b_lo, double b_hi) {
  if (energy >= b_lo && energy < b_hi) return idx;
  while (idx > 0 && energy < __ldg(&bounds[idx])) --idx;
  while (idx < 999 && energy >= __ldg(&bounds[idx + 1])) ++idx;
  return idx;
}
__device__ __forceinline__ int guess_histogram_index(double loudness) {
  const int idx = (int)floor((loudness + 70.0) * 10.0);
  return idx < 0 ? 0 : (idx > 999 ? 999 : idx);
}

// All 32 lanes call this; lanes 16 g .. 16 g + 15 serve stream s (valid: the group has a stream).  `stg`: this group's
// kLeanSlots * C doubles of shared memory.  finish_row: write o[3] from the cache and publish the row (false when an
// LRA scan follows).  Control flow is warp-uniform around every shuffle; loads and stores are predicated on `valid`.
struct LeanSt {        // what a lean call hands to lra_scan_fast (equal in the 16 lanes of a stream)
  int st_bin;          // bin of the 3 s entry this call gated (-1: none, or below the absolute gate)
  unsigned long long n_st;
  double sum_st;       // StreamCache::n_st / sum_st including that entry
};

__device__ __forceinline__ LeanSt results_lean(const GateParams& g, const ResultsArgs& ra, double* __restrict__ stg,
                                               const size_t s, const bool valid, const int lane, const bool finish_row) {
  const int C = g.channels;
  const int gl = lane & 15, gbase = lane & 16;
  const unsigned gmask = 0xffffu << gbase;
  const size_t stride = 4 + 2 * (size_t)C;
  const double NaN = __longlong_as_double(0x7ff8000000000000ll);
  const double NEG_INF = __longlong_as_double(0xfff0000000000000ll);
  const bool want_s = (ra.mode & SSB_MODE_S) == SSB_MODE_S, want_lra = (ra.mode & SSB_MODE_LRA) == SSB_MODE_LRA;
  const bool pending = ra.gate_last >= ra.gate_first;
  const uint64_t J = ra.buckets_done + kNB - 1;   // the last completed bucket, biased to stay non-negative (mod kNB)

  // ---- 1. every load that depends on nothing ----
  const double* bkp = ra.bucket + s * (size_t)C * kNB;
  const int nval = kLeanSlots * C;
  constexpr int NV = (kLeanSlots * kLeanMaxChannels + 15) / 16;
  double v[NV];
#pragma unroll
  for (int t = 0; t < NV; t++) {
    const int i = gl + 16 * t;
    const int c = i / kLeanSlots, m = i - c * kLeanSlots;
    v[t] = (valid && i < nval) ? __ldcg(&bkp[c * kNB + (int)((J - (uint64_t)(kLeanSlots - 1) + (uint64_t)m) % kNB)]) : 0.0;
  }
  unsigned long long cw = 0;
  if (valid && gl < 8) cw = __ldcg(reinterpret_cast<const unsigned long long*>(&ra.cache[s]) + gl);
  float spv = 0.f, tpv = 0.f;
  if (valid && gl < C) {
    spv = __ldcg(&ra.speak[s * C + gl]);
    tpv = __ldcg(&ra.tpeak[s * C + gl]);
  }
#pragma unroll
  for (int t = 0; t < NV; t++) {
    const int i = gl + 16 * t;
    if (i < nval) stg[i] = v[t];
  }
  __syncwarp();

  // ---- 2. one window per lane, in window_energy_t's order ----
  int len = 4, back = 0;   // window of `len` buckets ending `back` buckets before the last completed one
  bool act = false;
  if (gl == 0) { len = 30; act = ra.aligned && want_s; }
  else if (gl == 1) act = ra.aligned != 0;
  else if (gl < 2 + kLeanPending) {
    const uint64_t j = ra.gate_first + (uint64_t)(gl - 2);
    act = pending && j <= ra.gate_last && j >= 3 && g.do_i;
    back = act ? (int)(ra.buckets_done - 1 - j) : 0;
  } else if (gl == 2 + kLeanPending) {   // the 3 s entry among the pending buckets (lra_fast launches)
    len = 30;
    act = ra.lra_fast && ra.st_back >= 0 && g.do_lra;
    back = act ? ra.st_back : 0;
  }
  double e = 0.0;
  for (int c = 0; c < C; c++) {
    const float w = g.weight[c];
    if (w == 0.0f) continue;
    const double* p = stg + c * kLeanSlots + (kLeanSlots - back - len);
    double ch = 0.0;
#pragma unroll 1
    for (int k0 = 0; k0 < 30; k0 += 6) {
      if (k0 >= len) break;   // (lane 0 runs all five rounds; the warp follows it)
      double u[6];
#pragma unroll
      for (int k = 0; k < 6; k++) u[k] = (k0 + k < len) ? p[k0 + k] : 0.0;
#pragma unroll
      for (int k = 0; k < 6; k++) if (k0 + k < len) ch += u[k];
    }
    if (w != 1.0f) ch *= 1.41;
    e += ch;
  }
  e = e / (double)((uint64_t)len * g.s100);

  // ---- 3. loudness of my window; block lanes: histogram bin ----
  const double l = energy_to_loudness(e);
  const double lout = act ? (e <= 0.0 ? NEG_INF : l) : NaN;   // lanes 0 / 1: short-term / momentary
  const bool ok_bin = act && gl >= 2 && e >= ra.bound0;   // lanes that enter a histogram: blocks and the 3 s entry
  const bool ok_b = ok_bin && gl < 2 + kLeanPending;
  int idx = ok_bin ? guess_histogram_index(l) : 0;
  double en = 0.0;
  if (ok_bin) {
    const double b_lo = __ldg(&ra.bounds[idx]), b_hi = __ldg(&ra.bounds[idx + 1]);
    en = __ldg(&ra.energies[idx]);
    const int fixed = fix_histogram_index(ra.bounds, e, idx, b_lo, b_hi);
    if (fixed != idx) { idx = fixed; en = __ldg(&ra.energies[idx]); }
  }

  // ---- 4. integrated loudness from the running sums ----
  unsigned long long n_all = __shfl_sync(0xffffffffu, cw, gbase + 0);
  double sum_all = __longlong_as_double((long long)__shfl_sync(0xffffffffu, cw, gbase + 1));
  unsigned long long n_above = __shfl_sync(0xffffffffu, cw, gbase + 2);
  double sum_above = __longlong_as_double((long long)__shfl_sync(0xffffffffu, cw, gbase + 3));
  const int start = (int)(long long)__shfl_sync(0xffffffffu, cw, gbase + 4);
  const double lra_cached = __longlong_as_double((long long)__shfl_sync(0xffffffffu, cw, gbase + 5));
  LeanSt st;
  st.n_st = __shfl_sync(0xffffffffu, cw, gbase + 6);
  st.sum_st = __longlong_as_double((long long)__shfl_sync(0xffffffffu, cw, gbase + 7));
  {
    const bool st_ok = __shfl_sync(0xffffffffu, (int)ok_bin, gbase + 2 + kLeanPending) != 0;
    const int st_idx = __shfl_sync(0xffffffffu, idx, gbase + 2 + kLeanPending);
    const double st_en = __shfl_sync(0xffffffffu, en, gbase + 2 + kLeanPending);
    st.st_bin = st_ok ? st_idx : -1;
    if (st_ok) { st.n_st += 1; st.sum_st += st_en; }
  }
  n_all += (unsigned long long)__popc(__ballot_sync(0xffffffffu, ok_b) & gmask);
  sum_all += group_sum16(ok_b ? en : 0.0);
  int start_new = start;
  if (n_all) {
    double rel = sum_all / (double)n_all;
    rel *= 0.1;  // 10^(-10/10)
    if (rel < ra.bound0) start_new = 0;
    else {
      int i2 = guess_histogram_index(energy_to_loudness(rel));
      const double b_lo = __ldg(&ra.bounds[i2]), b_hi = __ldg(&ra.bounds[i2 + 1]);
      i2 = fix_histogram_index(ra.bounds, rel, i2, b_lo, b_hi);
      if (rel > __ldg(&ra.energies[i2])) ++i2;
      start_new = i2;
    }
  }
  {
    // the bins the gate moved across change sides (read before this call's atomics go out)
    const int lo = min(start, start_new), hi = min(max(start, start_new), kHistBins);
    const int span = hi - lo;
    const int span_w = max(span, __shfl_xor_sync(0xffffffffu, span, 16));
    double ds = 0.0;
    unsigned long long dn = 0;
    for (int b0 = 0; b0 < span_w; b0 += 16) {
      const int b = lo + b0 + gl;
      if (valid && b0 + gl < span) {
        const uint32_t hcount = __ldcg(&ra.block_hist[s * kHistBins + b]);
        if (hcount) {
          ds = fma((double)hcount, __ldg(&ra.energies[b]), ds);
          dn += hcount;
        }
      }
    }
    if (span_w) {
      ds = group_sum16(ds);
      dn = group_sum16_u64(dn);
      if (start_new > start) { n_above -= dn; sum_above -= ds; }
      else { n_above += dn; sum_above += ds; }
    }
  }
  {
    const bool ok_a = ok_b && idx >= start_new;
    n_above += (unsigned long long)__popc(__ballot_sync(0xffffffffu, ok_a) & gmask);
    sum_above += group_sum16(ok_a ? en : 0.0);
  }
  if (n_above == 0) sum_above = 0.0;   // nothing above the gate: drop the rounding residue of the moves
  double integrated = NaN;
  if ((ra.mode & SSB_MODE_I) == SSB_MODE_I) integrated = (n_all && n_above) ? energy_to_loudness(sum_above / (double)n_above) : NEG_INF;

  // ---- 5. the row, one value per lane: [M, S, I, LRA, max(true, sample) peak[C], sample peak[C]] ----
  const double l_m = __shfl_sync(0xffffffffu, lout, gbase + 1), l_s = __shfl_sync(0xffffffffu, lout, gbase + 0);
  const float pk = fmaxf(spv, tpv);
  const float r_pk = __shfl_sync(0xffffffffu, pk, gbase + ((gl - 4) & 15));
  const float r_sp = __shfl_sync(0xffffffffu, spv, gbase + ((gl - 4 - C) & 15));
  const double rowv = gl == 0 ? l_m : gl == 1 ? l_s : gl == 2 ? integrated : gl == 3 ? (want_lra ? lra_cached : NaN)
                      : gl < 4 + C ? (double)r_pk : (double)r_sp;
  if (valid && gl < (int)stride && (finish_row || gl != 3)) {
    double* o = ra.out + s * stride;
    o[gl] = rowv;
    if (finish_row) {
      for (int p = 0; p < ra.ga.world; p++) {
        double* dst = ra.ga.rows[p] + s * stride + gl;
        if (dst != &o[gl]) *dst = rowv;
      }
    }
  }
  if (valid && (gl < 5 || ((gl == 6 || gl == 7) && st.st_bin >= 0))) {
    const unsigned long long wv = gl == 0 ? n_all : gl == 1 ? (unsigned long long)__double_as_longlong(sum_all)
                                  : gl == 2 ? n_above : gl == 3 ? (unsigned long long)__double_as_longlong(sum_above)
                                  : gl == 4 ? (unsigned long long)(long long)start_new
                                  : gl == 6 ? st.n_st : (unsigned long long)__double_as_longlong(st.sum_st);
    reinterpret_cast<unsigned long long*>(&ra.cache[s])[gl] = wv;
  }
  __syncwarp();
  if (valid && ok_b) atomicAdd(&ra.block_hist_rw[s * kHistBins + idx], 1u);
  return st;
}

// The loudness range of stream s on a launch whose 3 s entry the lean code has gated (ebur128 loudness_range, histogram
// branch, EBU Tech 3342): one warp per stream.  The first pass over the histogram (count and energy sum of all entries)
// comes from the cache, so the chain is: histogram loads -> relative gate's bin (one table round trip) -> prefix scan and
// percentile walk in registers -> the two bin energies (one round trip).  Writes o[3], the cache, and publishes the row.
__device__ __forceinline__ void lra_scan_fast(const ResultsArgs& ra, const size_t s, const int lane, const int st_bin,
                                              const unsigned long long n_st, const double sum_st, const size_t stride) {
  const int bin0 = lane * 32;
  uint32_t hl_[32];
  {
    const uint4* hs4 = reinterpret_cast<const uint4*>(ra.st_hist + s * kHistBins) + lane * 8;
#pragma unroll
    for (int q = 0; q < 8; q++) {
      uint4 w = make_uint4(0, 0, 0, 0);
      if (bin0 + 4 * q < kHistBins) w = __ldcg(hs4 + q);
      hl_[4 * q] = w.x; hl_[4 * q + 1] = w.y; hl_[4 * q + 2] = w.z; hl_[4 * q + 3] = w.w;
    }
  }
  double lra = 0.0;
  if (n_st) {
    const double stl_integrated = 0.01 * (sum_st / (double)n_st);  // 10^(-20/10)
    int index = 0;
    if (!(stl_integrated < ra.bound0)) {
      index = guess_histogram_index(energy_to_loudness(stl_integrated));
      const double b_lo = __ldg(&ra.bounds[index]), b_hi = __ldg(&ra.bounds[index + 1]);
      index = fix_histogram_index(ra.bounds, stl_integrated, index, b_lo, b_hi);
      if (stl_integrated > __ldg(&ra.energies[index])) ++index;
    }
    // the new entry goes into the register copy of its bin (the atomic below is not read back)
#pragma unroll
    for (int t = 0; t < 32; t++) hl_[t] += (bin0 + t == st_bin) ? 1u : 0u;
    unsigned long long mine = 0;
#pragma unroll
    for (int t = 0; t < 32; t++) if (bin0 + t >= index) mine += hl_[t];
    unsigned long long incl = mine;
#pragma unroll
    for (int o2 = 1; o2 < 32; o2 <<= 1) {
      const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o2);
      if (lane >= o2) incl += up;
    }
    const unsigned long long above = __shfl_sync(0xffffffffu, incl, 31);
    if (above) {
      const unsigned long long excl = incl - mine;
      const unsigned long long lo = (unsigned long long)((double)(above - 1) * 0.1 + 0.5);
      const unsigned long long hi = (unsigned long long)((double)(above - 1) * 0.95 + 0.5);
      int lo_bin = -1, hi_bin = -1;
      unsigned long long run = excl;
#pragma unroll
      for (int t = 0; t < 32; t++) {
        if (bin0 + t >= index) {
          run += hl_[t];
          if (lo_bin < 0 && run > lo && excl <= lo) lo_bin = bin0 + t;
          if (hi_bin < 0 && run > hi && excl <= hi) hi_bin = bin0 + t;
        }
      }
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) {
        lo_bin = max(lo_bin, __shfl_xor_sync(0xffffffffu, lo_bin, o2));
        hi_bin = max(hi_bin, __shfl_xor_sync(0xffffffffu, hi_bin, o2));
      }
      lra = energy_to_loudness(__ldg(&ra.energies[hi_bin])) - energy_to_loudness(__ldg(&ra.energies[lo_bin]));
    }
  }
  __syncwarp();   // every lane's histogram loads have returned
  double* o = ra.out + s * stride;
  if (lane == 0) {
    if (st_bin >= 0) atomicAdd(&ra.st_hist_rw[s * kHistBins + st_bin], 1u);
    o[3] = lra;
    ra.cache[s].lra = lra;
  }
  if (ra.ga.world > 0) {
    __syncwarp();
    for (size_t i = lane; i < stride; i += 32) {
      const double v = __ldcg(&o[i]);
      for (int p = 0; p < ra.ga.world; p++) {
        double* dst = ra.ga.rows[p] + s * stride + i;
        if (dst != &o[i]) *dst = v;
      }
    }
  }
}

// The rows of streams s0 and s0 + 1 (valid0 / valid1) by one warp: the lean pass for both at once, then — on launches
// that gate a 3 s entry — each stream's loudness range.  `stg`: 2 * kLeanSlots * C doubles of shared memory.
__device__ __forceinline__ void results_lean_pair(const GateParams& g, const ResultsArgs& ra, double* __restrict__ stg,
                                                  const size_t s0, const bool valid0, const bool valid1, const int lane) {
  const bool lra_scan = (ra.mode & SSB_MODE_LRA) == SSB_MODE_LRA && !ra.lra_from_cache;
  const int grp = lane >> 4;
  const bool valid = grp ? valid1 : valid0;
  const LeanSt st = results_lean(g, ra, stg + grp * (kLeanSlots * g.channels), valid ? s0 + grp : s0, valid, lane, !lra_scan);
  if (lra_scan) {
    const size_t stride = 4 + 2 * (size_t)g.channels;
    for (int q = 0; q < 2; q++) {
      const int st_bin = __shfl_sync(0xffffffffu, st.st_bin, q * 16);
      const unsigned long long n_st = __shfl_sync(0xffffffffu, st.n_st, q * 16);
      const double sum_st = __shfl_sync(0xffffffffu, st.sum_st, q * 16);
      if (q ? valid1 : valid0) {
        if (ra.lra_fast) lra_scan_fast(ra, s0 + q, lane, st_bin, n_st, sum_st, stride);
        else results_for_stream<R_LRA | R_GATHER>(g, ra, ra.energies, ra.bounds, s0 + q, lane);
      }
    }
  }
}

}  // namespace ssb
