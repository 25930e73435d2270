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
__device__ __forceinline__ void results_for_stream(const GateParams& g, const ResultsArgs& ra, const double* __restrict__ energies,
                                                   const double* __restrict__ bounds, const size_t s, const int lane) {
  const int C = g.channels;
  const size_t stride = 4 + 2 * (size_t)C;
  double* o = ra.out + s * stride;
  const double NaN = __longlong_as_double(0x7ff8000000000000ll);
  const double NEG_INF = __longlong_as_double(0xfff0000000000000ll);
  const bool want_i = (ra.mode & SSB_MODE_I) == SSB_MODE_I, want_lra = (ra.mode & SSB_MODE_LRA) == SSB_MODE_LRA;
  const int bin0 = lane * 32;
  const bool pending = ra.gate_last >= ra.gate_first;
  // more pending buckets than lanes (only after a long unqueried feed): gate them the slow way first — atomics, fence,
  // then the histogram loads see them
  const bool one_round = !pending || ra.gate_last - ra.gate_first < 32;
  const double* bkp = ra.bucket + s * (size_t)C * kNB;
  if (!one_round) {
    for (uint64_t j = ra.gate_first + lane; j <= ra.gate_last; j += 32) {
      if (g.do_i && j >= 3) {
        const double e = window_energy(bkp, g, j, 4);
        if (e >= bounds[0]) atomicAdd(&ra.block_hist_rw[s * kHistBins + find_histogram_index(bounds, e)], 1u);
      }
      if (g.do_lra && j >= 29 && (j - 29) % 10 == 0) {
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
      if (want_lra && bin0 + 4 * q < kHistBins) w = __ldcg(hs4 + q);
      hi_[4 * q] = v.x; hi_[4 * q + 1] = v.y; hi_[4 * q + 2] = v.z; hi_[4 * q + 3] = v.w;
      hl_[4 * q] = w.x; hl_[4 * q + 1] = w.y; hl_[4 * q + 2] = w.z; hl_[4 * q + 3] = w.w;
    }
  }
  // --- gating of the pending buckets, one lane each: my_b / my_s = the bin this lane's bucket entered (-1: none) ---
  int my_b = -1, my_s = -1;
  if (pending && one_round) {
    const uint64_t j = ra.gate_first + lane;
    if (j <= ra.gate_last) {
      if (g.do_i && j >= 3) {
        const double e = window_energy(bkp, g, j, 4);
        if (e >= bounds[0]) {
          my_b = find_histogram_index(bounds, e);
          atomicAdd(&ra.block_hist_rw[s * kHistBins + my_b], 1u);
        }
      }
      if (g.do_lra && j >= 29 && (j - 29) % 10 == 0) {
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
  if (ra.ring_e) {
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
  if (lane == 0) {
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
    if (!cnt) integrated = NEG_INF;
    else {
      double rel = pw / (double)cnt;
      rel *= 0.1;  // 10^(-10/10)
      int start;
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
    }
  }
  // --- loudness range: ebur128 loudness_range, histogram branch (EBU Tech 3342) ---
  double lra = NaN;
  if (want_lra) {
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
    o[2] = integrated;
    o[3] = lra;
  }
  // --- peaks: EbuR128::true_peak = max(true_peak, sample_peak) ---
  for (int c = lane; c < C; c += 32) {
    const float spv = __ldcg(&ra.speak[s * C + c]), tpv = __ldcg(&ra.tpeak[s * C + c]);
    o[4 + c] = (double)fmaxf(spv, tpv);
    o[4 + C + c] = (double)spv;
  }
  // --- multi-GPU: the same row into block `rank` of every rank's gather buffer (peer stores over NVLink) ---
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

// Round-1 ordering of the same computation (gating with atomics, fence, then the histogram loads); kept for A/B timing
// of the standalone k_results kernel (SSB_RESULTS_V0=1).
//
// Latency is what this costs (every step is a dependent global round trip), so the loads that do not depend on
// anything are issued first: both 1000-bin histograms go to registers lane-major (lane l owns bins [32 l, 32 l + 32):
// eight 16-byte loads each), then the pending buckets are gated (one lane per bucket).  A gated block goes to the
// histogram in memory with a fire-and-forget atomic and to this call's sums directly (the lane that gated it adds its
// bin's energy to its partial sums; a new short-term entry is patched into the register copy), so nothing is read back
// and no fence is needed.
__device__ __forceinline__ void results_for_stream_v0(const GateParams& g, const ResultsArgs& ra, const size_t s, const int lane) {
  const int C = g.channels;
  const size_t stride = 4 + 2 * (size_t)C;
  double* o = ra.out + s * stride;
  const double NaN = __longlong_as_double(0x7ff8000000000000ll);
  const double NEG_INF = __longlong_as_double(0xfff0000000000000ll);
  const bool want_i = (ra.mode & SSB_MODE_I) == SSB_MODE_I, want_lra = (ra.mode & SSB_MODE_LRA) == SSB_MODE_LRA;
  const int bin0 = lane * 32;
  const bool pending = ra.gate_last >= ra.gate_first;
  // more pending buckets than lanes (only after a long unqueried feed): gate them the slow way first — atomics, fence,
  // then the histogram loads see them
  const bool one_round = !pending || ra.gate_last - ra.gate_first < 32;
  const double* bkp = ra.bucket + s * (size_t)C * kNB;
  if (!one_round) {
    for (uint64_t j = ra.gate_first + lane; j <= ra.gate_last; j += 32) {
      if (g.do_i && j >= 3) {
        const double e = window_energy(bkp, g, j, 4);
        if (e >= ra.bounds[0]) atomicAdd(&ra.block_hist_rw[s * kHistBins + find_histogram_index(ra.bounds, e)], 1u);
      }
      if (g.do_lra && j >= 29 && (j - 29) % 10 == 0) {
        const double e = window_energy(bkp, g, j, 30);
        if (e >= ra.bounds[0]) atomicAdd(&ra.st_hist_rw[s * kHistBins + find_histogram_index(ra.bounds, e)], 1u);
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
      if (want_lra && bin0 + 4 * q < kHistBins) w = __ldcg(hs4 + q);
      hi_[4 * q] = v.x; hi_[4 * q + 1] = v.y; hi_[4 * q + 2] = v.z; hi_[4 * q + 3] = v.w;
      hl_[4 * q] = w.x; hl_[4 * q + 1] = w.y; hl_[4 * q + 2] = w.z; hl_[4 * q + 3] = w.w;
    }
  }
  // --- gating of the pending buckets, one lane each: my_b / my_s = the bin this lane's bucket entered (-1: none) ---
  int my_b = -1, my_s = -1;
  if (pending && one_round) {
    const uint64_t j = ra.gate_first + lane;
    if (j <= ra.gate_last) {
      if (g.do_i && j >= 3) {
        const double e = window_energy(bkp, g, j, 4);
        if (e >= ra.bounds[0]) {
          my_b = find_histogram_index(ra.bounds, e);
          atomicAdd(&ra.block_hist_rw[s * kHistBins + my_b], 1u);
        }
      }
      if (g.do_lra && j >= 29 && (j - 29) % 10 == 0) {
        const double e = window_energy(bkp, g, j, 30);
        if (e >= ra.bounds[0]) {
          my_s = find_histogram_index(ra.bounds, e);
          atomicAdd(&ra.st_hist_rw[s * kHistBins + my_s], 1u);
        }
      }
    }
  }

  // --- momentary / short-term ---
  double e_m = NaN, e_s = NaN;
  if (ra.ring_e) {
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
  if (lane == 0) {
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
      if (hi_[t]) pw = fma((double)hi_[t], ra.energies[bin0 + t], pw);
      cnt += hi_[t];
    }
    const double my_e = my_b >= 0 ? ra.energies[my_b] : 0.0;   // the block this lane has just gated
    if (my_b >= 0) { pw += my_e; cnt += 1; }
    pw = warp_sum(pw);
    cnt = warp_sum_u64(cnt);
    if (!cnt) integrated = NEG_INF;
    else {
      double rel = pw / (double)cnt;
      rel *= 0.1;  // 10^(-10/10)
      int start;
      if (rel < ra.bounds[0]) start = 0;
      else {
        start = find_histogram_index(ra.bounds, rel);
        if (rel > ra.energies[start]) ++start;
      }
      double gp = 0.0;
      unsigned long long gc = 0;
#pragma unroll
      for (int t = 0; t < 32; t++) {
        if (hi_[t] && bin0 + t >= start) {
          gp = fma((double)hi_[t], ra.energies[bin0 + t], gp);
          gc += hi_[t];
        }
      }
      if (my_b >= start) { gp += my_e; gc += 1; }
      gp = warp_sum(gp);
      gc = warp_sum_u64(gc);
      integrated = gc ? energy_to_loudness(gp / (double)gc) : NEG_INF;
    }
  }
  // --- loudness range: ebur128 loudness_range, histogram branch (EBU Tech 3342) ---
  double lra = NaN;
  if (want_lra) {
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
      if (hl_[t]) pw = fma((double)hl_[t], ra.energies[bin0 + t], pw);
      cnt += hl_[t];
    }
    pw = warp_sum(pw);
    cnt = warp_sum_u64(cnt);
    if (!cnt) lra = 0.0;
    else {
      const double stl_integrated = 0.01 * (pw / (double)cnt);  // 10^(-20/10)
      int index;
      if (stl_integrated < ra.bounds[0]) index = 0;
      else {
        index = find_histogram_index(ra.bounds, stl_integrated);
        if (stl_integrated > ra.energies[index]) ++index;
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
        lra = energy_to_loudness(ra.energies[hi_bin]) - energy_to_loudness(ra.energies[lo_bin]);
      }
    }
  }
  if (lane == 0) {
    o[2] = integrated;
    o[3] = lra;
  }
  // --- peaks: EbuR128::true_peak = max(true_peak, sample_peak) ---
  for (int c = lane; c < C; c += 32) {
    const float spv = __ldcg(&ra.speak[s * C + c]), tpv = __ldcg(&ra.tpeak[s * C + c]);
    o[4 + c] = (double)fmaxf(spv, tpv);
    o[4 + C + c] = (double)spv;
  }
}


}  // namespace ssb
