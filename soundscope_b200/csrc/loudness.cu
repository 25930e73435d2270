// loudness.cu — BS.1770 / EBU R128 meter kernels (generic path), gating and result kernels.
//
// Replaces, for n_streams independent meters, what `EbuR128::add_frames_f32` and the loudness
// queries do for one (reference src/analyzer.rs:139-164; crate ebur128 0.1.10 = libebur128):
//   k_loudness_generic  one thread per (stream, channel): sample peak, polyphase true peak (f32),
//                       4th-order DF-II K-weighting in f64, y^2 accumulated into per-100 ms buckets,
//                       optional ring of y (SSB_FLAG_RING)
//   k_gating            per stream: 400 ms block energies -> block histogram; 3 s energies -> short-term histogram
//   k_results           per stream (one warp): momentary, short-term, integrated (histogram gating), LRA, peaks
//
// State layout in HBM (all stream-major):
//   filt[n][C][4] f64 | bucket[n][C][64] f64 | block_hist[n][1000] u32 | st_hist[n][1000] u32
//   speak/tpeak[n][C] f32 | tphist[n][C][24] f32 | ring[n][ring_frames][C] f64 (optional)
#include <stdlib.h>

#include "loudness_results.cuh"

namespace ssb {

// ------------------------------------------------------------------------------------------------
// generic filter kernel: fully general in channels / rate / chunking; thread per (stream, channel)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_loudness_generic(const LoudParams p, const float* __restrict__ in, size_t frames, size_t in_stride_frames,
                   size_t n_streams, double* __restrict__ filt, double* __restrict__ bucket,
                   float* __restrict__ speak, float* __restrict__ tpeak, float* __restrict__ tphist,
                   double* __restrict__ ring, size_t ring_frames, size_t ring_pos, uint32_t pos0,
                   uint32_t slot0) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int C = p.channels;
  if (idx >= n_streams * (size_t)C) return;
  const size_t stream = idx / C;
  const int c = (int)(idx % C);
  const float* x_ptr = in + stream * in_stride_frames * C + c;
  const bool active = p.do_filter && ((p.active_mask >> c) & 1ull);

  double v1 = 0, v2 = 0, v3 = 0, v4 = 0;
  if (active) {
    const double* f = filt + idx * 4;
    v1 = f[0]; v2 = f[1]; v3 = f[2]; v4 = f[3];
  }
  double* bk = bucket + idx * kNB;
  uint32_t slot = slot0, pos = pos0;
  double acc = (active && pos0 > 0) ? bk[slot] : 0.0;

  float h[kTpHist];  // h[t] = x[n-1-t]
  float sp = 0.f, tp = 0.f;
  const bool do_tp = p.do_true_peak && p.tp_factor != 0;
  if (do_tp) {
#pragma unroll
    for (int t = 0; t < kTpHist; t++) h[t] = tphist[idx * kTpHist + t];
  }
  double* rg = ring ? ring + stream * ring_frames * C + c : nullptr;
  size_t rpos = ring_pos;

  for (size_t i = 0; i < frames; i++) {
    const float xf = x_ptr[i * C];
    if (p.do_sample_peak) sp = fmaxf(sp, fabsf(xf));
    if (do_tp) {
      if (p.tp_factor == 4) {
#pragma unroll
        for (int f = 0; f < 3; f++) {
          float a = xf * p.tp4[f][0];
#pragma unroll
          for (int t = 1; t < 12; t++) a = fmaf(h[t - 1], p.tp4[f][t], a);
          tp = fmaxf(tp, fabsf(a));
        }
      } else {
        float a = xf * p.tp2[0];
#pragma unroll
        for (int t = 1; t < 24; t++) a = fmaf(h[t - 1], p.tp2[t], a);
        tp = fmaxf(tp, fabsf(a));
      }
#pragma unroll
      for (int t = kTpHist - 1; t > 0; t--) h[t] = h[t - 1];
      h[0] = xf;
    }
    if (active) {
      // v0 = x - a1 v1 - a2 v2 - a3 v3 - a4 v4, the newest state last to keep the dependent chain short
      double t = fma(p.na[4], v4, (double)xf);
      t = fma(p.na[3], v3, t);
      t = fma(p.na[2], v2, t);
      const double v0 = fma(p.na[1], v1, t);
      double y = p.b[4] * v4;
      y = fma(p.b[3], v3, y);
      y = fma(p.b[2], v2, y);
      y = fma(p.b[1], v1, y);
      y = fma(p.b[0], v0, y);
      v4 = v3; v3 = v2; v2 = v1; v1 = v0;
      acc = fma(y, y, acc);
      if (rg) {
        rg[rpos * C] = y;
        if (++rpos == ring_frames) rpos = 0;
      }
    } else if (rg) {
      rg[rpos * C] = 0.0;
      if (++rpos == ring_frames) rpos = 0;
    }
    if (++pos == p.s100) {
      bk[slot] = acc;
      acc = 0.0;
      pos = 0;
      slot = (slot + 1) % kNB;
    }
  }
  bk[slot] = acc;  // running sum of the (possibly empty) bucket in progress
  if (active) {
    // libebur128 flushes denormal filter state at the end of every call
    double* f = filt + idx * 4;
    f[0] = fabs(v1) < 2.2250738585072014e-308 ? 0.0 : v1;
    f[1] = fabs(v2) < 2.2250738585072014e-308 ? 0.0 : v2;
    f[2] = fabs(v3) < 2.2250738585072014e-308 ? 0.0 : v3;
    f[3] = fabs(v4) < 2.2250738585072014e-308 ? 0.0 : v4;
  }
  if (p.do_sample_peak) speak[idx] = fmaxf(speak[idx], sp);
  if (do_tp) {
    tpeak[idx] = fmaxf(tpeak[idx], tp);
#pragma unroll
    for (int t = 0; t < kTpHist; t++) tphist[idx * kTpHist + t] = h[t];
  }
}

cudaError_t launch_loudness_generic(const LoudParams& p, const LoudState& st, const float* d_in,
                                    size_t frames, size_t in_stride_frames, uint32_t pos0,
                                    uint64_t bucket0, size_t ring_pos, cudaStream_t s, uint64_t* launches) {
  const size_t chains = st.n_streams * (size_t)p.channels;
  if (!chains || !frames) return cudaSuccess;
  const int tpb = 128;
  const unsigned blocks = (unsigned)((chains + tpb - 1) / tpb);
  k_loudness_generic<<<blocks, tpb, 0, s>>>(p, d_in, frames, in_stride_frames, st.n_streams, st.filt,
                                            st.bucket, st.speak, st.tpeak, st.tphist, st.ring,
                                            st.ring_frames, ring_pos, pos0, (uint32_t)(bucket0 % kNB));
  if (launches) ++*launches;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// gating: block / short-term energies of the buckets completed by the last filter launch
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_gating(const GateParams g, size_t n_streams, const double* __restrict__ bucket,
         uint32_t* __restrict__ block_hist, uint32_t* __restrict__ st_hist,
         const double* __restrict__ bounds, uint64_t j_first, uint64_t j_last) {
  const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_streams) return;
  const double* bk = bucket + s * (size_t)g.channels * kNB;
  for (uint64_t j = j_first; j <= j_last; j++) {
    if (g.do_i && j >= 3) {
      const double e = window_energy(bk, g, j, 4);
      if (e >= bounds[0]) block_hist[s * kHistBins + find_histogram_index(bounds, e)]++;
    }
    if (g.do_lra && j >= 29 && (j - 29) % 10 == 0) {
      const double e = window_energy(bk, g, j, 30);
      if (e >= bounds[0]) st_hist[s * kHistBins + find_histogram_index(bounds, e)]++;
    }
  }
}

cudaError_t launch_gating(const GateParams& g, const LoudState& st, uint64_t j_first, uint64_t j_last,
                          cudaStream_t s, uint64_t* launches) {
  if (!st.n_streams || j_last < j_first || !(g.do_i || g.do_lra)) return cudaSuccess;
  const int tpb = 128;
  k_gating<<<(unsigned)((st.n_streams + tpb - 1) / tpb), tpb, 0, s>>>(g, st.n_streams, st.bucket, st.block_hist,
                                                                      st.st_hist, st.hist_boundaries, j_first, j_last);
  if (launches) ++*launches;
  return cudaGetLastError();
}

// Whole-file gating: one thread per 100 ms hop of stream 0, bucket sums in a linear array (no ring).  Same summation
// order as window_energy_t (oldest bucket first, channel by channel), so a block energy is the number k_gating computes.
template <int NB>
__device__ __forceinline__ double file_window_energy(const double* __restrict__ fb, size_t stride, const GateParams& g,
                                                     uint64_t j) {
  double sum = 0.0;
  for (int c = 0; c < g.channels; c++) {
    const float w = g.weight[c];
    if (w == 0.0f) continue;
    const double* b = fb + (size_t)c * stride + (j - (uint64_t)(NB - 1));
    double ch = 0.0;
#pragma unroll
    for (int k = 0; k < NB; k++) ch += b[k];
    if (w != 1.0f) ch *= 1.41;
    sum += ch;
  }
  return sum / (double)((uint64_t)NB * g.s100);
}

__global__ void __launch_bounds__(128)
k_file_gating(const GateParams g, const double* __restrict__ fb, size_t stride, uint64_t n_buckets,
              const double* __restrict__ bounds, uint32_t* __restrict__ block_hist, uint32_t* __restrict__ st_hist) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_buckets) return;
  if (g.do_i && j >= 3) {
    const double e = file_window_energy<4>(fb, stride, g, j);
    if (e >= bounds[0]) atomicAdd(&block_hist[find_histogram_index(bounds, e)], 1u);
  }
  if (g.do_lra && j >= 29 && (j - 29) % 10 == 0) {
    const double e = file_window_energy<30>(fb, stride, g, j);
    if (e >= bounds[0]) atomicAdd(&st_hist[find_histogram_index(bounds, e)], 1u);
  }
}

cudaError_t launch_file_gating(const GateParams& g, const LoudState& st, const double* d_file_buckets, size_t bucket_stride,
                               uint64_t n_buckets, cudaStream_t s, uint64_t* launches) {
  if (!n_buckets || !(g.do_i || g.do_lra)) return cudaSuccess;
  k_file_gating<<<(unsigned)((n_buckets + 127) / 128), 128, 0, s>>>(g, d_file_buckets, bucket_stride, n_buckets,
                                                                    st.hist_boundaries, st.block_hist, st.st_hist);
  if (launches) ++*launches;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// results: one warp per stream
// ------------------------------------------------------------------------------------------------
// Momentary / short-term energies from the ring of K-weighted samples (SSB_FLAG_RING handles: few streams,
// so one 512-thread block per stream; fixed-order tree reduction).  Window sums follow ebur128
// calc_gating_block: per channel sum of y^2 over the last W frames ending at ring_pos, surround x1.41.
__global__ void __launch_bounds__(512)
k_ring_energy(const GateParams g, const double* __restrict__ ring, size_t ring_frames, size_t ring_pos, int want_s,
              double* __restrict__ e_out /* [n][2] */) {
  __shared__ double red[512];
  const size_t s = blockIdx.x;
  const double* rg = ring + s * ring_frames * g.channels;
  for (int which = 0; which < (want_s ? 2 : 1); which++) {
    const size_t win = (size_t)g.s100 * (which ? 30 : 4);
    double total = 0.0;
    for (int c = 0; c < g.channels; c++) {
      const float w = g.weight[c];
      if (w == 0.0f) continue;
      double part = 0.0;
      for (size_t i = threadIdx.x; i < win; i += blockDim.x) {
        size_t idx = ring_pos + ring_frames - win + i;
        if (idx >= ring_frames) idx -= ring_frames;
        const double y = rg[idx * g.channels + c];
        part = fma(y, y, part);
      }
      red[threadIdx.x] = part;
      __syncthreads();
      for (int o = 256; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
      }
      double ch = red[0];
      __syncthreads();
      if (w != 1.0f) ch *= 1.41;
      total += ch;
    }
    if (threadIdx.x == 0) e_out[s * 2 + which] = total / (double)win;
  }
}

// 128 registers per thread: 16 resident warps per SM instead of 12 (measured 38 us against 54 us per 4096-stream query
// after a cfg2 launch)
__global__ void __launch_bounds__(128, 4)
k_results(const __grid_constant__ GateParams g, const __grid_constant__ ResultsArgs ra, size_t n_streams) {
  const size_t s = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s < n_streams) results_for_stream<R_ALL>(g, ra, ra.energies, ra.bounds, s, lane);
}

// The lean path (loudness_results.cuh): two streams per warp, eight per CTA; the launches that gate a 3 s entry finish
// each stream's row with the short-term histogram scan.
__global__ void __launch_bounds__(128)
k_results_lean(const __grid_constant__ GateParams g, const __grid_constant__ ResultsArgs ra, size_t n_streams) {
  __shared__ double stg[4][2][kLeanSlots * kLeanMaxChannels];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t s0 = ((size_t)blockIdx.x * 4 + warp) * 2;
  results_lean_pair(g, ra, &stg[warp][0][0], s0 < n_streams ? s0 : 0, s0 < n_streams, s0 + 1 < n_streams, lane);
}

cudaError_t launch_results(const GateParams& g, const LoudState& st, uint64_t buckets_done, int aligned,
                           size_t ring_pos, int mode, double* d_out, cudaStream_t s, uint64_t* launches,
                           uint64_t gate_first, uint64_t gate_last, const GatherArgs* ga, int lra_from_cache, int lean,
                           int st_back) {
  if (!st.n_streams) return cudaSuccess;
  const double* ring_e = nullptr;
  if (st.ring && st.ring_e) {
    k_ring_energy<<<(unsigned)st.n_streams, 512, 0, s>>>(g, st.ring, st.ring_frames, ring_pos,
                                                         (mode & SSB_MODE_S) == SSB_MODE_S, st.ring_e);
    if (launches) ++*launches;
    ring_e = st.ring_e;
  }
  const int tpb = 128;
  ResultsArgs ra = make_results_args(st, buckets_done, aligned, ring_pos, mode, d_out, gate_first, gate_last, ring_e);
  if (ga) ra.ga = *ga;
  ra.lra_from_cache = lra_from_cache;
  ra.lean = (lean && !st.ring && g.channels <= kLeanMaxChannels) ? 1 : 0;
  ra.st_back = ra.lean ? st_back : -1;
  ra.lra_fast = ra.st_back >= 0 ? 1 : 0;
  if (ra.lean) {
    k_results_lean<<<(unsigned)((st.n_streams + 7) / 8), tpb, 0, s>>>(g, ra, st.n_streams);
  } else {
    const size_t threads = st.n_streams * 32;
    k_results<<<(unsigned)((threads + tpb - 1) / tpb), tpb, 0, s>>>(g, ra, st.n_streams);
  }
  if (launches) ++*launches;
  return cudaGetLastError();
}

// find_histogram_index on an array of energies (tests: bin-edge cases against the crate's bisection); -1 below the
// absolute gate, as the `e >= boundaries[0]` tests of the gating code
__global__ void k_histogram_index(const double* __restrict__ e, size_t n, const double* __restrict__ bounds, int32_t* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = e[i] >= bounds[0] ? find_histogram_index(bounds, e[i]) : -1;
}
cudaError_t launch_histogram_index(const LoudState& st, const double* d_e, size_t n, int32_t* d_out, cudaStream_t s) {
  if (!n) return cudaSuccess;
  k_histogram_index<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_e, n, st.hist_boundaries, d_out);
  return cudaGetLastError();
}

cudaError_t launch_reset(const LoudState& st, int channels, cudaStream_t s, uint64_t* launches) {
  const size_t chains = st.n_streams * (size_t)channels;
  cudaError_t e;
  if ((e = cudaMemsetAsync(st.filt, 0, chains * 4 * sizeof(double), s))) return e;
  if ((e = cudaMemsetAsync(st.bucket, 0, chains * kNB * sizeof(double), s))) return e;
  if ((e = cudaMemsetAsync(st.block_hist, 0, st.n_streams * kHistBins * sizeof(uint32_t), s))) return e;
  if ((e = cudaMemsetAsync(st.st_hist, 0, st.n_streams * kHistBins * sizeof(uint32_t), s))) return e;
  if ((e = cudaMemsetAsync(st.speak, 0, chains * sizeof(float), s))) return e;
  if ((e = cudaMemsetAsync(st.tpeak, 0, chains * sizeof(float), s))) return e;
  if ((e = cudaMemsetAsync(st.tphist, 0, chains * kTpHist * sizeof(float), s))) return e;
  if (st.cache && (e = cudaMemsetAsync(st.cache, 0, st.n_streams * sizeof(StreamCache), s))) return e;
  if (st.ring && (e = cudaMemsetAsync(st.ring, 0, st.n_streams * st.ring_frames * channels * sizeof(double), s))) return e;
  (void)launches;
  return cudaSuccess;
}

}  // namespace ssb
