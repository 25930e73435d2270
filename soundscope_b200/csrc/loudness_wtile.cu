// loudness_wtile.cu — the batch K-weighting + gated-RMS (+ peaks) kernel for a few thousand streams per GPU
// (BASELINE config 2), second generation: warp-private TMA pipelines, balanced sub-partitions, fused results.
//
// Same arithmetic as the reference's `add_samples` (src/analyzer.rs:139-141 -> ebur128 add_frames_f32: 4th-order
// DF-II K-weighting in f64, y^2 summed per 100 ms, sample peak, polyphase true peak) followed, optionally in the
// same launch, by the queries of src/analyzer.rs:147-164 (loudness_results.cuh).
//
// What bounds it: the FP64 pipe (16 lanes per SM sub-partition: a warp DFMA holds it for 2 cycles) and the issue port, not
// HBM — a sample costs 9 DFMA (recursion 4, output 4, square 1), and splitting time across lanes adds a zero-state pass
// (4 DFMA) and a state hand-off.  So the design rule is "every sub-partition issues the same number of DFMA":
//   * one persistent CTA of 16 warps per SM, no producer warp.  Streams are grouped into 8 sets; a set owns a 3-stage
//     shared-memory ring with its own mbarriers and is served by TWO warps specialised by pass (see the banner above
//     run_p1): the P1 warp runs the zero-state recursion, the hand-off algebra and (Mode::all) the peak detectors, the
//     P2 warp the full filter and the bucket sums; the P2 warp's lane 0 issues the set's TMA box loads
//     (cp.async.bulk.tensor.3d, SWIZZLE_128B).  Sets never wait for each other inside the main loop.
//   * 4096 stereo streams over 148 SMs is 6.9 streams per sub-partition.  A warp's time per tile is its segment length L
//     whatever the number of active lanes, so a sub-partition's two sets are typed: type A takes 4 streams x 2 channels
//     x T=4 time segments (L = 80 frames of a 320-frame tile), type B takes 3 streams x 2 x T=5 (L = 64).  7 streams
//     cost 80 + 64 = 144 sample steps per 320 frames instead of the 2 x 80 a uniform T = 4 layout pays (the round-1
//     kernel: 7 of 8 warps live, 2-2-2-1 over the sub-partitions).
//   * per sample the output tap is computed as y / b0 = x + sum (b_i / b0 - a_i) v_i: 4 DFMA instead of 5, the
//     b0^2 applied once per tile to the partial sum; both recursions are software-pipelined (partial sums of the next
//     three samples in registers: one dependent DFMA per sample instead of four).
//   * after a set's last tile its two warps meet on a named barrier, gate the completed 100 ms buckets and write the
//     result rows of the set's streams (loudness_results.cuh: the lean path, two streams per warp), so "feed 400 ms,
//     read the meters" is one launch; with a stale cache the CTA falls back to the full histogram scan.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "loudness_results.cuh"
#include "ssb_internal.cuh"
#include "tma_ptx.cuh"

namespace ssb {

namespace {

constexpr int kWStages = 3;
constexpr int kWPairs = 8;  // sets of streams per CTA: 0..3 type A, 4..7 type B; every set has a P2 warp and a P1 warp
constexpr int kWWarps = 16;
constexpr int kWBars = 5;   // per pair: TMA full x 3, dk_full, dk_free

struct WArgs {
  double na[5];     // -a[i]
  double cy[5];     // b[i] / b[0] - a[i]  (cy[0] unused)
  double b0sq;      // b[0]^2
  double PA[16];    // D A^LA D, row-major: hand-off of a type-A segment
  double PB[16];    // D A^LB D
  double* filt;     // [n][C][4]
  double* bucket;   // [n][C][kNB]
  float* speak;     // [n][C]
  float* tpeak;     // [n][C]
  float* tphist;    // [n][C][kTpHist]
  float tp4[3][12];
  float tp2[24];
  float2 tp12[12];
  uint64_t active_mask;
  unsigned n_streams;
  unsigned n_tiles;  // tiles of F frames in this launch
  unsigned s100;
  unsigned pos0;
  unsigned slot0;
  int do_sample_peak;
  int fused_results;
};

// One warp type: T time segments of L frames, R streams of C channels; lane = k * (R*C) + row * C + channel.
template <int C_, int T_, int R_, int L_>
struct WType {
  static constexpr int C = C_, T = T_, R = R_, L = L_;
  static constexpr int Q = R * C;           // lanes per segment
  static constexpr int F = T * L;           // frames per tile
  static constexpr int NQ = L * C / 4;      // 16-byte quads per segment
  static constexpr int FPQ = 4 / C;         // frames per quad
  static constexpr int U = 16 / FPQ;        // quads per unrolled group (16 frames)
  static constexpr int SHQ = 4 / FPQ;       // quads of lookahead: the recursions consume samples four ahead
  static constexpr int NL = F * C / 32;     // 128-byte lines per stream per tile
  static constexpr unsigned TX_BYTES = R * NL * 128;
  static constexpr unsigned STAGE_STRIDE = (TX_BYTES + 1023u) / 1024u * 1024u;  // SWIZZLE_128B: 1 KB aligned stages
  static_assert(Q * T <= 32, "a warp holds all segments of its streams");
  static_assert(NQ % U == 0, "segments are whole 16-frame groups");
  static_assert((F * C) % 32 == 0, "tiles are whole 128-byte lines");
};

template <int C>
__device__ __forceinline__ float pickc(const float4& q, int f, int c) {
  if (C == 1) return f == 0 ? q.x : (f == 1 ? q.y : (f == 2 ? q.z : q.w));
  return f == 0 ? (c ? q.y : q.x) : (c ? q.w : q.z);
}

// A TMA box {32 floats, R streams, NL lines} lands as [line][stream][128 B]; SWIZZLE_128B XORs the 16-byte chunk
// index with bits 7..9 of the offset, i.e. with (line * R + stream) & 7.
template <class W>
__device__ __forceinline__ const unsigned char* float_addr(const unsigned char* stage, int rr, int fi) {
  const int qd = fi >> 2;
  const int rl = (qd >> 3) * W::R + rr;
  return stage + (rl << 7) + (((qd ^ rl) & 7) << 4) + ((fi & 3) << 2);
}

__device__ __forceinline__ void to_diff(double v1, double v2, double v3, double v4, double& d0, double& d1, double& d2,
                                        double& d3) {
  const double e1 = v1 - v2, e2 = v2 - v3, e3 = v3 - v4;
  d0 = v1;
  d1 = e1;
  d2 = e1 - e2;
  d3 = (e1 - e2) - (e2 - e3);
}
__device__ __forceinline__ void from_diff(double d0, double d1, double d2, double d3, double& v1, double& v2, double& v3,
                                          double& v4) {
  const double e2 = d1 - d2;
  const double e3 = e2 - (d2 - d3);
  v1 = d0;
  v2 = d0 - d1;
  v3 = v2 - e2;
  v4 = v3 - e3;
}

#define SSBW_CVT(x) ((double)(x))

// The recursions in software-pipelined form.  The reference's sample step is
//     v0 = (((x + na4 v4) + na3 v3) + na2 v2) + na1 v1          y/b0 = (((x + cy4 v4) + cy3 v3) + cy2 v2) + cy1 v1
// — per sample a chain of four dependent DFMA (33 cycles of latency) of which only the last link needs the newest
// state.  Written that way the compiler issues the chain in order and a warp advances one sample per 33 cycles
// (measured: the zero-state pass alone took 109 us for cfg2, an FP64 pipe busy 30 % of the time).  Here the partial
// sums of the next three samples are carried in registers instead of the older states:
//     pA = x[n]   + na4 v[n-4] + na3 v[n-3] + na2 v[n-2]     (complete but for the newest state)
//     pB = x[n+1] + na4 v[n-3] + na3 v[n-2]
//     pC = x[n+2] + na4 v[n-2]                  xh = x[n+3]
// and one step is  v0 = fma(na1, v1, pA); pA = fma(na2, v1, pB); pB = fma(na3, v1, pC); pC = fma(na4, v1, xh):
// four DFMA that depend on v1 only, the same additions in the same order (bit-identical results), a recurrence of
// ONE DFMA per sample.  The output tap is pipelined the same way (yA, yB, yC).  Samples are consumed four ahead.
#define SSBW_P2_INIT(x0, x1, x2, x3)                                                             \
  double pA = fma(a.na[2], v2, fma(a.na[3], v3, fma(a.na[4], v4, (double)(x0))));                \
  double pB = fma(a.na[3], v2, fma(a.na[4], v3, (double)(x1)));                                  \
  double pC = fma(a.na[4], v2, (double)(x2));                                                    \
  double yA = fma(a.cy[2], v2, fma(a.cy[3], v3, fma(a.cy[4], v4, (double)(x0))));                \
  double yB = fma(a.cy[3], v2, fma(a.cy[4], v3, (double)(x1)));                                  \
  double yC = fma(a.cy[4], v2, (double)(x2));                                                    \
  double xh = (double)(x3);                                                                      \
  float xd0 = (x0), xd1 = (x1), xd2 = (x2), xd3 = (x3);   /* x[n] .. x[n+3] for the peak detectors */
// xs: the sample four ahead, x[n+4]; ACC(y) takes y[n] / b0
#define SSBW_P2_STEP(xs, ACC)                                                                    \
  {                                                                                              \
    const float xf_ = xd0;                                                                       \
    xd0 = xd1; xd1 = xd2; xd2 = xd3; xd3 = (xs);                                                 \
    if (TPF != 0) sp = fmaxf(sp, fabsf(xf_));                                                    \
    const double v0_ = fma(a.na[1], v1, pA);                                                     \
    const double y_ = fma(a.cy[1], v1, yA);                                                      \
    pA = fma(a.na[2], v1, pB); pB = fma(a.na[3], v1, pC); pC = fma(a.na[4], v1, xh);             \
    yA = fma(a.cy[2], v1, yB); yB = fma(a.cy[3], v1, yC); yC = fma(a.cy[4], v1, xh);             \
    xh = (double)(xs);                                                                           \
    v1 = v0_;                                                                                    \
    ACC(y_)                                                                                      \
    SSBW_TP_STEP(xf_)                                                                            \
  }
#define SSBW_P1_INIT(x0, x1, x2, x3) \
  double pA = (double)(x0), pB = (double)(x1), pC = (double)(x2), xh = (double)(x3);
#define SSBW_P1_STEP(xs)                                                                         \
  {                                                                                              \
    const double z0_ = fma(a.na[1], z1, pA);                                                     \
    pA = fma(a.na[2], z1, pB); pB = fma(a.na[3], z1, pC); pC = fma(a.na[4], z1, xh);             \
    xh = (double)(xs);                                                                           \
    z4 = z3; z3 = z2; z2 = z1; z1 = z0_;                                                         \
  }

// One true-peak sample: ebur128's polyphase interpolator as f32 FMAs over the register window w2[t] = x[n-1-t]
// (same tap order in every kernel, so all kernels report identical true peaks).
#define SSBW_TP_STEP(xf)                                                        \
  if (TPF == 4) {                                                               \
    const float2 xx_ = make_float2((xf), (xf));                                 \
    float2 acc12_ = make_float2((xf) * a.tp12[0].x, (xf) * a.tp12[0].y);        \
    float acc3_ = (xf) * a.tp4[2][0];                                           \
    _Pragma("unroll") for (int t = 1; t < 12; t++) {                            \
      acc12_ = __ffma2_rn(w2[t - 1], a.tp12[t], acc12_);                        \
      acc3_ = fmaf(w2[t - 1].x, a.tp4[2][t], acc3_);                            \
    }                                                                           \
    tp = fmaxf(tp, fmaxf(fabsf(acc12_.x), fmaxf(fabsf(acc12_.y), fabsf(acc3_)))); \
    _Pragma("unroll") for (int t = TPW - 1; t > 0; t--) w2[t] = w2[t - 1];      \
    w2[0] = xx_;                                                                \
  } else if (TPF == 2) {                                                        \
    float acc_ = (xf) * a.tp2[0];                                               \
    _Pragma("unroll") for (int t = 1; t < 24; t++) acc_ = fmaf(w2[t - 1].x, a.tp2[t], acc_); \
    tp = fmaxf(tp, fabsf(acc_));                                                \
    _Pragma("unroll") for (int t = TPW - 1; t > 0; t--) w2[t] = w2[t - 1];      \
    w2[0] = make_float2((xf), (xf));                                            \
  }

// fixed-order sum over the T segments of a chain: ((k0 + k1) + (k2 + k3)) [+ k4]
template <class W>
__device__ __forceinline__ double seg_sum(double v, int q) {
  if (W::T == 4 && W::Q == 8) {   // lanes k * 8 + q: two butterfly rounds give every lane (k0 + k1) + (k2 + k3)
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
  }
  double s[W::T];
#pragma unroll
  for (int k = 0; k < W::T; k++) s[k] = __shfl_sync(0xffffffffu, v, k * W::Q + q);
  double r = (s[0] + s[1]) + (s[2] + s[3]);
  if (W::T == 5) r += s[4];
  return r;
}
template <class W>
__device__ __forceinline__ float seg_max(float v, int q) {
  float r = 0.f;
#pragma unroll
  for (int k = 0; k < W::T; k++) r = fmaxf(r, __shfl_sync(0xffffffffu, v, k * W::Q + q));
  return r;
}

// ------------------------------------------------------------------------------------------------------------------
// Two warps per set of streams, specialised by pass, reading the same shared-memory tiles:
//   the P1 warp runs the zero-state recursion (4 DFMA / sample) over every tile as soon as it lands, then the hand-off:
//       the carry e (state at the tile start, in difference coordinates) advances by e <- Pt e + D z_j, j = 0..T-1, lane
//       k keeping e after k steps as its segment's start state — linear algebra on the zero-state end states only, so
//       the true states never have to come out of the full filter — and publishes the start states through a
//       1 KB shared-memory slot (mbarrier dk_full / dk_free);
//   the P2 warp picks the start states up, runs the full filter (9 DFMA / sample, + peaks), sums the buckets and hands
//       the stage back to the TMA ring.
// Four warps per sub-partition (P1 and P2 of a type-A and of a type-B set) progress independently: while one is in
// its per-tile serial work (hand-off chain, reductions, barrier waits, TMA issue) the others keep the FP64 pipe busy.
// The first generation of this kernel ran both passes in one warp (two warps per sub-partition) and lost a quarter of
// every warp's time to that serial work (profiles/r2_wtile_v1.md).
// ------------------------------------------------------------------------------------------------------------------
struct PairGeom {
  unsigned cta_row0, cta_rows, rows_per_pass, n_pass, warp_off;
};

template <class W>
__device__ __forceinline__ unsigned count_tasks(const PairGeom& gm) {
  // passes in which this pair owns at least one stream (the last pass may be shorter: the count is monotone)
  unsigned n_task = 0;
  for (unsigned p = 0; p < gm.n_pass; p++) {
    const unsigned rp = min(gm.rows_per_pass, gm.cta_rows - p * gm.rows_per_pass);
    if (rp > gm.warp_off) n_task++;
  }
  return n_task;
}

// quad qd of stream rr inside a stage (lines past the tile are clamped: only lookahead samples that are never used
// come from there)
template <class W>
__device__ __forceinline__ const unsigned char* quad_ptr(const unsigned char* stage, int rr, int qd) {
  const int li = min(qd >> 3, W::NL - 1);
  const int rl = li * W::R + rr;
  return stage + (rl << 7) + (((qd ^ rl) & 7) << 4);
}
// The sample stream of a segment runs four samples (SHQ quads) ahead of the recursion.
// Stereo: a lane needs ONE float (its channel) per frame, so it reads 4-byte words: the byte offset of its word in
// quad chunk `ch` of line `la` is ((la * R + rr) << 7) + ((ch ^ key) << 4) + 4 c with key = (la * R + rr) & 7, i.e.
// line_off ^ (ch << 4) with line_off = (rl << 7) ^ (key << 4) | 4 c — one XOR per quad, the second frame of the quad
// at +8 (an immediate), no channel select.  A group's 8 quads are chunks 2..7 of line LA and chunks 0..1 of line
// LA + 1 (segments start on line boundaries).  Mono: 4 quads per group, 16-byte loads, generic addressing.
#define SSBW_GROUP_BASES(stage, g)                                                               \
  unsigned offA_ = 0, offB_ = 0;                                                                 \
  if (W::C == 2) {                                                                               \
    const int la_ = (q0 >> 3) + (g), lb_ = min(la_ + 1, W::NL - 1);                              \
    const unsigned ra_ = (unsigned)(la_ * W::R + rr), rb_ = (unsigned)(lb_ * W::R + rr);         \
    const unsigned sb_ = smem_u32(stage);   /* 1 KB aligned: the XORs below touch bits 2..6 only */ \
    offA_ = ((sb_ + (ra_ << 7)) ^ ((ra_ & 7u) << 4)) | ((unsigned)c << 2);                       \
    offB_ = ((sb_ + (rb_ << 7)) ^ ((rb_ & 7u) << 4)) | ((unsigned)c << 2);                       \
  }
// sample f (0 .. FPQ-1) of quad j of group g, for this lane's channel
#define SSBW_GROUP_SAMPLE(stage, g, j, f, qv)                                                                      \
  (W::C == 2 ? *reinterpret_cast<const float*>(__cvta_shared_to_generic(((j) < 6 ? offA_ ^ (((j) + 2u) << 4) : offB_ ^ (((j) - 6u) << 4)) + 8u * (f))) \
             : pickc<1>(qv, f, 0))
#define SSBW_GROUP_QUADLOAD(stage, g, j)                                                                           \
  (W::C == 2 ? make_float4(0.f, 0.f, 0.f, 0.f)                                                                     \
             : *reinterpret_cast<const float4*>(quad_ptr<W>((stage), rr, q0 + W::SHQ + (g) * W::U + (j))))
// the first four samples of my segment
#define SSBW_FIRST4(stage, x0, x1, x2, x3)                                                       \
  float x0, x1, x2, x3;                                                                          \
  {                                                                                              \
    const float4 qa_ = *reinterpret_cast<const float4*>(quad_ptr<W>((stage), rr, q0));           \
    if (W::C == 2) {                                                                             \
      const float4 qb_ = *reinterpret_cast<const float4*>(quad_ptr<W>((stage), rr, q0 + 1));     \
      x0 = pickc<2>(qa_, 0, c); x1 = pickc<2>(qa_, 1, c); x2 = pickc<2>(qb_, 0, c); x3 = pickc<2>(qb_, 1, c); \
    } else {                                                                                     \
      x0 = qa_.x; x1 = qa_.y; x2 = qa_.z; x3 = qa_.w;                                            \
    }                                                                                            \
  }

template <class W, bool IS_B, int TPF>
__device__ __forceinline__ void run_p1(const WArgs& a, unsigned char* stages, uint64_t* full, uint64_t* dk_full,
                                       uint64_t* dk_free, double* dk_slot, const PairGeom gm, const int lane) {
  constexpr int C = W::C, T = W::T, Q = W::Q, L = W::L;
  constexpr int NG = L / 16;
  constexpr int TPW = TPF == 4 ? 11 : (TPF == 2 ? 23 : 1);  // true-peak FIR history length
  static_assert(TPW <= L, "the FIR history of a segment lies inside the previous one");
#define SSBW_P(i) (IS_B ? a.PB[i] : a.PA[i])
  const unsigned n_task = count_tasks<W>(gm);
  if (n_task * a.n_tiles == 0) return;
  const bool lane_used = lane < T * Q;
  const int k = lane_used ? lane / Q : 0;
  const int q = lane_used ? lane - k * Q : 0;
  const int rr = q / C;
  const int c = q - rr * C;
  const int q0 = k * W::NQ;  // first quad of my segment

  for (unsigned task = 0; task < n_task; task++) {
    const unsigned rp = min(gm.rows_per_pass, gm.cta_rows - task * gm.rows_per_pass);
    const unsigned nrows = min((unsigned)W::R, rp - gm.warp_off);
    const unsigned row_g = gm.cta_row0 + task * gm.rows_per_pass + gm.warp_off;
    const bool row_ok = lane_used && (unsigned)rr < nrows;
    const bool live = row_ok && ((a.active_mask >> c) & 1ull);
    const size_t gidx = ((size_t)(row_g + (row_ok ? rr : 0))) * C + c;
    const unsigned g0 = task * a.n_tiles;
    // e: the chain's true state at the tile start, in difference coordinates, replicated in the chain's T lanes
    double e0 = 0, e1 = 0, e2 = 0, e3 = 0;
    if (live) {
      const double* f = a.filt + gidx * 4;
      to_diff(f[0], f[1], f[2], f[3], e0, e1, e2, e3);
    }
    // The peak detectors depend on the input only, so they ride with this (lighter) pass: sample peak and the polyphase
    // true-peak FIR over the register window w2[t] = x[n-1-t]
    float sp = 0.f, tp = 0.f;
    float hist[TPW];   // the TPW samples before the next tile (meaningful in the k == 0 lanes)
#pragma unroll
    for (int t = 0; t < TPW; t++) hist[t] = (TPF >= 2 && row_ok) ? a.tphist[gidx * kTpHist + t] : 0.f;
    for (unsigned tile = 0; tile < a.n_tiles; tile++) {
      const unsigned g = g0 + tile;
      mbar_wait_warp(&full[g % kWStages], (g / kWStages) & 1);
      const unsigned char* st = stages + (size_t)(g % kWStages) * W::STAGE_STRIDE;
      double z1 = 0, z2 = 0, z3 = 0, z4 = 0;
      // FIR window: the TPW samples before my segment (previous segment's tail in the same tile, or, for the first
      // segment, the previous tile's tail carried in `hist`); both halves equal so a tap feeds two phases in one FFMA2
      float2 w2[TPW];
      if (TPF >= 2) {
#pragma unroll
        for (int t = 0; t < TPW; t++) {
          const int fr = k > 0 ? k * L - 1 - t : 0;
          const float prev = *reinterpret_cast<const float*>(float_addr<W>(st, rr, fr * C + c));
          const float wv = k > 0 ? prev : hist[t];
          w2[t] = make_float2(wv, wv);
        }
      }
      {
        SSBW_FIRST4(st, x0, x1, x2, x3)
        SSBW_P1_INIT(x0, x1, x2, x3)
        float xd0 = x0, xd1 = x1, xd2 = x2, xd3 = x3;   // x[n] .. x[n+3] for the peak detectors
#ifdef SSBW_DIAG_NO_P1
        for (int gi = 0; gi < 0; gi++) {
#else
#pragma unroll 1
        for (int gi = 0; gi < NG; gi++) {
#endif
          SSBW_GROUP_BASES(st, gi)
#pragma unroll
          for (int j = 0; j < W::U; j++) {
            const float4 qn = SSBW_GROUP_QUADLOAD(st, gi, j);
#pragma unroll
            for (int f = 0; f < W::FPQ; f++) {
              const float xs_ = SSBW_GROUP_SAMPLE(st, gi, j, f, qn);
              SSBW_P1_STEP(xs_)
              if (TPF != 0) {
                const float xf_ = xd0;
                xd0 = xd1; xd1 = xd2; xd2 = xd3; xd3 = xs_;
                sp = fmaxf(sp, fabsf(xf_));
                SSBW_TP_STEP(xf_)
              }
            }
          }
        }
      }
      if (TPF >= 2) {
        // the last segment's tail is the history of the next tile's first segment
#pragma unroll
        for (int t = 0; t < TPW; t++) hist[t] = __shfl_sync(0xffffffffu, w2[t].x, (T - 1) * Q + q);
      }
      // hand-off: lane k keeps the carry after k links as its own start state; after T links e is the next tile's carry
      double zd0, zd1, zd2, zd3;
      to_diff(z1, z2, z3, z4, zd0, zd1, zd2, zd3);
      double dk0 = e0, dk1 = e1, dk2 = e2, dk3 = e3;
#pragma unroll
      for (int j = 0; j < T; j++) {
        const int src = j * Q + q;
        const double zj0 = __shfl_sync(0xffffffffu, zd0, src), zj1 = __shfl_sync(0xffffffffu, zd1, src);
        const double zj2 = __shfl_sync(0xffffffffu, zd2, src), zj3 = __shfl_sync(0xffffffffu, zd3, src);
        const double n0 = fma(SSBW_P(0), e0, fma(SSBW_P(1), e1, fma(SSBW_P(2), e2, fma(SSBW_P(3), e3, zj0))));
        const double n1 = fma(SSBW_P(4), e0, fma(SSBW_P(5), e1, fma(SSBW_P(6), e2, fma(SSBW_P(7), e3, zj1))));
        const double n2 = fma(SSBW_P(8), e0, fma(SSBW_P(9), e1, fma(SSBW_P(10), e2, fma(SSBW_P(11), e3, zj2))));
        const double n3 = fma(SSBW_P(12), e0, fma(SSBW_P(13), e1, fma(SSBW_P(14), e2, fma(SSBW_P(15), e3, zj3))));
        e0 = n0; e1 = n1; e2 = n2; e3 = n3;
        if (k == j + 1) { dk0 = n0; dk1 = n1; dk2 = n2; dk3 = n3; }
      }
      // publish the start states once the P2 warp has taken the previous tile's
      if (g >= 1) mbar_wait_warp(dk_free, (g - 1) & 1);
      reinterpret_cast<double2*>(dk_slot)[lane * 2] = make_double2(dk0, dk1);
      reinterpret_cast<double2*>(dk_slot)[lane * 2 + 1] = make_double2(dk2, dk3);
      __syncwarp();
      if (lane == 0) mbar_arrive(dk_full);
    }
    // the carry after the last hand-off is the state at the end of the last tile
    if (row_ok && k == 0 && live) {
      double c1, c2, c3, c4;
      from_diff(e0, e1, e2, e3, c1, c2, c3, c4);
      double* f = a.filt + gidx * 4;
      const double tiny = 2.2250738585072014e-308;  // libebur128: flush denormal state at the end of a call
      f[0] = fabs(c1) < tiny ? 0.0 : c1;
      f[1] = fabs(c2) < tiny ? 0.0 : c2;
      f[2] = fabs(c3) < tiny ? 0.0 : c3;
      f[3] = fabs(c4) < tiny ? 0.0 : c4;
    }
    sp = seg_max<W>(sp, q);
    tp = seg_max<W>(tp, q);
    if (row_ok && k == 0) {
      if (TPF != 0) a.speak[gidx] = fmaxf(a.speak[gidx], sp);
      if (TPF >= 2) {
        a.tpeak[gidx] = fmaxf(a.tpeak[gidx], tp);
#pragma unroll
        for (int t = 0; t < TPW; t++) a.tphist[gidx * kTpHist + t] = hist[t];
      }
    }
  }
#undef SSBW_P
}

template <class W>
__device__ __forceinline__ void run_p2(const WArgs& a, const CUtensorMap* tmap, unsigned char* stages, uint64_t* full,
                                       uint64_t* dk_full, uint64_t* dk_free, const double* dk_slot, const PairGeom gm,
                                       const int lane) {
  constexpr int C = W::C, T = W::T, Q = W::Q, L = W::L, F = W::F;
  constexpr int NG = L / 16;
  constexpr int TPF = 0;   // the peak detectors live in the P1 warp
  static_assert(L % 16 == 0, "segments are whole groups");
  const unsigned n_task = count_tasks<W>(gm);
  const unsigned total_tiles = n_task * a.n_tiles;
  if (total_tiles == 0) return;

  // linear tile counter g = task * n_tiles + tile: ring stage g % 3, barrier phase (g / 3) & 1
  auto issue = [&](unsigned g) {
    const unsigned p = g / a.n_tiles, t = g - p * a.n_tiles;
    const unsigned s = g % kWStages;
    mbar_expect_tx(&full[s], W::TX_BYTES);
    tma_load_3d(stages + (size_t)s * W::STAGE_STRIDE, tmap, &full[s], 0,
                (int)(gm.cta_row0 + p * gm.rows_per_pass + gm.warp_off), (int)(t * W::NL));
  };
  if (lane == 0) {
    for (unsigned g = 0; g < (unsigned)kWStages && g < total_tiles; g++) issue(g);
  }

  const bool lane_used = lane < T * Q;
  const int k = lane_used ? lane / Q : 0;
  const int q = lane_used ? lane - k * Q : 0;
  const int rr = q / C;
  const int c = q - rr * C;
  const int q0 = k * W::NQ;  // first quad of my segment

  for (unsigned task = 0; task < n_task; task++) {
    const unsigned rp = min(gm.rows_per_pass, gm.cta_rows - task * gm.rows_per_pass);
    const unsigned nrows = min((unsigned)W::R, rp - gm.warp_off);
    const unsigned row_g = gm.cta_row0 + task * gm.rows_per_pass + gm.warp_off;  // first stream of this pair's box
    const bool row_ok = lane_used && (unsigned)rr < nrows;
    const bool live = row_ok && ((a.active_mask >> c) & 1ull);
    const bool owner = row_ok && k == 0;
    const size_t gidx = ((size_t)(row_g + (row_ok ? rr : 0))) * C + c;
    const unsigned g0 = task * a.n_tiles;

    double acc_cur = 0.0;  // owner lane: running sum (already times b0^2) of the bucket in progress
    unsigned slot = a.slot0;
    if (owner && live && a.pos0 > 0) acc_cur = a.bucket[gidx * kNB + slot];
    float sp = 0.f, tp = 0.f;   // unused here (TPF == 0): the step macro is shared with the peak-carrying pass
    constexpr int TPW = 1;
    float2 w2[TPW];
    w2[0] = make_float2(0.f, 0.f);
    unsigned pos_tile = a.pos0;   // position of the tile start inside the bucket in progress

    for (unsigned tile = 0; tile < a.n_tiles; tile++) {
      const unsigned g = g0 + tile;
      const unsigned char* st0 = stages + (size_t)(g % kWStages) * W::STAGE_STRIDE;
      // ---- start states from the P1 warp.  It has observed the tile's mbarrier phase (which is what makes the TMA writes
      //      visible to it) before it computed them, and dk_full's release / acquire pair orders everything it has seen
      //      before this warp's reads: no second wait on `full` (it cost 1.5 us per launch) ----
      mbar_wait_warp(dk_full, g & 1);
      const double2 da = reinterpret_cast<const double2*>(dk_slot)[lane * 2];
      const double2 db = reinterpret_cast<const double2*>(dk_slot)[lane * 2 + 1];
      __syncwarp();
      if (lane == 0) mbar_arrive(dk_free);
#ifdef SSBW_SECOND_WAIT
      mbar_wait_warp(&full[g % kWStages], (g / kWStages) & 1);
#endif
      double v1, v2, v3, v4;
      from_diff(da.x, da.y, db.x, db.y, v1, v2, v3, v4);
      const unsigned to_boundary = a.s100 - pos_tile;  // frames of this tile before the bucket boundary (>= F: none inside)
      int lb = (int)to_boundary - k * L;               // my samples [0, lb) belong to the bucket in progress
      lb = lb < 0 ? 0 : (lb > L ? L : lb);
      const unsigned mixed = __reduce_or_sync(0xffffffffu, (unsigned)(lb & 15));  // a boundary strictly inside a group
      double accA = 0.0, accB = 0.0;
      {
        SSBW_FIRST4(st0, x0, x1, x2, x3)
        SSBW_P2_INIT(x0, x1, x2, x3)
        if (!mixed) {
#define SSBW_ACC_FAST(y) acc = fma((y), (y), acc);
#ifdef SSBW_DIAG_NO_P2
          for (int gi = 0; gi < 0; gi++) {
#else
#pragma unroll 1
          for (int gi = 0; gi < NG; gi++) {
#endif
            SSBW_GROUP_BASES(st0, gi)
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < W::U; j++) {
              const float4 qv = SSBW_GROUP_QUADLOAD(st0, gi, j);
#pragma unroll
              for (int f = 0; f < W::FPQ; f++) SSBW_P2_STEP(SSBW_GROUP_SAMPLE(st0, gi, j, f, qv), SSBW_ACC_FAST)
            }
            if (16 * (gi + 1) <= lb) accA += acc; else accB += acc;
          }
#undef SSBW_ACC_FAST
        } else {
          // a bucket boundary strictly inside some lane's group (rates whose 100 ms is not whole tiles): per-sample test
          int i_ = 0;
#define SSBW_ACC_SLOW(y) if (i_ < lb) accA = fma((y), (y), accA); else accB = fma((y), (y), accB); i_++;
#pragma unroll 1
          for (int qi = 0; qi < W::NQ; qi++) {
            const float4 qv = *reinterpret_cast<const float4*>(quad_ptr<W>(st0, rr, q0 + W::SHQ + qi));
#pragma unroll
            for (int f = 0; f < W::FPQ; f++) SSBW_P2_STEP(pickc<C>(qv, f, c), SSBW_ACC_SLOW)
          }
#undef SSBW_ACC_SLOW
        }
      }
      // ---- the tile is done: FIR history, stage back to the TMA ring (the tile three ahead, possibly the next pass's),
      //      bucket sums (fixed-order reduction over the segments, times b0^2) ----
      __syncwarp();
      if (lane == 0 && g + kWStages < total_tiles) issue(g + kWStages);
      const double sA = seg_sum<W>(accA, q);
      if (to_boundary <= (unsigned)F) {
        const double sB = seg_sum<W>(accB, q);
        if (k == 0) {
          acc_cur = fma(a.b0sq, sA, acc_cur);
          if (owner && live) a.bucket[gidx * kNB + slot] = acc_cur;
          acc_cur = a.b0sq * sB;
          slot = (slot + 1) % kNB;
        }
      } else if (k == 0) {
        acc_cur = fma(a.b0sq, sA, acc_cur);
      }
      pos_tile += F;
      if (pos_tile >= a.s100) pos_tile -= a.s100;
    }

    // ---------------- end of this pair's streams for this pass: bucket in progress, peaks ----------------
    (void)sp; (void)tp; (void)w2;
    if (owner) a.bucket[gidx * kNB + slot] = live ? acc_cur : 0.0;
  }
}
#undef SSBW_GROUP_BASES
#undef SSBW_GROUP_SAMPLE
#undef SSBW_GROUP_QUADLOAD
#undef SSBW_FIRST4

// MIXED: type A = (T 4, L 80), type B = (T 5, L 64), 320-frame tiles, 28 stereo streams per pass;
// !MIXED: every set (T 4, L 64), 256-frame tiles, 32 stereo streams per pass.
template <int C, bool MIXED>
struct WCfg {
#ifdef SSBW_UNIFORM_L80   // experiment: both set types T = 4, L = 80 (4 + 3 streams): equal steps per tile, 11 % more FP64 work
  using A = WType<C, 4, 8 / C, MIXED ? 80 : 64>;
  using B = WType<C, 4, (MIXED ? 6 : 8) / C, MIXED ? 80 : 64>;
#else
  using A = WType<C, 4, 8 / C, MIXED ? 80 : 64>;
  using B = WType<C, MIXED ? 5 : 4, (MIXED ? 6 : 8) / C, 64>;
#endif
  static_assert(A::F == B::F, "both set types walk the same tiles");
  static constexpr int F = A::F;
  static constexpr int CAP = 4 * A::R + 4 * B::R;  // streams per CTA pass
  static constexpr size_t STAGES = (size_t)4 * kWStages * (A::STAGE_STRIDE + B::STAGE_STRIDE);
  static constexpr size_t DK = (size_t)kWPairs * 32 * 4 * sizeof(double);        // one start-state slot per pair
  static constexpr size_t BARS = (size_t)kWPairs * kWBars * sizeof(uint64_t);
  static constexpr size_t SMEM = STAGES + DK + BARS + 1024;
};

template <int C, int TPF, bool MIXED>
__global__ void __launch_bounds__(kWWarps * 32, 1)
k_loudness_wtile(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
                 const __grid_constant__ WArgs a, const __grid_constant__ GateParams g,
                 const __grid_constant__ ResultsArgs ra) {
  using Cfg = WCfg<C, MIXED>;
  using WA = typename Cfg::A;
  using WB = typename Cfg::B;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  double* dk_all = reinterpret_cast<double*>(smem + Cfg::STAGES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES + Cfg::DK);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = warp & (kWPairs - 1);     // warps 0..7: the P2 warps, 8..15: the P1 warps of the same pairs
  // pair p and p + 4, P1 and P2: warps p, p+4, p+8, p+12 share sub-partition p % 4.  The P1 warps sit on the LOW warp
  // ids: measured 1.5 % faster than the other way round (196.5 -> 192.8 us with the change below).
  const bool is_p1 = warp < kWPairs;
  const unsigned n_ctas = gridDim.x;
  const unsigned row0 = (unsigned)(((unsigned long long)blockIdx.x * a.n_streams) / n_ctas);
  const unsigned row1 = (unsigned)(((unsigned long long)(blockIdx.x + 1) * a.n_streams) / n_ctas);
  const unsigned cta_rows = row1 - row0;
  const unsigned n_pass = (cta_rows + Cfg::CAP - 1) / Cfg::CAP;
  const unsigned rows_per_pass = n_pass ? (cta_rows + n_pass - 1) / n_pass : 0;

  uint64_t* full = bars + pair * kWBars;    // [0..2] TMA full, [3] dk_full, [4] dk_free
  if (warp < kWPairs && lane == 0) {
    for (int s = 0; s < kWBars; s++) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();

  if (cta_rows) {
    PairGeom gm;
    gm.cta_row0 = row0;
    gm.cta_rows = cta_rows;
    gm.rows_per_pass = rows_per_pass;
    gm.n_pass = n_pass;
    double* dk_slot = dk_all + (size_t)pair * 32 * 4;
    if (pair < 4) {
      unsigned char* stages = smem + (size_t)pair * kWStages * WA::STAGE_STRIDE;
      gm.warp_off = (unsigned)pair * WA::R;
      if (is_p1) run_p1<WA, false, TPF>(a, stages, full, full + 3, full + 4, dk_slot, gm, lane);
      else run_p2<WA>(a, &tmapA, stages, full, full + 3, full + 4, dk_slot, gm, lane);
    } else {
      unsigned char* stages = smem + (size_t)4 * kWStages * WA::STAGE_STRIDE + (size_t)(pair - 4) * kWStages * WB::STAGE_STRIDE;
      gm.warp_off = 4u * WA::R + (unsigned)(pair - 4) * WB::R;
      if (is_p1) run_p1<WB, true, TPF>(a, stages, full, full + 3, full + 4, dk_slot, gm, lane);
      else run_p2<WB>(a, &tmapB, stages, full, full + 3, full + 4, dk_slot, gm, lane);
    }
  }
  if (a.fused_results && ra.lean) {
    // gating + result rows (analyzer.rs:147-164), lean form: the two warps of a pair meet on their own named barrier as
    // soon as the pair's last tile is done and split the pair's streams, two per warp at a time (16 lanes each); the
    // pairs that finish early (the lighter type-B sets) do this while the others still filter.  No CTA-wide barrier.
    if (cta_rows) {
      PairGeom gm;
      gm.cta_row0 = row0;
      gm.cta_rows = cta_rows;
      gm.rows_per_pass = rows_per_pass;
      gm.n_pass = n_pass;
      const unsigned R = pair < 4 ? (unsigned)WA::R : (unsigned)WB::R;
      gm.warp_off = pair < 4 ? (unsigned)pair * WA::R : 4u * WA::R + (unsigned)(pair - 4) * WB::R;
      unsigned char* stages = pair < 4 ? smem + (size_t)pair * kWStages * WA::STAGE_STRIDE
                                       : smem + (size_t)4 * kWStages * WA::STAGE_STRIDE + (size_t)(pair - 4) * kWStages * WB::STAGE_STRIDE;
      const unsigned n_task = count_tasks<WA>(gm);
      if (n_task) {
        // the bucket sums (P2 warp) and peaks (P1 warp) of the pair's streams are visible to both warps after this
        asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");
        double* stg = reinterpret_cast<double*>(stages) + (is_p1 ? 2 : 0) * (kLeanSlots * C);
        for (unsigned task = 0; task < n_task; task++) {
          const unsigned rp = min(rows_per_pass, cta_rows - task * rows_per_pass);
          const unsigned nrows = min(R, rp - gm.warp_off);
          const size_t row_g = (size_t)row0 + task * rows_per_pass + gm.warp_off;
          for (unsigned item = (task + (is_p1 ? 1u : 0u)) & 1u; item * 2 < nrows; item += 2) {
            __syncwarp();   // the previous item's reads of the staging slots are done
            results_lean_pair(g, ra, stg, row_g + item * 2, true, item * 2 + 1 < nrows, lane);
          }
        }
      }
    }
  } else if (a.fused_results) {
    // the full form (stale cache, many pending buckets): one warp per stream after a CTA-wide barrier; the bucket sums
    // and peaks written above by other warps of this CTA are visible after the barrier
    __syncthreads();
    // the histogram tables (energies[1000] | boundaries[1001], contiguous) into the now idle stage memory
    double* tab = reinterpret_cast<double*>(smem);
    for (int i = threadIdx.x; i < 2 * kHistBins + 1; i += kWWarps * 32) tab[i] = __ldg(ra.energies + i);
    __syncthreads();
    for (unsigned r = warp; r < cta_rows; r += kWWarps) results_for_stream<R_ALL>(g, ra, tab, tab + kHistBins, (size_t)row0 + r, lane);
  }
}
#undef SSBW_P2_INIT
#undef SSBW_P2_STEP
#undef SSBW_P1_INIT
#undef SSBW_P1_STEP
#undef SSBW_TP_STEP
#undef SSBW_CVT

template <int C, int TPF, bool MIXED>
cudaError_t launch_cfg(const CUtensorMap& tA, const CUtensorMap& tB, const WArgs& a, const GateParams& g,
                       const ResultsArgs& ra, unsigned n_ctas, int device, cudaStream_t s) {
  auto kern = k_loudness_wtile<C, TPF, MIXED>;
  const size_t smem = WCfg<C, MIXED>::SMEM;
  static bool configured_dev[64] = {false};  // per instantiation and device
  bool& configured = configured_dev[device & 63];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e) return e;
    configured = true;
  }
  kern<<<n_ctas, kWWarps * 32, smem, s>>>(tA, tB, a, g, ra);
  return cudaGetLastError();
}

template <int C, bool MIXED>
cudaError_t launch_c(const CUtensorMap& tA, const CUtensorMap& tB, const WArgs& a, const GateParams& g,
                     const ResultsArgs& ra, unsigned n_ctas, int tpf, int device, cudaStream_t s) {
  if (tpf == 4) return launch_cfg<C, 4, MIXED>(tA, tB, a, g, ra, n_ctas, device, s);
  if (tpf == 2) return launch_cfg<C, 2, MIXED>(tA, tB, a, g, ra, n_ctas, device, s);
  if (tpf == 1) return launch_cfg<C, 1, MIXED>(tA, tB, a, g, ra, n_ctas, device, s);
  return launch_cfg<C, 0, MIXED>(tA, tB, a, g, ra, n_ctas, device, s);
}

template <class W>
bool encode_box(CUtensorMap* tm, const float* d_in, size_t n_streams, size_t row_floats, size_t used_floats) {
  // 3-D view of the [stream][frame][channel] input: dim0 = 32 floats of one 128-byte line, dim1 = stream (row pitch),
  // dim2 = line index along the row
  cuuint64_t gdim[3] = {32, (cuuint64_t)n_streams, (cuuint64_t)(used_floats / 32)};
  cuuint64_t gstride[2] = {(cuuint64_t)(row_floats * sizeof(float)), 128};
  cuuint32_t box[3] = {32, (cuuint32_t)W::R, (cuuint32_t)W::NL};
  cuuint32_t estride[3] = {1, 1, 1};
  return tma_encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(d_in), gdim, gstride, box, estride,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int C, bool MIXED>
cudaError_t launch_shape(const LoudParams& p, const LoudState& st, const float* d_in, size_t frames, size_t in_stride_frames,
                         WArgs& a, const GateParams& g, const ResultsArgs& ra, int sm_count, int device, cudaStream_t s,
                         size_t* consumed) {
  using Cfg = WCfg<C, MIXED>;
  const size_t n_tiles = frames / Cfg::F;
  if (!n_tiles) return cudaSuccess;
  const size_t row_floats = in_stride_frames * C;
  const size_t used_floats = n_tiles * Cfg::F * C;
  CUtensorMap tA, tB;
  if (!encode_box<typename Cfg::A>(&tA, d_in, st.n_streams, row_floats, used_floats) ||
      !encode_box<typename Cfg::B>(&tB, d_in, st.n_streams, row_floats, used_floats))
    return cudaErrorInvalidValue;
  memcpy(a.PA, Cfg::A::L == 80 ? p.handoff80 : p.handoff, sizeof(a.PA));   // cached per meter (init_meter)
  memcpy(a.PB, Cfg::B::L == 80 ? p.handoff80 : p.handoff, sizeof(a.PB));
  static_assert((Cfg::B::L == 64 || Cfg::B::L == 80) && (Cfg::A::L == 80 || Cfg::A::L == 64), "hand-off matrices cached for 64 and 80 frames");
  a.n_tiles = (unsigned)n_tiles;
  const unsigned n_ctas = (unsigned)(st.n_streams < (size_t)sm_count ? st.n_streams : (size_t)sm_count);
  // 4 / 2: true-peak FIR (ebur128's rate rule) + sample peak; 1: sample peak only; 0: neither
  const int tpf = (p.do_true_peak && p.tp_factor) ? p.tp_factor : (p.do_sample_peak ? 1 : 0);
  cudaError_t e = launch_c<C, MIXED>(tA, tB, a, g, ra, n_ctas, tpf, device, s);
  if (e) return e;
  *consumed = n_tiles * Cfg::F;
  return cudaSuccess;
}

}  // namespace

int wtile_frames(int variant) { return variant == 1 ? 256 : 320; }

bool wtile_path_usable(const LoudParams& p, const LoudState& st, const float* d_in, size_t frames,
                       size_t in_stride_frames, int variant) {
  if (p.channels < 1 || p.channels > 2) return false;
  if (st.ring) return false;  // the ring of y is written by the generic / scan kernels
  const size_t F = (size_t)wtile_frames(variant);
  if (p.s100 < F || frames < F) return false;
  if (((uintptr_t)d_in & 15) != 0) return false;
  if ((in_stride_frames * p.channels * sizeof(float)) % 16 != 0) return false;
  if (st.n_streams > 0x7fffffffu) return false;
  return tma_encode_fn() != nullptr;
}

// Filters the leading floor(frames / F) * F frames of every stream (F = 320, or 256 for variant 1) and, when `ra` is
// given and the whole chunk was consumed, gates the completed buckets and writes the result rows in the same launch.
cudaError_t launch_loudness_wtile(const LoudParams& p, const LoudState& st, const GateParams& gp, const float* d_in,
                                  size_t frames, size_t in_stride_frames, uint32_t pos0, uint64_t bucket0, int variant,
                                  const ResultsArgs* ra, int sm_count, int device, cudaStream_t s, uint64_t* launches,
                                  size_t* consumed, bool* results_written) {
  *consumed = 0;
  if (results_written) *results_written = false;
  const int C = p.channels;
  const size_t F = (size_t)wtile_frames(variant);
  const bool fuse = ra != nullptr && frames % F == 0;
  WArgs a{};
  for (int i = 0; i < 5; i++) {
    a.na[i] = -p.a[i];
    // b[i] / b[0] - a[i] rounded ONCE: the taps nearly cancel against states ~1e6 times the output (the high-pass
    // poles), so a tap that is off by the 4e-16 a plain double division leaves shifts a block energy by ~1e-8 LU at
    // 96 kHz; with the quotient's remainder carried along the scaled tap agrees with the reference's five-tap form to
    // the recursion's own rounding noise (1e-11 LU at 48 kHz)
    const double qh = p.b[i] / p.b[0];
    const double ql = fma(-qh, p.b[0], p.b[i]) / p.b[0];
    const double sm = qh - p.a[i];
    const double bb = sm - qh;
    const double er = (qh - (sm - bb)) + (-p.a[i] - bb);
    a.cy[i] = sm + (er + ql);
  }
  a.b0sq = p.b[0] * p.b[0];
  a.filt = st.filt;
  a.bucket = st.bucket;
  a.speak = st.speak;
  a.tpeak = st.tpeak;
  a.tphist = st.tphist;
  memcpy(a.tp4, p.tp4, sizeof(a.tp4));
  memcpy(a.tp2, p.tp2, sizeof(a.tp2));
  for (int t = 0; t < 12; t++) a.tp12[t] = make_float2(p.tp4[0][t], p.tp4[1][t]);
  a.active_mask = p.do_filter ? p.active_mask : 0;
  a.n_streams = (unsigned)st.n_streams;
  a.s100 = p.s100;
  a.pos0 = pos0;
  a.slot0 = (unsigned)(bucket0 % kNB);
  a.do_sample_peak = p.do_sample_peak;
  a.fused_results = fuse ? 1 : 0;
  ResultsArgs none{};
  const ResultsArgs& r = fuse ? *ra : none;
  cudaError_t e;
  if (variant == 1)
    e = C == 1 ? launch_shape<1, false>(p, st, d_in, frames, in_stride_frames, a, gp, r, sm_count, device, s, consumed)
               : launch_shape<2, false>(p, st, d_in, frames, in_stride_frames, a, gp, r, sm_count, device, s, consumed);
  else
    e = C == 1 ? launch_shape<1, true>(p, st, d_in, frames, in_stride_frames, a, gp, r, sm_count, device, s, consumed)
               : launch_shape<2, true>(p, st, d_in, frames, in_stride_frames, a, gp, r, sm_count, device, s, consumed);
  if (e) return e;
  if (*consumed) {
    if (launches) ++*launches;
    if (results_written) *results_written = fuse;
  }
  return cudaSuccess;
}

}  // namespace ssb
