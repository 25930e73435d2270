// loudness_tile.cu — the batch K-weighting + gated-RMS kernel for sm_100a (BASELINE config 2/4).
//
// Same arithmetic as k_loudness_generic (reference src/analyzer.rs:139-141 -> ebur128 add_frames_f32:
// 4th-order DF-II K-weighting in f64, y^2 summed into 100 ms buckets, sample peak), reorganised for the
// machine:
//   * input tiles [32 streams x F frames] are staged HBM -> shared memory by TMA (one 3-D
//     cp.async.bulk.tensor per stage, SWIZZLE_128B, mbarrier full/empty ring) by a producer warp, so every
//     HBM sector is fetched once, fully used, and the compute warps issue no global loads;
//   * each compute thread owns one (stream, channel, time-segment): lanes of a quarter-warp read different
//     streams' 128-byte lines, which the swizzle spreads over distinct banks (conflict-free LDS.128);
//   * the IIR is serial in time, so with few streams the FP64 pipes would idle.  T time-segments per
//     tile run concurrently: pass 1 runs the recursion from zero state (4 DFMA/sample), the true incoming
//     state of segment k is s_k = P s_{k-1} + z_{k-1} with P = A^Ls (exact linear algebra, host-computed in
//     extended precision), pass 2 runs the full filter (10 DFMA/sample) from s_k.  Both passes read the same
//     shared-memory tile, so HBM traffic stays at the algorithmic 4 B/sample.
//   * the four segments of a chain live in one warp: hand-off, carried state and bucket partial sums move by
//     warp shuffles (fixed reduction order, deterministic, no atomics, no CTA barrier in the main loop).
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "ssb_internal.cuh"
#include "tma_ptx.cuh"

namespace ssb {

namespace {

constexpr int kRows = 32;       // streams per CTA tile (multiple of 8: swizzle key = row & 7)
constexpr int kStages = 3;

struct TileArgs {
  double na[5];        // negated feedback coefficients -a[i]
  double b[5];
  double P[16];        // D * A^Ls * D (state hand-off matrix in difference coordinates), row-major
  const float* in;     // unused by the TMA path (kept for debugging)
  double* filt;        // [n][C][4]
  double* bucket;      // [n][C][kNB]
  float* speak;        // [n][C]
  float* tpeak;        // [n][C]
  float* tphist;       // [n][C][kTpHist]: x[n-1-t]
  float tp4[3][12];    // factor-4 interpolator phases 1..3 (tap t multiplies x[n-t])
  float tp2[24];       // factor-2 interpolator phase 1
  float2 tp12[12];     // (phase 1, phase 2) taps of the factor-4 interpolator, paired for FFMA2
  uint64_t active_mask;
  unsigned n_streams;
  unsigned n_tiles;    // tiles of F frames in this launch
  unsigned s100;
  unsigned pos0;       // frames already in the bucket in progress
  unsigned slot0;      // its ring slot
  int do_sample_peak;
};

__device__ __forceinline__ void bar_sync_compute(int n) { asm volatile("bar.sync 1, %0;" ::"r"(n) : "memory"); }

// d = D v with D = [[1,0,0,0],[1,-1,0,0],[1,-2,1,0],[1,-3,3,-1]] (finite differences; D is its own inverse).
// Differences of neighbouring state samples are (nearly) exact in floating point, and the hand-off matrix
// D A^Ls D acts on the small differences instead of cancelling four huge, nearly equal terms.
__device__ __forceinline__ void to_diff(double v1, double v2, double v3, double v4, double& d0, double& d1,
                                        double& d2, double& d3) {
  const double e1 = v1 - v2, e2 = v2 - v3, e3 = v3 - v4;
  d0 = v1;
  d1 = e1;
  d2 = e1 - e2;
  d3 = (e1 - e2) - (e2 - e3);
}
__device__ __forceinline__ void from_diff(double d0, double d1, double d2, double d3, double& v1, double& v2,
                                          double& v3, double& v4) {
  const double e2 = d1 - d2;
  const double e3 = e2 - (d2 - d3);
  v1 = d0;
  v2 = d0 - d1;
  v3 = v2 - e2;
  v4 = v3 - e3;
}
__device__ __forceinline__ void store_carry(double* cr, double v1, double v2, double v3, double v4) {
  cr[0] = v1; cr[1] = v2; cr[2] = v3; cr[3] = v4;
  to_diff(v1, v2, v3, v4, cr[4], cr[5], cr[6], cr[7]);
}

template <int C>
__device__ __forceinline__ float pick(const float4& q, int f, int c) {
  // element (frame f of the quad, channel c); C frames-per-quad = 4 / C
  if (C == 1) return f == 0 ? q.x : (f == 1 ? q.y : (f == 2 ? q.z : q.w));
  // C == 2
  return f == 0 ? (c ? q.y : q.x) : (c ? q.w : q.z);
}

// f32 -> f64 widening with integer ops (exact for normal numbers and zero; subnormal / inf / nan inputs take
// the hardware conversion).  F2F.F64.F32 issues at 16 lanes/clk/SM; this trades it for ALU-pipe work.
__device__ __forceinline__ double widen_int(float x) {
  const unsigned u = __float_as_uint(x);
  const unsigned e = u & 0x7f800000u;
  if (__builtin_expect((e == 0u && (u << 1) != 0u) || e == 0x7f800000u, 0)) return (double)x;
  const unsigned hi = e ? ((u & 0x80000000u) | (((u & 0x7fffffffu) >> 3) + 0x38000000u)) : (u & 0x80000000u);
  return __hiloint2double((int)hi, (int)(u << 29));
}
#if defined(SSB_EXPERIMENT_NOF2F)  // timing experiment only (wrong numerics): no conversion at all
#define SSB_CVT(x) __hiloint2double(__float_as_int(x), 0)
#elif defined(SSB_WIDEN_INT)
#define SSB_CVT(x) widen_int(x)
#else
#define SSB_CVT(x) ((double)(x))
#endif

// One filter sample: DF-II recursion + output taps; the newest state enters last (short dependent chain).
#define SSB_FILTER_STEP(x)                       \
  double t_ = fma(a.na[4], v4, (x));             \
  t_ = fma(a.na[3], v3, t_);                     \
  t_ = fma(a.na[2], v2, t_);                     \
  const double v0_ = fma(a.na[1], v1, t_);       \
  double y_ = a.b[4] * v4;                       \
  y_ = fma(a.b[3], v3, y_);                      \
  y_ = fma(a.b[2], v2, y_);                      \
  y_ = fma(a.b[1], v1, y_);                      \
  y_ = fma(a.b[0], v0_, y_);                     \
  v4 = v3; v3 = v2; v2 = v1; v1 = v0_;

// pass 1 of the time-segmented kernel: zero-state recursion over one lane's segment (4 DFMA / sample)
template <int C, int SEG_CHUNKS>
__device__ __forceinline__ void pass1_segment(const unsigned char* line0, unsigned key, int c, const TileArgs& a,
                                              double& z1, double& z2, double& z3, double& z4) {
  constexpr int FPQ = 4 / C;
  z1 = z2 = z3 = z4 = 0.0;
#pragma unroll 1
  for (int ch = 0; ch < SEG_CHUNKS; ch++) {
    const unsigned char* line = line0 + (size_t)ch * kRows * 128;
#pragma unroll
    for (int qi = 0; qi < 8; qi++) {
      const float4 q = *reinterpret_cast<const float4*>(line + ((qi ^ key) << 4));
#pragma unroll
      for (int f = 0; f < FPQ; f++) {
        const double x = SSB_CVT(pick<C>(q, f, c));
        double t = fma(a.na[4], z4, x);
        t = fma(a.na[3], z3, t);
        t = fma(a.na[2], z2, t);
        const double z0 = fma(a.na[1], z1, t);
        z4 = z3; z3 = z2; z2 = z1; z1 = z0;
      }
    }
  }
}

// One true-peak sample: ebur128's polyphase interpolator as f32 FMAs over the register window w[t] = x[n-1-t]
// (same tap order in every kernel, so all kernels report identical true peaks).
#define SSB_TP_STEP(xf)                                                         \
  if (TPF == 4) {                                                               \
    /* phases 1 and 2 as one packed FFMA2 per tap (two IEEE FMAs, same order as the scalar kernels), phase 3 scalar */ \
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

// C channels (1 or 2); T = 4 time segments per tile of F frames.
// Warp w owns rows [w*RW, (w+1)*RW) of the CTA's 32-row box, RW = 8 / C; lane = k*8 + rr*C + c holds
// (segment k, row rr, channel c).  All four segments of a chain sit in one warp, so the segment hand-off,
// the carried state and the bucket partial sums travel by warp shuffles: compute warps never meet at a CTA
// barrier and synchronise only through the TMA full/empty mbarriers.
template <int C, int F, int TPF>
__global__ void __launch_bounds__(kRows* C * 4 + 32, 1)
k_loudness_tile(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TileArgs a) {
  constexpr int T = 4;
  constexpr int RW = 8 / C;                  // rows per warp
  constexpr int NW = kRows / RW;             // compute warps
  constexpr int LS = F / T;                  // frames per segment
  constexpr int CHUNKS = F * C / 32;         // 128-byte lines per row per stage
  constexpr int SEG_CHUNKS = LS * C / 32;    // lines per segment
  constexpr int FPQ = 4 / C;                 // frames per 16-byte quad
  constexpr unsigned STAGE_BYTES = CHUNKS * kRows * 128;
  static_assert(LS * C % 32 == 0, "segment must be whole 128-byte lines");

  extern __shared__ unsigned char smem_raw[];
  // SWIZZLE_128B needs 1024-byte aligned stage bases (1 KB of slack is included in the launch size)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* stages = smem;                                                   // kStages * STAGE_BYTES
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * STAGE_BYTES);     // [kStages]
  uint64_t* empty = full + kStages;                                               // [kStages]

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const unsigned n_ctas = gridDim.x;
  const unsigned row0 = (unsigned)(((unsigned long long)blockIdx.x * a.n_streams) / n_ctas);
  const unsigned row1 = (unsigned)(((unsigned long long)(blockIdx.x + 1) * a.n_streams) / n_ctas);
  const unsigned nrows = row1 - row0;                        // <= kRows by construction of the grid
  const unsigned live_warps = (nrows + RW - 1) / RW;         // warps that own at least one stream

  if (tid == 0) {
    for (int s = 0; s < kStages; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], live_warps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

#ifndef SSB_PRODUCER_IDLE_WARP
#define SSB_PRODUCER_IDLE_WARP 1
#endif
  // The producer is the first warp without streams: with 27-28 streams per CTA (cfg2) that is warp 7, which shares its
  // SM sub-partition with one compute warp only; the extra warp 8 would sit on sub-partition 0 beside two of them.
  const int producer_warp = (SSB_PRODUCER_IDLE_WARP && live_warps < (unsigned)NW) ? (int)live_warps : NW;
  if (warp == producer_warp) {
    // ---------------- producer warp: one elected lane drives TMA ----------------
    if (lane == 0 && live_warps > 0) {
      for (unsigned tile = 0; tile < a.n_tiles; tile++) {
        const unsigned s = tile % kStages;
        if (tile >= (unsigned)kStages) mbar_wait(&empty[s], ((tile / kStages) - 1) & 1);
        mbar_expect_tx(&full[s], STAGE_BYTES);
        tma_load_3d(stages + (size_t)s * STAGE_BYTES, &tmap, &full[s], 0, (int)row0, (int)(tile * CHUNKS));
      }
    }
    return;
  }
  if ((unsigned)warp >= live_warps) return;  // no stream in this warp's rows

  // ---------------- compute warp: lane = (segment k, row rr, channel c) ----------------
  const int k = lane >> 3;
  const int q8 = lane & 7;
  const int rr = q8 / C;
  const int c = q8 - rr * C;
  const int r = warp * RW + rr;
  const bool row_ok = (unsigned)r < nrows;
  const bool live = row_ok && ((a.active_mask >> c) & 1ull);
  const size_t gidx = ((size_t)(row0 + r)) * C + c;
  const unsigned key = r & 7;

  // c1..c4: the chain's true state at the start of the tile, replicated in its four lanes
  double c1 = 0, c2 = 0, c3 = 0, c4 = 0;
  if (live) {
    const double* f = a.filt + gidx * 4;
    c1 = f[0]; c2 = f[1]; c3 = f[2]; c4 = f[3];
  }
  double acc_cur = 0.0;  // lane k == 0: running sum of the bucket in progress
  unsigned slot = a.slot0;
  if (k == 0 && live && a.pos0 > 0) acc_cur = a.bucket[gidx * kNB + slot];
  float sp = 0.f, tp = 0.f;
  constexpr int TPW = TPF == 4 ? 11 : (TPF == 2 ? 23 : 1);  // true-peak FIR history length
  float hist[TPW];  // lane k == 0: the TPW samples before the tile, hist[t] = x[n-1-t]
#pragma unroll
  for (int t = 0; t < TPW; t++) hist[t] = (TPF != 0 && row_ok) ? a.tphist[gidx * kTpHist + t] : 0.f;
  unsigned pos_tile = a.pos0;  // position of the tile start inside the bucket in progress

  // Software pipeline over tiles: while a lane runs pass 2 on tile t (throughput-bound, 10 DFMA / sample) it
  // also runs pass 1 on tile t+1 (one dependent DFMA chain) in the same loop, so every warp carries two
  // independent dependency chains and the zero-state pass hides in pass 2's issue gaps.
  const size_t seg_off = ((size_t)(k * SEG_CHUNKS) * kRows + r) * 128;
  double z1 = 0, z2 = 0, z3 = 0, z4 = 0;
  if (a.n_tiles > 0) {
    mbar_wait_warp(&full[0], 0);
    pass1_segment<C, SEG_CHUNKS>(stages + seg_off, key, c, a, z1, z2, z3, z4);
  }
  for (unsigned tile = 0; tile < a.n_tiles; tile++) {
    const unsigned s = tile % kStages;
    const unsigned char* line0 = stages + (size_t)s * STAGE_BYTES + seg_off;
    const bool has_next = tile + 1 < a.n_tiles;
    const unsigned char* line1 = stages + (size_t)((tile + 1) % kStages) * STAGE_BYTES + seg_off;
    // ---- hand-off in difference coordinates: d_k = Pt d_{k-1} + D z_{k-1}, d_0 = D carry ----
    double v1, v2, v3, v4;
    {
      double dz0, dz1, dz2, dz3, d0, d1, d2, d3;
      to_diff(z1, z2, z3, z4, dz0, dz1, dz2, dz3);
      to_diff(c1, c2, c3, c4, d0, d1, d2, d3);
#pragma unroll
      for (int j = 0; j < T - 1; j++) {
        const int src = j * 8 + q8;
        const double zj0 = __shfl_sync(0xffffffffu, dz0, src), zj1 = __shfl_sync(0xffffffffu, dz1, src);
        const double zj2 = __shfl_sync(0xffffffffu, dz2, src), zj3 = __shfl_sync(0xffffffffu, dz3, src);
        const double n0 = fma(a.P[0], d0, fma(a.P[1], d1, fma(a.P[2], d2, fma(a.P[3], d3, zj0))));
        const double n1 = fma(a.P[4], d0, fma(a.P[5], d1, fma(a.P[6], d2, fma(a.P[7], d3, zj1))));
        const double n2 = fma(a.P[8], d0, fma(a.P[9], d1, fma(a.P[10], d2, fma(a.P[11], d3, zj2))));
        const double n3 = fma(a.P[12], d0, fma(a.P[13], d1, fma(a.P[14], d2, fma(a.P[15], d3, zj3))));
        if (j < k) { d0 = n0; d1 = n1; d2 = n2; d3 = n3; }
      }
      from_diff(d0, d1, d2, d3, v1, v2, v3, v4);
      if (k == 0) { v1 = c1; v2 = c2; v3 = c3; v4 = c4; }  // exact hand-over from the previous tile
    }

    // true-peak window: the TPW samples before my segment (previous segment's tail in the same tile, or, for
    // the first segment, the previous tile's tail carried in `hist`)
    float2 w2[TPW];  // w2[t] = (x[n-1-t], x[n-1-t]): both halves equal so a tap feeds two phases in one FFMA2
    if (TPF != 0) {
      const unsigned char* row_base = stages + (size_t)s * STAGE_BYTES + (size_t)r * 128;
#pragma unroll
      for (int t = 0; t < TPW; t++) {
        const int fr = k > 0 ? k * LS - 1 - t : 0;
        const int fi = fr * C + c;
        const float prev = *reinterpret_cast<const float*>(row_base + (size_t)(fi >> 5) * kRows * 128 +
                                                            ((((fi >> 2) & 7) ^ key) << 4) + ((fi & 3) << 2));
        const float wv = k > 0 ? prev : hist[t];
        w2[t] = make_float2(wv, wv);
      }
    }
    // ---- pass 2: full filter from the true state (10 DFMA / sample) with the true-peak FIR (36 or 24 FFMA /
    //      sample) interleaved in the same loop, so FP64 and FP32 issue slots fill each other's gaps ----
    const unsigned to_boundary = a.s100 - pos_tile;  // frames of this tile before the boundary (>= F: none)
    int lb = (int)to_boundary - k * LS;              // my samples [0, lb) belong to the bucket in progress
    lb = lb < 0 ? 0 : (lb > LS ? LS : lb);
    double accA = 0.0, accB = 0.0;
    double n1 = 0, n2 = 0, n3 = 0, n4 = 0;           // zero-state response of my segment of the NEXT tile
    if (has_next) mbar_wait_warp(&full[(tile + 1) % kStages], ((tile + 1) / kStages) & 1);
    if (has_next && to_boundary >= (unsigned)F) {
      // fused: pass 2 on tile t and pass 1 on tile t+1, sample by sample (warp-uniform branch)
      double acc = 0.0;
#pragma unroll 1
      for (int ch = 0; ch < SEG_CHUNKS; ch++) {
        const unsigned char* line = line0 + (size_t)ch * kRows * 128;
        const unsigned char* lnext = line1 + (size_t)ch * kRows * 128;
#pragma unroll
        for (int qi = 0; qi < 8; qi++) {
          const float4 q = *reinterpret_cast<const float4*>(line + ((qi ^ key) << 4));
          const float4 qn = *reinterpret_cast<const float4*>(lnext + ((qi ^ key) << 4));
#pragma unroll
          for (int f = 0; f < FPQ; f++) {
            const float xf = pick<C>(q, f, c);
            sp = fmaxf(sp, fabsf(xf));
            SSB_FILTER_STEP(SSB_CVT(xf))
            acc = fma(y_, y_, acc);
            SSB_TP_STEP(xf)
            const double xn = SSB_CVT(pick<C>(qn, f, c));
            double tn = fma(a.na[4], n4, xn);
            tn = fma(a.na[3], n3, tn);
            tn = fma(a.na[2], n2, tn);
            const double n0 = fma(a.na[1], n1, tn);
            n4 = n3; n3 = n2; n2 = n1; n1 = n0;
          }
        }
      }
      accA = acc;
    } else {
      if (lb == LS || lb == 0) {
        double acc = 0.0;
#pragma unroll 1
        for (int ch = 0; ch < SEG_CHUNKS; ch++) {
          const unsigned char* line = line0 + (size_t)ch * kRows * 128;
#pragma unroll
          for (int qi = 0; qi < 8; qi++) {
            const float4 q = *reinterpret_cast<const float4*>(line + ((qi ^ key) << 4));
#pragma unroll
            for (int f = 0; f < FPQ; f++) {
              const float xf = pick<C>(q, f, c);
              sp = fmaxf(sp, fabsf(xf));
              SSB_FILTER_STEP(SSB_CVT(xf))
              acc = fma(y_, y_, acc);
              SSB_TP_STEP(xf)
            }
          }
        }
        if (lb == LS) accA = acc; else accB = acc;
      } else {
        int i = 0;
#pragma unroll 1
        for (int ch = 0; ch < SEG_CHUNKS; ch++) {
          const unsigned char* line = line0 + (size_t)ch * kRows * 128;
#pragma unroll 1
          for (int qi = 0; qi < 8; qi++) {
            const float4 q = *reinterpret_cast<const float4*>(line + ((qi ^ key) << 4));
#pragma unroll
            for (int f = 0; f < FPQ; f++, i++) {
              const float xf = pick<C>(q, f, c);
              sp = fmaxf(sp, fabsf(xf));
              SSB_FILTER_STEP(SSB_CVT(xf))
              if (i < lb) accA = fma(y_, y_, accA); else accB = fma(y_, y_, accB);
              SSB_TP_STEP(xf)
            }
          }
        }
      }
      __syncwarp();
      if (has_next) pass1_segment<C, SEG_CHUNKS>(line1, key, c, a, n1, n2, n3, n4);
    }
    if (TPF != 0) {
      // the last segment's tail is the history of the next tile's first segment
#pragma unroll
      for (int t = 0; t < TPW; t++) hist[t] = __shfl_sync(0xffffffffu, w2[t].x, (T - 1) * 8 + q8);
    }
    // this warp is done reading the stage
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);

    // ---- the last segment's end state is the chain's state for the next tile ----
    {
      const int src = (T - 1) * 8 + q8;
      c1 = __shfl_sync(0xffffffffu, v1, src);
      c2 = __shfl_sync(0xffffffffu, v2, src);
      c3 = __shfl_sync(0xffffffffu, v3, src);
      c4 = __shfl_sync(0xffffffffu, v4, src);
    }
    // ---- bucket sums: fixed-order reduction over the four segments: (k0 + k1) + (k2 + k3) ----
    accA += __shfl_xor_sync(0xffffffffu, accA, 8);
    accB += __shfl_xor_sync(0xffffffffu, accB, 8);
    accA += __shfl_xor_sync(0xffffffffu, accA, 16);
    accB += __shfl_xor_sync(0xffffffffu, accB, 16);
    if (k == 0) {
      acc_cur += accA;
      if (to_boundary <= (unsigned)F) {
        if (live) a.bucket[gidx * kNB + slot] = acc_cur;
        acc_cur = accB;
        slot = (slot + 1) % kNB;
      }
    }
    pos_tile += F;
    if (pos_tile >= a.s100) pos_tile -= a.s100;
    z1 = n1; z2 = n2; z3 = n3; z4 = n4;
  }

  // ---------------- epilogue: state, bucket in progress, peak ----------------
  sp = fmaxf(sp, __shfl_xor_sync(0xffffffffu, sp, 8));
  sp = fmaxf(sp, __shfl_xor_sync(0xffffffffu, sp, 16));
  tp = fmaxf(tp, __shfl_xor_sync(0xffffffffu, tp, 8));
  tp = fmaxf(tp, __shfl_xor_sync(0xffffffffu, tp, 16));
  if (k == 0) {
    if (live) {
      a.bucket[gidx * kNB + slot] = acc_cur;
      double* f = a.filt + gidx * 4;
      const double tiny = 2.2250738585072014e-308;  // libebur128: flush denormal state at the end of a call
      f[0] = fabs(c1) < tiny ? 0.0 : c1;
      f[1] = fabs(c2) < tiny ? 0.0 : c2;
      f[2] = fabs(c3) < tiny ? 0.0 : c3;
      f[3] = fabs(c4) < tiny ? 0.0 : c4;
    } else if (row_ok) {
      a.bucket[gidx * kNB + slot] = 0.0;
    }
    if (row_ok && a.do_sample_peak) a.speak[gidx] = fmaxf(a.speak[gidx], sp);
    if (TPF != 0 && row_ok) {
      a.tpeak[gidx] = fmaxf(a.tpeak[gidx], tp);
#pragma unroll
      for (int t = 0; t < TPW; t++) a.tphist[gidx * kTpHist + t] = hist[t];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// k_loudness_rows: the many-streams variant (>= 16384 streams per GPU, BASELINE config 4).  With that many
// independent recursions there is no need to split time: one lane per (stream, channel) runs the reference's
// recursion serially (10 DFMA + 1 F2F per sample: no zero-state pass, no hand-off), on the same TMA /
// SWIZZLE_128B staging with [128 streams x 64 frames] tiles and a register-resident true-peak history.
// ------------------------------------------------------------------------------------------------
#ifndef SSB_SERIAL_F
#define SSB_SERIAL_F 64
#endif
#ifndef SSB_SERIAL_MINB
#define SSB_SERIAL_MINB 1
#endif
constexpr int kRowsSerial = 128;
constexpr int kSerialF = SSB_SERIAL_F;

template <int C, int TPF>
__global__ void __launch_bounds__(kRowsSerial* C + 32, SSB_SERIAL_MINB)
k_loudness_rows(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TileArgs a) {
  constexpr int F = kSerialF;
  constexpr int RW = 32 / C;                       // rows per warp
  constexpr int NW = kRowsSerial / RW;             // compute warps
  constexpr int CHUNKS = F * C / 32;               // 128-byte lines per row per stage
  constexpr int FPQ = 4 / C;
  constexpr unsigned STAGE_BYTES = CHUNKS * kRowsSerial * 128;

  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* stages = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * STAGE_BYTES);
  uint64_t* empty = full + kStages;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const unsigned row0 = blockIdx.x * (unsigned)kRowsSerial;
  const unsigned nrows = min((unsigned)kRowsSerial, a.n_streams - row0);
  const unsigned live_warps = (nrows + RW - 1) / RW;

  if (tid == 0) {
    for (int s = 0; s < kStages; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], live_warps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp == NW) {
    if (lane == 0 && live_warps > 0) {
      for (unsigned tile = 0; tile < a.n_tiles; tile++) {
        const unsigned s = tile % kStages;
        if (tile >= (unsigned)kStages) mbar_wait(&empty[s], ((tile / kStages) - 1) & 1);
        mbar_expect_tx(&full[s], STAGE_BYTES);
        tma_load_3d(stages + (size_t)s * STAGE_BYTES, &tmap, &full[s], 0, (int)row0, (int)(tile * CHUNKS));
      }
    }
    return;
  }
  if ((unsigned)warp >= live_warps) return;

  const int rr = lane / C;
  const int c = lane - rr * C;
  const int r = warp * RW + rr;
  const bool row_ok = (unsigned)r < nrows;
  const bool live = row_ok && ((a.active_mask >> c) & 1ull);
  const size_t gidx = ((size_t)(row0 + r)) * C + c;
  const unsigned key = r & 7;

  double v1 = 0, v2 = 0, v3 = 0, v4 = 0;
  if (live) {
    const double* f = a.filt + gidx * 4;
    v1 = f[0]; v2 = f[1]; v3 = f[2]; v4 = f[3];
  }
  unsigned slot = a.slot0;
  double acc = (live && a.pos0 > 0) ? a.bucket[gidx * kNB + slot] : 0.0;
  float sp = 0.f, tp = 0.f;
  constexpr int TPW = TPF == 4 ? 11 : (TPF == 2 ? 23 : 1);
  float2 w2[TPW];  // true-peak window, both halves equal (FFMA2 operand); lives in registers across tiles
#pragma unroll
  for (int t = 0; t < TPW; t++) {
    const float wv = (TPF != 0 && row_ok) ? a.tphist[gidx * kTpHist + t] : 0.f;
    w2[t] = make_float2(wv, wv);
  }
  unsigned pos = a.pos0;  // frames already in the bucket in progress

  for (unsigned tile = 0; tile < a.n_tiles; tile++) {
    const unsigned s = tile % kStages;
    mbar_wait_warp(&full[s], (tile / kStages) & 1);
    const unsigned char* line0 = stages + (size_t)s * STAGE_BYTES + (size_t)r * 128;
    const unsigned to_boundary = a.s100 - pos;   // > F: the bucket does not end inside this tile
    if (to_boundary > (unsigned)F) {
#pragma unroll 1
      for (int ch = 0; ch < CHUNKS; ch++) {
        const unsigned char* line = line0 + (size_t)ch * kRowsSerial * 128;
#pragma unroll
        for (int qi = 0; qi < 8; qi++) {
          const float4 q = *reinterpret_cast<const float4*>(line + ((qi ^ key) << 4));
#pragma unroll
          for (int f = 0; f < FPQ; f++) {
            const float xf = pick<C>(q, f, c);
            sp = fmaxf(sp, fabsf(xf));
            SSB_FILTER_STEP(SSB_CVT(xf))
            acc = fma(y_, y_, acc);
            SSB_TP_STEP(xf)
          }
        }
      }
      pos += F;
    } else {
      int i = 0;
#pragma unroll 1
      for (int ch = 0; ch < CHUNKS; ch++) {
        const unsigned char* line = line0 + (size_t)ch * kRowsSerial * 128;
#pragma unroll 1
        for (int qi = 0; qi < 8; qi++) {
          const float4 q = *reinterpret_cast<const float4*>(line + ((qi ^ key) << 4));
#pragma unroll
          for (int f = 0; f < FPQ; f++, i++) {
            const float xf = pick<C>(q, f, c);
            sp = fmaxf(sp, fabsf(xf));
            SSB_FILTER_STEP(SSB_CVT(xf))
            acc = fma(y_, y_, acc);
            SSB_TP_STEP(xf)
            if (i + 1 == (int)to_boundary) {   // the bucket in progress is complete
              if (live) a.bucket[gidx * kNB + slot] = acc;
              acc = 0.0;
              slot = (slot + 1) % kNB;
            }
          }
        }
      }
      pos = (unsigned)F - to_boundary;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  if (live) {
    a.bucket[gidx * kNB + slot] = acc;
    double* f = a.filt + gidx * 4;
    const double tiny = 2.2250738585072014e-308;
    f[0] = fabs(v1) < tiny ? 0.0 : v1;
    f[1] = fabs(v2) < tiny ? 0.0 : v2;
    f[2] = fabs(v3) < tiny ? 0.0 : v3;
    f[3] = fabs(v4) < tiny ? 0.0 : v4;
  } else if (row_ok) {
    a.bucket[gidx * kNB + slot] = 0.0;
  }
  if (row_ok && a.do_sample_peak) a.speak[gidx] = fmaxf(a.speak[gidx], sp);
  if (TPF != 0 && row_ok) {
    a.tpeak[gidx] = fmaxf(a.tpeak[gidx], tp);
#pragma unroll
    for (int t = 0; t < TPW; t++) a.tphist[gidx * kTpHist + t] = w2[t].x;
  }
}

// ------------------------------------------------------------------------------------------------
// k_loudness_rows_any: the serial kernel for any channel count 3..32 (5.1 at 96 kHz is BASELINE config 5).
// Frames of C channels do not align with 16-byte quads, so each lane reads its samples one float at a time
// from the swizzled tile; a warp holds floor(32/C) streams (lane = row*C + channel), a CTA 8 such warps, a
// tile 32 frames (32*C floats per stream = exactly C lines, so every C is supported).
// ------------------------------------------------------------------------------------------------
constexpr int kAnyF = 32;
constexpr int kAnyWarps = 8;        // compute warps per CTA without the true-peak FIR (3 CTAs per SM)
constexpr int kAnyWarpsTp = 15;     // with it: one CTA of 15 + 1 warps (512 threads) per SM, 128 registers per thread

// NS TMA stages of 32 KB, MINB CTAs per SM: without the true-peak FIR the kernel is latency-bound (one dependent
// recursion per lane), so it trades a stage for a third resident CTA (24 compute warps per SM)
template <int TPF, int NS, int MINB, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32 + 32, MINB)
k_loudness_rows_any(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TileArgs a, const int C,
                    const int ROWS) {   // ROWS: streams per TMA box (multiple of 8, <= NWARPS * RW), >= what a CTA owns
  constexpr int F = kAnyF;
  const int RW = 32 / C;                 // streams per warp
  const unsigned STAGE_BYTES = (unsigned)C * ROWS * 128u;

  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* stages = smem;
  constexpr unsigned SLOT = NWARPS * 4096u;   // bytes reserved per stage: NWARPS * RW rows x C lines x 128 B <= NWARPS * 4 KB
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NS * SLOT);
  uint64_t* empty = full + NS;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  // streams are split evenly over the grid (a multiple of the resident CTA slots): <= ROWS per CTA by construction
  const unsigned row0 = (unsigned)(((unsigned long long)blockIdx.x * a.n_streams) / gridDim.x);
  const unsigned row1 = (unsigned)(((unsigned long long)(blockIdx.x + 1) * a.n_streams) / gridDim.x);
  const unsigned nrows = row1 - row0;
  const unsigned live_warps = (nrows + RW - 1) / RW;

  if (tid == 0) {
    for (int s = 0; s < NS; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], live_warps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp == NWARPS) {
    if (lane == 0 && live_warps > 0) {
      for (unsigned tile = 0; tile < a.n_tiles; tile++) {
        const unsigned s = tile % NS;
        if (tile >= (unsigned)NS) mbar_wait(&empty[s], ((tile / NS) - 1) & 1);
        mbar_expect_tx(&full[s], STAGE_BYTES);
        tma_load_3d(stages + (size_t)s * SLOT, &tmap, &full[s], 0, (int)row0, (int)(tile * C));
      }
    }
    return;
  }
  if ((unsigned)warp >= live_warps) return;

  const int rr = lane / C;
  const int c = lane - rr * C;
  const int r = warp * RW + rr;
  const bool lane_ok = rr < RW;                     // lanes past RW*C idle
  const bool row_ok = lane_ok && (unsigned)r < nrows;
  const bool live = row_ok && ((a.active_mask >> c) & 1ull);
  const size_t gidx = ((size_t)(row0 + (lane_ok ? r : 0))) * C + (lane_ok ? c : 0);
  const unsigned key = r & 7;

  double v1 = 0, v2 = 0, v3 = 0, v4 = 0;
  if (live) {
    const double* f = a.filt + gidx * 4;
    v1 = f[0]; v2 = f[1]; v3 = f[2]; v4 = f[3];
  }
  unsigned slot = a.slot0;
  double acc = (live && a.pos0 > 0) ? a.bucket[gidx * kNB + slot] : 0.0;
  float sp = 0.f, tp = 0.f;
  constexpr int TPW = TPF == 4 ? 11 : (TPF == 2 ? 23 : 1);
  float w[TPW];
#pragma unroll
  for (int t = 0; t < TPW; t++) w[t] = (TPF != 0 && row_ok) ? a.tphist[gidx * kTpHist + t] : 0.f;
  unsigned pos = a.pos0;

  // Where this lane's 32 samples of a tile sit inside a stage (swizzled) depends on (row, channel) only: the byte
  // offsets live in registers, two 16-bit offsets per register (a stage is <= 64 KB), and the fully unrolled tile loop
  // spends one extract instead of ~8 address instructions per sample; full unrolling also turns the true-peak
  // window shift into register renaming (23 moves per tile instead of per 4 samples).
  constexpr bool kPackedOffsets = true;
  unsigned offp[F / 2];
  if (kPackedOffsets) {
#pragma unroll
    for (int j = 0; j < F / 2; j++) {
      unsigned o2[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int fi = (2 * j + h) * C + (lane_ok ? c : 0);
        o2[h] = (unsigned)(lane_ok ? r : 0) * 128u + (unsigned)(fi >> 5) * (unsigned)ROWS * 128u +
                (unsigned)((((fi >> 2) & 7) ^ key) << 4) + (unsigned)((fi & 3) << 2);
      }
      offp[j] = o2[0] | (o2[1] << 16);
    }
  }

  for (unsigned tile = 0; tile < a.n_tiles; tile++) {
    const unsigned s = tile % NS;
    mbar_wait_warp(&full[s], (tile / NS) & 1);
    const unsigned char* stage_base = stages + (size_t)s * SLOT;
    const unsigned char* row_base = stage_base + (size_t)(lane_ok ? r : 0) * 128;
    const unsigned to_boundary = a.s100 - pos;
    // one sample of this lane's channel: peaks, K-weighting step, y^2 into the bucket in progress
#define SSB_ANY_LOAD_CALC(f)                                                                                 \
    *reinterpret_cast<const float*>(row_base + (size_t)(((f) * C + (lane_ok ? c : 0)) >> 5) * ROWS * 128 +    \
                                    ((((((f) * C + (lane_ok ? c : 0)) >> 2) & 7) ^ key) << 4) +               \
                                    ((((f) * C + (lane_ok ? c : 0)) & 3) << 2))
#define SSB_ANY_LOAD_PACKED(f) \
    *reinterpret_cast<const float*>(stage_base + (((f) & 1) ? (offp[(f) >> 1] >> 16) : (offp[(f) >> 1] & 0xffffu)))
#define SSB_ANY_SAMPLE(f, LOAD)                                                                             \
    {                                                                                                       \
      const float xf = LOAD(f);                                                                             \
      sp = fmaxf(sp, fabsf(xf));                                                                            \
      if (TPF == 4) {                                                                                       \
        _Pragma("unroll") for (int ph = 0; ph < 3; ph++) {                                                  \
          float accf = xf * a.tp4[ph][0];                                                                   \
          _Pragma("unroll") for (int t = 1; t < 12; t++) accf = fmaf(w[t - 1], a.tp4[ph][t], accf);         \
          tp = fmaxf(tp, fabsf(accf));                                                                      \
        }                                                                                                   \
      } else if (TPF == 2) {                                                                                \
        float accf = xf * a.tp2[0];                                                                         \
        _Pragma("unroll") for (int t = 1; t < 24; t++) accf = fmaf(w[t - 1], a.tp2[t], accf);               \
        tp = fmaxf(tp, fabsf(accf));                                                                        \
      }                                                                                                     \
      if (TPF != 0) {                                                                                       \
        _Pragma("unroll") for (int t = TPW - 1; t > 0; t--) w[t] = w[t - 1];                                \
        w[0] = xf;                                                                                          \
      }                                                                                                     \
      SSB_FILTER_STEP(SSB_CVT(xf))                                                                          \
      acc = fma(y_, y_, acc);                                                                               \
    }
    if (to_boundary > (unsigned)F) {
      // the usual tile (299 of 300 at 96 kHz): no bucket ends inside it, the loop carries no boundary test
#pragma unroll
      for (int f = 0; f < F; f++) SSB_ANY_SAMPLE(f, SSB_ANY_LOAD_PACKED)
    } else {
#pragma unroll 1
      for (int f = 0; f < F; f++) {
        SSB_ANY_SAMPLE(f, SSB_ANY_LOAD_CALC)
        if (f + 1 == (int)to_boundary) {   // the bucket in progress is complete
          if (live) a.bucket[gidx * kNB + slot] = acc;
          acc = 0.0;
          slot = (slot + 1) % kNB;
        }
      }
    }
#undef SSB_ANY_SAMPLE
#undef SSB_ANY_LOAD_CALC
#undef SSB_ANY_LOAD_PACKED
    pos = to_boundary > (unsigned)F ? pos + F : (unsigned)F - to_boundary;
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  if (live) {
    a.bucket[gidx * kNB + slot] = acc;
    double* f = a.filt + gidx * 4;
    const double tiny = 2.2250738585072014e-308;
    f[0] = fabs(v1) < tiny ? 0.0 : v1;
    f[1] = fabs(v2) < tiny ? 0.0 : v2;
    f[2] = fabs(v3) < tiny ? 0.0 : v3;
    f[3] = fabs(v4) < tiny ? 0.0 : v4;
  } else if (row_ok) {
    a.bucket[gidx * kNB + slot] = 0.0;
  }
  if (row_ok && a.do_sample_peak) a.speak[gidx] = fmaxf(a.speak[gidx], sp);
  if (TPF != 0 && row_ok) {
    a.tpeak[gidx] = fmaxf(a.tpeak[gidx], tp);
#pragma unroll
    for (int t = 0; t < TPW; t++) a.tphist[gidx * kTpHist + t] = w[t];
  }
}
#undef SSB_FILTER_STEP
#undef SSB_TP_STEP

using EncodeTiledFn = TmaEncodeTiledFn;
EncodeTiledFn encode_fn() { return tma_encode_fn(); }

// ---- host: D * A^n * D in double-double arithmetic, rounded once to double ----------------------
struct dd { double hi, lo; };
static inline dd dd_from(double x) { return {x, 0.0}; }
static inline dd dd_renorm(double s, double e) { const double hi = s + e; return {hi, e - (hi - s)}; }
static inline dd dd_add(dd x, dd y) {
  const double s = x.hi + y.hi, bb = s - x.hi;
  double e = (x.hi - (s - bb)) + (y.hi - bb);
  e += x.lo + y.lo;
  return dd_renorm(s, e);
}
static inline dd dd_mul_d(dd x, double b) {
  const double p = x.hi * b;
  double e = fma(x.hi, b, -p);
  e = fma(x.lo, b, e);
  return dd_renorm(p, e);
}
static inline dd dd_neg(dd x) { return {-x.hi, -x.lo}; }

void handoff_matrix(const double a[5], int n, double P[16]) {
  // R = A^n by n left-multiplications with the companion matrix A (row 0 = -a1..-a4, rows 1..3 shift)
  dd R[4][4];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) R[i][j] = dd_from(i == j ? 1.0 : 0.0);
  for (int it = 0; it < n; it++) {
    dd row0[4];
    for (int j = 0; j < 4; j++) {
      dd acc = dd_from(0.0);
      for (int k = 0; k < 4; k++) acc = dd_add(acc, dd_mul_d(R[k][j], -a[k + 1]));
      row0[j] = acc;
    }
    for (int i = 3; i > 0; i--) for (int j = 0; j < 4; j++) R[i][j] = R[i - 1][j];
    for (int j = 0; j < 4; j++) R[0][j] = row0[j];
  }
  const double D[4][4] = {{1, 0, 0, 0}, {1, -1, 0, 0}, {1, -2, 1, 0}, {1, -3, 3, -1}};
  dd Tm[4][4];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      dd acc = dd_from(0.0);
      for (int k = 0; k < 4; k++) acc = dd_add(acc, dd_mul_d(R[i][k], D[k][j]));  // R * D
      Tm[i][j] = acc;
    }
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      dd acc = dd_from(0.0);
      for (int k = 0; k < 4; k++) acc = dd_add(acc, dd_mul_d(Tm[k][j], D[i][k]));  // D * (R * D)
      P[i * 4 + j] = acc.hi + acc.lo;
    }
  (void)dd_neg;
}

template <int C, int F>
size_t tile_smem_bytes() {
  const size_t stage = (size_t)(F * C / 32) * kRows * 128;
  return kStages * stage + 2 * kStages * sizeof(uint64_t) + 1024;
}

template <int C, int F, int TPF>
cudaError_t launch_tile_cfg(const CUtensorMap& tmap, const TileArgs& args, unsigned n_ctas, cudaStream_t s) {
  auto kern = k_loudness_tile<C, F, TPF>;
  const size_t smem = tile_smem_bytes<C, F>();
  static bool configured_dev[64] = {false};  // per instantiation and device
  int dev_ = 0;
  cudaGetDevice(&dev_);
  bool& configured = configured_dev[dev_ & 63];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e) return e;
    configured = true;
  }
  kern<<<n_ctas, kRows * C * 4 + 32, smem, s>>>(tmap, args);
  return cudaGetLastError();
}

template <int C>
cudaError_t launch_tile_c(const CUtensorMap& tmap, const TileArgs& args, unsigned n_ctas, int tpf, cudaStream_t s) {
  if (tpf == 4) return launch_tile_cfg<C, 256, 4>(tmap, args, n_ctas, s);
  if (tpf == 2) return launch_tile_cfg<C, 256, 2>(tmap, args, n_ctas, s);
  return launch_tile_cfg<C, 256, 0>(tmap, args, n_ctas, s);
}


template <int C, int TPF>
cudaError_t launch_rows_cfg(const CUtensorMap& tmap, const TileArgs& args, unsigned n_ctas, cudaStream_t s) {
  auto kern = k_loudness_rows<C, TPF>;
  const size_t smem = (size_t)kStages * (kSerialF * C / 32) * kRowsSerial * 128 + 2 * kStages * sizeof(uint64_t) + 1024;
  static bool configured_dev[64] = {false};
  int dev_ = 0;
  cudaGetDevice(&dev_);
  bool& configured = configured_dev[dev_ & 63];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e) return e;
    configured = true;
  }
  kern<<<n_ctas, kRowsSerial * C + 32, smem, s>>>(tmap, args);
  return cudaGetLastError();
}

template <int C>
cudaError_t launch_rows_c(const CUtensorMap& tmap, const TileArgs& args, unsigned n_ctas, int tpf, cudaStream_t s) {
  if (tpf == 4) return launch_rows_cfg<C, 4>(tmap, args, n_ctas, s);
  if (tpf == 2) return launch_rows_cfg<C, 2>(tmap, args, n_ctas, s);
  return launch_rows_cfg<C, 0>(tmap, args, n_ctas, s);
}


template <int TPF, int NS, int MINB, int NWARPS>
cudaError_t launch_any_cfg(const CUtensorMap& tmap, const TileArgs& args, unsigned n_ctas, int C, int box_rows,
                           cudaStream_t s) {
  auto kern = k_loudness_rows_any<TPF, NS, MINB, NWARPS>;
  const size_t smem = (size_t)NS * NWARPS * 4096 + 2 * NS * sizeof(uint64_t) + 1024;
  static bool configured_dev[64] = {false};
  int dev_ = 0;
  cudaGetDevice(&dev_);
  bool& configured = configured_dev[dev_ & 63];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e) return e;
    configured = true;
  }
  kern<<<n_ctas, NWARPS * 32 + 32, smem, s>>>(tmap, args, C, box_rows);
  return cudaGetLastError();
}

}  // namespace

TmaEncodeTiledFn tma_encode_fn() {
  static TmaEncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<TmaEncodeTiledFn>(p);
    return (TmaEncodeTiledFn) nullptr;
  }();
  return fn;
}

constexpr int kTileT = 4;
constexpr int kTileFMax = 256;

// stream count from which the serial many-streams kernel replaces the time-segmented ones
// (SSB_SERIAL_MIN overrides it; the tests use that to run both kernels on small batches); read once, thread-safe
size_t serial_min_streams() {
  static const size_t v = [] {
    const char* e = getenv("SSB_SERIAL_MIN");
    size_t x = e ? (size_t)strtoull(e, nullptr, 10) : 16384;
    return x ? x : (size_t)1;
  }();
  return v;
}

void tile_handoff_matrix(const double a[5], double P[16]) { handoff_matrix(a, kTileFMax / kTileT, P); }
void tile_handoff_power(const double a[5], int n, double P[16]) { handoff_matrix(a, n, P); }

bool tile_path_usable(const LoudParams& p, const LoudState& st, const float* d_in, size_t frames,
                      size_t in_stride_frames) {
  if (p.channels < 1 || p.channels > 32) return false;       // 3..32 channels: k_loudness_rows_any
  if (st.ring) return false;                      // the ring of y is written by the generic kernel
  if (p.s100 < (unsigned)kTileFMax) return false;
  if (frames < (size_t)kTileFMax) return false;
  if (((uintptr_t)d_in & 15) != 0) return false;
  if ((in_stride_frames * p.channels * sizeof(float)) % 16 != 0) return false;
  if (st.n_streams > 0x7fffffffu) return false;
  return encode_fn() != nullptr;
}

// Filters the first floor(frames / F) * F frames; returns how many frames were consumed in *consumed.
cudaError_t launch_loudness_tile(const LoudParams& p, const LoudState& st, const float* d_in, size_t frames,
                                 size_t in_stride_frames, uint32_t pos0, uint64_t bucket0, cudaStream_t s,
                                 uint64_t* launches, size_t* consumed, int force_kernel, int sm_count) {
  *consumed = 0;
  const int C = p.channels;
  // force_kernel: 0 = choose by stream count, 2 = serial rows kernel, 3 = time-segmented kernel (tests)
  const bool any_c = C > 2;
  const bool serial = any_c || force_kernel == 2 || (force_kernel != 3 && st.n_streams >= serial_min_streams());
  const int tile_f = any_c ? kAnyF : (serial ? kSerialF : kTileFMax);
  const int sms = sm_count > 0 ? sm_count : 148;   // of the handle's device (queried at create)
  const bool any_tp = p.do_true_peak && p.tp_factor;
  // multichannel kernel: a box holds whole warps' worth of streams, rounded down to the swizzle period of 8 rows
  int box_rows = any_c ? (any_tp ? kAnyWarpsTp : kAnyWarps) * (32 / C) / 8 * 8 : (serial ? kRowsSerial : kRows);
  size_t any_ctas = 0;
  if (any_c) {
    // three CTAs of 8 warps (one of 16 warps with the true-peak FIR) fit an SM: the grid is the smallest multiple of the resident slots that
    // keeps <= box_rows streams per CTA, so every SM carries the same number of streams (few streams: one CTA per
    // box), and the TMA box shrinks to what a CTA owns
    any_ctas = (st.n_streams + box_rows - 1) / box_rows;
    const size_t slots = (size_t)(any_tp ? 1 : 3) * (size_t)sms;
    if (st.n_streams >= slots) {
      any_ctas = (any_ctas + slots - 1) / slots * slots;
    }
    // SWIZZLE_128B keys on shared-memory address bits 7..9 = (line * box_rows + row) & 7: a box of a multiple of 8 rows
    // keeps that equal to row & 7 for every line
    box_rows = (int)(((st.n_streams + any_ctas - 1) / any_ctas + 7) / 8 * 8);
  }
  const size_t n_tiles = frames / tile_f;
  if (!n_tiles) return cudaSuccess;
  const size_t row_floats = in_stride_frames * C;
  const size_t used_floats = n_tiles * tile_f * C;  // multiple of 32
  CUtensorMap tmap;
  // 3-D view of the [stream][frame][channel] input: dim0 = 32 floats of one 128-byte line,
  // dim1 = stream (row pitch), dim2 = line index along the row
  cuuint64_t gdim[3] = {32, (cuuint64_t)st.n_streams, (cuuint64_t)(used_floats / 32)};
  cuuint64_t gstride[2] = {(cuuint64_t)(row_floats * sizeof(float)), 128};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, (cuuint32_t)(tile_f * C / 32)};
  cuuint32_t estride[3] = {1, 1, 1};
  CUresult cr = encode_fn()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(d_in), gdim, gstride, box,
                            estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return cudaErrorInvalidValue;

  TileArgs a{};
  for (int i = 0; i < 5; i++) a.na[i] = -p.a[i];
  memcpy(a.b, p.b, sizeof(a.b));
  if (!serial) memcpy(a.P, p.handoff, sizeof(a.P));
  a.in = d_in;
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
  a.n_tiles = (unsigned)n_tiles;
  a.s100 = p.s100;
  a.pos0 = pos0;
  a.slot0 = (unsigned)(bucket0 % kNB);
  a.do_sample_peak = p.do_sample_peak;

  const int tpf = p.do_true_peak ? p.tp_factor : 0;
  cudaError_t e;
  if (any_c) {
    const size_t n_ctas = any_ctas;
    e = tpf == 4 ? launch_any_cfg<4, 3, 1, kAnyWarpsTp>(tmap, a, (unsigned)n_ctas, C, box_rows, s)
                 : (tpf == 2 ? launch_any_cfg<2, 3, 1, kAnyWarpsTp>(tmap, a, (unsigned)n_ctas, C, box_rows, s)
                             : launch_any_cfg<0, 2, 3, kAnyWarps>(tmap, a, (unsigned)n_ctas, C, box_rows, s));
  } else if (serial) {
    const size_t n_ctas = (st.n_streams + kRowsSerial - 1) / kRowsSerial;
    e = C == 1 ? launch_rows_c<1>(tmap, a, (unsigned)n_ctas, tpf, s) : launch_rows_c<2>(tmap, a, (unsigned)n_ctas, tpf, s);
  } else {
    size_t n_ctas = (st.n_streams + kRows - 1) / kRows;
    if (n_ctas < (size_t)sms && st.n_streams >= (size_t)sms) n_ctas = sms;  // spread rows over every SM
    e = C == 1 ? launch_tile_c<1>(tmap, a, (unsigned)n_ctas, tpf, s) : launch_tile_c<2>(tmap, a, (unsigned)n_ctas, tpf, s);
  }
  if (e) return e;
  if (launches) ++*launches;
  *consumed = n_tiles * tile_f;
  return cudaSuccess;
}

}  // namespace ssb
