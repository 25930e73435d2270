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
//   * bucket sums are reduced across segments in a fixed order (deterministic, no atomics).
#include <cuda.h>
#include <math.h>
#include <string.h>

#include "ssb_internal.cuh"

namespace ssb {

namespace {

constexpr int kRows = 32;       // streams per CTA tile (multiple of 8: swizzle key = row & 7)
constexpr int kStages = 3;

struct TileArgs {
  double a[5];
  double b[5];
  double P[16];        // D * A^Ls * D (state hand-off matrix in difference coordinates), row-major
  const float* in;     // unused by the TMA path (kept for debugging)
  double* filt;        // [n][C][4]
  double* bucket;      // [n][C][kNB]
  float* speak;        // [n][C]
  uint64_t active_mask;
  unsigned n_streams;
  unsigned n_tiles;    // tiles of F frames in this launch
  unsigned s100;
  unsigned pos0;       // frames already in the bucket in progress
  unsigned slot0;      // its ring slot
  int do_sample_peak;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned mbar_try(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// bounded spin: a broken pipeline traps instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned spins = 0;
  while (!mbar_try(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void bar_sync_compute(int n) { asm volatile("bar.sync 1, %0;" ::"r"(n) : "memory"); }

// f32 -> f64 widening with integer ops (exact for normal numbers and zero); subnormal / inf / nan
// inputs take the hardware conversion.
__device__ __forceinline__ double widen(float x) {
  const unsigned u = __float_as_uint(x);
  const unsigned e = u & 0x7f800000u;
  if (__builtin_expect(e == 0u || e == 0x7f800000u, 0)) return (double)x;
  const unsigned hi = (u & 0x80000000u) | (((u & 0x7fffffffu) >> 3) + 0x38000000u);
  return __hiloint2double((int)hi, (int)(u << 29));
}

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

// C channels (1 or 2), T time segments per tile, F frames per tile
template <int C, int T, int F, bool WIDEN_INT>
__global__ void __launch_bounds__(kRows* C* T + 32, 1)
k_loudness_tile(const __grid_constant__ CUtensorMap tmap, const TileArgs a) {
  constexpr int NC = kRows * C * T;          // compute threads
  constexpr int RC = kRows * C;              // chains per CTA
  constexpr int LS = F / T;                  // frames per segment
  constexpr int CHUNKS = F * C / 32;         // 128-byte lines per row per stage
  constexpr int SEG_CHUNKS = LS * C / 32;    // lines per segment
  constexpr int FPQ = 4 / C;                 // frames per 16-byte quad
  constexpr unsigned STAGE_BYTES = CHUNKS * kRows * 128;
  static_assert(LS * C % 32 == 0, "segment must be whole 128-byte lines");

  extern __shared__ unsigned char smem_raw[];
  // SWIZZLE_128B needs 1024-byte aligned stage bases (1 KB of slack is included in the launch size)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* stages = smem;                                                 // kStages * STAGE_BYTES
  double* zbuf = reinterpret_cast<double*>(smem + kStages * STAGE_BYTES);       // [2][T][RC][4]
  double* carry = zbuf + 2 * T * RC * 4;                                        // [2][RC][8]: raw v1..v4 | D v
  double* part = carry + 2 * RC * 8;                                            // [2][T][RC][2]
  float* pk = reinterpret_cast<float*>(part + 2 * T * RC * 2);                  // [T][RC]
  uint64_t* full = reinterpret_cast<uint64_t*>(pk + T * RC);                    // [kStages]
  uint64_t* empty = full + kStages;                                             // [kStages]

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const unsigned n_ctas = gridDim.x;
  const unsigned row0 = (unsigned)(((unsigned long long)blockIdx.x * a.n_streams) / n_ctas);
  const unsigned row1 = (unsigned)(((unsigned long long)(blockIdx.x + 1) * a.n_streams) / n_ctas);
  const unsigned nrows = row1 - row0;  // <= kRows by construction of the grid

  if (tid == 0) {
    for (int s = 0; s < kStages; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NC / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp == NC / 32) {
    // ---------------- producer warp: one elected lane drives TMA ----------------
    if (lane == 0) {
      for (unsigned tile = 0; tile < a.n_tiles; tile++) {
        const unsigned s = tile % kStages;
        if (tile >= (unsigned)kStages) mbar_wait(&empty[s], ((tile / kStages) - 1) & 1);
        mbar_expect_tx(&full[s], STAGE_BYTES);
        tma_load_3d(stages + (size_t)s * STAGE_BYTES, &tmap, &full[s], 0, (int)row0, (int)(tile * CHUNKS));
      }
    }
    return;
  }

  // ---------------- compute threads: (segment k, row r, channel c) ----------------
  const int k = tid / RC;
  const int rc = tid - k * RC;
  const int r = rc / C;
  const int c = rc - r * C;
  const bool row_ok = (unsigned)r < nrows;
  const bool chan_ok = (a.active_mask >> c) & 1ull;
  const bool live = row_ok && chan_ok;
  const size_t gidx = ((size_t)(row0 + r)) * C + c;
  const unsigned key = r & 7;

  // state carried across tiles lives in carry[parity][rc][4]; thread k==0 seeds it
  if (k == 0) {
    double4 s0 = make_double4(0, 0, 0, 0);
    if (live) {
      const double* f = a.filt + gidx * 4;
      s0 = make_double4(f[0], f[1], f[2], f[3]);
    }
    store_carry(carry + (size_t)rc * 8, s0.x, s0.y, s0.z, s0.w);
  }
  double acc_cur = 0.0;   // k == 0 only: running sum of the bucket in progress
  unsigned slot = a.slot0;
  if (k == 0 && live && a.pos0 > 0) acc_cur = a.bucket[gidx * kNB + slot];
  float sp = 0.f;
  double v1 = 0, v2 = 0, v3 = 0, v4 = 0;
  if (T == 1 && live) {
    const double* f = a.filt + gidx * 4;
    v1 = f[0]; v2 = f[1]; v3 = f[2]; v4 = f[3];
  }
  const double a1 = a.a[1], a2 = a.a[2], a3 = a.a[3], a4 = a.a[4];
  const double b0 = a.b[0], b1 = a.b[1], b2 = a.b[2], b3 = a.b[3], b4 = a.b[4];
  unsigned pos_tile = a.pos0;  // position of the tile start inside the bucket in progress

  for (unsigned tile = 0; tile < a.n_tiles; tile++) {
    const unsigned s = tile % kStages;
    const unsigned par = tile & 1;
    mbar_wait(&full[s], (tile / kStages) & 1);
    const unsigned char* line0 = stages + (size_t)s * STAGE_BYTES + ((size_t)(k * SEG_CHUNKS) * kRows + r) * 128;

    if (T > 1) {
      // ---- pass 1: zero-state recursion over my segment -> z ----
      double z1 = 0, z2 = 0, z3 = 0, z4 = 0;
#pragma unroll 1
      for (int ch = 0; ch < SEG_CHUNKS; ch++) {
        const unsigned char* line = line0 + (size_t)ch * kRows * 128;
#pragma unroll
        for (int qi = 0; qi < 8; qi++) {
          const float4 q = *reinterpret_cast<const float4*>(line + ((qi ^ key) << 4));
#pragma unroll
          for (int f = 0; f < FPQ; f++) {
            const float xf = pick<C>(q, f, c);
            const double x = WIDEN_INT ? widen(xf) : (double)xf;
            double t = fma(-a4, z4, x);
            t = fma(-a3, z3, t);
            t = fma(-a2, z2, t);
            const double z0 = fma(-a1, z1, t);
            z4 = z3; z3 = z2; z2 = z1; z1 = z0;
          }
        }
      }
      {
        // hand-off runs in difference coordinates d = D v (see to_diff): the DF-II state is four consecutive
        // samples of a large, smooth internal signal, and A^Ls applied to it directly cancels ~1e3-fold
        double* zp = zbuf + (((size_t)par * T + k) * RC + rc) * 4;
        double d0, d1, d2, d3;
        to_diff(z1, z2, z3, z4, d0, d1, d2, d3);
        zp[0] = d0; zp[1] = d1; zp[2] = d2; zp[3] = d3;
      }
      bar_sync_compute(NC);
      // ---- combine: true incoming state of segment k ----
      const double* cr = carry + ((size_t)par * RC + rc) * 8;
      if (k == 0) {
        v1 = cr[0]; v2 = cr[1]; v3 = cr[2]; v4 = cr[3];   // exact hand-over from the previous tile
      } else {
        double d0 = cr[4], d1 = cr[5], d2 = cr[6], d3 = cr[7];
        for (int j = 0; j < k; j++) {
          const double* zj = zbuf + (((size_t)par * T + j) * RC + rc) * 4;
          const double n0 = fma(a.P[0], d0, fma(a.P[1], d1, fma(a.P[2], d2, fma(a.P[3], d3, zj[0]))));
          const double n1 = fma(a.P[4], d0, fma(a.P[5], d1, fma(a.P[6], d2, fma(a.P[7], d3, zj[1]))));
          const double n2 = fma(a.P[8], d0, fma(a.P[9], d1, fma(a.P[10], d2, fma(a.P[11], d3, zj[2]))));
          const double n3 = fma(a.P[12], d0, fma(a.P[13], d1, fma(a.P[14], d2, fma(a.P[15], d3, zj[3]))));
          d0 = n0; d1 = n1; d2 = n2; d3 = n3;
        }
        from_diff(d0, d1, d2, d3, v1, v2, v3, v4);
      }
      // ---- k == 0 folds the previous tile's partial sums into the bucket accumulator (fixed order) ----
      if (k == 0 && tile > 0) {
        const unsigned prev_pos = pos_tile >= (unsigned)F ? pos_tile - F : pos_tile + a.s100 - F;
        const bool had_boundary = prev_pos + F >= a.s100;
        const double* pp = part + ((size_t)(par ^ 1) * T * RC + rc) * 2;
        double sa = 0.0, sb = 0.0;
#pragma unroll
        for (int j = 0; j < T; j++) { sa += pp[(size_t)j * RC * 2]; sb += pp[(size_t)j * RC * 2 + 1]; }
        acc_cur += sa;
        if (had_boundary) {
          if (live) a.bucket[gidx * kNB + slot] = acc_cur;
          acc_cur = sb;
          slot = (slot + 1) % kNB;
        }
      }
    }

    // ---- pass 2: full filter from the true state; y^2 split at the bucket boundary ----
    // frames of this tile before the boundary of the bucket in progress (>= F: no boundary in this tile)
    const unsigned to_boundary = a.s100 - pos_tile;
    int lb = (int)to_boundary - k * LS;            // my samples [0, lb) belong to the current bucket
    lb = lb < 0 ? 0 : (lb > LS ? LS : lb);
    double accA = 0.0, accB = 0.0;
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
            const double x = WIDEN_INT ? widen(xf) : (double)xf;
            double t = fma(-a4, v4, x);
            t = fma(-a3, v3, t);
            t = fma(-a2, v2, t);
            const double v0 = fma(-a1, v1, t);
            double y = b4 * v4;
            y = fma(b3, v3, y);
            y = fma(b2, v2, y);
            y = fma(b1, v1, y);
            y = fma(b0, v0, y);
            v4 = v3; v3 = v2; v2 = v1; v1 = v0;
            acc = fma(y, y, acc);
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
            const double x = WIDEN_INT ? widen(xf) : (double)xf;
            double t = fma(-a4, v4, x);
            t = fma(-a3, v3, t);
            t = fma(-a2, v2, t);
            const double v0 = fma(-a1, v1, t);
            double y = b4 * v4;
            y = fma(b3, v3, y);
            y = fma(b2, v2, y);
            y = fma(b1, v1, y);
            y = fma(b0, v0, y);
            v4 = v3; v3 = v2; v2 = v1; v1 = v0;
            if (i < lb) accA = fma(y, y, accA); else accB = fma(y, y, accB);
          }
        }
      }
    }
    // this stage's shared memory is no longer needed by this warp
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);

    if (T > 1) {
      double* pp = part + (((size_t)par * T + k) * RC + rc) * 2;
      pp[0] = live ? accA : 0.0;
      pp[1] = live ? accB : 0.0;
      if (k == T - 1) store_carry(carry + ((size_t)(par ^ 1) * RC + rc) * 8, v1, v2, v3, v4);
    } else {
      acc_cur += accA;
      if (to_boundary <= (unsigned)F) {
        if (live) a.bucket[gidx * kNB + slot] = acc_cur;
        acc_cur = accB;
        slot = (slot + 1) % kNB;
      }
    }
    pos_tile += F;
    if (pos_tile >= a.s100) pos_tile -= a.s100;
  }

  // ---------------- epilogue: last tile's partials, state, peaks ----------------
  if (T > 1) {
    if (a.do_sample_peak) pk[(size_t)k * RC + rc] = sp;
    bar_sync_compute(NC);
    if (k == 0) {
      const unsigned par = a.n_tiles & 1;  // parity the next tile would have had
      if (a.n_tiles > 0) {
        const unsigned prev_pos = pos_tile >= (unsigned)F ? pos_tile - F : pos_tile + a.s100 - F;
        const bool had_boundary = prev_pos + F >= a.s100;
        const double* pp = part + ((size_t)(par ^ 1) * T * RC + rc) * 2;
        double sa = 0.0, sb = 0.0;
#pragma unroll
        for (int j = 0; j < T; j++) { sa += pp[(size_t)j * RC * 2]; sb += pp[(size_t)j * RC * 2 + 1]; }
        acc_cur += sa;
        if (had_boundary) {
          if (live) a.bucket[gidx * kNB + slot] = acc_cur;
          acc_cur = sb;
          slot = (slot + 1) % kNB;
        }
      }
      const double* cr = carry + ((size_t)par * RC + rc) * 8;
      v1 = cr[0]; v2 = cr[1]; v3 = cr[2]; v4 = cr[3];
      if (a.do_sample_peak) {
#pragma unroll
        for (int j = 1; j < T; j++) sp = fmaxf(sp, pk[(size_t)j * RC + rc]);
      }
    }
  }
  if (k == 0) {
    if (live) {
      a.bucket[gidx * kNB + slot] = acc_cur;
      double* f = a.filt + gidx * 4;
      const double tiny = 2.2250738585072014e-308;  // libebur128: flush denormal state at the end of a call
      f[0] = fabs(v1) < tiny ? 0.0 : v1;
      f[1] = fabs(v2) < tiny ? 0.0 : v2;
      f[2] = fabs(v3) < tiny ? 0.0 : v3;
      f[3] = fabs(v4) < tiny ? 0.0 : v4;
    } else if (row_ok) {
      a.bucket[gidx * kNB + slot] = 0.0;
    }
    if (row_ok && a.do_sample_peak) a.speak[gidx] = fmaxf(a.speak[gidx], sp);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

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

template <int C, int T, int F>
size_t tile_smem_bytes() {
  constexpr int RC = kRows * C;
  const size_t stage = (size_t)(F * C / 32) * kRows * 128;
  return kStages * stage + (size_t)(2 * T * RC * 4 + 2 * RC * 8 + 2 * T * RC * 2) * sizeof(double) +
         (size_t)T * RC * sizeof(float) + 2 * kStages * sizeof(uint64_t) + 1024;
}

template <int C, int T, int F>
cudaError_t launch_tile_cfg(const CUtensorMap& tmap, const TileArgs& args, unsigned n_ctas, cudaStream_t s) {
  auto kern = k_loudness_tile<C, T, F, false>;  // hardware F2F: the integer widening costs more issue slots (tools/microbench)
  const size_t smem = tile_smem_bytes<C, T, F>();
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e) return e;
  kern<<<n_ctas, kRows * C * T + 32, smem, s>>>(tmap, args);
  return cudaGetLastError();
}

}  // namespace

constexpr int kTileF = 256;
constexpr int kTileT = 4;

bool tile_path_usable(const LoudParams& p, const LoudState& st, const float* d_in, size_t frames,
                      size_t in_stride_frames) {
  if (p.channels != 1 && p.channels != 2) return false;
  if (st.ring) return false;                      // the ring of y is written by the generic kernel
  if (p.do_true_peak && p.tp_factor) return false;  // true-peak FIR not in this kernel yet
  if (p.s100 < (unsigned)kTileF) return false;
  if (frames < (size_t)kTileF) return false;
  if (((uintptr_t)d_in & 15) != 0) return false;
  if ((in_stride_frames * p.channels * sizeof(float)) % 16 != 0) return false;
  if (st.n_streams > 0x7fffffffu) return false;
  return encode_fn() != nullptr;
}

// Filters the first floor(frames / F) * F frames; returns how many frames were consumed in *consumed.
cudaError_t launch_loudness_tile(const LoudParams& p, const LoudState& st, const float* d_in, size_t frames,
                                 size_t in_stride_frames, uint32_t pos0, uint64_t bucket0, cudaStream_t s,
                                 uint64_t* launches, size_t* consumed) {
  *consumed = 0;
  const int C = p.channels;
  const size_t n_tiles = frames / kTileF;
  if (!n_tiles) return cudaSuccess;
  const size_t row_floats = in_stride_frames * C;
  const size_t used_floats = n_tiles * kTileF * C;  // multiple of 32
  CUtensorMap tmap;
  // 3-D view of the [stream][frame][channel] input: dim0 = 32 floats of one 128-byte line,
  // dim1 = stream (row pitch), dim2 = line index along the row
  cuuint64_t gdim[3] = {32, (cuuint64_t)st.n_streams, (cuuint64_t)(used_floats / 32)};
  cuuint64_t gstride[2] = {(cuuint64_t)(row_floats * sizeof(float)), 128};
  cuuint32_t box[3] = {32, (cuuint32_t)kRows, (cuuint32_t)(kTileF * C / 32)};
  cuuint32_t estride[3] = {1, 1, 1};
  CUresult cr = encode_fn()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(d_in), gdim, gstride, box,
                            estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return cudaErrorInvalidValue;

  TileArgs a;
  memcpy(a.a, p.a, sizeof(a.a));
  memcpy(a.b, p.b, sizeof(a.b));
  handoff_matrix(p.a, kTileF / kTileT, a.P);
  a.in = d_in;
  a.filt = st.filt;
  a.bucket = st.bucket;
  a.speak = st.speak;
  a.active_mask = p.do_filter ? p.active_mask : 0;
  a.n_streams = (unsigned)st.n_streams;
  a.n_tiles = (unsigned)n_tiles;
  a.s100 = p.s100;
  a.pos0 = pos0;
  a.slot0 = (unsigned)(bucket0 % kNB);
  a.do_sample_peak = p.do_sample_peak;

  int sms = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  size_t n_ctas = (st.n_streams + kRows - 1) / kRows;
  if (n_ctas < (size_t)sms && st.n_streams >= (size_t)sms) n_ctas = sms;  // spread rows over every SM
  cudaError_t e = C == 1 ? launch_tile_cfg<1, kTileT, kTileF>(tmap, a, (unsigned)n_ctas, s)
                         : launch_tile_cfg<2, kTileT, kTileF>(tmap, a, (unsigned)n_ctas, s);
  if (e) return e;
  if (launches) ++*launches;
  *consumed = n_tiles * kTileF;
  return cudaSuccess;
}

}  // namespace ssb
