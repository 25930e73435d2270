// loudness_scan.cu — the few-streams / long-audio kernel (one reference `Analyzer`, BASELINE config 1 on the
// GPU; the per-tick `add_samples` of tui.rs:1528-1543 and the whole-file pass of analyzer.rs:170-182).
//
// With one stream there is nothing to parallelise but time.  One CTA per stream; each lane owns a 64-frame
// segment of one channel, so a 256-thread CTA covers 8192 stereo frames per sweep:
//   pass 1   zero-state recursion over the lane's segment (4 DFMA / sample)                      -> z_k
//   scan     the state entering every segment, d_{k+1} = Pt d_k + D z_k, as a parallel prefix over the
//            segments: Hillis-Steele with __shfl_up_sync inside a warp (matrices Pt^1, Pt^2, Pt^4, ...),
//            a short serial combine over the 8 warp totals (Pt^w), then one lane-dependent Pt^m matvec
//   pass 2   the reference's full recursion from the true state (10 DFMA / sample): y -> optional 3 s ring
//            (SSB_FLAG_RING), y^2 -> per-segment partial sums split at the 100 ms bucket boundary
//   reduce   one thread per channel folds the segment partials into the bucket ring in time order
//   peaks    sample peak, and ebur128's polyphase true-peak FIR (embarrassingly parallel in time)
// Pt^m = D A^(64 m) D are computed on the host in double-double (tile_handoff_power) and rounded once; the
// difference coordinates D keep the hand-off at the accuracy of the serial recursion (see loudness_tile.cu).
#include <math.h>
#include <string.h>

#include "ssb_internal.cuh"

namespace ssb {

namespace {

constexpr int kScanLS = 64;        // frames per segment
constexpr int kScanThreads = 256;
constexpr int kMaxPow = 32;        // Pt^0 .. Pt^32 (mono: 32 segments per warp)
constexpr int kFileWarmBuckets = 4;  // 0.4 s run-in of a file-mode chunk: |A^n| < e^-95 there at every sample rate

struct ScanArgs {
  double na[5];
  double b[5];
  float tp4[3][12];
  float tp2[24];
  const float* in;      // [n][in_stride_frames][C]
  const double* powers; // device: [kMaxPow + 1][16], Pt^m
  double* filt;
  double* bucket;
  float* speak;
  float* tpeak;
  float* tphist;
  double* ring;         // [n][ring_frames][C] or nullptr
  size_t in_stride_frames;
  size_t frames;
  size_t ring_frames;
  size_t ring_pos;
  uint64_t active_mask;
  unsigned s100;
  unsigned pos0;
  unsigned slot0;
  int tp_factor;        // 0, 2, 4
  int do_sample_peak;
  // file mode (whole-file one-shot, one stream): blockIdx.x is a time chunk instead of a stream
  double* file_buckets;       // [C][file_bucket_stride] sum of y^2 per 100 ms bucket by GLOBAL bucket index; nullptr: streaming
  size_t file_bucket_stride;
  size_t chunk_frames;        // frames a CTA owns (multiple of s100)
  size_t warm_frames;         // zero-state run-in before the owned range (multiple of s100)
};

__device__ __forceinline__ void to_diff(double v1, double v2, double v3, double v4, double& d0, double& d1,
                                        double& d2, double& d3) {
  const double e1 = v1 - v2, e2 = v2 - v3, e3 = v3 - v4;
  d0 = v1; d1 = e1; d2 = e1 - e2; d3 = (e1 - e2) - (e2 - e3);
}
__device__ __forceinline__ void from_diff(double d0, double d1, double d2, double d3, double& v1, double& v2,
                                          double& v3, double& v4) {
  const double e2 = d1 - d2;
  const double e3 = e2 - (d2 - d3);
  v1 = d0; v2 = d0 - d1; v3 = v2 - e2; v4 = v3 - e3;
}
// r = M x (+ add)
__device__ __forceinline__ void matvec(const double* __restrict__ M, double x0, double x1, double x2, double x3,
                                       double& r0, double& r1, double& r2, double& r3) {
  r0 = fma(M[0], x0, fma(M[1], x1, fma(M[2], x2, M[3] * x3)));
  r1 = fma(M[4], x0, fma(M[5], x1, fma(M[6], x2, M[7] * x3)));
  r2 = fma(M[8], x0, fma(M[9], x1, fma(M[10], x2, M[11] * x3)));
  r3 = fma(M[12], x0, fma(M[13], x1, fma(M[14], x2, M[15] * x3)));
}

template <int C>
__global__ void __launch_bounds__(kScanThreads)
k_loudness_scan(const __grid_constant__ ScanArgs a) {
  constexpr int SPW = 32 / C;                 // segments per warp (per channel)
  constexpr int NSEG = kScanThreads / C;      // segments per sweep (per channel)
  constexpr int NWARP = kScanThreads / 32;
  __shared__ double s_pow[(kMaxPow + 1) * 16];
  __shared__ double s_warp[NWARP][C][4];      // end state of each warp's span (zero incoming state), then prefix
  __shared__ double s_carry[C][4];            // true state entering the sweep (raw v1..v4)
  __shared__ double s_part[NSEG][C][2];
  __shared__ float s_pk[2][kScanThreads];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = tid / C;                      // segment of the sweep
  const int c = tid - k * C;                  // channel
  const int kw = (lane / C);                  // segment index inside the warp
  // File mode: the K-weighting poles sit at radius exp(-~240 / rate) per sample (the 38 Hz high-pass), so the state a
  // chunk inherits from audio more than 0.4 s back is below 1e-40 of full scale: a CTA that starts `warm_frames`
  // early from zero state reproduces the serial recursion down to that recursion's own rounding noise (~2e-11 of full
  // scale at 48 kHz: two f64 runs of it never re-synchronise their roundings; tests/test_design_claims_cpu.py), and the
  // chunks of one file run on different SMs.  Only the owned buckets are written.
  const bool file_mode = a.file_buckets != nullptr;
  const size_t stream = file_mode ? 0 : blockIdx.x;
  const size_t own0 = file_mode ? (size_t)blockIdx.x * a.chunk_frames : 0;
  const size_t f_begin = (file_mode && own0 > a.warm_frames) ? own0 - a.warm_frames : 0;
  const size_t f_end = file_mode ? (own0 + a.chunk_frames < a.frames ? own0 + a.chunk_frames : a.frames) : a.frames;
  const size_t gidx = stream * C + c;
  const bool live = (a.active_mask >> c) & 1ull;
  const float* x_base = a.in + stream * a.in_stride_frames * C + c;

  for (int i = tid; i < (kMaxPow + 1) * 16; i += kScanThreads) s_pow[i] = a.powers[i];
  if (k == 0) {
    const double* f = a.filt + gidx * 4;
    const bool from_state = live && f_begin == 0;   // a run-in starts from zero state
    s_carry[c][0] = from_state ? f[0] : 0.0; s_carry[c][1] = from_state ? f[1] : 0.0;
    s_carry[c][2] = from_state ? f[2] : 0.0; s_carry[c][3] = from_state ? f[3] : 0.0;
  }
  // bucket bookkeeping lives in thread (k == 0, c)
  double acc_cur = 0.0;
  unsigned slot = a.slot0;
  size_t gbucket = f_begin / a.s100;                 // file mode: global index of the bucket in progress
  const size_t own_bucket0 = own0 / a.s100;
  if (!file_mode && k == 0 && live && a.pos0 > 0) acc_cur = a.bucket[gidx * kNB + slot];
  float sp = 0.f, tp = 0.f;
  __syncthreads();

  const size_t sweep_frames = (size_t)NSEG * kScanLS;
  unsigned pos_sweep = file_mode ? 0u : a.pos0;  // position of the sweep start inside the bucket in progress
  for (size_t f0 = f_begin; f0 < f_end; f0 += sweep_frames) {
    const size_t seg0 = f0 + (size_t)k * kScanLS;
    const int lv = seg0 >= f_end ? 0 : (int)((f_end - seg0) < (size_t)kScanLS ? (f_end - seg0) : kScanLS);
    const float* xs = x_base + seg0 * C;

    // ---- pass 1: zero-state response of my segment ----
    double z1 = 0, z2 = 0, z3 = 0, z4 = 0;
    for (int i = 0; i < lv; i++) {
      const double x = (double)xs[(size_t)i * C];
      double t = fma(a.na[4], z4, x);
      t = fma(a.na[3], z3, t);
      t = fma(a.na[2], z2, t);
      const double z0 = fma(a.na[1], z1, t);
      z4 = z3; z3 = z2; z2 = z1; z1 = z0;
    }
    // ---- scan: S_k = state at the END of segment k (difference coordinates) ----
    double u0, u1, u2, u3;
    to_diff(z1, z2, z3, z4, u0, u1, u2, u3);
    if (k == 0) {  // fold the carried state into the first element: S_0 = Pt D carry + D z_0
      double d0, d1, d2, d3, r0, r1, r2, r3;
      to_diff(s_carry[c][0], s_carry[c][1], s_carry[c][2], s_carry[c][3], d0, d1, d2, d3);
      matvec(&s_pow[16], d0, d1, d2, d3, r0, r1, r2, r3);
      u0 += r0; u1 += r1; u2 += r2; u3 += r3;
    }
    // inclusive scan inside the warp over the SPW segments of this channel (lane stride C)
#pragma unroll
    for (int o = 1; o < SPW; o <<= 1) {
      const double p0 = __shfl_up_sync(0xffffffffu, u0, o * C), p1 = __shfl_up_sync(0xffffffffu, u1, o * C);
      const double p2 = __shfl_up_sync(0xffffffffu, u2, o * C), p3 = __shfl_up_sync(0xffffffffu, u3, o * C);
      if (kw >= o) {
        double r0, r1, r2, r3;
        matvec(&s_pow[o * 16], p0, p1, p2, p3, r0, r1, r2, r3);
        u0 += r0; u1 += r1; u2 += r2; u3 += r3;
      }
    }
    if (kw == SPW - 1) { s_warp[warp][c][0] = u0; s_warp[warp][c][1] = u1; s_warp[warp][c][2] = u2; s_warp[warp][c][3] = u3; }
    __syncthreads();
    if (tid < C) {  // serial combine of the warp totals: W_w = Pt^SPW W_{w-1} + local_w
      double w0 = s_warp[0][tid][0], w1 = s_warp[0][tid][1], w2 = s_warp[0][tid][2], w3 = s_warp[0][tid][3];
      for (int w = 1; w < NWARP; w++) {
        double r0, r1, r2, r3;
        matvec(&s_pow[SPW * 16], w0, w1, w2, w3, r0, r1, r2, r3);
        w0 = r0 + s_warp[w][tid][0]; w1 = r1 + s_warp[w][tid][1];
        w2 = r2 + s_warp[w][tid][2]; w3 = r3 + s_warp[w][tid][3];
        s_warp[w][tid][0] = w0; s_warp[w][tid][1] = w1; s_warp[w][tid][2] = w2; s_warp[w][tid][3] = w3;
      }
    }
    __syncthreads();
    if (warp > 0) {  // add the state entering this warp's span, advanced through kw + 1 segments
      double r0, r1, r2, r3;
      matvec(&s_pow[(kw + 1) * 16], s_warp[warp - 1][c][0], s_warp[warp - 1][c][1], s_warp[warp - 1][c][2],
             s_warp[warp - 1][c][3], r0, r1, r2, r3);
      u0 += r0; u1 += r1; u2 += r2; u3 += r3;
    }
    // state entering my segment = end state of the previous one
    double d0 = __shfl_up_sync(0xffffffffu, u0, C), d1 = __shfl_up_sync(0xffffffffu, u1, C);
    double d2 = __shfl_up_sync(0xffffffffu, u2, C), d3 = __shfl_up_sync(0xffffffffu, u3, C);
    if (kw == 0 && warp > 0) {
      d0 = s_warp[warp - 1][c][0]; d1 = s_warp[warp - 1][c][1]; d2 = s_warp[warp - 1][c][2]; d3 = s_warp[warp - 1][c][3];
    }
    double v1, v2, v3, v4;
    from_diff(d0, d1, d2, d3, v1, v2, v3, v4);
    if (k == 0) { v1 = s_carry[c][0]; v2 = s_carry[c][1]; v3 = s_carry[c][2]; v4 = s_carry[c][3]; }

    // ---- pass 2: the reference's recursion from the true state ----
    const unsigned seg_pos = (unsigned)((pos_sweep + (unsigned long long)k * kScanLS) % a.s100);
    const int lb = (int)min((unsigned)lv, a.s100 - seg_pos);   // samples [0, lb) belong to the bucket in progress
    double accA = 0.0, accB = 0.0;
    size_t rpos = a.ring ? (a.ring_pos + seg0) % a.ring_frames : 0;
    double* rg = a.ring ? a.ring + stream * a.ring_frames * C + c : nullptr;
    for (int i = 0; i < lv; i++) {
      const float xf = xs[(size_t)i * C];
      sp = fmaxf(sp, fabsf(xf));
      double y = 0.0;
      if (live) {
        double t = fma(a.na[4], v4, (double)xf);
        t = fma(a.na[3], v3, t);
        t = fma(a.na[2], v2, t);
        const double v0 = fma(a.na[1], v1, t);
        y = a.b[4] * v4;
        y = fma(a.b[3], v3, y);
        y = fma(a.b[2], v2, y);
        y = fma(a.b[1], v1, y);
        y = fma(a.b[0], v0, y);
        v4 = v3; v3 = v2; v2 = v1; v1 = v0;
        if (i < lb) accA = fma(y, y, accA); else accB = fma(y, y, accB);
      }
      if (rg) {
        rg[rpos * C] = y;
        if (++rpos == a.ring_frames) rpos = 0;
      }
    }
    s_part[k][c][0] = accA;
    s_part[k][c][1] = accB;
    // ---- true peak: f32 polyphase FIR, history straight from the input (or the stored tail for the first taps) ----
    if (a.tp_factor) {
      const int W = a.tp_factor == 4 ? 11 : 23;
      for (int i = 0; i < lv; i++) {
        const long long n = (long long)seg0 + i;
        const float x0 = xs[(size_t)i * C];
        if (a.tp_factor == 4) {
          float acc[3] = {x0 * a.tp4[0][0], x0 * a.tp4[1][0], x0 * a.tp4[2][0]};
          for (int t = 1; t <= W; t++) {
            const long long m = n - t;
            const float xm = m >= 0 ? x_base[(size_t)m * C] : a.tphist[gidx * kTpHist + (int)(-m - 1)];
            acc[0] = fmaf(xm, a.tp4[0][t], acc[0]);
            acc[1] = fmaf(xm, a.tp4[1][t], acc[1]);
            acc[2] = fmaf(xm, a.tp4[2][t], acc[2]);
          }
          tp = fmaxf(tp, fmaxf(fabsf(acc[0]), fmaxf(fabsf(acc[1]), fabsf(acc[2]))));
        } else {
          float acc = x0 * a.tp2[0];
          for (int t = 1; t <= W; t++) {
            const long long m = n - t;
            const float xm = m >= 0 ? x_base[(size_t)m * C] : a.tphist[gidx * kTpHist + (int)(-m - 1)];
            acc = fmaf(xm, a.tp2[t], acc);
          }
          tp = fmaxf(tp, fabsf(acc));
        }
      }
    }
    // the last non-empty segment of the sweep owns the state carried to the next sweep / next call
    const bool last_seg = lv > 0 && (seg0 + kScanLS >= f_end || k == NSEG - 1);
    __syncthreads();  // s_carry / s_warp / s_part of this sweep fully consumed / produced
    if (last_seg) { s_carry[c][0] = v1; s_carry[c][1] = v2; s_carry[c][2] = v3; s_carry[c][3] = v4; }
    // ---- fold the segment partials into the bucket ring, in time order ----
    if (k == 0) {
      unsigned p = pos_sweep;
      for (int j = 0; j < NSEG; j++) {
        const size_t sj = f0 + (size_t)j * kScanLS;
        if (sj >= f_end) break;
        const unsigned lvj = (unsigned)((f_end - sj) < (size_t)kScanLS ? (f_end - sj) : kScanLS);
        acc_cur += s_part[j][c][0];
        if (p + lvj >= a.s100) {   // the bucket in progress ends inside (or at the end of) this segment
          if (file_mode) {
            if (gbucket >= own_bucket0) a.file_buckets[(size_t)c * a.file_bucket_stride + gbucket] = live ? acc_cur : 0.0;
            ++gbucket;
          } else {
            if (live) a.bucket[gidx * kNB + slot] = acc_cur;
            slot = (slot + 1) % kNB;
          }
          acc_cur = s_part[j][c][1];
          p = p + lvj - a.s100;
        } else {
          p += lvj;
        }
      }
    }
    pos_sweep = (unsigned)((pos_sweep + sweep_frames) % a.s100);
    __syncthreads();
  }

  // ---------------- epilogue ----------------
  if (file_mode) return;   // the one-shot meter keeps nothing but the bucket energies (an unfinished bucket is never gated)
  s_pk[0][tid] = sp;
  s_pk[1][tid] = tp;
  __syncthreads();
  if (k == 0) {
    for (int j = 1; j < NSEG; j++) { sp = fmaxf(sp, s_pk[0][j * C + c]); tp = fmaxf(tp, s_pk[1][j * C + c]); }
    a.bucket[gidx * kNB + slot] = live ? acc_cur : 0.0;
    if (live) {
      double* f = a.filt + gidx * 4;
      const double tiny = 2.2250738585072014e-308;
      for (int i = 0; i < 4; i++) f[i] = fabs(s_carry[c][i]) < tiny ? 0.0 : s_carry[c][i];
    }
    if (a.do_sample_peak) a.speak[gidx] = fmaxf(a.speak[gidx], sp);
    if (a.tp_factor) {
      a.tpeak[gidx] = fmaxf(a.tpeak[gidx], tp);
      // new history: the last kTpHist - 1 inputs, newest first; older entries slide when the call was short
      float nh[kTpHist];
      for (int t = 0; t < kTpHist; t++) {
        const long long m = (long long)a.frames - 1 - t;
        nh[t] = m >= 0 ? x_base[(size_t)m * C] : a.tphist[gidx * kTpHist + (int)(-m - 1)];
      }
      for (int t = 0; t < kTpHist; t++) a.tphist[gidx * kTpHist + t] = nh[t];
    }
  }
}

}  // namespace

// Pt^m for m = 0..kMaxPow, device-resident per handle (capi owns the buffer)
void scan_power_table(const double a[5], double* host_table /* [(kMaxPow+1)*16] */) {
  for (int m = 0; m <= kMaxPow; m++) tile_handoff_power(a, kScanLS * m, host_table + 16 * m);
}
int scan_power_table_doubles() { return (kMaxPow + 1) * 16; }

bool scan_path_usable(const LoudParams& p, const LoudState& st, size_t frames) {
  if (p.channels != 1 && p.channels != 2) return false;
  if (st.n_streams > 64) return false;          // many streams: the batch kernels win
  if (p.s100 < (unsigned)kScanLS) return false;
  if (frames < 512) return false;               // short calls: the serial kernel's latency is lower
  // A sweep covers 256 / C segments of kScanLS frames whose lanes write the ring concurrently: if the ring is shorter
  // than a sweep (rates below ~2.7 kHz stereo / 5.4 kHz mono) two segments of one sweep would land on the same slot in
  // an undefined order — those handles take the serial kernel
  if (st.ring && st.ring_frames < (size_t)(256 / p.channels) * kScanLS) return false;
  return true;
}

cudaError_t launch_loudness_scan(const LoudParams& p, const LoudState& st, const double* d_powers, const float* d_in,
                                 size_t frames, size_t in_stride_frames, uint32_t pos0, uint64_t bucket0,
                                 size_t ring_pos, cudaStream_t s, uint64_t* launches) {
  ScanArgs a{};
  memcpy(a.na, p.na, sizeof(a.na));
  memcpy(a.b, p.b, sizeof(a.b));
  memcpy(a.tp4, p.tp4, sizeof(a.tp4));
  memcpy(a.tp2, p.tp2, sizeof(a.tp2));
  a.in = d_in;
  a.powers = d_powers;
  a.filt = st.filt;
  a.bucket = st.bucket;
  a.speak = st.speak;
  a.tpeak = st.tpeak;
  a.tphist = st.tphist;
  a.ring = st.ring;
  a.in_stride_frames = in_stride_frames;
  a.frames = frames;
  a.ring_frames = st.ring_frames;
  a.ring_pos = ring_pos;
  a.active_mask = p.do_filter ? p.active_mask : 0;
  a.s100 = p.s100;
  a.pos0 = pos0;
  a.slot0 = (unsigned)(bucket0 % kNB);
  a.tp_factor = p.do_true_peak ? p.tp_factor : 0;
  a.do_sample_peak = p.do_sample_peak;
  if (p.channels == 1) k_loudness_scan<1><<<(unsigned)st.n_streams, kScanThreads, 0, s>>>(a);
  else k_loudness_scan<2><<<(unsigned)st.n_streams, kScanThreads, 0, s>>>(a);
  if (launches) ++*launches;
  return cudaGetLastError();
}

// Whole-file one-shot (Analyzer::calculate_integrated_lufs, analyzer.rs:170-182): one stream starting at the meter's
// reset state, split into time chunks of `chunk_buckets` 100 ms buckets, one CTA each, every chunk but the first
// preceded by a zero-state run-in of kFileWarmBuckets buckets.  Fills d_file_buckets[C][stride] with the energy sums of
// every COMPLETE bucket; the caller gates them (launch_file_gating).
cudaError_t launch_loudness_scan_file(const LoudParams& p, const LoudState& st, const double* d_powers, const float* d_in,
                                      size_t frames, double* d_file_buckets, size_t bucket_stride, size_t chunk_buckets,
                                      cudaStream_t s, uint64_t* launches) {
  if (!frames) return cudaSuccess;
  ScanArgs a{};
  memcpy(a.na, p.na, sizeof(a.na));
  memcpy(a.b, p.b, sizeof(a.b));
  a.in = d_in;
  a.powers = d_powers;
  a.filt = st.filt;
  a.bucket = st.bucket;
  a.speak = st.speak;
  a.tpeak = st.tpeak;
  a.tphist = st.tphist;
  a.ring = nullptr;
  a.in_stride_frames = frames;
  a.frames = frames;
  a.active_mask = p.do_filter ? p.active_mask : 0;
  a.s100 = p.s100;
  a.tp_factor = 0;         // only the integrated loudness leaves the one-shot meter
  a.do_sample_peak = 0;
  a.file_buckets = d_file_buckets;
  a.file_bucket_stride = bucket_stride;
  a.chunk_frames = chunk_buckets * (size_t)p.s100;
  a.warm_frames = (size_t)kFileWarmBuckets * p.s100;
  const size_t n_chunks = (frames + a.chunk_frames - 1) / a.chunk_frames;
  if (p.channels == 1) k_loudness_scan<1><<<(unsigned)n_chunks, kScanThreads, 0, s>>>(a);
  else k_loudness_scan<2><<<(unsigned)n_chunks, kScanThreads, 0, s>>>(a);
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace ssb
