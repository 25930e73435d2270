// capture.cu — the data formats either side of the analyzer path:
//
//   * decoded PCM -> interleaved f32: what AudioFile::decode_file (reference src/audio_player.rs:169-267)
//     produces for WAV / AIFF input.  The container's samples are already interleaved, symphonia's PCM decoder
//     de-interleaves them into planes and `SampleBuffer::<f32>::copy_interleaved_ref` (audio_player.rs:248)
//     re-interleaves while converting with symphonia-core's `FromSample`, so end to end it is an element-wise
//     conversion of the interleaved stream: HBM-bound byte work, `bytes_per_sample` read + 4 written per sample.
//   * the capture ring (`RBuffer`, reference src/tui.rs:37; main.rs:63-65; audio_capture.rs:31-59) and the
//     microphone tick that reads it (tui.rs:1427-1480).
//
// No compute happens on the host: the producer side of the ring only stores samples into pinned memory.
#include <string.h>

#include <atomic>
#include <new>

#include "ssb_handle.cuh"

using namespace ssb;

// ================================================================================================
// PCM -> f32
// ================================================================================================
namespace {

__host__ __device__ constexpr int pcm_bytes(int fmt) {
  return fmt == SSB_PCM_U8 || fmt == SSB_PCM_S8                            ? 1
         : fmt == SSB_PCM_S16LE || fmt == SSB_PCM_S16BE                    ? 2
         : fmt == SSB_PCM_S24LE || fmt == SSB_PCM_S24BE                    ? 3
         : fmt == SSB_PCM_S32LE || fmt == SSB_PCM_S32BE || fmt == SSB_PCM_F32LE || fmt == SSB_PCM_F32BE ? 4
         : fmt == SSB_PCM_F64LE || fmt == SSB_PCM_F64BE                    ? 8
                                                                           : 0;
}

__device__ __forceinline__ unsigned bswap32(unsigned v) { return __byte_perm(v, 0, 0x0123); }

// symphonia-core 0.5.5 conv.rs `impl FromSample<S> for f32` (un-vendored; restated in oracle/capture_ref.py).
// Integer inputs up to 24 bits convert exactly and the division is by a power of two; the 32-bit rule goes
// through f64 and rounds once, which is what one I2F.RN followed by an exact scaling does.
__device__ __forceinline__ float cvt_u8(unsigned b) { return __int2float_rn((int)b - 128) * 0.0078125f; }
__device__ __forceinline__ float cvt_s8(unsigned b) { return __int2float_rn((int)(signed char)b) * 0.0078125f; }
__device__ __forceinline__ float cvt_s16(unsigned v) { return __int2float_rn((int)(short)v) * (1.0f / 32768.0f); }
__device__ __forceinline__ float cvt_s24(unsigned v) {  // v: 24 significant bits, little-endian order
  return __int2float_rn(((int)(v << 8)) >> 8) * (1.0f / 8388608.0f);
}
__device__ __forceinline__ float cvt_s32(unsigned v) { return __int2float_rn((int)v) * (1.0f / 2147483648.0f); }

// One sample by byte loads: any alignment, used for ragged tails and unaligned device pointers.
template <int FMT>
__device__ __forceinline__ float pcm_scalar(const unsigned char* p, size_t i) {
  constexpr int B = pcm_bytes(FMT);
  const unsigned char* q = p + i * B;
  if (FMT == SSB_PCM_U8) return cvt_u8(q[0]);
  if (FMT == SSB_PCM_S8) return cvt_s8(q[0]);
  if (FMT == SSB_PCM_S16LE) return cvt_s16(q[0] | (q[1] << 8));
  if (FMT == SSB_PCM_S16BE) return cvt_s16(q[1] | (q[0] << 8));
  if (FMT == SSB_PCM_S24LE) return cvt_s24(q[0] | (q[1] << 8) | (q[2] << 16));
  if (FMT == SSB_PCM_S24BE) return cvt_s24(q[2] | (q[1] << 8) | (q[0] << 16));
  if (B == 4) {
    const unsigned le = q[0] | (q[1] << 8) | (q[2] << 16) | ((unsigned)q[3] << 24);
    const unsigned v = (FMT == SSB_PCM_S32BE || FMT == SSB_PCM_F32BE) ? bswap32(le) : le;
    return (FMT == SSB_PCM_S32LE || FMT == SSB_PCM_S32BE) ? cvt_s32(v) : __uint_as_float(v);
  }
  unsigned lo = q[0] | (q[1] << 8) | (q[2] << 16) | ((unsigned)q[3] << 24);
  unsigned hi = q[4] | (q[5] << 8) | (q[6] << 16) | ((unsigned)q[7] << 24);
  if (FMT == SSB_PCM_F64BE) {
    const unsigned t = bswap32(lo);
    lo = bswap32(hi);
    hi = t;
  }
  return __double2float_rn(__hiloint2double((int)hi, (int)lo));
}

// Four consecutive samples per thread per step: 4*B bytes in (one vector load, three words for 24-bit audio),
// one float4 out.  A warp reads 128*B contiguous bytes and writes 512 contiguous bytes per step.
template <int FMT>
__device__ __forceinline__ float4 pcm_quad(const unsigned char* p, size_t quad) {
  constexpr int B = pcm_bytes(FMT);
  float4 o;
  if (B == 1) {
    const unsigned w = __ldg(reinterpret_cast<const unsigned*>(p) + quad);
    if (FMT == SSB_PCM_U8) {
      o = make_float4(cvt_u8(w & 0xff), cvt_u8((w >> 8) & 0xff), cvt_u8((w >> 16) & 0xff), cvt_u8(w >> 24));
    } else {
      o = make_float4(cvt_s8(w & 0xff), cvt_s8((w >> 8) & 0xff), cvt_s8((w >> 16) & 0xff), cvt_s8(w >> 24));
    }
  } else if (B == 2) {
    uint2 w = __ldg(reinterpret_cast<const uint2*>(p) + quad);
    if (FMT == SSB_PCM_S16BE) {
      w.x = __byte_perm(w.x, 0, 0x2301);
      w.y = __byte_perm(w.y, 0, 0x2301);
    }
    o = make_float4(cvt_s16(w.x & 0xffff), cvt_s16(w.x >> 16), cvt_s16(w.y & 0xffff), cvt_s16(w.y >> 16));
  } else if (B == 3) {
    const unsigned* q = reinterpret_cast<const unsigned*>(p) + quad * 3;
    const unsigned a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);  // bytes 0-3, 4-7, 8-11 of the group
    unsigned s0 = a & 0xffffff, s1 = (a >> 24) | ((b & 0xffff) << 8), s2 = (b >> 16) | ((c & 0xff) << 16), s3 = c >> 8;
    if (FMT == SSB_PCM_S24BE) {
      s0 = __byte_perm(s0, 0, 0x4012);
      s1 = __byte_perm(s1, 0, 0x4012);
      s2 = __byte_perm(s2, 0, 0x4012);
      s3 = __byte_perm(s3, 0, 0x4012);
    }
    o = make_float4(cvt_s24(s0), cvt_s24(s1), cvt_s24(s2), cvt_s24(s3));
  } else if (B == 4) {
    uint4 w = __ldg(reinterpret_cast<const uint4*>(p) + quad);
    if (FMT == SSB_PCM_S32BE || FMT == SSB_PCM_F32BE) {
      w.x = bswap32(w.x); w.y = bswap32(w.y); w.z = bswap32(w.z); w.w = bswap32(w.w);
    }
    if (FMT == SSB_PCM_S32LE || FMT == SSB_PCM_S32BE) o = make_float4(cvt_s32(w.x), cvt_s32(w.y), cvt_s32(w.z), cvt_s32(w.w));
    else o = make_float4(__uint_as_float(w.x), __uint_as_float(w.y), __uint_as_float(w.z), __uint_as_float(w.w));
  } else {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(p) + 2 * quad);
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + 2 * quad + 1);
    if (FMT == SSB_PCM_F64BE) {
      unsigned t;
      t = bswap32(u.x); u.x = bswap32(u.y); u.y = t;
      t = bswap32(u.z); u.z = bswap32(u.w); u.w = t;
      t = bswap32(v.x); v.x = bswap32(v.y); v.y = t;
      t = bswap32(v.z); v.z = bswap32(v.w); v.w = t;
    }
    o = make_float4(__double2float_rn(__hiloint2double((int)u.y, (int)u.x)),
                    __double2float_rn(__hiloint2double((int)u.w, (int)u.z)),
                    __double2float_rn(__hiloint2double((int)v.y, (int)v.x)),
                    __double2float_rn(__hiloint2double((int)v.w, (int)v.z)));
  }
  return o;
}

template <int FMT>
__global__ void __launch_bounds__(256)
k_pcm_to_f32(const unsigned char* __restrict__ in, size_t n, float* __restrict__ out, int vector_ok) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nthreads = (size_t)gridDim.x * blockDim.x;
  size_t done = 0;
  if (vector_ok) {
    const size_t quads = n / 4;
    float4* o4 = reinterpret_cast<float4*>(out);
    size_t q = tid;
    // four independent loads in flight per thread before the first store
    for (; q + 3 * nthreads < quads; q += 4 * nthreads) {
      const float4 a = pcm_quad<FMT>(in, q), b = pcm_quad<FMT>(in, q + nthreads);
      const float4 c = pcm_quad<FMT>(in, q + 2 * nthreads), d = pcm_quad<FMT>(in, q + 3 * nthreads);
      o4[q] = a; o4[q + nthreads] = b; o4[q + 2 * nthreads] = c; o4[q + 3 * nthreads] = d;
    }
    for (; q < quads; q += nthreads) o4[q] = pcm_quad<FMT>(in, q);
    done = quads * 4;
  }
  for (size_t i = done + tid; i < n; i += nthreads) out[i] = pcm_scalar<FMT>(in, i);
}

template <int FMT>
cudaError_t launch_pcm_fmt(const void* d_in, size_t n, float* d_out, cudaStream_t s) {
  constexpr int B = pcm_bytes(FMT);
  // vector path: the input must be aligned for its widest load, the output for float4
  const uintptr_t in_align = B == 1 ? 4 : (B == 2 ? 8 : (B == 3 ? 4 : 16));
  const int vector_ok = (reinterpret_cast<uintptr_t>(d_in) % in_align == 0) && (reinterpret_cast<uintptr_t>(d_out) % 16 == 0);
  size_t blocks = (n / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 16) blocks = 148 * 16;  // grid-stride over a whole number of waves (8 CTAs of 256 per SM x 2)
  k_pcm_to_f32<FMT><<<(unsigned)blocks, 256, 0, s>>>(static_cast<const unsigned char*>(d_in), n, d_out, vector_ok);
  return cudaGetLastError();
}

cudaError_t launch_pcm(const void* d_in, size_t n, int fmt, float* d_out, cudaStream_t s, uint64_t* launches) {
  if (!n) return cudaSuccess;
  if (launches) ++*launches;
  switch (fmt) {
    case SSB_PCM_U8: return launch_pcm_fmt<SSB_PCM_U8>(d_in, n, d_out, s);
    case SSB_PCM_S8: return launch_pcm_fmt<SSB_PCM_S8>(d_in, n, d_out, s);
    case SSB_PCM_S16LE: return launch_pcm_fmt<SSB_PCM_S16LE>(d_in, n, d_out, s);
    case SSB_PCM_S16BE: return launch_pcm_fmt<SSB_PCM_S16BE>(d_in, n, d_out, s);
    case SSB_PCM_S24LE: return launch_pcm_fmt<SSB_PCM_S24LE>(d_in, n, d_out, s);
    case SSB_PCM_S24BE: return launch_pcm_fmt<SSB_PCM_S24BE>(d_in, n, d_out, s);
    case SSB_PCM_S32LE: return launch_pcm_fmt<SSB_PCM_S32LE>(d_in, n, d_out, s);
    case SSB_PCM_S32BE: return launch_pcm_fmt<SSB_PCM_S32BE>(d_in, n, d_out, s);
    case SSB_PCM_F32LE: return launch_pcm_fmt<SSB_PCM_F32LE>(d_in, n, d_out, s);
    case SSB_PCM_F32BE: return launch_pcm_fmt<SSB_PCM_F32BE>(d_in, n, d_out, s);
    case SSB_PCM_F64LE: return launch_pcm_fmt<SSB_PCM_F64LE>(d_in, n, d_out, s);
    case SSB_PCM_F64BE: return launch_pcm_fmt<SSB_PCM_F64BE>(d_in, n, d_out, s);
  }
  return cudaErrorInvalidValue;
}

// ================================================================================================
// capture ring -> one microphone tick
// ================================================================================================
// The device mirror holds the ring in PHYSICAL order; `oldest` is the physical index of the oldest value, so
// `to_vec()[i] == ring[(oldest + i) % cap]`.  One launch does everything the tick needs from the ring:
//   blocks [0, wave_blocks): get_waveform over mid = (l + r) / 2 of every stereo pair of the logical vector
//                            (audio_player.rs:400-419 fused into analyzer.rs:107-137), one warp per column;
//   the remaining blocks:    the logical spans the FFT and the meter read, copied out contiguously.
struct RingTickArgs {
  const float* ring;
  unsigned long long cap, oldest;
  unsigned long long frames;      // cap / 2: length of mid
  double spp;                     // frames / window (the reference's samples_per_pixel)
  unsigned long long columns;
  float* wave;                    // [columns][2] (min, max)
  unsigned wave_blocks;
  unsigned long long fft_v0, fft_n;    // logical value range -> d_fft (interleaved stereo window)
  float* d_fft;
  unsigned long long lufs_v0, lufs_n;  // logical value range -> d_lufs
  float* d_lufs;
};

__device__ __forceinline__ float ring_at(const RingTickArgs& a, unsigned long long logical) {
  unsigned long long p = a.oldest + logical;
  if (p >= a.cap) p -= a.cap;
  return __ldg(a.ring + p);
}
__device__ __forceinline__ float ring_mid(const RingTickArgs& a, unsigned long long frame) {
  const float l = ring_at(a, 2 * frame), r = ring_at(a, 2 * frame + 1);
  return __fmul_rn(__fadd_rn(l, r), 0.5f);  // (l + r) / 2.
}

__global__ void __launch_bounds__(256) k_ring_tick(const __grid_constant__ RingTickArgs a) {
  if (blockIdx.x < a.wave_blocks) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long n_warps = ((unsigned long long)a.wave_blocks * blockDim.x) >> 5;
    for (unsigned long long i = warp; i < a.columns; i += n_warps) {
      // analyzer.rs:118-120: start = (i as f64 * spp) as usize, end = ((i+1) as f64 * spp).ceil() as usize .min(len)
      const unsigned long long start = (unsigned long long)__dmul_rn((double)i, a.spp);
      unsigned long long end = (unsigned long long)ceil(__dmul_rn((double)(i + 1), a.spp));
      if (end > a.frames) end = a.frames;
      float mn = 0.0f, mx = 0.0f;
      if (end > start) {
        mn = mx = ring_mid(a, start);  // reduce(): the first element seeds; f32::min/max ignore NaN like fminf/fmaxf
        for (unsigned long long j = start + lane; j < end; j += 32) {
          const float v = ring_mid(a, j);
          mn = fminf(mn, v);
          mx = fmaxf(mx, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
      }
      if (lane == 0) { a.wave[2 * i] = mn; a.wave[2 * i + 1] = mx; }
    }
    return;
  }
  const unsigned long long t = (unsigned long long)(blockIdx.x - a.wave_blocks) * blockDim.x + threadIdx.x;
  const unsigned long long nt = (unsigned long long)(gridDim.x - a.wave_blocks) * blockDim.x;
  for (unsigned long long i = t; i < a.fft_n; i += nt) a.d_fft[i] = ring_at(a, a.fft_v0 + i);
  for (unsigned long long i = t; i < a.lufs_n; i += nt) a.d_lufs[i] = ring_at(a, a.lufs_v0 + i);
}

}  // namespace

struct ssb_capture_ring {
  int device = 0;
  size_t cap = 0;
  float* h_ring = nullptr;  // pinned; physical order; the producer's only target
  float* d_ring = nullptr;  // device mirror, valid for values [.., mirrored)
  std::atomic<uint64_t> claimed{0};  // producer: set to the end of a push BEFORE its stores (seqlock "begin")
  std::atomic<uint64_t> written{0};  // values pushed so far; published with release order after the stores
  uint64_t mirrored = 0;             // consumer side: values already copied to d_ring
};

namespace {

// Bring the device mirror up to the producer's published position: copies only the physical spans written since
// the last call.  Returns the snapshot position W (to_vec() is then the cap values ending at W).
int32_t ring_sync_mirror(ssb_analyzer* h, ssb_capture_ring* r, uint64_t* w_out) {
  // The reference's to_vec runs under the ring's mutex and cannot fail; this one is lock-free, so a copy that a push
  // overlapped is retried.  Each attempt first waits (bounded) for the push in flight to publish: with a live audio
  // callback (a few hundred microseconds of stores every ~10 ms) the second attempt practically always succeeds.
  for (int attempt = 0; attempt < 64; attempt++) {
    for (int spin = 0; spin < (1 << 16) && r->claimed.load(std::memory_order_acquire) != r->written.load(std::memory_order_acquire); spin++) {
#if defined(__x86_64__) || defined(__i386__)
      __builtin_ia32_pause();
#endif
    }
    const uint64_t w = r->written.load(std::memory_order_acquire);
    uint64_t from = r->mirrored;
    if (w - from >= r->cap) from = w - r->cap;  // everything older has been overwritten
    uint64_t left = w - from;
    size_t p = (size_t)(from % r->cap);
    while (left) {
      const size_t n = left < r->cap - p ? (size_t)left : r->cap - p;
      CK(cudaMemcpyAsync(r->d_ring + p, r->h_ring + p, n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
      left -= n;
      p = 0;
    }
    if (w != r->mirrored) CK(cudaStreamSynchronize(h->stream));
    // the spans were being read by the copy engine while the producer kept running: they are intact unless it
    // lapped the ring (wrote a value index >= from + cap) in the meantime
    std::atomic_thread_fence(std::memory_order_seq_cst);
    const uint64_t w2 = r->claimed.load(std::memory_order_acquire);
    if (w2 - from <= r->cap) {
      r->mirrored = w;
      *w_out = w;
      return SSB_OK;
    }
    r->mirrored = from;  // torn: take a fresh snapshot
  }
  return fail(h, SSB_ERR_INVALID_ARG, "capture ring: the producer lapped the ring four times during one snapshot");
}

}  // namespace

extern "C" {

size_t ssb_pcm_bytes_per_sample(int32_t format) { return (size_t)pcm_bytes(format); }

int32_t ssb_pcm_to_f32_device(ssb_analyzer* h, const void* d_pcm, size_t n_samples, int32_t format, float* d_out) {
  if (!h) return SSB_ERR_INVALID_ARG;
  if (!pcm_bytes(format)) return fail(h, SSB_ERR_INVALID_ARG, "unknown PCM format %d", format);
  if (!n_samples) return SSB_OK;
  if (!d_pcm || !d_out) return fail(h, SSB_ERR_INVALID_ARG, "null buffer");
  DeviceGuard g(h->device);
  CK(launch_pcm(d_pcm, n_samples, format, d_out, h->stream, &h->launches));
  return SSB_OK;
}

int32_t ssb_pcm_to_f32(ssb_analyzer* h, const void* pcm, size_t n_samples, int32_t format, float* out) {
  if (!h) return SSB_ERR_INVALID_ARG;
  const size_t B = (size_t)pcm_bytes(format);
  if (!B) return fail(h, SSB_ERR_INVALID_ARG, "unknown PCM format %d", format);
  if (!n_samples) return SSB_OK;
  if (!pcm || !out) return fail(h, SSB_ERR_INVALID_ARG, "null buffer");
  DeviceGuard g(h->device);
  const size_t in_bytes = n_samples * B;
  const size_t out_off = (in_bytes + 255) & ~(size_t)255;
  int32_t rc = ensure_scratch(h, out_off + n_samples * sizeof(float));
  if (rc) return rc;
  char* base = reinterpret_cast<char*>(h->d_scratch);
  float* d_out = reinterpret_cast<float*>(base + out_off);
  CK(cudaMemcpyAsync(base, pcm, in_bytes, cudaMemcpyHostToDevice, h->stream));
  CK(launch_pcm(base, n_samples, format, d_out, h->stream, &h->launches));
  CK(cudaMemcpyAsync(out, d_out, n_samples * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return SSB_OK;
}

int32_t ssb_add_frames_pcm_device(ssb_analyzer* h, const void* d_pcm, int32_t format, size_t frames_per_stream) {
  if (!h) return SSB_ERR_INVALID_ARG;
  if (!pcm_bytes(format)) return fail(h, SSB_ERR_INVALID_ARG, "unknown PCM format %d", format);
  if (!frames_per_stream) return SSB_OK;
  if (!d_pcm) return fail(h, SSB_ERR_INVALID_ARG, "null input");
  DeviceGuard g(h->device);
  const size_t n = h->n_streams * frames_per_stream * h->channels;
  int32_t rc = ensure_stage(h, n);
  if (rc) return rc;
  const int i = h->stage_idx;
  h->stage_idx ^= 1;
  CK(cudaEventSynchronize(h->ev_consumed[i]));  // the f32 stage may still be read by the kernels of two calls ago
  CK(launch_pcm(d_pcm, n, format, h->d_stage[i], h->stream, &h->launches));
  rc = feed_device(h, h->d_stage[i], frames_per_stream, frames_per_stream);
  if (rc) return rc;
  CK(cudaEventRecord(h->ev_consumed[i], h->stream));
  return SSB_OK;
}

int32_t ssb_add_frames_pcm(ssb_analyzer* h, const void* pcm, int32_t format, size_t frames_per_stream) {
  if (!h) return SSB_ERR_INVALID_ARG;
  const size_t B = (size_t)pcm_bytes(format);
  if (!B) return fail(h, SSB_ERR_INVALID_ARG, "unknown PCM format %d", format);
  if (!frames_per_stream) return SSB_OK;
  if (!pcm) return fail(h, SSB_ERR_INVALID_ARG, "null input");
  DeviceGuard g(h->device);
  const size_t n = h->n_streams * frames_per_stream * h->channels;
  // raw bytes go to their own device buffer, the converted f32 to the meter's stage buffer
  CK(cudaStreamSynchronize(h->stream));  // the raw buffer may still feed the previous call's conversion
  if (n * B > h->pcm_cap) {
    cudaFree(h->d_pcm);
    h->d_pcm = nullptr;
    h->pcm_cap = 0;
    CK(cudaMalloc(&h->d_pcm, n * B));
    h->pcm_cap = n * B;
  }
  CK(cudaMemcpyAsync(h->d_pcm, pcm, n * B, cudaMemcpyHostToDevice, h->stream));
  int32_t rc = ssb_add_frames_pcm_device(h, h->d_pcm, format, frames_per_stream);
  if (rc) return rc;
  CK(cudaStreamSynchronize(h->stream));  // caller may reuse its buffer; results are consistent at return
  return SSB_OK;
}

// ---- capture ring -------------------------------------------------------------------------------

int32_t ssb_capture_ring_create(ssb_capture_ring** out, size_t capacity_values, int32_t device) {
  if (!out) return SSB_ERR_INVALID_ARG;
  *out = nullptr;
  if (!capacity_values) return SSB_ERR_INVALID_ARG;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return SSB_ERR_NO_DEVICE;
  int dev = device;
  if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) return SSB_ERR_NO_DEVICE;
  if (dev >= count) return SSB_ERR_INVALID_ARG;
  ssb_capture_ring* r = new (std::nothrow) ssb_capture_ring();
  if (!r) return SSB_ERR_NOMEM;
  r->device = dev;
  r->cap = capacity_values;
  DeviceGuard g(dev);
  cudaError_t e = cudaHostAlloc(&r->h_ring, capacity_values * sizeof(float), cudaHostAllocPortable);
  if (e == cudaSuccess) e = cudaMalloc(&r->d_ring, capacity_values * sizeof(float));
  if (e == cudaSuccess) e = cudaMemset(r->d_ring, 0, capacity_values * sizeof(float));  // buf.fill(0.0)
  if (e != cudaSuccess) {
    if (r->h_ring) cudaFreeHost(r->h_ring);
    cudaFree(r->d_ring);
    delete r;
    return SSB_ERR_CUDA + (int32_t)e;
  }
  memset(r->h_ring, 0, capacity_values * sizeof(float));
  *out = r;
  return SSB_OK;
}

void ssb_capture_ring_destroy(ssb_capture_ring* r) {
  if (!r) return;
  DeviceGuard g(r->device);
  cudaFree(r->d_ring);
  if (r->h_ring) cudaFreeHost(r->h_ring);
  delete r;
}

size_t ssb_capture_ring_capacity(const ssb_capture_ring* r) { return r ? r->cap : 0; }
uint64_t ssb_capture_ring_written(const ssb_capture_ring* r) {
  return r ? r->written.load(std::memory_order_acquire) : 0;
}

int32_t ssb_capture_ring_push(ssb_capture_ring* r, const float* data, size_t n, int32_t is_mono) {
  if (!r || (!data && n)) return SSB_ERR_INVALID_ARG;
  if (!n) return SSB_OK;
  const uint64_t w0 = r->written.load(std::memory_order_relaxed);
  const size_t cap = r->cap;
  // seqlock: announce the end position before touching the ring so a concurrent snapshot can tell it was lapped
  r->claimed.store(w0 + (is_mono ? 2 * (uint64_t)n - 1 : (uint64_t)n), std::memory_order_relaxed);
  std::atomic_thread_fence(std::memory_order_seq_cst);
  if (!is_mono) {
    // audio_buf.extend(data): only the last `cap` values of an oversized push survive
    const size_t skip = n > cap ? n - cap : 0;
    size_t p = (size_t)((w0 + skip) % cap);
    size_t left = n - skip;
    const float* src = data + skip;
    while (left) {
      const size_t m = left < cap - p ? left : cap - p;
      memcpy(r->h_ring + p, src, m * sizeof(float));
      src += m;
      left -= m;
      p = 0;
    }
    r->written.store(w0 + n, std::memory_order_release);
  } else {
    // audio_capture.rs:43-48: i == 0 -> [x], i > 0 -> [0., x]  =>  value j of the 2n-1: j even -> x[j/2], j odd -> 0
    const uint64_t total = 2 * (uint64_t)n - 1;
    const uint64_t skip = total > cap ? total - cap : 0;
    size_t p = (size_t)((w0 + skip) % cap);
    for (uint64_t j = skip; j < total; j++) {
      r->h_ring[p] = (j & 1) ? 0.0f : data[j >> 1];
      if (++p == cap) p = 0;
    }
    r->written.store(w0 + total, std::memory_order_release);
  }
  return SSB_OK;
}

int32_t ssb_capture_ring_to_vec(ssb_capture_ring* r, float* out, size_t cap) {
  if (!r || !out || cap < r->cap) return SSB_ERR_INVALID_ARG;
  // The reference's to_vec runs under the ring's mutex and cannot fail; this one is lock-free, so a copy that a push
  // overlapped is retried.  Each attempt first waits (bounded) for the push in flight to publish: with a live audio
  // callback (a few hundred microseconds of stores every ~10 ms) the second attempt practically always succeeds.
  for (int attempt = 0; attempt < 64; attempt++) {
    for (int spin = 0; spin < (1 << 16) && r->claimed.load(std::memory_order_acquire) != r->written.load(std::memory_order_acquire); spin++) {
#if defined(__x86_64__) || defined(__i386__)
      __builtin_ia32_pause();
#endif
    }
    const uint64_t w = r->written.load(std::memory_order_acquire);
    const size_t oldest = (size_t)(w % r->cap);
    memcpy(out, r->h_ring + oldest, (r->cap - oldest) * sizeof(float));
    memcpy(out + (r->cap - oldest), r->h_ring, oldest * sizeof(float));
    std::atomic_thread_fence(std::memory_order_seq_cst);
    if (r->claimed.load(std::memory_order_acquire) == w) return SSB_OK;  // no push started during the copy
  }
  return SSB_ERR_BUSY;
}

int32_t ssb_mic_tick(ssb_analyzer* h, ssb_capture_ring* ring, size_t n_fft, size_t lufs_samples, double waveform_window,
                     double* xy_mid, double* xy_side, size_t cap, size_t* n_points, double* xy_wave, size_t wave_cap,
                     size_t* n_wave_points, double* shortterm_lufs, int32_t* fft_status, int32_t* lufs_status) {
  if (!h || !ring || !n_points || !n_wave_points || !shortterm_lufs || !fft_status || !lufs_status)
    return SSB_ERR_INVALID_ARG;
  if (h->n_streams != 1) return fail(h, SSB_ERR_INVALID_ARG, "mic_tick needs a one-stream handle");
  if (!h->meter_ok) return fail(h, SSB_ERR_NOMEM, "the loudness meter is not initialised (a previous create/reinit failed)");
  if (ring->device != h->device) return fail(h, SSB_ERR_INVALID_ARG, "ring and analyzer live on different devices");
  *n_points = 0;
  *n_wave_points = 0;
  const size_t rate = h->rate;
  const size_t frames = ring->cap / 2;  // get_mid_and_side_samples: zip() drops an odd trailing value
  // tui.rs:1431-1446, 1465-1469: slices at fixed offsets from the oldest value; out of range = a panic there
  if (15 * rate > frames || n_fft > 15 * rate || lufs_samples > 30 * rate || 30 * rate > ring->cap)
    return fail(h, SSB_ERR_INVALID_ARG, "mic_tick: the reference's slices [15*rate - n_fft, 15*rate) / "
                "[30*rate - lufs_samples, 30*rate) do not fit a ring of %zu values at rate %zu", ring->cap, rate);
  *fft_status = fft_shape_check(n_fft, h->rate);
  h->tick_fft_status[0] = h->tick_fft_status[1] = *fft_status;
  *lufs_status = (lufs_samples % h->channels != 0) ? SSB_ERR_NOMEM : SSB_OK;  // add_frames_f32: ragged -> NoMem
  DeviceGuard g(h->device);
  FftPlan* plan = nullptr;
  size_t nb = 0;
  if (*fft_status == SSB_OK) {
    int32_t rc = get_plan(h, n_fft, h->rate, &plan);
    if (rc) return rc;
    nb = plan->n_bins;
    *n_points = nb;
    if (nb > cap || !xy_mid || !xy_side) return fail(h, SSB_ERR_CAPACITY, "mic_tick: need room for %zu points", nb);
  }
  size_t window = 0;
  const size_t cols = waveform_window_columns(waveform_window, frames, &window);
  *n_wave_points = 2 * cols;
  if (2 * cols > wave_cap || (cols && !xy_wave))
    return fail(h, SSB_ERR_CAPACITY, "mic_tick: need room for %zu waveform points", 2 * cols);

  // scratch: [stereo fft window | meter input | outputs: waveform min/max, dB planes, two status words]
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t fft_off = 0, lufs_off = up(2 * n_fft * sizeof(float));
  const size_t out_off = lufs_off + up(lufs_samples * sizeof(float));
  const size_t wave_bytes = up(2 * cols * sizeof(float));
  const size_t db_bytes = 2 * nb * sizeof(float) + 2 * sizeof(int32_t);
  int32_t rc = ensure_scratch(h, out_off + wave_bytes + db_bytes + 256);
  if (rc) return rc;
  char* base = reinterpret_cast<char*>(h->d_scratch);
  float* d_fft = reinterpret_cast<float*>(base + fft_off);
  float* d_lufs = reinterpret_cast<float*>(base + lufs_off);
  float* d_wave = reinterpret_cast<float*>(base + out_off);
  float* d_db = reinterpret_cast<float*>(base + out_off + wave_bytes);
  int32_t* d_status = reinterpret_cast<int32_t*>(d_db + 2 * nb);

  uint64_t w = 0;
  rc = ring_sync_mirror(h, ring, &w);
  if (rc) return rc;

  RingTickArgs a{};
  a.ring = ring->d_ring;
  a.cap = ring->cap;
  a.oldest = w % ring->cap;
  a.frames = frames;
  a.spp = window ? (double)frames / (double)window : 0.0;
  a.columns = cols;
  a.wave = d_wave;
  a.wave_blocks = (unsigned)((cols * 32 + 255) / 256 < 148 * 8 ? (cols * 32 + 255) / 256 : 148 * 8);
  a.fft_v0 = 2 * (15 * rate - n_fft);
  a.fft_n = plan ? 2 * n_fft : 0;
  a.d_fft = d_fft;
  const bool feed = *lufs_status == SSB_OK && lufs_samples;
  a.lufs_v0 = 30 * rate - lufs_samples;
  a.lufs_n = feed ? lufs_samples : 0;
  a.d_lufs = d_lufs;
  const size_t copy_values = (size_t)(a.fft_n > a.lufs_n ? a.fft_n : a.lufs_n);
  const unsigned copy_blocks = copy_values ? (unsigned)((copy_values + 1023) / 1024) : 0;
  if (a.wave_blocks + copy_blocks) {
    k_ring_tick<<<a.wave_blocks + copy_blocks, 256, 0, h->stream>>>(a);
    ++h->launches;
    CK(cudaGetLastError());
  }
  if (plan) CK(launch_fft(*plan, d_fft, SSB_FFT_MID_SIDE, 1, d_db, d_status, h->stream, &h->launches));
  if (feed) {
    rc = feed_device(h, d_lufs, lufs_samples / h->channels, lufs_samples / h->channels);
    if (rc) return rc;
  }
  const int aligned = (h->total_frames % h->lp.s100) == 0;
  if (!h->st.ring && !aligned) *lufs_status = *lufs_status ? *lufs_status : SSB_ERR_UNALIGNED_QUERY;
  rc = launch_results_now(h);
  if (rc) return rc;
  const size_t stride = 4 + 2 * (size_t)h->channels;
  CK(cudaMemcpyAsync(h->h_scratch, d_wave, wave_bytes + db_bytes, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(h->h_results, h->d_results, stride * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->results_valid = true;
  *shortterm_lufs = h->h_results[1];
  if ((h->mode & SSB_MODE_S) != SSB_MODE_S && *lufs_status == SSB_OK) *lufs_status = SSB_ERR_INVALID_MODE;

  const float* mm = static_cast<const float*>(h->h_scratch);
  for (size_t i = 0; i < cols; i++) {  // analyzer.rs:131-132: (i, min), (i, max)
    xy_wave[4 * i + 0] = (double)i;
    xy_wave[4 * i + 1] = (double)mm[2 * i];
    xy_wave[4 * i + 2] = (double)i;
    xy_wave[4 * i + 3] = (double)mm[2 * i + 1];
  }
  if (plan) {
    const float* db = reinterpret_cast<const float*>(static_cast<const char*>(h->h_scratch) + wave_bytes);
    const int32_t* st = reinterpret_cast<const int32_t*>(db + 2 * nb);
    *fft_status = st[0] ? st[0] : st[1];
    h->tick_fft_status[0] = st[0];   // the reference handles the two get_fft results independently (tui.rs:1505-1523)
    h->tick_fft_status[1] = st[1];
    const auto& axes = fft_axis_cached(h, n_fft, h->rate);
    const std::vector<double>& ax = axes.first;
    const std::vector<double>& tilt = axes.second;
    for (size_t i = 0; i < nb; i++) {
      xy_mid[2 * i] = ax[i];
      xy_mid[2 * i + 1] = (double)db[i] + tilt[i];
      xy_side[2 * i] = ax[i];
      xy_side[2 * i + 1] = (double)db[nb + i] + tilt[i];
    }
  }
  return SSB_OK;
}

}  // extern "C"
