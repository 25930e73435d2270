// gather.cu — the one exchange step of the multi-GPU path: every rank's per-stream result rows written straight into
// every other rank's gather buffer over NVLink by the kernel that computes them (loudness_results.cuh), no collective
// kernel, no SMs taken from the filter kernels.
//
// streams shard over ranks with no data-path exchange; "the only collective is the gather of the
// per-stream result struct".  One cudaMalloc per rank holds [2 parities][world][n_streams][stride] f64 rows + one flag
// word per rank; the allocation is exported with cudaIpcGetMemHandle, the 64-byte handles travel through the caller's
// process group (torch.distributed / MPI / files), every rank opens the others' (cudaIpcOpenMemHandle maps the peer
// memory: stores then ride NVLink / NVSwitch).  A results launch on rank r writes its rows into block r of the selected
// parity on EVERY rank with plain fire-and-forget stores; ssb_gather_wait enqueues a one-warp kernel (after the results
// launches in stream order) that fences, bumps this rank's flag on every rank and spins until every rank's flag has
// reached the same wait count.  Double buffering by parity is the caller's (sharding.PeerGather flips it after every wait).
#include "ssb_handle.cuh"

using namespace ssb;

namespace ssb {

struct GatherFlags {
  unsigned long long* peer[kMaxGatherRanks];   // per rank: its flag word for THIS rank
};

// One warp, after the last results launch in stream order: lane r tells rank r that this rank's rows up to `epoch` have
// been stored (the kernel boundary has retired those stores; the system-scope fence orders them before the flag), then
// waits until rank r has said the same.
__global__ void k_gather_wait(const unsigned long long* flags, const __grid_constant__ GatherFlags gf, int world,
                              unsigned long long epoch) {
  const int r = threadIdx.x;
  if (r >= world) return;
  __threadfence_system();
  *reinterpret_cast<volatile unsigned long long*>(gf.peer[r]) = epoch;
  const volatile unsigned long long* f = flags + r;
  unsigned long long spins = 0;
  while (*f < epoch) {
    __nanosleep(200);
    if (++spins > (1ull << 26)) __trap();   // a peer that never arrives: fail loudly instead of hanging the box
  }
  __threadfence_system();
}

}  // namespace ssb

extern "C" {

int32_t ssb_gather_create(ssb_analyzer* h, uint32_t world, uint32_t rank, void* ipc_handle_out) {
  if (!h || !ipc_handle_out || world < 1 || world > (uint32_t)kMaxGatherRanks || rank >= world) return SSB_ERR_INVALID_ARG;
  if (!h->meter_ok) return fail(h, SSB_ERR_NOMEM, "the loudness meter is not initialised");
  if (h->gather.base) return fail(h, SSB_ERR_INVALID_ARG, "gather already created");
  DeviceGuard g(h->device);
  const size_t stride = 4 + 2 * (size_t)h->channels;
  const size_t rows_bytes = 2 * (size_t)world * h->n_streams * stride * sizeof(double);
  const size_t total = rows_bytes + 256;
  CK(cudaMalloc(&h->gather.base, total));
  CK(cudaMemset(h->gather.base, 0, total));
  h->gather.world = (int)world;
  h->gather.rank = (int)rank;
  h->gather.rows_bytes = rows_bytes;
  h->gather.epoch = 0;
  h->gather.parity = 0;
  h->gather.peer_base[rank] = h->gather.base;
  cudaIpcMemHandle_t mh;
  CK(cudaIpcGetMemHandle(&mh, h->gather.base));
  static_assert(sizeof(mh) == 64, "the ABI exchanges 64-byte handles");
  memcpy(ipc_handle_out, &mh, sizeof(mh));
  CK(cudaDeviceSynchronize());
  return SSB_OK;
}

int32_t ssb_gather_open(ssb_analyzer* h, const void* ipc_handles) {
  if (!h || !ipc_handles) return SSB_ERR_INVALID_ARG;
  if (!h->gather.base) return fail(h, SSB_ERR_INVALID_ARG, "ssb_gather_create first");
  DeviceGuard g(h->device);
  for (int p = 0; p < h->gather.world; p++) {
    if (p == h->gather.rank) continue;
    cudaIpcMemHandle_t mh;
    memcpy(&mh, static_cast<const char*>(ipc_handles) + (size_t)p * sizeof(mh), sizeof(mh));
    void* ptr = nullptr;
    CK(cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess));
    h->gather.peer_base[p] = ptr;
  }
  h->gather.open = true;
  return SSB_OK;
}

int32_t ssb_gather_select(ssb_analyzer* h, int32_t parity) {
  if (!h || !h->gather.base) return SSB_ERR_INVALID_ARG;
  h->gather.parity = parity & 1;
  return SSB_OK;
}

double* ssb_gather_rows(ssb_analyzer* h, int32_t parity) {
  if (!h || !h->gather.base) return nullptr;
  return reinterpret_cast<double*>(static_cast<char*>(h->gather.base) + (size_t)(parity & 1) * (h->gather.rows_bytes / 2));
}

uint64_t ssb_gather_epoch(const ssb_analyzer* h) { return h ? h->gather.epoch : 0; }

int32_t ssb_gather_wait(ssb_analyzer* h) {
  if (!h || !h->gather.base) return SSB_ERR_INVALID_ARG;
  if (!h->gather.open && h->gather.world > 1) return fail(h, SSB_ERR_INVALID_ARG, "ssb_gather_open first");
  DeviceGuard g(h->device);
  const unsigned long long* flags =
      reinterpret_cast<const unsigned long long*>(static_cast<char*>(h->gather.base) + h->gather.rows_bytes);
  GatherFlags gf{};
  for (int p = 0; p < h->gather.world; p++)
    gf.peer[p] = reinterpret_cast<unsigned long long*>(static_cast<char*>(h->gather.peer_base[p]) + h->gather.rows_bytes) + h->gather.rank;
  ++h->gather.epoch;   // one per wait: every rank calls ssb_gather_wait the same number of times
  k_gather_wait<<<1, 32, 0, h->stream>>>(flags, gf, h->gather.world, h->gather.epoch);
  h->launches++;
  CK(cudaGetLastError());
  return SSB_OK;
}

int32_t ssb_gather_destroy(ssb_analyzer* h) {
  if (!h) return SSB_ERR_INVALID_ARG;
  if (!h->gather.base) return SSB_OK;
  DeviceGuard g(h->device);
  cudaStreamSynchronize(h->stream);
  for (int p = 0; p < h->gather.world; p++)
    if (p != h->gather.rank && h->gather.peer_base[p]) cudaIpcCloseMemHandle(h->gather.peer_base[p]);
  cudaFree(h->gather.base);
  h->gather = ssb_analyzer::Gather{};
  return SSB_OK;
}

}  // extern "C"

namespace ssb {

// The GatherArgs of this handle's results launches (world == 0 when no gather is open).
GatherArgs peek_gather_args(ssb_analyzer* h) {
  GatherArgs ga{};
  if (!h->gather.base || (!h->gather.open && h->gather.world > 1)) return ga;
  const size_t stride = 4 + 2 * (size_t)h->channels;
  const size_t half = h->gather.rows_bytes / 2;
  ga.world = h->gather.world;
  ga.rank = h->gather.rank;
  for (int p = 0; p < ga.world; p++) {
    char* base = static_cast<char*>(h->gather.peer_base[p]);
    ga.rows[p] = reinterpret_cast<double*>(base + (size_t)h->gather.parity * half) + (size_t)ga.rank * h->n_streams * stride;
  }
  return ga;
}

}  // namespace ssb
