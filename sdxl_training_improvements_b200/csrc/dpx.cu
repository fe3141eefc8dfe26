// Data-parallel gradient exchange over NVSwitch peer memory, overlapped with the backward pass.
//
// Replaces DistributedDataParallel's bucketed NCCL all-reduce (reference: src/core/distributed.py:142-163,
// `convert_model_to_ddp`) for the flat bf16 gradient buffer.  Why not NCCL here: the GEMM / attention kernels are
// persistent, one CTA (pair) per SM with a static tile schedule, so an NCCL kernel that takes even a few SMs during the
// backward pass doubles the time of every GEMM it overlaps.  This exchange needs NO resident SMs:
//
//   chunk c of the gradient buffer is final (rank r has passed cut c of its backward pass)
//     1. r signals "ready(c)" into every peer's flag page (one tiny kernel: st.release.sys into IPC-mapped memory);
//     2. reduce-scatter by COPY ENGINE: r pulls shard r of chunk c from every peer into local staging slots
//        (cudaMemcpyAsync peer -> local, one stream per few peers, each gated by cuStreamWaitValue32 on that peer's
//        ready flag — a stream memory op, no SM);
//     3. one short, shared-memory-free reduce kernel: grad[shard r] = bf16(fp32(grad) + sum of the staged shards);
//     4. all-gather by COPY ENGINE: r pushes its reduced shard into every peer's gradient buffer, then signals
//        "delivered(c)".
//   Before the optimizer step the main stream waits (stream memory op) for delivered(c) from every peer for every chunk.
//
// All of it runs on side streams owned by this object; the only coupling to the compute stream is one event at the cut
// and one event before the optimizer.  Every rank ends with bit-identical reduced gradients (the owner of a shard is the
// only one that sums it).  Host-side runtime in C++ behind the C ABI (include/sdxl_b200.h, "b2_dpx_*").
#include <cuda.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace b2 {

constexpr int DPX_MAX_WORLD = 16;
constexpr int DPX_MAX_CHUNKS = 64;
constexpr int DPX_FLAG_WORDS = 2 * DPX_MAX_CHUNKS * DPX_MAX_WORLD;

typedef CUresult (*WaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*GetAddressRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);

static void* driver_entry(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult st;
  if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return fn;
}

struct PeerFlags {
  uint32_t* p[DPX_MAX_WORLD];
};

// One warp; lane l < world writes `value` into peer l's flag word.  Runs after the stream's preceding copies completed,
// the fence + release store make them visible to the peer's stream memory op / kernels before the flag is.
__global__ void dpx_signal_kernel(PeerFlags flags, int world, int self, int word, uint32_t value) {
  const int l = threadIdx.x;
  if (l < world && l != self) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags.p[l] + word), "r"(value) : "memory");
  }
}

// grad[i] = bf16( fp32(grad[i]) + sum_s fp32(staging[s * slot_stride + i]) ), 8 elements per 16-byte access, four
// vectors per thread.  No shared memory and few registers, so its CTAs co-reside with the persistent GEMM / attention
// CTAs instead of waiting for an SM; many short CTAs rather than a persistent grid for the same reason.
__global__ void __launch_bounds__(128) dpx_reduce_kernel(bf16* __restrict__ grad, const bf16* __restrict__ staging,
                                                         long long slot_stride, int nslots, long long nvec) {
  const long long base = ((long long)blockIdx.x * blockDim.x) * 4 + threadIdx.x;
  float acc[4][8];
  bool live[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long v = base + (long long)u * blockDim.x;
    live[u] = v < nvec;
    if (live[u]) unpack8(ld8(grad + v * 8), acc[u]);
  }
  for (int s = 0; s < nslots; ++s) {
    const bf16* src = staging + (long long)s * slot_stride;
    bf16x8 t[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long v = base + (long long)u * blockDim.x;
      if (live[u]) t[u] = ld8(src + v * 8);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (live[u]) {
        float f[8];
        unpack8(t[u], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[u][e] += f[e];
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long v = base + (long long)u * blockDim.x;
    if (live[u]) st8(grad + v * 8, pack8(acc[u]));
  }
}

struct Dpx {
  int rank = 0, world = 1;
  bf16* grad[DPX_MAX_WORLD] = {};      // [rank] = local buffer, others IPC-mapped
  uint32_t* flags[DPX_MAX_WORLD] = {};  // [rank] = local flag page
  bf16* staging = nullptr;
  long long slot_stride = 0;
  cudaStream_t xs = nullptr;            // orchestration stream (signals, reduce kernel)
  std::vector<cudaStream_t> cs;         // copy streams
  cudaEvent_t ev_in = nullptr, ev_red = nullptr, ev_done = nullptr;
  std::vector<cudaEvent_t> ev_cs;
  WaitValue32Fn wait32 = nullptr;
  std::vector<int> pending;  // chunks exchanged since the last finish()
  uint32_t pending_seq = 0;
};

static inline int flag_word(int kind, int chunk, int src) { return (kind * DPX_MAX_CHUNKS + chunk) * DPX_MAX_WORLD + src; }

#define DPX_CUDA(call)                                                       \
  do {                                                                       \
    cudaError_t e_ = (call);                                                 \
    if (e_ != cudaSuccess) {                                                 \
      cudaGetLastError();                                                    \
      set_error("dpx: %s: %s", #call, cudaGetErrorString(e_));               \
      return B2_ERR_CUDA;                                                    \
    }                                                                        \
  } while (0)

static std::mutex g_ipc_mu;
static std::map<std::string, void*> g_ipc_open;  // handle bytes -> mapped base (a handle may be opened once per process)

}  // namespace b2

using namespace b2;

// ---- IPC plumbing ----------------------------------------------------------------------------------------------
extern "C" int b2_dpx_ipc_export(const void* dev_ptr, unsigned char* handle_out /*64 bytes*/, int64_t* offset_out) {
  B2_REQUIRE(dev_ptr && handle_out && offset_out, "b2_dpx_ipc_export: bad args");
  static GetAddressRangeFn get_range = (GetAddressRangeFn)driver_entry("cuMemGetAddressRange");
  B2_REQUIRE(get_range, "b2_dpx_ipc_export: cuMemGetAddressRange not available");
  CUdeviceptr base = 0;
  size_t size = 0;
  CUresult r = get_range(&base, &size, (CUdeviceptr)dev_ptr);
  if (r != CUDA_SUCCESS) {
    set_error("b2_dpx_ipc_export: cuMemGetAddressRange failed (%d)", (int)r);
    return B2_ERR_CUDA;
  }
  cudaIpcMemHandle_t h;
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  DPX_CUDA(cudaIpcGetMemHandle(&h, (void*)base));
  memcpy(handle_out, &h, 64);
  *offset_out = (int64_t)((CUdeviceptr)dev_ptr - base);
  return B2_OK;
}

extern "C" int b2_dpx_ipc_import(const unsigned char* handle /*64 bytes*/, int64_t offset, void** dev_ptr_out) {
  B2_REQUIRE(handle && dev_ptr_out && offset >= 0, "b2_dpx_ipc_import: bad args");
  std::lock_guard<std::mutex> lk(g_ipc_mu);
  std::string key((const char*)handle, 64);
  auto it = g_ipc_open.find(key);
  void* base = nullptr;
  if (it != g_ipc_open.end()) {
    base = it->second;
  } else {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    DPX_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    g_ipc_open[key] = base;
  }
  *dev_ptr_out = (char*)base + offset;
  return B2_OK;
}

// The flag page is this library's own cudaMalloc (never a sub-block of a caching allocator), zero-initialised.
extern "C" int b2_dpx_alloc_flags(void** flags_out) {
  B2_REQUIRE(flags_out, "b2_dpx_alloc_flags: bad args");
  void* p = nullptr;
  DPX_CUDA(cudaMalloc(&p, DPX_FLAG_WORDS * sizeof(uint32_t)));
  DPX_CUDA(cudaMemset(p, 0, DPX_FLAG_WORDS * sizeof(uint32_t)));
  DPX_CUDA(cudaDeviceSynchronize());
  *flags_out = p;
  return B2_OK;
}

extern "C" int b2_dpx_max_chunks(void) { return DPX_MAX_CHUNKS; }

// ---- object ----------------------------------------------------------------------------------------------------
extern "C" int b2_dpx_create(int rank, int world, void* const* grad_ptrs, void* const* flag_ptrs, void* staging,
                             int64_t staging_slot_elems, int n_copy_streams, void** handle_out) {
  B2_REQUIRE(world >= 2 && world <= DPX_MAX_WORLD && rank >= 0 && rank < world, "b2_dpx_create: bad rank/world");
  B2_REQUIRE(grad_ptrs && flag_ptrs && staging && handle_out && staging_slot_elems > 0, "b2_dpx_create: bad args");
  B2_REQUIRE(staging_slot_elems % 8 == 0 && ((uintptr_t)staging & 15) == 0, "b2_dpx_create: staging must be 16-byte aligned");
  Dpx* d = new Dpx();
  d->rank = rank;
  d->world = world;
  for (int p = 0; p < world; ++p) {
    B2_REQUIRE(grad_ptrs[p] && flag_ptrs[p], "b2_dpx_create: null peer pointer");
    B2_REQUIRE(((uintptr_t)grad_ptrs[p] & 15) == 0, "b2_dpx_create: gradient buffers must be 16-byte aligned");
    d->grad[p] = (bf16*)grad_ptrs[p];
    d->flags[p] = (uint32_t*)flag_ptrs[p];
  }
  d->staging = (bf16*)staging;
  d->slot_stride = staging_slot_elems;
  d->wait32 = (WaitValue32Fn)driver_entry("cuStreamWaitValue32");
  if (!d->wait32) {
    delete d;
    set_error("b2_dpx_create: cuStreamWaitValue32 not available from this driver");
    return B2_ERR_CUDA;
  }
  int ncs = n_copy_streams > 0 ? n_copy_streams : (world - 1 < 4 ? world - 1 : 4);
  if (ncs > world - 1) ncs = world - 1;
  DPX_CUDA(cudaStreamCreateWithFlags(&d->xs, cudaStreamNonBlocking));
  d->cs.resize(ncs);
  d->ev_cs.resize(ncs);
  for (int i = 0; i < ncs; ++i) {
    DPX_CUDA(cudaStreamCreateWithFlags(&d->cs[i], cudaStreamNonBlocking));
    DPX_CUDA(cudaEventCreateWithFlags(&d->ev_cs[i], cudaEventDisableTiming));
  }
  DPX_CUDA(cudaEventCreateWithFlags(&d->ev_in, cudaEventDisableTiming));
  DPX_CUDA(cudaEventCreateWithFlags(&d->ev_red, cudaEventDisableTiming));
  DPX_CUDA(cudaEventCreateWithFlags(&d->ev_done, cudaEventDisableTiming));
  *handle_out = d;
  return B2_OK;
}

extern "C" int b2_dpx_destroy(void* handle) {
  Dpx* d = (Dpx*)handle;
  if (!d) return B2_OK;
  cudaStreamSynchronize(d->xs);
  for (auto s : d->cs) {
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
  }
  cudaStreamDestroy(d->xs);
  for (auto e : d->ev_cs) cudaEventDestroy(e);
  cudaEventDestroy(d->ev_in);
  cudaEventDestroy(d->ev_red);
  cudaEventDestroy(d->ev_done);
  delete d;
  return B2_OK;
}

// Exchange (sum over ranks) of this rank's view of one chunk: `n_ranges` pieces of the gradient buffer, for each the
// element offset / length of THIS RANK'S SHARD (multiples of 8; length may be 0) and its offset inside a staging slot.
// Work issued before the call on `main_stream` is ordered before the exchange; nothing is made to wait for it.
extern "C" int b2_dpx_exchange(void* handle, int chunk, uint32_t seq, int n_ranges, const int64_t* shard_off,
                               const int64_t* shard_len, const int64_t* staging_off, void* main_stream) {
  Dpx* d = (Dpx*)handle;
  B2_REQUIRE(d && chunk >= 0 && chunk < DPX_MAX_CHUNKS && n_ranges >= 0, "b2_dpx_exchange: bad args");
  for (int i = 0; i < n_ranges; ++i)
    B2_REQUIRE(shard_off[i] % 8 == 0 && shard_len[i] % 8 == 0 && staging_off[i] % 8 == 0 && shard_len[i] >= 0 &&
                   staging_off[i] + shard_len[i] <= d->slot_stride,
               "b2_dpx_exchange: shard %d not 16-byte granular or outside the staging slot", i);
  if (!d->pending.empty() && d->pending_seq != seq) {
    set_error("b2_dpx_exchange: chunks of sequence %u are still pending (call b2_dpx_finish first)", d->pending_seq);
    return B2_ERR_ARG;
  }
  const int W = d->world, R = d->rank, ncs = (int)d->cs.size();
  PeerFlags pf;
  for (int p = 0; p < DPX_MAX_WORLD; ++p) pf.p[p] = p < W ? d->flags[p] : nullptr;

  DPX_CUDA(cudaEventRecord(d->ev_in, (cudaStream_t)main_stream));
  DPX_CUDA(cudaStreamWaitEvent(d->xs, d->ev_in, 0));
  // 1. ready(c): my part of chunk c is final
  dpx_signal_kernel<<<1, 32, 0, d->xs>>>(pf, W, R, flag_word(0, chunk, R), seq);
  if (int rc = check_launch("dpx_signal(ready)")) return rc;
  DPX_CUDA(cudaEventRecord(d->ev_red, d->xs));  // reused below; here: "ready signalled" (orders copy streams after the cut)
  // 2. reduce-scatter by copy engine: slot j holds peer (R + 1 + j) % W's copy of my shard
  for (int i = 0; i < ncs; ++i) DPX_CUDA(cudaStreamWaitEvent(d->cs[i], d->ev_red, 0));
  for (int j = 0; j < W - 1; ++j) {
    const int p = (R + 1 + j) % W;
    cudaStream_t s = d->cs[j % ncs];
    CUresult r = d->wait32((CUstream)s, (CUdeviceptr)(d->flags[R] + flag_word(0, chunk, p)), seq, CU_STREAM_WAIT_VALUE_GEQ);
    if (r != CUDA_SUCCESS) {
      set_error("b2_dpx_exchange: cuStreamWaitValue32 failed (%d)", (int)r);
      return B2_ERR_CUDA;
    }
    for (int i = 0; i < n_ranges; ++i) {
      if (shard_len[i] == 0) continue;
      DPX_CUDA(cudaMemcpyAsync(d->staging + (long long)j * d->slot_stride + staging_off[i], d->grad[p] + shard_off[i],
                               (size_t)shard_len[i] * sizeof(bf16), cudaMemcpyDefault, s));
    }
  }
  for (int i = 0; i < ncs; ++i) {
    DPX_CUDA(cudaEventRecord(d->ev_cs[i], d->cs[i]));
    DPX_CUDA(cudaStreamWaitEvent(d->xs, d->ev_cs[i], 0));
  }
  // 3. reduce my shard
  for (int i = 0; i < n_ranges; ++i) {
    if (shard_len[i] == 0) continue;
    const long long nvec = shard_len[i] / 8;
    const long long blocks = (nvec + 128 * 4 - 1) / (128 * 4);
    dpx_reduce_kernel<<<(unsigned)blocks, 128, 0, d->xs>>>(d->grad[R] + shard_off[i], d->staging + staging_off[i],
                                                           d->slot_stride, W - 1, nvec);
    if (int rc = check_launch("dpx_reduce")) return rc;
  }
  DPX_CUDA(cudaEventRecord(d->ev_red, d->xs));
  // 4. all-gather by copy engine: push the reduced shard into every peer's buffer
  for (int i = 0; i < ncs; ++i) DPX_CUDA(cudaStreamWaitEvent(d->cs[i], d->ev_red, 0));
  for (int j = 0; j < W - 1; ++j) {
    const int p = (R + 1 + j) % W;
    cudaStream_t s = d->cs[j % ncs];
    for (int i = 0; i < n_ranges; ++i) {
      if (shard_len[i] == 0) continue;
      DPX_CUDA(cudaMemcpyAsync(d->grad[p] + shard_off[i], d->grad[R] + shard_off[i], (size_t)shard_len[i] * sizeof(bf16),
                               cudaMemcpyDefault, s));
    }
  }
  for (int i = 0; i < ncs; ++i) {
    DPX_CUDA(cudaEventRecord(d->ev_cs[i], d->cs[i]));
    DPX_CUDA(cudaStreamWaitEvent(d->xs, d->ev_cs[i], 0));
  }
  dpx_signal_kernel<<<1, 32, 0, d->xs>>>(pf, W, R, flag_word(1, chunk, R), seq);
  if (int rc = check_launch("dpx_signal(delivered)")) return rc;
  d->pending.push_back(chunk);
  d->pending_seq = seq;
  return B2_OK;
}

// Make `main_stream` wait until every chunk exchanged with `seq` has been delivered into the local gradient buffer by
// every peer (and this rank's own pulls / pushes are complete, so the buffer may be overwritten afterwards).
extern "C" int b2_dpx_finish(void* handle, uint32_t seq, void* main_stream) {
  Dpx* d = (Dpx*)handle;
  B2_REQUIRE(d, "b2_dpx_finish: bad args");
  if (d->pending.empty()) return B2_OK;
  B2_REQUIRE(d->pending_seq == seq, "b2_dpx_finish: sequence %u does not match the pending exchange %u", seq, d->pending_seq);
  for (int chunk : d->pending) {
    for (int p = 0; p < d->world; ++p) {
      if (p == d->rank) continue;
      CUresult r = d->wait32((CUstream)d->xs, (CUdeviceptr)(d->flags[d->rank] + flag_word(1, chunk, p)), seq,
                             CU_STREAM_WAIT_VALUE_GEQ);
      if (r != CUDA_SUCCESS) {
        set_error("b2_dpx_finish: cuStreamWaitValue32 failed (%d)", (int)r);
        return B2_ERR_CUDA;
      }
    }
  }
  d->pending.clear();
  DPX_CUDA(cudaEventRecord(d->ev_done, d->xs));
  DPX_CUDA(cudaStreamWaitEvent((cudaStream_t)main_stream, d->ev_done, 0));
  return B2_OK;
}
