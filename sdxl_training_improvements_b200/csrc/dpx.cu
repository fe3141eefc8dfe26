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
// Three transports share that protocol (b2_dpx_create `mode`; dp.py picks one by timing them at start-up):
//   0 "ce_pull": as above.
//   1 "ce_push": the reduce-scatter is also a PUSH — r copies its piece of peer p's shard into p's staging slot and signals
//      "pushed(c)"; NVLink then carries posted writes only (no read round trips).  Every chunk has its own staging region,
//      so a sender never needs to ask whether the receiver's slot is free.
//   2 "sm": one shared-memory-free kernel of many short 128-thread CTAs (they co-reside with the persistent GEMM /
//      attention CTAs: 0 smem, ~5 K registers) reads shard r from every peer with 16-byte loads, sums in fp32 and stores
//      the bf16 result into every peer's buffer — reduce-scatter and all-gather fused, no staging.
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
constexpr int DPX_FLAG_WORDS = 3 * DPX_MAX_CHUNKS * DPX_MAX_WORLD;  // kinds: 0 ready / pushed, 1 delivered, 2 norm published
constexpr int DPX_SQ_ACCS = 64;                                      // partial-sum accumulators of the reduce kernel
// behind the flag words of every rank's flag page: [2 parities][world] doubles = each rank's sum of squares of its shards
constexpr size_t DPX_NORM_OFF = DPX_FLAG_WORDS * sizeof(uint32_t);
constexpr size_t DPX_PAGE_BYTES = DPX_NORM_OFF + 2 * DPX_MAX_WORLD * sizeof(double);

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

// Streaming (evict-first) accesses: the exchange moves gigabytes that are touched once, while the GEMMs running beside it
// live off their operand tiles staying in the 126 MB L2.
__device__ __forceinline__ bf16x8 ld8_stream(const bf16* p) {
  bf16x8 r;
  r.u = __ldcs(reinterpret_cast<const uint4*>(p));
  return r;
}
__device__ __forceinline__ void st8_stream(bf16* p, const bf16x8& v) { __stcs(reinterpret_cast<uint4*>(p), v.u); }

// grad[i] = bf16( fp32(grad[i]) + sum_s fp32(staging[s * slot_stride + i]) ), 8 elements per 16-byte access, four
// vectors per thread.  No shared memory and few registers, so its CTAs co-reside with the persistent GEMM / attention
// CTAs instead of waiting for an SM; many short CTAs rather than a persistent grid for the same reason.
// sq_acc (optional): DPX_SQ_ACCS doubles; every warp adds the sum of squares of the bf16 values it stores — the
// global gradient norm of the REDUCED gradients falls out of the exchange (each element is summed by exactly one rank).
__global__ void __launch_bounds__(128) dpx_reduce_kernel(bf16* __restrict__ grad, const bf16* __restrict__ staging,
                                                         long long slot_stride, int nslots, long long nvec,
                                                         double* __restrict__ sq_acc) {
  const long long base = ((long long)blockIdx.x * blockDim.x) * 4 + threadIdx.x;
  float acc[4][8];
  bool live[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long v = base + (long long)u * blockDim.x;
    live[u] = v < nvec;
    if (live[u]) unpack8(ld8_stream(grad + v * 8), acc[u]);
  }
  for (int s = 0; s < nslots; ++s) {
    const bf16* src = staging + (long long)s * slot_stride;
    bf16x8 t[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long v = base + (long long)u * blockDim.x;
      if (live[u]) t[u] = ld8_stream(src + v * 8);
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
  float sq = 0.f;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long v = base + (long long)u * blockDim.x;
    if (live[u]) {
      const bf16x8 o = pack8(acc[u]);
      st8_stream(grad + v * 8, o);
      float f[8];
      unpack8(o, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) sq = fmaf(f[e], f[e], sq);
    }
  }
  if (sq_acc) {
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) atomicAdd(sq_acc + ((blockIdx.x * 4 + (threadIdx.x >> 5)) & (DPX_SQ_ACCS - 1)), (double)sq);
  }
}

struct PeerNorm {
  double* p[DPX_MAX_WORLD];
};
// one thread: this rank's sum of squares -> slot [parity][self] of every rank's flag page (own included); accumulators cleared
__global__ void dpx_norm_publish_kernel(double* sq_acc, PeerNorm slots, int world, int self, int parity) {
  double t = 0.0;
  for (int i = 0; i < DPX_SQ_ACCS; ++i) {
    t += sq_acc[i];
    sq_acc[i] = 0.0;
  }
  for (int p = 0; p < world; ++p) slots.p[p][parity * DPX_MAX_WORLD + self] = t;
  __threadfence_system();
}
// one thread: total over ranks in rank order (identical on every rank) -> out
__global__ void dpx_norm_total_kernel(const double* my_slots, int world, int parity, double* out) {
  double t = 0.0;
  for (int p = 0; p < world; ++p) t += my_slots[parity * DPX_MAX_WORLD + p];
  *out = t;
}

struct PeerBufs {
  bf16* p[DPX_MAX_WORLD];
};

// "sm" transport: out_p[i] = bf16( sum_q fp32(in_q[i]) ) for every peer p, i over this rank's shard.  Each thread owns
// its vectors from load to store, so reading and writing the same addresses in place is ordered by data dependence.
template <int W, int U>
__global__ void __launch_bounds__(128) dpx_fused_kernel(PeerBufs bufs, long long off, long long nvec) {
  const long long base = ((long long)blockIdx.x * blockDim.x) * U + threadIdx.x;
  bf16x8 t[U][W];
  bool live[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const long long v = base + (long long)u * blockDim.x;
    live[u] = v < nvec;
    if (live[u]) {
#pragma unroll
      for (int q = 0; q < W; ++q) t[u][q] = ld8_stream(bufs.p[q] + off + v * 8);
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (!live[u]) continue;
    const long long v = base + (long long)u * blockDim.x;
    float acc[8];
    unpack8(t[u][0], acc);
#pragma unroll
    for (int q = 1; q < W; ++q) {
      float f[8];
      unpack8(t[u][q], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[e];
    }
    const bf16x8 r = pack8(acc);
#pragma unroll
    for (int q = 0; q < W; ++q) st8_stream(bufs.p[q] + off + v * 8, r);
  }
}

// U vectors per thread: 2 up to four ranks, 1 beyond (register budget: the CTA must fit beside a 320-thread x 168-register
// attention CTA, i.e. stay under ~11 K registers)
template <int W>
static void launch_fused(const PeerBufs& b, long long off, long long nvec, cudaStream_t s) {
  constexpr int U = W <= 4 ? 2 : 1;
  const long long blocks = (nvec + 128 * U - 1) / (128 * U);
  dpx_fused_kernel<W, U><<<(unsigned)blocks, 128, 0, s>>>(b, off, nvec);
}
static bool launch_fused_any(int world, const PeerBufs& b, long long off, long long nvec, cudaStream_t s) {
  switch (world) {
    case 2: launch_fused<2>(b, off, nvec, s); return true;
    case 3: launch_fused<3>(b, off, nvec, s); return true;
    case 4: launch_fused<4>(b, off, nvec, s); return true;
    case 5: launch_fused<5>(b, off, nvec, s); return true;
    case 6: launch_fused<6>(b, off, nvec, s); return true;
    case 7: launch_fused<7>(b, off, nvec, s); return true;
    case 8: launch_fused<8>(b, off, nvec, s); return true;
    default: return false;
  }
}

// reduce-scatter ownership, the same arithmetic as dp.shard_plan(): a piece of `len` elements is cut into `world` shards
// of ceil(len / world) rounded up to 8 elements
static inline long long shard_size(long long len, int world) { return ((len + world - 1) / world + 7) / 8 * 8; }
static inline void shard_of(long long off, long long len, int world, int r, long long* so, long long* sl) {
  const long long ss = shard_size(len, world);
  long long lo = (long long)r * ss;
  if (lo > len) lo = len;
  long long hi = lo + ss;
  if (hi > len) hi = len;
  *so = off + lo;
  *sl = hi - lo;
}

struct Dpx {
  int mode = 0;                         // 0 ce_pull, 1 ce_push, 2 sm
  bf16* peer_staging[DPX_MAX_WORLD] = {};  // ce_push: every rank's staging base
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
  double* sq_acc = nullptr;  // DPX_SQ_ACCS partial sums of squares (reduce kernel)
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
  DPX_CUDA(cudaMalloc(&p, DPX_PAGE_BYTES));
  DPX_CUDA(cudaMemset(p, 0, DPX_PAGE_BYTES));
  DPX_CUDA(cudaDeviceSynchronize());
  *flags_out = p;
  return B2_OK;
}

extern "C" int b2_dpx_max_chunks(void) { return DPX_MAX_CHUNKS; }

// ---- object ----------------------------------------------------------------------------------------------------
extern "C" int b2_dpx_create(int rank, int world, int mode, void* const* grad_ptrs, void* const* flag_ptrs,
                             void* const* staging_ptrs, int64_t staging_slot_elems, int n_copy_streams, void** handle_out) {
  B2_REQUIRE(world >= 2 && world <= DPX_MAX_WORLD && rank >= 0 && rank < world, "b2_dpx_create: bad rank/world");
  B2_REQUIRE(mode >= 0 && mode <= 2, "b2_dpx_create: mode must be 0 (ce_pull), 1 (ce_push) or 2 (sm)");
  B2_REQUIRE(mode != 2 || world <= 8, "b2_dpx_create: the sm transport is instantiated for up to 8 ranks");
  B2_REQUIRE(grad_ptrs && flag_ptrs && staging_ptrs && handle_out && staging_slot_elems > 0, "b2_dpx_create: bad args");
  B2_REQUIRE(staging_slot_elems % 8 == 0, "b2_dpx_create: staging slot must be a multiple of 8 elements");
  Dpx* d = new Dpx();
  d->rank = rank;
  d->world = world;
  d->mode = mode;
  for (int p = 0; p < world; ++p) {
    B2_REQUIRE(grad_ptrs[p] && flag_ptrs[p], "b2_dpx_create: null peer pointer");
    B2_REQUIRE(((uintptr_t)grad_ptrs[p] & 15) == 0, "b2_dpx_create: gradient buffers must be 16-byte aligned");
    B2_REQUIRE(mode == 2 || (staging_ptrs[p] && ((uintptr_t)staging_ptrs[p] & 15) == 0) || (mode == 0 && p != rank),
               "b2_dpx_create: staging pointers must be 16-byte aligned (ce_push needs every rank's)");
    d->grad[p] = (bf16*)grad_ptrs[p];
    d->flags[p] = (uint32_t*)flag_ptrs[p];
    d->peer_staging[p] = (bf16*)staging_ptrs[p];
  }
  d->staging = (bf16*)staging_ptrs[rank];
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
  DPX_CUDA(cudaMalloc(&d->sq_acc, DPX_SQ_ACCS * sizeof(double)));
  DPX_CUDA(cudaMemset(d->sq_acc, 0, DPX_SQ_ACCS * sizeof(double)));
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
  cudaFree(d->sq_acc);
  delete d;
  return B2_OK;
}

namespace b2 {
static int dpx_wait_flag(Dpx* d, cudaStream_t s, int kind, int chunk, int src, uint32_t seq) {
  CUresult r = d->wait32((CUstream)s, (CUdeviceptr)(d->flags[d->rank] + flag_word(kind, chunk, src)), seq,
                         CU_STREAM_WAIT_VALUE_GEQ);
  if (r != CUDA_SUCCESS) {
    set_error("dpx: cuStreamWaitValue32 failed (%d)", (int)r);
    return B2_ERR_CUDA;
  }
  return B2_OK;
}
static int dpx_signal(Dpx* d, int kind, int chunk, uint32_t seq) {
  PeerFlags pf;
  for (int p = 0; p < DPX_MAX_WORLD; ++p) pf.p[p] = p < d->world ? d->flags[p] : nullptr;
  dpx_signal_kernel<<<1, 32, 0, d->xs>>>(pf, d->world, d->rank, flag_word(kind, chunk, d->rank), seq);
  return check_launch("dpx_signal");
}
// every copy stream waits for what has been enqueued on xs so far
static int dpx_fork(Dpx* d) {
  DPX_CUDA(cudaEventRecord(d->ev_red, d->xs));
  for (auto s : d->cs) DPX_CUDA(cudaStreamWaitEvent(s, d->ev_red, 0));
  return B2_OK;
}
// xs waits for everything enqueued on the copy streams so far
static int dpx_join(Dpx* d) {
  for (size_t i = 0; i < d->cs.size(); ++i) {
    DPX_CUDA(cudaEventRecord(d->ev_cs[i], d->cs[i]));
    DPX_CUDA(cudaStreamWaitEvent(d->xs, d->ev_cs[i], 0));
  }
  return B2_OK;
}
}  // namespace b2

// Exchange (sum over ranks) of one chunk: `n_ranges` pieces [range_off, range_off + range_len) of the gradient buffer
// (elements, multiples of 8).  Each piece is cut into `world` shards (shard_of); rank r owns shard r.  staging_base:
// element offset of this chunk's region inside a staging slot (ce_push: regions of different chunks must not overlap).
// Work issued before the call on `main_stream` is ordered before the exchange; nothing is made to wait for it.
extern "C" int b2_dpx_exchange(void* handle, int chunk, uint32_t seq, int n_ranges, const int64_t* range_off,
                               const int64_t* range_len, int64_t staging_base, void* main_stream) {
  Dpx* d = (Dpx*)handle;
  B2_REQUIRE(d && chunk >= 0 && chunk < DPX_MAX_CHUNKS && n_ranges >= 0 && staging_base >= 0 && staging_base % 8 == 0,
             "b2_dpx_exchange: bad args");
  const int W = d->world, R = d->rank, ncs = (int)d->cs.size();
  long long used = staging_base;
  for (int i = 0; i < n_ranges; ++i) {
    B2_REQUIRE(range_off[i] % 8 == 0 && range_len[i] % 8 == 0 && range_len[i] >= 0,
               "b2_dpx_exchange: piece %d is not 16-byte granular", i);
    used += shard_size(range_len[i], W);
  }
  B2_REQUIRE(d->mode == 2 || used <= d->slot_stride, "b2_dpx_exchange: chunk does not fit its staging slot");
  if (!d->pending.empty() && d->pending_seq != seq) {
    set_error("b2_dpx_exchange: chunks of sequence %u are still pending (call b2_dpx_finish first)", d->pending_seq);
    return B2_ERR_ARG;
  }
  int rc;
  DPX_CUDA(cudaEventRecord(d->ev_in, (cudaStream_t)main_stream));
  DPX_CUDA(cudaStreamWaitEvent(d->xs, d->ev_in, 0));

  if (d->mode == 2) {
    // ---- sm: ready -> wait all -> fused reduce + broadcast kernel -> delivered
    if ((rc = dpx_signal(d, 0, chunk, seq))) return rc;
    for (int p = 0; p < W; ++p)
      if (p != R && (rc = dpx_wait_flag(d, d->xs, 0, chunk, p, seq))) return rc;
    PeerBufs pb;
    for (int q = 0; q < DPX_MAX_WORLD; ++q) pb.p[q] = q < W ? d->grad[(R + q) % W] : nullptr;  // own copy first
    for (int i = 0; i < n_ranges; ++i) {
      long long so, sl;
      shard_of(range_off[i], range_len[i], W, R, &so, &sl);
      if (sl == 0) continue;
      if (!launch_fused_any(W, pb, so, sl / 8, d->xs)) {
        set_error("b2_dpx_exchange: sm transport has no instantiation for %d ranks", W);
        return B2_ERR_ARG;
      }
      if ((rc = check_launch("dpx_fused"))) return rc;
    }
    if ((rc = dpx_signal(d, 1, chunk, seq))) return rc;
    d->pending.push_back(chunk);
    d->pending_seq = seq;
    return B2_OK;
  }

  if (d->mode == 0) {
    // ---- ce_pull: ready -> (per peer: wait its ready, pull my shard from it)
    if ((rc = dpx_signal(d, 0, chunk, seq))) return rc;
    if ((rc = dpx_fork(d))) return rc;
    for (int j = 0; j < W - 1; ++j) {
      const int p = (R + 1 + j) % W;
      cudaStream_t s = d->cs[j % ncs];
      if ((rc = dpx_wait_flag(d, s, 0, chunk, p, seq))) return rc;
      long long so_stage = staging_base;
      for (int i = 0; i < n_ranges; ++i) {
        long long so, sl;
        shard_of(range_off[i], range_len[i], W, R, &so, &sl);
        if (sl)
          DPX_CUDA(cudaMemcpyAsync(d->staging + (long long)j * d->slot_stride + so_stage, d->grad[p] + so,
                                   (size_t)sl * sizeof(bf16), cudaMemcpyDefault, s));
        so_stage += shard_size(range_len[i], W);
      }
    }
    if ((rc = dpx_join(d))) return rc;
  } else {
    // ---- ce_push: push my copy of peer p's shard into p's staging slot, then tell everyone; wait for everyone's pushes.
    // My slot at receiver p is j = (R - p - 1) mod W  (slot j of rank p holds the copy of rank (p + 1 + j) mod W).
    if ((rc = dpx_fork(d))) return rc;
    for (int jj = 0; jj < W - 1; ++jj) {
      const int p = (R + 1 + jj) % W;
      const int j = ((R - p - 1) % W + W) % W;
      cudaStream_t s = d->cs[jj % ncs];
      long long so_stage = staging_base;
      for (int i = 0; i < n_ranges; ++i) {
        long long so, sl;
        shard_of(range_off[i], range_len[i], W, p, &so, &sl);
        if (sl)
          DPX_CUDA(cudaMemcpyAsync(d->peer_staging[p] + (long long)j * d->slot_stride + so_stage, d->grad[R] + so,
                                   (size_t)sl * sizeof(bf16), cudaMemcpyDefault, s));
        so_stage += shard_size(range_len[i], W);
      }
    }
    if ((rc = dpx_join(d))) return rc;
    if ((rc = dpx_signal(d, 0, chunk, seq))) return rc;  // "pushed(c)"
    for (int p = 0; p < W; ++p)
      if (p != R && (rc = dpx_wait_flag(d, d->xs, 0, chunk, p, seq))) return rc;
  }
  // ---- reduce my shard (staged copies of every peer are in my slots)
  {
    long long so_stage = staging_base;
    for (int i = 0; i < n_ranges; ++i) {
      long long so, sl;
      shard_of(range_off[i], range_len[i], W, R, &so, &sl);
      if (sl) {
        const long long nvec = sl / 8;
        const long long blocks = (nvec + 128 * 4 - 1) / (128 * 4);
        dpx_reduce_kernel<<<(unsigned)blocks, 128, 0, d->xs>>>(d->grad[R] + so, d->staging + so_stage, d->slot_stride, W - 1,
                                                               nvec, d->sq_acc);
        if ((rc = check_launch("dpx_reduce"))) return rc;
      }
      so_stage += shard_size(range_len[i], W);
    }
  }
  // ---- all-gather by copy engine: push the reduced shard into every peer's gradient buffer, then "delivered(c)"
  if ((rc = dpx_fork(d))) return rc;
  for (int j = 0; j < W - 1; ++j) {
    const int p = (R + 1 + j) % W;
    cudaStream_t s = d->cs[j % ncs];
    for (int i = 0; i < n_ranges; ++i) {
      long long so, sl;
      shard_of(range_off[i], range_len[i], W, R, &so, &sl);
      if (sl)
        DPX_CUDA(cudaMemcpyAsync(d->grad[p] + so, d->grad[R] + so, (size_t)sl * sizeof(bf16), cudaMemcpyDefault, s));
    }
  }
  if ((rc = dpx_join(d))) return rc;
  if ((rc = dpx_signal(d, 1, chunk, seq))) return rc;
  d->pending.push_back(chunk);
  d->pending_seq = seq;
  return B2_OK;
}

// Make `main_stream` wait until every chunk exchanged with `seq` has been delivered into the local gradient buffer by
// every peer (and this rank's own pulls / pushes are complete, so the buffer may be overwritten afterwards).
// gnorm_sq_out (optional, copy-engine transports only): receives sum over ALL ranks' shards of (reduced bf16 gradient)^2, i.e.
// what b2_sumsq over the exchanged buffer would return, without reading the buffer again: partial sums come out of the reduce
// kernel, every rank publishes its own into every peer's flag page, and all ranks add them in rank order (bit-identical).
extern "C" int b2_dpx_finish_norm(void* handle, uint32_t seq, void* main_stream, double* gnorm_sq_out);
extern "C" int b2_dpx_finish(void* handle, uint32_t seq, void* main_stream) {
  return b2_dpx_finish_norm(handle, seq, main_stream, nullptr);
}
extern "C" int b2_dpx_norm_supported(void* handle) {
  Dpx* d = (Dpx*)handle;
  return d && d->mode != 2 ? 1 : 0;
}
extern "C" int b2_dpx_finish_norm(void* handle, uint32_t seq, void* main_stream, double* gnorm_sq_out) {
  Dpx* d = (Dpx*)handle;
  B2_REQUIRE(d, "b2_dpx_finish: bad args");
  B2_REQUIRE(!gnorm_sq_out || d->mode != 2, "b2_dpx_finish_norm: the sm transport does not produce the gradient norm");
  if (d->pending.empty()) {
    B2_REQUIRE(!gnorm_sq_out, "b2_dpx_finish_norm: nothing was exchanged, there is no norm to return");
    return B2_OK;
  }
  B2_REQUIRE(d->pending_seq == seq, "b2_dpx_finish: sequence %u does not match the pending exchange %u", seq, d->pending_seq);
  int rc;
  for (int chunk : d->pending)
    for (int p = 0; p < d->world; ++p)
      if (p != d->rank && (rc = dpx_wait_flag(d, d->xs, 1, chunk, p, seq))) return rc;
  d->pending.clear();
  if (gnorm_sq_out) {
    const int parity = (int)(seq & 1u);
    PeerNorm pn;
    for (int p = 0; p < DPX_MAX_WORLD; ++p)
      pn.p[p] = p < d->world ? reinterpret_cast<double*>(reinterpret_cast<char*>(d->flags[p]) + DPX_NORM_OFF) : nullptr;
    dpx_norm_publish_kernel<<<1, 1, 0, d->xs>>>(d->sq_acc, pn, d->world, d->rank, parity);
    if ((rc = check_launch("dpx_norm_publish"))) return rc;
    if ((rc = dpx_signal(d, 2, 0, seq))) return rc;
    for (int p = 0; p < d->world; ++p)
      if (p != d->rank && (rc = dpx_wait_flag(d, d->xs, 2, 0, p, seq))) return rc;
    dpx_norm_total_kernel<<<1, 1, 0, d->xs>>>(pn.p[d->rank], d->world, parity, gnorm_sq_out);
    if ((rc = check_launch("dpx_norm_total"))) return rc;
  } else if (d->mode != 2) {
    DPX_CUDA(cudaMemsetAsync(d->sq_acc, 0, DPX_SQ_ACCS * sizeof(double), d->xs));  // self-test / autotune / callers without clip
  }
  DPX_CUDA(cudaEventRecord(d->ev_done, d->xs));
  DPX_CUDA(cudaStreamWaitEvent((cudaStream_t)main_stream, d->ev_done, 0));
  return B2_OK;
}

// Raw async copy between two device pointers of this process (local or IPC-mapped): tools/dp_copy_probe.py times the
// copy engines with it.
extern "C" int b2_dpx_memcpy_async(void* dst, const void* src, int64_t bytes, void* stream) {
  B2_REQUIRE(dst && src && bytes > 0, "b2_dpx_memcpy_async: bad args");
  DPX_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return B2_OK;
}
