// b2_gemm: warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[z][m,n] = alpha * sum_k A[z][m,k] * B[z][n,k] (+bias) (+residual) (+D_old)
//
// One 128 x BN output tile per CTA.  Warp 0 = TMA producer (cp.async.bulk.tensor.4d, SWIZZLE_128B boxes),
// warp 1 = TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, fp32 accumulators in TMEM),
// warps 2..5 = epilogue (tcgen05.ld 32x32b -> registers -> fused bias/residual/accumulate -> 16-byte stores).
// smem ring of STAGES x (A 128x64 + B BNx64) bf16 tiles guarded by full/empty mbarriers; tcgen05.commit
// releases a stage back to the producer and finally signals the epilogue.  Both operands may be K-major or
// MN-major (runtime flag -> different TMA box + UMMA smem-descriptor strides), which is what lets one kernel
// serve Linear fwd (K,K), dgrad (K,MN) and wgrad (MN,MN) without any transposed copies.
#include "tc.cuh"

namespace b2 {

// ---------------------------------------------------------------------------------------------
// Kernel
// ---------------------------------------------------------------------------------------------
struct GemmEpi {
  void* D;
  const bf16* bias;
  const bf16* R;
  int M, N, K;
  int nb_lo;
  long long ldd, ldr;
  long long d_bs_lo, d_bs_hi, r_bs_lo, r_bs_hi;
  float alpha;
  int accumulate, out_fp32;
  int bias_rows_per_group;
  long long bias_group_stride;
  int a_mn, b_mn;
};

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_TILE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int MN_BLOCK_BYTES = BK * 128;   // one 64-wide MN-major block: BK k-rows of 128 B

template <int BN>
constexpr int stage_bytes() { return A_TILE_BYTES + BN * BK * 2; }

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, (BN <= 128 ? 2 : 1))
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmEpi p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[STAGES];
  __shared__ __align__(8) uint64_t bar_empty[STAGES];
  __shared__ __align__(8) uint64_t bar_acc;
  __shared__ uint32_t tmem_slot;

  constexpr int STAGE = stage_bytes<BN>();
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024-B aligned tiles
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * BM;
  const int zlo = blockIdx.z % p.nb_lo;
  const int zhi = blockIdx.z / p.nb_lo;
  const int num_kb = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_acc), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
        const uint32_t full = smem_u32(&bar_full[s]);
        mbar_expect_tx(full, STAGE);
        const uint32_t sa = smem_base + s * STAGE;
        const uint32_t sb = sa + A_TILE_BYTES;
        const int k0 = kb * BK;
        if (!p.a_mn) {
          tma_load_4d(sa, &tmA, full, k0, m0, zlo, zhi);
        } else {
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) tma_load_4d(sa + j * MN_BLOCK_BYTES, &tmA, full, m0 + 64 * j, k0, zlo, zhi);
        }
        if (!p.b_mn) {
          tma_load_4d(sb, &tmB, full, k0, n0, zlo, zhi);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_4d(sb + j * MN_BLOCK_BYTES, &tmB, full, n0 + 64 * j, k0, zlo, zhi);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) /*D=f32*/ | (1u << 7) /*A=bf16*/ | (1u << 10) /*B=bf16*/ |
                             (uint32_t(p.a_mn) << 15) | (uint32_t(p.b_mn) << 16) | (uint32_t(BN >> 3) << 17) |
                             (uint32_t(BM >> 4) << 24);
      const uint32_t a_lbo = p.a_mn ? MN_BLOCK_BYTES : 16, a_kstep = p.a_mn ? 2048 : 32;
      const uint32_t b_lbo = p.b_mn ? MN_BLOCK_BYTES : 16, b_kstep = p.b_mn ? 2048 : 32;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(smem_u32(&bar_full[s]), ph);
        tc_fence_after();
        const uint32_t sa = smem_base + s * STAGE;
        const uint32_t sb = sa + A_TILE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t ad = umma_desc(sa + k * a_kstep, a_lbo, 1024);
          const uint64_t bd = umma_desc(sb + k * b_kstep, b_lbo, 1024);
          umma_bf16(tmem_base, ad, bd, idesc, (kb | k) != 0);
        }
        umma_commit(smem_u32(&bar_empty[s]));  // stage reusable once these MMAs have read it
      }
      umma_commit(smem_u32(&bar_acc));  // accumulator complete
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may touch
    mbar_wait(smem_u32(&bar_acc), 0);
    tc_fence_after();
    const int gm = m0 + q * 32 + lane;
    const bool row_ok = gm < p.M;
    const long long doff = (long long)zlo * p.d_bs_lo + (long long)zhi * p.d_bs_hi + (long long)gm * p.ldd;
    const long long roff = (long long)zlo * p.r_bs_lo + (long long)zhi * p.r_bs_hi + (long long)gm * p.ldr;
    const bf16* bias_row = p.bias ? p.bias + (long long)(gm / p.bias_rows_per_group) * p.bias_group_stride : nullptr;
    bf16* Db = reinterpret_cast<bf16*>(p.D);
    float* Df = reinterpret_cast<float*>(p.D);
    const bool vec_d = !p.out_fp32 && ((p.ldd & 7) == 0) && ((p.d_bs_lo & 7) == 0) && ((p.d_bs_hi & 7) == 0) &&
                       ((reinterpret_cast<uintptr_t>(p.D) & 15) == 0);
    const bool vec_r = !p.R || (((p.ldr & 7) == 0) && ((p.r_bs_lo & 7) == 0) && ((p.r_bs_hi & 7) == 0) &&
                                ((reinterpret_cast<uintptr_t>(p.R) & 15) == 0));
    const bool vec_b = !p.bias || (((p.bias_group_stride & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0));
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t raw[32];
      tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(c * 32), raw);
      const int gn = n0 + c * 32;
      if (!row_ok || gn >= p.N) continue;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int n = gn + g * 8;
        if (n >= p.N) break;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(raw[g * 8 + j]) * p.alpha;
        if (n + 8 <= p.N && vec_d && vec_r && vec_b) {
          float t[8];
          if (bias_row) {
            unpack8(ld8(bias_row + n), t);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] += t[j];
          }
          if (p.R) {
            unpack8(ld8(p.R + roff + n), t);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] += t[j];
          }
          if (p.accumulate) {
            unpack8(ld8(Db + doff + n), t);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] += t[j];
          }
          st8(Db + doff + n, pack8(f));
        } else {
          for (int j = 0; j < 8 && n + j < p.N; ++j) {
            float x = f[j];
            if (bias_row) x += __bfloat162float(bias_row[n + j]);
            if (p.R) x += __bfloat162float(p.R[roff + n + j]);
            if (p.out_fp32) {
              if (p.accumulate) x += Df[doff + n + j];
              Df[doff + n + j] = x;
            } else {
              if (p.accumulate) x += __bfloat162float(Db[doff + n + j]);
              Db[doff + n + j] = __float2bfloat16(x);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// ---------------------------------------------------------------------------------------------
// Host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------
// Operand described as logical [rows(MN), K] per batch; mn_major selects which is contiguous.
static int make_operand_map(CUtensorMap* tm, const void* ptr, int mn_extent, int k_extent, int mn_major, long long ld,
                            int nb_lo, long long bs_lo, int nb_hi, long long bs_hi, int box_mn, const char* name) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return B2_ERR_TMAP;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld & 7) || (nb_lo > 1 && (bs_lo & 7)) || (nb_hi > 1 && (bs_hi & 7))) {
    set_error("b2_gemm: operand %s violates TMA alignment (ptr %p ld %lld bs %lld %lld)", name, ptr, ld, bs_lo, bs_hi);
    return B2_ERR_ARG;
  }
  cuuint64_t dims[4];
  cuuint64_t strides[3];
  cuuint32_t box[4] = {64, 1, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (!mn_major) {
    dims[0] = (cuuint64_t)k_extent;
    dims[1] = (cuuint64_t)mn_extent;
    box[1] = (cuuint32_t)box_mn;
  } else {
    dims[0] = (cuuint64_t)mn_extent;
    dims[1] = (cuuint64_t)k_extent;
    box[1] = 64;
  }
  dims[2] = (cuuint64_t)nb_lo;
  dims[3] = (cuuint64_t)nb_hi;
  const long long min_stride = 16;
  strides[0] = (cuuint64_t)(ld * 2);
  strides[1] = (cuuint64_t)((nb_lo > 1 ? bs_lo * 2 : min_stride));
  strides[2] = (cuuint64_t)((nb_hi > 1 ? bs_hi * 2 : min_stride));
  if (strides[0] < 16) strides[0] = 16;
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s) failed: CUresult %d (dims %llu %llu %llu %llu, strides %llu %llu %llu)", name,
              (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
              (unsigned long long)dims[3], (unsigned long long)strides[0], (unsigned long long)strides[1],
              (unsigned long long)strides[2]);
    return B2_ERR_TMAP;
  }
  return B2_OK;
}

template <int BN, int STAGES>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpi& e, int nbatch, cudaStream_t st) {
  constexpr int smem = STAGES * stage_bytes<BN>() + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) {
      set_error("b2_gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(err));
      return B2_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid((e.N + BN - 1) / BN, (e.M + BM - 1) / BM, nbatch);
  gemm_tcgen05_kernel<BN, STAGES><<<grid, 192, smem, st>>>(ta, tb, e);
  return check_launch("b2_gemm");
}

int gemm2_try(const b2_gemm_args* a, cudaStream_t st, int* out_rc);  // gemm2.cu: persistent CTA-pair fast path
bool smallm_try(const b2_gemm_args* a, cudaStream_t st, int* out_rc);  // smallm.cu: few-row linears on CUDA cores

}  // namespace b2

extern "C" int b2_gemm(const b2_gemm_args* a, void* stream) {
  using namespace b2;
  B2_REQUIRE(a && a->A && a->B && a->D, "b2_gemm: null pointer");
  B2_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "b2_gemm: bad shape %d %d %d", a->M, a->N, a->K);
  B2_REQUIRE(a->nb_lo >= 1 && a->nb_hi >= 1 && (long long)a->nb_lo * a->nb_hi <= 65535, "b2_gemm: bad batch");
  {
    int rc2 = 0;
    if (smallm_try(a, reinterpret_cast<cudaStream_t>(stream), &rc2)) return rc2;
    if (gemm2_try(a, reinterpret_cast<cudaStream_t>(stream), &rc2)) return rc2;
  }
  int bn = a->tile_n;
  if (bn == 0) bn = (a->N <= 64) ? 64 : 128;
  B2_REQUIRE(bn == 64 || bn == 128 || bn == 256, "b2_gemm: tile_n must be 64/128/256");
  B2_REQUIRE((a->M + BM - 1) / BM <= 65535, "b2_gemm: M too large");

  CUtensorMap ta, tb;
  int rc = make_operand_map(&ta, a->A, a->M, a->K, a->a_mn, a->lda, a->nb_lo, a->a_bs_lo, a->nb_hi, a->a_bs_hi, BM, "A");
  if (rc) return rc;
  rc = make_operand_map(&tb, a->B, a->N, a->K, a->b_mn, a->ldb, a->nb_lo, a->b_bs_lo, a->nb_hi, a->b_bs_hi, bn, "B");
  if (rc) return rc;

  GemmEpi e;
  e.D = a->D;
  e.bias = reinterpret_cast<const bf16*>(a->bias);
  e.R = reinterpret_cast<const bf16*>(a->residual);
  e.M = a->M; e.N = a->N; e.K = a->K;
  e.nb_lo = a->nb_lo;
  e.ldd = a->ldd; e.ldr = a->ldr;
  e.d_bs_lo = a->d_bs_lo; e.d_bs_hi = a->d_bs_hi; e.r_bs_lo = a->r_bs_lo; e.r_bs_hi = a->r_bs_hi;
  e.alpha = a->alpha;
  e.accumulate = a->accumulate; e.out_fp32 = a->out_fp32;
  e.bias_rows_per_group = a->bias_rows_per_group > 0 ? a->bias_rows_per_group : 0x7fffffff;
  e.bias_group_stride = a->bias_group_stride;
  e.a_mn = a->a_mn ? 1 : 0; e.b_mn = a->b_mn ? 1 : 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nb = a->nb_lo * a->nb_hi;
  if (bn == 64) return launch_gemm<64, 4>(ta, tb, e, nb, st);
  if (bn == 128) return launch_gemm<128, 3>(ta, tb, e, nb, st);
  return launch_gemm<256, 4>(ta, tb, e, nb, st);
}
