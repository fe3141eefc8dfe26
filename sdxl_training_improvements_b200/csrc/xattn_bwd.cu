// b2_xattn_bwd: cross-attention backward (<= 80 keys) in ONE pass — dQ, dK and dV from a single recomputation of P.
//
//   S = Q K^T, P = exp2(S c - LSE), dP = dO V^T, dS = P * (dP - D) with D_i = sum_d dO_id O_id = sum_j P_ij dP_ij
//   (the second form needs no O: with ONE key block the whole row of P and dP is in this thread's registers),
//   dQ = scale * dS K,   dK = scale * dS^T Q,   dV = P^T dO            (per sample b and 64-channel head h)
//
// replaces, for attn2 of every BasicTransformerBlock (autograd of F.scaled_dot_product_attention in diffusers'
// AttnProcessor2_0; loss.backward() at /root/reference/src/training/trainers/methods/ddpm_trainer.py:271): the generic
// two-kernel backward (attn_bwd_dq3 + attn_bwd_dkv3), which for 77 keys launched 640 CTAs that each set up TMEM / barriers
// for two 64-key blocks plus 80 CTAs that walked all queries again, and evaluated every exponential twice:
// 41-46 us per layer call at n = 1024, 91-99 us at n = 4096 (profiles/r2_attention_notes.md).
//
// Why one pass works here: with one key block the whole dS^T / P^T for a query tile fits in shared memory, so after the
// softmax warps (thread = query row) have written bf16 P and dS rows into two swizzled [128 queries x 80 keys] tiles, the SAME
// tiles serve as  the K-major A operand of dQ = dS K  and as the MN-major A operand (M = keys) of dK += dS^T Q and
// dV += P^T dO — the transposition is a descriptor, not a data movement.  dK / dV accumulate in TMEM over all query tiles of
// the (sample, head) pair the CTA owns and are written once: no atomics, deterministic.
//
// One CTA per (b, h) (SDXL at B = 4: 80 CTAs at C = 1280, 40 at C = 640; the work per pair is small and HBM-light, the
// kernel is latency-bound).  Warps (192 threads): 0 = TMA producer, 1 = MMA issuer + TMEM allocator,
// 2..5 = softmax / epilogue (warp w owns TMEM lanes 32*(w%4)..+31).
// TMEM (512 columns): S | dP slots [0,160) and [160,320) (80 columns each), dQ [320,384), dK [384,448), dV [448,512).
#include <math.h>
#include <stdlib.h>

#include "tc.cuh"

namespace b2 {

constexpr int XB_THREADS = 192;
constexpr int XB_T = 128 * 64 * 2;              // a [128 rows x 64 bf16] swizzled tile: 16 KiB
constexpr int XB_KV = 80 * 64 * 2;              // K or V: 80 keys x 128 B = 10 KiB
constexpr int XB_STAGES = 3;                    // {Q, dO} tiles per query tile
constexpr int XB_OFF_KV = XB_STAGES * 2 * XB_T;           // 96 KiB
constexpr int XB_OFF_P = XB_OFF_KV + 2 * XB_KV;           // 116 KiB: P  = block 0 (keys 0..63) | block 1 (keys 64..79 in the first 32 B of each row)
constexpr int XB_OFF_DS = XB_OFF_P + 2 * XB_T;            // 148 KiB: dS, same layout
constexpr int XB_SMEM = XB_OFF_DS + 2 * XB_T + 1024;      // 181 KiB

struct XbP {
  int H, n_q, n_k, n_pad;
  float c, scale;
  const float* LSE;
  bf16 *dQ, *dK, *dV;
  long long lddq, dq_bs, lddk, dk_bs, lddv, dv_bs;
};

__global__ void __launch_bounds__(XB_THREADS, 1)
xattn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                 const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const XbP p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_kv, bar_full[XB_STAGES], bar_empty[XB_STAGES], bar_s[2], bar_sfree[2], bar_pd, bar_pdfree,
      bar_dq, bar_dqfree, bar_final;
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int h = blockIdx.x % p.H, b = blockIdx.x / p.H;
  const int nqt = (p.n_q + 127) / 128;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(smem_u32(&bar_kv), 1);
#pragma unroll
    for (int s = 0; s < XB_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_s[i]), 1);
      mbar_init(smem_u32(&bar_sfree[i]), 128);
    }
    mbar_init(smem_u32(&bar_pd), 128);
    mbar_init(smem_u32(&bar_pdfree), 1);
    mbar_init(smem_u32(&bar_dq), 1);
    mbar_init(smem_u32(&bar_dqfree), 128);
    mbar_init(smem_u32(&bar_final), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(tmem_slot);
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ---------------- TMA producer: K, V once; {Q, dO} per query tile
    const bool el = elect_one();
    if (el) {
      mbar_expect_tx(smem_u32(&bar_kv), 2 * XB_KV);
      tma_load_4d(smem_base + XB_OFF_KV, &tmK, smem_u32(&bar_kv), 0, 0, h, b);
      tma_load_4d(smem_base + XB_OFF_KV + XB_KV, &tmV, smem_u32(&bar_kv), 0, 0, h, b);
    }
    int s = 0;
    uint32_t ph = 0;
    for (int t = 0; t < nqt; ++t) {
      mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
      if (el) {
        const uint32_t full = smem_u32(&bar_full[s]);
        const uint32_t st = smem_base + s * 2 * XB_T;
        mbar_expect_tx(full, 2 * XB_T);
        tma_load_4d(st, &tmQ, full, 0, t * 128, h, b);
        tma_load_4d(st + XB_T, &tmdO, full, 0, t * 128, h, b);
      }
      if (++s == XB_STAGES) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer
    const bool el = elect_one();
    constexpr uint32_t idS = umma_idesc(128, 80, 0, 0);    // S / dP: A K-major (Q / dO), B K-major (K / V), N = 80 keys
    constexpr uint32_t idQ = umma_idesc(128, 64, 0, 1);    // dQ = dS K: A K-major (dS), B MN-major (K as [key][d])
    constexpr uint32_t idK = umma_idesc(128, 64, 1, 1);    // dK += dS^T Q, dV += P^T dO: A MN-major (M = keys), B MN-major
    const uint64_t dKk = umma_desc(smem_base + XB_OFF_KV, 16, 1024);            // K, K-major B
    const uint64_t dVk = umma_desc(smem_base + XB_OFF_KV + XB_KV, 16, 1024);    // V, K-major B
    const uint64_t dKm = umma_desc(smem_base + XB_OFF_KV, 8192, 1024);          // K, MN-major B: k-step = 16 keys = 2048 B
    const uint64_t dSa = umma_desc(smem_base + XB_OFF_DS, 16, 1024);            // dS, K-major A (block 1 at + XB_T)
    const uint64_t dSm = umma_desc(smem_base + XB_OFF_DS, XB_T, 1024);          // dS, MN-major A: second 64-key block LBO = XB_T apart
    const uint64_t dPm = umma_desc(smem_base + XB_OFF_P, XB_T, 1024);           // P,  MN-major A
    mbar_wait(smem_u32(&bar_kv), 0);
    auto issue_SdP = [&](int t) {
      const int s = t % XB_STAGES, slot = t & 1;
      mbar_wait(smem_u32(&bar_full[s]), (uint32_t)(t / XB_STAGES) & 1u);
      if (t >= 2) mbar_wait(smem_u32(&bar_sfree[slot]), (uint32_t)((t >> 1) - 1) & 1u);  // softmax has read this slot's last S / dP
      tc_fence_after();
      const uint32_t tS = tmem_base + slot * 160, tdP = tS + 80;
      const uint64_t dq = umma_desc(smem_base + s * 2 * XB_T, 16, 1024), ddo = desc_adv(dq, XB_T);
      if (el) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tS, desc_adv(dq, k * 32), desc_adv(dKk, k * 32), idS, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tdP, desc_adv(ddo, k * 32), desc_adv(dVk, k * 32), idS, k != 0);
        umma_commit(smem_u32(&bar_s[slot]));
      }
    };
    issue_SdP(0);
    for (int t = 0; t < nqt; ++t) {
      const int s = t % XB_STAGES;
      if (t + 1 < nqt) issue_SdP(t + 1);  // run-ahead: the next tile's logits form while the softmax warps work on this one
      mbar_wait(smem_u32(&bar_pd), (uint32_t)t & 1u);                       // P, dS of tile t are in shared memory
      if (t > 0) mbar_wait(smem_u32(&bar_dqfree), (uint32_t)(t - 1) & 1u);  // dQ of tile t-1 has been read out
      tc_fence_after();
      const uint64_t dqm = umma_desc(smem_base + s * 2 * XB_T, 8192, 1024);  // Q tile, MN-major B ([q][d], k-step = 16 queries)
      const uint64_t ddom = desc_adv(dqm, XB_T);                             // dO tile, MN-major B
      if (el) {
        // dQ = dS K: K-dim = 80 keys: four k-steps inside block 0 (32 B apart), the fifth in block 1
#pragma unroll
        for (int k = 0; k < 5; ++k)
          umma_bf16(tmem_base + 320, desc_adv(dSa, k < 4 ? k * 32 : XB_T), desc_adv(dKm, k * 2048), idQ, k != 0);
        umma_commit(smem_u32(&bar_dq));
        // dK += dS^T Q, dV += P^T dO: M = 128 key rows (80 real), K-dim = 128 queries = 8 k-steps of 16 rows (2048 B)
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tmem_base + 384, desc_adv(dSm, k * 2048), desc_adv(dqm, k * 2048), idK, (t | k) != 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tmem_base + 448, desc_adv(dPm, k * 2048), desc_adv(ddom, k * 2048), idK, (t | k) != 0);
        umma_commit(smem_u32(&bar_empty[s]));   // {Q, dO} stage free
        umma_commit(smem_u32(&bar_pdfree));     // P / dS tiles free
      }
    }
    if (el) umma_commit(smem_u32(&bar_final));
  } else {
    // ---------------- softmax / epilogue warps: thread = query row (tiles), later = key row (dK / dV)
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = uint32_t(qd * 32) << 16;
    const uint32_t sw = (uint32_t)(row & 7);
    const long long lse_base = ((long long)b * p.H + h) * p.n_pad;
    uint32_t pp[40], pd[40];  // bf16 P and dS of the tile being prepared (this thread's query row, 80 keys)
    // P / dS of tile t into registers: row constants, the stage's dO . O dot product, S / dP from TMEM, the exponentials
    float L_next = p.LSE[lse_base + row];  // LSE of this thread's row in the NEXT tile to prepare: loaded one tile ahead, its
                                           // ~1 us global-load latency is otherwise on the per-tile critical path
    auto compute_pd = [&](int t) {
      const int slot = t & 1;
      const float L2 = L_next;                // n_pad is a multiple of 128: in bounds; +inf on pad rows -> P = 0
      if (t + 1 < nqt) L_next = p.LSE[lse_base + (t + 1) * 128 + row];
      mbar_wait(smem_u32(&bar_s[slot]), (uint32_t)(t >> 1) & 1u);
      tc_fence_after();
      const uint32_t tS = tmem_base + slot * 160 + lane_off, tdP = tS + 80;
      uint32_t rs[80], rd[80];
      tmem_ld32_nowait(tS, rs);
      tmem_ld32_nowait(tS + 32, rs + 32);
      tmem_ld16_nowait(tS + 64, rs + 64);
      tmem_ld32_nowait(tdP, rd);
      tmem_ld32_nowait(tdP + 32, rd + 32);
      tmem_ld16_nowait(tdP + 64, rd + 64);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_sfree[slot]));
      // pass 1: P (fp32, kept in place of S) and D = sum_j P_j dP_j;  pass 2: dS = P (dP - D)
      const float2 c2 = make_float2(p.c, p.c), nl2 = make_float2(-L2, -L2);
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int j = 0; j < 80; j += 2) {
        const float2 x = __ffma2_rn(make_float2(__uint_as_float(rs[j]), __uint_as_float(rs[j + 1])), c2, nl2);
        float p0 = fast_exp2(x.x), p1 = fast_exp2(x.y);
        if (j >= p.n_k) p0 = 0.f;       // padding keys (K rows are TMA zero fill: S = 0, not -inf)
        if (j + 1 >= p.n_k) p1 = 0.f;
        rs[j] = __float_as_uint(p0);
        rs[j + 1] = __float_as_uint(p1);
        d0 = fmaf(p0, __uint_as_float(rd[j]), d0);
        d1 = fmaf(p1, __uint_as_float(rd[j + 1]), d1);
        pp[j / 2] = pack_bf16x2(p0, p1);
      }
      const float Dr = d0 + d1;
      const float2 nd2 = make_float2(-Dr, -Dr);
#pragma unroll
      for (int j = 0; j < 80; j += 2) {
        const float2 ds = __fmul2_rn(make_float2(__uint_as_float(rs[j]), __uint_as_float(rs[j + 1])),
                                     __fadd2_rn(make_float2(__uint_as_float(rd[j]), __uint_as_float(rd[j + 1])), nd2));
        pd[j / 2] = pack_bf16x2(ds.x, ds.y);
      }
    };
    // registers -> the swizzled P / dS tiles (operands of the dQ / dK / dV MMAs), then signal the MMA thread
    auto write_pd = [&]() {
      const uint32_t prow = smem_base + XB_OFF_P + row * 128, drow = smem_base + XB_OFF_DS + row * 128;
#pragma unroll
      for (int cidx = 0; cidx < 8; ++cidx) {  // block 0: keys 0..63
        const uint32_t off = (cidx ^ sw) << 4;
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(prow + off), "r"(pp[4 * cidx]), "r"(pp[4 * cidx + 1]),
                     "r"(pp[4 * cidx + 2]), "r"(pp[4 * cidx + 3]) : "memory");
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(drow + off), "r"(pd[4 * cidx]), "r"(pd[4 * cidx + 1]),
                     "r"(pd[4 * cidx + 2]), "r"(pd[4 * cidx + 3]) : "memory");
      }
#pragma unroll
      for (int cidx = 0; cidx < 2; ++cidx) {  // block 1: keys 64..79 (chunks 0, 1 of the row); the rest of the row is never used
        const uint32_t off = (cidx ^ sw) << 4;
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(prow + XB_T + off), "r"(pp[32 + 4 * cidx]),
                     "r"(pp[33 + 4 * cidx]), "r"(pp[34 + 4 * cidx]), "r"(pp[35 + 4 * cidx]) : "memory");
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(drow + XB_T + off), "r"(pd[32 + 4 * cidx]),
                     "r"(pd[33 + 4 * cidx]), "r"(pd[34 + 4 * cidx]), "r"(pd[35 + 4 * cidx]) : "memory");
      }
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&bar_pd));
    };
    // Software pipeline: while the tensor pipe works on tile t (dQ, dK +=, dV +=), this thread already prepares P / dS of
    // tile t + 1 in registers; it then drains dQ(t) and, once the MMAs of tile t have released the tiles, writes P / dS(t+1).
    compute_pd(0);
    write_pd();
    for (int t = 0; t < nqt; ++t) {
      const int gq = t * 128 + row;
      if (t + 1 < nqt) compute_pd(t + 1);
      // ---- dQ of tile t
      mbar_wait(smem_u32(&bar_dq), (uint32_t)t & 1u);
      tc_fence_after();
      uint32_t q0[32], q1[32];
      tmem_ld32_nowait(tmem_base + 320 + lane_off, q0);
      tmem_ld32_nowait(tmem_base + 352 + lane_off, q1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_dqfree));
      if (t + 1 < nqt) {
        mbar_wait(smem_u32(&bar_pdfree), (uint32_t)t & 1u);  // the MMAs of tile t have read the P / dS tiles
        write_pd();
      }
      // the dQ rows leave AFTER the hand-off above: an mbarrier arrive has release semantics and would wait for these stores
      if (gq < p.n_q) {
        bf16* dst = p.dQ + (long long)b * p.dq_bs + (long long)gq * p.lddq + h * 64;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const uint32_t* v = g < 4 ? q0 + g * 8 : q1 + (g - 4) * 8;
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[j]) * p.scale;
          st8(dst + g * 8, pack8(f));
        }
      }
    }
    // ---- dK / dV of this (b, h): TMEM lane = key row
    mbar_wait(smem_u32(&bar_final), 0);
    tc_fence_after();
    uint32_t k0[32], k1[32], v0[32], v1[32];
    tmem_ld32_nowait(tmem_base + 384 + lane_off, k0);
    tmem_ld32_nowait(tmem_base + 416 + lane_off, k1);
    tmem_ld32_nowait(tmem_base + 448 + lane_off, v0);
    tmem_ld32_nowait(tmem_base + 480 + lane_off, v1);
    tmem_ld_wait();
    if (row < p.n_k) {
      bf16* dk = p.dK + (long long)b * p.dk_bs + (long long)row * p.lddk + h * 64;
      bf16* dv = p.dV + (long long)b * p.dv_bs + (long long)row * p.lddv + h * 64;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const uint32_t* a = g < 4 ? k0 + g * 8 : k1 + (g - 4) * 8;
        const uint32_t* c = g < 4 ? v0 + g * 8 : v1 + (g - 4) * 8;
        float fk[8], fv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { fk[j] = __uint_as_float(a[j]) * p.scale; fv[j] = __uint_as_float(c[j]); }
        st8(dk + g * 8, pack8(fk));
        st8(dv + g * 8, pack8(fv));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_xattn_bwd_ok(int B, int H, int n_q, int n_k) {
  if (getenv("B2_XATTN_BWD_GENERIC")) return 0;
  return B > 0 && H > 0 && n_q > 0 && n_k > 0 && n_k <= 80 && (long long)B * H <= 65535 * 32ll;
}

extern "C" int b2_xattn_bwd(const b2_attn_args* a, void* stream) {
  B2_REQUIRE(a && a->Q && a->K && a->V && a->O && a->LSE && a->dO && a->dQ && a->dK && a->dV, "b2_xattn_bwd: null pointer");
  B2_REQUIRE(b2_xattn_bwd_ok(a->B, a->H, a->n_q, a->n_k), "b2_xattn_bwd: unsupported shape B=%d H=%d n_q=%d n_k=%d", a->B, a->H,
             a->n_q, a->n_k);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CUtensorMap tq, tdo, tk, tv;
  int rc;
  if ((rc = make_map_bf16_4d(&tq, a->Q, 64, a->n_q, a->H, a->B, a->ldq, 64, a->q_bs, 64, 128, "xbwd Q"))) return rc;
  if ((rc = make_map_bf16_4d(&tdo, a->dO, 64, a->n_q, a->H, a->B, a->lddo, 64, a->do_bs, 64, 128, "xbwd dO"))) return rc;
  if ((rc = make_map_bf16_4d(&tk, a->K, 64, a->n_k, a->H, a->B, a->ldk, 64, a->k_bs, 64, 80, "xbwd K"))) return rc;
  if ((rc = make_map_bf16_4d(&tv, a->V, 64, a->n_k, a->H, a->B, a->ldv, 64, a->v_bs, 64, 80, "xbwd V"))) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(xattn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XB_SMEM);
    if (err != cudaSuccess) {
      set_error("b2_xattn_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(err));
      return B2_ERR_CUDA;
    }
    configured = true;
  }
  XbP p{};
  p.H = a->H; p.n_q = a->n_q; p.n_k = a->n_k; p.n_pad = (a->n_q + 127) / 128 * 128;
  p.scale = a->scale; p.c = a->scale * 1.4426950408889634f;
  p.LSE = a->LSE;
  p.dQ = (bf16*)a->dQ; p.lddq = a->lddq; p.dq_bs = a->dq_bs;
  p.dK = (bf16*)a->dK; p.lddk = a->lddk; p.dk_bs = a->dk_bs;
  p.dV = (bf16*)a->dV; p.lddv = a->lddv; p.dv_bs = a->dv_bs;
  cudaError_t le = launch_pdl(xattn_bwd_kernel, dim3(a->B * a->H), dim3(XB_THREADS), (size_t)XB_SMEM, st, tq, tdo, tk, tv, p);
  if (le != cudaSuccess) {
    set_error("b2_xattn_bwd: launch: %s", cudaGetErrorString(le));
    return B2_ERR_CUDA;
  }
  return check_launch("b2_xattn_bwd");
}
