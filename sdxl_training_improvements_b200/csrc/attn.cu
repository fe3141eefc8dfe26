// Fused attention for sm_100a: softmax(Q K^T * scale) V, head dim 64, bf16 in/out, fp32 accumulation in TMEM.
// Replaces F.scaled_dot_product_attention inside diffusers' AttnProcessor2_0 (self-attention, n keys; cross-attention,
// 77 keys) and its autograd backward — SURVEY.md §8 a5.5 / a5.6.  No n x n tensor ever touches HBM.
//
// Three kernels with one shape: TMA-staged bf16 tiles in SWIZZLE_128B smem -> tcgen05.mma (SS) into TMEM -> the
// four "row" warps read their TMEM lane with tcgen05.ld, do the exp2 / rescale math in registers, write the bf16
// result back to TMEM with tcgen05.st -> tcgen05.mma (TS: A operand from TMEM) accumulates the output tile in TMEM.
//   fwd    : CTA = 128 queries x (all keys in blocks of 128);  S -> P -> O += P V, online softmax with lazy rescale.
//   bwd_dq : CTA = 128 queries x (keys in blocks of 64);        S, dP -> dS -> dQ += dS K.
//   bwd_dkv: CTA = 128 keys x (queries in blocks of 64);        S^T, dP^T -> P^T, dS^T -> dV += P^T dO, dK += dS^T Q.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = row warps
// (warp w owns TMEM lanes 32*(w%4)..+31).  256 TMEM columns and <= 82 KB smem per CTA -> two CTAs per SM, which is
// what overlaps one CTA's exp2 phase with the other CTA's MMA phase.
//
// LSE is kept in the log2 domain: L2[i] = m_i + log2(sum_j 2^(s_ij*c - m_i)), c = scale*log2(e); P_ij = 2^(s_ij*c - L2[i]).
// LSE / D are laid out [B, H, n_pad] with n_pad = ceil(n_q/128)*128; pad rows hold L2 = +inf (P = 0) and D = 0.
#include <math.h>
#include <stdlib.h>

#include "tc.cuh"

namespace b2 {

constexpr int AT_THREADS = 192;
constexpr int AT_D = 64;                        // head dim
constexpr int AT_TILE128 = 128 * AT_D * 2;      // 16 KiB: 128 rows x 64 bf16
constexpr int AT_TILE64 = 64 * AT_D * 2;        // 8 KiB
constexpr int AT_TMEM_COLS = 256;
constexpr uint32_t AT_FULL = 0xffffffffu;

struct AttnP {
  int H, n_q, n_k, n_pad;
  float c;      // scale * log2(e)
  float scale;
  float* LSE;   // [B,H,n_pad]
  float* D;     // [B,H,n_pad]
  bf16* out0;   // fwd: O ; bwd_dq: dQ ; bwd_dkv: dK
  bf16* out1;   //                        bwd_dkv: dV
  long long ld0, bs0, ld1, bs1;
};

__device__ __forceinline__ void store_row64(bf16* dst, const uint32_t* r0, const uint32_t* r1, float mul) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const uint32_t* r = g < 4 ? r0 + g * 8 : r1 + (g - 4) * 8;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(r[j]) * mul;
    st8(dst + g * 8, pack8(f));
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
constexpr int FWD_STAGES = 2;
constexpr int FWD_SMEM = AT_TILE128 + FWD_STAGES * 2 * AT_TILE128 + 1024;

__global__ void __launch_bounds__(AT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnP p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_q, bar_full[FWD_STAGES], bar_empty[FWD_STAGES], bar_s, bar_p, bar_o;
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sKV = smem_base + AT_TILE128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nkb = (p.n_k + 127) / 128;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(smem_u32(&bar_q), 1);
#pragma unroll
    for (int s = 0; s < FWD_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_s), 1);
    mbar_init(smem_u32(&bar_p), 128);
    mbar_init(smem_u32(&bar_o), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), AT_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t tS = tmem_base, tP = tmem_base + 128, tO = tmem_base + 192;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(smem_u32(&bar_q), AT_TILE128);
      tma_load_4d(sQ, &tmQ, smem_u32(&bar_q), 0, q0, h, b);
      for (int j = 0; j < nkb; ++j) {
        const int s = j % FWD_STAGES;
        const uint32_t ph = (j / FWD_STAGES) & 1;
        mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
        const uint32_t full = smem_u32(&bar_full[s]);
        mbar_expect_tx(full, 2 * AT_TILE128);
        tma_load_4d(sKV + s * 2 * AT_TILE128, &tmK, full, 0, j * 128, h, b);
        tma_load_4d(sKV + s * 2 * AT_TILE128 + AT_TILE128, &tmV, full, 0, j * 128, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idS = umma_idesc(128, 128, 0, 0);
      constexpr uint32_t idO = umma_idesc(128, AT_D, 0, 1);
      mbar_wait(smem_u32(&bar_q), 0);
      for (int j = 0; j < nkb; ++j) {
        const int s = j % FWD_STAGES;
        const uint32_t ph = (j / FWD_STAGES) & 1;
        mbar_wait(smem_u32(&bar_full[s]), ph);
        tc_fence_after();
        const uint32_t sK = sKV + s * 2 * AT_TILE128, sV = sK + AT_TILE128;
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k)
          umma_bf16(tS, umma_desc(sQ + k * 32, 16, 1024), umma_desc(sK + k * 32, 16, 1024), idS, k != 0);
        umma_commit(smem_u32(&bar_s));
        mbar_wait(smem_u32(&bar_p), j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 128 / 16; ++k)
          umma_bf16_ts(tO, tP + k * 8, umma_desc(sV + k * 2048, 16384, 1024), idO, (j | k) != 0);
        umma_commit(smem_u32(&bar_empty[s]));
      }
      umma_commit(smem_u32(&bar_o));
    }
  } else {
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = uint32_t(qd * 32) << 16;
    float m_used = 0.f, l = 0.f;
    for (int j = 0; j < nkb; ++j) {
      mbar_wait(smem_u32(&bar_s), j & 1);
      tc_fence_after();
      const int valid = min(128, p.n_k - j * 128);
      // pass 1: row max of the raw logits
      float mx = -INFINITY;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t r[32];
        tmem_ld32(tS + lane_off + cc * 32, r);
        if (valid >= 128) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (cc * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      mx *= p.c;
      float factor = 1.f;
      if (j == 0) {
        m_used = mx;
      } else if (mx > m_used + 8.f) {  // lazy rescale: stale max is fine while 2^(s - m) <= 2^8
        factor = fast_exp2(m_used - mx);
        m_used = mx;
      }
      if (j > 0 && __any_sync(AT_FULL, factor != 1.f)) {
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t r[32];
          tmem_ld32(tO + lane_off + cc * 32, r);
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * factor);
          tmem_st32(tO + lane_off + cc * 32, r);
        }
        l *= factor;
      }
      // pass 2: P = 2^(s*c - m), row sum, bf16 P -> TMEM (A operand of the PV MMA)
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t r[32], pk[16];
        tmem_ld32(tS + lane_off + cc * 32, r);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = fast_exp2(fmaf(__uint_as_float(r[2 * i]), p.c, -m_used));
          float p1 = fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), p.c, -m_used));
          if (valid < 128) {
            if (cc * 32 + 2 * i >= valid) p0 = 0.f;
            if (cc * 32 + 2 * i + 1 >= valid) p1 = 0.f;
          }
          l += p0 + p1;
          pk[i] = pack_bf16x2(p0, p1);
        }
        tmem_st16(tP + lane_off + cc * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_p));
    }
    mbar_wait(smem_u32(&bar_o), 0);
    tc_fence_after();
    uint32_t r0[32], r1[32];
    tmem_ld32_nowait(tO + lane_off, r0);
    tmem_ld32_nowait(tO + lane_off + 32, r1);
    tmem_ld_wait();
    const int gq = q0 + row;
    const float inv = 1.f / l;
    if (gq < p.n_q) store_row64(p.out0 + (long long)b * p.bs0 + (long long)gq * p.ld0 + h * AT_D, r0, r1, inv);
    if (gq < p.n_pad) p.LSE[((long long)b * p.H + h) * p.n_pad + gq] = gq < p.n_q ? m_used + log2f(l) : INFINITY;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AT_TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// forward, v2: two query tiles per CTA, ping-pong between the tensor pipe and two softmax warp groups
//   The v1 kernel alternates "MMA, then exp2" inside one CTA and leans on a second resident CTA for overlap; its
//   tensor pipe sat idle for most of every key block (ncu: 19 % tensor-pipe active, 80 us per n = 1024 layer call).
//   Here one CTA owns the whole SM: 256 queries (tiles t = 0, 1), TMEM = S0 | S1 | O0 | O1 (384 columns), the bf16
//   probabilities P_t are written IN PLACE over the first 64 columns of S_t, and the single MMA thread interleaves
//       PV_0(j), S_0(j+1), PV_1(j), S_1(j+1)
//   so that while softmax group 0 chews on S_0 the tensor pipe serves tile 1 and vice versa.  The exp2 throughput
//   of the SFUs (16 / clk / SM) is the bound for head dim 64: 32768 exp2 per 128-key block pair = 2048 cycles against
//   1024 cycles of MMA.
//   Ordering argument for the in-place P / O rescale: tcgen05.commit arrives only after ALL previously issued MMAs of
//   the issuing thread completed, and S_t(j+1) is issued after PV_t(j); hence "S_t(j+1) ready" implies PV_t(j) has
//   finished reading P_t(j) and accumulating into O_t, so group t may overwrite S_t / rescale O_t.  PV_t(j+1) is issued
//   only after group t signalled P_t(j+1).
// warps: 0 = TMA, 1 = MMA issuer + TMEM allocator, 2..5 = softmax group 0, 6..9 = softmax group 1.
// ---------------------------------------------------------------------------------------------
constexpr int F2_THREADS = 320;
constexpr int F2_STAGES = 3;
constexpr int F2_SMEM = 2 * AT_TILE128 + F2_STAGES * 2 * AT_TILE128 + 1024;
constexpr int F2_TMEM_COLS = 512;

__global__ void __launch_bounds__(F2_THREADS, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnP p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_q, bar_full[F2_STAGES], bar_empty[F2_STAGES], bar_s[2], bar_p[2], bar_o;
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sKV = smem_base + 2 * AT_TILE128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256, h = blockIdx.y, b = blockIdx.z;
  const int nkb = (p.n_k + 127) / 128;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(smem_u32(&bar_q), 1);
#pragma unroll
    for (int s = 0; s < F2_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      mbar_init(smem_u32(&bar_s[t]), 1);
      mbar_init(smem_u32(&bar_p[t]), 128);
    }
    mbar_init(smem_u32(&bar_o), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), F2_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(smem_u32(&bar_q), 2 * AT_TILE128);
      tma_load_4d(sQ, &tmQ, smem_u32(&bar_q), 0, q0, h, b);
      tma_load_4d(sQ + AT_TILE128, &tmQ, smem_u32(&bar_q), 0, q0 + 128, h, b);
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nkb; ++j) {
        mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
        const uint32_t full = smem_u32(&bar_full[s]);
        mbar_expect_tx(full, 2 * AT_TILE128);
        tma_load_4d(sKV + s * 2 * AT_TILE128, &tmK, full, 0, j * 128, h, b);
        tma_load_4d(sKV + s * 2 * AT_TILE128 + AT_TILE128, &tmV, full, 0, j * 128, h, b);
        if (++s == F2_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idS = umma_idesc(128, 128, 0, 0);
      constexpr uint32_t idO = umma_idesc(128, AT_D, 0, 1);
      const uint32_t tS0 = tmem_base, tS1 = tmem_base + 128, tO0 = tmem_base + 256, tO1 = tmem_base + 320;
      auto issue_S = [&](uint32_t tS, uint32_t sQt, uint32_t sK) {
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k)
          umma_bf16(tS, umma_desc(sQt + k * 32, 16, 1024), umma_desc(sK + k * 32, 16, 1024), idS, k != 0);
      };
      auto issue_PV = [&](uint32_t tO, uint32_t tP, uint32_t sV, bool first) {
#pragma unroll
        for (int k = 0; k < 128 / 16; ++k)
          umma_bf16_ts(tO, tP + k * 8, umma_desc(sV + k * 2048, 16384, 1024), idO, (!first) || k != 0);
      };
      mbar_wait(smem_u32(&bar_q), 0);
      mbar_wait(smem_u32(&bar_full[0]), 0);
      tc_fence_after();
      issue_S(tS0, sQ, sKV);
      umma_commit(smem_u32(&bar_s[0]));
      issue_S(tS1, sQ + AT_TILE128, sKV);
      umma_commit(smem_u32(&bar_s[1]));
      int s = 0;           // stage of block j
      uint32_t ph = 0;     // its phase
      for (int j = 0; j < nkb; ++j) {
        int sn = s + 1;    // stage / phase of block j + 1
        uint32_t phn = ph;
        if (sn == F2_STAGES) { sn = 0; phn ^= 1u; }
        const uint32_t sV = sKV + s * 2 * AT_TILE128 + AT_TILE128;
        const uint32_t sKn = sKV + sn * 2 * AT_TILE128;
        const bool more = j + 1 < nkb;
        // ---- tile 0
        mbar_wait(smem_u32(&bar_p[0]), j & 1);
        tc_fence_after();
        issue_PV(tO0, tS0, sV, j == 0);
        if (more) {
          mbar_wait(smem_u32(&bar_full[sn]), phn);
          tc_fence_after();
          issue_S(tS0, sQ, sKn);
          umma_commit(smem_u32(&bar_s[0]));
        }
        // ---- tile 1
        mbar_wait(smem_u32(&bar_p[1]), j & 1);
        tc_fence_after();
        issue_PV(tO1, tS1, sV, j == 0);
        umma_commit(smem_u32(&bar_empty[s]));  // K_j (both S MMAs) and V_j (both PV MMAs) consumed
        if (more) {
          issue_S(tS1, sQ + AT_TILE128, sKn);
          umma_commit(smem_u32(&bar_s[1]));
        }
        s = sn;
        ph = phn;
      }
      umma_commit(smem_u32(&bar_o));
    }
  } else {
    const int t = (warp - 2) >> 2;             // softmax group == query tile
    const int qd = warp & 3;                   // TMEM lane quarter this warp may touch
    const int row = qd * 32 + lane;
    const uint32_t lane_off = uint32_t(qd * 32) << 16;
    const uint32_t tS = tmem_base + t * 128 + lane_off;   // P_t aliases the first 64 columns
    const uint32_t tO = tmem_base + 256 + t * 64 + lane_off;
    const uint32_t bs = smem_u32(&bar_s[t]), bp = smem_u32(&bar_p[t]);
    float m_used = 0.f, l = 0.f;
    for (int j = 0; j < nkb; ++j) {
      mbar_wait(bs, j & 1);
      tc_fence_after();
      const int valid = min(128, p.n_k - j * 128);
      // The whole 128-key row lives in registers: ONE TMEM round trip per block (the two-pass version paid eight), and
      // the max / sum reductions run as four independent chains (two warps per scheduler cannot hide a 128-long one).
      uint32_t r[128];
      tmem_ld32_nowait(tS, r);
      tmem_ld32_nowait(tS + 32, r + 32);
      tmem_ld32_nowait(tS + 64, r + 64);
      tmem_ld32_nowait(tS + 96, r + 96);
      tmem_ld_wait();
      if (valid < 128) {  // ragged last key block: -inf logits give P = 0 without per-element predicates later
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= valid) r[i] = 0xff800000u;
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 128; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(r[i]));
        mx1 = fmaxf(mx1, __uint_as_float(r[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(r[i + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(r[i + 3]));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.c;
      float factor = 1.f;
      if (j == 0) {
        m_used = mx;
      } else if (mx > m_used + 8.f) {  // lazy rescale: a stale max is fine while 2^(s - m) <= 2^8
        factor = fast_exp2(m_used - mx);
        m_used = mx;
      }
      if (j > 0 && __any_sync(AT_FULL, factor != 1.f)) {
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t o[32];
          tmem_ld32(tO + cc * 32, o);
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
          tmem_st32(tO + cc * 32, o);
        }
        l *= factor;
      }
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(r[cc * 32 + 2 * i]), p.c, -m_used));
          const float p1 = fast_exp2(fmaf(__uint_as_float(r[cc * 32 + 2 * i + 1]), p.c, -m_used));
          const float p2 = fast_exp2(fmaf(__uint_as_float(r[cc * 32 + 2 * i + 2]), p.c, -m_used));
          const float p3 = fast_exp2(fmaf(__uint_as_float(r[cc * 32 + 2 * i + 3]), p.c, -m_used));
          l0 += p0; l1 += p1; l2 += p2; l3 += p3;
          pk[i] = pack_bf16x2(p0, p1);
          pk[i + 1] = pack_bf16x2(p2, p3);
        }
        tmem_st16(tS + cc * 16, pk);   // in place over S: this thread's row was read out completely above
      }
      l += (l0 + l1) + (l2 + l3);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bp);
    }
    mbar_wait(smem_u32(&bar_o), 0);
    tc_fence_after();
    uint32_t r0[32], r1[32];
    tmem_ld32_nowait(tO, r0);
    tmem_ld32_nowait(tO + 32, r1);
    tmem_ld_wait();
    const int gq = q0 + t * 128 + row;
    const float inv = 1.f / l;
    if (gq < p.n_q) store_row64(p.out0 + (long long)b * p.bs0 + (long long)gq * p.ld0 + h * AT_D, r0, r1, inv);
    if (gq < p.n_pad) p.LSE[((long long)b * p.H + h) * p.n_pad + gq] = gq < p.n_q ? m_used + log2f(l) : INFINITY;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, F2_TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// backward, part 0: D[i] = sum_d dO[i,d] * O[i,d]   (8 lanes per (row, head))
// ---------------------------------------------------------------------------------------------
__global__ void attn_bwd_prep_kernel(const bf16* __restrict__ O, const bf16* __restrict__ dO, float* __restrict__ D, int B,
                                     int H, int n_q, int n_pad, long long ldo, long long o_bs, long long lddo,
                                     long long do_bs) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long item = gid >> 3;
  const int sub = (int)(gid & 7);
  const long long total = (long long)B * n_pad * H;
  const bool live = item < total;
  const int h = (int)(item % H);
  const long long t = item / H;
  const int i = (int)(t % n_pad);
  const int b = (int)(t / n_pad);
  float acc = 0.f;
  if (live && i < n_q) {
    float a[8], g[8];
    unpack8(ld8(O + (long long)b * o_bs + (long long)i * ldo + h * AT_D + sub * 8), a);
    unpack8(ld8(dO + (long long)b * do_bs + (long long)i * lddo + h * AT_D + sub * 8), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += a[j] * g[j];
  }
  acc += __shfl_xor_sync(AT_FULL, acc, 1);
  acc += __shfl_xor_sync(AT_FULL, acc, 2);
  acc += __shfl_xor_sync(AT_FULL, acc, 4);
  if (live && sub == 0) D[((long long)b * H + h) * n_pad + i] = acc;
}

// ---------------------------------------------------------------------------------------------
// backward, dQ: CTA = 128 queries, loops over key blocks of 64
//   TMEM: S [0,64)  dP [64,128)  dS(bf16) [128,160)  dQ [160,224)
// ---------------------------------------------------------------------------------------------
constexpr int BWD_STAGES = 3;
constexpr int BWD_SMEM = 2 * AT_TILE128 + BWD_STAGES * 2 * AT_TILE64 + 1024;

__global__ void __launch_bounds__(AT_THREADS, 2)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const AttnP p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_q, bar_full[BWD_STAGES], bar_empty[BWD_STAGES], bar_s, bar_p, bar_o;
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base, sdO = smem_base + AT_TILE128, sKV = smem_base + 2 * AT_TILE128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nkb = (p.n_k + 63) / 64;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(smem_u32(&bar_q), 1);
#pragma unroll
    for (int s = 0; s < BWD_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_s), 1);
    mbar_init(smem_u32(&bar_p), 128);
    mbar_init(smem_u32(&bar_o), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), AT_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t tS = tmem_base, tdP = tmem_base + 64, tdS = tmem_base + 128, tdQ = tmem_base + 160;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(smem_u32(&bar_q), 2 * AT_TILE128);
      tma_load_4d(sQ, &tmQ, smem_u32(&bar_q), 0, q0, h, b);
      tma_load_4d(sdO, &tmdO, smem_u32(&bar_q), 0, q0, h, b);
      for (int j = 0; j < nkb; ++j) {
        const int s = j % BWD_STAGES;
        const uint32_t ph = (j / BWD_STAGES) & 1;
        mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
        const uint32_t full = smem_u32(&bar_full[s]);
        mbar_expect_tx(full, 2 * AT_TILE64);
        tma_load_4d(sKV + s * 2 * AT_TILE64, &tmK, full, 0, j * 64, h, b);
        tma_load_4d(sKV + s * 2 * AT_TILE64 + AT_TILE64, &tmV, full, 0, j * 64, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idS = umma_idesc(128, 64, 0, 0);
      constexpr uint32_t idQ = umma_idesc(128, AT_D, 0, 1);
      mbar_wait(smem_u32(&bar_q), 0);
      for (int j = 0; j < nkb; ++j) {
        const int s = j % BWD_STAGES;
        const uint32_t ph = (j / BWD_STAGES) & 1;
        mbar_wait(smem_u32(&bar_full[s]), ph);
        tc_fence_after();
        const uint32_t sK = sKV + s * 2 * AT_TILE64, sV = sK + AT_TILE64;
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k)
          umma_bf16(tS, umma_desc(sQ + k * 32, 16, 1024), umma_desc(sK + k * 32, 16, 1024), idS, k != 0);
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k)
          umma_bf16(tdP, umma_desc(sdO + k * 32, 16, 1024), umma_desc(sV + k * 32, 16, 1024), idS, k != 0);
        umma_commit(smem_u32(&bar_s));
        mbar_wait(smem_u32(&bar_p), j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 64 / 16; ++k)
          umma_bf16_ts(tdQ, tdS + k * 8, umma_desc(sK + k * 2048, 8192, 1024), idQ, (j | k) != 0);
        umma_commit(smem_u32(&bar_empty[s]));
      }
      umma_commit(smem_u32(&bar_o));
    }
  } else {
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = uint32_t(qd * 32) << 16;
    const int gq = q0 + row;
    const long long sidx = ((long long)b * p.H + h) * p.n_pad + gq;
    const float L2 = p.LSE[sidx];   // +inf on pad rows -> P = 0
    const float Dr = p.D[sidx];
    for (int j = 0; j < nkb; ++j) {
      mbar_wait(smem_u32(&bar_s), j & 1);
      tc_fence_after();
      const int valid = min(64, p.n_k - j * 64);
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t rs[32], rd[32], pk[16];
        tmem_ld32_nowait(tS + lane_off + cc * 32, rs);
        tmem_ld32_nowait(tdP + lane_off + cc * 32, rd);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = fast_exp2(fmaf(__uint_as_float(rs[2 * i]), p.c, -L2));
          float p1 = fast_exp2(fmaf(__uint_as_float(rs[2 * i + 1]), p.c, -L2));
          if (valid < 64) {
            if (cc * 32 + 2 * i >= valid) p0 = 0.f;
            if (cc * 32 + 2 * i + 1 >= valid) p1 = 0.f;
          }
          pk[i] = pack_bf16x2(p0 * (__uint_as_float(rd[2 * i]) - Dr), p1 * (__uint_as_float(rd[2 * i + 1]) - Dr));
        }
        tmem_st16(tdS + lane_off + cc * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_p));
    }
    mbar_wait(smem_u32(&bar_o), 0);
    tc_fence_after();
    uint32_t r0[32], r1[32];
    tmem_ld32_nowait(tdQ + lane_off, r0);
    tmem_ld32_nowait(tdQ + lane_off + 32, r1);
    tmem_ld_wait();
    if (gq < p.n_q) store_row64(p.out0 + (long long)b * p.bs0 + (long long)gq * p.ld0 + h * AT_D, r0, r1, p.scale);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AT_TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// backward, dK / dV: CTA = 128 keys, loops over query blocks of 64
//   TMEM: S^T -> P^T(bf16, in place) [0,64)   dP^T -> dS^T(bf16, in place) [64,128)   dV [128,192)   dK [192,256)
//   The next S^T / dP^T MMAs overwrite columns the previous dV / dK MMAs read as their A operand; tcgen05.mma
//   instructions of one thread execute in issue order, so the pipe itself orders that WAR.  kDrain = true inserts an
//   explicit completion wait instead (debug / A-B check).
// ---------------------------------------------------------------------------------------------
template <bool kDrain>
__global__ void __launch_bounds__(AT_THREADS, 2)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                    const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO, const AttnP p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_kv, bar_full[BWD_STAGES], bar_empty[BWD_STAGES], bar_s, bar_p, bar_o, bar_x;
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = smem_base, sV = smem_base + AT_TILE128, sQdO = smem_base + 2 * AT_TILE128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nqb = (p.n_q + 63) / 64;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmdO);
    mbar_init(smem_u32(&bar_kv), 1);
#pragma unroll
    for (int s = 0; s < BWD_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_s), 1);
    mbar_init(smem_u32(&bar_p), 128);
    mbar_init(smem_u32(&bar_o), 1);
    mbar_init(smem_u32(&bar_x), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), AT_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t tS = tmem_base, tdP = tmem_base + 64, tdV = tmem_base + 128, tdK = tmem_base + 192;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(smem_u32(&bar_kv), 2 * AT_TILE128);
      tma_load_4d(sK, &tmK, smem_u32(&bar_kv), 0, k0, h, b);
      tma_load_4d(sV, &tmV, smem_u32(&bar_kv), 0, k0, h, b);
      for (int i = 0; i < nqb; ++i) {
        const int s = i % BWD_STAGES;
        const uint32_t ph = (i / BWD_STAGES) & 1;
        mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
        const uint32_t full = smem_u32(&bar_full[s]);
        mbar_expect_tx(full, 2 * AT_TILE64);
        tma_load_4d(sQdO + s * 2 * AT_TILE64, &tmQ, full, 0, i * 64, h, b);
        tma_load_4d(sQdO + s * 2 * AT_TILE64 + AT_TILE64, &tmdO, full, 0, i * 64, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idS = umma_idesc(128, 64, 0, 0);
      constexpr uint32_t idG = umma_idesc(128, AT_D, 0, 1);
      mbar_wait(smem_u32(&bar_kv), 0);
      for (int i = 0; i < nqb; ++i) {
        const int s = i % BWD_STAGES;
        const uint32_t ph = (i / BWD_STAGES) & 1;
        mbar_wait(smem_u32(&bar_full[s]), ph);
        if (kDrain && i > 0) mbar_wait(smem_u32(&bar_x), (i - 1) & 1);
        tc_fence_after();
        const uint32_t sQ = sQdO + s * 2 * AT_TILE64, sdO = sQ + AT_TILE64;
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k)
          umma_bf16(tS, umma_desc(sK + k * 32, 16, 1024), umma_desc(sQ + k * 32, 16, 1024), idS, k != 0);
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k)
          umma_bf16(tdP, umma_desc(sV + k * 32, 16, 1024), umma_desc(sdO + k * 32, 16, 1024), idS, k != 0);
        umma_commit(smem_u32(&bar_s));
        mbar_wait(smem_u32(&bar_p), i & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 64 / 16; ++k)
          umma_bf16_ts(tdV, tS + k * 8, umma_desc(sdO + k * 2048, 8192, 1024), idG, (i | k) != 0);
#pragma unroll
        for (int k = 0; k < 64 / 16; ++k)
          umma_bf16_ts(tdK, tdP + k * 8, umma_desc(sQ + k * 2048, 8192, 1024), idG, (i | k) != 0);
        umma_commit(smem_u32(&bar_empty[s]));
        if (kDrain) umma_commit(smem_u32(&bar_x));
      }
      umma_commit(smem_u32(&bar_o));
    }
  } else {
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = uint32_t(qd * 32) << 16;
    const float* Lp = p.LSE + ((long long)b * p.H + h) * p.n_pad;
    const float* Dp = p.D + ((long long)b * p.H + h) * p.n_pad;
    for (int i = 0; i < nqb; ++i) {
      mbar_wait(smem_u32(&bar_s), i & 1);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t rs[32], rd[32], pp[16], pd[16];
        tmem_ld32_nowait(tS + lane_off + cc * 32, rs);
        tmem_ld32_nowait(tdP + lane_off + cc * 32, rd);
        const float4* L4 = reinterpret_cast<const float4*>(Lp + i * 64 + cc * 32);
        const float4* D4 = reinterpret_cast<const float4*>(Dp + i * 64 + cc * 32);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 lv = __ldg(L4 + g), dv = __ldg(D4 + g);
          const float p0 = fast_exp2(fmaf(__uint_as_float(rs[4 * g + 0]), p.c, -lv.x));
          const float p1 = fast_exp2(fmaf(__uint_as_float(rs[4 * g + 1]), p.c, -lv.y));
          const float p2 = fast_exp2(fmaf(__uint_as_float(rs[4 * g + 2]), p.c, -lv.z));
          const float p3 = fast_exp2(fmaf(__uint_as_float(rs[4 * g + 3]), p.c, -lv.w));
          pp[2 * g] = pack_bf16x2(p0, p1);
          pp[2 * g + 1] = pack_bf16x2(p2, p3);
          pd[2 * g] = pack_bf16x2(p0 * (__uint_as_float(rd[4 * g + 0]) - dv.x), p1 * (__uint_as_float(rd[4 * g + 1]) - dv.y));
          pd[2 * g + 1] = pack_bf16x2(p2 * (__uint_as_float(rd[4 * g + 2]) - dv.z), p3 * (__uint_as_float(rd[4 * g + 3]) - dv.w));
        }
        tmem_st16(tS + lane_off + cc * 16, pp);
        tmem_st16(tdP + lane_off + cc * 16, pd);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_p));
    }
    mbar_wait(smem_u32(&bar_o), 0);
    tc_fence_after();
    const int gk = k0 + row;
    uint32_t r0[32], r1[32];
    tmem_ld32_nowait(tdV + lane_off, r0);
    tmem_ld32_nowait(tdV + lane_off + 32, r1);
    tmem_ld_wait();
    if (gk < p.n_k) store_row64(p.out1 + (long long)b * p.bs1 + (long long)gk * p.ld1 + h * AT_D, r0, r1, 1.f);
    tmem_ld32_nowait(tdK + lane_off, r0);
    tmem_ld32_nowait(tdK + lane_off + 32, r1);
    tmem_ld_wait();
    if (gk < p.n_k) store_row64(p.out0 + (long long)b * p.bs0 + (long long)gk * p.ld0 + h * AT_D, r0, r1, p.scale);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AT_TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// backward v2: same ping-pong structure as attn_fwd2_kernel (one CTA per SM, two tiles, two softmax warp groups, the MMA
// thread interleaving the tiles), bf16 dS / P^T written in place over the fp32 S / dP columns.
//   dq2 : CTA = 256 queries (2 tiles) x key blocks of 64.    TMEM per tile: S | dP | dQ            (3 x 64 columns)
//   dkv2: CTA = 256 keys    (2 tiles) x query blocks of 64.  TMEM per tile: S^T | dP^T | dV | dK   (4 x 64 columns)
// Per 64-wide block and tile the tensor pipe needs 384 (dq) / 512 (dkv) cycles and the SFUs 512 cycles (8192 exp2).
// ---------------------------------------------------------------------------------------------
constexpr int B2_THREADS = 320;
constexpr int B2_STAGES = 4;
constexpr int B2_SMEM = 4 * AT_TILE128 + B2_STAGES * 2 * AT_TILE64 + 1024;
constexpr int B2_TMEM_COLS = 512;

__global__ void __launch_bounds__(B2_THREADS, 1)
attn_bwd_dq2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                    const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const AttnP p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_q, bar_full[B2_STAGES], bar_empty[B2_STAGES], bar_s[2], bar_p[2], bar_o;
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base, sdO = smem_base + 2 * AT_TILE128, sKV = smem_base + 4 * AT_TILE128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256, h = blockIdx.y, b = blockIdx.z;
  const int nkb = (p.n_k + 63) / 64;
  const bool two = q0 + 128 < p.n_q;  // second query tile holds real rows

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(smem_u32(&bar_q), 1);
#pragma unroll
    for (int s = 0; s < B2_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      mbar_init(smem_u32(&bar_s[t]), 1);
      mbar_init(smem_u32(&bar_p[t]), 128);
    }
    mbar_init(smem_u32(&bar_o), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), B2_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(smem_u32(&bar_q), 4 * AT_TILE128);
      tma_load_4d(sQ, &tmQ, smem_u32(&bar_q), 0, q0, h, b);
      tma_load_4d(sQ + AT_TILE128, &tmQ, smem_u32(&bar_q), 0, q0 + 128, h, b);
      tma_load_4d(sdO, &tmdO, smem_u32(&bar_q), 0, q0, h, b);
      tma_load_4d(sdO + AT_TILE128, &tmdO, smem_u32(&bar_q), 0, q0 + 128, h, b);
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nkb; ++j) {
        mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
        const uint32_t full = smem_u32(&bar_full[s]);
        mbar_expect_tx(full, 2 * AT_TILE64);
        tma_load_4d(sKV + s * 2 * AT_TILE64, &tmK, full, 0, j * 64, h, b);
        tma_load_4d(sKV + s * 2 * AT_TILE64 + AT_TILE64, &tmV, full, 0, j * 64, h, b);
        if (++s == B2_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idS = umma_idesc(128, 64, 0, 0);
      constexpr uint32_t idQ = umma_idesc(128, AT_D, 0, 1);
      auto issue_SdP = [&](int t, uint32_t sK, uint32_t sV) {
        const uint32_t tS = tmem_base + t * 192, tdP = tS + 64;
        const uint32_t sQt = sQ + t * AT_TILE128, sdOt = sdO + t * AT_TILE128;
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k)
          umma_bf16(tS, umma_desc(sQt + k * 32, 16, 1024), umma_desc(sK + k * 32, 16, 1024), idS, k != 0);
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k)
          umma_bf16(tdP, umma_desc(sdOt + k * 32, 16, 1024), umma_desc(sV + k * 32, 16, 1024), idS, k != 0);
      };
      auto issue_dQ = [&](int t, uint32_t sK, bool first) {
        const uint32_t tdS = tmem_base + t * 192, tdQ = tdS + 128;
#pragma unroll
        for (int k = 0; k < 64 / 16; ++k)
          umma_bf16_ts(tdQ, tdS + k * 8, umma_desc(sK + k * 2048, 8192, 1024), idQ, (!first) || k != 0);
      };
      mbar_wait(smem_u32(&bar_q), 0);
      mbar_wait(smem_u32(&bar_full[0]), 0);
      tc_fence_after();
      issue_SdP(0, sKV, sKV + AT_TILE64);
      umma_commit(smem_u32(&bar_s[0]));
      if (two) {
        issue_SdP(1, sKV, sKV + AT_TILE64);
        umma_commit(smem_u32(&bar_s[1]));
      }
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nkb; ++j) {
        int sn = s + 1;
        uint32_t phn = ph;
        if (sn == B2_STAGES) { sn = 0; phn ^= 1u; }
        const uint32_t sK = sKV + s * 2 * AT_TILE64;
        const uint32_t sKn = sKV + sn * 2 * AT_TILE64, sVn = sKn + AT_TILE64;
        const bool more = j + 1 < nkb;
        mbar_wait(smem_u32(&bar_p[0]), j & 1);
        tc_fence_after();
        issue_dQ(0, sK, j == 0);
        if (more) {
          mbar_wait(smem_u32(&bar_full[sn]), phn);
          tc_fence_after();
          issue_SdP(0, sKn, sVn);
          umma_commit(smem_u32(&bar_s[0]));
        }
        if (two) {
          mbar_wait(smem_u32(&bar_p[1]), j & 1);
          tc_fence_after();
          issue_dQ(1, sK, j == 0);
        }
        umma_commit(smem_u32(&bar_empty[s]));
        if (more && two) {
          issue_SdP(1, sKn, sVn);
          umma_commit(smem_u32(&bar_s[1]));
        }
        s = sn;
        ph = phn;
      }
      umma_commit(smem_u32(&bar_o));
    }
  } else {
    const int t = (warp - 2) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = uint32_t(qd * 32) << 16;
    const int gq = q0 + t * 128 + row;
    if (t == 0 || two) {
      const uint32_t tS = tmem_base + t * 192 + lane_off, tdP = tS + 64, tdQ = tS + 128;
      const uint32_t bs = smem_u32(&bar_s[t]), bp = smem_u32(&bar_p[t]);
      const long long sidx = ((long long)b * p.H + h) * p.n_pad + gq;
      const float L2 = gq < p.n_pad ? p.LSE[sidx] : INFINITY;  // +inf on pad rows -> P = 0
      const float Dr = gq < p.n_pad ? p.D[sidx] : 0.f;
      for (int j = 0; j < nkb; ++j) {
        mbar_wait(bs, j & 1);
        tc_fence_after();
        const int valid = min(64, p.n_k - j * 64);
        // whole 64-key block in registers: one TMEM round trip, then 64 independent exp2 / dS evaluations
        uint32_t rs[64], rd[64];
        tmem_ld32_nowait(tS, rs);
        tmem_ld32_nowait(tS + 32, rs + 32);
        tmem_ld32_nowait(tdP, rd);
        tmem_ld32_nowait(tdP + 32, rd + 32);
        tmem_ld_wait();
        if (valid < 64) {
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i >= valid) rs[i] = 0xff800000u;  // -inf logit -> P = 0 -> dS = 0
        }
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(rs[cc * 32 + 2 * i]), p.c, -L2));
            const float p1 = fast_exp2(fmaf(__uint_as_float(rs[cc * 32 + 2 * i + 1]), p.c, -L2));
            pk[i] = pack_bf16x2(p0 * (__uint_as_float(rd[cc * 32 + 2 * i]) - Dr),
                                p1 * (__uint_as_float(rd[cc * 32 + 2 * i + 1]) - Dr));
          }
          tmem_st16(tS + cc * 16, pk);  // dS (bf16) in place over the consumed S columns
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bp);
      }
      mbar_wait(smem_u32(&bar_o), 0);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld32_nowait(tdQ, r0);
      tmem_ld32_nowait(tdQ + 32, r1);
      tmem_ld_wait();
      if (gq < p.n_q) store_row64(p.out0 + (long long)b * p.bs0 + (long long)gq * p.ld0 + h * AT_D, r0, r1, p.scale);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, B2_TMEM_COLS);
  }
}

__global__ void __launch_bounds__(B2_THREADS, 1)
attn_bwd_dkv2_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                     const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO, const AttnP p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_kv, bar_full[B2_STAGES], bar_empty[B2_STAGES], bar_s[2], bar_p[2], bar_o;
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = smem_base, sV = smem_base + 2 * AT_TILE128, sQdO = smem_base + 4 * AT_TILE128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 256, h = blockIdx.y, b = blockIdx.z;
  const int nqb = (p.n_q + 63) / 64;
  const bool two = k0 + 128 < p.n_k;  // cross-attention (77 keys): only tile 0 holds real keys

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmdO);
    mbar_init(smem_u32(&bar_kv), 1);
#pragma unroll
    for (int s = 0; s < B2_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      mbar_init(smem_u32(&bar_s[t]), 1);
      mbar_init(smem_u32(&bar_p[t]), 128);
    }
    mbar_init(smem_u32(&bar_o), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), B2_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(smem_u32(&bar_kv), 4 * AT_TILE128);
      tma_load_4d(sK, &tmK, smem_u32(&bar_kv), 0, k0, h, b);
      tma_load_4d(sK + AT_TILE128, &tmK, smem_u32(&bar_kv), 0, k0 + 128, h, b);
      tma_load_4d(sV, &tmV, smem_u32(&bar_kv), 0, k0, h, b);
      tma_load_4d(sV + AT_TILE128, &tmV, smem_u32(&bar_kv), 0, k0 + 128, h, b);
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nqb; ++i) {
        mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
        const uint32_t full = smem_u32(&bar_full[s]);
        mbar_expect_tx(full, 2 * AT_TILE64);
        tma_load_4d(sQdO + s * 2 * AT_TILE64, &tmQ, full, 0, i * 64, h, b);
        tma_load_4d(sQdO + s * 2 * AT_TILE64 + AT_TILE64, &tmdO, full, 0, i * 64, h, b);
        if (++s == B2_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idS = umma_idesc(128, 64, 0, 0);
      constexpr uint32_t idG = umma_idesc(128, AT_D, 0, 1);
      auto issue_SdP = [&](int t, uint32_t sQb, uint32_t sdOb) {
        const uint32_t tS = tmem_base + t * 256, tdP = tS + 64;
        const uint32_t sKt = sK + t * AT_TILE128, sVt = sV + t * AT_TILE128;
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k)
          umma_bf16(tS, umma_desc(sKt + k * 32, 16, 1024), umma_desc(sQb + k * 32, 16, 1024), idS, k != 0);
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k)
          umma_bf16(tdP, umma_desc(sVt + k * 32, 16, 1024), umma_desc(sdOb + k * 32, 16, 1024), idS, k != 0);
      };
      auto issue_dVdK = [&](int t, uint32_t sQb, uint32_t sdOb, bool first) {
        const uint32_t tS = tmem_base + t * 256, tdP = tS + 64, tdV = tS + 128, tdK = tS + 192;
#pragma unroll
        for (int k = 0; k < 64 / 16; ++k)
          umma_bf16_ts(tdV, tS + k * 8, umma_desc(sdOb + k * 2048, 8192, 1024), idG, (!first) || k != 0);
#pragma unroll
        for (int k = 0; k < 64 / 16; ++k)
          umma_bf16_ts(tdK, tdP + k * 8, umma_desc(sQb + k * 2048, 8192, 1024), idG, (!first) || k != 0);
      };
      mbar_wait(smem_u32(&bar_kv), 0);
      mbar_wait(smem_u32(&bar_full[0]), 0);
      tc_fence_after();
      issue_SdP(0, sQdO, sQdO + AT_TILE64);
      umma_commit(smem_u32(&bar_s[0]));
      if (two) {
        issue_SdP(1, sQdO, sQdO + AT_TILE64);
        umma_commit(smem_u32(&bar_s[1]));
      }
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nqb; ++i) {
        int sn = s + 1;
        uint32_t phn = ph;
        if (sn == B2_STAGES) { sn = 0; phn ^= 1u; }
        const uint32_t sQb = sQdO + s * 2 * AT_TILE64, sdOb = sQb + AT_TILE64;
        const uint32_t sQn = sQdO + sn * 2 * AT_TILE64, sdOn = sQn + AT_TILE64;
        const bool more = i + 1 < nqb;
        mbar_wait(smem_u32(&bar_p[0]), i & 1);
        tc_fence_after();
        issue_dVdK(0, sQb, sdOb, i == 0);
        if (more) {
          mbar_wait(smem_u32(&bar_full[sn]), phn);
          tc_fence_after();
          issue_SdP(0, sQn, sdOn);
          umma_commit(smem_u32(&bar_s[0]));
        }
        if (two) {
          mbar_wait(smem_u32(&bar_p[1]), i & 1);
          tc_fence_after();
          issue_dVdK(1, sQb, sdOb, i == 0);
        }
        umma_commit(smem_u32(&bar_empty[s]));
        if (more && two) {
          issue_SdP(1, sQn, sdOn);
          umma_commit(smem_u32(&bar_s[1]));
        }
        s = sn;
        ph = phn;
      }
      umma_commit(smem_u32(&bar_o));
    }
  } else {
    const int t = (warp - 2) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = uint32_t(qd * 32) << 16;
    if (t == 0 || two) {
      const uint32_t tS = tmem_base + t * 256 + lane_off, tdP = tS + 64, tdV = tS + 128, tdK = tS + 192;
      const uint32_t bs = smem_u32(&bar_s[t]), bp = smem_u32(&bar_p[t]);
      const float* Lp = p.LSE + ((long long)b * p.H + h) * p.n_pad;
      const float* Dp = p.D + ((long long)b * p.H + h) * p.n_pad;
      for (int i = 0; i < nqb; ++i) {
        mbar_wait(bs, i & 1);
        tc_fence_after();
        // whole 64-query block in registers (one TMEM round trip); LSE / D of the 64 queries are warp-uniform loads
        uint32_t rs[64], rd[64];
        tmem_ld32_nowait(tS, rs);
        tmem_ld32_nowait(tS + 32, rs + 32);
        tmem_ld32_nowait(tdP, rd);
        tmem_ld32_nowait(tdP + 32, rd + 32);
        // n_pad is a multiple of 128 and the last 64-query block starts below n_q <= n_pad: always in bounds
        const float4* L4 = reinterpret_cast<const float4*>(Lp + i * 64);
        const float4* D4 = reinterpret_cast<const float4*>(Dp + i * 64);
        tmem_ld_wait();
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t pp[16], pd[16];
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 lv = __ldg(L4 + cc * 8 + g), dv = __ldg(D4 + cc * 8 + g);
            const int o = cc * 32 + 4 * g;
            const float p0 = fast_exp2(fmaf(__uint_as_float(rs[o + 0]), p.c, -lv.x));
            const float p1 = fast_exp2(fmaf(__uint_as_float(rs[o + 1]), p.c, -lv.y));
            const float p2 = fast_exp2(fmaf(__uint_as_float(rs[o + 2]), p.c, -lv.z));
            const float p3 = fast_exp2(fmaf(__uint_as_float(rs[o + 3]), p.c, -lv.w));
            pp[2 * g] = pack_bf16x2(p0, p1);
            pp[2 * g + 1] = pack_bf16x2(p2, p3);
            pd[2 * g] = pack_bf16x2(p0 * (__uint_as_float(rd[o + 0]) - dv.x), p1 * (__uint_as_float(rd[o + 1]) - dv.y));
            pd[2 * g + 1] = pack_bf16x2(p2 * (__uint_as_float(rd[o + 2]) - dv.z), p3 * (__uint_as_float(rd[o + 3]) - dv.w));
          }
          tmem_st16(tS + cc * 16, pp);
          tmem_st16(tdP + cc * 16, pd);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bp);
      }
      mbar_wait(smem_u32(&bar_o), 0);
      tc_fence_after();
      const int gk = k0 + t * 128 + row;
      uint32_t r0[32], r1[32];
      tmem_ld32_nowait(tdV, r0);
      tmem_ld32_nowait(tdV + 32, r1);
      tmem_ld_wait();
      if (gk < p.n_k) store_row64(p.out1 + (long long)b * p.bs1 + (long long)gk * p.ld1 + h * AT_D, r0, r1, 1.f);
      tmem_ld32_nowait(tdK, r0);
      tmem_ld32_nowait(tdK + 32, r1);
      tmem_ld_wait();
      if (gk < p.n_k) store_row64(p.out0 + (long long)b * p.bs0 + (long long)gk * p.ld0 + h * AT_D, r0, r1, p.scale);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, B2_TMEM_COLS);
  }
}

template <typename K>
static int set_smem(K kernel, int bytes, const char* what) {
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (err != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(err));
    return B2_ERR_CUDA;
  }
  return B2_OK;
}

static int check_common(const b2_attn_args* a, const char* what) {
  B2_REQUIRE(a && a->Q && a->K && a->V && a->O && a->LSE, "%s: null pointer", what);
  B2_REQUIRE(a->B > 0 && a->H > 0 && a->n_q > 0 && a->n_k > 0, "%s: bad shape", what);
  B2_REQUIRE(a->H <= 65535 && a->B <= 65535, "%s: H/B too large for the grid", what);
  return B2_OK;
}

}  // namespace b2

using namespace b2;

extern "C" int b2_attn_lse_rows(int n_q) { return (n_q + 127) / 128 * 128; }

extern "C" int b2_attn_fwd(const b2_attn_args* a, void* stream) {
  int rc = check_common(a, "b2_attn_fwd");
  if (rc) return rc;
  CUtensorMap tq, tk, tv;
  if ((rc = make_map_bf16_4d(&tq, a->Q, AT_D, a->n_q, a->H, a->B, a->ldq, AT_D, a->q_bs, 64, 128, "attn Q"))) return rc;
  if ((rc = make_map_bf16_4d(&tk, a->K, AT_D, a->n_k, a->H, a->B, a->ldk, AT_D, a->k_bs, 64, 128, "attn K"))) return rc;
  if ((rc = make_map_bf16_4d(&tv, a->V, AT_D, a->n_k, a->H, a->B, a->ldv, AT_D, a->v_bs, 64, 128, "attn V"))) return rc;
  static bool configured = false;
  if (!configured) {
    if ((rc = set_smem(attn_fwd_kernel, FWD_SMEM, "b2_attn_fwd"))) return rc;
    configured = true;
  }
  AttnP p{};
  p.H = a->H; p.n_q = a->n_q; p.n_k = a->n_k; p.n_pad = b2_attn_lse_rows(a->n_q);
  p.scale = a->scale; p.c = a->scale * 1.4426950408889634f;
  p.LSE = a->LSE; p.D = nullptr;
  p.out0 = (bf16*)a->O; p.ld0 = a->ldo; p.bs0 = a->o_bs;
  static const bool legacy = getenv("B2_ATTN_LEGACY") != nullptr;
  if (!legacy) {
    static bool configured2 = false;
    if (!configured2) {
      if ((rc = set_smem(attn_fwd2_kernel, F2_SMEM, "b2_attn_fwd"))) return rc;
      configured2 = true;
    }
    dim3 grid2((a->n_q + 255) / 256, a->H, a->B);
    attn_fwd2_kernel<<<grid2, F2_THREADS, F2_SMEM, (cudaStream_t)stream>>>(tq, tk, tv, p);
    return check_launch("b2_attn_fwd");
  }
  dim3 grid((a->n_q + 127) / 128, a->H, a->B);
  attn_fwd_kernel<<<grid, AT_THREADS, FWD_SMEM, (cudaStream_t)stream>>>(tq, tk, tv, p);
  return check_launch("b2_attn_fwd");
}

extern "C" int b2_attn_bwd(const b2_attn_args* a, void* stream) {
  int rc = check_common(a, "b2_attn_bwd");
  if (rc) return rc;
  B2_REQUIRE(a->dO && a->D && a->dQ && a->dK && a->dV, "b2_attn_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int n_pad = b2_attn_lse_rows(a->n_q);
  {
    const long long threads = (long long)a->B * n_pad * a->H * 8;
    attn_bwd_prep_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
        (const bf16*)a->O, (const bf16*)a->dO, a->D, a->B, a->H, a->n_q, n_pad, a->ldo, a->o_bs, a->lddo, a->do_bs);
    if ((rc = check_launch("b2_attn_bwd prep"))) return rc;
  }
  static bool configured = false;
  if (!configured) {
    if ((rc = set_smem(attn_bwd_dq_kernel, BWD_SMEM, "b2_attn_bwd"))) return rc;
    if ((rc = set_smem(attn_bwd_dkv_kernel<false>, BWD_SMEM, "b2_attn_bwd"))) return rc;
    if ((rc = set_smem(attn_bwd_dkv_kernel<true>, BWD_SMEM, "b2_attn_bwd"))) return rc;
    configured = true;
  }
  AttnP p{};
  p.H = a->H; p.n_q = a->n_q; p.n_k = a->n_k; p.n_pad = n_pad;
  p.scale = a->scale; p.c = a->scale * 1.4426950408889634f;
  p.LSE = a->LSE; p.D = a->D;
  static const bool legacy = getenv("B2_ATTN_LEGACY") != nullptr;
  if (!legacy && !(a->flags & 1)) {
    static bool configured2 = false;
    if (!configured2) {
      if ((rc = set_smem(attn_bwd_dq2_kernel, B2_SMEM, "b2_attn_bwd"))) return rc;
      if ((rc = set_smem(attn_bwd_dkv2_kernel, B2_SMEM, "b2_attn_bwd"))) return rc;
      configured2 = true;
    }
    CUtensorMap tq128, tdo128, tk64, tv64, tk128, tv128, tq64, tdo64;
    if ((rc = make_map_bf16_4d(&tq128, a->Q, AT_D, a->n_q, a->H, a->B, a->ldq, AT_D, a->q_bs, 64, 128, "attn Q"))) return rc;
    if ((rc = make_map_bf16_4d(&tdo128, a->dO, AT_D, a->n_q, a->H, a->B, a->lddo, AT_D, a->do_bs, 64, 128, "attn dO"))) return rc;
    if ((rc = make_map_bf16_4d(&tk64, a->K, AT_D, a->n_k, a->H, a->B, a->ldk, AT_D, a->k_bs, 64, 64, "attn K64"))) return rc;
    if ((rc = make_map_bf16_4d(&tv64, a->V, AT_D, a->n_k, a->H, a->B, a->ldv, AT_D, a->v_bs, 64, 64, "attn V64"))) return rc;
    if ((rc = make_map_bf16_4d(&tk128, a->K, AT_D, a->n_k, a->H, a->B, a->ldk, AT_D, a->k_bs, 64, 128, "attn K"))) return rc;
    if ((rc = make_map_bf16_4d(&tv128, a->V, AT_D, a->n_k, a->H, a->B, a->ldv, AT_D, a->v_bs, 64, 128, "attn V"))) return rc;
    if ((rc = make_map_bf16_4d(&tq64, a->Q, AT_D, a->n_q, a->H, a->B, a->ldq, AT_D, a->q_bs, 64, 64, "attn Q64"))) return rc;
    if ((rc = make_map_bf16_4d(&tdo64, a->dO, AT_D, a->n_q, a->H, a->B, a->lddo, AT_D, a->do_bs, 64, 64, "attn dO64"))) return rc;
    AttnP pq = p;
    pq.out0 = (bf16*)a->dQ; pq.ld0 = a->lddq; pq.bs0 = a->dq_bs;
    attn_bwd_dq2_kernel<<<dim3((a->n_q + 255) / 256, a->H, a->B), B2_THREADS, B2_SMEM, st>>>(tq128, tdo128, tk64, tv64, pq);
    if ((rc = check_launch("b2_attn_bwd dq2"))) return rc;
    AttnP pk = p;
    pk.out0 = (bf16*)a->dK; pk.ld0 = a->lddk; pk.bs0 = a->dk_bs;
    pk.out1 = (bf16*)a->dV; pk.ld1 = a->lddv; pk.bs1 = a->dv_bs;
    attn_bwd_dkv2_kernel<<<dim3((a->n_k + 255) / 256, a->H, a->B), B2_THREADS, B2_SMEM, st>>>(tk128, tv128, tq64, tdo64, pk);
    return check_launch("b2_attn_bwd dkv2");
  }
  {
    CUtensorMap tq, tdo, tk, tv;
    if ((rc = make_map_bf16_4d(&tq, a->Q, AT_D, a->n_q, a->H, a->B, a->ldq, AT_D, a->q_bs, 64, 128, "attn Q"))) return rc;
    if ((rc = make_map_bf16_4d(&tdo, a->dO, AT_D, a->n_q, a->H, a->B, a->lddo, AT_D, a->do_bs, 64, 128, "attn dO"))) return rc;
    if ((rc = make_map_bf16_4d(&tk, a->K, AT_D, a->n_k, a->H, a->B, a->ldk, AT_D, a->k_bs, 64, 64, "attn K64"))) return rc;
    if ((rc = make_map_bf16_4d(&tv, a->V, AT_D, a->n_k, a->H, a->B, a->ldv, AT_D, a->v_bs, 64, 64, "attn V64"))) return rc;
    AttnP pq = p;
    pq.out0 = (bf16*)a->dQ; pq.ld0 = a->lddq; pq.bs0 = a->dq_bs;
    dim3 grid((a->n_q + 127) / 128, a->H, a->B);
    attn_bwd_dq_kernel<<<grid, AT_THREADS, BWD_SMEM, st>>>(tq, tdo, tk, tv, pq);
    if ((rc = check_launch("b2_attn_bwd dq"))) return rc;
  }
  {
    CUtensorMap tk, tv, tq, tdo;
    if ((rc = make_map_bf16_4d(&tk, a->K, AT_D, a->n_k, a->H, a->B, a->ldk, AT_D, a->k_bs, 64, 128, "attn K"))) return rc;
    if ((rc = make_map_bf16_4d(&tv, a->V, AT_D, a->n_k, a->H, a->B, a->ldv, AT_D, a->v_bs, 64, 128, "attn V"))) return rc;
    if ((rc = make_map_bf16_4d(&tq, a->Q, AT_D, a->n_q, a->H, a->B, a->ldq, AT_D, a->q_bs, 64, 64, "attn Q64"))) return rc;
    if ((rc = make_map_bf16_4d(&tdo, a->dO, AT_D, a->n_q, a->H, a->B, a->lddo, AT_D, a->do_bs, 64, 64, "attn dO64"))) return rc;
    AttnP pk = p;
    pk.out0 = (bf16*)a->dK; pk.ld0 = a->lddk; pk.bs0 = a->dk_bs;
    pk.out1 = (bf16*)a->dV; pk.ld1 = a->lddv; pk.bs1 = a->dv_bs;
    dim3 grid((a->n_k + 127) / 128, a->H, a->B);
    if (a->flags & 1)
      attn_bwd_dkv_kernel<true><<<grid, AT_THREADS, BWD_SMEM, st>>>(tk, tv, tq, tdo, pk);
    else
      attn_bwd_dkv_kernel<false><<<grid, AT_THREADS, BWD_SMEM, st>>>(tk, tv, tq, tdo, pk);
    if ((rc = check_launch("b2_attn_bwd dkv"))) return rc;
  }
  return B2_OK;
}
