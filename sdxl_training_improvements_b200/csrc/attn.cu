// Fused attention for sm_100a: softmax(Q K^T * scale) V, head dim 64, bf16 in/out, fp32 accumulation in TMEM.
// Replaces F.scaled_dot_product_attention inside diffusers' AttnProcessor2_0 (self-attention, n keys; cross-attention,
// 77 keys) and its autograd backward — SURVEY.md §8 a5.5 / a5.6.  No n x n tensor ever touches HBM.
//
// Kernels (all: TMA-staged bf16 tiles in SWIZZLE_128B smem -> tcgen05.mma (SS) into TMEM -> softmax warps read their TMEM
// lane with tcgen05.ld, do the exp2 / rescale math in registers, write the bf16 result back IN PLACE with tcgen05.st ->
// tcgen05.mma (TS: A operand from TMEM) accumulates the output tile in TMEM):
//   attn_pfwd2  : persistent CTA per SM over (sample, head, 128-query tile) x key blocks of 128; ring of three S accumulators,
//                 two softmax groups on alternate blocks, deferred two-group merge            (attn_fwd3: its per-tile ancestor)
//   attn_xfwd   : cross-attention (<= 96 keys): persistent CTA per SM over (sample, head, query-tile) work items
//   attn_bwd_dq3: CTA = 128 queries x key blocks of 64;   S, dP -> dS -> dQ += dS K
//   attn_bwd_dkv3: CTA = 128 keys x query blocks of 64;   S^T, dP^T -> P^T, dS^T -> dV += P^T dO, dK += dS^T Q
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 / 6..9 = the two
// softmax groups (warp w owns TMEM lanes 32*(w%4)..+31).  One CTA per SM (512 TMEM columns).
//
// LSE is kept in the log2 domain: L2[i] = m_i + log2(sum_j 2^(s_ij*c - m_i)), c = scale*log2(e); P_ij = 2^(s_ij*c - L2[i]).
// LSE / D are laid out [B, H, n_pad] with n_pad = ceil(n_q/128)*128; pad rows hold L2 = +inf (P = 0) and D = 0.
#include <math.h>
#include <stdlib.h>

#include "tc.cuh"

namespace b2 {

static unsigned long long* g_attn_dbg = nullptr;
__device__ __forceinline__ unsigned long long clk() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
  return t;
}

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
  unsigned long long* dbg;  // optional per-phase cycle counters (b2_attn_set_debug), NULL in production
};

// Packed fp32 pairs (sm_100 FFMA2 / FADD2 / FMUL2: one instruction for two lanes of work).  The softmax warps of the backward
// kernels walk a dependent chain LDTM -> FFMA -> ex2 -> FADD -> FMUL -> F2FP -> STTM with one warp per scheduler per group;
// fewer instructions on that chain shortened it measurably (458 -> 350 per 64-query block, backward 145 -> 138 us at n = 1024,
// 777 -> 707 us at n = 4096; profiles/r2_attention_notes.md).  ncu: issue slots are only ~33 % busy — the chain's latency,
// not its issue rate, is what the packing buys back.
__device__ __forceinline__ float2 f2(uint32_t a, uint32_t b) { return make_float2(__uint_as_float(a), __uint_as_float(b)); }

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
// v3 kernels: one CTA per SM, a RING of three S (or S/dP) accumulators in TMEM, two softmax warp groups that take
// alternate key (query) blocks.
//   What the measurements said (profiles/r1_attention_notes.md): the v1 kernels (MMA -> exp2 -> MMA inside a CTA, two
//   CTAs per SM) and a first ping-pong rewrite both sat at ~2x the exp2 bound; the time went to the HAND-OFF, not to the
//   math: "P written -> mbarrier -> MMA thread -> PV + next S MMA -> commit -> mbarrier -> softmax warps" costs about
//   as much as the exp2 work of a whole block, and every block paid it.  With three S buffers the MMA thread runs up to
//   three blocks ahead, so a group that finishes block j finds S(j+2) already waiting; the hand-off only gates how far
//   ahead the tensor pipe may run.
//   Softmax loops hold a whole block row in registers (ONE tcgen05.ld round trip per block), reduce with four
//   independent chains, and write the bf16 result in place over the columns they have just consumed.
// warps (320 threads): 0 = TMA producer, 1 = MMA issuer + TMEM allocator, 2..5 = group 0, 6..9 = group 1.
// ---------------------------------------------------------------------------------------------
constexpr int A3_THREADS = 320;
constexpr int A3_TMEM_COLS = 512;

__device__ __forceinline__ void a3_group_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void store_row32(bf16* dst, const float* f) {
#pragma unroll
  for (int g = 0; g < 4; ++g) st8(dst + g * 8, pack8(f + g * 8));
}

// forward: CTA = 128 queries x key blocks of 128.   TMEM: S ring [0,384) (P in place, 64 columns each), O_0 [384,448),
// O_1 [448,512): each group keeps its own online-softmax state (m, l, O) over its alternate blocks; the two partial
// results are merged at the end like a split-KV decode (exact: O = (w0 O_0 + w1 O_1) / (w0 l_0 + w1 l_1)).
constexpr int F3_STAGES = 5;
constexpr int F3_SMEM = AT_TILE128 + F3_STAGES * 2 * AT_TILE128 + 1024;

__global__ void __launch_bounds__(A3_THREADS, 1)
attn_fwd3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnP p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_q, bar_full[F3_STAGES], bar_empty[F3_STAGES], bar_s[3], bar_p[3], bar_pv[2], bar_o;
  __shared__ uint32_t tmem_slot;
  __shared__ float2 ml[2][128];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sKV = smem_base + AT_TILE128;
  const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nkb = (p.n_k + 127) / 128;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(smem_u32(&bar_q), 1);
#pragma unroll
    for (int s = 0; s < F3_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      mbar_init(smem_u32(&bar_s[i]), 1);
      mbar_init(smem_u32(&bar_p[i]), 128);
    }
    mbar_init(smem_u32(&bar_pv[0]), 1);
    mbar_init(smem_u32(&bar_pv[1]), 1);
    mbar_init(smem_u32(&bar_o), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), A3_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(tmem_slot);
  pdl_wait();  // set-up above overlaps the previous kernel's tail (PDL); its outputs are visible from here on
  pdl_launch_dependents();

  if (warp == 0) {
    const bool el = elect_one();  // warp-uniform loop, elected lane issues (see tc.cuh)
    {
      if (el) {
      mbar_expect_tx(smem_u32(&bar_q), AT_TILE128);
        tma_load_4d(sQ, &tmQ, smem_u32(&bar_q), 0, q0, h, b);
      }
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nkb; ++j) {
        mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
        const uint32_t full = smem_u32(&bar_full[s]);
        if (el) {
        mbar_expect_tx(full, 2 * AT_TILE128);
          tma_load_4d(sKV + s * 2 * AT_TILE128, &tmK, full, 0, j * 128, h, b);
          tma_load_4d(sKV + s * 2 * AT_TILE128 + AT_TILE128, &tmV, full, 0, j * 128, h, b);
        }
        if (++s == F3_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const bool el = elect_one();
    {
      constexpr uint32_t idS = umma_idesc(128, 128, 0, 0);
      constexpr uint32_t idO = umma_idesc(128, AT_D, 0, 1);
      auto issue_S = [&](int buf, int stage) {
        const uint32_t tS = tmem_base + buf * 128, sK = sKV + stage * 2 * AT_TILE128;
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k)
          umma_bf16(tS, umma_desc(sQ + k * 32, 16, 1024), umma_desc(sK + k * 32, 16, 1024), idS, k != 0);
      };
      mbar_wait(smem_u32(&bar_q), 0);
      // run-ahead: S(0), S(1), S(2)
      int ls = 0;          // stage / phase of the next block whose S gets issued
      uint32_t lph = 0;
      int issued = 0;
      for (; issued < 3 && issued < nkb; ++issued) {
        mbar_wait(smem_u32(&bar_full[ls]), lph);
        tc_fence_after();
        if (el) issue_S(issued, ls);
        if (el) umma_commit(smem_u32(&bar_s[issued]));
        if (++ls == F3_STAGES) { ls = 0; lph ^= 1u; }
      }
      int buf = 0, cs = 0;  // ring slot and KV stage of block j
      uint32_t ppar = 0;    // parity of bar_p[buf] for block j (flips when buf wraps)
      unsigned long long m_waitp = 0, m_waitf = 0, m_pv = 0, m_s = 0, m_cm = 0, mt0 = 0, mt1 = 0;
      const bool dbg = p.dbg != nullptr;
      for (int j = 0; j < nkb; ++j) {
        const int g = j & 1;
        if (dbg) mt0 = clk();
        mbar_wait(smem_u32(&bar_p[buf]), ppar);
        tc_fence_after();
        if (dbg) { mt1 = clk(); m_waitp += mt1 - mt0; mt0 = mt1; }
        const uint32_t tP = tmem_base + buf * 128, tO = tmem_base + 384 + g * 64;
        const uint32_t sV = sKV + cs * 2 * AT_TILE128 + AT_TILE128;
        if (el) {
#pragma unroll
          for (int k = 0; k < 128 / 16; ++k)
            umma_bf16_ts(tO, tP + k * 8, umma_desc(sV + k * 2048, 16384, 1024), idO, (j >= 2) || k != 0);
        }
        if (dbg) { mt1 = clk(); m_pv += mt1 - mt0; mt0 = mt1; }
        if (el) umma_commit(smem_u32(&bar_pv[g]));
        if (el) umma_commit(smem_u32(&bar_empty[cs]));
        if (dbg) { mt1 = clk(); m_cm += mt1 - mt0; mt0 = mt1; }
        if (issued < nkb) {  // refill this ring slot with S(j + 3)
          if (dbg) mt0 = clk();
          mbar_wait(smem_u32(&bar_full[ls]), lph);
          tc_fence_after();
          if (dbg) { mt1 = clk(); m_waitf += mt1 - mt0; mt0 = mt1; }
          if (el) issue_S(buf, ls);
          if (dbg) { mt1 = clk(); m_s += mt1 - mt0; mt0 = mt1; }
          if (el) umma_commit(smem_u32(&bar_s[buf]));
          if (dbg) { mt1 = clk(); m_cm += mt1 - mt0; mt0 = mt1; }
          if (++ls == F3_STAGES) { ls = 0; lph ^= 1u; }
          ++issued;
        }
        if (++cs == F3_STAGES) cs = 0;
        if (++buf == 3) { buf = 0; ppar ^= 1u; }
      }
      if (el) umma_commit(smem_u32(&bar_o));
      if (dbg && el) { atomicAdd(p.dbg + 16, m_waitp); atomicAdd(p.dbg + 17, m_waitf); atomicAdd(p.dbg + 18, 1ull);
                 atomicAdd(p.dbg + 19, m_pv); atomicAdd(p.dbg + 20, m_s); atomicAdd(p.dbg + 21, m_cm); }
    }
  } else {
    const int g = (warp - 2) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = uint32_t(qd * 32) << 16;
    const uint32_t tOg = tmem_base + 384 + g * 64 + lane_off;
    float m_used = -INFINITY, l = 0.f;
    int buf = g;          // ring slot of block j = g, g + 2, ...
    uint32_t spar = 0;    // parity of bar_s[buf] for that block
    int kown = 0;         // own blocks done
    unsigned long long d_wait = 0, d_ld = 0, d_cmp = 0, d_st = 0, d_arr = 0, t0 = 0, t1 = 0;
    const bool dbg = p.dbg != nullptr;
    const unsigned long long t_begin = dbg ? clk() : 0;
    for (int j = g; j < nkb; j += 2, ++kown) {
      if (dbg) t0 = clk();
      mbar_wait(smem_u32(&bar_s[buf]), spar);
      tc_fence_after();
      if (dbg) { t1 = clk(); d_wait += t1 - t0; t0 = t1; }
      const uint32_t tS = tmem_base + buf * 128 + lane_off;
      const int valid = min(128, p.n_k - j * 128);
      uint32_t r[128];
      tmem_ld32_nowait(tS, r);
      tmem_ld32_nowait(tS + 32, r + 32);
      tmem_ld32_nowait(tS + 64, r + 64);
      tmem_ld32_nowait(tS + 96, r + 96);
      tmem_ld_wait();
      if (dbg) { t1 = clk(); d_ld += t1 - t0; t0 = t1; }
      if (valid < 128) {  // ragged last key block: -inf logits -> P = 0
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
      if (kown == 0) {
        m_used = mx;
      } else if (mx > m_used + 8.f) {  // lazy rescale: a stale max is fine while 2^(s - m) <= 2^8
        factor = fast_exp2(m_used - mx);
        m_used = mx;
      }
      if (kown > 0 && __any_sync(AT_FULL, factor != 1.f)) {
        // O_g is being accumulated by this group's previous PV MMA: wait for it before touching the accumulator
        mbar_wait(smem_u32(&bar_pv[g]), (uint32_t)(kown - 1) & 1u);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t o[32];
          tmem_ld32(tOg + cc * 32, o);
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
          tmem_st32(tOg + cc * 32, o);
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
        tmem_st16(tS + cc * 16, pk);  // in place: this thread's row was read out completely above
      }
      l += (l0 + l1) + (l2 + l3);
      if (dbg) { t1 = clk(); d_cmp += t1 - t0; t0 = t1; }
      tmem_st_wait();
      tc_fence_before();
      if (dbg) { t1 = clk(); d_st += t1 - t0; t0 = t1; }
      mbar_arrive(smem_u32(&bar_p[buf]));
      if (dbg) { t1 = clk(); d_arr += t1 - t0; t0 = t1; }
      buf += 2;
      if (buf >= 3) { buf -= 3; spar ^= 1u; }
    }
    // ---- merge the two groups' partial results (split-KV combine), 32 output columns per group
    ml[g][row] = make_float2(m_used, l);
    const unsigned long long t_loop_end = dbg ? clk() : 0;
    mbar_wait(smem_u32(&bar_o), 0);
    tc_fence_after();
    a3_group_sync();
    const float2 s0 = ml[0][row], s1 = ml[1][row];
    const bool has1 = nkb > 1;
    const float m = has1 ? fmaxf(s0.x, s1.x) : s0.x;
    const float w0 = fast_exp2(s0.x - m), w1 = has1 ? fast_exp2(s1.x - m) : 0.f;
    const float lt = s0.y * w0 + s1.y * w1;
    const float inv = 1.f / lt;
    const uint32_t tO0 = tmem_base + 384 + g * 32 + lane_off, tO1 = tO0 + 64;
    uint32_t a[32], c[32];
    float f[32];
    tmem_ld32_nowait(tO0, a);
    if (has1) tmem_ld32_nowait(tO1, c);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i)
      f[i] = (__uint_as_float(a[i]) * w0 + (has1 ? __uint_as_float(c[i]) * w1 : 0.f)) * inv;
    const int gq = q0 + row;
    if (gq < p.n_q) store_row32(p.out0 + (long long)b * p.bs0 + (long long)gq * p.ld0 + h * AT_D + g * 32, f);
    if (g == 0 && gq < p.n_pad) p.LSE[((long long)b * p.H + h) * p.n_pad + gq] = gq < p.n_q ? m + log2f(lt) : INFINITY;
    if (dbg && lane == 0 && qd == 0) {
      const unsigned long long t_end = clk();
      unsigned long long* d = p.dbg + g * 8;
      atomicAdd(d + 0, d_wait); atomicAdd(d + 1, d_ld); atomicAdd(d + 2, d_cmp); atomicAdd(d + 3, d_st);
      atomicAdd(d + 4, d_arr); atomicAdd(d + 5, t_loop_end - t_begin); atomicAdd(d + 6, t_end - t_loop_end);
      atomicAdd(d + 7, 1ull);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, A3_TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// forward, PERSISTENT (attn_pfwd2): the v3 pipeline with the query tile as an outer loop inside the CTA.
//   What round 2's first GPU run said about the first persistent draft (profiles/r2_attention_notes.md): parity green but
//   SLOWER than v3 (n = 1024: 70 vs 55 us; n = 4096: 366 vs 262 us) for two reasons visible without a profiler —
//   ptxas spilled 512 bytes in the softmax loop (168 registers per thread at 320 threads, one TMEM row of 128 logits live),
//   and the two-group merge sat between tiles un-overlapped (the group that finishes first idles for a whole block).
//   This version fixes both and trims the MMA thread's critical path:
//   * 384 threads = three warpgroups: softmax group 0, softmax group 1, {TMA producer, MMA issuer, 2 idle warps}.
//     setmaxnreg moves registers from the third warpgroup (56) to the softmax warpgroups (224): no spills.
//   * DEFERRED MERGE: a group that has finished its last key block of tile t goes straight on to its first key block of
//     tile t+1 (whose S was issued ahead) and only then merges tile t; the MMA thread holds back the first PV of tile t+1
//     (which overwrites O) until both groups have read O of tile t (bar_ofree).  The skew between the groups and the
//     PV-completion latency are hidden behind a block of useful softmax work.
//   * MMA thread: descriptors built once and advanced with one add (tc.cuh desc_adv), the K/V-stage wait for the S that
//     will be issued after a PV is taken BEFORE waiting for P (off the critical path), and no separate "PV done" barrier:
//     the lazy-rescale path waits on the K/V stage's bar_empty, which the same PV commits.
// Every ring keeps running across tiles on a global block counter G: S ring slot G % 3, K/V stage G % 5.
// warps: 0..3 = softmax group 0, 4..7 = group 1 (warp w owns TMEM lanes 32*(w%4)..+31), 8 = TMA producer, 9 = MMA issuer +
// TMEM allocator, 10..11 = idle.
// ---------------------------------------------------------------------------------------------
constexpr int P2_THREADS = 384;
constexpr int P2_SMEM = 2 * AT_TILE128 + F3_STAGES * 2 * AT_TILE128 + 1024;

__device__ __forceinline__ void p2_groups_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// RAGGED: n_k is not a multiple of 128 (aspect buckets: 576, 1200, ... keys) — the last key block's out-of-range columns are
// masked to -inf.  Compiled out for the multiple-of-128 case (every 1024^2 shape): the masking made ptxas copy and spill a
// quarter of the logits row in BOTH paths.
template <bool RAGGED>
__global__ void __launch_bounds__(P2_THREADS, 1)
attn_pfwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const AttnP p, int nqt, int total_tiles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_q[2], bar_qfree[2], bar_full[F3_STAGES], bar_empty[F3_STAGES], bar_s[3], bar_p[3],
      bar_o, bar_ofree;
  __shared__ uint32_t tmem_slot;
  __shared__ float2 ml[2][2][128];  // [tile parity][group][row]
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ0 = smem_base;                     // two Q buffers
  const uint32_t sKV = smem_base + 2 * AT_TILE128;
  const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int t_begin = (int)(((long long)blockIdx.x * total_tiles) / gridDim.x);
  const int t_end = (int)(((long long)(blockIdx.x + 1) * total_tiles) / gridDim.x);
  const int ntiles = t_end - t_begin;
  const int nkb = (p.n_k + 127) / 128;

  if (threadIdx.x == 256) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_q[i]), 1);
      mbar_init(smem_u32(&bar_qfree[i]), 1);
    }
#pragma unroll
    for (int s = 0; s < F3_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      mbar_init(smem_u32(&bar_s[i]), 1);
      mbar_init(smem_u32(&bar_p[i]), 128);
    }
    mbar_init(smem_u32(&bar_o), 1);
    mbar_init(smem_u32(&bar_ofree), 256);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(smem_u32(&tmem_slot), A3_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(tmem_slot);
  pdl_wait();
  pdl_launch_dependents();

  if (warp >= 8) {
    setmaxnreg_dec<56>();
    if (warp == 8) {
      // ---------------- TMA producer: Q(t), then the K/V blocks of tile t, for every tile of this CTA
      const bool el = elect_one();
      int s = 0;
      uint32_t ph = 0;
      for (int tl = 0; tl < ntiles; ++tl) {
        const int gt = t_begin + tl;
        const int qt = gt % nqt, bh = gt / nqt;
        const int h = bh % p.H, b = bh / p.H;
        const int qb = tl & 1;
        if (tl >= 2) mbar_wait(smem_u32(&bar_qfree[qb]), (uint32_t)((tl >> 1) - 1) & 1u);  // S MMAs of tile tl-2 are done
        if (el) {
          mbar_expect_tx(smem_u32(&bar_q[qb]), AT_TILE128);
          tma_load_4d(sQ0 + qb * AT_TILE128, &tmQ, smem_u32(&bar_q[qb]), 0, qt * 128, h, b);
        }
        for (int j = 0; j < nkb; ++j) {
          mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
          const uint32_t full = smem_u32(&bar_full[s]);
          if (el) {
            mbar_expect_tx(full, 2 * AT_TILE128);
            tma_load_4d(sKV + s * 2 * AT_TILE128, &tmK, full, 0, j * 128, h, b);
            tma_load_4d(sKV + s * 2 * AT_TILE128 + AT_TILE128, &tmV, full, 0, j * 128, h, b);
          }
          if (++s == F3_STAGES) { s = 0; ph ^= 1u; }
        }
      }
    } else if (warp == 9) {
      // ---------------- MMA issuer: PV cursor (tile tl, block j) and an S cursor up to three blocks ahead
      const bool el = elect_one();
      constexpr uint32_t idS = umma_idesc(128, 128, 0, 0);
      constexpr uint32_t idO = umma_idesc(128, AT_D, 0, 1);
      constexpr uint32_t STAGE_B = 2 * AT_TILE128;
      const uint64_t dQ = umma_desc(sQ0, 16, 1024);                    // + qb * TILE128; k-step + 32 B
      const uint64_t dK = umma_desc(sKV, 16, 1024);                    // + stage * STAGE_B; k-step + 32 B
      const uint64_t dV = umma_desc(sKV + AT_TILE128, 16384, 1024);    // + stage * STAGE_B; k-step + 2048 B
      int s_tl = 0, s_j = 0;   // S cursor: tile, block within the tile
      int s_stage = 0;         // K/V stage of the S cursor's block (= global block index % F3_STAGES)
      uint32_t s_sph = 0;      // parity of bar_full[s_stage] for that block
      int s_buf = 0;           // S ring slot of the S cursor's block (= global block index % 3)
      // the waits an S issue needs (its tile's Q, its K/V stage); both complete long before they are needed in steady state
      auto wait_next_S = [&]() {
        if (s_tl >= ntiles) return;
        if (s_j == 0) mbar_wait(smem_u32(&bar_q[s_tl & 1]), (uint32_t)(s_tl >> 1) & 1u);
        mbar_wait(smem_u32(&bar_full[s_stage]), s_sph);
      };
      auto issue_next_S = [&]() {
        if (s_tl >= ntiles) return;
        const int qb = s_tl & 1;
        const uint32_t tS = tmem_base + s_buf * 128;
        const uint64_t dq = desc_adv(dQ, qb * AT_TILE128), dk = desc_adv(dK, s_stage * STAGE_B);
        if (el) {
#pragma unroll
          for (int k = 0; k < AT_D / 16; ++k) umma_bf16(tS, desc_adv(dq, k * 32), desc_adv(dk, k * 32), idS, k != 0);
          umma_commit(smem_u32(&bar_s[s_buf]));
        }
        if (++s_stage == F3_STAGES) { s_stage = 0; s_sph ^= 1u; }
        if (++s_buf == 3) s_buf = 0;
        if (++s_j == nkb) {
          if (el) umma_commit(smem_u32(&bar_qfree[qb]));  // this tile's Q buffer may be reloaded once its S MMAs completed
          s_j = 0;
          ++s_tl;
        }
      };
      for (int i = 0; i < 3; ++i) {
        wait_next_S();
        tc_fence_after();
        issue_next_S();
      }
      int buf = 0, cs = 0;   // ring slot / K/V stage of the PV cursor's block
      uint32_t ppar = 0;     // parity of bar_p[buf] for that block
      for (int tl = 0; tl < ntiles; ++tl) {
        for (int j = 0; j < nkb; ++j) {
          const int g = j & 1;
          wait_next_S();  // off the critical path: taken while the softmax warps are still working on this block's P
          mbar_wait(smem_u32(&bar_p[buf]), ppar);
          if (j == 0 && tl > 0) mbar_wait(smem_u32(&bar_ofree), (uint32_t)(tl - 1) & 1u);  // previous tile's merge has read O
          tc_fence_after();
          const uint32_t tP = tmem_base + buf * 128, tO = tmem_base + 384 + g * 64;
          const uint64_t dv = desc_adv(dV, cs * STAGE_B);
          if (el) {
#pragma unroll
            for (int k = 0; k < 128 / 16; ++k) umma_bf16_ts(tO, tP + k * 8, desc_adv(dv, k * 2048), idO, (j >= 2) || k != 0);
            umma_commit(smem_u32(&bar_empty[cs]));  // K/V stage free; also "PV of this block done" for the lazy rescale
            if (j == nkb - 1) umma_commit(smem_u32(&bar_o));  // every PV of this tile has been issued
          }
          issue_next_S();  // refills ring slot `buf` (global block + 3), possibly with a block of the NEXT tile
          if (++cs == F3_STAGES) cs = 0;
          if (++buf == 3) { buf = 0; ppar ^= 1u; }
        }
      }
    }
  } else {
    setmaxnreg_inc<224>();
    // ---------------- softmax groups
    const int g = warp >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = uint32_t(qd * 32) << 16;
    const uint32_t tOg = tmem_base + 384 + g * 64 + lane_off;
    const bool has1 = nkb > 1;
    // merge of tile `tl` (whose partial state this thread saved as m_t, l_t): O = (w0 O_0 + w1 O_1) / (w0 l_0 + w1 l_1),
    // 32 output columns per group; hands the O accumulators back to the MMA thread as soon as they are in registers
    auto merge_tile = [&](int tl, float m_t, float l_t) {
      const int gt = t_begin + tl;
      const int qt = gt % nqt, bh = gt / nqt;
      const int h = bh % p.H, b = bh / p.H;
      ml[tl & 1][g][row] = make_float2(m_t, l_t);
      mbar_wait(smem_u32(&bar_o), (uint32_t)tl & 1u);
      tc_fence_after();
      p2_groups_sync();
      const float2 s0 = ml[tl & 1][0][row], s1 = ml[tl & 1][1][row];
      const float m = has1 ? fmaxf(s0.x, s1.x) : s0.x;
      const float w0 = fast_exp2(s0.x - m), w1 = has1 ? fast_exp2(s1.x - m) : 0.f;
      const float lt = s0.y * w0 + s1.y * w1;
      const float inv = 1.f / lt;
      const uint32_t tO0 = tmem_base + 384 + g * 32 + lane_off, tO1 = tO0 + 64;
      uint32_t a[32], c[32];
      float f[32];
      tmem_ld32_nowait(tO0, a);
      if (has1) tmem_ld32_nowait(tO1, c);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_ofree));  // the MMA warp may start the next tile's PV into O_0 / O_1
#pragma unroll
      for (int i = 0; i < 32; ++i)
        f[i] = (__uint_as_float(a[i]) * w0 + (has1 ? __uint_as_float(c[i]) * w1 : 0.f)) * inv;
      const int gq = qt * 128 + row;
      if (gq < p.n_q) store_row32(p.out0 + (long long)b * p.bs0 + (long long)gq * p.ld0 + h * AT_D + g * 32, f);
      if (g == 0 && gq < p.n_pad) p.LSE[((long long)b * p.H + h) * p.n_pad + gq] = gq < p.n_q ? m + log2f(lt) : INFINITY;
    };
    int gbase = 0;   // global index of the current tile's first block
    float m_prev = 0.f, l_prev = 0.f;
    for (int tl = 0; tl < ntiles; ++tl, gbase += nkb) {
      float m_used = -INFINITY, l = 0.f;
      int kown = 0;
      bool merged_prev = tl == 0;
      for (int j = g; j < nkb; j += 2, ++kown) {
        const int G = gbase + j;
        const int buf = G % 3;
        const uint32_t spar = (uint32_t)(G / 3) & 1u;
        mbar_wait(smem_u32(&bar_s[buf]), spar);
        tc_fence_after();
        const uint32_t tS = tmem_base + buf * 128 + lane_off;
        const int valid = min(128, p.n_k - j * 128);
        uint32_t r[128];
        tmem_ld32_nowait(tS, r);
        tmem_ld32_nowait(tS + 32, r + 32);
        tmem_ld32_nowait(tS + 64, r + 64);
        tmem_ld32_nowait(tS + 96, r + 96);
        tmem_ld_wait();
        if (RAGGED && valid < 128) {
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
        if (kown == 0) {
          m_used = mx;
        } else if (mx > m_used + 8.f) {  // lazy rescale: a stale max is fine while 2^(s - m) <= 2^8
          factor = fast_exp2(m_used - mx);
          m_used = mx;
        }
        if (kown > 0 && __any_sync(AT_FULL, factor != 1.f)) {
          // O_g is being accumulated by the PV of this group's previous block (global index G - 2): its completion is what
          // frees that block's K/V stage.  (That PV was issued after bar_ofree of the previous tile, so the merge — which
          // reads both groups' accumulators — is over too; no later completion of the same barrier can have happened:
          // PV(G + 3) is issued after PV(G), which needs this block's P.)
          const int Gp = G - 2;
          mbar_wait(smem_u32(&bar_empty[Gp % F3_STAGES]), (uint32_t)(Gp / F3_STAGES) & 1u);
          tc_fence_after();
#pragma unroll 1
          for (int cc = 0; cc < 2; ++cc) {
            uint32_t o[32];
            tmem_ld32(tOg + cc * 32, o);
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            tmem_st32(tOg + cc * 32, o);
          }
          l *= factor;
        }
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
        const float2 c2 = make_float2(p.c, p.c), nm2 = make_float2(-m_used, -m_used);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            // packed fp32 pairs (FFMA2) halve the issue slots of the scale; the row sum stays scalar: with FADD2 the
            // compiler moves every ex2 result into an aligned register pair first (one MOV per element, no gain)
            const float2 x01 = __ffma2_rn(f2(r[cc * 32 + 2 * i], r[cc * 32 + 2 * i + 1]), c2, nm2);
            const float2 x23 = __ffma2_rn(f2(r[cc * 32 + 2 * i + 2], r[cc * 32 + 2 * i + 3]), c2, nm2);
            const float p0 = fast_exp2(x01.x), p1 = fast_exp2(x01.y), p2 = fast_exp2(x23.x), p3 = fast_exp2(x23.y);
            l0 += p0; l1 += p1; l2 += p2; l3 += p3;
            pk[i] = pack_bf16x2(p0, p1);
            pk[i + 1] = pack_bf16x2(p2, p3);
          }
          tmem_st16(tS + cc * 16, pk);  // in place: this thread's row was read out completely above
        }
        l += (l0 + l1) + (l2 + l3);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(smem_u32(&bar_p[buf]));
        if (!merged_prev) {  // deferred merge of the previous tile, behind this tile's first block
          merge_tile(tl - 1, m_prev, l_prev);
          merged_prev = true;
        }
      }
      if (!merged_prev) merge_tile(tl - 1, m_prev, l_prev);  // this group has no block in the tile (one key block, group 1)
      m_prev = m_used;
      l_prev = l;
    }
    if (ntiles > 0) merge_tile(ntiles - 1, m_prev, l_prev);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, A3_TMEM_COLS);
  }
}


// ---------------------------------------------------------------------------------------------
// Cross-attention kernels (n_k <= 96 keys: SDXL's 77 text tokens = ONE key block).
//   With a single key block the v3 kernels spend their time in per-CTA set-up (640 CTAs x {TMEM alloc, barrier init, one
//   128x128 tile}): 29 us per layer call for 1.6 GFLOP / 22.6 MB (HBM floor 3.5 us).  Here the loop dimension is the
//   QUERY tile: a persistent CTA per SM walks a contiguous range of (sample, head, query-tile) work items, every stage of
//   the TMA ring carries {Q tile, K, V} (K/V are L2 hits after the first tile of a head), the MMA thread runs up to three
//   tiles ahead through a ring of S accumulators, and the two softmax groups take alternate tiles, each finishing its
//   tile (O / l, LSE, store) from its own O accumulator.  Keys are padded to a multiple of 16 (80), not to 128.
// ---------------------------------------------------------------------------------------------
constexpr int X_STAGES = 4;
constexpr int X_KV_ROWS = 96;
constexpr int X_KV_BYTES = X_KV_ROWS * AT_D * 2;                 // 12 KiB
constexpr int X_STAGE_BYTES = AT_TILE128 + 2 * X_KV_BYTES;       // Q | K | V
constexpr int X_SMEM = X_STAGES * X_STAGE_BYTES + 1024;

__global__ void __launch_bounds__(A3_THREADS, 1)
attn_xfwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnP p, int nqt, int total_tiles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[X_STAGES], bar_empty[X_STAGES], bar_s[3], bar_p[3], bar_o[2], bar_ofree[2];
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int t_begin = (int)(((long long)blockIdx.x * total_tiles) / gridDim.x);
  const int t_end = (int)(((long long)(blockIdx.x + 1) * total_tiles) / gridDim.x);
  const int ntiles = t_end - t_begin;
  const int nkp = (p.n_k + 15) & ~15;  // keys padded to the UMMA N granularity

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
#pragma unroll
    for (int s = 0; s < X_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      mbar_init(smem_u32(&bar_s[i]), 1);
      mbar_init(smem_u32(&bar_p[i]), 128);
    }
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      mbar_init(smem_u32(&bar_o[g]), 1);
      mbar_init(smem_u32(&bar_ofree[g]), 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), A3_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(tmem_slot);
  pdl_wait();  // set-up above overlaps the previous kernel's tail (PDL); its outputs are visible from here on
  pdl_launch_dependents();

  if (warp == 0) {
    const bool el = elect_one();
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < ntiles; ++i) {
      const int gt = t_begin + i;
      const int qt = gt % nqt, bh = gt / nqt;
      const int h = bh % p.H, b = bh / p.H;
      mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
      if (el) {
        const uint32_t full = smem_u32(&bar_full[s]);
        const uint32_t st = smem_base + s * X_STAGE_BYTES;
        mbar_expect_tx(full, X_STAGE_BYTES);
        tma_load_4d(st, &tmQ, full, 0, qt * 128, h, b);
        tma_load_4d(st + AT_TILE128, &tmK, full, 0, 0, h, b);
        tma_load_4d(st + AT_TILE128 + X_KV_BYTES, &tmV, full, 0, 0, h, b);
      }
      if (++s == X_STAGES) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    const bool el = elect_one();
    const uint32_t idS = umma_idesc(128, nkp, 0, 0);
    constexpr uint32_t idO = umma_idesc(128, AT_D, 0, 1);
    const int ksteps = nkp >> 4;
    // descriptors built once, advanced with one add per MMA (tc.cuh desc_adv)
    const uint64_t dQ = umma_desc(smem_base, 16, 1024), dK = umma_desc(smem_base + AT_TILE128, 16, 1024);
    const uint64_t dV = umma_desc(smem_base + AT_TILE128 + X_KV_BYTES, 16384, 1024);
    auto issue_S = [&](int buf, int stage) {
      const uint32_t tS = tmem_base + buf * 128;
      const uint64_t dq = desc_adv(dQ, stage * X_STAGE_BYTES), dk = desc_adv(dK, stage * X_STAGE_BYTES);
#pragma unroll
      for (int k = 0; k < AT_D / 16; ++k) umma_bf16(tS, desc_adv(dq, k * 32), desc_adv(dk, k * 32), idS, k != 0);
    };
    int ls = 0, issued = 0;
    uint32_t lph = 0;
    for (; issued < 3 && issued < ntiles; ++issued) {
      mbar_wait(smem_u32(&bar_full[ls]), lph);
      tc_fence_after();
      if (el) issue_S(issued, ls);
      if (el) umma_commit(smem_u32(&bar_s[issued]));
      if (++ls == X_STAGES) { ls = 0; lph ^= 1u; }
    }
    int buf = 0, cs = 0;
    uint32_t ppar = 0;
    for (int i = 0; i < ntiles; ++i) {
      const int g = i & 1, kown = i >> 1;
      if (issued < ntiles) mbar_wait(smem_u32(&bar_full[ls]), lph);  // stage of the S issued below: off the critical path
      mbar_wait(smem_u32(&bar_p[buf]), ppar);
      if (kown > 0) mbar_wait(smem_u32(&bar_ofree[g]), (uint32_t)(kown - 1) & 1u);  // group g drained O_g of its last tile
      tc_fence_after();
      const uint32_t tP = tmem_base + buf * 128, tO = tmem_base + 384 + g * 64;
      const uint64_t dv = desc_adv(dV, cs * X_STAGE_BYTES);
      if (el) {
        for (int k = 0; k < ksteps; ++k) umma_bf16_ts(tO, tP + k * 8, desc_adv(dv, k * 2048), idO, k != 0);
        umma_commit(smem_u32(&bar_o[g]));
        umma_commit(smem_u32(&bar_empty[cs]));
      }
      if (issued < ntiles) {
        if (el) issue_S(buf, ls);
        if (el) umma_commit(smem_u32(&bar_s[buf]));
        if (++ls == X_STAGES) { ls = 0; lph ^= 1u; }
        ++issued;
      }
      if (++cs == X_STAGES) cs = 0;
      if (++buf == 3) { buf = 0; ppar ^= 1u; }
    }
  } else {
    const int g = (warp - 2) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = uint32_t(qd * 32) << 16;
    const uint32_t tOg = tmem_base + 384 + g * 64 + lane_off;
    int buf = g;
    uint32_t spar = 0;
    int kown = 0;
    for (int i = g; i < ntiles; i += 2, ++kown) {
      const int gt = t_begin + i;
      const int qt = gt % nqt, bh = gt / nqt;
      const int h = bh % p.H, b = bh / p.H;
      mbar_wait(smem_u32(&bar_s[buf]), spar);
      tc_fence_after();
      const uint32_t tS = tmem_base + buf * 128 + lane_off;
      uint32_t r[96];
      tmem_ld32_nowait(tS, r);
      tmem_ld32_nowait(tS + 32, r + 32);
      if (nkp > 64) tmem_ld32_nowait(tS + 64, r + 64);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 96; ++j)
        if (j >= p.n_k) r[j] = 0xff800000u;  // padding keys (and never-written columns): -inf -> P = 0
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 96; j += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(r[j]));
        mx1 = fmaxf(mx1, __uint_as_float(r[j + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(r[j + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(r[j + 3]));
      }
      const float m = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.c;
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
      const float2 c2 = make_float2(p.c, p.c), nm2 = make_float2(-m, -m);
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float2 x01 = __ffma2_rn(f2(r[cc * 32 + 2 * j], r[cc * 32 + 2 * j + 1]), c2, nm2);
          const float2 x23 = __ffma2_rn(f2(r[cc * 32 + 2 * j + 2], r[cc * 32 + 2 * j + 3]), c2, nm2);
          const float p0 = fast_exp2(x01.x), p1 = fast_exp2(x01.y), p2 = fast_exp2(x23.x), p3 = fast_exp2(x23.y);
          l0 += p0; l1 += p1; l2 += p2; l3 += p3;
          pk[j] = pack_bf16x2(p0, p1);
          pk[j + 1] = pack_bf16x2(p2, p3);
        }
        if (cc * 32 < nkp) tmem_st16(tS + cc * 16, pk);  // bf16 P in place over the consumed S columns
      }
      const float l = (l0 + l1) + (l2 + l3);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_p[buf]));
      // ---- finish this tile: O_g = P V is complete when bar_o[g] fires
      mbar_wait(smem_u32(&bar_o[g]), (uint32_t)kown & 1u);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld32_nowait(tOg, r0);
      tmem_ld32_nowait(tOg + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_ofree[g]));  // the MMA thread may overwrite O_g with this group's next tile
      const int gq = qt * 128 + row;
      const float inv = 1.f / l;
      if (gq < p.n_q) store_row64(p.out0 + (long long)b * p.bs0 + (long long)gq * p.ld0 + h * AT_D, r0, r1, inv);
      if (gq < p.n_pad) p.LSE[((long long)b * p.H + h) * p.n_pad + gq] = gq < p.n_q ? m + log2f(l) : INFINITY;
      buf += 2;
      if (buf >= 3) { buf -= 3; spar ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, A3_TMEM_COLS);
  }
}

// backward dQ: CTA = 128 queries x key blocks of 64.  TMEM: ring of three {S | dP} pairs [0,384) (bf16 dS written in
// place over S), dQ accumulator [384,448).  Both groups feed the same dQ accumulator (the blocks are independent given
// LSE and D), so no merge is needed.
constexpr int Q3_STAGES = 6;
constexpr int Q3_SMEM = 2 * AT_TILE128 + Q3_STAGES * 2 * AT_TILE64 + 1024;
constexpr int DQ3_SMEM = Q3_SMEM + AT_TILE128;  // + the O tile: D = rowsum(dO * O) is computed here, not by a separate kernel

__global__ void __launch_bounds__(A3_THREADS, 1)
attn_bwd_dq3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                    const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnP p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_q, bar_t, bar_full[Q3_STAGES], bar_empty[Q3_STAGES], bar_s[3], bar_p[3], bar_o;
  __shared__ uint32_t tmem_slot;
  __shared__ float d_row[128];  // D[i] = sum_d dO[i,d] * O[i,d] of this CTA's 128 queries (was: attn_bwd_prep_kernel)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base, sdO = smem_base + AT_TILE128, sKV = smem_base + 2 * AT_TILE128;
  const uint32_t sO = sKV + Q3_STAGES * 2 * AT_TILE64;
  const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nkb = (p.n_k + 63) / 64;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(smem_u32(&bar_q), 1);
    mbar_init(smem_u32(&bar_t), 256);
#pragma unroll
    for (int s = 0; s < Q3_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      mbar_init(smem_u32(&bar_s[i]), 1);
      mbar_init(smem_u32(&bar_p[i]), 128);
    }
    mbar_init(smem_u32(&bar_o), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), A3_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(tmem_slot);
  pdl_wait();  // set-up above overlaps the previous kernel's tail (PDL); its outputs are visible from here on
  pdl_launch_dependents();

  if (warp == 0) {
    const bool el = elect_one();  // warp-uniform loop, elected lane issues (see tc.cuh)
    {
      if (el) {
      mbar_expect_tx(smem_u32(&bar_q), 3 * AT_TILE128);
        tma_load_4d(sQ, &tmQ, smem_u32(&bar_q), 0, q0, h, b);
        tma_load_4d(sdO, &tmdO, smem_u32(&bar_q), 0, q0, h, b);
        tma_load_4d(sO, &tmO, smem_u32(&bar_q), 0, q0, h, b);
      }
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nkb; ++j) {
        mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
        const uint32_t full = smem_u32(&bar_full[s]);
        if (el) {
        mbar_expect_tx(full, 2 * AT_TILE64);
          tma_load_4d(sKV + s * 2 * AT_TILE64, &tmK, full, 0, j * 64, h, b);
          tma_load_4d(sKV + s * 2 * AT_TILE64 + AT_TILE64, &tmV, full, 0, j * 64, h, b);
        }
        if (++s == Q3_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const bool el = elect_one();
    {
      constexpr uint32_t idS = umma_idesc(128, 64, 0, 0);
      constexpr uint32_t idQ = umma_idesc(128, AT_D, 0, 1);
      constexpr uint32_t STAGE_B = 2 * AT_TILE64;
      // descriptors are built once and advanced with one add per MMA (tc.cuh desc_adv): this thread's issue rate bounds the kernel
      // A of S / dP = this CTA's Q / dO tile: the same 128 x 64 operand for every key block.  Read from shared memory (SS mode)
      // an N = 64 MMA moves 4 KB of A + 2 KB of B per 32 tensor cycles = 192 B/clk against the 128 B/clk the SM's shared
      // memory delivers — measured 47 cycles per MMA instead of 32 (tools/attn_trace.py).  The softmax warps therefore copy
      // Q and dO once into TMEM columns [448, 512) as bf16 pairs and the MMAs take A from there (TS mode).
      const uint32_t tQ = tmem_base + 448, tdO = tmem_base + 480;
      const uint64_t dKk = umma_desc(sKV, 16, 1024), dVk = umma_desc(sKV + AT_TILE64, 16, 1024);  // B of S / dP (K-major)
      const uint64_t dKm = umma_desc(sKV, 8192, 1024);                                          // B of dQ += dS K (MN-major)
      auto issue_SdP = [&](int buf, int stage) {
        const uint32_t tS = tmem_base + buf * 128, tdP = tS + 64;
        const uint64_t dk = desc_adv(dKk, stage * STAGE_B), dv = desc_adv(dVk, stage * STAGE_B);
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k) umma_bf16_ts(tS, tQ + k * 8, desc_adv(dk, k * 32), idS, k != 0);
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k) umma_bf16_ts(tdP, tdO + k * 8, desc_adv(dv, k * 32), idS, k != 0);
      };
      mbar_wait(smem_u32(&bar_t), 0);  // Q and dO are in TMEM
      int ls = 0, issued = 0;
      uint32_t lph = 0;
      for (; issued < 3 && issued < nkb; ++issued) {
        mbar_wait(smem_u32(&bar_full[ls]), lph);
        tc_fence_after();
        if (el) issue_SdP(issued, ls);
        if (el) umma_commit(smem_u32(&bar_s[issued]));
        if (++ls == Q3_STAGES) { ls = 0; lph ^= 1u; }
      }
      int buf = 0, cs = 0;
      uint32_t ppar = 0;
      const uint32_t tdQ = tmem_base + 384;
      for (int j = 0; j < nkb; ++j) {
        if (issued < nkb) mbar_wait(smem_u32(&bar_full[ls]), lph);  // K/V of the S / dP issued below: off the critical path
        mbar_wait(smem_u32(&bar_p[buf]), ppar);
        tc_fence_after();
        const uint32_t tdS = tmem_base + buf * 128;
        const uint64_t dk = desc_adv(dKm, cs * STAGE_B);
        if (el) {
#pragma unroll
          for (int k = 0; k < 64 / 16; ++k) umma_bf16_ts(tdQ, tdS + k * 8, desc_adv(dk, k * 2048), idQ, (j | k) != 0);
          umma_commit(smem_u32(&bar_empty[cs]));
        }
        if (issued < nkb) {
          if (el) issue_SdP(buf, ls);
          if (el) umma_commit(smem_u32(&bar_s[buf]));
          if (++ls == Q3_STAGES) { ls = 0; lph ^= 1u; }
          ++issued;
        }
        if (++cs == Q3_STAGES) cs = 0;
        if (++buf == 3) { buf = 0; ppar ^= 1u; }
      }
      if (el) umma_commit(smem_u32(&bar_o));
    }
  } else {
    const int g = (warp - 2) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = uint32_t(qd * 32) << 16;
    const int gq = q0 + row;
    const long long sidx = ((long long)b * p.H + h) * p.n_pad + gq;
    const float L2 = p.LSE[sidx];   // n_pad is a multiple of 128: in bounds; +inf on pad rows -> P = 0
    {  // group 0 moves row `row` of Q, group 1 of dO, from the swizzled TMA tile into TMEM (bf16 pairs, K-major A operand);
       // group 1 also forms D[row] = sum_d dO * O from the row it is holding and publishes it (smem for this CTA, global
       // for the dK / dV kernel that runs next on this stream).  Pad rows are TMA zero fill: D = 0.
      mbar_wait(smem_u32(&bar_q), 0);
      const uint32_t src = (g == 0 ? sQ : sdO) + row * 128;
      uint32_t w[32];
#pragma unroll
      for (int c = 0; c < 8; ++c)
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(w[4 * c]), "=r"(w[4 * c + 1]), "=r"(w[4 * c + 2]), "=r"(w[4 * c + 3])
                     : "r"(src + ((c ^ (row & 7)) << 4)));
      if (g == 1) {
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t o4[4];
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(o4[0]), "=r"(o4[1]), "=r"(o4[2]), "=r"(o4[3])
                       : "r"(sO + row * 128 + ((c ^ (row & 7)) << 4)));
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 a2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[4 * c + e]));
            const float2 b2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&o4[e]));
            acc0 = fmaf(a2.x, b2.x, acc0);
            acc1 = fmaf(a2.y, b2.y, acc1);
          }
        }
        const float dsum = acc0 + acc1;
        d_row[row] = dsum;
        if (gq < p.n_pad) p.D[sidx] = dsum;
      }
      tmem_st32(tmem_base + 448 + g * 32 + lane_off, w);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_t));
    }
    a3_group_sync();  // d_row[] written by group 1 is visible to group 0
    const float Dr = d_row[row];
    int buf = g;
    uint32_t spar = 0;
    for (int j = g; j < nkb; j += 2) {
      mbar_wait(smem_u32(&bar_s[buf]), spar);
      tc_fence_after();
      const uint32_t tS = tmem_base + buf * 128 + lane_off, tdP = tS + 64;
      const int valid = min(64, p.n_k - j * 64);
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
          const float2 x = __ffma2_rn(f2(rs[cc * 32 + 2 * i], rs[cc * 32 + 2 * i + 1]), make_float2(p.c, p.c), make_float2(-L2, -L2));
          const float2 pe = make_float2(fast_exp2(x.x), fast_exp2(x.y));
          const float2 ds = __fmul2_rn(pe, __fadd2_rn(f2(rd[cc * 32 + 2 * i], rd[cc * 32 + 2 * i + 1]), make_float2(-Dr, -Dr)));
          pk[i] = pack_bf16x2(ds.x, ds.y);
        }
        tmem_st16(tS + cc * 16, pk);  // dS (bf16) in place over the consumed S columns
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_p[buf]));
      buf += 2;
      if (buf >= 3) { buf -= 3; spar ^= 1u; }
    }
    mbar_wait(smem_u32(&bar_o), 0);
    tc_fence_after();
    uint32_t a[32];
    float f[32];
    tmem_ld32(tmem_base + 384 + g * 32 + lane_off, a);
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(a[i]) * p.scale;
    if (gq < p.n_q) store_row32(p.out0 + (long long)b * p.bs0 + (long long)gq * p.ld0 + h * AT_D + g * 32, f);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, A3_TMEM_COLS);
  }
}

// backward dK / dV: CTA = 128 keys x query blocks of 64.  TMEM: ring of three {S^T | dP^T} pairs [0,384) (bf16 P^T / dS^T
// in place), dV [384,448), dK [448,512).  The S^T / dP^T MMAs that refill a ring slot are issued after the dV / dK MMAs
// that read it as their A operand; tcgen05.mma instructions of one thread execute in issue order.
__global__ void __launch_bounds__(A3_THREADS, 1)
attn_bwd_dkv3_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                     const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO, const AttnP p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_kv, bar_full[Q3_STAGES], bar_empty[Q3_STAGES], bar_s[3], bar_p[3], bar_o;
  __shared__ uint32_t tmem_slot;
  // LSE / D of each stage's 64 queries, staged by the producer with a bulk copy: reading them with __ldg from the softmax
  // warps cost ~2000 cycles per block (32 L2-latency loads the compiler cannot hoist past 128 live accumulator registers)
  __shared__ __align__(16) float ld_s[Q3_STAGES][128];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = smem_base, sV = smem_base + AT_TILE128, sQdO = smem_base + 2 * AT_TILE128;
  const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nqb = (p.n_q + 63) / 64;
  // b2_attn_set_debug(buffer of >= 320 uint64): besides the phase counters, CTA (1,0,0) leaves a clock64 trace of its first
  // 64 query blocks — [64 + 2i] P(i) seen by the MMA thread, [65 + 2i] block i issued, [192 + 2i] S(i) seen by its softmax
  // group, [193 + 2i] P(i) arrived (tools/attn_trace.py)
  const bool trace = p.dbg != nullptr && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmdO);
    mbar_init(smem_u32(&bar_kv), 1);
#pragma unroll
    for (int s = 0; s < Q3_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      mbar_init(smem_u32(&bar_s[i]), 1);
      mbar_init(smem_u32(&bar_p[i]), 128);
    }
    mbar_init(smem_u32(&bar_o), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), A3_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(tmem_slot);
  pdl_wait();  // set-up above overlaps the previous kernel's tail (PDL); its outputs are visible from here on
  pdl_launch_dependents();

  if (warp == 0) {
    const bool el = elect_one();  // warp-uniform loop, elected lane issues (see tc.cuh)
    {
      if (el) {
      mbar_expect_tx(smem_u32(&bar_kv), 2 * AT_TILE128);
        tma_load_4d(sK, &tmK, smem_u32(&bar_kv), 0, k0, h, b);
        tma_load_4d(sV, &tmV, smem_u32(&bar_kv), 0, k0, h, b);
      }
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nqb; ++i) {
        mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
        const uint32_t full = smem_u32(&bar_full[s]);
        if (el) {
          mbar_expect_tx(full, 2 * AT_TILE64 + 512);
          tma_load_4d(sQdO + s * 2 * AT_TILE64, &tmQ, full, 0, i * 64, h, b);
          tma_load_4d(sQdO + s * 2 * AT_TILE64 + AT_TILE64, &tmdO, full, 0, i * 64, h, b);
          const long long off = ((long long)b * p.H + h) * p.n_pad + (long long)i * 64;
          bulk_load_1d(smem_u32(&ld_s[s][0]), p.LSE + off, 256, full);
          bulk_load_1d(smem_u32(&ld_s[s][64]), p.D + off, 256, full);
        }
        if (++s == Q3_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const bool el = elect_one();
    {
      constexpr uint32_t idS = umma_idesc(128, 64, 0, 0);
      constexpr uint32_t idG = umma_idesc(128, AT_D, 0, 1);
      constexpr uint32_t STAGE_B = 2 * AT_TILE64;
      // descriptors are built once and advanced with one add per MMA (tc.cuh desc_adv): 16 MMAs per 64-query block go through
      // this one thread, and rebuilding both descriptors from addresses (~11 instructions per MMA) made it the kernel's bound
      const uint64_t dKa = umma_desc(sK, 16, 1024), dVa = umma_desc(sV, 16, 1024);                 // A of S^T / dP^T (K-major)
      const uint64_t dQk = umma_desc(sQdO, 16, 1024), ddOk = umma_desc(sQdO + AT_TILE64, 16, 1024);  // B of S^T / dP^T (K-major)
      const uint64_t dQm = umma_desc(sQdO, 8192, 1024), ddOm = umma_desc(sQdO + AT_TILE64, 8192, 1024);  // B of dK / dV (MN-major)
      auto issue_SdP = [&](int buf, int stage) {
        const uint32_t tS = tmem_base + buf * 128, tdP = tS + 64;
        const uint64_t dq = desc_adv(dQk, stage * STAGE_B), ddo = desc_adv(ddOk, stage * STAGE_B);
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k) umma_bf16(tS, desc_adv(dKa, k * 32), desc_adv(dq, k * 32), idS, k != 0);
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k) umma_bf16(tdP, desc_adv(dVa, k * 32), desc_adv(ddo, k * 32), idS, k != 0);
      };
      mbar_wait(smem_u32(&bar_kv), 0);
      int ls = 0, issued = 0;
      uint32_t lph = 0;
      for (; issued < 3 && issued < nqb; ++issued) {
        mbar_wait(smem_u32(&bar_full[ls]), lph);
        tc_fence_after();
        if (el) issue_SdP(issued, ls);
        if (el) umma_commit(smem_u32(&bar_s[issued]));
        if (++ls == Q3_STAGES) { ls = 0; lph ^= 1u; }
      }
      int buf = 0, cs = 0;
      uint32_t ppar = 0;
      const uint32_t tdV = tmem_base + 384, tdK = tmem_base + 448;
      for (int i = 0; i < nqb; ++i) {
        if (issued < nqb) mbar_wait(smem_u32(&bar_full[ls]), lph);  // {Q, dO} of the S^T / dP^T issued below: off the critical path
        mbar_wait(smem_u32(&bar_p[buf]), ppar);
        tc_fence_after();
        if (trace && el && i < 64) p.dbg[64 + 2 * i] = clk();        // event trace of CTA (0,0,0): P(i) seen by the MMA thread
        const uint32_t tP = tmem_base + buf * 128, tdS = tP + 64;
        const uint64_t ddo = desc_adv(ddOm, cs * STAGE_B), dq = desc_adv(dQm, cs * STAGE_B);
        if (el) {
#pragma unroll
          for (int k = 0; k < 64 / 16; ++k) umma_bf16_ts(tdV, tP + k * 8, desc_adv(ddo, k * 2048), idG, (i | k) != 0);
#pragma unroll
          for (int k = 0; k < 64 / 16; ++k) umma_bf16_ts(tdK, tdS + k * 8, desc_adv(dq, k * 2048), idG, (i | k) != 0);
          umma_commit(smem_u32(&bar_empty[cs]));
        }
        if (issued < nqb) {
          if (el) issue_SdP(buf, ls);
          if (el) umma_commit(smem_u32(&bar_s[buf]));
          if (++ls == Q3_STAGES) { ls = 0; lph ^= 1u; }
          ++issued;
        }
        if (trace && el && i < 64) p.dbg[65 + 2 * i] = clk();        // ... and everything for block i issued
        if (++cs == Q3_STAGES) cs = 0;
        if (++buf == 3) { buf = 0; ppar ^= 1u; }
      }
      if (el) umma_commit(smem_u32(&bar_o));
    }
  } else {
    const int g = (warp - 2) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = uint32_t(qd * 32) << 16;
    int buf = g, stg = g;   // ring slot and TMA stage of this group's current block
    uint32_t spar = 0, fpar = 0;
    unsigned long long d_wait = 0, d_ld = 0, d_cmp = 0, d_st = 0, t0 = 0, t1 = 0;
    const bool dbg = p.dbg != nullptr;
    for (int i = g; i < nqb; i += 2) {
      if (dbg) t0 = clk();
      mbar_wait(smem_u32(&bar_s[buf]), spar);
      tc_fence_after();
      if (dbg) { t1 = clk(); d_wait += t1 - t0; t0 = t1; }
      if (trace && qd == 0 && lane == 0 && i < 64) p.dbg[192 + 2 * i] = clk();   // S(i) seen by its softmax group
      const uint32_t tS = tmem_base + buf * 128 + lane_off, tdP = tS + 64;
      uint32_t rs[64], rd[64];
      tmem_ld32_nowait(tS, rs);
      tmem_ld32_nowait(tS + 32, rs + 32);
      tmem_ld32_nowait(tdP, rd);
      tmem_ld32_nowait(tdP + 32, rd + 32);
      // pad rows hold LSE = +inf (P = 0) and D = 0; the stage's full barrier (observed by the MMA thread before it issued
      // S^T) covers these bytes too, and bar_s was committed after that
      mbar_wait(smem_u32(&bar_full[stg]), fpar);  // already complete: makes the bulk-copied bytes visible to this thread
      const float4* L4 = reinterpret_cast<const float4*>(&ld_s[stg][0]);
      const float4* D4 = reinterpret_cast<const float4*>(&ld_s[stg][64]);
      tmem_ld_wait();
      if (dbg) { t1 = clk(); d_ld += t1 - t0; t0 = t1; }
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t pp[16], pd[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 lv = L4[cc * 8 + q], dv = D4[cc * 8 + q];
          const int o = cc * 32 + 4 * q;
          const float2 cc2 = make_float2(p.c, p.c);
          const float2 x01 = __ffma2_rn(f2(rs[o + 0], rs[o + 1]), cc2, make_float2(-lv.x, -lv.y));
          const float2 x23 = __ffma2_rn(f2(rs[o + 2], rs[o + 3]), cc2, make_float2(-lv.z, -lv.w));
          const float2 p01 = make_float2(fast_exp2(x01.x), fast_exp2(x01.y));
          const float2 p23 = make_float2(fast_exp2(x23.x), fast_exp2(x23.y));
          const float2 s01 = __fmul2_rn(p01, __fadd2_rn(f2(rd[o + 0], rd[o + 1]), make_float2(-dv.x, -dv.y)));
          const float2 s23 = __fmul2_rn(p23, __fadd2_rn(f2(rd[o + 2], rd[o + 3]), make_float2(-dv.z, -dv.w)));
          pp[2 * q] = pack_bf16x2(p01.x, p01.y);
          pp[2 * q + 1] = pack_bf16x2(p23.x, p23.y);
          pd[2 * q] = pack_bf16x2(s01.x, s01.y);
          pd[2 * q + 1] = pack_bf16x2(s23.x, s23.y);
        }
        tmem_st16(tS + cc * 16, pp);
        tmem_st16(tdP + cc * 16, pd);
      }
      if (dbg) { t1 = clk(); d_cmp += t1 - t0; t0 = t1; }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_p[buf]));
      if (dbg) { t1 = clk(); d_st += t1 - t0; t0 = t1; }
      if (trace && qd == 0 && lane == 0 && i < 64) p.dbg[193 + 2 * i] = clk();   // ... P(i) / dS(i) written, arrived
      buf += 2;
      if (buf >= 3) { buf -= 3; spar ^= 1u; }
      stg += 2;
      if (stg >= Q3_STAGES) { stg -= Q3_STAGES; fpar ^= 1u; }
    }
    if (dbg && lane == 0 && qd == 0) {
      unsigned long long* d = p.dbg + 24 + g * 4;
      atomicAdd(d + 0, d_wait); atomicAdd(d + 1, d_ld); atomicAdd(d + 2, d_cmp); atomicAdd(d + 3, d_st);
      if (g == 0) atomicAdd(p.dbg + 23, 1ull);
    }
    mbar_wait(smem_u32(&bar_o), 0);
    tc_fence_after();
    const int gk = k0 + row;
    uint32_t a[32];
    float f[32];
    tmem_ld32(tmem_base + 384 + g * 32 + lane_off, a);  // dV
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(a[i]);
    if (gk < p.n_k) store_row32(p.out1 + (long long)b * p.bs1 + (long long)gk * p.ld1 + h * AT_D + g * 32, f);
    tmem_ld32(tmem_base + 448 + g * 32 + lane_off, a);  // dK
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(a[i]) * p.scale;
    if (gk < p.n_k) store_row32(p.out0 + (long long)b * p.bs0 + (long long)gk * p.ld0 + h * AT_D + g * 32, f);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, A3_TMEM_COLS);
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

/* profiling hook (tools/attn_phase_timing.py): 32 uint64 counters accumulated by attn_fwd3_kernel (B2_ATTN_FWD3=1) and
   attn_bwd_dkv3_kernel, NULL disables */
extern "C" int b2_attn_set_debug(void* counters) {
  b2::g_attn_dbg = reinterpret_cast<unsigned long long*>(counters);
  return B2_OK;
}

extern "C" int b2_attn_fwd(const b2_attn_args* a, void* stream) {
  int rc = check_common(a, "b2_attn_fwd");
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tq;
  if ((rc = make_map_bf16_4d(&tq, a->Q, AT_D, a->n_q, a->H, a->B, a->ldq, AT_D, a->q_bs, 64, 128, "attn Q"))) return rc;
  AttnP p{};
  p.H = a->H; p.n_q = a->n_q; p.n_k = a->n_k; p.n_pad = b2_attn_lse_rows(a->n_q);
  p.scale = a->scale; p.c = a->scale * 1.4426950408889634f;
  p.LSE = a->LSE; p.D = nullptr;
  p.dbg = g_attn_dbg;
  p.out0 = (bf16*)a->O; p.ld0 = a->ldo; p.bs0 = a->o_bs;
  static const bool no_cross = getenv("B2_ATTN_NO_CROSS") != nullptr;
  if (!no_cross && a->n_k <= X_KV_ROWS) {  // cross-attention: query-tile-persistent kernel
    static bool configured_x = false;
    if (!configured_x) {
      if ((rc = set_smem(attn_xfwd_kernel, X_SMEM, "b2_attn_fwd"))) return rc;
      configured_x = true;
    }
    CUtensorMap tkx, tvx;
    if ((rc = make_map_bf16_4d(&tkx, a->K, AT_D, a->n_k, a->H, a->B, a->ldk, AT_D, a->k_bs, 64, X_KV_ROWS, "attn K96"))) return rc;
    if ((rc = make_map_bf16_4d(&tvx, a->V, AT_D, a->n_k, a->H, a->B, a->ldv, AT_D, a->v_bs, 64, X_KV_ROWS, "attn V96"))) return rc;
    const int nqt = (a->n_q + 127) / 128;
    const long long total = (long long)nqt * a->H * a->B;
    B2_REQUIRE(total < (1ll << 31), "b2_attn_fwd: too many tiles");
    const int grid = (int)(total < num_sms() ? total : num_sms());
    (void)launch_pdl(attn_xfwd_kernel, dim3(grid), dim3(A3_THREADS), (size_t)X_SMEM, st, tq, tkx, tvx, p, nqt, (int)total);
    return check_launch("b2_attn_fwd(cross)");
  }
  CUtensorMap tk, tv;
  if ((rc = make_map_bf16_4d(&tk, a->K, AT_D, a->n_k, a->H, a->B, a->ldk, AT_D, a->k_bs, 64, 128, "attn K"))) return rc;
  if ((rc = make_map_bf16_4d(&tv, a->V, AT_D, a->n_k, a->H, a->B, a->ldv, AT_D, a->v_bs, 64, 128, "attn V"))) return rc;
  static bool configured = false;
  if (!configured) {
    if ((rc = set_smem(attn_fwd3_kernel, F3_SMEM, "b2_attn_fwd"))) return rc;
    if ((rc = set_smem(attn_pfwd2_kernel<false>, P2_SMEM, "b2_attn_fwd"))) return rc;
    if ((rc = set_smem(attn_pfwd2_kernel<true>, P2_SMEM, "b2_attn_fwd"))) return rc;
    configured = true;
  }
  if (getenv("B2_ATTN_FWD3")) {  // the round-1 one-CTA-per-query-tile kernel, kept for A/B timing (read per call: tests switch it)
    const dim3 grid3((a->n_q + 127) / 128, a->H, a->B);
    (void)launch_pdl(attn_fwd3_kernel, grid3, dim3(A3_THREADS), (size_t)F3_SMEM, st, tq, tk, tv, p);
    return check_launch("b2_attn_fwd(v3)");
  }
  const int nqt = (a->n_q + 127) / 128;
  const long long total = (long long)nqt * a->H * a->B;
  B2_REQUIRE(total < (1ll << 31), "b2_attn_fwd: too many tiles");
  const int grid = (int)(total < num_sms() ? total : num_sms());
  if (a->n_k % 128)
    (void)launch_pdl(attn_pfwd2_kernel<true>, dim3(grid), dim3(P2_THREADS), (size_t)P2_SMEM, st, tq, tk, tv, p, nqt, (int)total);
  else
    (void)launch_pdl(attn_pfwd2_kernel<false>, dim3(grid), dim3(P2_THREADS), (size_t)P2_SMEM, st, tq, tk, tv, p, nqt, (int)total);
  return check_launch("b2_attn_fwd(persistent)");
}

extern "C" int b2_xattn_bwd_ok(int B, int H, int n_q, int n_k);
extern "C" int b2_xattn_bwd(const b2_attn_args* a, void* stream);

extern "C" int b2_attn_bwd(const b2_attn_args* a, void* stream) {
  int rc = check_common(a, "b2_attn_bwd");
  if (rc) return rc;
  B2_REQUIRE(a->dO && a->D && a->dQ && a->dK && a->dV, "b2_attn_bwd: null pointer");
  if (b2_xattn_bwd_ok(a->B, a->H, a->n_q, a->n_k)) return b2_xattn_bwd(a, stream);  // cross-attention: one-pass kernel (xattn_bwd.cu)
  cudaStream_t st = (cudaStream_t)stream;
  const int n_pad = b2_attn_lse_rows(a->n_q);
  // D = rowsum(dO * O) is formed inside the dQ kernel (which owns whole query tiles) and written to a->D for the dK / dV
  // kernel: the separate attn_bwd_prep_kernel launch (140 per step, 1.1 ms in profiles/r1_step_launches_v8_final.txt) is gone
  static bool configured = false;
  if (!configured) {
    if ((rc = set_smem(attn_bwd_dq3_kernel, DQ3_SMEM, "b2_attn_bwd"))) return rc;
    if ((rc = set_smem(attn_bwd_dkv3_kernel, Q3_SMEM, "b2_attn_bwd"))) return rc;
    configured = true;
  }
  AttnP p{};
  p.H = a->H; p.n_q = a->n_q; p.n_k = a->n_k; p.n_pad = n_pad;
  p.scale = a->scale; p.c = a->scale * 1.4426950408889634f;
  p.LSE = a->LSE; p.D = a->D;
  p.dbg = g_attn_dbg;
  CUtensorMap tq128, tdo128, to128, tk64, tv64, tk128, tv128, tq64, tdo64;
  if ((rc = make_map_bf16_4d(&tq128, a->Q, AT_D, a->n_q, a->H, a->B, a->ldq, AT_D, a->q_bs, 64, 128, "attn Q"))) return rc;
  if ((rc = make_map_bf16_4d(&tdo128, a->dO, AT_D, a->n_q, a->H, a->B, a->lddo, AT_D, a->do_bs, 64, 128, "attn dO"))) return rc;
  if ((rc = make_map_bf16_4d(&to128, a->O, AT_D, a->n_q, a->H, a->B, a->ldo, AT_D, a->o_bs, 64, 128, "attn O"))) return rc;
  if ((rc = make_map_bf16_4d(&tk64, a->K, AT_D, a->n_k, a->H, a->B, a->ldk, AT_D, a->k_bs, 64, 64, "attn K64"))) return rc;
  if ((rc = make_map_bf16_4d(&tv64, a->V, AT_D, a->n_k, a->H, a->B, a->ldv, AT_D, a->v_bs, 64, 64, "attn V64"))) return rc;
  if ((rc = make_map_bf16_4d(&tk128, a->K, AT_D, a->n_k, a->H, a->B, a->ldk, AT_D, a->k_bs, 64, 128, "attn K"))) return rc;
  if ((rc = make_map_bf16_4d(&tv128, a->V, AT_D, a->n_k, a->H, a->B, a->ldv, AT_D, a->v_bs, 64, 128, "attn V"))) return rc;
  if ((rc = make_map_bf16_4d(&tq64, a->Q, AT_D, a->n_q, a->H, a->B, a->ldq, AT_D, a->q_bs, 64, 64, "attn Q64"))) return rc;
  if ((rc = make_map_bf16_4d(&tdo64, a->dO, AT_D, a->n_q, a->H, a->B, a->lddo, AT_D, a->do_bs, 64, 64, "attn dO64"))) return rc;
  AttnP pq = p;
  pq.out0 = (bf16*)a->dQ; pq.ld0 = a->lddq; pq.bs0 = a->dq_bs;
  (void)launch_pdl(attn_bwd_dq3_kernel, dim3((a->n_q + 127) / 128, a->H, a->B), dim3(A3_THREADS), (size_t)DQ3_SMEM, st, tq128,
                   tdo128, to128, tk64, tv64, pq);
  if ((rc = check_launch("b2_attn_bwd dq3"))) return rc;
  AttnP pk = p;
  pk.out0 = (bf16*)a->dK; pk.ld0 = a->lddk; pk.bs0 = a->dk_bs;
  pk.out1 = (bf16*)a->dV; pk.ld1 = a->lddv; pk.bs1 = a->dv_bs;
  (void)launch_pdl(attn_bwd_dkv3_kernel, dim3((a->n_k + 127) / 128, a->H, a->B), dim3(A3_THREADS), (size_t)Q3_SMEM, st, tk128,
                   tv128, tq64, tdo64, pk);
  return check_launch("b2_attn_bwd dkv3");
}
