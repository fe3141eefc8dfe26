// b2_gemm fast path: persistent CTA-pair tcgen05 GEMM for sm_100a.
//
//   D[m,n] = alpha * sum_k A[m,k] * B[n,k] (+bias) (+residual | +D_old)          bf16 in / bf16 out, fp32 accumulate
//
// Why this shape (measured on B200, profiles/r1_gemm_notes.md): a 128x128 single-CTA tile reads 32 KB of smem per
// 256 MMA cycles and re-fetches every operand tile from L2 once per 128 output rows/cols; it topped out at 27 % of
// the bf16 peak on the K=1280 shapes that dominate the SDXL step.  Here a cluster of two CTAs (one TPC) computes a
// 256 x BN tile with cta_group::2 UMMA (M = 256): each CTA stages only its 128 rows of A and its BN/2 rows of B,
// so smem traffic per flop halves and every operand byte fetched from L2 feeds twice the math.
//
//   * persistent: grid = one cluster per SM pair, tiles assigned round-robin; BN (multiple of 16, <= 256) is chosen
//     on the host so that the tile count fills whole rounds of the 74 clusters (wave quantisation).
//   * warp 0 (both CTAs)  : TMA producer, 6-stage ring; all transaction bytes are signalled on the LEADER's barrier.
//   * warp 1 (leader CTA) : single-thread tcgen05.mma.cta_group::2 issuer; tcgen05.commit multicasts the
//     "stage free" / "accumulator ready" arrivals to both CTAs.  Warp 1 of both CTAs owns the TMEM allocation.
//   * warps 2..5          : epilogue.  TMEM accumulators are double-buffered (2 x 256 columns) so the epilogue of
//     tile i overlaps the mainloop of tile i+1.  tcgen05.ld -> registers -> (+bias, +residual) -> bf16 -> SWIZZLE_128B
//     smem staging -> TMA store (coalesced 128-byte rows, clipped at the M/N edges by the tensor map).  The residual
//     / accumulate operand comes in through a TMA load into the same staging buffer.
#include <stdlib.h>

#include "tc.cuh"

namespace b2 {

constexpr int G2_THREADS = 320;  // warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 and 6-9: two epilogue groups
constexpr int G2_STAGES = 6;
constexpr int G2_BK = 64;
constexpr int G2_A_BYTES = 128 * G2_BK * 2;        // 16 KiB: this CTA's 128 rows of A
constexpr int G2_STAGE_BYTES = 2 * G2_A_BYTES;     // A + up to 128 rows of B
constexpr int G2_EPI_BYTES = 128 * 64 * 2;         // one 128 x 64 bf16 staging tile
constexpr int G2_SMEM = G2_STAGES * G2_STAGE_BYTES + 2 * G2_EPI_BYTES + 1024;
constexpr int G2_WIDE_BN = 320;
constexpr int G2_WIDE_STAGES = 5;
constexpr int G2_WIDE_STAGE_BYTES = G2_A_BYTES + (G2_WIDE_BN / 2) * G2_BK * 2;  // 16 KiB A + 20 KiB B half
static_assert(G2_WIDE_STAGES * G2_WIDE_STAGE_BYTES + 2 * G2_EPI_BYTES + 1024 <= G2_SMEM, "wide mode must fit the same smem budget");
static_assert((G2_STAGES - 1) * G2_STAGE_BYTES + 4 * G2_EPI_BYTES + 1024 <= G2_SMEM, "GEGLU modes: 5 stages + 3 (forward) / 4 (backward) staging tiles");
constexpr int G2_TMEM_COLS = 512;

struct Gemm2P {
  int M, N, K;
  int BN;
  int a_mn, b_mn;
  int tiles_m, tiles_n;
  float alpha;
  const bf16* bias;
  int bias_rows_per_group;
  long long bias_group_stride;
  int has_res;
  bf16* D;
  long long ldd;
  const bf16* R;
  long long ldr;
  // implicit 3x3 convolution (stride 1, pad 1): one operand is gathered from an NHWC activation [B,H,W,C] through a
  // 4-D tensor map {C, W, H, B}; a pixel block is a box {64, bw, bh, 1} shifted by the tap, halo zero-filled by TMA.
  //   conv = 1: A = activation (K-major, k = (tap, c)), 128 pixels per CTA slab      (forward: sign +1, dgrad: sign -1)
  //   conv = 2: B = activation (MN-major, n = (tap, c), k = pixel), 64 pixels per k-block   (wgrad)
  int conv;
  int cv_W, cv_HW;
  int cv_cpb;    // 64-channel blocks per tap of the gathered operand
  int cv_sign;   // +1: input pixel = output pixel + (kh-1, kw-1);  -1: minus (dgrad)
  int cv_btap;   // conv = 1 with MN-major B (dgrad): B column offset per tap (= Cin of the weight)
  int cv_C;      // conv = 2: channels of the gathered activation (n -> tap = n / C, c = n % C)
  // split-K (gradient GEMMs only): work unit = (tile, split); each unit reduces k-blocks [split*kb_per, +kb_per) and
  // its epilogue adds the bf16 partial into D with a TMA reduce (cp.reduce.async.bulk.tensor .add), so units of one
  // tile may run concurrently on different clusters.  D must hold the value to accumulate onto (zeros if none).
  int splits, kb_per;
  // wide mode (BN = 320): one 256 x 320 tile per CTA pair issued as two N = 160 UMMAs per k-step into a single-buffered
  // 320-column accumulator.  Used when it turns "80 tiles on 74 clusters" (N = 1280 at M = 4096: two rounds, the second
  // 8 % full) into 64 tiles in ONE round; each CTA stages its B half as two 80-row boxes so that accumulator columns
  // stay in natural order.
  int wide;
  int stages, stage_bytes;
  // GEGLU mode (b2_linear_geglu): D = u[M, 2F] = x W1^T + b1 with u = [h | g], and Z[M, F] = h * gelu(g) from the same
  // accumulators.  A 256-column tile is 128 h-columns (this pair's CTA 0 stages those rows of W1) next to the MATCHING 128
  // g-columns (CTA 1 stages rows F + ...), so the epilogue holds h_j and g_j of the same feature j in one thread; it
  // rounds both to bf16 exactly as the un-fused path stores them, writes them to u (the backward pass needs both) and
  // writes bf16(h * gelu(g)) to Z — the separate GEGLU kernel's 2F-wide re-read of u disappears.
  int geglu, gg_F;
  // GEGLU-backward mode (b2_linear_dgrad_geglu): the GEMM is the down-projection's input gradient dz[M, F] = dy W2; the
  // epilogue never stores dz — per 64-column chunk it fetches the matching h and g tiles of u = [h | g] by TMA (tmR maps u),
  // forms dh = dz g Phi(g), dg = dz h (Phi(g) + g phi(g)) in place and stores both into du (tmD maps du[M, 2F]): the separate
  // GEGLU-backward kernel's pass over dz, u and du (210 MB per transformer block at 1024 px) shrinks to the u read + du write.
  // Two staging tiles per epilogue group (four in all, mainloop on 5 stages); needs both epilogue groups.  Traced
  // (tools/geglu_trace.py): per chunk ~1.5k cycles store drain + 2.5-3.4k until the u tiles arrive + 4.7k math (issue-bound:
  // two epilogue warps per scheduler) and the mainloop slows to ~18k cycles per tile under the extra shared-memory / TMA
  // traffic; an L2 prefetch of the u boxes did not shorten the arrival (it is TMA-queue / smem-write bound, not HBM latency)
  // and was removed.  75 us against 41 (dgrad) + 53 (gate-backward kernel) at M=4096, F=5120, C=1280.
  int gbwd;
  int res_prefetch;  // 1: residual tiles travel one chunk ahead (default); B2_GEMM_NO_RES_PREFETCH=1 restores the per-chunk load
  int epi_groups;    // 2 (default): both epilogue warp groups work; 1: warps 6-9 idle (B2_GEMM_EPI_GROUPS=1, for A/B runs)
  unsigned long long* dbg;  // optional clock64 trace of cluster 0 / CTA 0 (b2_gemm2_set_debug, tools/geglu_trace.py); NULL in production
};

static unsigned long long* g_g2_dbg = nullptr;
__device__ __forceinline__ unsigned long long g2_clk() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
  return t;
}

// One full 64-column chunk of the plain epilogue: accumulator columns (fp32, in registers) * alpha (+ bias columns staged in
// shared memory) (+ residual tile already in the staging buffer) -> bf16 -> staging buffer (128-byte swizzled rows).
// Specialised on what is present: written with run-time flags inside the column-group loop the same work was 41 branches
// and 16 reconvergence pairs per chunk, and the lone epilogue warp of each scheduler needed 1,400 cycles for ~520
// instructions (clock64 trace, tools/geglu_trace.py).  All shared-memory reads are issued before the first st.shared.
template <bool RES, bool BIAS>
__device__ __forceinline__ void epi_chunk_fast(const uint32_t* v0, const uint32_t* v1, float alpha, uint32_t srow, int row,
                                               const bf16* bias_cols) {
  bf16x8 rv[8], bv[8];
  if (RES) {
#pragma unroll
    for (int g = 0; g < 8; ++g)
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(rv[g].u.x), "=r"(rv[g].u.y), "=r"(rv[g].u.z), "=r"(rv[g].u.w)
                   : "r"(srow + ((g ^ (row & 7)) << 4)));
  }
  if (BIAS) {
#pragma unroll
    for (int g = 0; g < 8; ++g) bv[g].u = *reinterpret_cast<const uint4*>(bias_cols + g * 8);
  }
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const uint32_t* v = g < 4 ? v0 + g * 8 : v1 + (g - 4) * 8;
    float f[8], t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[j]) * alpha;
    if (BIAS) {
      unpack8(bv[g], t);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += t[j];
    }
    if (RES) {
      unpack8(rv[g], t);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += t[j];
    }
    const bf16x8 o = pack8(f);
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((g ^ (row & 7)) << 4)), "r"(o.u.x), "r"(o.u.y),
                 "r"(o.u.z), "r"(o.u.w)
                 : "memory");
  }
}

// 2-D TMA load / store / reduce, cluster address mapping and the epilogue named barrier: tc.cuh (shared with xattn.cu)

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmR,
             const __grid_constant__ CUtensorMap tmZ, const Gemm2P p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[G2_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[G2_STAGES];
  __shared__ __align__(8) uint64_t bar_acc_full[2];
  __shared__ __align__(8) uint64_t bar_acc_empty[2];
  __shared__ __align__(8) uint64_t bar_res[2];
  __shared__ __align__(16) bf16 bias_s[2][G2_WIDE_BN];  // this tile's bias columns (double-buffered by tile parity)
  __shared__ uint32_t tmem_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_epi = smem_base + p.stages * p.stage_bytes;
  const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;
  const uint32_t rank = uniform_u32(cluster_ctarank());
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int num_units = num_tiles * p.splits;
  const int num_kb_total = (p.K + G2_BK - 1) / G2_BK;
  const int halfn = p.BN >> 1;
  const uint32_t stage_tx = G2_A_BYTES + halfn * 128;  // bytes this CTA's two operand tiles occupy

  if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[187] = g2_clk();  // kernel entry
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
    if (p.has_res || p.gbwd) tma_prefetch_desc(&tmR);
#pragma unroll
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_acc_full[b]), 1);
      mbar_init(smem_u32(&bar_acc_empty[b]), 8 * p.epi_groups);  // 4 epilogue warps per group x 2 CTAs
      mbar_init(smem_u32(&bar_res[b]), 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(smem_u32(&tmem_slot), G2_TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(tmem_slot);
  if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[191] = g2_clk();  // set-up done (barriers, TMEM, cluster sync)
  pdl_wait();
  if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[190] = g2_clk();  // predecessor kernel complete               // everything above overlapped the previous kernel's tail; from here on we touch its outputs
  pdl_launch_dependents();  // let the next kernel begin ITS set-up as our CTAs retire

  if (warp == 0) {
    // ===================== TMA producer (both CTAs): warp-uniform loop, elected lane issues =====================
    const bool el = elect_one();
    {
      int stage = 0;
      uint32_t phase = 0;
      const int cv_cend = p.cv_cpb * 64;
      const int nblk = halfn / 64;  // 64-column blocks of an MN-major B half-tile (<= 2)
      for (int u = cluster_id; u < num_units; u += num_clusters) {
        const int t = u % num_tiles, split = u / num_tiles;
        const int kb0 = split * p.kb_per, kb1 = min(num_kb_total, kb0 + p.kb_per);
        const int mb = t % p.tiles_m, nb = t / p.tiles_m;
        const int m_base = mb * 256 + (int)rank * 128;
        // GEGLU mode: CTA 0 stages W1's h rows [nb*128, +128), CTA 1 the matching g rows [F + nb*128, +128)
        const int n_base = p.geglu ? nb * 128 + (int)rank * p.gg_F : nb * p.BN + (int)rank * halfn;
        const int n_wide = nb * p.BN + (int)rank * 80;  // wide mode: rows [n_wide, +80) and [n_wide + 160, +80)
        // Implicit-conv address state.  Everything below is strength-reduced to counters: this loop runs on ONE
        // thread and must issue a k-block's TMAs in well under the k-block's MMA time (256..512 cycles); runtime
        // integer divisions here (~40 dependent instructions each) made the first version producer-bound.
        int cb = 0, ch0 = 0, cw0 = 0;          // conv = 1: this CTA's pixel slab -> (image, row, col)
        int tapc[2] = {0, 0}, tdw[2] = {0, 0}, tdh[2] = {0, 0};  // conv = 2: per 64-column block (channel, dw, dh)
        if (p.conv == 1) {
          cb = m_base / p.cv_HW;
          const int rem = m_base - cb * p.cv_HW;
          ch0 = rem / p.cv_W;
          cw0 = rem - ch0 * p.cv_W;
        } else if (p.conv == 2) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (j >= nblk) break;
            const int n = n_base + 64 * j;
            const int tap = n / p.cv_C;
            tapc[j] = n - tap * p.cv_C;
            const int kh = tap / 3;
            tdh[j] = kh - 1;
            tdw[j] = tap - kh * 3 - 1;
          }
        }
        int c0 = 0, kw = 0, kh = 0, tapoff = 0;  // conv = 1 k-block walk: channel block, tap
        int pw0 = 0, ph0 = 0, pb = 0;            // conv = 2 k-block walk: 64-pixel block -> (col, row, image)
        const int rows_per_kb = p.cv_W >= 64 ? 1 : 64 / (p.cv_W > 0 ? p.cv_W : 64);
        const int cv_H = p.conv ? p.cv_HW / p.cv_W : 0;
        if (kb0 > 0) {  // split-K: start the walk at this unit's first k-block
          if (p.conv == 1) {
            const int tap = kb0 / p.cv_cpb;
            c0 = (kb0 - tap * p.cv_cpb) * 64;
            kh = tap / 3;
            kw = tap - kh * 3;
            tapoff = tap * p.cv_btap;
          } else if (p.conv == 2) {
            const int px = kb0 * 64;
            pb = px / p.cv_HW;
            const int rem = px - pb * p.cv_HW;
            ph0 = rem / p.cv_W;
            pw0 = rem - ph0 * p.cv_W;
          }
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait<true>(smem_u32(&bar_empty[stage]), phase ^ 1u);
          const uint32_t full_local = smem_u32(&bar_full[stage]);
          if (leader && el) mbar_expect_tx(full_local, 2 * stage_tx);
          const uint32_t full = mapa_u32(full_local, 0);
          const uint32_t sa = smem_base + stage * p.stage_bytes;
          const uint32_t sb = sa + G2_A_BYTES;
          const int k0 = kb * G2_BK;
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          if (p.conv == 1) {
            if (el) {
              tma_load_4d_2sm(sa, &tmA, full, c0, cw0 + p.cv_sign * (kw - 1), ch0 + p.cv_sign * (kh - 1), cb);
              if (!p.b_mn) {
                if (p.wide) {
                  tma_load_2d_2sm(sb, &tmB, full, k0, n_wide);
                  tma_load_2d_2sm(sb + 10240, &tmB, full, k0, n_wide + 160);
                } else {
                  tma_load_2d_2sm(sb, &tmB, full, k0, n_base);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 2; ++j)
                  if (j < nblk) tma_load_2d_2sm(sb + j * 8192, &tmB, full, tapoff + n_base + 64 * j, c0);
              }
            }
            c0 += 64;
            if (c0 == cv_cend) {
              c0 = 0;
              tapoff += p.cv_btap;
              if (++kw == 3) { kw = 0; ++kh; }
            }
            continue;
          }
          if (el) {
            if (!p.a_mn) {
              tma_load_2d_2sm(sa, &tmA, full, k0, m_base);
            } else {
              tma_load_2d_2sm(sa, &tmA, full, m_base, k0);
              tma_load_2d_2sm(sa + 8192, &tmA, full, m_base + 64, k0);
            }
          }
          if (p.conv == 2) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
              if (j < nblk && el) tma_load_4d_2sm(sb + j * 8192, &tmB, full, tapc[j], pw0 + tdw[j], ph0 + tdh[j], pb);
            if (p.cv_W >= 64) {
              pw0 += 64;
              if (pw0 == p.cv_W) { pw0 = 0; ++ph0; }
            } else {
              ph0 += rows_per_kb;
            }
            if (ph0 == cv_H) { ph0 = 0; ++pb; }
          } else if (!p.b_mn) {
            if (el) {
              if (p.wide) {
                tma_load_2d_2sm(sb, &tmB, full, k0, n_wide);
                tma_load_2d_2sm(sb + 10240, &tmB, full, k0, n_wide + 160);
              } else {
                tma_load_2d_2sm(sb, &tmB, full, k0, n_base);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j)
              if (j < nblk && el) tma_load_2d_2sm(sb + j * 8192, &tmB, full, n_base + 64 * j, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA): warp-uniform loop, elected lane issues =====================
    const bool el = elect_one();
    if (leader) {
      const uint32_t idesc = umma_idesc(256, p.wide ? 160 : p.BN, p.a_mn, p.b_mn);
      const uint32_t a_lbo = p.a_mn ? 8192 : 16, a_kstep = p.a_mn ? 2048 : 32;
      const uint32_t b_lbo = p.b_mn ? 8192 : 16, b_kstep = p.b_mn ? 2048 : 32;
      // descriptors differ between stages / k-steps only in their 14-bit start-address field: one add each
      const uint64_t adesc0 = umma_desc(smem_base, a_lbo, 1024);
      const uint64_t bdesc0 = umma_desc(smem_base + G2_A_BYTES, b_lbo, 1024);
      const uint32_t a_k16 = a_kstep >> 4, b_k16 = b_kstep >> 4;
      int tile_i = 0, stage = 0;
      uint32_t phase = 0;
      for (int u = cluster_id; u < num_units; u += num_clusters, ++tile_i) {
        const int split = u / num_tiles;
        const int kb0 = split * p.kb_per, kb1 = min(num_kb_total, kb0 + p.kb_per);
        const int buf = tile_i & 1;
        const uint32_t use = (uint32_t)tile_i >> 1;
        const bool tr = p.dbg && cluster_id == 0 && el && tile_i < 8;
        if (tr) p.dbg[128 + tile_i * 4] = g2_clk();
        mbar_wait<true>(smem_u32(&bar_acc_empty[buf]), (use & 1) ^ 1u);
        tc_fence_after();
        if (tr) p.dbg[128 + tile_i * 4 + 1] = g2_clk();
        const uint32_t tacc = tmem_base + buf * 256;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait<true>(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          const uint64_t so = (uint64_t)((uint32_t)(stage * p.stage_bytes) >> 4);
          if (el) {
            if (p.wide) {
#pragma unroll
              for (int k = 0; k < G2_BK / 16; ++k) {
                const uint32_t acc = ((kb - kb0) | k) != 0;
                umma_bf16_2sm(tacc, adesc0 + so + k * a_k16, bdesc0 + so + k * b_k16, idesc, acc);
                umma_bf16_2sm(tacc + 160, adesc0 + so + k * a_k16, bdesc0 + so + k * b_k16 + (10240 >> 4), idesc, acc);
              }
            } else {
#pragma unroll
              for (int k = 0; k < G2_BK / 16; ++k)
                umma_bf16_2sm(tacc, adesc0 + so + k * a_k16, bdesc0 + so + k * b_k16, idesc, ((kb - kb0) | k) != 0);
            }
            umma_commit_2sm(smem_u32(&bar_empty[stage]), 3);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        if (el) umma_commit_2sm(smem_u32(&bar_acc_full[buf]), 3);
        if (tr) p.dbg[128 + tile_i * 4 + 2] = g2_clk();
      }
    }
  } else if (((warp - 2) >> 2) < p.epi_groups) {
    // ===================== epilogue (both CTAs): TMEM -> regs -> smem -> TMA store =====================
    // Two groups of four warps (one warp of each group per scheduler).  Plain mode: the groups take ALTERNATE 64-column
    // chunks of the tile, each with its own staging buffer, named barrier and store-issuing thread, so two chunks are in
    // flight and the lone-warp latencies of one group hide behind the other (a single group needed ~1,250 cycles per chunk,
    // more than half of it barriers / fences / TMEM-load waits: tools/geglu_trace.py).  GEGLU mode: both groups work on
    // the same chunk, 32 of its 64 columns each.
    const int grp = (warp - 2) >> 2;
    const bool two = p.epi_groups == 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;            // TMEM lane == row of this CTA's 128-row slab
    const int et = row;                        // epilogue thread id 0..127
    const uint32_t lane_off = uint32_t(qd * 32) << 16;
    const uint32_t acc_empty_leader = mapa_u32(smem_u32(&bar_acc_empty[0]), 0);
    // Hands accumulator buffer `b` back to the MMA issuer (leader CTA) as soon as its last columns are in registers — the
    // math and the stores of the last chunk no longer hold it.  Relaxed: the tcgen05 fence orders this thread's TMEM reads
    // before the arrive and nothing else is published through this barrier; the `.release.cluster` arrive that used to sit
    // at the end of the tile stalled the issuing thread ~3,000 cycles behind its outstanding stores (tools/geglu_trace.py).
    auto release_acc = [&](int b) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0)
        asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(acc_empty_leader + b * 8) : "memory");
    };
    auto gbar = [&]() {  // this group's 128 threads
      if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");
    };
    auto allbar = [&]() {  // every epilogue thread of the CTA
      if (two) asm volatile("bar.sync 3, 256;" ::: "memory");
      else asm volatile("bar.sync 1, 128;" ::: "memory");
    };
    const int c_first = two ? grp * 64 : 0, c_step = two ? 128 : 64;  // plain mode: this group's chunks of a tile
    int tile_i = 0;
    uint32_t chunk_i = 0;  // chunks this group has staged so far (single group: parity selects the staging buffer)
    uint32_t res_uses[2] = {0, 0};
    const bool reduce_out = p.splits > 1;
    for (int u = cluster_id; u < num_units; u += num_clusters, ++tile_i) {
      const int t = u % num_tiles;
      const int mb = t % p.tiles_m, nb = t / p.tiles_m;
      const int m_base = mb * 256 + (int)rank * 128;
      const int n_tile = nb * p.BN;
      const int ncols = min(p.BN, p.N - n_tile);
      const int buf = tile_i & 1;
      const uint32_t use = (uint32_t)tile_i >> 1;
      const int gm = m_base + row;
      const bool row_ok = gm < p.M;
      const bf16* bias_row =
          p.bias ? p.bias + (long long)((row_ok ? gm : 0) / p.bias_rows_per_group) * p.bias_group_stride : nullptr;
      // The bias columns of the tile go through shared memory: read straight from global inside the chunk loop, each of the
      // 8 (GEGLU: 16) 16-byte loads per chunk exposed its full L1-miss latency to the ONE epilogue warp of its scheduler
      // (clock64 trace, tools/geglu_trace.py: 7,500 cycles per 64-column GEGLU chunk for 1,800 instructions).  Loaded here,
      // before the wait for the accumulator; the chunk loop's first named barrier orders the writes before the reads.
      // Only when the CTA's 128 rows share one bias row (always, except per-sample conv bias with H*W < 128).
      const int tb = tile_i & 1;
      bool bias_smem = false;
      if (p.bias) {
        const int g_lo = m_base / p.bias_rows_per_group, g_hi = min(m_base + 127, p.M - 1) / p.bias_rows_per_group;
        bias_smem = g_lo == g_hi;
        if (bias_smem && grp == 0) {
          const bf16* brow = p.bias + (long long)g_lo * p.bias_group_stride;
          const int c = et * 8;
          if (p.geglu) {
            if (c < 256) *reinterpret_cast<uint4*>(&bias_s[tb][c]) = ld8(brow + (c < 128 ? nb * 128 + c : p.gg_F + nb * 128 + c - 128)).u;
          } else if (c < ncols) {
            *reinterpret_cast<uint4*>(&bias_s[tb][c]) = ld8(brow + n_tile + c).u;
          }
        }
        if (two) allbar();  // group 1 reads what group 0 staged (and has left the previous tile's bias buffer)
      }
      // Residual / accumulate operand: its 16 KB tile per 64-column chunk comes in by TMA.  Issued at the chunk's own start, the
      // load's latency (L2 hit ~0.4 us, HBM ~0.8 us) was exposed once per chunk — five times per 256 x 320 tile, on every
      // "+= gradient" GEMM and every Linear with a fused residual.  Now the first chunk's load leaves BEFORE the wait for the
      // accumulator and each chunk prefetches the next one's tile into the other staging buffer.
      auto chunk_full = [&](int c0) { return (min(64, ncols - c0) == 64) || (n_tile + p.BN >= p.N); };
      const bool first_res_early = p.has_res && p.res_prefetch && !p.geglu && c_first < ncols && chunk_full(c_first);
      if (first_res_early && et == 0) {
        const uint32_t sb = two ? (uint32_t)grp : (chunk_i & 1);
        if (two) tma_store_wait_read<0>();  // this group's previous store has read the buffer
        else tma_store_wait_read<1>();      // the store that last read this staging buffer (two chunks ago) has drained
        mbar_expect_tx(smem_u32(&bar_res[sb]), G2_EPI_BYTES);
        tma_load_2d(smem_epi + sb * G2_EPI_BYTES, &tmR, smem_u32(&bar_res[sb]), n_tile + c_first, m_base);
      }
      const bool tr = p.dbg && cluster_id == 0 && leader && et == 0 && grp == 0 && tile_i < 8;
      if (tr) p.dbg[tile_i * 8] = g2_clk();
      mbar_wait<true>(smem_u32(&bar_acc_full[buf]), use & 1);
      tc_fence_after();
      if (tr) p.dbg[tile_i * 8 + 1] = g2_clk();
      const uint32_t tacc = tmem_base + buf * 256 + lane_off;
      if (!p.geglu && c_first >= ncols) release_acc(buf);  // narrow tile: this group has no chunk in it
      if (p.geglu) {
        // accumulator columns [0,128) = h features nb*128 .., [128,256) = the matching g features.  Three staging tiles
        // (h, g, z) per 64-column chunk: GEGLU mode runs the mainloop on 5 stages, which frees a third 16 KB buffer.
        const int nh = nb * 128;
        for (int c0 = 0; c0 < 128; c0 += 64) {
          if (et == 0 && grp == 0) tma_store_wait_read<0>();  // the previous chunk's three TMA stores have read their tiles
          allbar();
          if (tr) p.dbg[tile_i * 8 + 2 + 3 * (c0 >> 6)] = g2_clk();
          const uint32_t srow = smem_epi + row * 128;
          // 32 columns of h and of g at a time: with the whole 64 + 64-column chunk in registers (128 + temporaries) ptxas
          // had no registers left to interleave the eight independent GELU chains, and the lone epilogue warp of each
          // scheduler ran them at ~3 cycles per instruction (tools/geglu_trace.py); TMEM loads are cheap (~50 cycles).
          for (int hf = two ? grp : 0; hf < (two ? grp + 1 : 2); ++hf) {  // two groups: one 32-column half each
            uint32_t vh[32], vg[32];
            tmem_ld32_nowait(tacc + c0 + hf * 32, vh);
            tmem_ld32_nowait(tacc + 128 + c0 + hf * 32, vg);
            tmem_ld_wait();
            if (c0 == 64 && (two || hf == 1)) release_acc(buf);  // this warp has read all it needs of the accumulator
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              const int g = hf * 4 + g4;
              float fh[8], fg[8], tbv[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) { fh[j] = __uint_as_float(vh[g4 * 8 + j]); fg[j] = __uint_as_float(vg[g4 * 8 + j]); }
              if (p.bias) {  // GEGLU mode has one bias row for all rows: always staged
                bf16x8 bv;
                bv.u = *reinterpret_cast<const uint4*>(&bias_s[tb][c0 + g * 8]);
                unpack8(bv, tbv);
#pragma unroll
                for (int j = 0; j < 8; ++j) fh[j] += tbv[j];
                bv.u = *reinterpret_cast<const uint4*>(&bias_s[tb][128 + c0 + g * 8]);
                unpack8(bv, tbv);
#pragma unroll
                for (int j = 0; j < 8; ++j) fg[j] += tbv[j];
              }
              const bf16x8 ph = pack8(fh), pg = pack8(fg);
              unpack8(ph, fh);   // the GEGLU product is formed from the bf16 values u holds (as geglu_fwd_kernel does)
              unpack8(pg, fg);
              float cdf[8], pdf[8];
              gelu_parts8(fg, cdf, pdf);
#pragma unroll
              for (int j = 0; j < 8; ++j) fh[j] *= fg[j] * cdf[j];  // h * gelu(g), gelu(g) = g * Phi(g) as gelu_erf forms it
              const bf16x8 pz = pack8(fh);
              const uint32_t saddr = srow + ((g ^ (row & 7)) << 4);
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(ph.u.x), "r"(ph.u.y), "r"(ph.u.z),
                           "r"(ph.u.w) : "memory");
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr + G2_EPI_BYTES), "r"(pg.u.x), "r"(pg.u.y),
                           "r"(pg.u.z), "r"(pg.u.w) : "memory");
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr + 2 * G2_EPI_BYTES), "r"(pz.u.x), "r"(pz.u.y),
                           "r"(pz.u.z), "r"(pz.u.w) : "memory");
            }
          }
          if (tr) p.dbg[tile_i * 8 + 3 + 3 * (c0 >> 6)] = g2_clk();
          fence_proxy_async_smem();
          allbar();
          if (et == 0 && grp == 0) {
            tma_store_2d(&tmD, smem_epi, nh + c0, m_base);
            tma_store_2d(&tmD, smem_epi + G2_EPI_BYTES, p.gg_F + nh + c0, m_base);
            tma_store_2d(&tmZ, smem_epi + 2 * G2_EPI_BYTES, nh + c0, m_base);
            tma_store_commit();
          }
          if (tr) p.dbg[tile_i * 8 + 4 + 3 * (c0 >> 6)] = g2_clk();
        }
      } else if (p.gbwd) {
        for (int c0 = c_first; c0 < ncols; c0 += c_step) {  // ncols is a multiple of 64 (F % 128 == 0): full chunks only
          const int n0 = n_tile + c0;
          uint32_t v0[32], v1[32];
          tmem_ld32_nowait(tacc + c0, v0);
          tmem_ld32_nowait(tacc + c0 + 32, v1);
          const uint32_t stage_h = smem_epi + (uint32_t)(grp * 2) * G2_EPI_BYTES, stage_g = stage_h + G2_EPI_BYTES;
          if (et == 0) {
            tma_store_wait_read<0>();  // this group's previous two stores have read the tiles
            mbar_expect_tx(smem_u32(&bar_res[grp]), 2 * G2_EPI_BYTES);
            tma_load_2d(stage_h, &tmR, smem_u32(&bar_res[grp]), n0, m_base);
            tma_load_2d(stage_g, &tmR, smem_u32(&bar_res[grp]), p.gg_F + n0, m_base);
          }
          if (tr && c0 < 256) p.dbg[tile_i * 8 + 2 + 3 * (c0 >> 7)] = g2_clk();
          mbar_wait(smem_u32(&bar_res[grp]), res_uses[grp] & 1);
          res_uses[grp]++;
          tmem_ld_wait();
          if (c0 + c_step >= ncols) release_acc(buf);
          if (tr && c0 < 256) p.dbg[tile_i * 8 + 3 + 3 * (c0 >> 7)] = g2_clk();
          const uint32_t srow = row * 128;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {  // four column groups at a time: their eight shared-memory reads up front
            bf16x8 hv[4], gv[4];
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              const uint32_t so = srow + (((hf * 4 + g4) ^ (row & 7)) << 4);
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(hv[g4].u.x), "=r"(hv[g4].u.y), "=r"(hv[g4].u.z), "=r"(hv[g4].u.w) : "r"(stage_h + so));
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(gv[g4].u.x), "=r"(gv[g4].u.y), "=r"(gv[g4].u.z), "=r"(gv[g4].u.w) : "r"(stage_g + so));
            }
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              const uint32_t* v = (hf ? v1 : v0) + g4 * 8;
              float d[8], fh[8], fg[8], dh[8], dg[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) d[j] = __uint_as_float(v[j]);
              unpack8(pack8(d), d);  // dz as the un-fused path stores it (bf16): same bits downstream
              unpack8(hv[g4], fh);
              unpack8(gv[g4], fg);
              float cdf[8], pdf[8];
              gelu_parts8(fg, cdf, pdf);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                dh[j] = d[j] * fg[j] * cdf[j];
                dg[j] = d[j] * fh[j] * fmaf(fg[j], pdf[j], cdf[j]);
              }
              const bf16x8 oh = pack8(dh), og = pack8(dg);
              const uint32_t so = srow + (((hf * 4 + g4) ^ (row & 7)) << 4);
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(stage_h + so), "r"(oh.u.x), "r"(oh.u.y), "r"(oh.u.z),
                           "r"(oh.u.w) : "memory");
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(stage_g + so), "r"(og.u.x), "r"(og.u.y), "r"(og.u.z),
                           "r"(og.u.w) : "memory");
            }
          }
          if (tr && c0 < 256) p.dbg[tile_i * 8 + 4 + 3 * (c0 >> 7)] = g2_clk();
          fence_proxy_async_smem();
          gbar();
          if (et == 0) {
            tma_store_2d(&tmD, stage_h, n0, m_base);
            tma_store_2d(&tmD, stage_g, p.gg_F + n0, m_base);
            tma_store_commit();
          }
        }
      } else
      for (int c0 = c_first; c0 < ncols; c0 += c_step) {
        const int cw = min(64, ncols - c0);
        const int n0 = n_tile + c0;
        uint32_t v0[32], v1[32];
        tmem_ld32_nowait(tacc + c0, v0);
        if (cw > 32) tmem_ld32_nowait(tacc + c0 + 32, v1);
        const bool full_chunk = (cw == 64) || (n_tile + p.BN >= p.N);  // TMA may write the whole 64-wide box
        if (full_chunk) {
          const uint32_t sbuf = two ? (uint32_t)grp : (chunk_i & 1);
          const uint32_t stage = smem_epi + sbuf * G2_EPI_BYTES;
          if (et == 0) {
            if (p.has_res && (two || !p.res_prefetch)) {
              // one staging buffer per group: the residual tile of this chunk can only be fetched now (the group's first
              // chunk of the tile was fetched before the accumulator wait); the other group's chunk runs meanwhile
              if (!(first_res_early && c0 == c_first)) {
                if (two) tma_store_wait_read<0>(); else tma_store_wait_read<1>();
                mbar_expect_tx(smem_u32(&bar_res[sbuf]), G2_EPI_BYTES);
                tma_load_2d(stage, &tmR, smem_u32(&bar_res[sbuf]), n0, m_base);
              }
            } else if (p.has_res) {
              // this chunk's residual tile is already in flight (tile start / previous chunk); send the next chunk's after
              // the store that last read the OTHER staging buffer (the previous chunk's) has drained
              if (c0 + 64 < ncols && chunk_full(c0 + 64)) {
                tma_store_wait_read<0>();
                mbar_expect_tx(smem_u32(&bar_res[sbuf ^ 1]), G2_EPI_BYTES);
                tma_load_2d(smem_epi + (sbuf ^ 1) * G2_EPI_BYTES, &tmR, smem_u32(&bar_res[sbuf ^ 1]), n0 + 64, m_base);
              }
            } else {
              // the previous TMA store that read this staging buffer must have drained
              if (two) tma_store_wait_read<0>(); else tma_store_wait_read<1>();
            }
          }
          gbar();
          if (p.has_res) {
            mbar_wait(smem_u32(&bar_res[sbuf]), res_uses[sbuf] & 1);
            res_uses[sbuf]++;
          }
          tmem_ld_wait();
          if (c0 + c_step >= ncols) release_acc(buf);
          if (tr && c0 < 128) p.dbg[tile_i * 8 + 2 + 3 * (c0 >> 6)] = g2_clk();
          const uint32_t srow = stage + row * 128;
          if (cw == 64 && (!bias_row || bias_smem)) {
            const bf16* bias_cols = &bias_s[tb][c0];
            if (p.has_res) {
              if (bias_row) epi_chunk_fast<true, true>(v0, v1, p.alpha, srow, row, bias_cols);
              else epi_chunk_fast<true, false>(v0, v1, p.alpha, srow, row, bias_cols);
            } else {
              if (bias_row) epi_chunk_fast<false, true>(v0, v1, p.alpha, srow, row, bias_cols);
              else epi_chunk_fast<false, false>(v0, v1, p.alpha, srow, row, bias_cols);
            }
          } else {
            // narrow last chunk of the matrix, or a per-sample bias whose rows change inside this CTA's slab
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const uint32_t* v = g < 4 ? v0 + g * 8 : v1 + (g - 4) * 8;
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[j]) * p.alpha;
              const bool col_ok = g * 8 < cw;
              if (bias_row && col_ok) {
                float tbv[8];
                unpack8(bias_smem ? ld8(&bias_s[tb][c0 + g * 8]) : ld8(bias_row + n0 + g * 8), tbv);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] += tbv[j];
              }
              const uint32_t saddr = srow + ((g ^ (row & 7)) << 4);
              if (p.has_res) {
                bf16x8 rv;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(rv.u.x), "=r"(rv.u.y), "=r"(rv.u.z), "=r"(rv.u.w)
                             : "r"(saddr));
                float tr8[8];
                unpack8(rv, tr8);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] += tr8[j];
              }
              const bf16x8 o = pack8(f);
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(o.u.x), "r"(o.u.y), "r"(o.u.z),
                           "r"(o.u.w)
                           : "memory");
            }
          }
          if (tr && c0 < 128) p.dbg[tile_i * 8 + 3 + 3 * (c0 >> 6)] = g2_clk();
          fence_proxy_async_smem();
          gbar();
          if (et == 0) {
            if (reduce_out)
              tma_reduce_add_2d(&tmD, stage, n0, m_base);
            else
              tma_store_2d(&tmD, stage, n0, m_base);
            tma_store_commit();
          }
          if (tr && c0 < 128) p.dbg[tile_i * 8 + 4 + 3 * (c0 >> 6)] = g2_clk();
          ++chunk_i;
        } else {
          // tile-bounded partial chunk (BN not a multiple of 64): direct 16-byte stores of the valid columns
          tmem_ld_wait();
          if (c0 + c_step >= ncols) release_acc(buf);
          if (row_ok) {
            bf16* drow = p.D + (long long)gm * p.ldd + n0;
            const bf16* rrow = p.R ? p.R + (long long)gm * p.ldr + n0 : nullptr;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              if (g * 8 < cw) {
                const uint32_t* v = g < 4 ? v0 + g * 8 : v1 + (g - 4) * 8;
                float f[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[j]) * p.alpha;
                if (bias_row) {
                  float tb[8];
                  unpack8(ld8(bias_row + n0 + g * 8), tb);  // partial-chunk path: no named barrier here, read from global
#pragma unroll
                  for (int j = 0; j < 8; ++j) f[j] += tb[j];
                }
                if (rrow) {
                  float tr[8];
                  unpack8(ld8(rrow + g * 8), tr);
#pragma unroll
                  for (int j = 0; j < 8; ++j) f[j] += tr[j];
                }
                st8(drow + g * 8, pack8(f));
              }
            }
          }
        }
      }
    }
    if (p.dbg && cluster_id == 0 && leader && et == 0 && grp == 0) p.dbg[189] = g2_clk();  // last store issued
    if (et == 0) tma_store_wait_all();
    if (p.dbg && cluster_id == 0 && leader && et == 0 && grp == 0) p.dbg[188] = g2_clk();  // all stores complete
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, G2_TMEM_COLS);
  }
}

// make_map_2d(): tc.cuh

// Predicted cycles per k-block of one 256 x bn tile: MMA issue floor vs smem port (TMA write + UMMA read).
static double tile_cost(int bn) {
  const double mma = 2.0 * bn;
  const double smem = (16.0 + bn / 16.0) * 16.0;
  return (mma > smem ? mma : smem) + 16.0;
}

int gemm2_pick_bn(int M, int N, int K, int b_mn, int num_clusters) {
  (void)K;
  const char* env = getenv("B2_GEMM_BN");
  if (env && atoi(env) > 0) return atoi(env);
  const int tiles_m = (M + 255) / 256;
  int best = 256;
  double best_cost = 1e30;
  for (int bn = 256; bn >= 64; bn -= (b_mn ? 128 : 16)) {
    if (bn > 64 && bn - 16 >= N && !b_mn) continue;  // no point in a tile wider than N rounded up
    const int tiles_n = (N + bn - 1) / bn;
    const long long tiles = (long long)tiles_m * tiles_n;
    const long long rounds = (tiles + num_clusters - 1) / num_clusters;
    const double cost = rounds * tile_cost(bn);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

// 256 x 320 single-round tiles: K-major B only, N a multiple of 320, every tile gets its own cluster, and the best
// <= 256-wide plan needs at least two rounds.
static bool gemm2_use_wide(int M, int N, int b_mn, int bn_narrow, int num_clusters) {
  static const bool off = getenv("B2_GEMM_NOWIDE") != nullptr;
  if (off || b_mn || N % G2_WIDE_BN) return false;
  const long long tiles_m = (M + 255) / 256;
  if (tiles_m * (N / G2_WIDE_BN) > num_clusters) return false;
  const long long narrow_tiles = tiles_m * ((N + bn_narrow - 1) / bn_narrow);
  return narrow_tiles > num_clusters;
}

struct G2Plan { int bn, splits, kb_per; };

// Tile width + K split for a gradient GEMM.  Model: a unit costs kb_per * tile_cost(bn) + a fixed hand-over bubble;
// units run in rounds of `num_clusters`.  Splitting pays when the tile grid cannot fill the machine (wgrad: small
// output, long reduction) or quantises badly (80 tiles on 74 clusters).
static G2Plan gemm2_plan(int M, int N, int K, int b_mn, int num_clusters, bool may_split, bool needs_zero_fill) {
  G2Plan best{gemm2_pick_bn(M, N, K, b_mn, num_clusters), 1, (K + G2_BK - 1) / G2_BK};
  const char* env = getenv("B2_GEMM_SPLITK");
  if (!may_split || (env && atoi(env) == 0)) return best;
  const int num_kb = (K + G2_BK - 1) / G2_BK;
  const int tiles_m = (M + 255) / 256;
  auto unit_cost = [&](int bn, int splits, int kb_per) {
    const long long units = (long long)tiles_m * ((N + bn - 1) / bn) * splits;
    const long long rounds = (units + num_clusters - 1) / num_clusters;
    return rounds * (kb_per * tile_cost(bn) + 1200.0 + (splits > 1 ? 600.0 : 0.0)) +
           ((splits > 1 && needs_zero_fill) ? 3000.0 : 0.0);
  };
  double best_cost = unit_cost(best.bn, 1, num_kb);
  static const int cand[] = {2, 3, 4, 5, 6, 8, 10, 12, 16};  // every split adds one bf16 rounding of the running sum
  for (int bn = 256; bn >= 128; bn -= 64) {
    if (b_mn && (bn & 127)) continue;
    if (bn > 128 && bn - 64 >= N) continue;
    for (int s : cand) {
      const int kb_per = (num_kb + s - 1) / s;
      if (kb_per < 8) break;
      const int splits = (num_kb + kb_per - 1) / kb_per;
      const double c = unit_cost(bn, splits, kb_per);
      if (c < best_cost * 0.93) {  // demand a real gain before giving up the single-rounding epilogue
        best_cost = c;
        best = G2Plan{bn, splits, kb_per};
      }
    }
  }
  if (env && atoi(env) > 1) {
    const int s = atoi(env);
    const int kb_per = (num_kb + s - 1) / s;
    best = G2Plan{b_mn ? 256 : 256, (num_kb + kb_per - 1) / kb_per, kb_per};
  }
  return best;
}

// split-K accumulates into D with reduce-adds: when the caller did not ask for `+=`, D starts from zero
static int gemm2_zero_fill(void* D, long long ldd, int M, int N, cudaStream_t st) {
  cudaError_t e = cudaMemset2DAsync(D, (size_t)ldd * 2, 0, (size_t)N * 2, (size_t)M, st);
  if (e != cudaSuccess) {
    set_error("b2_gemm split-K zero fill: %s", cudaGetErrorString(e));
    return B2_ERR_CUDA;
  }
  return B2_OK;
}

static int gemm2_launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const CUtensorMap& tr,
                        Gemm2P& p, cudaStream_t st, const char* what, const CUtensorMap* tz = nullptr) {
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM);
    if (err != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute(gemm2): %s", what, cudaGetErrorString(err));
      return B2_ERR_CUDA;
    }
    configured = true;
  }
  const int num_clusters = num_sms() / 2;
  p.tiles_m = (p.M + 255) / 256;
  p.tiles_n = p.geglu ? p.gg_F / 128 : (p.N + p.BN - 1) / p.BN;
  if (p.splits < 1) {
    p.splits = 1;
    p.kb_per = (p.K + G2_BK - 1) / G2_BK;
  }
  static const bool no_res_prefetch = getenv("B2_GEMM_NO_RES_PREFETCH") != nullptr;
  p.res_prefetch = no_res_prefetch ? 0 : 1;
  p.wide = p.BN == G2_WIDE_BN ? 1 : 0;
  p.stages = p.wide ? G2_WIDE_STAGES : (p.geglu || p.gbwd) ? G2_STAGES - 1 : G2_STAGES;  // GEGLU modes: 5 stages + extra staging tiles
  p.stage_bytes = p.wide ? G2_WIDE_STAGE_BYTES : G2_STAGE_BYTES;
  if (p.splits > 1) p.has_res = 0;  // the reduce-add epilogue is the accumulation
  p.dbg = g_g2_dbg;
  static const int epi_groups = (getenv("B2_GEMM_EPI_GROUPS") && atoi(getenv("B2_GEMM_EPI_GROUPS")) == 1) ? 1 : 2;
  p.epi_groups = epi_groups;
  static const bool log_calls = getenv("B2_GEMM_LOG") != nullptr;
  if (log_calls)
    fprintf(stderr, "B2GEMM %s M=%d N=%d K=%d a_mn=%d b_mn=%d conv=%d BN=%d splits=%d bias=%d res=%d\n", what, p.M, p.N, p.K,
            p.a_mn, p.b_mn, p.conv, p.BN, p.splits, p.bias != nullptr, p.has_res);
  const long long tiles = (long long)p.tiles_m * p.tiles_n * p.splits;
  const int clusters = (int)(tiles < num_clusters ? tiles : num_clusters);
  cudaError_t le = launch_pdl(gemm2_kernel, dim3(2 * clusters), dim3(G2_THREADS), (size_t)G2_SMEM, st, ta, tb, td, tr,
                              tz ? *tz : td, p);
  if (le != cudaSuccess) {
    set_error("%s: launch: %s", what, cudaGetErrorString(le));
    return B2_ERR_CUDA;
  }
  return check_launch(what);
}

// Returns 1 if the fast path handled the call (rc in *out_rc), 0 if the caller must use the generic kernel.
int gemm2_try(const b2_gemm_args* a, cudaStream_t st, int* out_rc) {
  if (getenv("B2_GEMM_LEGACY")) return 0;
  if (a->nb_lo != 1 || a->nb_hi != 1 || a->out_fp32) return 0;
  if (a->M < 256 || a->N < 64 || (a->N & 7) || (a->ldd & 7) || a->K < 64) return 0;
  if (reinterpret_cast<uintptr_t>(a->D) & 15) return 0;
  if (a->b_mn && a->N < 128) return 0;
  const void* R = a->accumulate ? a->D : a->residual;
  const long long ldr = a->accumulate ? a->ldd : a->ldr;
  if (a->accumulate && a->residual) return 0;
  if (R && ((ldr & 7) || (reinterpret_cast<uintptr_t>(R) & 15))) return 0;
  if (a->bias && ((a->bias_group_stride & 7) || (reinterpret_cast<uintptr_t>(a->bias) & 15))) return 0;

  const int num_clusters = num_sms() / 2;
  const bool may_split = a->allow_split_k && !a->bias && !a->residual && a->alpha == 1.f;
  const G2Plan plan = gemm2_plan(a->M, a->N, a->K, a->b_mn, num_clusters, may_split, !a->accumulate);
  int bn = plan.bn;
  if (plan.splits == 1 && gemm2_use_wide(a->M, a->N, a->b_mn, bn, num_clusters)) bn = G2_WIDE_BN;
  if (bn != G2_WIDE_BN && (bn < 32 || bn > 256 || (bn & 15) || (a->b_mn && (bn & 127)))) {
    set_error("b2_gemm: bad BN %d", bn);
    *out_rc = B2_ERR_ARG;
    return 1;
  }
  CUtensorMap ta, tb, td, tr;
  int rc;
  if (!a->a_mn)
    rc = make_map_2d(&ta, a->A, a->K, a->M, a->lda, 64, 128, "A");
  else
    rc = make_map_2d(&ta, a->A, a->M, a->K, a->lda, 64, 64, "A(mn)");
  if (rc) { *out_rc = rc; return 1; }
  if (!a->b_mn)
    rc = make_map_2d(&tb, a->B, a->K, a->N, a->ldb, 64, bn == G2_WIDE_BN ? 80 : bn / 2, "B");
  else
    rc = make_map_2d(&tb, a->B, a->N, a->K, a->ldb, 64, 64, "B(mn)");
  if (rc) { *out_rc = rc; return 1; }
  rc = make_map_2d(&td, a->D, a->N, a->M, a->ldd, 64, 128, "D");
  if (rc) { *out_rc = rc; return 1; }
  if (R) {
    rc = make_map_2d(&tr, R, a->N, a->M, ldr, 64, 128, "R");
    if (rc) { *out_rc = rc; return 1; }
  } else {
    tr = td;
  }
  Gemm2P p{};
  p.M = a->M; p.N = a->N; p.K = a->K; p.BN = bn;
  p.a_mn = a->a_mn ? 1 : 0; p.b_mn = a->b_mn ? 1 : 0;
  p.alpha = a->alpha;
  p.bias = reinterpret_cast<const bf16*>(a->bias);
  p.bias_rows_per_group = a->bias_rows_per_group > 0 ? a->bias_rows_per_group : 0x7fffffff;
  p.bias_group_stride = a->bias_group_stride;
  p.has_res = R ? 1 : 0;
  p.D = reinterpret_cast<bf16*>(a->D); p.ldd = a->ldd;
  p.R = reinterpret_cast<const bf16*>(R); p.ldr = ldr;
  p.splits = plan.splits; p.kb_per = plan.kb_per;
  if (plan.splits > 1 && !a->accumulate) {
    if ((rc = gemm2_zero_fill(a->D, a->ldd, a->M, a->N, st))) { *out_rc = rc; return 1; }
  }
  *out_rc = gemm2_launch(ta, tb, td, tr, p, st, "b2_gemm(pair)");
  return 1;
}

// bf16 4-D map over an NHWC activation: dims {C, W, H, B}; box {64, bw, bh, 1} = `pix` consecutive pixels of one image
// row-major (bw = min(W, pix), bh = pix / bw).  Halo / out-of-image pixels are zero-filled.
static int make_map_conv(CUtensorMap* tm, const void* ptr, int B, int H, int W, int C, long long ld, int pix,
                         const char* name) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return B2_ERR_TMAP;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld & 7)) {
    set_error("b2_conv3x3: operand %s violates TMA alignment (ptr %p ld %lld)", name, ptr, ld);
    return B2_ERR_ARG;
  }
  const int bw = W < pix ? W : pix;
  const int bh = pix / bw;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)(ld * 2), (cuuint64_t)(ld * 2 * W), (cuuint64_t)(ld * 2 * W * H)};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s) failed: CUresult %d (B %d H %d W %d C %d ld %lld box %d x %d)", name, (int)r, B, H,
              W, C, ld, bw, bh);
    return B2_ERR_TMAP;
  }
  return B2_OK;
}

static bool conv_geom_ok(int B, int H, int W, int pix) {
  const long long HW = (long long)H * W;
  if (HW % pix) return false;
  if (W <= pix) return (pix % W) == 0 && (pix / W) <= 256;
  return (W % pix) == 0;
}

}  // namespace b2

using namespace b2;

/* profiling hook (tools/geglu_trace.py): >= 192 uint64 clock64 stamps of cluster 0 (epilogue thread 0 + MMA thread), NULL disables */
extern "C" int b2_gemm2_set_debug(void* buf) {
  b2::g_g2_dbg = reinterpret_cast<unsigned long long*>(buf);
  return B2_OK;
}

extern "C" int b2_linear_geglu_ok(int M, int F, int K) {
  if (getenv("B2_GEGLU_UNFUSED")) return 0;
  return M >= 128 && F >= 128 && (F % 128) == 0 && K >= 64 && (K % 8) == 0;
}

extern "C" int b2_linear_geglu(const void* x, const void* W1, const void* b1, void* u, void* z, int M, int F, int K,
                               int64_t ldx, int64_t ldw, int64_t ldu, int64_t ldz, void* stream) {
  B2_REQUIRE(x && W1 && u && z, "b2_linear_geglu: null pointer");
  B2_REQUIRE(b2_linear_geglu_ok(M, F, K), "b2_linear_geglu: unsupported shape M=%d F=%d K=%d (use b2_gemm + b2_geglu_fwd)", M,
             F, K);
  B2_REQUIRE(!b1 || !(reinterpret_cast<uintptr_t>(b1) & 15), "b2_linear_geglu: bias must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CUtensorMap ta, tb, td, tz;
  int rc;
  if ((rc = make_map_2d(&ta, x, K, M, ldx, 64, 128, "geglu x"))) return rc;
  if ((rc = make_map_2d(&tb, W1, K, 2ull * F, ldw, 64, 128, "geglu W1"))) return rc;
  if ((rc = make_map_2d(&td, u, 2ull * F, M, ldu, 64, 128, "geglu u"))) return rc;
  if ((rc = make_map_2d(&tz, z, F, M, ldz, 64, 128, "geglu z"))) return rc;
  Gemm2P p{};
  p.M = M; p.N = 2 * F; p.K = K; p.BN = 256;
  p.alpha = 1.f;
  p.bias = reinterpret_cast<const bf16*>(b1);
  p.bias_rows_per_group = 0x7fffffff;
  p.D = reinterpret_cast<bf16*>(u); p.ldd = ldu;
  p.geglu = 1; p.gg_F = F;
  return gemm2_launch(ta, tb, td, td, p, st, "b2_linear_geglu", &tz);
}

extern "C" int b2_linear_dgrad_geglu_ok(int M, int F, int C) {
  if (getenv("B2_GEGLU_BWD_UNFUSED")) return 0;
  if (getenv("B2_GEMM_EPI_GROUPS") && atoi(getenv("B2_GEMM_EPI_GROUPS")) == 1) return 0;  // needs both epilogue groups
  return M >= 256 && F >= 256 && (F % 128) == 0 && C >= 64 && (C % 8) == 0;
}

// du[M, 2F] = GEGLU-backward( dz = dy[M, C] @ W2[C, F], u[M, 2F] ): the down-projection's dgrad GEMM with the gate's backward
// in its epilogue (replaces b2_gemm(dgrad) + b2_geglu_bwd; diffusers GEGLU.forward: hidden_states * gelu(gate)).
extern "C" int b2_linear_dgrad_geglu(const void* dy, const void* W2, const void* u, void* du, int M, int F, int C, int64_t ldy,
                                     int64_t ldw, int64_t ldu, int64_t lddu, void* stream) {
  B2_REQUIRE(dy && W2 && u && du, "b2_linear_dgrad_geglu: null pointer");
  B2_REQUIRE(b2_linear_dgrad_geglu_ok(M, F, C), "b2_linear_dgrad_geglu: unsupported shape M=%d F=%d C=%d (use b2_gemm + b2_geglu_bwd)",
             M, F, C);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CUtensorMap ta, tb, td, tr;
  int rc;
  if ((rc = make_map_2d(&ta, dy, C, M, ldy, 64, 128, "dgrad_geglu dy"))) return rc;
  if ((rc = make_map_2d(&tb, W2, F, C, ldw, 64, 64, "dgrad_geglu W2(mn)"))) return rc;
  if ((rc = make_map_2d(&td, du, 2ull * F, M, lddu, 64, 128, "dgrad_geglu du"))) return rc;
  if ((rc = make_map_2d(&tr, u, 2ull * F, M, ldu, 64, 128, "dgrad_geglu u"))) return rc;
  Gemm2P p{};
  p.M = M; p.N = F; p.K = C; p.BN = 256;
  p.a_mn = 0; p.b_mn = 1;
  p.alpha = 1.f;
  p.bias_rows_per_group = 0x7fffffff;
  p.D = reinterpret_cast<bf16*>(du); p.ldd = lddu;
  p.gbwd = 1; p.gg_F = F;
  return gemm2_launch(ta, tb, td, tr, p, st, "b2_linear_dgrad_geglu");
}

extern "C" int b2_conv3x3_implicit_ok(int B, int H, int W, int Cin, int Cout) {
  if (getenv("B2_CONV_EXPLICIT")) return 0;
  if (B <= 0 || H <= 0 || W <= 0 || Cin < 128 || Cout < 64 || (Cin & 63) || (Cout & 63)) return 0;
  if (!conv_geom_ok(B, H, W, 128) || !conv_geom_ok(B, H, W, 64)) return 0;
  return 1;
}

extern "C" int b2_conv3x3(const b2_conv3x3_args* a, void* stream) {
  B2_REQUIRE(a && a->x && a->w && a->y, "b2_conv3x3: null pointer");
  B2_REQUIRE(a->mode >= 0 && a->mode <= 2, "b2_conv3x3: bad mode %d", a->mode);
  B2_REQUIRE(b2_conv3x3_implicit_ok(a->B, a->H, a->W, a->Cin, a->Cout),
             "b2_conv3x3: shape B=%d H=%d W=%d Cin=%d Cout=%d not supported by the implicit path (use im2col + b2_gemm)",
             a->B, a->H, a->W, a->Cin, a->Cout);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int B = a->B, H = a->H, W = a->W, Cin = a->Cin, Cout = a->Cout;
  const long long pixels = (long long)B * H * W;
  B2_REQUIRE(pixels < (1ll << 31), "b2_conv3x3: too many pixels");
  const long long ldx = a->ldx > 0 ? a->ldx : Cin, ldy = a->ldy > 0 ? a->ldy : Cout, ldw = 9ll * Cin;
  const int num_clusters = num_sms() / 2;
  CUtensorMap ta, tb, td, tr;
  Gemm2P p{};
  p.alpha = 1.f;
  p.bias_rows_per_group = 0x7fffffff;
  p.cv_W = W;
  p.cv_HW = H * W;
  int rc;
  if (a->mode == 0) {  // y[pixels, Cout] = conv(x, w) (+bias | +rowbias[b]) (+residual)
    p.M = (int)pixels; p.N = Cout; p.K = 9 * Cin;
    p.a_mn = 0; p.b_mn = 0; p.conv = 1; p.cv_sign = 1; p.cv_cpb = Cin / 64;
    p.BN = gemm2_pick_bn(p.M, p.N, p.K, 0, num_clusters);
    if (gemm2_use_wide(p.M, p.N, 0, p.BN, num_clusters)) p.BN = G2_WIDE_BN;
    if ((rc = make_map_conv(&ta, a->x, B, H, W, Cin, ldx, 128, "conv x"))) return rc;
    if ((rc = make_map_2d(&tb, a->w, 9ull * Cin, Cout, ldw, 64, p.BN == G2_WIDE_BN ? 80 : p.BN / 2, "conv w"))) return rc;
    if ((rc = make_map_2d(&td, a->y, Cout, pixels, ldy, 64, 128, "conv y"))) return rc;
    if (a->residual) {
      const long long ldr = a->ldr > 0 ? a->ldr : Cout;
      if ((rc = make_map_2d(&tr, a->residual, Cout, pixels, ldr, 64, 128, "conv residual"))) return rc;
      p.has_res = 1; p.R = reinterpret_cast<const bf16*>(a->residual); p.ldr = ldr;
    } else {
      tr = td;
    }
    if (a->bias) {
      B2_REQUIRE(!(reinterpret_cast<uintptr_t>(a->bias) & 15), "b2_conv3x3: bias must be 16-byte aligned");
      p.bias = reinterpret_cast<const bf16*>(a->bias);
      if (a->bias_per_sample) {
        p.bias_rows_per_group = H * W;
        p.bias_group_stride = Cout;
      }
    }
    p.D = reinterpret_cast<bf16*>(a->y); p.ldd = ldy;
  } else if (a->mode == 1) {  // dx[pixels, Cin] (+)= conv_transpose(dy, w)
    B2_REQUIRE(Cin >= 128, "b2_conv3x3 dgrad: Cin < 128");
    p.M = (int)pixels; p.N = Cin; p.K = 9 * Cout;
    p.a_mn = 0; p.b_mn = 1; p.conv = 1; p.cv_sign = -1; p.cv_cpb = Cout / 64; p.cv_btap = Cin;
    {
      const G2Plan plan = gemm2_plan(p.M, p.N, p.K, 1, num_clusters, true, !a->accumulate);
      p.BN = plan.bn; p.splits = plan.splits; p.kb_per = plan.kb_per;
      if (plan.splits > 1 && !a->accumulate && (rc = gemm2_zero_fill(a->x, ldx, p.M, p.N, st))) return rc;
    }
    if ((rc = make_map_conv(&ta, a->y, B, H, W, Cout, ldy, 128, "conv dy"))) return rc;
    if ((rc = make_map_2d(&tb, a->w, 9ull * Cin, Cout, ldw, 64, 64, "conv w(mn)"))) return rc;
    if ((rc = make_map_2d(&td, a->x, Cin, pixels, ldx, 64, 128, "conv dx"))) return rc;
    tr = td;
    if (a->accumulate) { p.has_res = 1; p.R = reinterpret_cast<const bf16*>(a->x); p.ldr = ldx; }
    p.D = reinterpret_cast<bf16*>(a->x); p.ldd = ldx;
  } else {  // dw[Cout, 9*Cin] (+)= dy^T (*) x
    p.M = Cout; p.N = 9 * Cin; p.K = (int)pixels;
    p.a_mn = 1; p.b_mn = 1; p.conv = 2; p.cv_sign = 1; p.cv_C = Cin;
    {
      const G2Plan plan = gemm2_plan(p.M, p.N, p.K, 1, num_clusters, true, !a->accumulate);
      p.BN = plan.bn; p.splits = plan.splits; p.kb_per = plan.kb_per;
      if (plan.splits > 1 && !a->accumulate && (rc = gemm2_zero_fill(a->w, ldw, p.M, p.N, st))) return rc;
    }
    if ((rc = make_map_2d(&ta, a->y, Cout, pixels, ldy, 64, 64, "conv dy(mn)"))) return rc;
    if ((rc = make_map_conv(&tb, a->x, B, H, W, Cin, ldx, 64, "conv x(mn)"))) return rc;
    if ((rc = make_map_2d(&td, a->w, 9ull * Cin, Cout, ldw, 64, 128, "conv dw"))) return rc;
    tr = td;
    if (a->accumulate) { p.has_res = 1; p.R = reinterpret_cast<const bf16*>(a->w); p.ldr = ldw; }
    p.D = reinterpret_cast<bf16*>(a->w); p.ldd = ldw;
  }
  return gemm2_launch(ta, tb, td, tr, p, st, "b2_conv3x3");
}
