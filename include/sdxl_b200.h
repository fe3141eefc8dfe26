/*
 * sdxl_b200.h — C ABI of libsdxl_b200.so: hand-written sm_100a kernels for the SDXL UNet training step.
 *
 * Drop-in boundary.  The reference has no native code and no FFI: its hot path is
 * `self.model.unet(...)` + `loss.backward()` (src/training/trainers/methods/ddpm_trainer.py:320-325,:271;
 * flow_matching_trainer.py:400-405,:252), which dispatch through diffusers -> torch -> cuDNN/cuBLAS/SDPA.
 * Each entry below replaces the library kernel(s) that path reaches for one operator; the "replaces" note
 * on every group names the reference call site / diffusers module it stands in for.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; all pointers are DEVICE pointers unless named host_*.
 *   - every launch entry returns 0 on success, a negative code on failure; b2_last_error() (thread-local)
 *     describes the last failure.  No allocation inside, no ownership transfer, no hidden syncs; all work
 *     is enqueued on `stream` (a cudaStream_t passed as void*), so calls are CUDA-graph capturable.
 *   - activations are token-major / NHWC: [B, H*W, C] bf16.  "ld*" / strides are in ELEMENTS.
 *   - bf16 = __nv_bfloat16 (uint16_t storage).
 */
#ifndef SDXL_B200_H
#define SDXL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2_OK 0
#define B2_ERR_ARG (-1)
#define B2_ERR_CUDA (-2)
#define B2_ERR_TMAP (-3)

int b2_version(void);
const char* b2_last_error(void);
/* Number of kernel launches issued through this library since load (for bench.py's gpu_launches). */
long long b2_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM  (tcgen05.mma, TMEM accumulators, TMA-staged operands, 128xBN tiles)
 *   D[z][m,n] = alpha * sum_k A[z][m,k] * B[z][n,k]  (+ bias) (+ residual) (+ D_old if accumulate)
 * replaces: every torch.nn.Linear / 1x1 conv / (with b2_im2col3x3) 3x3 conv of the diffusers UNet, its
 *   autograd dgrad/wgrad GEMMs, and the QK^T / PV batched matmuls of attention (cuBLASLt / cuDNN today).
 *   a_mn / b_mn = 0: operand is K-major (k contiguous, ld = row stride);
 *   a_mn / b_mn = 1: operand is MN-major (m or n contiguous, ld = stride between consecutive k).
 *   Linear fwd: A=x[M,K] (K-major), B=W[N,K] (K-major).  dgrad: A=dY (K-major), B=W as [K_in rows... ] MN-major.
 *   wgrad: A=dY^T (MN-major), B=x^T (MN-major).
 *   batch index z = z_hi * nb_lo + z_lo with independent strides (attention: lo = head, hi = sample).
 *   bias: bf16, indexed bias[(m / bias_rows_per_group) * bias_group_stride + n]  (plain bias: group_stride 0;
 *         per-sample time-embedding row: rows_per_group = H*W, group_stride = N).
 *   TMA constraints (checked): A/B base 16-byte aligned, lda/ldb/batch strides multiples of 8 elements.
 * ------------------------------------------------------------------------------------------------ */
typedef struct b2_gemm_args {
  const void* A;
  const void* B;
  void* D;
  const void* bias;      /* bf16 or NULL */
  const void* residual;  /* bf16, same indexing as D with ldr / r_bs_*; or NULL */
  int32_t M, N, K;
  int32_t nb_lo, nb_hi;  /* batch counts (>=1) */
  int32_t a_mn, b_mn;
  int64_t lda, ldb, ldd, ldr;
  int64_t a_bs_lo, a_bs_hi, b_bs_lo, b_bs_hi, d_bs_lo, d_bs_hi, r_bs_lo, r_bs_hi;
  float alpha;
  int32_t accumulate;    /* 1: D += result (gradient accumulation into bf16/fp32 D) */
  int32_t out_fp32;      /* 1: D is float, else bf16 */
  int32_t bias_rows_per_group;
  int64_t bias_group_stride;
  int32_t tile_n;        /* 0 = auto, else 64 / 128 / 256 */
  int32_t allow_split_k; /* 1: gradient GEMM — the kernel may split K across CTA pairs and combine the bf16 partials
                            with TMA reduce-adds into D (one extra bf16 rounding per split; D zero-filled first when
                            !accumulate).  0: single-pass, single-rounding epilogue (forward activations). */
} b2_gemm_args;

int b2_gemm(const b2_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 3x3 convolution support (implicit-GEMM operand staging; K ordering = (kh, kw, cin))
 * replaces: torch Conv2d 3x3 (pad 1, stride 1|2) and Upsample2D's interpolate+conv in ResnetBlock2D /
 *   Downsample2D / Upsample2D / conv_in / conv_out, and their cuDNN dgrad paths.
 *   x: [B,H,W,C] bf16 -> col: [B*Ho*Wo, ldc] (ldc >= 9*C, multiple of 8; pad columns written as zero)
 *   stride: 1 or 2.  upsample: 1 => the conv sees nearest-2x of x (Ho=2H, Wo=2W) without materialising it.
 *   col2im is the exact adjoint (gather form): dx[B,H,W,C] (+)= sum over taps of dcol.
 * ------------------------------------------------------------------------------------------------ */
int b2_im2col3x3(const void* x, void* col, int B, int H, int W, int C, int stride, int upsample,
                 int64_t ldc, void* stream);
int b2_col2im3x3(const void* dcol, void* dx, int B, int H, int W, int C, int stride, int upsample,
                 int64_t ldc, int accumulate, void* stream);

/* Implicit-GEMM 3x3 convolution (stride 1, pad 1) — no im2col buffer in HBM.  The activation operand is gathered by
 * TMA straight from the NHWC tensor: a 4-D tensor map {C, W, H, B} with a {64, bw, bh, 1} box per pixel block,
 * shifted by the filter tap; the zero padding is TMA's out-of-bounds fill.  Same persistent CTA-pair tcgen05 kernel
 * as b2_gemm (gemm2_kernel), same epilogues.
 * replaces: ResnetBlock2D.conv1 / conv2 (torch Conv2d -> cuDNN fprop / dgrad / wgrad), 34 of the UNet's 40 3x3 convs.
 *   mode 0 (fwd)  : y[B*H*W, Cout] = conv(x[B,H,W,Cin], w[Cout,3,3,Cin]) + bias (+ residual)
 *                   bias_per_sample = 1: bias is [B, Cout] (time-embedding row of each sample), else [Cout]
 *   mode 1 (dgrad): x  (+)= conv_transpose(y = dL/dy, w)          (accumulate selects +=)
 *   mode 2 (wgrad): w  (+)= sum over pixels of y^T (x) shifted x    (w is the [Cout, 9*Cin] gradient buffer)
 * Supported when b2_conv3x3_implicit_ok() returns 1: Cin, Cout multiples of 64 (Cin >= 128), H*W a multiple of 128,
 * W a divisor or a multiple of 128 and of 64.  Other shapes (stride 2, folded upsample, conv_in / conv_out) go through
 * b2_im2col3x3 + b2_gemm. */
typedef struct b2_conv3x3_args {
  void* x;              /* [B*H*W, ldx] bf16 */
  void* w;              /* [Cout, 9*Cin] bf16, K order (kh, kw, cin) */
  void* y;              /* [B*H*W, ldy] bf16 */
  const void* bias;     /* fwd only; bf16 or NULL */
  const void* residual; /* fwd only; [B*H*W, ldr] bf16 or NULL */
  int32_t mode;
  int32_t B, H, W, Cin, Cout;
  int64_t ldx, ldy, ldr; /* 0 = dense */
  int32_t bias_per_sample;
  int32_t accumulate;
} b2_conv3x3_args;
int b2_conv3x3_implicit_ok(int B, int H, int W, int Cin, int Cout);
int b2_conv3x3(const b2_conv3x3_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GroupNorm (+SiLU), NHWC.  replaces torch.nn.GroupNorm + F.silu in ResnetBlock2D / Transformer2DModel.norm /
 *   conv_norm_out and their backward.
 *   stats: ws = double[B*G*2] scratch (zeroed by the call); mean,rstd = float[B*G].
 *   bwd:   dx = d/dx of silu?(gn(x)); dgamma/dbeta accumulated as fp32 into dgb[2*C] (caller zeroes,
 *          then folds into the bf16 grads with b2_accum_f32_to_bf16).
 * ------------------------------------------------------------------------------------------------ */
int b2_gn_stats(const void* x, int B, int HW, int C, int G, float eps, void* ws, float* mean, float* rstd,
                void* stream);
int b2_gn_apply(const void* x, void* y, const float* mean, const float* rstd, const void* gamma,
                const void* beta, int B, int HW, int C, int G, int silu, void* stream);
int b2_gn_bwd(const void* x, const void* dy, void* dx, const float* mean, const float* rstd,
              const void* gamma, const void* beta, int B, int HW, int C, int G, int silu,
              void* ws /* double[B*G*2] */, float* dgb /* float[2*C], accumulated */, int accumulate_dx,
              void* stream);

/* LayerNorm over the last dim. replaces torch.nn.LayerNorm(C, eps=1e-5) x210 and backward. */
int b2_ln_fwd(const void* x, void* y, const void* gamma, const void* beta, float* mean, float* rstd,
              int M, int C, float eps, void* stream);
int b2_ln_bwd(const void* x, const void* dy, void* dx, const void* gamma, const float* mean,
              const float* rstd, float* dgb /* float[2*C], accumulated */, int M, int C, int accumulate_dx,
              void* stream);
/* The same in two independently launchable halves (parts: 1 = dx, 2 = dgamma / dbeta accumulation, 3 = both): the host runs
 * them on two streams (they read the same x / dy). */
int b2_ln_bwd_parts(const void* x, const void* dy, void* dx, const void* gamma, const float* mean, const float* rstd,
                    float* dgb, int M, int C, int accumulate_dx, int parts, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Attention softmax over materialised logits (round-1 path; the fused flash kernel supersedes it).
 * replaces the softmax inside F.scaled_dot_product_attention (attention_processor.AttnProcessor2_0).
 *   S: float [rows, lds] (already scaled); P: bf16 [rows, ldp]; columns >= n_valid of P are written 0.
 *   bwd: dS = P * (dP - rowsum(dP*P)) * scale ; dP float [rows, lds], dS bf16 [rows, ldp].
 * ------------------------------------------------------------------------------------------------ */
int b2_softmax_fwd(const float* S, void* P, int64_t rows, int n_valid, int64_t lds, int64_t ldp, void* stream);
int b2_softmax_bwd(const void* P, const float* dP, void* dS, int64_t rows, int n_valid, int64_t lds,
                   int64_t ldp, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused attention (flash style, head dim 64): O = softmax(Q K^T * scale) V with no n x n tensor in HBM.
 * replaces F.scaled_dot_product_attention in diffusers' AttnProcessor2_0 (BasicTransformerBlock.attn1 / attn2) and
 *   its autograd backward.  tcgen05: S = Q K^T (SS MMA) -> TMEM; exp2 / rescale in registers; bf16 P written back to
 *   TMEM and used as the A operand of the PV MMA (TS MMA); O accumulates in TMEM.
 *   Q/K/V/O (and dO, dQ, dK, dV): bf16, logical [B, n, H, 64]; element (b, i, h, d) at base + b*bs + i*ld + h*64 + d,
 *   so the fused QKV projection output [B*n, 3C] is consumed in place (ld = 3C, bs = n*3C, K = Q + C, V = Q + 2C).
 *   TMA constraints (checked): bases 16-byte aligned, ld / bs multiples of 8 elements.
 *   LSE: float [B, H, n_pad], n_pad = b2_attn_lse_rows(n_q) (= n_q rounded up to 128), log2 domain, written by fwd,
 *   read by bwd.  D: float [B, H, n_pad] scratch written by bwd.  flags: reserved (0).
 *   n_k <= 96 (cross-attention, 77 text tokens) takes a query-tile-persistent forward kernel.
 * ------------------------------------------------------------------------------------------------ */
typedef struct b2_attn_args {
  const void* Q;
  const void* K;
  const void* V;
  void* O;        /* fwd: output; bwd: input */
  float* LSE;
  const void* dO; /* bwd only from here */
  float* D;
  void* dQ;
  void* dK;
  void* dV;
  int32_t B, H, n_q, n_k;
  int64_t ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
  int64_t q_bs, k_bs, v_bs, o_bs, do_bs, dq_bs, dk_bs, dv_bs;
  float scale;
  int32_t flags;
} b2_attn_args;

int b2_attn_lse_rows(int n_q);
int b2_attn_fwd(const b2_attn_args* args, void* stream);
int b2_attn_bwd(const b2_attn_args* args, void* stream);
/* Cross-attention backward (n_k <= 80 keys) in one pass: dQ, dK, dV from a single recomputation of P; b2_attn_bwd routes to it
 * when b2_xattn_bwd_ok().  Same argument struct and semantics as b2_attn_bwd (D is not used).  B2_XATTN_BWD_GENERIC=1 forces the
 * generic two-kernel path. */
int b2_xattn_bwd_ok(int B, int H, int n_q, int n_k);
int b2_xattn_bwd(const b2_attn_args* args, void* stream);
/* Profiling hook: device buffer of 32 uint64 per-phase cycle counters accumulated by the forward kernel (NULL = off). */
int b2_attn_set_debug(void* counters);

/* Cross-attention query projection fused with the 77-key attention core (xattn.cu; the north-star kernel of SURVEY.md 8d):
 *   Q[M, C] = xn[M, C] Wq[C, C]^T          (M = B * n_q; written once for the backward pass, never read back here)
 *   O[:, h*64:(h+1)*64] = softmax(Q_h K_h^T * scale) V_h   per head, keys = the sample's n_k text tokens
 *   LSE[B, H, n_pad] (log2 domain, as b2_attn_fwd)
 * replaces: diffusers Attention.to_q (Linear) + F.scaled_dot_product_attention of attn2 in BasicTransformerBlock (call site
 * of the UNet forward: src/training/trainers/methods/ddpm_trainer.py:320-325) — two launches and the HBM round trip of Q.
 * K / V: [B * n_k, ld] row-major views (head h = columns h*64..), sample stride k_bs / v_bs elements.
 * b2_xattn_q_core_ok(): C a multiple of 320, n_q a multiple of 256, n_k <= 80; other shapes: b2_gemm + b2_attn_fwd. */
int b2_xattn_q_core_ok(int B, int n_q, int n_k, int C);
/* Down-projection input gradient with the GEGLU backward in its epilogue: du[M, 2F] from dy[M, C], W2[C, F] (row-major Linear
 * weight [out = C, in = F]) and u[M, 2F] = [h | g]; dz is never stored.  Bit-identical to b2_gemm(dgrad) + b2_geglu_bwd.
 * replaces: autograd of diffusers FeedForward.net[2] (Linear) + GEGLU.forward's `hidden_states * self.gelu(gate)`. */
int b2_linear_dgrad_geglu_ok(int M, int F, int C);
int b2_linear_dgrad_geglu(const void* dy, const void* W2, const void* u, void* du, int M, int F, int C, int64_t ldy,
                          int64_t ldw, int64_t ldu, int64_t lddu, void* stream);
int b2_gemm2_set_debug(void* buf);  /* clock64 trace buffer (>= 192 uint64) for tools/geglu_trace.py; NULL disables */
int b2_xattn_set_debug(void* stamps); /* profiling hook: >= 32 uint64 clock64 stamps of CTA 0's first tile; NULL = off */
int b2_xattn_q_core(const void* xn, const void* Wq, const void* K, const void* V, void* Q, void* O, float* LSE, int B, int n_q,
                    int n_k, int C, int64_t ldx, int64_t ldw, int64_t ldq, int64_t ldo, int64_t ldk, int64_t ldv, int64_t k_bs,
                    int64_t v_bs, float scale, void* stream);

/* GEGLU: z[m, j] = u[m, j] * gelu_erf(u[m, F + j]), u: [M, 2F]. replaces diffusers GEGLU.forward + backward. */
int b2_geglu_fwd(const void* u, void* z, int64_t M, int F, void* stream);
/* GEGLU up-projection with the gate in the GEMM epilogue (gemm2_kernel, GEGLU mode):
 *   u[M, 2F] = x[M, K] W1[2F, K]^T + b1   (both halves rounded to bf16 and stored: the backward pass reads them)
 *   z[M, F]  = u[:, :F] * gelu_erf(u[:, F:])   formed from the bf16 values of u, i.e. bit-identical to b2_gemm + b2_geglu_fwd
 * replaces: diffusers GEGLU.forward (attention.py / activations.py: proj -> chunk(2) -> hidden * F.gelu(gate)), 70 per UNet
 * forward; saves the separate kernel's re-read of u (84 MB per block at C = 1280, B = 4, 1024 px).
 * b2_linear_geglu_ok(): M >= 128, F a multiple of 128, K a multiple of 8 and >= 64. */
int b2_linear_geglu_ok(int M, int F, int K);
int b2_linear_geglu(const void* x, const void* W1, const void* b1, void* u, void* z, int M, int F, int K, int64_t ldx,
                    int64_t ldw, int64_t ldu, int64_t ldz, void* stream);
int b2_geglu_bwd(const void* u, const void* dz, void* du, int64_t M, int F, void* stream);
/* same + the ff1 bias gradient: db32[c] += sum_m du[m, c] (fp32, 2F entries; replaces b2_colsum_f32 over du) */
int b2_geglu_bwd_bias(const void* u, const void* dz, void* du, float* db32, int64_t M, int F, void* stream);

/* Elementwise / plumbing */
int b2_silu_fwd(const void* x, void* y, int64_t n, void* stream);
int b2_silu_bwd(const void* x, const void* dy, void* dx, int64_t n, int accumulate, void* stream);
int b2_add(const void* a, const void* b, void* out, int64_t n, void* stream);              /* out = a + b */
int b2_copy2d(const void* src, void* dst, int64_t rows, int64_t cols, int64_t lds, int64_t ldd,
              int accumulate, void* stream);                                                  /* bf16 */
int b2_copy2d_any(const void* src, void* dst, int64_t rows, int64_t cols, int64_t lds, int64_t ldd,
                  int accumulate, void* stream);                   /* bf16, no alignment requirements (scalar) */
int b2_colsum(const void* dy, void* db, int64_t M, int N, int64_t ld, int accumulate, float* ws /* float[N] */,
              void* stream);                                                                  /* db[n] (+)= sum_m dy[m,n] */
int b2_accum_f32_to_bf16(const float* src, void* dst, int64_t n, int accumulate, void* stream);
/* Small-parameter gradients (biases, norm scale / shift; replaces the autograd bias / affine gradient reductions):
 *   b2_colsum_f32: out[n] += sum_m dy[m,n] with fp32 atomics into a staging buffer (no zero-fill, no conversion);
 *   b2_flush_small_grads: for every segment (staging offset, gradient offset, length; int64 triples on the device)
 *   grad_bf16[dst + i] += staging[src + i], then staging is cleared — one launch per backward pass. */
int b2_colsum_f32(const void* dy, float* out, int64_t M, int N, int64_t ld, void* stream);
/* out[g, n] (+)= sum over the rows of group g (rows [g*rows_per_group, (g+1)*rows_per_group)) of dy[m, n]; out bf16
 * [groups, N] dense.  The per-sample gradient of ResnetBlock2D's time-embedding row (conv1 epilogue bias). */
int b2_colsum_groups(const void* dy, void* out, int groups, int64_t rows_per_group, int N, int64_t ld, int accumulate,
                     float* ws /* float[groups*N] */, void* stream);
int b2_flush_small_grads(float* staging, void* grad_bf16, const int64_t* segments, int nseg, void* stream);
int b2_nchw_to_nhwc(const void* x, int x_fp32, void* y, int B, int C, int HW, int Cpad, void* stream);
int b2_nhwc_to_nchw(const void* x, void* y, int y_fp32, int B, int C, int HW, int Cpad, void* stream);

/* Timestep sinusoid (diffusers get_timestep_embedding, flip_sin_to_cos=True, freq_shift=0):
 *   out[i, :] = [cos(t_i f_k), sin(t_i f_k)], f_k = exp(-ln(1e4) k / half); t float[n]; out bf16 [n, ldo]. */
int b2_timestep_embedding(const float* t, void* out, int n, int dim, int64_t ldo, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Noise / noising / loss   (reference: ddpm_trainer.py:303-345,380-384; flow_matching_trainer.py:298-335;
 *   novelai_v3.py:111-127).  Philox4x32-10 counter RNG, Box-Muller; `seed`,`offset` read from device memory
 *   (uint64[2]) so that a captured CUDA graph draws fresh noise per replay.
 *   b2_randn:  out fp32[n] ~ N(0,1), optionally rounded through bf16 (the reference draws in model dtype).
 *   b2_make_noisy: mode 0 (ddpm): noisy = clamp?(x + sigma[t_b] * eps) ; target = eps | (eps - x)/sigma
 *                  mode 1 (flow): noisy = (1-t_b) eps + t_b x          ; target = x - eps
 *       x fp32 NCHW [B, CHW]; writes noisy as bf16 NHWC-padded [B,HW,Cpad] and target fp32 NCHW.
 *   b2_mse_loss: pred bf16 NHWC-padded vs target fp32 NCHW; weight[b] optional;
 *       loss_sum += sum w_b (pred-target)^2 (double); dpred (bf16 NHWC-padded) = 2 w_b (pred-target) * gscale.
 *   b2_finalize_loss: loss = sum/count * scale; non-finite or > 1000 -> 1000, ok=0 and dpred zeroed (the
 *       reference's clamp / fallback pass no gradient); else loss, ok=1.
 * ------------------------------------------------------------------------------------------------ */
int b2_randn(float* out, int64_t n, const uint64_t* seed_offset, uint64_t stream_id, int round_bf16, void* stream);
int b2_philox_advance(uint64_t* seed_offset, uint64_t inc, void* stream);   /* seed_offset[1] += inc (graph-safe) */
int b2_make_noisy(const float* x, const float* eps, const float* sigma_or_t, int mode, int v_prediction,
                  int clamp_ztsnr, void* noisy, float* target, int B, int C, int HW, int Cpad, void* stream);
int b2_mse_loss(const void* pred, const float* target, const float* weight, double* loss_sum, void* dpred,
                float gscale, int B, int C, int HW, int Cpad, void* stream);
int b2_finalize_loss(const double* loss_sum, double count, float scale, float* loss_out, int32_t* ok,
                     void* dpred /* zeroed when !ok; may be NULL */, int64_t n_dpred, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer side: global grad-norm and fused AdamW over the flat bf16 buffers.
 *   b2_sumsq: out(double) += sum g^2.  b2_adamw: decoupled weight decay, bias-corrected, fp32 m/v,
 *   clip coefficient computed on device from *gnorm_sq (max_norm <= 0: no clipping).
 * ------------------------------------------------------------------------------------------------ */
int b2_sumsq(const void* g, int64_t n, double* out, void* stream);
/* out[0] += sum|x|, out[1] += sum x^2 over elements with (i % period) < valid (period 0: all). Metrics of the
 * reference's step dicts (ddpm_trainer.py:386-396, flow_matching_trainer.py:338-347) without host syncs. */
int b2_abs_sq_sums(const void* x, int is_fp32, int64_t n, int period, int valid, double* out, void* stream);
/* x *= (*dev_scale if given) * host_scale  — applies autograd's upstream scalar (loss / accum) to dL/dpred. */
int b2_scale_bf16(void* x, int64_t n, const float* dev_scale, float host_scale, void* stream);
int b2_adamw(void* p, float* master /* fp32 master weights or NULL */, const void* g, float* m, float* v, int64_t n,
             float lr, float beta1, float beta2, float eps, float weight_decay, int step, const double* gnorm_sq,
             float max_norm, float grad_scale, const uint64_t* dev_step /* used when step <= 0: the step counter is read
             from device memory, so a captured CUDA graph advances it per replay (b2_philox_advance) */, void* stream);

/* AdamWBF16 — the reference's default optimizer (src/config.yaml `optimizer_type: adamw_bf16`;
 * src/training/optimizers/adamw_bfloat16/__init__.py:86-197, stochastic/__init__.py:46-124) as one fused launch over
 * the flat buffers: bf16 p / exp_avg / exp_avg_sq / shift, stochastic rounding of every add (16 counter-hash bits each),
 * global-norm clip folded in (clip coefficient from *gnorm_sq, applied to the bf16 gradient with one rounding like
 * torch.nn.utils.clip_grad_norm_).  18 bytes of HBM traffic per parameter.
 *   as_written = 1 keeps add_stochastic_'s operand order as the reference wrote it (exp_avg <- SR(g + (1-b1) b1 m)).
 *   rng_mode: 0 counter hash of (element, seed_offset[0], step); 1/2/3 deterministic patterns for parity tests (3: int32 [4,n] buffer).
 *   zero_grad = 1: g is overwritten with zeros after it has been read (optimizer.zero_grad() in the same pass).
 *   step <= 0: the step is read from seed_offset[1] on the device (CUDA-graph replays; advance with b2_philox_advance).
 * b2_axpy_bf16: y <- bf16(y + alpha x) — the deferred weight decay `shift.add_(p, alpha=-decay)` (:191-192).
 * b2_adamw_denom_test: test hook — the kernel's denominator bf16(bf16(sqrt v) + eps) as computed (fast) and as IEEE. */
int b2_adamw_bf16(void* p, void* g, void* m, void* v, void* shift, int64_t n, double lr, double beta1, double beta2,
                  double eps, int step, const double* gnorm_sq, float max_norm, float grad_scale,
                  const uint64_t* seed_offset, int as_written, int rng_mode, const int32_t* test_rand16, int zero_grad,
                  void* stream);
int b2_axpy_bf16(void* y, const void* x, int64_t n, float alpha, void* stream);
int b2_adamw_denom_test(const void* v, void* fast, void* ieee, int n, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange over NVSwitch peer memory.
 * replaces: DistributedDataParallel's bucketed NCCL all-reduce of the UNet gradients
 *   (src/core/distributed.py:142-163 `convert_model_to_ddp`; the reducer fires during `loss.backward()`,
 *   ddpm_trainer.py:271).  Sum over ranks of pieces ("chunks") of the flat bf16 gradient buffer, started as soon
 *   as the backward pass has finished writing a chunk and running beside the rest of it WITHOUT occupying SMs:
 *   reduce-scatter and all-gather are copy-engine transfers between IPC-mapped peer buffers, ordering across
 *   processes is by flag words (st.release.sys from a one-warp kernel; cuStreamWaitValue32 on the waiting side),
 *   and the only arithmetic is one shared-memory-free reduce kernel over the owner's shard.  Every rank ends
 *   with bit-identical sums (fp32 accumulation, one rounding).  csrc/dpx.cu.
 *   Three transports, same protocol (b2_dpx_create `mode`): 0 "ce_pull" (reduce-scatter = copy-engine reads from the
 *   peers), 1 "ce_push" (reduce-scatter = copy-engine writes into the owner's staging slots; NVLink carries posted
 *   writes only), 2 "sm" (one shared-memory-free kernel of short 128-thread CTAs that co-reside with the persistent
 *   GEMM / attention CTAs: 16-byte loads of the shard from every peer, fp32 sum, 16-byte stores to every peer; <= 8
 *   ranks, no staging).
 *   b2_dpx_ipc_export / _import: cudaIpc handle (64 bytes) + byte offset of `dev_ptr` inside its allocation.
 *   b2_dpx_alloc_flags: this rank's zeroed flag page (library-owned cudaMalloc; export it with _ipc_export).
 *   b2_dpx_create: grad_ptrs / flag_ptrs / staging_ptrs are [world] device pointers valid IN THIS PROCESS (own entry =
 *       local buffer, the others IPC-mapped; peer staging pointers are used by mode 1 only); every staging buffer is
 *       (world-1) slots of staging_slot_elems bf16.
 *   b2_dpx_exchange: enqueue the exchange of chunk `chunk` (< b2_dpx_max_chunks()) after everything already
 *       enqueued on main_stream.  The chunk is n_ranges pieces [range_off, range_off + range_len) of the buffer
 *       (elements, multiples of 8); each piece is cut into `world` shards of ceil(len / world) rounded up to 8, rank r
 *       reduces shard r.  staging_base: element offset of the chunk's region inside a staging slot (regions of
 *       chunks exchanged in the same step must not overlap).  `seq` must increase by one per optimizer step.
 *   b2_dpx_finish: main_stream waits until all chunks exchanged with `seq` are complete in the local buffer.
 *   b2_dpx_memcpy_async: raw async copy between local / IPC-mapped pointers (copy-engine probe).
 * ------------------------------------------------------------------------------------------------ */
int b2_dpx_ipc_export(const void* dev_ptr, unsigned char* handle_out /* 64 bytes */, int64_t* offset_out);
int b2_dpx_ipc_import(const unsigned char* handle /* 64 bytes */, int64_t offset, void** dev_ptr_out);
int b2_dpx_alloc_flags(void** flags_out);
int b2_dpx_max_chunks(void);
int b2_dpx_create(int rank, int world, int mode, void* const* grad_ptrs, void* const* flag_ptrs,
                  void* const* staging_ptrs, int64_t staging_slot_elems, int n_copy_streams /* 0: default */,
                  void** handle_out);
int b2_dpx_exchange(void* handle, int chunk, uint32_t seq, int n_ranges, const int64_t* range_off,
                    const int64_t* range_len, int64_t staging_base, void* main_stream);
int b2_dpx_finish(void* handle, uint32_t seq, void* main_stream);
/* b2_dpx_finish + the global gradient norm for free: *gnorm_sq_out = sum over the whole exchanged buffer of (reduced bf16
 * gradient)^2 — what b2_sumsq over the buffer would return (replaces the 5 GB read of clip_grad_norm_'s norm pass,
 * flow_matching_trainer.py:181-186).  Copy-engine transports only (b2_dpx_norm_supported). */
int b2_dpx_finish_norm(void* handle, uint32_t seq, void* main_stream, double* gnorm_sq_out);
int b2_dpx_norm_supported(void* handle);
int b2_dpx_memcpy_async(void* dst, const void* src, int64_t bytes, void* stream);
int b2_dpx_destroy(void* handle);

#ifdef __cplusplus
}
#endif
#endif /* SDXL_B200_H */
