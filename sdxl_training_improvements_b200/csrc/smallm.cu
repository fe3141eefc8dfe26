// Few-row linears on CUDA cores: the 21 Linear layers that act on ONE ROW PER SAMPLE — `time_embedding.linear_1/2`,
// `add_embedding.linear_1/2` and the 17 ResnetBlock2D `time_emb_proj` (diffusers UNet2DConditionModel.get_time_embed /
// ResnetBlock2D.forward: `temb = self.time_emb_proj(self.nonlinearity(temb))`), forward, input gradient and weight gradient:
// 63 launches per training step with M = batch size = 4.  On a 128-row tensor-core tile 97 % of every MMA is padding and the
// launch costs 26 us (set-up, TMA pipeline fill, a 1280-deep k loop on one or two CTAs); here the weight matrix (<= 7 MB)
// is simply streamed once by the whole GPU: ~3 us, HBM / latency bound.
//
//   fwd   (a_mn 0, b_mn 0, M <= 8): D[m,n] = alpha sum_k A[m,k] B[n,k] (+bias) (+residual) (+D_old)   one warp per n
//   dgrad (a_mn 0, b_mn 1, M <= 8): D[m,n] = alpha sum_k A[m,k] B[k,n]  (+D_old)                      64 columns per CTA
//   wgrad (a_mn 1, b_mn 1, K <= 8): D[m,n] = alpha sum_k A[k,m] B[k,n]  (+D_old)                      one 8-vector per thread
// fp32 accumulation in a fixed order (deterministic); same epilogue order as the GEMM kernels (alpha, bias, residual, D_old).
#include "../../include/sdxl_b200.h"
#include "common.cuh"

namespace b2 {

constexpr int SM_MAXM = 8;

struct SmallMP {
  const bf16* A;
  const bf16* B;
  void* D;
  const bf16* bias;
  const bf16* R;
  int M, N, K;
  long long lda, ldb, ldd, ldr;
  float alpha;
  int accumulate, out_fp32;
  int bias_rows_per_group;
  long long bias_group_stride;
};

__device__ __forceinline__ void smallm_store(const SmallMP& p, int m, int n, float x) {
  x *= p.alpha;
  if (p.bias) x += __bfloat162float(p.bias[(long long)(m / p.bias_rows_per_group) * p.bias_group_stride + n]);
  if (p.R) x += __bfloat162float(p.R[(long long)m * p.ldr + n]);
  const long long o = (long long)m * p.ldd + n;
  if (p.out_fp32) {
    float* Df = reinterpret_cast<float*>(p.D);
    if (p.accumulate) x += Df[o];
    Df[o] = x;
  } else {
    bf16* Db = reinterpret_cast<bf16*>(p.D);
    if (p.accumulate) x += __bfloat162float(Db[o]);
    Db[o] = __float2bfloat16(x);
  }
}

// ---- forward: one warp per output column, lanes stride over k in 16-byte vectors, A rows come from L1
__global__ void __launch_bounds__(256) smallm_fwd_kernel(SmallMP p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n >= p.N) return;
  float acc[SM_MAXM];
#pragma unroll
  for (int m = 0; m < SM_MAXM; ++m) acc[m] = 0.f;
  const bf16* brow = p.B + (long long)n * p.ldb;
  const int kv = p.K >> 3;
  for (int v = lane; v < kv; v += 32) {
    float b[8];
    unpack8(ld8(brow + v * 8), b);
#pragma unroll
    for (int m = 0; m < SM_MAXM; ++m) {
      if (m < p.M) {
        float a[8];
        unpack8(ld8(p.A + (long long)m * p.lda + v * 8), a);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[m] = fmaf(a[j], b[j], acc[m]);
      }
    }
  }
#pragma unroll
  for (int m = 0; m < SM_MAXM; ++m) {
    if (m < p.M) {
      const float s = warp_sum(acc[m]);
      if (lane == m) smallm_store(p, m, n, s);
    }
  }
}

// ---- input gradient: CTA = 64 output columns (one bf16 pair per lane, 128-byte rows), 32 warps split the k rows,
//      partial sums meet in shared memory and are added in warp order
__global__ void __launch_bounds__(1024) smallm_dgrad_kernel(SmallMP p) {
  extern __shared__ float red[];  // [32 warps][M][64]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 64 + lane * 2;
  float acc[SM_MAXM][2];
#pragma unroll
  for (int m = 0; m < SM_MAXM; ++m) acc[m][0] = acc[m][1] = 0.f;
  if (n < p.N) {
#pragma unroll 8
    for (int k = warp; k < p.K; k += 32) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(p.B + (long long)k * p.ldb + n);
      const float w0 = __uint_as_float(w << 16), w1 = __uint_as_float(w & 0xffff0000u);
#pragma unroll
      for (int m = 0; m < SM_MAXM; ++m) {
        if (m < p.M) {
          const float a = __bfloat162float(p.A[(long long)m * p.lda + k]);
          acc[m][0] = fmaf(a, w0, acc[m][0]);
          acc[m][1] = fmaf(a, w1, acc[m][1]);
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < SM_MAXM; ++m) {
    if (m < p.M) {
      red[(warp * p.M + m) * 64 + lane * 2] = acc[m][0];
      red[(warp * p.M + m) * 64 + lane * 2 + 1] = acc[m][1];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.M * 64; i += blockDim.x) {
    const int m = i >> 6, c = i & 63;
    const int gn = blockIdx.x * 64 + c;
    if (gn >= p.N) continue;
    float s = 0.f;
#pragma unroll 8
    for (int w = 0; w < 32; ++w) s += red[(w * p.M + m) * 64 + c];
    smallm_store(p, m, gn, s);
  }
}

// ---- weight gradient: D[m, n..n+8) (+)= sum_{k < K <= 8} A[k,m] B[k, n..n+8): one 16-byte vector of D per thread
__global__ void __launch_bounds__(256) smallm_wgrad_kernel(SmallMP p) {
  const int nv = p.N >> 3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)p.M * nv) return;
  const int m = (int)(i / nv), v = (int)(i % nv);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
  for (int k = 0; k < SM_MAXM; ++k) {
    if (k < p.K) {
      const float a = __bfloat162float(p.A[(long long)k * p.lda + m]);
      float b[8];
      unpack8(ld8(p.B + (long long)k * p.ldb + v * 8), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(a, b[j], acc[j]);
    }
  }
  const long long o = (long long)m * p.ldd + v * 8;
  if (p.out_fp32) {
    float* Df = reinterpret_cast<float*>(p.D) + o;
#pragma unroll
    for (int j = 0; j < 8; ++j) Df[j] = acc[j] * p.alpha + (p.accumulate ? Df[j] : 0.f);
  } else {
    bf16* Db = reinterpret_cast<bf16*>(p.D) + o;
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= p.alpha;
    if (p.accumulate) {
      unpack8(ld8(Db), t);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += t[j];
    }
    st8(Db, pack8(acc));
  }
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Returns true when the call was taken (out_rc = its status); false = not a few-row shape, the caller goes on to the GEMMs.
bool smallm_try(const b2_gemm_args* a, cudaStream_t st, int* out_rc) {
  static const bool off = getenv("B2_NO_SMALLM") != nullptr;
  if (off || a->nb_lo != 1 || a->nb_hi != 1) return false;
  SmallMP p;
  p.A = reinterpret_cast<const bf16*>(a->A);
  p.B = reinterpret_cast<const bf16*>(a->B);
  p.D = a->D;
  p.bias = reinterpret_cast<const bf16*>(a->bias);
  p.R = reinterpret_cast<const bf16*>(a->residual);
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.lda = a->lda; p.ldb = a->ldb; p.ldd = a->ldd; p.ldr = a->ldr;
  p.alpha = a->alpha;
  p.accumulate = a->accumulate; p.out_fp32 = a->out_fp32;
  p.bias_rows_per_group = a->bias_rows_per_group > 0 ? a->bias_rows_per_group : 0x7fffffff;
  p.bias_group_stride = a->bias_group_stride;
  if (!a->a_mn && !a->b_mn && a->M <= SM_MAXM && a->N >= 64) {
    if ((a->K & 7) || (a->lda & 7) || (a->ldb & 7) || !al16(a->A) || !al16(a->B)) return false;
    smallm_fwd_kernel<<<(a->N + 7) / 8, 256, 0, st>>>(p);
    *out_rc = check_launch("smallm_fwd");
    return true;
  }
  if (!a->a_mn && a->b_mn && a->M <= SM_MAXM && a->N >= 64 && !a->bias && !a->residual) {
    if ((a->N & 1) || (a->ldb & 1) || (reinterpret_cast<uintptr_t>(a->B) & 3)) return false;
    static bool attr_done = false;
    if (!attr_done) {
      cudaFuncSetAttribute(smallm_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * SM_MAXM * 64 * 4);
      attr_done = true;
    }
    smallm_dgrad_kernel<<<(a->N + 63) / 64, 1024, (size_t)32 * a->M * 64 * 4, st>>>(p);
    *out_rc = check_launch("smallm_dgrad");
    return true;
  }
  if (a->a_mn && a->b_mn && a->K <= SM_MAXM && a->N >= 64 && !a->bias && !a->residual) {
    if ((a->N & 7) || (a->ldb & 7) || (a->ldd & 7) || !al16(a->B) || !al16(a->D)) return false;
    const long long threads = (long long)a->M * (a->N >> 3);
    smallm_wgrad_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(p);
    *out_rc = check_launch("smallm_wgrad");
    return true;
  }
  return false;
}

}  // namespace b2
