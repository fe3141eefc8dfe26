// Noise sampling, noising, target construction, MSE loss (+ dL/dpred) and the optimizer-side kernels.
//
// Arithmetic follows the reference's loss code:
//   ddpm: noisy = clamp(x + sigma_t * eps, +-20000)      (novelai_v3.py:111-120)
//         target = eps | (eps - x) / sqrt(sigma_t^2)      (ddpm_trainer.py:328-333, novelai_v3.py:122-127)
//   flow: xt = (1 - t) x0 + t x1, target = x1 - x0       (flow_matching_trainer.py:387-390, :414)
//         evaluated op-by-op in bf16 exactly as torch does for bf16 tensors (:288-306 cast everything to bf16)
//   loss = mean((pred - target)^2 [* w_b]) (* tag weight), non-finite -> 1000, else min(loss, 1000)
//         (ddpm_trainer.py:336-384, flow_matching_trainer.py:419-335); accumulated in fp64, reported fp32 (B23).
#include "common.cuh"

namespace b2 {

// ---------------------------------------------------------------- Philox4x32-10
struct u4 { uint32_t x, y, z, w; };
__device__ __forceinline__ u4 philox4x32_10(u4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = u4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ float u01(uint32_t u) { return ((float)(u >> 8) + 0.5f) * (1.0f / 16777216.0f); }

__global__ void randn_kernel(float* out, long long n, const uint64_t* seed_offset, uint64_t stream_id, int round_bf16) {
  const uint64_t seed = seed_offset[0], off = seed_offset[1];
  const long long nq = (n + 3) / 4;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (long long)gridDim.x * blockDim.x) {
    u4 c{(uint32_t)q, (uint32_t)((uint64_t)q >> 32), (uint32_t)(off + stream_id * 0x9E3779B97F4A7C15ull),
         (uint32_t)((off + stream_id * 0x9E3779B97F4A7C15ull) >> 32)};
    u4 r = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    float z[4];
    {
      const float r0 = sqrtf(-2.f * logf(u01(r.x))), a0 = 6.283185307179586f * u01(r.y);
      const float r1 = sqrtf(-2.f * logf(u01(r.z))), a1 = 6.283185307179586f * u01(r.w);
      z[0] = r0 * cosf(a0); z[1] = r0 * sinf(a0); z[2] = r1 * cosf(a1); z[3] = r1 * sinf(a1);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long e = q * 4 + j;
      if (e < n) out[e] = round_bf16 ? __bfloat162float(__float2bfloat16(z[j])) : z[j];
    }
  }
}

__global__ void philox_advance_kernel(uint64_t* seed_offset, uint64_t inc) { seed_offset[1] += inc; }

__device__ __forceinline__ float rbf(float x) { return __bfloat162float(__float2bfloat16(x)); }

// x, eps: fp32 NCHW [B,C,HW];  noisy: bf16 [B,HW,Cpad] (pad channels zero);  target: fp32 NCHW
__global__ void make_noisy_kernel(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ st,
                                  int mode, int vpred, int clampz, bf16* __restrict__ noisy, float* __restrict__ target, int B,
                                  int C, int HW, int Cpad) {
  const long long total = (long long)B * HW * Cpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cpad);
    const long long t = i / Cpad;
    const int p = (int)(t % HW);
    const int b = (int)(t / HW);
    if (c >= C) {
      noisy[i] = __float2bfloat16(0.f);
      continue;
    }
    const long long src = ((long long)b * C + c) * HW + p;
    const float xv = x[src], ev = eps[src], s = st[b];
    float nz, tg;
    if (mode == 0) {
      nz = xv + s * ev;
      if (clampz) nz = fminf(fmaxf(nz, -20000.f), 20000.f);
      tg = vpred ? (ev - xv) / sqrtf(s * s) : ev;
    } else {
      // bf16 op-by-op, as torch evaluates (1 - t) * x0 + t * x1 on bf16 tensors
      const float one_minus = rbf(1.f - s);
      nz = rbf(rbf(one_minus * ev) + rbf(s * xv));
      tg = rbf(xv - ev);
    }
    noisy[i] = __float2bfloat16(nz);
    target[src] = tg;
  }
}

__global__ void mse_loss_kernel(const bf16* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ weight,
                                double* loss_sum, bf16* __restrict__ dpred, float gscale, int B, int C, int HW, int Cpad) {
  const long long total = (long long)B * HW * Cpad;
  float local = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cpad);
    const long long t = i / Cpad;
    const int p = (int)(t % HW);
    const int b = (int)(t / HW);
    float g = 0.f;
    if (c < C) {
      const float w = weight ? weight[b] : 1.f;
      const float d = __bfloat162float(pred[i]) - target[((long long)b * C + c) * HW + p];
      local += w * d * d;
      g = 2.f * w * d * gscale;
    }
    if (dpred) dpred[i] = __float2bfloat16(g);
  }
  local = warp_sum(local);
  __shared__ float sm[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sm[warp] = local;
  __syncthreads();
  if (warp == 0) {
    float v = lane < (blockDim.x >> 5) ? sm[lane] : 0.f;
    v = warp_sum(v);
    if (lane == 0) atomicAdd(loss_sum, (double)v);
  }
}

__global__ void finalize_loss_kernel(const double* loss_sum, double count, float scale, float* loss_out, int32_t* ok) {
  const double l = loss_sum[0] / count * (double)scale;
  const float lf = (float)l;
  if (!isfinite(lf)) {
    *loss_out = 1000.f;  // ddpm_trainer.py:380-382 (B22: contributes no gradient)
    *ok = 0;
  } else if (lf > 1000.f) {
    *loss_out = 1000.f;  // clamp(max=1000): zero gradient through the clamp
    *ok = 0;
  } else {
    *loss_out = lf;
    *ok = 1;
  }
}
__global__ void mask_grad_kernel(bf16* g, long long n, const int32_t* ok) {
  if (*ok) return;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    g[i] = __float2bfloat16(0.f);
}

// ---------------------------------------------------------------- optimizer side
__global__ void sumsq_kernel(const bf16* __restrict__ g, long long n, double* out) {
  float local = 0.f;
  const long long nv = n >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    unpack8(ld8(g + i * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) local += f[j] * f[j];
  }
  if (blockIdx.x == 0)
    for (long long i = nv * 8 + threadIdx.x; i < n; i += blockDim.x) {
      const float f = __bfloat162float(g[i]);
      local += f * f;
    }
  local = warp_sum(local);
  __shared__ float sm[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sm[warp] = local;
  __syncthreads();
  if (warp == 0) {
    float v = lane < (blockDim.x >> 5) ? sm[lane] : 0.f;
    v = warp_sum(v);
    if (lane == 0) atomicAdd(out, (double)v);
  }
}

__global__ void adamw_kernel(bf16* __restrict__ p, float* __restrict__ master, const bf16* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float wd, float bc1,
                             float bc2, const double* gnorm_sq, float max_norm, float grad_scale,
                             const uint64_t* __restrict__ dev_step) {
  if (dev_step) {  // step counter lives on the device so that a captured CUDA graph advances it on every replay
    const float st = (float)dev_step[0];
    bc1 = 1.f - powf(b1, st);
    bc2 = 1.f - powf(b2, st);
  }
  float clip = grad_scale;
  if (gnorm_sq && max_norm > 0.f) {
    const float norm = (float)sqrt(*gnorm_sq) * grad_scale;
    const float coef = max_norm / (norm + 1e-6f);  // torch.nn.utils.clip_grad_norm_
    if (coef < 1.f) clip *= coef;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = __bfloat162float(g[i]) * clip;
    float w = master ? master[i] : __bfloat162float(p[i]);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    w = w * (1.f - lr * wd);
    w -= lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
    if (master) master[i] = w;
    p[i] = __float2bfloat16(w);
  }
}

// out[0] += sum |x|, out[1] += sum x^2   (metrics: noise_scale / pred_scale / *_norm of the reference's dicts)
template <typename T>
__global__ void abs_sq_sums_kernel(const T* __restrict__ x, long long n, int period, int valid, double* out) {
  float sa = 0.f, sq = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (period > 0 && (int)(i % period) >= valid) continue;  // skip channel padding
    const float f = (float)x[i];
    sa += fabsf(f);
    sq += f * f;
  }
  sa = warp_sum(sa);
  sq = warp_sum(sq);
  __shared__ float sm[2][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sm[0][warp] = sa; sm[1][warp] = sq; }
  __syncthreads();
  if (warp == 0) {
    float a = lane < (blockDim.x >> 5) ? sm[0][lane] : 0.f;
    float q = lane < (blockDim.x >> 5) ? sm[1][lane] : 0.f;
    a = warp_sum(a);
    q = warp_sum(q);
    if (lane == 0) { atomicAdd(&out[0], (double)a); atomicAdd(&out[1], (double)q); }
  }
}

__global__ void scale_bf16_kernel(bf16* x, long long n, const float* scale, float host_scale) {
  const float s = (scale ? *scale : 1.f) * host_scale;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = __float2bfloat16(__bfloat162float(x[i]) * s);
}

static inline int red_blocks(long long items) {
  long long b = (items + 255) / 256;
  const long long cap = 8LL * num_sms();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace b2

using namespace b2;

extern "C" int b2_randn(float* out, int64_t n, const uint64_t* seed_offset, uint64_t stream_id, int round_bf16, void* stream) {
  B2_REQUIRE(out && seed_offset && n > 0, "b2_randn: bad args");
  randn_kernel<<<red_blocks((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(out, n, seed_offset, stream_id, round_bf16);
  return check_launch("randn");
}
extern "C" int b2_philox_advance(uint64_t* seed_offset, uint64_t inc, void* stream) {
  B2_REQUIRE(seed_offset, "b2_philox_advance: null");
  philox_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(seed_offset, inc);
  return check_launch("philox_advance");
}
extern "C" int b2_make_noisy(const float* x, const float* eps, const float* sigma_or_t, int mode, int v_prediction,
                             int clamp_ztsnr, void* noisy, float* target, int B, int C, int HW, int Cpad, void* stream) {
  B2_REQUIRE(x && eps && sigma_or_t && noisy && target && Cpad >= C, "b2_make_noisy: bad args");
  make_noisy_kernel<<<red_blocks((long long)B * HW * Cpad), 256, 0, (cudaStream_t)stream>>>(
      x, eps, sigma_or_t, mode, v_prediction, clamp_ztsnr, (bf16*)noisy, target, B, C, HW, Cpad);
  return check_launch("make_noisy");
}
extern "C" int b2_mse_loss(const void* pred, const float* target, const float* weight, double* loss_sum, void* dpred,
                           float gscale, int B, int C, int HW, int Cpad, void* stream) {
  B2_REQUIRE(pred && target && loss_sum && Cpad >= C, "b2_mse_loss: bad args");
  mse_loss_kernel<<<red_blocks((long long)B * HW * Cpad), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)pred, target, weight, loss_sum, (bf16*)dpred, gscale, B, C, HW, Cpad);
  return check_launch("mse_loss");
}
extern "C" int b2_finalize_loss(const double* loss_sum, double count, float scale, float* loss_out, int32_t* ok,
                                void* dpred, int64_t n_dpred, void* stream) {
  B2_REQUIRE(loss_sum && loss_out && ok && count > 0, "b2_finalize_loss: bad args");
  finalize_loss_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(loss_sum, count, scale, loss_out, ok);
  int rc = check_launch("finalize_loss");
  if (rc || !dpred) return rc;
  mask_grad_kernel<<<red_blocks(n_dpred), 256, 0, (cudaStream_t)stream>>>((bf16*)dpred, n_dpred, ok);
  return check_launch("mask_grad");
}
extern "C" int b2_abs_sq_sums(const void* x, int is_fp32, int64_t n, int period, int valid, double* out, void* stream) {
  B2_REQUIRE(x && out && n > 0, "b2_abs_sq_sums: bad args");
  if (is_fp32)
    abs_sq_sums_kernel<float><<<red_blocks(n), 256, 0, (cudaStream_t)stream>>>((const float*)x, n, period, valid, out);
  else
    abs_sq_sums_kernel<bf16><<<red_blocks(n), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, n, period, valid, out);
  return check_launch("abs_sq_sums");
}
extern "C" int b2_scale_bf16(void* x, int64_t n, const float* dev_scale, float host_scale, void* stream) {
  B2_REQUIRE(x && n > 0, "b2_scale_bf16: bad args");
  scale_bf16_kernel<<<red_blocks(n), 256, 0, (cudaStream_t)stream>>>((bf16*)x, n, dev_scale, host_scale);
  return check_launch("scale_bf16");
}
extern "C" int b2_sumsq(const void* g, int64_t n, double* out, void* stream) {
  B2_REQUIRE(g && out && n > 0 && !((uintptr_t)g & 15), "b2_sumsq: bad args");
  sumsq_kernel<<<red_blocks(n / 8 + 1), 256, 0, (cudaStream_t)stream>>>((const bf16*)g, n, out);
  return check_launch("sumsq");
}
extern "C" int b2_adamw(void* p, float* master, const void* g, float* m, float* v, int64_t n, float lr, float beta1,
                        float beta2, float eps, float weight_decay, int step, const double* gnorm_sq, float max_norm,
                        float grad_scale, const uint64_t* dev_step, void* stream) {
  B2_REQUIRE(p && g && m && v && n > 0 && (step >= 1 || dev_step), "b2_adamw: bad args");
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  adamw_kernel<<<red_blocks(n), 256, 0, (cudaStream_t)stream>>>((bf16*)p, master, (const bf16*)g, m, v, n, lr, beta1, beta2,
                                                                eps, weight_decay, bc1, bc2, gnorm_sq, max_norm, grad_scale,
                                                                step >= 1 ? nullptr : dev_step);
  return check_launch("adamw");
}
