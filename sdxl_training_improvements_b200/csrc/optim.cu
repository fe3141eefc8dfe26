// AdamWBF16 — the reference's default optimizer (`optimizer_type: adamw_bf16`, src/config.yaml) as ONE fused kernel over
// the flat bf16 buffers: bf16 parameters, bf16 exp_avg / exp_avg_sq, and a bf16 "shift" residual that carries what the
// bf16 parameter could not absorb, every add rounded stochastically with 16 random low bits.
//
// Follows src/training/optimizers/adamw_bfloat16/__init__.py:150-197 (`_make_step`) and
// stochastic/__init__.py:46-124 op by op, including the operand order of `add_stochastic_` AS WRITTEN
// (result = other + alpha * input, stochastic/__init__.py:96-107 — so the first-moment update the reference actually
// computes is  exp_avg <- SR(grad + (1 - beta1) * (beta1 * exp_avg)),  not the textbook EMA; `as_written = 0`
// selects the documented intent instead).  Each torch bf16 op rounds to nearest-even; the three `*_stochastic_`
// helpers compute in fp32 and round with SR(x) = (bits(x) + rand16) & 0xFFFF0000.
//
// Traffic: reads p, g, m, v, shift and writes p, m, v, shift = 18 B / parameter (46 GB for the SDXL UNet), HBM-bound.
// Random bits: 64 per element (4 roundings x 16 bits) from a counter-based hash — two evaluations of the "triple32" integer
// mixer (3 multiplies + 4 xor-shifts, avalanche bias 0.02 bits) of (element index, optimizer step, seed).  Round 1 used
// Philox4x32-7 here: ~35 integer instructions per element out of ~80 in a kernel whose ALU time (5.7 ms at perfect issue)
// sits right under its HBM time (7.0 ms for 46 GB) — measured 10.8 ms.  Stochastic rounding needs unbiased, de-correlated
// low bits, not a cryptographic stream; the hash costs ~20 instructions per element.  Reproducible and independent of the
// launch geometry: the counter is the element index, the key is (seed, step).
#include "common.cuh"

namespace b2 {

struct pu4 { uint32_t x, y, z, w; };
__device__ __forceinline__ uint32_t lowbias32(uint32_t x) {  // "triple32": three multiply / xor-shift rounds
  x ^= x >> 17;
  x *= 0xed5ad4bbu;
  x ^= x >> 11;
  x *= 0xac4c1b51u;
  x ^= x >> 15;
  x *= 0x31848babu;
  x ^= x >> 14;
  return x;
}
// two independent 32-bit words for element `e` under the per-launch keys (ka, kb)
__device__ __forceinline__ void rand64(uint64_t e, uint32_t ka, uint32_t kb, uint32_t& r01, uint32_t& r23) {
  const uint32_t lo = (uint32_t)e, hi = (uint32_t)(e >> 32) * 0xC2B2AE35u;
  r01 = lowbias32(lo ^ ka ^ hi);
  r23 = lowbias32(lo ^ kb ^ hi);
}

__device__ __forceinline__ float rn_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
// stochastic/__init__.py:46-71: add 16 random bits below the bf16 mantissa, truncate
__device__ __forceinline__ float sr_bf16(float x, uint32_t r16) {
  return __uint_as_float((__float_as_uint(x) + r16) & 0xffff0000u);
}

struct AdamBF16P {
  float b1, b2, eps;
  float omb1, omb2;  // (float)(1.0 - beta) evaluated in double like the Python scalars `1 - beta1`, `1 - beta2`
  float step_size;  // -lr * sqrt(1 - beta2^step)   (host step)
  double lr, b2d;   // device step (seed_offset[1]): step_size is recomputed in the kernel
  int dev_step;
  float max_norm, grad_scale;
  int as_written;
  int rng_mode;     // 0: Philox; 1: rand16 = 0 (truncate); 2: rand16 = 0xFFFF; 3: rand16 read from `test_rand16`
                    // (int32 [4, n]: test hooks for bit-exact parity with the reference's own functions)
};

__device__ __forceinline__ void adam_bf16_elem(float& p, float g, float& m, float& v, float& s, const AdamBF16P& a,
                                               float clip, uint32_t r01, uint32_t r23) {
  // clip_grad_norm_ scales the bf16 gradient in place (flow_matching_trainer.py:181-186) -> one bf16 rounding
  const float gi = clip == 1.f ? g : rn_bf16(__fmul_rn(g, clip));
  // exp_avg.mul_(beta1); add_stochastic_(exp_avg, grad, alpha=1-beta1)
  const float m1 = rn_bf16(__fmul_rn(m, a.b1));
  const float one_b1 = a.omb1;
  // torch's add-with-alpha is a fused multiply-add (vec::fmadd on CPU, nvcc contraction on CUDA)
  const float mr = a.as_written ? fmaf(one_b1, m1, gi) : fmaf(one_b1, gi, m1);
  m = sr_bf16(mr, r01 & 0xffffu);
  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
  const float v1 = rn_bf16(__fmul_rn(v, a.b2));
  v = rn_bf16(__fadd_rn(v1, __fmul_rn(__fmul_rn(a.omb2, gi), gi)));
  // denom = exp_avg_sq.sqrt().add_(eps)
  const float den = rn_bf16(__fadd_rn(rn_bf16(__fsqrt_rn(v)), a.eps));
  // addcdiv_stochastic_(shift, exp_avg, denom, value=-lr*denom_correction)
  const float s1 = sr_bf16(__fadd_rn(s, __fdiv_rn(__fmul_rn(a.step_size, m), den)), r01 >> 16);
  // buffer = p.clone(); add_stochastic_(p, shift); add_stochastic_(shift, buffer.sub_(p))
  const float p1 = sr_bf16(__fadd_rn(s1, p), r23 & 0xffffu);
  const float d = rn_bf16(__fsub_rn(p, p1));
  s = sr_bf16(__fadd_rn(d, s1), r23 >> 16);
  p = p1;
}

__global__ void __launch_bounds__(256)
adamw_bf16_kernel(bf16* __restrict__ p, const bf16* __restrict__ g, bf16* __restrict__ m, bf16* __restrict__ v,
                  bf16* __restrict__ sh, long long n, AdamBF16P a, const double* __restrict__ gnorm_sq,
                  const uint64_t* __restrict__ seed_offset, uint64_t step, const int32_t* __restrict__ test_rand16) {
  float clip = a.grad_scale;
  if (gnorm_sq && a.max_norm > 0.f) {
    const float norm = (float)sqrt(*gnorm_sq) * a.grad_scale;
    const float coef = a.max_norm / (norm + 1e-6f);  // torch.nn.utils.clip_grad_norm_
    if (coef < 1.f) clip *= coef;
  }
  const uint64_t seed = seed_offset ? seed_offset[0] : 0;
  if (a.dev_step) {  // CUDA-graph replays advance seed_offset[1] on the device
    step = seed_offset[1];
    a.step_size = (float)(-a.lr * sqrt(1.0 - pow(a.b2d, (double)step)));
  }
  // per-launch keys: (seed, step) scrambled twice with different constants
  const uint32_t kmix = lowbias32((uint32_t)seed ^ lowbias32((uint32_t)(seed >> 32) + 0x9E3779B9u) ^
                                  lowbias32((uint32_t)step * 0x85EBCA6Bu + (uint32_t)(step >> 32)));
  const uint32_t ka = kmix, kb = lowbias32(kmix ^ 0x68E31DA4u) | 1u;
  const long long nv = n >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // two 16-byte vectors per array in flight per thread (10 independent loads): the loop is latency-bound otherwise
  for (long long q0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; q0 < nv; q0 += 2 * stride) {
    bf16x8 vp[2], vg[2], vm[2], vv[2], vs[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long q = q0 + u * stride;
      if (q < nv) {
        vp[u] = ld8(p + q * 8); vg[u] = ld8(g + q * 8); vm[u] = ld8(m + q * 8); vv[u] = ld8(v + q * 8); vs[u] = ld8(sh + q * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long q = q0 + u * stride;
      if (q >= nv) break;
      float fp[8], fg[8], fm[8], fv[8], fs[8];
      unpack8(vp[u], fp);
      unpack8(vg[u], fg);
      unpack8(vm[u], fm);
      unpack8(vv[u], fv);
      unpack8(vs[u], fs);
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        pu4 r;
        if (a.rng_mode == 0) {
          rand64((uint64_t)q * 8 + 2 * h, ka, kb, r.x, r.y);
          rand64((uint64_t)q * 8 + 2 * h + 1, ka, kb, r.z, r.w);
        } else if (a.rng_mode == 3) {
          const long long e = q * 8 + 2 * h;
          r.x = (uint32_t)test_rand16[e] | ((uint32_t)test_rand16[n + e] << 16);
          r.y = (uint32_t)test_rand16[2 * n + e] | ((uint32_t)test_rand16[3 * n + e] << 16);
          r.z = (uint32_t)test_rand16[e + 1] | ((uint32_t)test_rand16[n + e + 1] << 16);
          r.w = (uint32_t)test_rand16[2 * n + e + 1] | ((uint32_t)test_rand16[3 * n + e + 1] << 16);
        } else {
          const uint32_t f = a.rng_mode == 1 ? 0u : 0xffffffffu;
          r = pu4{f, f, f, f};
        }
        adam_bf16_elem(fp[2 * h], fg[2 * h], fm[2 * h], fv[2 * h], fs[2 * h], a, clip, r.x, r.y);
        adam_bf16_elem(fp[2 * h + 1], fg[2 * h + 1], fm[2 * h + 1], fv[2 * h + 1], fs[2 * h + 1], a, clip, r.z, r.w);
      }
      // values are exact bf16 (low 16 bits zero): packing is a truncation, not a second rounding
      st8(p + q * 8, pack8(fp));
      st8(m + q * 8, pack8(fm));
      st8(v + q * 8, pack8(fv));
      st8(sh + q * 8, pack8(fs));
    }
  }
  // tail (n % 8 elements), one thread
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (long long i = nv * 8; i < n; ++i) {
      pu4 r;
      if (a.rng_mode == 0) {
        rand64((uint64_t)i, ka, kb, r.x, r.y);
        r.z = r.w = 0;
      } else if (a.rng_mode == 3) {
        r.x = (uint32_t)test_rand16[i] | ((uint32_t)test_rand16[n + i] << 16);
        r.y = (uint32_t)test_rand16[2 * n + i] | ((uint32_t)test_rand16[3 * n + i] << 16);
        r.z = r.w = 0;
      } else {
        const uint32_t f = a.rng_mode == 1 ? 0u : 0xffffffffu;
        r = pu4{f, f, f, f};
      }
      float fp = __bfloat162float(p[i]), fm = __bfloat162float(m[i]), fv = __bfloat162float(v[i]),
            fs = __bfloat162float(sh[i]);
      adam_bf16_elem(fp, __bfloat162float(g[i]), fm, fv, fs, a, clip, r.x, r.y);
      p[i] = __float2bfloat16_rn(fp);
      m[i] = __float2bfloat16_rn(fm);
      v[i] = __float2bfloat16_rn(fv);
      sh[i] = __float2bfloat16_rn(fs);
    }
  }
}

// y <- bf16(y + alpha * x): the deferred weight decay `shift.add_(p, alpha=-decay)` (adamw_bfloat16/__init__.py:191-192)
__global__ void axpy_bf16_kernel(bf16* __restrict__ y, const bf16* __restrict__ x, long long n, float alpha) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16_rn(__fadd_rn(__bfloat162float(y[i]), __fmul_rn(alpha, __bfloat162float(x[i]))));
}

}  // namespace b2

using namespace b2;

extern "C" int b2_adamw_bf16(void* p, const void* g, void* m, void* v, void* shift, int64_t n, double lr, double beta1,
                             double beta2, double eps, int step, const double* gnorm_sq, float max_norm, float grad_scale,
                             const uint64_t* seed_offset, int as_written, int rng_mode, const int32_t* test_rand16,
                             void* stream) {
  B2_REQUIRE(p && g && m && v && shift && n > 0 && (step >= 1 || seed_offset), "b2_adamw_bf16: bad args");
  B2_REQUIRE(!((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(shift)) & 15),
             "b2_adamw_bf16: buffers must be 16-byte aligned");
  B2_REQUIRE(rng_mode >= 0 && rng_mode <= 3 && (rng_mode != 3 || test_rand16), "b2_adamw_bf16: bad rng_mode");
  AdamBF16P a;
  // hyper-parameters arrive as doubles (Python floats) and are narrowed exactly where torch narrows them
  a.b1 = (float)beta1; a.b2 = (float)beta2; a.eps = (float)eps;
  a.omb1 = (float)(1.0 - beta1);
  a.omb2 = (float)(1.0 - beta2);
  // python: value = -lr * (1 - beta2**step) ** 0.5  (double), passed to a float kernel argument
  a.step_size = (float)(-lr * sqrt(1.0 - pow(beta2, (double)step)));
  a.lr = lr; a.b2d = beta2; a.dev_step = step >= 1 ? 0 : 1;
  a.max_norm = max_norm; a.grad_scale = grad_scale;
  a.as_written = as_written; a.rng_mode = rng_mode;
  long long blocks = ((n >> 3) + 255) / 256;
  const long long cap = 16LL * num_sms();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adamw_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((bf16*)p, (const bf16*)g, (bf16*)m, (bf16*)v,
                                                                        (bf16*)shift, n, a, gnorm_sq, seed_offset,
                                                                        (uint64_t)step, test_rand16);
  return check_launch("adamw_bf16");
}

extern "C" int b2_axpy_bf16(void* y, const void* x, int64_t n, float alpha, void* stream) {
  B2_REQUIRE(y && x && n > 0, "b2_axpy_bf16: bad args");
  long long blocks = (n + 255) / 256;
  const long long cap = 8LL * num_sms();
  if (blocks > cap) blocks = cap;
  axpy_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((bf16*)y, (const bf16*)x, n, alpha);
  return check_launch("axpy_bf16");
}
