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
// Traffic: reads p, g, m, v, shift and writes p, m, v, shift = 18 B / parameter (46 GB for the SDXL UNet), HBM-bound
// (+2 B when the gradient is zeroed in the same pass).
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
// 64 random bits for element `lo` (+ the pre-multiplied high index word `hi`) under the per-launch keys (ka, kb): the first
// word is the full mixer of the counter, the second a one-round multiply / xor-shift of the first under the other key —
// each 16-bit field is uniform, and the four fields of one element only have to be de-correlated, not independent streams
__device__ __forceinline__ void rand64(uint32_t lo, uint32_t hi, uint32_t ka, uint32_t kb, uint32_t& r01, uint32_t& r23) {
  r01 = lowbias32(lo ^ ka ^ hi);
  const uint32_t t = (r01 ^ kb) * 0x9E3779B1u;
  r23 = t ^ (t >> 15);
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
  int zero_grad;
  int rng_mode;     // 0: counter hash; 1: rand16 = 0 (truncate); 2: rand16 = 0xFFFF; 3: rand16 read from `test_rand16`
                    // (int32 [4, n]: test hooks for bit-exact parity with the reference's own functions)
};

// two roundings to bf16 (nearest-even) through ONE packed convert
__device__ __forceinline__ void rn2_bf16(float& a, float& b) {
  const uint32_t u = pack2_bf16(a, b);
  a = __uint_as_float(u << 16);
  b = __uint_as_float(u & 0xffff0000u);
}
// sqrt of a NON-NEGATIVE bf16 VALUE, to be rounded to bf16 next: `sqrt.approx.f32` (relative error <= 2^-22, subnormals
// handled) is enough for the bf16 result to equal the correctly rounded one — an 8-bit significand's square root is either
// exactly representable in 8 bits or at least 2^-19 (relative) away from every 9-bit rounding boundary.  Checked over all
// 65 536 bf16 patterns by tests/test_gpu_optim.py::test_adamw_denominator_all_bf16_patterns (b2_adamw_denom_test).
__device__ __forceinline__ float sqrt_of_bf16(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// denom = exp_avg_sq.sqrt().add_(eps) on bf16 tensors: two bf16 roundings
__device__ __forceinline__ void adam_denom2(float v0, float v1, float eps, float& d0, float& d1) {
  float q0 = sqrt_of_bf16(v0), q1 = sqrt_of_bf16(v1);
  rn2_bf16(q0, q1);
  d0 = __fadd_rn(q0, eps);
  d1 = __fadd_rn(q1, eps);
  rn2_bf16(d0, d1);
}

// Two neighbouring elements at a time (so each pair of bf16 roundings is one F2FP).  `g` is already clipped.
// AW = the reference's operand order as written (see the header); r01 / r23 carry the four 16-bit random words of each
// element: r01 low -> exp_avg, r01 high -> shift (addcdiv), r23 low -> p, r23 high -> shift (residual).
template <bool AW>
__device__ __forceinline__ void adam_bf16_pair(float* p, const float* g, float* m, float* v, float* s, const AdamBF16P& a,
                                               const uint32_t* r01, const uint32_t* r23) {
  // exp_avg.mul_(beta1); add_stochastic_(exp_avg, grad, alpha=1-beta1)
  float m1a = __fmul_rn(m[0], a.b1), m1b = __fmul_rn(m[1], a.b1);
  rn2_bf16(m1a, m1b);
  // torch's add-with-alpha is a fused multiply-add (vec::fmadd on CPU, nvcc contraction on CUDA)
  const float mra = AW ? fmaf(a.omb1, m1a, g[0]) : fmaf(a.omb1, g[0], m1a);
  const float mrb = AW ? fmaf(a.omb1, m1b, g[1]) : fmaf(a.omb1, g[1], m1b);
  m[0] = sr_bf16(mra, r01[0] & 0xffffu);
  m[1] = sr_bf16(mrb, r01[1] & 0xffffu);
  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
  float v1a = __fmul_rn(v[0], a.b2), v1b = __fmul_rn(v[1], a.b2);
  rn2_bf16(v1a, v1b);
  float va = __fadd_rn(v1a, __fmul_rn(__fmul_rn(a.omb2, g[0]), g[0]));
  float vb = __fadd_rn(v1b, __fmul_rn(__fmul_rn(a.omb2, g[1]), g[1]));
  rn2_bf16(va, vb);
  v[0] = va;
  v[1] = vb;
  float dena, denb;
  adam_denom2(va, vb, a.eps, dena, denb);
  // addcdiv_stochastic_(shift, exp_avg, denom, value=-lr*denom_correction)
  const float s1a = sr_bf16(__fadd_rn(s[0], __fdiv_rn(__fmul_rn(a.step_size, m[0]), dena)), r01[0] >> 16);
  const float s1b = sr_bf16(__fadd_rn(s[1], __fdiv_rn(__fmul_rn(a.step_size, m[1]), denb)), r01[1] >> 16);
  // buffer = p.clone(); add_stochastic_(p, shift); add_stochastic_(shift, buffer.sub_(p))
  const float p1a = sr_bf16(__fadd_rn(s1a, p[0]), r23[0] & 0xffffu);
  const float p1b = sr_bf16(__fadd_rn(s1b, p[1]), r23[1] & 0xffffu);
  float da = __fsub_rn(p[0], p1a), db = __fsub_rn(p[1], p1b);
  rn2_bf16(da, db);
  s[0] = sr_bf16(__fadd_rn(da, s1a), r23[0] >> 16);
  s[1] = sr_bf16(__fadd_rn(db, s1b), r23[1] >> 16);
  p[0] = p1a;
  p[1] = p1b;
}

struct AdamVec { bf16x8 p, g, m, v, s; };

// FAST = the production path (hash random bits, operand order as written) with both choices compiled in; the generic
// instantiation keeps the test hooks (fixed / supplied random words, documented-intent order) on the SAME arithmetic.
template <bool FAST>
__global__ void __launch_bounds__(256, 2)
adamw_bf16_kernel(bf16* __restrict__ p, bf16* __restrict__ g, bf16* __restrict__ m, bf16* __restrict__ v,
                  bf16* __restrict__ sh, long long n, AdamBF16P a, const double* __restrict__ gnorm_sq,
                  const uint64_t* __restrict__ seed_offset, uint64_t step, const int32_t* __restrict__ test_rand16) {
  // a.zero_grad: the gradient vector is overwritten with zeros once it has been read (optimizer.zero_grad() folded in:
  // +2 B / parameter of writes here instead of a separate 5 GB fill)
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
  const bool aw = FAST ? true : a.as_written != 0;
  const int mode = FAST ? 0 : a.rng_mode;
  const long long nv = n >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // the next vector of each of the five arrays is loaded before this one is computed: five 16-byte loads per thread stay
  // in flight through the ~500-instruction compute phase, so HBM and the ALUs overlap inside one warp
  AdamVec cur;
  if (q < nv) { cur.p = ld8(p + q * 8); cur.g = ld8(g + q * 8); cur.m = ld8(m + q * 8); cur.v = ld8(v + q * 8); cur.s = ld8(sh + q * 8); }
  for (; q < nv; q += stride) {
    AdamVec nxt;
    const long long qn = q + stride;
    if (qn < nv) { nxt.p = ld8(p + qn * 8); nxt.g = ld8(g + qn * 8); nxt.m = ld8(m + qn * 8); nxt.v = ld8(v + qn * 8); nxt.s = ld8(sh + qn * 8); }
    float fp[8], fg[8], fm[8], fv[8], fs[8];
    unpack8(cur.p, fp);
    unpack8(cur.g, fg);
    unpack8(cur.m, fm);
    unpack8(cur.v, fv);
    unpack8(cur.s, fs);
    if (clip != 1.f) {  // clip_grad_norm_ scales the bf16 gradient in place (flow_matching_trainer.py:181-186): one rounding
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        fg[2 * h] = __fmul_rn(fg[2 * h], clip);
        fg[2 * h + 1] = __fmul_rn(fg[2 * h + 1], clip);
        rn2_bf16(fg[2 * h], fg[2 * h + 1]);
      }
    }
    const uint64_t e0 = (uint64_t)q * 8;
    const uint32_t ehi = (uint32_t)(e0 >> 32) * 0xC2B2AE35u;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      uint32_t r01[2], r23[2];
      if (mode == 0) {
        rand64((uint32_t)e0 + 2 * h, ehi, ka, kb, r01[0], r23[0]);
        rand64((uint32_t)e0 + 2 * h + 1, ehi, ka, kb, r01[1], r23[1]);
      } else if (mode == 3) {
        const long long e = q * 8 + 2 * h;
        r01[0] = (uint32_t)test_rand16[e] | ((uint32_t)test_rand16[n + e] << 16);
        r23[0] = (uint32_t)test_rand16[2 * n + e] | ((uint32_t)test_rand16[3 * n + e] << 16);
        r01[1] = (uint32_t)test_rand16[e + 1] | ((uint32_t)test_rand16[n + e + 1] << 16);
        r23[1] = (uint32_t)test_rand16[2 * n + e + 1] | ((uint32_t)test_rand16[3 * n + e + 1] << 16);
      } else {
        const uint32_t f = mode == 1 ? 0u : 0xffffffffu;
        r01[0] = r01[1] = r23[0] = r23[1] = f;
      }
      if (aw) adam_bf16_pair<true>(fp + 2 * h, fg + 2 * h, fm + 2 * h, fv + 2 * h, fs + 2 * h, a, r01, r23);
      else adam_bf16_pair<false>(fp + 2 * h, fg + 2 * h, fm + 2 * h, fv + 2 * h, fs + 2 * h, a, r01, r23);
    }
    // values are exact bf16 (low 16 bits zero): packing is a truncation, not a second rounding
    st8(p + q * 8, pack8(fp));
    st8(m + q * 8, pack8(fm));
    st8(v + q * 8, pack8(fv));
    st8(sh + q * 8, pack8(fs));
    if (a.zero_grad) st8(g + q * 8, zero8());
    cur = nxt;
  }
  // tail (n % 8 elements), one thread
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (long long i = nv * 8; i < n; ++i) {
      uint32_t r01[2] = {0, 0}, r23[2] = {0, 0};
      if (mode == 0) {
        rand64((uint32_t)i, (uint32_t)((uint64_t)i >> 32) * 0xC2B2AE35u, ka, kb, r01[0], r23[0]);
      } else if (mode == 3) {
        r01[0] = (uint32_t)test_rand16[i] | ((uint32_t)test_rand16[n + i] << 16);
        r23[0] = (uint32_t)test_rand16[2 * n + i] | ((uint32_t)test_rand16[3 * n + i] << 16);
      } else {
        r01[0] = r23[0] = mode == 1 ? 0u : 0xffffffffu;
      }
      float fp[2] = {__bfloat162float(p[i]), 0.f}, fm[2] = {__bfloat162float(m[i]), 0.f},
            fv[2] = {__bfloat162float(v[i]), 0.f}, fs[2] = {__bfloat162float(sh[i]), 0.f};
      float fg[2] = {__bfloat162float(g[i]), 0.f};
      if (clip != 1.f) fg[0] = rn_bf16(__fmul_rn(fg[0], clip));
      if (aw) adam_bf16_pair<true>(fp, fg, fm, fv, fs, a, r01, r23);
      else adam_bf16_pair<false>(fp, fg, fm, fv, fs, a, r01, r23);
      p[i] = __float2bfloat16_rn(fp[0]);
      m[i] = __float2bfloat16_rn(fm[0]);
      v[i] = __float2bfloat16_rn(fv[0]);
      sh[i] = __float2bfloat16_rn(fs[0]);
      if (a.zero_grad) g[i] = __float2bfloat16_rn(0.f);
    }
  }
}

// test hook: the denominator function above next to its IEEE statement, element-wise over a bf16 array
__global__ void adamw_denom_test_kernel(const bf16* __restrict__ v, bf16* __restrict__ fast, bf16* __restrict__ ieee, int n,
                                        float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = __bfloat162float(v[i]);
  float d0, d1;
  adam_denom2(x, x, eps, d0, d1);
  fast[i] = __float2bfloat16_rn(d0);
  ieee[i] = __float2bfloat16_rn(rn_bf16(__fadd_rn(rn_bf16(__fsqrt_rn(x)), eps)));
}

// y <- bf16(y + alpha * x): the deferred weight decay `shift.add_(p, alpha=-decay)` (adamw_bfloat16/__init__.py:191-192)
__global__ void axpy_bf16_kernel(bf16* __restrict__ y, const bf16* __restrict__ x, long long n, float alpha) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16_rn(__fadd_rn(__bfloat162float(y[i]), __fmul_rn(alpha, __bfloat162float(x[i]))));
}

}  // namespace b2

using namespace b2;

extern "C" int b2_adamw_bf16(void* p, void* g, void* m, void* v, void* shift, int64_t n, double lr, double beta1,
                             double beta2, double eps, int step, const double* gnorm_sq, float max_norm, float grad_scale,
                             const uint64_t* seed_offset, int as_written, int rng_mode, const int32_t* test_rand16,
                             int zero_grad, void* stream) {
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
  a.as_written = as_written; a.rng_mode = rng_mode; a.zero_grad = zero_grad;
  long long blocks = ((n >> 3) + 255) / 256;
  const long long cap = 2LL * num_sms();  // resident CTAs only: the prefetch pipeline runs across grid-stride iterations
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  auto kern = (rng_mode == 0 && as_written) ? adamw_bf16_kernel<true> : adamw_bf16_kernel<false>;
  kern<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((bf16*)p, (bf16*)g, (bf16*)m, (bf16*)v, (bf16*)shift, n, a,
                                                           gnorm_sq, seed_offset, (uint64_t)step, test_rand16);
  return check_launch("adamw_bf16");
}

extern "C" int b2_adamw_denom_test(const void* v, void* fast, void* ieee, int n, float eps, void* stream) {
  B2_REQUIRE(v && fast && ieee && n > 0, "b2_adamw_denom_test: bad args");
  adamw_denom_test_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const bf16*)v, (bf16*)fast, (bf16*)ieee, n, eps);
  return check_launch("adamw_denom_test");
}

extern "C" int b2_axpy_bf16(void* y, const void* x, int64_t n, float alpha, void* stream) {
  B2_REQUIRE(y && x && n > 0, "b2_axpy_bf16: bad args");
  long long blocks = (n + 255) / 256;
  const long long cap = 8LL * num_sms();
  if (blocks > cap) blocks = cap;
  axpy_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((bf16*)y, (const bf16*)x, n, alpha);
  return check_launch("axpy_bf16");
}
