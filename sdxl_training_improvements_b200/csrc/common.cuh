// Shared helpers for libsdxl_b200: error reporting, launch accounting, small device utilities.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sdxl_b200.h"

typedef __nv_bfloat16 bf16;

namespace b2 {

void set_error(const char* fmt, ...);
void count_launch();

inline int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: %s", what, cudaGetErrorString(e));
    return B2_ERR_CUDA;
  }
  count_launch();
  return B2_OK;
}

#define B2_REQUIRE(cond, ...)          \
  do {                                 \
    if (!(cond)) {                     \
      b2::set_error(__VA_ARGS__);      \
      return B2_ERR_ARG;               \
    }                                  \
  } while (0)

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 8 bf16 <-> 8 float through ONE 16-byte vector access (LDG.128 / STG.128; a struct of four bfloat162 members is
// copied member-wise by nvcc and degrades to 4 x 32-bit accesses, which costs 4x the LSU wavefronts and L2 sectors).
struct __align__(16) bf16x8 {
  uint4 u;
};
__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
  const uint32_t w[4] = {p.u.x, p.u.y, p.u.z, p.u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 p;
  p.u = make_uint4(pack2_bf16(f[0], f[1]), pack2_bf16(f[2], f[3]), pack2_bf16(f[4], f[5]), pack2_bf16(f[6], f[7]));
  return p;
}
__device__ __forceinline__ bf16x8 zero8() {
  bf16x8 p;
  p.u = make_uint4(0u, 0u, 0u, 0u);
  return p;
}
__device__ __forceinline__ bf16x8 ld8(const bf16* p) {
  bf16x8 r;
  r.u = *reinterpret_cast<const uint4*>(p);
  return r;
}
__device__ __forceinline__ void st8(bf16* p, const bf16x8& v) { *reinterpret_cast<uint4*>(p) = v.u; }

// fast division (MUFU.RCP + multiply, 2 ulp): the IEEE divide is ~10 instructions in kernels that sit next to their HBM bound
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.f + __expf(-x)); }
__device__ __forceinline__ float dsilu_f(float x) {
  float s = __fdividef(1.f, 1.f + __expf(-x));
  return s * (1.f + x * (1.f - s));
}

// Exact-erf GELU (diffusers GEGLU uses F.gelu, not the tanh form).  erff() costs ~25 instructions and made the GEGLU
// kernels ALU-bound (41 / 53 us for 126 / 210 MB); Abramowitz & Stegun 7.1.26 (|abs error| <= 1.5e-7, far below one bf16
// ulp of the product) needs one reciprocal, five FMAs and ONE exponential — exp(-x^2/2) — which is also the Gaussian
// density the derivative needs:   Phi(x) = (1 + erf(x/sqrt2)) / 2,   gelu' = Phi + x * phi.
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& pdf) {
  const float ax = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, ax, 1.f));
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-ax * ax * 1.4426950408889634f));  // exp(-x^2 / 2)
  const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f),
                              0.254829592f);
  const float erf_abs = fmaf(-poly, e, 1.f);
  cdf = 0.5f * (1.f + copysignf(erf_abs, x));
  pdf = 0.3989422804014327f * e;
}
// Eight elements stage by stage (same operations per element as gelu_parts, bit-identical): eight independent reciprocals,
// exponentials and polynomial chains next to each other — for code where ONE warp per scheduler has to hide its own latencies
// (GEMM epilogues), where the element-at-a-time form ran at ~90 cycles per element.
__device__ __forceinline__ void gelu_parts8(const float* x, float* cdf, float* pdf) {
  float ax[8], t[8], e[8], poly[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) ax[j] = fabsf(x[j]) * 0.70710678118654752f;
#pragma unroll
  for (int j = 0; j < 8; ++j) t[j] = __fdividef(1.f, fmaf(0.3275911f, ax[j], 1.f));
#pragma unroll
  for (int j = 0; j < 8; ++j) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[j]) : "f"(-ax[j] * ax[j] * 1.4426950408889634f));
#pragma unroll
  for (int j = 0; j < 8; ++j)
    poly[j] = t[j] * fmaf(t[j], fmaf(t[j], fmaf(t[j], fmaf(t[j], 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f),
                          0.254829592f);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float erf_abs = fmaf(-poly[j], e[j], 1.f);
    cdf[j] = 0.5f * (1.f + copysignf(erf_abs, x[j]));
    pdf[j] = 0.3989422804014327f * e[j];
  }
}
__device__ __forceinline__ float gelu_erf(float x) {
  float c, d;
  gelu_parts(x, c, d);
  return x * c;
}
__device__ __forceinline__ float dgelu_erf(float x) {
  float c, d;
  gelu_parts(x, c, d);
  return fmaf(x, d, c);
}

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace b2
