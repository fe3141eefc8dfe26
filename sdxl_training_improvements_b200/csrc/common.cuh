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

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }
__device__ __forceinline__ float dsilu_f(float x) {
  float s = 1.f / (1.f + __expf(-x));
  return s * (1.f + x * (1.f - s));
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
