// GroupNorm(+SiLU) and LayerNorm, forward and backward, token-major (NHWC) bf16 activations, fp32 math.
//
// HBM-bound kernels.  Thread mapping for GroupNorm: a CTA owns a chunk of rows of one sample; thread t owns the
// fixed 16-byte channel vector (t % nvec) and walks rows (t / nvec), (t / nvec) + rpi, ... so every row is read
// as one fully coalesced burst and per-channel partial sums live in registers.  Partials are combined through
// shared memory, per-group totals go out as one double atomicAdd per (CTA, group).
#include <cstdlib>

#include "common.cuh"

namespace b2 {

static inline void gn_geometry(int B, int HW, int C, int* threads, int* rpi, int* rows_per_cta, int* chunks) {
  const int nvec = C / 8;
  int r = 512 / nvec;
  if (r < 1) r = 1;
  if (r > HW) r = HW;
  *rpi = r;
  *threads = nvec * r;
  // aim for ~4 CTAs per SM in total
  long long want_ctas = 4LL * num_sms();
  long long per = ((long long)B * HW + want_ctas - 1) / want_ctas;
  if (per < (long long)r * 4) per = (long long)r * 4;
  if (per > HW) per = HW;
  *rows_per_cta = (int)per;
  *chunks = (HW + *rows_per_cta - 1) / *rows_per_cta;
}

__global__ void gn_stats_kernel(const bf16* __restrict__ x, int HW, int C, int G, int rows_per_cta, double* ws) {
  extern __shared__ float sm[];
  const int nvec = C >> 3;
  const int v = threadIdx.x % nvec, slot = threadIdx.x / nvec, rpi = blockDim.x / nvec;
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(HW, r0 + rows_per_cta);
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
  const bf16* xb = x + (size_t)b * HW * C + v * 8;
  int r = r0 + slot;
  // four rows per trip with all four 16-byte loads issued before the first use: the plain loop keeps ONE load in flight
  // per thread (SASS: each LDG feeds the next instruction), which leaves these HBM-bound kernels latency-bound
  for (; r + 3 * rpi < r1; r += 4 * rpi) {
    bf16x8 t[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) t[u] = ld8(xb + (size_t)(r + u * rpi) * C);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
      unpack8(t[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += f[j];
        ss[j] += f[j] * f[j];
      }
    }
  }
  for (; r < r1; r += rpi) {
    float f[8];
    unpack8(ld8(xb + (size_t)r * C), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] += f[j];
      ss[j] += f[j] * f[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sm[(slot * 2 + 0) * C + v * 8 + j] = s[j];
    sm[(slot * 2 + 1) * C + v * 8 + j] = ss[j];
  }
  __syncthreads();
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double a = 0.0, q = 0.0;
    for (int sl = 0; sl < rpi; ++sl)
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        a += sm[(sl * 2 + 0) * C + c];
        q += sm[(sl * 2 + 1) * C + c];
      }
    atomicAdd(&ws[((size_t)b * G + g) * 2 + 0], a);
    atomicAdd(&ws[((size_t)b * G + g) * 2 + 1], q);
  }
}

__global__ void gn_finalize_kernel(const double* ws, float* mean, float* rstd, int n, double count, float eps) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double m = ws[2 * i] / count;
  double var = ws[2 * i + 1] / count - m * m;
  if (var < 0) var = 0;
  mean[i] = (float)m;
  rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
}

__global__ void gn_apply_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, const float* __restrict__ mean,
                                const float* __restrict__ rstd, const bf16* __restrict__ gamma,
                                const bf16* __restrict__ beta, int HW, int C, int G, int rows_per_cta, int silu) {
  const int nvec = C >> 3;
  const int v = threadIdx.x % nvec, slot = threadIdx.x / nvec, rpi = blockDim.x / nvec;
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(HW, r0 + rows_per_cta);
  const int cpg = C / G;
  float a[8], bb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = v * 8 + j;
    const int g = c / cpg;
    const float m = mean[b * G + g], rs = rstd[b * G + g];
    a[j] = rs * __bfloat162float(gamma[c]);
    bb[j] = __bfloat162float(beta[c]) - m * a[j];
  }
  const size_t base = (size_t)b * HW * C + v * 8;
  int r = r0 + slot;
  for (; r + 3 * rpi < r1; r += 4 * rpi) {  // four loads in flight per thread (see gn_stats_kernel)
    bf16x8 t4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) t4[u] = ld8(x + base + (size_t)(r + u * rpi) * C);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
      unpack8(t4[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = f[j] * a[j] + bb[j];
        f[j] = silu ? silu_f(t) : t;
      }
      st8(y + base + (size_t)(r + u * rpi) * C, pack8(f));
    }
  }
  for (; r < r1; r += rpi) {
    float f[8];
    unpack8(ld8(x + base + (size_t)r * C), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = f[j] * a[j] + bb[j];
      f[j] = silu ? silu_f(t) : t;
    }
    st8(y + base + (size_t)r * C, pack8(f));
  }
}

// backward pass 1: per-(sample, channel) sums of g and g*xhat  (g = dL/d(gn output), through SiLU if fused)
__global__ void gn_bwd_reduce_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                     const bf16* __restrict__ gamma, const bf16* __restrict__ beta, int HW, int C, int G,
                                     int rows_per_cta, int silu, double* ws, float* dgb) {
  extern __shared__ float sm[];
  const int nvec = C >> 3;
  const int v = threadIdx.x % nvec, slot = threadIdx.x / nvec, rpi = blockDim.x / nvec;
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(HW, r0 + rows_per_cta);
  const int cpg = C / G;
  float mu[8], rs[8], ga[8], be[8], dg[8], db[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = v * 8 + j;
    const int g = c / cpg;
    mu[j] = mean[b * G + g];
    rs[j] = rstd[b * G + g];
    ga[j] = __bfloat162float(gamma[c]);
    be[j] = __bfloat162float(beta[c]);
    dg[j] = db[j] = 0.f;
  }
  const size_t base = (size_t)b * HW * C + v * 8;
  int r = r0 + slot;
  for (; r + rpi < r1; r += 2 * rpi) {  // two rows = four loads in flight per thread (see gn_stats_kernel)
    bf16x8 tx[2], td[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      tx[u] = ld8(x + base + (size_t)(r + u * rpi) * C);
      td[u] = ld8(dy + base + (size_t)(r + u * rpi) * C);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float f[8], d[8];
      unpack8(tx[u], f);
      unpack8(td[u], d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (f[j] - mu[j]) * rs[j];
        float g = d[j];
        if (silu) g *= dsilu_f(xh * ga[j] + be[j]);
        dg[j] += g * xh;
        db[j] += g;
      }
    }
  }
  for (; r < r1; r += rpi) {
    float f[8], d[8];
    unpack8(ld8(x + base + (size_t)r * C), f);
    unpack8(ld8(dy + base + (size_t)r * C), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (f[j] - mu[j]) * rs[j];
      float g = d[j];
      if (silu) g *= dsilu_f(xh * ga[j] + be[j]);
      dg[j] += g * xh;
      db[j] += g;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sm[(slot * 2 + 0) * C + v * 8 + j] = dg[j];
    sm[(slot * 2 + 1) * C + v * 8 + j] = db[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float tg = 0.f, tb = 0.f;
    for (int sl = 0; sl < rpi; ++sl) {
      tg += sm[(sl * 2 + 0) * C + c];
      tb += sm[(sl * 2 + 1) * C + c];
    }
    sm[c] = tg;        // slot 0, plane 0
    sm[C + c] = tb;    // slot 0, plane 1
    atomicAdd(&dgb[c], tg);
    atomicAdd(&dgb[C + c], tb);
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double s1 = 0.0, s2 = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      const float gm = __bfloat162float(gamma[c]);
      s1 += (double)(gm * sm[C + c]);
      s2 += (double)(gm * sm[c]);
    }
    atomicAdd(&ws[((size_t)b * G + g) * 2 + 0], s1);
    atomicAdd(&ws[((size_t)b * G + g) * 2 + 1], s2);
  }
}

__global__ void gn_bwd_dx_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, bf16* __restrict__ dx,
                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                 const bf16* __restrict__ gamma, const bf16* __restrict__ beta, int HW, int C, int G,
                                 int rows_per_cta, int silu, const double* __restrict__ ws, int accumulate) {
  const int nvec = C >> 3;
  const int v = threadIdx.x % nvec, slot = threadIdx.x / nvec, rpi = blockDim.x / nvec;
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(HW, r0 + rows_per_cta);
  const int cpg = C / G;
  const float invn = 1.f / ((float)cpg * (float)HW);
  float mu[8], rs[8], ga[8], be[8], m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = v * 8 + j;
    const int g = c / cpg;
    mu[j] = mean[b * G + g];
    rs[j] = rstd[b * G + g];
    ga[j] = __bfloat162float(gamma[c]);
    be[j] = __bfloat162float(beta[c]);
    m1[j] = (float)(ws[((size_t)b * G + g) * 2 + 0]) * invn;
    m2[j] = (float)(ws[((size_t)b * G + g) * 2 + 1]) * invn;
  }
  const size_t base = (size_t)b * HW * C + v * 8;
  int r = r0 + slot;
  for (; r + rpi < r1; r += 2 * rpi) {  // two rows = four to six loads in flight per thread (see gn_stats_kernel)
    bf16x8 tx[2], td[2], to[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      tx[u] = ld8(x + base + (size_t)(r + u * rpi) * C);
      td[u] = ld8(dy + base + (size_t)(r + u * rpi) * C);
      if (accumulate) to[u] = ld8(dx + base + (size_t)(r + u * rpi) * C);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float f[8], d[8], o[8];
      unpack8(tx[u], f);
      unpack8(td[u], d);
      if (accumulate) unpack8(to[u], o);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (f[j] - mu[j]) * rs[j];
        float g = d[j];
        if (silu) g *= dsilu_f(xh * ga[j] + be[j]);
        const float r_ = rs[j] * (g * ga[j] - m1[j] - xh * m2[j]);
        f[j] = accumulate ? o[j] + r_ : r_;
      }
      st8(dx + base + (size_t)(r + u * rpi) * C, pack8(f));
    }
  }
  for (; r < r1; r += rpi) {
    float f[8], d[8], o[8];
    unpack8(ld8(x + base + (size_t)r * C), f);
    unpack8(ld8(dy + base + (size_t)r * C), d);
    if (accumulate) unpack8(ld8(dx + base + (size_t)r * C), o);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (f[j] - mu[j]) * rs[j];
      float g = d[j];
      if (silu) g *= dsilu_f(xh * ga[j] + be[j]);
      const float r_ = rs[j] * (g * ga[j] - m1[j] - xh * m2[j]);
      f[j] = accumulate ? o[j] + r_ : r_;
    }
    st8(dx + base + (size_t)r * C, pack8(f));
  }
}

// ------------------------------- LayerNorm -------------------------------------------------
__global__ void ln_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, const bf16* __restrict__ gamma,
                              const bf16* __restrict__ beta, float* __restrict__ mean, float* __restrict__ rstd, int M,
                              int C, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= M) return;
  const int nvec = C >> 3;
  const bf16* xr = x + (size_t)row * C;
  float s = 0.f;
  for (int v = lane; v < nvec; v += 32) {
    float f[8];
    unpack8(ld8(xr + v * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += f[j];
  }
  s = warp_sum(s);
  const float mu = s / (float)C;
  float q = 0.f;
  for (int v = lane; v < nvec; v += 32) {
    float f[8];
    unpack8(ld8(xr + v * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = f[j] - mu;
      q += d * d;
    }
  }
  q = warp_sum(q);
  const float rs = rsqrtf(q / (float)C + eps);
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
  bf16* yr = y + (size_t)row * C;
  for (int v = lane; v < nvec; v += 32) {
    float f[8], g[8], b[8];
    unpack8(ld8(xr + v * 8), f);
    unpack8(ld8(gamma + v * 8), g);
    unpack8(ld8(beta + v * 8), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (f[j] - mu) * rs * g[j] + b[j];
    st8(yr + v * 8, pack8(f));
  }
}

__global__ void ln_bwd_dx_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, bf16* __restrict__ dx,
                                 const bf16* __restrict__ gamma, const float* __restrict__ mean,
                                 const float* __restrict__ rstd, int M, int C, int accumulate) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= M) return;
  const int nvec = C >> 3;
  const bf16* xr = x + (size_t)row * C;
  const bf16* dr = dy + (size_t)row * C;
  const float mu = mean[row], rs = rstd[row];
  float c1 = 0.f, c2 = 0.f;
  for (int v = lane; v < nvec; v += 32) {
    float f[8], d[8], g[8];
    unpack8(ld8(xr + v * 8), f);
    unpack8(ld8(dr + v * 8), d);
    unpack8(ld8(gamma + v * 8), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float dxh = d[j] * g[j];
      c1 += dxh;
      c2 += dxh * (f[j] - mu) * rs;
    }
  }
  c1 = warp_sum(c1) / (float)C;
  c2 = warp_sum(c2) / (float)C;
  bf16* oxr = dx + (size_t)row * C;
  for (int v = lane; v < nvec; v += 32) {
    float f[8], d[8], g[8], o[8];
    unpack8(ld8(xr + v * 8), f);
    unpack8(ld8(dr + v * 8), d);
    unpack8(ld8(gamma + v * 8), g);
    if (accumulate) unpack8(ld8(oxr + v * 8), o);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (f[j] - mu) * rs;
      const float r_ = rs * (d[j] * g[j] - c1 - xh * c2);
      f[j] = accumulate ? o[j] + r_ : r_;
    }
    st8(oxr + v * 8, pack8(f));
  }
}

// Register-resident LayerNorm rows: one warp per row, each lane keeps its NV 16-byte vectors of the row (C <= 256 * NV)
// in registers, so x (and dy) are read from memory exactly once; the generic kernels above re-read the row per pass.
template <int NV>
__global__ void __launch_bounds__(256)
ln_fwd_reg_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, const bf16* __restrict__ gamma,
                  const bf16* __restrict__ beta, float* __restrict__ mean, float* __restrict__ rstd, int M, int C,
                  float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= M) return;
  const int nvec = C >> 3;
  const bf16* xr = x + (size_t)row * C;
  float f[NV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 32 * i;
    if (v < nvec) {
      unpack8(ld8(xr + v * 8), f[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[i][j];
    }
  }
  s = warp_sum(s);
  const float mu = s / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
    if (lane + 32 * i < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = f[i][j] - mu;
        q += d * d;
      }
    }
  q = warp_sum(q);
  const float rs = rsqrtf(q / (float)C + eps);
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
  bf16* yr = y + (size_t)row * C;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 32 * i;
    if (v < nvec) {
      float g[8], b[8];
      unpack8(ld8(gamma + v * 8), g);
      unpack8(ld8(beta + v * 8), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[i][j] = (f[i][j] - mu) * rs * g[j] + b[j];
      st8(yr + v * 8, pack8(f[i]));
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(256)
ln_bwd_dx_reg_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, bf16* __restrict__ dx,
                     const bf16* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd, int M,
                     int C, int accumulate) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= M) return;
  const int nvec = C >> 3;
  const bf16* xr = x + (size_t)row * C;
  const bf16* dr = dy + (size_t)row * C;
  const float mu = mean[row], rs = rstd[row];
  float xh[NV][8], dg[NV][8];
  float c1 = 0.f, c2 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 32 * i;
    if (v < nvec) {
      float d[8], g[8];
      unpack8(ld8(xr + v * 8), xh[i]);
      unpack8(ld8(dr + v * 8), d);
      unpack8(ld8(gamma + v * 8), g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[i][j] = (xh[i][j] - mu) * rs;
        dg[i][j] = d[j] * g[j];
        c1 += dg[i][j];
        c2 += dg[i][j] * xh[i][j];
      }
    }
  }
  c1 = warp_sum(c1) / (float)C;
  c2 = warp_sum(c2) / (float)C;
  bf16* oxr = dx + (size_t)row * C;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 32 * i;
    if (v < nvec) {
      float o[8];
      if (accumulate) unpack8(ld8(oxr + v * 8), o);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float r_ = rs * (dg[i][j] - c1 - xh[i][j] * c2);
        o[j] = accumulate ? o[j] + r_ : r_;
      }
      st8(oxr + v * 8, pack8(o));
    }
  }
}

// LayerNorm backward in ONE pass over x / dy: dx (register-resident rows as above) and the column sums dgamma / dbeta,
// which the two-kernel version re-read x and dy for.  A warp walks rows_per_cta / 8 rows and keeps its share of the column
// sums in registers (2 x NV x 8 floats); the 8 warps of a CTA combine through shared-memory atomics and the CTA adds 2C
// values to the fp32 staging buffer.  One CTA per SM-slot (grid sized to a single wave).
template <int NV>
__global__ void __launch_bounds__(256)
ln_bwd_fused_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, bf16* __restrict__ dx,
                    const bf16* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                    float* __restrict__ dgb, int M, int C, int accumulate, int rows_per_cta) {
  __shared__ float sred[2 * NV * 256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = C >> 3;
  for (int i = threadIdx.x; i < 2 * C; i += 256) sred[i] = 0.f;
  __syncthreads();
  float ag[NV][8], ab[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) ag[i][j] = ab[i][j] = 0.f;
  const int r_end = min(M, (blockIdx.x + 1) * rows_per_cta);
  for (int row = blockIdx.x * rows_per_cta + warp; row < r_end; row += 8) {
    const bf16* xr = x + (size_t)row * C;
    const bf16* dr = dy + (size_t)row * C;
    const float mu = mean[row], rs = rstd[row];
    float xh[NV][8], dg[NV][8];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        float d[8], g[8];
        unpack8(ld8(xr + v * 8), xh[i]);
        unpack8(ld8(dr + v * 8), d);
        unpack8(ld8(gamma + v * 8), g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] = (xh[i][j] - mu) * rs;
          ag[i][j] += d[j] * xh[i][j];
          ab[i][j] += d[j];
          dg[i][j] = d[j] * g[j];
          c1 += dg[i][j];
          c2 += dg[i][j] * xh[i][j];
        }
      }
    }
    c1 = warp_sum(c1) / (float)C;
    c2 = warp_sum(c2) / (float)C;
    bf16* oxr = dx + (size_t)row * C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        float o[8];
        if (accumulate) unpack8(ld8(oxr + v * 8), o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float r_ = rs * (dg[i][j] - c1 - xh[i][j] * c2);
          o[j] = accumulate ? o[j] + r_ : r_;
        }
        st8(oxr + v * 8, pack8(o));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 32 * i;
    if (v < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&sred[v * 8 + j], ag[i][j]);
        atomicAdd(&sred[C + v * 8 + j], ab[i][j]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) atomicAdd(&dgb[i], sred[i]);
}

// dgamma[c] += sum_m dy*xhat, dbeta[c] += sum_m dy.  CTA = 64 columns x a chunk of rows; 8 column-vectors x 32 rows.
__global__ void ln_bwd_dgb_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, const float* __restrict__ mean,
                                  const float* __restrict__ rstd, float* dgb, int M, int C, int rows_per_cta) {
  __shared__ float sm[2][32][65];
  const int cv = threadIdx.x & 7, rl = threadIdx.x >> 3;  // 256 threads
  const int c0 = blockIdx.x * 64 + cv * 8;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(M, r0 + rows_per_cta);
  float dg[8], db[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) dg[j] = db[j] = 0.f;
  if (c0 < C) {
    for (int r = r0 + rl; r < r1; r += 32) {
      float f[8], d[8];
      unpack8(ld8(x + (size_t)r * C + c0), f);
      unpack8(ld8(dy + (size_t)r * C + c0), d);
      const float mu = mean[r], rs = rstd[r];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dg[j] += d[j] * (f[j] - mu) * rs;
        db[j] += d[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sm[0][rl][cv * 8 + j] = dg[j];
    sm[1][rl][cv * 8 + j] = db[j];
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
    float t = 0.f;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) t += sm[which][r][c];
    const int gc = blockIdx.x * 64 + c;
    if (gc < C) atomicAdd(&dgb[which * C + gc], t);
  }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_gn_stats(const void* x, int B, int HW, int C, int G, float eps, void* ws, float* mean, float* rstd,
                           void* stream) {
  B2_REQUIRE(x && ws && mean && rstd, "b2_gn_stats: null pointer");
  B2_REQUIRE(C % 8 == 0 && C % G == 0 && C / 8 <= 1024, "b2_gn_stats: unsupported C=%d G=%d", C, G);
  cudaStream_t st = (cudaStream_t)stream;
  int threads, rpi, rpc, chunks;
  gn_geometry(B, HW, C, &threads, &rpi, &rpc, &chunks);
  cudaMemsetAsync(ws, 0, sizeof(double) * 2 * B * G, st);
  const size_t smem = (size_t)rpi * 2 * C * sizeof(float);
  gn_stats_kernel<<<dim3(chunks, B), threads, smem, st>>>((const bf16*)x, HW, C, G, rpc, (double*)ws);
  int rc = check_launch("gn_stats");
  if (rc) return rc;
  gn_finalize_kernel<<<(B * G + 127) / 128, 128, 0, st>>>((const double*)ws, mean, rstd, B * G,
                                                          (double)(C / G) * (double)HW, eps);
  return check_launch("gn_finalize");
}

extern "C" int b2_gn_apply(const void* x, void* y, const float* mean, const float* rstd, const void* gamma,
                           const void* beta, int B, int HW, int C, int G, int silu, void* stream) {
  B2_REQUIRE(x && y && mean && rstd && gamma && beta, "b2_gn_apply: null pointer");
  B2_REQUIRE(C % 8 == 0 && C % G == 0 && C / 8 <= 1024, "b2_gn_apply: unsupported C=%d G=%d", C, G);
  int threads, rpi, rpc, chunks;
  gn_geometry(B, HW, C, &threads, &rpi, &rpc, &chunks);
  gn_apply_kernel<<<dim3(chunks, B), threads, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, mean, rstd,
                                                                        (const bf16*)gamma, (const bf16*)beta, HW, C, G,
                                                                        rpc, silu);
  return check_launch("gn_apply");
}

extern "C" int b2_gn_bwd(const void* x, const void* dy, void* dx, const float* mean, const float* rstd,
                         const void* gamma, const void* beta, int B, int HW, int C, int G, int silu, void* ws,
                         float* dgb, int accumulate_dx, void* stream) {
  B2_REQUIRE(x && dy && dx && mean && rstd && gamma && beta && ws && dgb, "b2_gn_bwd: null pointer");
  B2_REQUIRE(C % 8 == 0 && C % G == 0 && C / 8 <= 1024, "b2_gn_bwd: unsupported C=%d G=%d", C, G);
  cudaStream_t st = (cudaStream_t)stream;
  int threads, rpi, rpc, chunks;
  gn_geometry(B, HW, C, &threads, &rpi, &rpc, &chunks);
  cudaMemsetAsync(ws, 0, sizeof(double) * 2 * B * G, st);
  const size_t smem = (size_t)rpi * 2 * C * sizeof(float);
  gn_bwd_reduce_kernel<<<dim3(chunks, B), threads, smem, st>>>((const bf16*)x, (const bf16*)dy, mean, rstd,
                                                               (const bf16*)gamma, (const bf16*)beta, HW, C, G, rpc,
                                                               silu, (double*)ws, dgb);
  int rc = check_launch("gn_bwd_reduce");
  if (rc) return rc;
  gn_bwd_dx_kernel<<<dim3(chunks, B), threads, 0, st>>>((const bf16*)x, (const bf16*)dy, (bf16*)dx, mean, rstd,
                                                        (const bf16*)gamma, (const bf16*)beta, HW, C, G, rpc, silu,
                                                        (const double*)ws, accumulate_dx);
  return check_launch("gn_bwd_dx");
}

extern "C" int b2_ln_fwd(const void* x, void* y, const void* gamma, const void* beta, float* mean, float* rstd, int M,
                         int C, float eps, void* stream) {
  B2_REQUIRE(x && y && gamma && beta && mean && rstd, "b2_ln_fwd: null pointer");
  B2_REQUIRE(C % 8 == 0, "b2_ln_fwd: C %% 8 != 0");
  cudaStream_t st = (cudaStream_t)stream;
  const int nv = (C / 8 + 31) / 32;
#define B2_LN_FWD(NV)                                                                                              \
  ln_fwd_reg_kernel<NV><<<(M + 7) / 8, 256, 0, st>>>((const bf16*)x, (bf16*)y, (const bf16*)gamma, (const bf16*)beta, \
                                                     mean, rstd, M, C, eps)
  if (nv <= 2) B2_LN_FWD(2);
  else if (nv <= 3) B2_LN_FWD(3);
  else if (nv <= 5) B2_LN_FWD(5);
  else if (nv <= 8) B2_LN_FWD(8);
  else
    ln_fwd_kernel<<<(M + 7) / 8, 256, 0, st>>>((const bf16*)x, (bf16*)y, (const bf16*)gamma, (const bf16*)beta, mean, rstd,
                                               M, C, eps);
#undef B2_LN_FWD
  return check_launch("ln_fwd");
}

extern "C" int b2_ln_bwd_parts(const void* x, const void* dy, void* dx, const void* gamma, const float* mean,
                               const float* rstd, float* dgb, int M, int C, int accumulate_dx, int parts, void* stream);
extern "C" int b2_ln_bwd(const void* x, const void* dy, void* dx, const void* gamma, const float* mean,
                         const float* rstd, float* dgb, int M, int C, int accumulate_dx, void* stream) {
  return b2_ln_bwd_parts(x, dy, dx, gamma, mean, rstd, dgb, M, C, accumulate_dx, 3, stream);
}

// parts: 1 = dx only, 2 = dgamma / dbeta only, 3 = both.  The two halves are independent kernels over the same x / dy: the
// host issues them on two streams so that they share their reads in L2 instead of running back to back (unet.py).
extern "C" int b2_ln_bwd_parts(const void* x, const void* dy, void* dx, const void* gamma, const float* mean,
                               const float* rstd, float* dgb, int M, int C, int accumulate_dx, int parts, void* stream) {
  B2_REQUIRE(x && dy && gamma && mean && rstd && (dx || !(parts & 1)) && (dgb || !(parts & 2)) && (parts & 3),
             "b2_ln_bwd: null pointer");
  B2_REQUIRE(C % 8 == 0, "b2_ln_bwd: C %% 8 != 0");
  cudaStream_t st = (cudaStream_t)stream;
  const int nv = (C / 8 + 31) / 32;
  // Opt-in (B2_LN_BWD_FUSED=1): one fused pass, one wave of CTAs.  Measured SLOWER inside the training step than the two
  // kernels below (130.9 vs 126.5 ms per step, profiles/r1_bench_n1_v8_notes.txt): at C = 1280 the column sums cost 80
  // more registers per thread, occupancy drops to 8 warps per SM and the row loop becomes latency-bound.
  const bool fused = getenv("B2_LN_BWD_FUSED") != nullptr && parts == 3;
  if (nv <= 5 && fused) {
    const int per_sm = nv <= 3 ? 2 : 1;
    const int slots = num_sms() * per_sm;
    int rows_per_cta = ((M + slots - 1) / slots + 7) / 8 * 8;
    if (rows_per_cta < 8) rows_per_cta = 8;
    const int ctas = (M + rows_per_cta - 1) / rows_per_cta;
#define B2_LN_BWD_F(NV)                                                                                          \
  ln_bwd_fused_kernel<NV><<<ctas, 256, 0, st>>>((const bf16*)x, (const bf16*)dy, (bf16*)dx, (const bf16*)gamma, mean, \
                                                rstd, dgb, M, C, accumulate_dx, rows_per_cta)
    if (nv <= 2) B2_LN_BWD_F(2);
    else if (nv <= 3) B2_LN_BWD_F(3);
    else B2_LN_BWD_F(5);
#undef B2_LN_BWD_F
    return check_launch("ln_bwd_fused");
  }
#define B2_LN_BWD(NV)                                                                                                 \
  ln_bwd_dx_reg_kernel<NV><<<(M + 7) / 8, 256, 0, st>>>((const bf16*)x, (const bf16*)dy, (bf16*)dx, (const bf16*)gamma, \
                                                        mean, rstd, M, C, accumulate_dx)
  if (parts & 1) {
    if (nv <= 2) B2_LN_BWD(2);
    else if (nv <= 3) B2_LN_BWD(3);
    else if (nv <= 5) B2_LN_BWD(5);
    else
      ln_bwd_dx_kernel<<<(M + 7) / 8, 256, 0, st>>>((const bf16*)x, (const bf16*)dy, (bf16*)dx, (const bf16*)gamma, mean,
                                                    rstd, M, C, accumulate_dx);
    int rc1 = check_launch("ln_bwd_dx");
    if (rc1) return rc1;
  }
#undef B2_LN_BWD
  if (!(parts & 2)) return B2_OK;
  int rc = B2_OK;
  const int colblocks = (C + 63) / 64;
  int rows_per_cta = (int)(((long long)M * colblocks + 4LL * num_sms() - 1) / (4LL * num_sms()));
  if (rows_per_cta < 128) rows_per_cta = 128;
  const int rchunks = (M + rows_per_cta - 1) / rows_per_cta;
  ln_bwd_dgb_kernel<<<dim3(colblocks, rchunks), 256, 0, st>>>((const bf16*)x, (const bf16*)dy, mean, rstd, dgb, M, C,
                                                             rows_per_cta);
  return check_launch("ln_bwd_dgb");
}
