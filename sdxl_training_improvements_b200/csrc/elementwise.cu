// HBM-bound glue kernels: softmax over materialised logits, GEGLU, SiLU, adds/copies, column sums, layout changes,
// sinusoidal timestep embedding.  All bf16 I/O with fp32 math, 16-byte vector accesses, grid-stride loops sized to
// a multiple of the SM count.
#include "common.cuh"

namespace b2 {

static inline int ew_blocks(long long work_items, int threads) {
  long long b = (work_items + threads - 1) / threads;
  long long cap = 8LL * num_sms();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ---------------------------------------------------------------- softmax
// One warp per row.  S fp32 [rows, lds] (pre-scaled logits), P bf16 [rows, ldp]; P[:, n_valid:ldp_pad] = 0.
__global__ void softmax_fwd_kernel(const float* __restrict__ S, bf16* __restrict__ P, long long rows, int n, long long lds,
                                   long long ldp, int npad) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* s = S + row * lds;
  float mx = -INFINITY;
  for (int i = lane; i < n; i += 32) mx = fmaxf(mx, s[i]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int i = lane; i < n; i += 32) sum += __expf(s[i] - mx);
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  bf16* p = P + row * ldp;
  for (int i = lane; i < npad; i += 32) p[i] = __float2bfloat16(i < n ? __expf(s[i] - mx) * inv : 0.f);
}

// dS = P * (dP - sum_j dP_j P_j) * scale
__global__ void softmax_bwd_kernel(const bf16* __restrict__ P, const float* __restrict__ dP, bf16* __restrict__ dS,
                                   long long rows, int n, long long lds, long long ldp, int npad, float scale) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const bf16* p = P + row * ldp;
  const float* dp = dP + row * lds;
  float dot = 0.f;
  for (int i = lane; i < n; i += 32) dot += dp[i] * __bfloat162float(p[i]);
  dot = warp_sum(dot);
  bf16* ds = dS + row * ldp;
  for (int i = lane; i < npad; i += 32)
    ds[i] = __float2bfloat16(i < n ? __bfloat162float(p[i]) * (dp[i] - dot) * scale : 0.f);
}

// ---------------------------------------------------------------- GEGLU
// exact-erf GELU helpers (gelu_parts / gelu_erf / dgelu_erf): common.cuh — shared with the GEGLU epilogue of gemm2.cu
// I = unsigned (M * F / 8 < 2^31, every SDXL shape) keeps the per-vector row / column split a 32-bit division: with the
// 64-bit one these kernels are co-bound by the integer pipe (~60 of ~270 instructions per 48 bytes of traffic)
template <typename I>
__global__ void geglu_fwd_kernel(const bf16* __restrict__ u, bf16* __restrict__ z, long long M, int F) {
  const I fv = (I)(F >> 3);
  const I total = (I)(M * (F >> 3));
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const long long m = (long long)(i / fv);
    const int c = (int)(i - (I)m * fv) * 8;
    float h[8], g[8];
    unpack8(ld8(u + m * 2 * F + c), h);
    unpack8(ld8(u + m * 2 * F + F + c), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] *= gelu_erf(g[j]);
    st8(z + m * F + c, pack8(h));
  }
}

template <typename I>
__global__ void geglu_bwd_kernel(const bf16* __restrict__ u, const bf16* __restrict__ dz, bf16* __restrict__ du,
                                 long long M, int F) {
  const I fv = (I)(F >> 3);
  const I total = (I)(M * (F >> 3));
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const long long m = (long long)(i / fv);
    const int c = (int)(i - (I)m * fv) * 8;
    float h[8], g[8], d[8], dh[8], dg[8];
    unpack8(ld8(u + m * 2 * F + c), h);
    unpack8(ld8(u + m * 2 * F + F + c), g);
    unpack8(ld8(dz + m * F + c), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float cdf, pdf;
      gelu_parts(g[j], cdf, pdf);
      dh[j] = d[j] * g[j] * cdf;
      dg[j] = d[j] * h[j] * fmaf(g[j], pdf, cdf);
    }
    st8(du + m * 2 * F + c, pack8(dh));
    st8(du + m * 2 * F + F + c, pack8(dg));
  }
}

// GEGLU backward that also produces the ff1 bias gradient: db[c] += sum_m du[m, c] over the bf16 values it stores.  The
// separate column-sum kernel re-read all of du (84 MB per transformer block at 1024 px) right after this kernel wrote it.
// CTA = 32 vector-columns x 8 row lanes; a warp covers 32 consecutive 16-byte vectors of one row (512 B each of h, g, dz);
// per-thread column sums in registers, reduced over the 8 row lanes in shared memory, one atomicAdd per column per CTA.
__global__ void __launch_bounds__(256) geglu_bwd_bias_kernel(const bf16* __restrict__ u, const bf16* __restrict__ dz,
                                                              bf16* __restrict__ du, float* __restrict__ db, long long M, int F,
                                                              long long rows_per_cta) {
  __shared__ float red[8][32 * 16 + 1];
  const int cvl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + cvl) * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float sh[8], sg[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[j] = sg[j] = 0.f;
  if (c < F) {
    for (long long m = r0 + rl; m < r1; m += 16) {
      // two rows per trip: six independent 16-byte loads in flight
      const bool two = m + 8 < r1;
      const long long m2 = two ? m + 8 : m;
      const bf16x8 vh0 = ld8(u + m * 2 * F + c), vg0 = ld8(u + m * 2 * F + F + c), vd0 = ld8(dz + m * F + c);
      const bf16x8 vh1 = ld8(u + m2 * 2 * F + c), vg1 = ld8(u + m2 * 2 * F + F + c), vd1 = ld8(dz + m2 * F + c);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (t == 1 && !two) break;
        float h[8], g[8], d[8], dh[8], dg[8];
        unpack8(t ? vh1 : vh0, h);
        unpack8(t ? vg1 : vg0, g);
        unpack8(t ? vd1 : vd0, d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float cdf, pdf;
          gelu_parts(g[j], cdf, pdf);
          dh[j] = d[j] * g[j] * cdf;
          dg[j] = d[j] * h[j] * fmaf(g[j], pdf, cdf);
        }
        const bf16x8 oh = pack8(dh), og = pack8(dg);
        const long long mm = t ? m2 : m;
        st8(du + mm * 2 * F + c, oh);
        st8(du + mm * 2 * F + F + c, og);
        unpack8(oh, dh);  // the sums are over what du holds (as colsum of du would see it)
        unpack8(og, dg);
#pragma unroll
        for (int j = 0; j < 8; ++j) { sh[j] += dh[j]; sg[j] += dg[j]; }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[rl][cvl * 16 + j] = sh[j];
    red[rl][cvl * 16 + 8 + j] = sg[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * 16; i += 256) {
    const int v = i >> 4, j = i & 15;
    const int cc = (blockIdx.x * 32 + v) * 8 + (j & 7);
    if (cc < F) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) t += red[r][i];
      atomicAdd(db + (j < 8 ? cc : F + cc), t);
    }
  }
}

// ---------------------------------------------------------------- silu / add / copy
__global__ void silu_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16(silu_f(__bfloat162float(x[i])));
}
__global__ void silu_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, bf16* __restrict__ dx,
                                long long n, int accumulate) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float r = __bfloat162float(dy[i]) * dsilu_f(__bfloat162float(x[i]));
    if (accumulate) r += __bfloat162float(dx[i]);
    dx[i] = __float2bfloat16(r);
  }
}
__global__ void add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ o, long long nv,
                           long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    float x[8], y[8];
    unpack8(ld8(a + i * 8), x);
    unpack8(ld8(b + i * 8), y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    st8(o + i * 8, pack8(x));
  }
  if (blockIdx.x == 0)
    for (long long i = nv * 8 + threadIdx.x; i < n; i += blockDim.x)
      o[i] = __float2bfloat16(__bfloat162float(a[i]) + __bfloat162float(b[i]));
}
__global__ void copy2d_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, long long rows, long long cols,
                              long long lds, long long ldd, int accumulate) {
  const long long cv = cols >> 3;
  const long long total = rows * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cv, c = (i - r * cv) * 8;
    bf16x8 v = ld8(src + r * lds + c);
    if (accumulate) {
      float a[8], b[8];
      unpack8(v, a);
      unpack8(ld8(dst + r * ldd + c), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += b[j];
      v = pack8(a);
    }
    st8(dst + r * ldd + c, v);
  }
}

// generic (unaligned / narrow) strided copy, one element per thread
__global__ void copy2d_any_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, long long rows, long long cols,
                                  long long lds, long long ldd, int accumulate) {
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols, c = i - r * cols;
    float v = __bfloat162float(src[r * lds + c]);
    if (accumulate) v += __bfloat162float(dst[r * ldd + c]);
    dst[r * ldd + c] = __float2bfloat16(v);
  }
}

// ---------------------------------------------------------------- column sum (bias gradients)
__global__ void colsum_kernel(const bf16* __restrict__ dy, float* ws, long long M, int N, long long ld,
                              long long rows_per_cta) {
  __shared__ float sm[32][65];
  dy += (long long)blockIdx.z * M * ld;  // row groups (per-sample sums): group z = rows [z*M, (z+1)*M), output row z
  ws += (long long)blockIdx.z * N;
  const int cv = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c0 = blockIdx.x * 64 + cv * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = min(M, r0 + rows_per_cta);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c0 + 8 <= N) {
    // four independent 16-byte loads in flight per thread: with one, the kernel is bound by DRAM latency x trip count
    long long r = r0 + rl;
    for (; r + 96 < r1; r += 128) {
      const bf16x8 v0 = ld8(dy + r * ld + c0), v1 = ld8(dy + (r + 32) * ld + c0), v2 = ld8(dy + (r + 64) * ld + c0),
                   v3 = ld8(dy + (r + 96) * ld + c0);
      float f0[8], f1[8], f2[8], f3[8];
      unpack8(v0, f0);
      unpack8(v1, f1);
      unpack8(v2, f2);
      unpack8(v3, f3);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += (f0[j] + f1[j]) + (f2[j] + f3[j]);
    }
    for (; r < r1; r += 32) {
      float f[8];
      unpack8(ld8(dy + r * ld + c0), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  } else if (c0 < N) {
    for (long long r = r0 + rl; r < r1; r += 32)
      for (int j = 0; j < 8 && c0 + j < N; ++j) acc[j] += __bfloat162float(dy[r * ld + c0 + j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[rl][cv * 8 + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) t += sm[r][threadIdx.x];
    const int gc = blockIdx.x * 64 + threadIdx.x;
    if (gc < N) atomicAdd(&ws[gc], t);
  }
}

// Small-parameter gradient staging (biases, norm scales / shifts): every producer kernel accumulates into ONE fp32 buffer
// with atomics during the backward pass; this kernel folds each segment into the flat bf16 gradient buffer (+=, which is
// what gradient accumulation needs) and clears the staging area for the next micro-step.  Replaces ~1,650 tiny launches
// per step (a zero-fill + an fp32->bf16 accumulate per tensor).
__global__ void flush_small_grads_kernel(float* __restrict__ src, bf16* __restrict__ dst, const long long* __restrict__ seg,
                                         int nseg) {
  for (int sgi = blockIdx.x; sgi < nseg; sgi += gridDim.x) {
    const long long so = seg[3 * sgi], d0 = seg[3 * sgi + 1], n = seg[3 * sgi + 2];
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      const float v = src[so + i];
      if (v != 0.f) {
        dst[d0 + i] = __float2bfloat16(__bfloat162float(dst[d0 + i]) + v);
        src[so + i] = 0.f;
      }
    }
  }
}

__global__ void accum_f32_to_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n, int accumulate) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float r = src[i];
    if (accumulate) r += __bfloat162float(dst[i]);
    dst[i] = __float2bfloat16(r);
  }
}

// ---------------------------------------------------------------- layout: NCHW <-> NHWC(+channel pad)
template <typename T>
__global__ void nchw_to_nhwc_kernel(const T* __restrict__ x, bf16* __restrict__ y, int B, int C, int HW, int Cpad) {
  const long long total = (long long)B * HW * Cpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cpad);
    const long long t = i / Cpad;
    const int p = (int)(t % HW);
    const int b = (int)(t / HW);
    float v = 0.f;
    if (c < C) v = (float)x[((long long)b * C + c) * HW + p];
    y[i] = __float2bfloat16(v);
  }
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const bf16* __restrict__ x, T* __restrict__ y, int B, int C, int HW, int Cpad) {
  const long long total = (long long)B * C * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const long long t = i / HW;
    const int c = (int)(t % C);
    const int b = (int)(t / C);
    y[i] = (T)(__bfloat162float(x[((long long)b * HW + p) * Cpad + c]));
  }
}

// ---------------------------------------------------------------- timestep sinusoid
__global__ void timestep_embedding_kernel(const float* __restrict__ t, bf16* __restrict__ out, int n, int dim, long long ldo) {
  const int half = dim >> 1;
  const int total = n * half;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / half, k = i - r * half;
    const float f = expf(-9.210340371976184f * (float)k / (float)half);  // ln(10000)
    const float a = t[r] * f;
    out[r * ldo + k] = __float2bfloat16(cosf(a));
    out[r * ldo + half + k] = __float2bfloat16(sinf(a));
  }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_softmax_fwd(const float* S, void* P, int64_t rows, int n_valid, int64_t lds, int64_t ldp, void* stream) {
  B2_REQUIRE(S && P && rows > 0 && n_valid > 0 && lds >= n_valid && ldp >= n_valid, "b2_softmax_fwd: bad args");
  const int npad = (int)ldp;
  softmax_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(S, (bf16*)P, rows, n_valid, lds, ldp, npad);
  return check_launch("softmax_fwd");
}
extern "C" int b2_softmax_bwd(const void* P, const float* dP, void* dS, int64_t rows, int n_valid, int64_t lds,
                              int64_t ldp, float scale, void* stream) {
  B2_REQUIRE(P && dP && dS && rows > 0 && n_valid > 0, "b2_softmax_bwd: bad args");
  softmax_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>((const bf16*)P, dP, (bf16*)dS, rows,
                                                                                   n_valid, lds, ldp, (int)ldp, scale);
  return check_launch("softmax_bwd");
}
extern "C" int b2_geglu_fwd(const void* u, void* z, int64_t M, int F, void* stream) {
  B2_REQUIRE(u && z && M > 0 && F % 8 == 0, "b2_geglu_fwd: bad args");
  const long long nvec = M * (F / 8);
  const long long span = nvec + (long long)ew_blocks(nvec, 256) * 256;  // the loop index may overshoot by one grid stride
  if (span < (1LL << 31))
    geglu_fwd_kernel<unsigned><<<ew_blocks(nvec, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)u, (bf16*)z, M, F);
  else
    geglu_fwd_kernel<long long><<<ew_blocks(nvec, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)u, (bf16*)z, M, F);
  return check_launch("geglu_fwd");
}
extern "C" int b2_geglu_bwd(const void* u, const void* dz, void* du, int64_t M, int F, void* stream) {
  B2_REQUIRE(u && dz && du && M > 0 && F % 8 == 0, "b2_geglu_bwd: bad args");
  const long long nvec = M * (F / 8);
  const long long span = nvec + (long long)ew_blocks(nvec, 256) * 256;
  if (span < (1LL << 31))
    geglu_bwd_kernel<unsigned><<<ew_blocks(nvec, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)u, (const bf16*)dz,
                                                                                       (bf16*)du, M, F);
  else
    geglu_bwd_kernel<long long><<<ew_blocks(nvec, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)u, (const bf16*)dz,
                                                                                        (bf16*)du, M, F);
  return check_launch("geglu_bwd");
}
extern "C" int b2_geglu_bwd_bias(const void* u, const void* dz, void* du, float* db32, int64_t M, int F, void* stream) {
  B2_REQUIRE(u && dz && du && db32 && M > 0 && F > 0 && F % 8 == 0, "b2_geglu_bwd_bias: bad args");
  const int colblocks = (F / 8 + 31) / 32;
  long long rchunks = (8LL * num_sms() + colblocks - 1) / colblocks;  // ~8 CTAs per SM in flight
  long long rows_per_cta = (M + rchunks - 1) / rchunks;
  rows_per_cta = (rows_per_cta + 15) / 16 * 16;
  rchunks = (M + rows_per_cta - 1) / rows_per_cta;
  geglu_bwd_bias_kernel<<<dim3(colblocks, (unsigned)rchunks), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)u, (const bf16*)dz, (bf16*)du, db32, M, F, rows_per_cta);
  return check_launch("geglu_bwd_bias");
}
extern "C" int b2_silu_fwd(const void* x, void* y, int64_t n, void* stream) {
  B2_REQUIRE(x && y && n > 0, "b2_silu_fwd: bad args");
  silu_fwd_kernel<<<ew_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, n);
  return check_launch("silu_fwd");
}
extern "C" int b2_silu_bwd(const void* x, const void* dy, void* dx, int64_t n, int accumulate, void* stream) {
  B2_REQUIRE(x && dy && dx && n > 0, "b2_silu_bwd: bad args");
  silu_bwd_kernel<<<ew_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)dy, (bf16*)dx, n,
                                                                       accumulate);
  return check_launch("silu_bwd");
}
extern "C" int b2_add(const void* a, const void* b, void* out, int64_t n, void* stream) {
  B2_REQUIRE(a && b && out && n > 0, "b2_add: bad args");
  const bool al = !(((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15);
  const long long nv = al ? n / 8 : 0;
  add_kernel<<<ew_blocks(nv > 0 ? nv : 1, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)a, (const bf16*)b,
                                                                                (bf16*)out, nv, n);
  return check_launch("add");
}
extern "C" int b2_copy2d(const void* src, void* dst, int64_t rows, int64_t cols, int64_t lds, int64_t ldd,
                         int accumulate, void* stream) {
  B2_REQUIRE(src && dst && rows > 0 && cols > 0, "b2_copy2d: bad args");
  B2_REQUIRE(cols % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0 && !(((uintptr_t)src | (uintptr_t)dst) & 15),
             "b2_copy2d: needs 16-byte aligned rows");
  copy2d_kernel<<<ew_blocks(rows * (cols / 8), 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)src, (bf16*)dst, rows,
                                                                                     cols, lds, ldd, accumulate);
  return check_launch("copy2d");
}
extern "C" int b2_copy2d_any(const void* src, void* dst, int64_t rows, int64_t cols, int64_t lds, int64_t ldd,
                             int accumulate, void* stream) {
  B2_REQUIRE(src && dst && rows > 0 && cols > 0, "b2_copy2d_any: bad args");
  copy2d_any_kernel<<<ew_blocks(rows * cols, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)src, (bf16*)dst, rows, cols,
                                                                                   lds, ldd, accumulate);
  return check_launch("copy2d_any");
}
extern "C" int b2_accum_f32_to_bf16(const float* src, void* dst, int64_t n, int accumulate, void* stream) {
  B2_REQUIRE(src && dst && n > 0, "b2_accum_f32_to_bf16: bad args");
  accum_f32_to_bf16_kernel<<<ew_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, n, accumulate);
  return check_launch("accum_f32_to_bf16");
}
extern "C" int b2_colsum(const void* dy, void* db, int64_t M, int N, int64_t ld, int accumulate, float* ws, void* stream) {
  B2_REQUIRE(dy && db && ws && M > 0 && N > 0, "b2_colsum: bad args");
  B2_REQUIRE(ld % 8 == 0 && !((uintptr_t)dy & 15), "b2_colsum: needs 16-byte aligned rows");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(ws, 0, sizeof(float) * N, st);
  const int colblocks = (N + 63) / 64;
  long long rows_per_cta = (M * colblocks + 4LL * num_sms() - 1) / (4LL * num_sms());
  if (rows_per_cta < 128) rows_per_cta = 128;
  const int rchunks = (int)((M + rows_per_cta - 1) / rows_per_cta);
  colsum_kernel<<<dim3(colblocks, rchunks), 256, 0, st>>>((const bf16*)dy, ws, M, N, ld, rows_per_cta);
  int rc = check_launch("colsum");
  if (rc) return rc;
  accum_f32_to_bf16_kernel<<<ew_blocks(N, 256), 256, 0, st>>>(ws, (bf16*)db, N, accumulate);
  return check_launch("colsum_finish");
}
extern "C" int b2_colsum_groups(const void* dy, void* out, int groups, int64_t rows_per_group, int N, int64_t ld,
                                int accumulate, float* ws, void* stream) {
  B2_REQUIRE(dy && out && ws && groups > 0 && groups <= 65535 && rows_per_group > 0 && N > 0, "b2_colsum_groups: bad args");
  B2_REQUIRE(ld % 8 == 0 && !((uintptr_t)dy & 15), "b2_colsum_groups: needs 16-byte aligned rows");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(ws, 0, sizeof(float) * (size_t)N * groups, st);
  const int colblocks = (N + 63) / 64;
  long long rows_per_cta = (rows_per_group * colblocks * groups + 4LL * num_sms() - 1) / (4LL * num_sms());
  rows_per_cta = (rows_per_cta + 127) / 128 * 128;
  const int rchunks = (int)((rows_per_group + rows_per_cta - 1) / rows_per_cta);
  colsum_kernel<<<dim3(colblocks, rchunks, groups), 256, 0, st>>>((const bf16*)dy, ws, rows_per_group, N, ld, rows_per_cta);
  int rc = check_launch("colsum_groups");
  if (rc) return rc;
  accum_f32_to_bf16_kernel<<<ew_blocks((long long)N * groups, 256), 256, 0, st>>>(ws, (bf16*)out, (long long)N * groups,
                                                                                 accumulate);
  return check_launch("colsum_groups_finish");
}
extern "C" int b2_colsum_f32(const void* dy, float* out, int64_t M, int N, int64_t ld, void* stream) {
  B2_REQUIRE(dy && out && M > 0 && N > 0, "b2_colsum_f32: bad args");
  B2_REQUIRE(ld % 8 == 0 && !((uintptr_t)dy & 15), "b2_colsum_f32: needs 16-byte aligned rows");
  const int colblocks = (N + 63) / 64;
  long long rows_per_cta = (M * colblocks + 4LL * num_sms() - 1) / (4LL * num_sms());
  rows_per_cta = (rows_per_cta + 127) / 128 * 128;  // whole 4-deep load groups
  const int rchunks = (int)((M + rows_per_cta - 1) / rows_per_cta);
  colsum_kernel<<<dim3(colblocks, rchunks), 256, 0, (cudaStream_t)stream>>>((const bf16*)dy, out, M, N, ld, rows_per_cta);
  return check_launch("colsum_f32");
}
extern "C" int b2_flush_small_grads(float* staging, void* grad_bf16, const int64_t* segments, int nseg, void* stream) {
  B2_REQUIRE(staging && grad_bf16 && segments && nseg > 0, "b2_flush_small_grads: bad args");
  int blocks = nseg < 4 * num_sms() ? nseg : 4 * num_sms();
  flush_small_grads_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(staging, (bf16*)grad_bf16,
                                                                     (const long long*)segments, nseg);
  return check_launch("flush_small_grads");
}
extern "C" int b2_nchw_to_nhwc(const void* x, int x_fp32, void* y, int B, int C, int HW, int Cpad, void* stream) {
  B2_REQUIRE(x && y && Cpad >= C, "b2_nchw_to_nhwc: bad args");
  const long long total = (long long)B * HW * Cpad;
  if (x_fp32)
    nchw_to_nhwc_kernel<float><<<ew_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>((const float*)x, (bf16*)y, B, C, HW, Cpad);
  else
    nchw_to_nhwc_kernel<bf16><<<ew_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, B, C, HW, Cpad);
  return check_launch("nchw_to_nhwc");
}
extern "C" int b2_nhwc_to_nchw(const void* x, void* y, int y_fp32, int B, int C, int HW, int Cpad, void* stream) {
  B2_REQUIRE(x && y && Cpad >= C, "b2_nhwc_to_nchw: bad args");
  const long long total = (long long)B * C * HW;
  if (y_fp32)
    nhwc_to_nchw_kernel<float><<<ew_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (float*)y, B, C, HW, Cpad);
  else
    nhwc_to_nchw_kernel<bf16><<<ew_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, B, C, HW, Cpad);
  return check_launch("nhwc_to_nchw");
}
extern "C" int b2_timestep_embedding(const float* t, void* out, int n, int dim, int64_t ldo, void* stream) {
  B2_REQUIRE(t && out && n > 0 && dim % 2 == 0, "b2_timestep_embedding: bad args");
  timestep_embedding_kernel<<<ew_blocks((long long)n * dim / 2, 128), 128, 0, (cudaStream_t)stream>>>(t, (bf16*)out, n, dim, ldo);
  return check_launch("timestep_embedding");
}
