// 3x3 convolution operand staging for the implicit-GEMM path: im2col (forward / wgrad operand) and its exact
// adjoint col2im in gather form (dgrad).  K ordering is (kh, kw, cin) so that a diffusers OIHW weight stored
// channels-last (O,H,W,I contiguous) is directly the K-major B operand [Cout, 9*Cin] of b2_gemm.
// stride 1|2, padding 1; `upsample` folds Upsample2D's nearest-2x into the address map (no 4x tensor).
#include "common.cuh"

namespace b2 {

__global__ void im2col3x3_kernel(const bf16* __restrict__ x, bf16* __restrict__ col, int B, int H, int W, int C, int stride,
                                 int up, int Ho, int Wo, long long ldc) {
  const int cv = C >> 3;
  const int kv = (int)(ldc >> 3);  // vectors per col row (9*cv data + pad)
  const long long total = (long long)B * Ho * Wo * kv;
  const int Hin = up ? 2 * H : H, Win = up ? 2 * W : W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / kv;
    const int kvi = (int)(i - m * kv);
    bf16x8 v = zero8();
    if (kvi < 9 * cv) {
      const int tap = kvi / cv, c = (kvi - tap * cv) * 8;
      const int kh = tap / 3, kw = tap - kh * 3;
      const int wo = (int)(m % Wo);
      const long long t = m / Wo;
      const int ho = (int)(t % Ho);
      const int b = (int)(t / Ho);
      int hi = ho * stride + kh - 1, wi = wo * stride + kw - 1;
      if (hi >= 0 && hi < Hin && wi >= 0 && wi < Win) {
        if (up) {
          hi >>= 1;
          wi >>= 1;
        }
        v = ld8(x + (((long long)b * H + hi) * W + wi) * C + c);
      }
    }
    st8(col + m * ldc + (long long)kvi * 8, v);
  }
}

__global__ void col2im3x3_kernel(const bf16* __restrict__ dcol, bf16* __restrict__ dx, int B, int H, int W, int C, int stride,
                                 int up, int Ho, int Wo, long long ldc, int accumulate) {
  const int cv = C >> 3;
  const long long total = (long long)B * H * W * cv;
  const int Hin = up ? 2 * H : H, Win = up ? 2 * W : W;
  (void)Hin; (void)Win;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * 8;
    long long t = i / cv;
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H);
    const int b = (int)(t / H);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int nsub = up ? 2 : 1;
    for (int a = 0; a < nsub; ++a)
      for (int bb = 0; bb < nsub; ++bb) {
        const int hv = up ? 2 * h + a : h, wv = up ? 2 * w + bb : w;  // pixel of the (virtual) conv input
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const int hn = hv + 1 - kh;  // = ho * stride
          if (hn < 0 || (hn % stride) != 0) continue;
          const int ho = hn / stride;
          if (ho >= Ho) continue;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int wn = wv + 1 - kw;
            if (wn < 0 || (wn % stride) != 0) continue;
            const int wo = wn / stride;
            if (wo >= Wo) continue;
            const long long m = ((long long)b * Ho + ho) * Wo + wo;
            float f[8];
            unpack8(ld8(dcol + m * ldc + (long long)(kh * 3 + kw) * C + c), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f[j];
          }
        }
      }
    bf16* o = dx + (((long long)b * H + h) * W + w) * C + c;
    if (accumulate) {
      float f[8];
      unpack8(ld8(o), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
    st8(o, pack8(acc));
  }
}

static inline void conv_out_hw(int H, int W, int stride, int up, int* Ho, int* Wo) {
  const int Hin = up ? 2 * H : H, Win = up ? 2 * W : W;
  *Ho = (Hin + 2 - 3) / stride + 1;
  *Wo = (Win + 2 - 3) / stride + 1;
}

}  // namespace b2

using namespace b2;

extern "C" int b2_im2col3x3(const void* x, void* col, int B, int H, int W, int C, int stride, int upsample, int64_t ldc,
                            void* stream) {
  B2_REQUIRE(x && col, "b2_im2col3x3: null pointer");
  B2_REQUIRE(C % 8 == 0 && ldc % 8 == 0 && ldc >= 9LL * C, "b2_im2col3x3: C=%d ldc=%lld unsupported", C, (long long)ldc);
  B2_REQUIRE((stride == 1 || stride == 2) && !(upsample && stride != 1), "b2_im2col3x3: bad stride/upsample");
  int Ho, Wo;
  conv_out_hw(H, W, stride, upsample, &Ho, &Wo);
  const long long total = (long long)B * Ho * Wo * (ldc / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = 16LL * num_sms();
  if (blocks > cap) blocks = cap;
  im2col3x3_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)col, B, H, W, C, stride,
                                                                       upsample, Ho, Wo, ldc);
  return check_launch("im2col3x3");
}

extern "C" int b2_col2im3x3(const void* dcol, void* dx, int B, int H, int W, int C, int stride, int upsample,
                            int64_t ldc, int accumulate, void* stream) {
  B2_REQUIRE(dcol && dx, "b2_col2im3x3: null pointer");
  B2_REQUIRE(C % 8 == 0 && ldc % 8 == 0 && ldc >= 9LL * C, "b2_col2im3x3: C=%d ldc=%lld unsupported", C, (long long)ldc);
  B2_REQUIRE((stride == 1 || stride == 2) && !(upsample && stride != 1), "b2_col2im3x3: bad stride/upsample");
  int Ho, Wo;
  conv_out_hw(H, W, stride, upsample, &Ho, &Wo);
  const long long total = (long long)B * H * W * (C / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = 16LL * num_sms();
  if (blocks > cap) blocks = cap;
  col2im3x3_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)dcol, (bf16*)dx, B, H, W, C, stride,
                                                                       upsample, Ho, Wo, ldc, accumulate);
  return check_launch("col2im3x3");
}
