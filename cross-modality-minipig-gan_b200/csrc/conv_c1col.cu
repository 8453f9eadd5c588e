// Weight gradient of the one-input-channel 3x3 convolutions through the tensor cores (SURVEY.md K6 layers: D's
// first convolution 1->64, the generator's 1->16 / 1->32 entry convolutions and its 32->1 ConvTranspose):
//   dw[c][t] = sum_p dY[p][c] * x[pix(p)*s - pad + t]
// is the weight gradient of a 1x1 convolution whose input is the im2col tensor xcol[p][16] (9 taps + 7 zeros), i.e. a
// 16-"channel" layer that tc::wgrad_kernel<16> already handles (both operands MN-major straight from memory, fp32
// accumulation in TMEM).  im2col_c1 writes xcol (32 B per output pixel -- 1/4 of dY's bytes for D layer 1), the
// tcgen05 kernel reduces it against dY, fold_dw16 adds the [cy][16] result into the [cy][9] gradient.
// The direct CUDA-core kernel (c1f::wgrad_kernel) needed 72 register accumulators per thread and ran at ~1.1 TB/s.
// Replaces cuDNN's backward-filter behind nn.Conv2d(1, 64, 3) (/root/reference/code/GAN/GAN_final.py:167-169).
#include "common.cuh"

namespace mpgan {
namespace c1col {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) im2col_kernel(const unsigned short* __restrict__ x, int n, int ih, int iw, int oh,
                                                          int ow, int s, int pad, uint4* __restrict__ xcol) {
  pdl_wait();
  pdl_launch();
  const int64_t P = (int64_t)n * oh * ow;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += stride) {
    const int w = (int)(p % ow);
    const int64_t r = p / ow;
    const int h = (int)(r % oh), img = (int)(r / oh);
    const int ih0 = h * s - pad, iw0 = w * s - pad;
    const unsigned short* xb = x + ((int64_t)img * ih + ih0) * iw + iw0;
    uint32_t v[9];
#pragma unroll
    for (int rh = 0; rh < 3; ++rh) {
      const bool okh = (unsigned)(ih0 + rh) < (unsigned)ih;
#pragma unroll
      for (int rw = 0; rw < 3; ++rw)
        v[rh * 3 + rw] = (okh && (unsigned)(iw0 + rw) < (unsigned)iw) ? (uint32_t)__ldg(xb + rh * iw + rw) : 0u;
    }
    xcol[2 * p] = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
    xcol[2 * p + 1] = make_uint4(v[8], 0u, 0u, 0u);
  }
}

__global__ void fold_dw16_kernel(const float* __restrict__ dw16, int cy, float* __restrict__ dw) {
  pdl_wait();
  pdl_launch();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cy * 9) dw[i] += dw16[(i / 9) * 16 + (i % 9)];
}

}  // namespace c1col
}  // namespace mpgan

using namespace mpgan;

extern "C" int mpgan_im2col_c1(const void* x_bf16, int32_t n, int32_t ih, int32_t iw, int32_t oh, int32_t ow,
                               int32_t stride, int32_t pad, void* xcol_bf16, void* stream) {
  MPGAN_REQUIRE(x_bf16 && xcol_bf16, MPGAN_ERR_SHAPE, "im2col_c1: null pointer");
  MPGAN_REQUIRE(n > 0 && ih > 0 && iw > 0 && oh > 0 && ow > 0, MPGAN_ERR_SHAPE, "im2col_c1: empty tensor");
  MPGAN_REQUIRE(stride == 1 || stride == 2, MPGAN_ERR_UNSUPPORTED, "im2col_c1: stride 1 or 2");
  MPGAN_REQUIRE(pad >= 0 && pad <= 1, MPGAN_ERR_UNSUPPORTED, "im2col_c1: pad 0 or 1");
  MPGAN_REQUIRE(oh <= (ih + 2 * pad - 3) / stride + 1 && ow <= (iw + 2 * pad - 3) / stride + 1, MPGAN_ERR_SHAPE,
                "im2col_c1: output extent exceeds the convolution's");
  MPGAN_REQUIRE(((uintptr_t)xcol_bf16 & 15) == 0, MPGAN_ERR_SHAPE, "im2col_c1: xcol not 16-byte aligned");
  const int64_t P = (int64_t)n * oh * ow;
  int64_t blocks = ceil_div(P, (int64_t)c1col::kThreads * 2);
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  launch_k(c1col::im2col_kernel, (int)blocks, c1col::kThreads, 0, (cudaStream_t)stream, (const unsigned short*)x_bf16,
           (int)n, (int)ih, (int)iw, (int)oh, (int)ow, (int)stride, (int)pad, (uint4*)xcol_bf16);
  MPGAN_CHECK_LAUNCH("im2col_kernel");
  return 0;
}

extern "C" int mpgan_fold_dw16(const float* dw16, int32_t cy, float* dw, void* stream) {
  MPGAN_REQUIRE(dw16 && dw && cy > 0, MPGAN_ERR_SHAPE, "fold_dw16: bad arguments");
  launch_k(c1col::fold_dw16_kernel, (cy * 9 + 127) / 128, 128, 0, (cudaStream_t)stream, dw16, (int)cy, dw);
  MPGAN_CHECK_LAUNCH("fold_dw16_kernel");
  return 0;
}
