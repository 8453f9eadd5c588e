// One-channel layers of the RANK-3 networks on the tensor cores (the reference's literal 128^3 volumes,
// /root/reference/code/GAN/GAN_final.py:107 dimensions=3, :167-169 Conv3d(1, 64, 3); MONAI UNet's 1 -> 16 entry
// convolutions, 32 -> 1 ConvTranspose and 1 -> 1 tail convolution).
//
// A 3 x 3 x 3 convolution with ONE input channel is a 1 x 1 x 1 convolution over the im2col tensor xcol[p][32]
// (27 taps of output voxel p + 5 zeros), i.e. a 32-"channel" layer the rank-3 tcgen05 kernels already run at HBM speed:
//   forward          y  = conv1x1(im2col(x), w32[cy][32]) (+ bias, + fused BatchNorm statistics in the tcgen05 epilogue)
//   weight gradient  dw = fold(wgrad1x1(im2col(x), dy))                  (fp32 [cy][32] -> accumulated into [cy][27])
//   data gradient    dx = col2im(conv1x1(dy, wT32[32][cy]))              (per-tap partial products, then a 27-point gather)
// im2col / col2im are pure bandwidth kernels (64 B per voxel).  The generic CUDA-core kernels they replace ran D layer 1 at
// 128^3 in 563 / 1957 / 977 us (fprop / dgrad / wgrad); the 1 -> 1 tail convolution (a 27-point stencil on a 4 MB
// volume) gets direct stencil kernels instead of an implicit GEMM with N = 1.
#include "common.cuh"

namespace mpgan {
namespace c1vol {

constexpr int kThreads = 256;
constexpr int TP = 32;   // padded tap count

struct Vol {
  int n, xd, xh, xw, yd, yh, yw, s, pad;
};

// x (n, xd, xh, xw) one channel -> xcol (n, yd, yh, yw, 32): tap t = (rd * 3 + rh) * 3 + rw of output voxel p
__global__ void __launch_bounds__(kThreads) im2col3_kernel(const unsigned short* __restrict__ x, const Vol v,
                                                           uint4* __restrict__ xcol) {
  pdl_wait();
  pdl_launch();
  const int64_t P = (int64_t)v.n * v.yd * v.yh * v.yw;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += stride) {
    const int w = (int)(p % v.yw);
    int64_t r = p / v.yw;
    const int h = (int)(r % v.yh); r /= v.yh;
    const int d = (int)(r % v.yd);
    const int img = (int)(r / v.yd);
    const int d0 = d * v.s - v.pad, h0 = h * v.s - v.pad, w0 = w * v.s - v.pad;
    const unsigned short* xb = x + (int64_t)img * v.xd * v.xh * v.xw;
    uint32_t t[TP];
#pragma unroll
    for (int i = 27; i < TP; ++i) t[i] = 0u;
#pragma unroll
    for (int rd = 0; rd < 3; ++rd) {
      const bool okd = (unsigned)(d0 + rd) < (unsigned)v.xd;
#pragma unroll
      for (int rh = 0; rh < 3; ++rh) {
        const bool okh = okd && (unsigned)(h0 + rh) < (unsigned)v.xh;
        const unsigned short* row = xb + ((int64_t)(d0 + rd) * v.xh + (h0 + rh)) * v.xw + w0;
#pragma unroll
        for (int rw = 0; rw < 3; ++rw)
          t[(rd * 3 + rh) * 3 + rw] = (okh && (unsigned)(w0 + rw) < (unsigned)v.xw) ? (uint32_t)__ldg(row + rw) : 0u;
      }
    }
    uint4* o = xcol + 4 * p;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      o[q] = make_uint4(t[8 * q] | (t[8 * q + 1] << 16), t[8 * q + 2] | (t[8 * q + 3] << 16), t[8 * q + 4] | (t[8 * q + 5] << 16),
                        t[8 * q + 6] | (t[8 * q + 7] << 16));
  }
}

// T (n, yd, yh, yw, 32) per-tap partial products -> x (n, xd, xh, xw):  x[q] = bias + res[q] + sum over taps r with
// (q + pad - r) divisible by s and inside Y of T[(q + pad - r) / s][r]
__global__ void __launch_bounds__(kThreads) col2im3_kernel(const bf16* __restrict__ T, const Vol v, const float* __restrict__ bias,
                                                           const bf16* __restrict__ res, bf16* __restrict__ x) {
  pdl_wait();
  pdl_launch();
  const int64_t Q = (int64_t)v.n * v.xd * v.xh * v.xw;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float b = bias ? __ldg(bias) : 0.f;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < Q; q += stride) {
    const int w = (int)(q % v.xw);
    int64_t r = q / v.xw;
    const int h = (int)(r % v.xh); r /= v.xh;
    const int d = (int)(r % v.xd);
    const int img = (int)(r / v.xd);
    const bf16* Tb = T + (int64_t)img * v.yd * v.yh * v.yw * TP;
    float acc = b;
#pragma unroll
    for (int rd = 0; rd < 3; ++rd) {
      const int ad = d + v.pad - rd;
      if (ad < 0 || (v.s == 2 && (ad & 1))) continue;
      const int yd_ = v.s == 2 ? ad >> 1 : ad;
      if (yd_ >= v.yd) continue;
#pragma unroll
      for (int rh = 0; rh < 3; ++rh) {
        const int ah = h + v.pad - rh;
        if (ah < 0 || (v.s == 2 && (ah & 1))) continue;
        const int yh_ = v.s == 2 ? ah >> 1 : ah;
        if (yh_ >= v.yh) continue;
#pragma unroll
        for (int rw = 0; rw < 3; ++rw) {
          const int aw = w + v.pad - rw;
          if (aw < 0 || (v.s == 2 && (aw & 1))) continue;
          const int yw_ = v.s == 2 ? aw >> 1 : aw;
          if (yw_ >= v.yw) continue;
          acc += to_f(Tb[(((int64_t)yd_ * v.yh + yh_) * v.yw + yw_) * TP + (rd * 3 + rh) * 3 + rw]);
        }
      }
    }
    if (res) acc += to_f(res[q]);
    x[q] = from_f<bf16>(acc);
  }
}

// dw[c][27] += dw32[c][32] (first 27 columns)
__global__ void fold_dw32_kernel(const float* __restrict__ dw32, int cy, float* __restrict__ dw) {
  pdl_wait();
  pdl_launch();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cy * 27) dw[i] += dw32[(i / 27) * TP + (i % 27)];
}

// ---- 1 -> 1 channel, k3 s1 p1: direct 27-point stencils ----
// MODE 0: y = b + sum_t w[t] x[p + off(t)];  MODE 1 (data gradient): dx = sum_t w[t] dy[p - off(t)] (+ res)
template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads) stencil3_kernel(const T* __restrict__ in, int n, int D, int H, int W,
                                                            const T* __restrict__ w27, const float* __restrict__ bias,
                                                            const T* __restrict__ res, T* __restrict__ out) {
  pdl_wait();
  pdl_launch();
  float wt[27];
#pragma unroll
  for (int t = 0; t < 27; ++t) wt[t] = to_f(w27[MODE == 0 ? t : 26 - t]);
  const float b = bias ? __ldg(bias) : 0.f;
  const int64_t P = (int64_t)n * D * H * W;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += stride) {
    const int x = (int)(p % W);
    int64_t r = p / W;
    const int y = (int)(r % H); r /= H;
    const int z = (int)(r % D);
    float acc = b;
#pragma unroll
    for (int rd = 0; rd < 3; ++rd) {
      const int zz = z + rd - 1;
      if ((unsigned)zz >= (unsigned)D) continue;
#pragma unroll
      for (int rh = 0; rh < 3; ++rh) {
        const int yy = y + rh - 1;
        if ((unsigned)yy >= (unsigned)H) continue;
        const T* row = in + p + ((int64_t)(rd - 1) * H + (rh - 1)) * W;
#pragma unroll
        for (int rw = 0; rw < 3; ++rw) {
          const int xx = x + rw - 1;
          if ((unsigned)xx < (unsigned)W) acc = fmaf(wt[(rd * 3 + rh) * 3 + rw], to_f(row[rw - 1]), acc);
        }
      }
    }
    if (res) acc += to_f(res[p]);
    out[p] = from_f<T>(acc);
  }
}

// dw[t] += sum_p dy[p] x[p + off(t)]
template <typename T>
__global__ void __launch_bounds__(kThreads) stencil3_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, int n, int D,
                                                                  int H, int W, float* __restrict__ dw, float* __restrict__ db) {
  pdl_wait();
  pdl_launch();
  __shared__ float red[32];
  float acc[27];
#pragma unroll
  for (int t = 0; t < 27; ++t) acc[t] = 0.f;
  float sdy = 0.f;
  const int64_t P = (int64_t)n * D * H * W;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += stride) {
    const int xw = (int)(p % W);
    int64_t r = p / W;
    const int y = (int)(r % H); r /= H;
    const int z = (int)(r % D);
    const float g = to_f(dy[p]);
    sdy += g;
#pragma unroll
    for (int rd = 0; rd < 3; ++rd) {
      const int zz = z + rd - 1;
      if ((unsigned)zz >= (unsigned)D) continue;
#pragma unroll
      for (int rh = 0; rh < 3; ++rh) {
        const int yy = y + rh - 1;
        if ((unsigned)yy >= (unsigned)H) continue;
        const T* row = x + p + ((int64_t)(rd - 1) * H + (rh - 1)) * W;
#pragma unroll
        for (int rw = 0; rw < 3; ++rw) {
          const int xx = xw + rw - 1;
          if ((unsigned)xx < (unsigned)W) acc[(rd * 3 + rh) * 3 + rw] = fmaf(g, to_f(row[rw - 1]), acc[(rd * 3 + rh) * 3 + rw]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 27; ++t) {
    const float s = block_sum(acc[t], red);
    if (threadIdx.x == 0 && s != 0.f) atomicAdd(&dw[t], s);
  }
  if (db) {
    const float s = block_sum(sdy, red);
    if (threadIdx.x == 0 && s != 0.f) atomicAdd(db, s);
  }
}

static int grid_for(int64_t items) {
  int64_t blocks = ceil_div(items, (int64_t)kThreads);
  const int64_t cap = (int64_t)num_sms() * 16;
  return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

static int check_vol(const char* what, int n, const int32_t* xs, const int32_t* ys, int stride, int pad, Vol* v) {
  MPGAN_REQUIRE(n > 0 && xs && ys, MPGAN_ERR_SHAPE, "%s: bad arguments", what);
  MPGAN_REQUIRE(stride == 1 || stride == 2, MPGAN_ERR_UNSUPPORTED, "%s: stride 1 or 2", what);
  MPGAN_REQUIRE(pad >= 0 && pad <= 1, MPGAN_ERR_UNSUPPORTED, "%s: pad 0 or 1", what);
  for (int i = 0; i < 3; ++i) {
    MPGAN_REQUIRE(xs[i] > 0 && ys[i] > 0, MPGAN_ERR_SHAPE, "%s: empty tensor", what);
    MPGAN_REQUIRE(ys[i] <= (xs[i] + 2 * pad - 3) / stride + 1, MPGAN_ERR_SHAPE, "%s: Y extent exceeds the convolution's", what);
  }
  v->n = n; v->xd = xs[0]; v->xh = xs[1]; v->xw = xs[2]; v->yd = ys[0]; v->yh = ys[1]; v->yw = ys[2]; v->s = stride; v->pad = pad;
  return 0;
}

}  // namespace c1vol
}  // namespace mpgan

using namespace mpgan;

extern "C" int mpgan_im2col_c1_vol(const void* x_bf16, int32_t n, const int32_t* xs3, const int32_t* ys3, int32_t stride,
                                   int32_t pad, void* xcol_bf16, void* stream) {
  c1vol::Vol v;
  int rc = c1vol::check_vol("im2col_c1_vol", n, xs3, ys3, stride, pad, &v);
  if (rc) return rc;
  MPGAN_REQUIRE(x_bf16 && xcol_bf16 && ((uintptr_t)xcol_bf16 & 15) == 0, MPGAN_ERR_SHAPE, "im2col_c1_vol: bad pointers");
  launch_k(c1vol::im2col3_kernel, c1vol::grid_for((int64_t)n * v.yd * v.yh * v.yw), c1vol::kThreads, 0, (cudaStream_t)stream,
           (const unsigned short*)x_bf16, v, (uint4*)xcol_bf16);
  MPGAN_CHECK_LAUNCH("im2col3_kernel");
  return 0;
}

extern "C" int mpgan_col2im_c1_vol(const void* t_bf16, int32_t n, const int32_t* xs3, const int32_t* ys3, int32_t stride,
                                   int32_t pad, const float* bias, const void* res_bf16, void* x_bf16, void* stream) {
  c1vol::Vol v;
  int rc = c1vol::check_vol("col2im_c1_vol", n, xs3, ys3, stride, pad, &v);
  if (rc) return rc;
  MPGAN_REQUIRE(t_bf16 && x_bf16, MPGAN_ERR_SHAPE, "col2im_c1_vol: null pointer");
  launch_k(c1vol::col2im3_kernel, c1vol::grid_for((int64_t)n * v.xd * v.xh * v.xw), c1vol::kThreads, 0, (cudaStream_t)stream,
           (const bf16*)t_bf16, v, bias, (const bf16*)res_bf16, (bf16*)x_bf16);
  MPGAN_CHECK_LAUNCH("col2im3_kernel");
  return 0;
}

extern "C" int mpgan_fold_dw32(const float* dw32, int32_t cy, float* dw, void* stream) {
  MPGAN_REQUIRE(dw32 && dw && cy > 0, MPGAN_ERR_SHAPE, "fold_dw32: bad arguments");
  launch_k(c1vol::fold_dw32_kernel, (cy * 27 + 127) / 128, 128, 0, (cudaStream_t)stream, dw32, (int)cy, dw);
  MPGAN_CHECK_LAUNCH("fold_dw32_kernel");
  return 0;
}

extern "C" int mpgan_stencil27(int dtype, int direction, const void* in, int32_t n, int32_t d, int32_t h, int32_t w,
                               const void* w27, const float* bias, const void* res, void* out, void* stream) {
  MPGAN_REQUIRE(in && w27 && out && n > 0 && d > 0 && h > 0 && w > 0, MPGAN_ERR_SHAPE, "stencil27: bad arguments");
  MPGAN_REQUIRE(direction == 0 || direction == 1, MPGAN_ERR_SHAPE, "stencil27: direction 0 (forward) or 1 (data gradient)");
  const int grid = c1vol::grid_for((int64_t)n * d * h * w);
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    if (direction == 0)
      launch_k(c1vol::stencil3_kernel<T, 0>, grid, c1vol::kThreads, 0, (cudaStream_t)stream, (const T*)in, (int)n, (int)d, (int)h,
               (int)w, (const T*)w27, bias, (const T*)res, (T*)out);
    else
      launch_k(c1vol::stencil3_kernel<T, 1>, grid, c1vol::kThreads, 0, (cudaStream_t)stream, (const T*)in, (int)n, (int)d, (int)h,
               (int)w, (const T*)w27, bias, (const T*)res, (T*)out);
    MPGAN_CHECK_LAUNCH("stencil3_kernel");
    return 0;
  });
}

extern "C" int mpgan_stencil27_wgrad(int dtype, const void* x, const void* dy, int32_t n, int32_t d, int32_t h, int32_t w,
                                     float* dw27, float* dbias, void* stream) {
  MPGAN_REQUIRE(x && dy && dw27 && n > 0 && d > 0 && h > 0 && w > 0, MPGAN_ERR_SHAPE, "stencil27_wgrad: bad arguments");
  const int64_t P = (int64_t)n * d * h * w;
  int64_t blocks = ceil_div(P, (int64_t)c1vol::kThreads * 8);
  const int64_t cap = (int64_t)num_sms() * 4;
  const int grid = (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    launch_k(c1vol::stencil3_wgrad_kernel<T>, grid, c1vol::kThreads, 0, (cudaStream_t)stream, (const T*)x, (const T*)dy, (int)n,
             (int)d, (int)h, (int)w, dw27, dbias);
    MPGAN_CHECK_LAUNCH("stencil3_wgrad_kernel");
    return 0;
  });
}
