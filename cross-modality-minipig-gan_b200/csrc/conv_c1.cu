// Direct convolution kernels for the one-channel edge layers: with Cin or Cout == 1 there is no GEMM
// worth issuing, the op is a bandwidth-bound stencil.
//   * conv_to1:   output has ONE channel (G's ConvTranspose 32->1 forward, the data gradients of 1->16 / 1->64 / 1->1
//                 convolutions): one thread per output pixel, vectorised channel dot products against weights held in
//                 shared memory (broadcast reads).
//   * wgrad_x1:   weight gradient when the X side has ONE channel: dw[cy][tap] += sum_p y[p][cy] * x[gather(p,tap)].
// Replaces the corresponding cuDNN calls behind nn.Conv2d / nn.ConvTranspose2d at
// /root/reference/code/GAN/GAN_final.py:167-169 (D first conv) and the MONAI UNet's first / last layers.
#include <stdlib.h>

#include "common.cuh"

namespace mpgan {

struct C1Geom {
  int rank, n;
  int xs[3], ys[3], k[3], st[3], pd[3];
  int cx, cy, taps;
};

static int make_c1(const MpganConvGeom* g, C1Geom* p) {
  MPGAN_REQUIRE(g && (g->rank == 2 || g->rank == 3), MPGAN_ERR_SHAPE, "bad geometry");
  p->rank = g->rank; p->n = g->n; p->cx = g->cx; p->cy = g->cy; p->taps = 1;
  for (int i = 0; i < 3; ++i) {
    p->xs[i] = g->xs[i]; p->ys[i] = g->ys[i]; p->k[i] = g->k[i]; p->st[i] = g->stride[i]; p->pd[i] = g->pad[i];
    MPGAN_REQUIRE(p->xs[i] > 0 && p->ys[i] > 0 && p->k[i] > 0 && p->st[i] > 0 && p->pd[i] >= 0, MPGAN_ERR_SHAPE,
                  "bad conv geometry");
    p->taps *= p->k[i];
  }
  return 0;
}

template <typename T> struct LoadVec;
template <> struct LoadVec<float> {
  static constexpr int V = 4;
  static __device__ __forceinline__ void load(const float* p, float* o) {
    float4 a = *reinterpret_cast<const float4*>(p);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
  }
};
template <> struct LoadVec<bf16> {
  static constexpr int V = 8;
  static __device__ __forceinline__ void load(const bf16* p, float* o) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
  }
};

// MODE 0: out = Y grid (cy == 1), gather X (C = cx):  xpos = ypos*st - pd + r,  w[t*cx + c]
// MODE 1: out = X grid (cx == 1), gather Y (C = cy):  ypos = (xpos + pd - r)/st,  w[c*taps + t]
template <typename T, int MODE, bool VEC>
__global__ void __launch_bounds__(256)
conv_to1_kernel(C1Geom p, const T* __restrict__ in, int64_t ldi, const T* __restrict__ w,
                const float* __restrict__ bias, T* __restrict__ out, int64_t ldo) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  extern __shared__ float ws[];  // [taps][C]
  const int C = MODE == 0 ? p.cx : p.cy;
  for (int i = threadIdx.x; i < p.taps * C; i += blockDim.x) {
    int t = i / C, c = i - t * C;
    ws[i] = to_f(MODE == 0 ? w[i] : w[(int64_t)c * p.taps + t]);
  }
  __syncthreads();
  const int* os = MODE == 0 ? p.ys : p.xs;
  const int* is = MODE == 0 ? p.xs : p.ys;
  const int64_t P = (int64_t)p.n * os[0] * os[1] * os[2];
  const float b = bias ? bias[0] : 0.f;
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < P; m += (int64_t)gridDim.x * blockDim.x) {
    int ow = (int)(m % os[2]); int64_t r = m / os[2];
    int oh = (int)(r % os[1]); r /= os[1];
    int od = (int)(r % os[0]);
    int64_t img = r / os[0];
    float acc = b;
    int t = 0;
    for (int rd = 0; rd < p.k[0]; ++rd)
      for (int rh = 0; rh < p.k[1]; ++rh)
        for (int rw = 0; rw < p.k[2]; ++rw, ++t) {
          int id, ih, iw;
          bool ok = true;
          if (MODE == 0) {
            id = od * p.st[0] - p.pd[0] + rd; ih = oh * p.st[1] - p.pd[1] + rh; iw = ow * p.st[2] - p.pd[2] + rw;
          } else {
            int qd = od + p.pd[0] - rd, qh = oh + p.pd[1] - rh, qw = ow + p.pd[2] - rw;
            ok = qd >= 0 && qh >= 0 && qw >= 0 && (qd % p.st[0] == 0) && (qh % p.st[1] == 0) && (qw % p.st[2] == 0);
            id = qd / p.st[0]; ih = qh / p.st[1]; iw = qw / p.st[2];
          }
          ok = ok && id >= 0 && id < is[0] && ih >= 0 && ih < is[1] && iw >= 0 && iw < is[2];
          if (!ok) continue;
          const T* src = in + ((((img * is[0] + id) * is[1] + ih) * is[2] + iw) * ldi);
          const float* wt = ws + t * C;
          if (VEC) {
            constexpr int V = LoadVec<T>::V;
            for (int c = 0; c < C; c += V) {
              float v[V];
              LoadVec<T>::load(src + c, v);
#pragma unroll
              for (int e = 0; e < V; ++e) acc = fmaf(v[e], wt[c + e], acc);
            }
          } else {
            for (int c = 0; c < C; ++c) acc = fmaf(to_f(src[c]), wt[c], acc);
          }
        }
    out[m * ldo] = from_f<T>(acc);
  }
}

// dw[cy][taps] (cx == 1) += sum over Y pixels y[p][cy] * x[gather(p, tap)]
// block: 256 pixels per iteration staged in shared memory, then a (cy*taps) x 256 product per block.
template <typename T>
__global__ void __launch_bounds__(256)
wgrad_x1_kernel(C1Geom p, const T* __restrict__ x, int64_t ldx, const T* __restrict__ y, int64_t ldy,
                float* __restrict__ dw, int64_t pix_per_block) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  extern __shared__ float sm[];
  const int CY = p.cy, TP = p.taps;
  float* ysm = sm;                    // [256][CY + 1]
  float* xsm = sm + 256 * (CY + 1);   // [TP][256]
  const int64_t P = (int64_t)p.n * p.ys[0] * p.ys[1] * p.ys[2];
  const int64_t pbeg = (int64_t)blockIdx.x * pix_per_block, pend = min(P, pbeg + pix_per_block);
  const int nout = CY * TP;
  constexpr int MAXO = 4;  // outputs per thread (cy*taps <= 1024)
  float acc[MAXO];
#pragma unroll
  for (int i = 0; i < MAXO; ++i) acc[i] = 0.f;
  for (int64_t pb = pbeg; pb < pend; pb += 256) {
    const int64_t m = pb + threadIdx.x;
    const bool pv = m < pend;
    int ow = 0, oh = 0, od = 0;
    int64_t img = 0;
    if (pv) {
      ow = (int)(m % p.ys[2]); int64_t r = m / p.ys[2];
      oh = (int)(r % p.ys[1]); r /= p.ys[1];
      od = (int)(r % p.ys[0]);
      img = r / p.ys[0];
    }
    int t = 0;
    for (int rd = 0; rd < p.k[0]; ++rd)
      for (int rh = 0; rh < p.k[1]; ++rh)
        for (int rw = 0; rw < p.k[2]; ++rw, ++t) {
          float v = 0.f;
          if (pv) {
            int id = od * p.st[0] - p.pd[0] + rd, ih = oh * p.st[1] - p.pd[1] + rh, iw = ow * p.st[2] - p.pd[2] + rw;
            if (id >= 0 && id < p.xs[0] && ih >= 0 && ih < p.xs[1] && iw >= 0 && iw < p.xs[2])
              v = to_f(x[(((img * p.xs[0] + id) * p.xs[1] + ih) * p.xs[2] + iw) * ldx]);
          }
          xsm[t * 256 + threadIdx.x] = v;
        }
    // y rows: coalesced over (pixel, channel)
    for (int i = threadIdx.x; i < 256 * CY; i += 256) {
      int pp = i / CY, c = i - pp * CY;
      int64_t mm = pb + pp;
      ysm[pp * (CY + 1) + c] = mm < pend ? to_f(y[mm * ldy + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int o = 0; o < MAXO; ++o) {
      int idx = threadIdx.x + o * 256;
      if (idx < nout) {
        int c = idx / TP, tt = idx - c * TP;
        float a = 0.f;
        const float* xr = xsm + tt * 256;
#pragma unroll 8
        for (int pp = 0; pp < 256; ++pp) a = fmaf(ysm[pp * (CY + 1) + c], xr[pp], a);
        acc[o] += a;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int o = 0; o < MAXO; ++o) {
    int idx = threadIdx.x + o * 256;
    if (idx < nout && pbeg < pend) atomicAdd(&dw[idx], acc[o]);
  }
}

}  // namespace mpgan

namespace mpgan {
// conv_c1_fast.cu: rank-2 3x3 fast paths; return 1 when they do not cover the call
int c1f_fprop(const MpganConvGeom* g, int dtype, const void* x, int64_t ldx, const void* w, const float* bias, void* y,
              int64_t ldy, double* stats, cudaStream_t s, const float* slope = nullptr);
int c1f_bprop(const MpganConvGeom* g, int dtype, const void* y, int64_t ldy, const void* w, const float* bias, void* x,
              int64_t ldx, double* stats, cudaStream_t s);
int c1f_wgrad(const MpganConvGeom* g, int dtype, const void* x, int64_t ldx, const void* y, int64_t ldy, float* dw,
              cudaStream_t s);
// conv_c1mma.cuh (conv_tc.cu): tcgen05 path for bf16 cx == 1 forward convolutions; returns 1 when not covered
int c1mma_fprop(const MpganConvGeom* g, int dtype, const void* x, int64_t ldx, const void* w, const float* bias, void* y,
                int64_t ldy, double* stats, cudaStream_t s);
}  // namespace mpgan

using namespace mpgan;

// 1 if the direct one-channel kernels cover (geometry, direction): 0 fprop with cy == 1, 1 bprop with cx == 1,
// 2 wgrad with cx == 1.
extern "C" int mpgan_c1_supported(const MpganConvGeom* g, int direction) {
  if (!g) return 0;
  int taps = g->k[0] * g->k[1] * g->k[2];
  if (direction == 0) {
    if (g->cx == 1 && g->rank == 2 && g->k[1] == 3 && g->k[2] == 3 && g->stride[1] == g->stride[2] &&
        g->stride[1] <= 2 && g->pad[1] == g->pad[2] && g->pad[1] <= 1 && (g->cy == 1 || (g->cy % 8 == 0 && (g->cy / 8 <= 32 ? 32 % (g->cy / 8) == 0 : 256 % (g->cy / 8) == 0))))
      return 1;   // rank-2 3x3 one-input-channel layer: conv_c1_fast.cu
    return g->cy == 1 && taps * g->cx * 4 <= 48 * 1024;
  }
  if (direction == 1) return g->cx == 1 && taps * g->cy * 4 <= 48 * 1024;
  if (direction == 2) return g->cx == 1 && g->cy * taps <= 1024 && (256 * (g->cy + 1) + taps * 256) * 4 <= 160 * 1024;
  return 0;
}

template <typename T, int MODE>
static int launch_to1(const C1Geom& p, const void* in, int64_t ldi, const void* w, const float* bias, void* out,
                      int64_t ldo, cudaStream_t s) {
  const int* os = MODE == 0 ? p.ys : p.xs;
  const int C = MODE == 0 ? p.cx : p.cy;
  const int64_t P = (int64_t)p.n * os[0] * os[1] * os[2];
  int64_t blocks = ceil_div(P, 256);
  int64_t cap = (int64_t)num_sms() * 32;
  int grid = (int)(blocks > cap ? cap : blocks);
  size_t smem = (size_t)p.taps * C * sizeof(float);
  const bool vec = (C % LoadVec<T>::V == 0) && (ldi % LoadVec<T>::V == 0) && (((uintptr_t)in) % 16 == 0);
  if (vec) launch_k(conv_to1_kernel<T, MODE, true>, grid, 256, smem, s, p, (const T*)in, ldi, (const T*)w, bias, (T*)out, ldo);
  else launch_k(conv_to1_kernel<T, MODE, false>, grid, 256, smem, s, p, (const T*)in, ldi, (const T*)w, bias, (T*)out, ldo);
  MPGAN_CHECK_LAUNCH("conv_to1_kernel");
  return 0;
}

extern "C" int mpgan_c1_conv_fprop(const MpganConvGeom* g, int dtype, const void* x, int64_t ldx, const void* w,
                                   const float* bias, void* y, int64_t ldy, double* stats, void* stream) {
  C1Geom p;
  int rc = make_c1(g, &p);
  if (rc) return rc;
  static const bool use_mma = !(getenv("MPGAN_NO_C1MMA") && getenv("MPGAN_NO_C1MMA")[0] == '1');
  rc = use_mma ? c1mma_fprop(g, dtype, x, ldx, w, bias, y, ldy, stats, (cudaStream_t)stream) : 1;   // bf16, cy in {16,32,64}
  if (rc != 1) return rc;
  rc = c1f_fprop(g, dtype, x, ldx, w, bias, y, ldy, stats, (cudaStream_t)stream);   // rank-2 3x3, cx == 1
  if (rc != 1) return rc;
  if (g->cy == 1 && (size_t)p.taps * g->cx * 4 <= 48 * 1024) {
    MPGAN_DISPATCH_DTYPE(dtype, T, rc = (launch_to1<T, 0>(p, x, ldx, w, bias, y, ldy, (cudaStream_t)stream)));
  } else {  // e.g. an unaligned channel slice: the generic implicit-GEMM kernel covers everything
    rc = mpgan_conv_fprop(g, dtype, x, ldx, w, bias, y, ldy, stream);
  }
  if (rc || !stats) return rc;
  return mpgan_bn_stats(dtype, y, ldy, (int64_t)p.n * p.ys[0] * p.ys[1] * p.ys[2], p.cy, stats, stream);
}

// Inference-mode fused one-input-channel layer: y = prelu(conv(x, w_folded) + bias_folded), BatchNorm folded into w / bias
// by the caller (MONAI Convolution in model.eval(), /root/reference/code/GAN/inferrence.py:107-109).  Covered: rank-2 3x3
// layers the run-based kernel takes (8 | cy; stride 1, or stride 2 on a contiguous bf16 image with 8 | width).
extern "C" int mpgan_c1_conv_act(const MpganConvGeom* g, int dtype, const void* x, int64_t ldx, const void* w,
                                 const float* bias, const float* slope, void* y, int64_t ldy, void* stream) {
  C1Geom p;
  int rc = make_c1(g, &p);
  if (rc) return rc;
  MPGAN_REQUIRE(slope != nullptr, MPGAN_ERR_SHAPE, "c1_conv_act: slope must be given (use mpgan_c1_conv_fprop otherwise)");
  rc = c1f_fprop(g, dtype, x, ldx, w, bias, y, ldy, nullptr, (cudaStream_t)stream, slope);
  MPGAN_REQUIRE(rc != 1, MPGAN_ERR_UNSUPPORTED, "c1_conv_act: layer not covered by the run-based one-channel kernel");
  return rc;
}

extern "C" int mpgan_c1_conv_bprop(const MpganConvGeom* g, int dtype, const void* y, int64_t ldy, const void* w,
                                   const float* bias, void* x, int64_t ldx, double* stats, void* stream) {
  C1Geom p;
  int rc = make_c1(g, &p);
  if (rc) return rc;
  MPGAN_REQUIRE(mpgan_c1_supported(g, 1), MPGAN_ERR_UNSUPPORTED, "c1 bprop needs cx == 1");
  rc = c1f_bprop(g, dtype, y, ldy, w, bias, x, ldx, stats, (cudaStream_t)stream);
  if (rc != 1) return rc;
  MPGAN_DISPATCH_DTYPE(dtype, T, rc = (launch_to1<T, 1>(p, y, ldy, w, bias, x, ldx, (cudaStream_t)stream)));
  if (rc || !stats) return rc;
  return mpgan_bn_stats(dtype, x, ldx, (int64_t)p.n * p.xs[0] * p.xs[1] * p.xs[2], p.cx, stats, stream);
}

extern "C" int mpgan_c1_conv_wgrad(const MpganConvGeom* g, int dtype, const void* x, int64_t ldx, const void* y,
                                   int64_t ldy, float* dw, void* stream) {
  C1Geom p;
  int rc = make_c1(g, &p);
  if (rc) return rc;
  MPGAN_REQUIRE(mpgan_c1_supported(g, 2), MPGAN_ERR_UNSUPPORTED, "c1 wgrad needs cx == 1 and cy*taps <= 1024");
  rc = c1f_wgrad(g, dtype, x, ldx, y, ldy, dw, (cudaStream_t)stream);
  if (rc != 1) return rc;
  const int64_t P = (int64_t)p.n * p.ys[0] * p.ys[1] * p.ys[2];
  int64_t want = (int64_t)num_sms() * 4;
  int64_t ppb = ceil_div(ceil_div(P, want), 256) * 256;
  if (ppb < 256) ppb = 256;
  int grid = (int)ceil_div(P, ppb);
  size_t smem = (size_t)(256 * (p.cy + 1) + p.taps * 256) * sizeof(float);
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    static bool attr_done = false;
    if (!attr_done && smem > 48 * 1024) {
      cudaFuncSetAttribute(wgrad_x1_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      attr_done = true;
    }
    launch_k(wgrad_x1_kernel<T>, grid, 256, smem, (cudaStream_t)stream, p, (const T*)x, ldx, (const T*)y, ldy, dw, ppb);
    MPGAN_CHECK_LAUNCH("wgrad_x1_kernel");
    return 0;
  });
}
