// Fast rank-2, 3x3 kernels for the one-channel edge layers: the layers with Cin == 1 (G's first
// convolutions 1->16 / 1->32, D's first convolution 1->64, the 1->1 tail) and Cout == 1 (G's ConvTranspose 32->1).
// They are bandwidth-bound stencils, not GEMMs:
//   * c1f_fprop  (X has 1 channel -> Y has C):  thread = (Y pixel, 8-channel group); the 72 weights of the group
//                live in registers, the nine X taps are broadcast loads, one 16-byte store per thread; the BatchNorm
//                statistics of the stored values are reduced in the same pass.
//   * c1f_bprop  (Y has C -> X has 1 channel):  C/8 lanes cooperate on one X pixel (coalesced 16-byte loads of the
//                Y rows), partial dot products are combined with warp shuffles.
//   * c1f_wgrad  (dw[c][tap], X has 1 channel): thread = (Y pixel, 8-channel group) with 72 register accumulators,
//                reduced over the block with shuffles + shared memory, one fp32 atomic per weight per block.
// Every thread keeps ONE channel group for its whole pixel walk (the grid is a multiple of the group count) and the
// (image, row, column) decomposition of the pixel index is advanced incrementally (no divisions in the loop).
// Replaces the cuDNN calls behind nn.Conv2d / nn.ConvTranspose2d at /root/reference/code/GAN/GAN_final.py:167-169
// (D first conv) and the MONAI UNet's first / last layers (call site GAN_final.py:106-114).
#include "common.cuh"

namespace mpgan {
namespace c1f {

constexpr int kThreads = 256;
constexpr int kTaps = 9;

struct Geom {
  int n, xh, xw, yh, yw, s, pad, C;
};

struct PixWalk {
  int img, h, w, dn, dh, dw, H, W;
  __device__ __forceinline__ void init(int64_t p, int64_t step, int H_, int W_) {
    H = H_; W = W_;
    w = (int)(p % W); p /= W; h = (int)(p % H); img = (int)(p / H);
    dw = (int)(step % W); step /= W; dh = (int)(step % H); dn = (int)(step / H);
  }
  __device__ __forceinline__ void next() {
    w += dw; int c = w >= W ? 1 : 0; w -= c * W;
    h += dh + c; c = h >= H ? 1 : 0; h -= c * H;
    img += dn + c;
  }
};

template <typename T, int V> struct IO;
template <> struct IO<float, 8> {
  static __device__ __forceinline__ void load(const float* p, float* o) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* o) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(o[4], o[5], o[6], o[7]);
  }
};
template <> struct IO<bf16, 8> {
  static __device__ __forceinline__ void load(const bf16* p, float* o) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
  }
  static __device__ __forceinline__ void store(bf16* p, const float* o) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <typename T> struct IO<T, 1> {
  static __device__ __forceinline__ void load(const T* p, float* o) { o[0] = to_f(*p); }
  static __device__ __forceinline__ void store(T* p, const float* o) { *p = from_f<T>(o[0]); }
};

// Sum NV per-thread values over the threads of the block that share channel group (threadIdx.x % cv), cv | 32;
// totals land in sm[g*NV + i] (valid for thread indices < cv*NV after the trailing __syncthreads()).
// sm: (kThreads/32) * cv * NV accumulators of type A, plus the result area cv*NV at the front (aliased safely).
template <typename A, int NV>
__device__ __forceinline__ void group_reduce(A (&a)[NV], int cv, A* sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int off = 16; off >= cv; off >>= 1) {
#pragma unroll
    for (int i = 0; i < NV; ++i) a[i] += __shfl_xor_sync(0xffffffffu, a[i], off);
  }
  if (lane < cv) {
#pragma unroll
    for (int i = 0; i < NV; ++i) sm[(warp * cv + lane) * NV + i] = a[i];
  }
  __syncthreads();
}
template <typename A>
__device__ __forceinline__ A group_total(const A* sm, int cv, int NV, int j) {  // j = g*NV + i
  A t = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) t += sm[w * cv * NV + j];
  return t;
}

// ------------------------------------------------------------------------------------------------------------
// fprop: y[p][c] = bias[c] + sum_t x[pix(p) * s - pad + t] * w[c][t]          (w: [C][9], cx == 1)
// ------------------------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(kThreads, sizeof(T) == 4 ? 1 : 2)
fprop_kernel(Geom g, const T* __restrict__ x, int64_t ldx, const T* __restrict__ w, const float* __restrict__ bias,
             T* __restrict__ y, int64_t ldy, double* __restrict__ stats) {
  __shared__ double sm[(kThreads / 32) * 32 * 2 * V];   // statistics reduce: up to 32 channel groups of 2V doubles per warp
  const int cv = g.C / V;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
  const int c0 = (int)(tid % cv) * V;
  const int64_t P = (int64_t)g.n * g.yh * g.yw;
  const int64_t pstep = nthr / cv;
  float wr[kTaps][V], b[V];
#pragma unroll
  for (int e = 0; e < V; ++e) {
    b[e] = bias ? bias[c0 + e] : 0.f;
#pragma unroll
    for (int t = 0; t < kTaps; ++t) wr[t][e] = to_f(w[(c0 + e) * kTaps + t]);
  }
  // bf16 storage: fp32 partials over this thread's (<= 256, see grid_for) pixels, fp64 from the block level on;
  // fp32 storage (the 1e-4 parity mode): the fp32 partials are flushed into fp64 every 4 pixels, because
  // var = E[x^2] - mean^2 cancels badly when |mean| >> std (un-normalised hand-over between cascaded UNets)
  constexpr bool kExact = sizeof(T) == 4;
  float s1[V], s2[V];
  double d1[kExact ? V : 1], d2[kExact ? V : 1];
#pragma unroll
  for (int e = 0; e < V; ++e) s1[e] = s2[e] = 0.f;
#pragma unroll
  for (int e = 0; e < (kExact ? V : 1); ++e) d1[e] = d2[e] = 0.0;
  int it = 0;
  PixWalk pw;
  int64_t p = tid / cv;
  pw.init(p, pstep, g.yh, g.yw);
  for (; p < P; p += pstep, pw.next()) {
    const T* ximg = x + (int64_t)pw.img * g.xh * g.xw * ldx;
    const int ih0 = pw.h * g.s - g.pad, iw0 = pw.w * g.s - g.pad;
    float xv[kTaps];
#pragma unroll
    for (int rh = 0; rh < 3; ++rh)
#pragma unroll
      for (int rw = 0; rw < 3; ++rw) {
        const int ih = ih0 + rh, iw = iw0 + rw;
        const bool ok = ih >= 0 && ih < g.xh && iw >= 0 && iw < g.xw;
        xv[rh * 3 + rw] = ok ? to_f(ximg[((int64_t)ih * g.xw + iw) * ldx]) : 0.f;
      }
    float o[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
      float a = b[e];
#pragma unroll
      for (int t = 0; t < kTaps; ++t) a = fmaf(xv[t], wr[t][e], a);
      o[e] = a;
    }
    IO<T, V>::store(y + p * ldy + c0, o);
    if (stats) {
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const float f = to_f(from_f<T>(o[e]));   // statistics of the values as stored
        s1[e] += f;
        s2[e] = fmaf(f, f, s2[e]);
      }
      if (kExact && (++it & 3) == 0) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
          d1[kExact ? e : 0] += (double)s1[e]; d2[kExact ? e : 0] += (double)s2[e];
          s1[e] = s2[e] = 0.f;
        }
      }
    }
  }
  if (stats) {
    double a[2 * V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
      a[e] = (double)s1[e] + (kExact ? d1[kExact ? e : 0] : 0.0);
      a[V + e] = (double)s2[e] + (kExact ? d2[kExact ? e : 0] : 0.0);
    }
    if (cv <= 32) {
      group_reduce<double, 2 * V>(a, cv, sm);
      for (int j = threadIdx.x; j < cv * 2 * V; j += kThreads) {
        const int gq = j / (2 * V), i = j - gq * 2 * V;
        atomicAdd(&stats[(i / V) * g.C + gq * V + (i % V)], group_total(sm, cv, 2 * V, j));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 2 * V; ++i) atomicAdd(&stats[(i / V) * g.C + c0 + (i % V)], a[i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// bprop: x[q] = bias + sum_t sum_c y[(q + pad - t)/s][c] * w[c][t]            (cx == 1; cv = C/V lanes per pixel)
// ------------------------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
bprop_kernel(Geom g, const T* __restrict__ y, int64_t ldy, const T* __restrict__ w, const float* __restrict__ bias,
             T* __restrict__ x, int64_t ldx, double* __restrict__ stats) {
  __shared__ double sred[2 * (kThreads / 32)];
  const int cv = g.C / V;   // divides 32
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
  const int gl = (int)(tid % cv);
  const int c0 = gl * V;
  const int64_t P = (int64_t)g.n * g.xh * g.xw;
  const int64_t pstep = nthr / cv;
  float wr[kTaps][V];
#pragma unroll
  for (int e = 0; e < V; ++e)
#pragma unroll
    for (int t = 0; t < kTaps; ++t) wr[t][e] = to_f(w[(c0 + e) * kTaps + t]);
  const float b = bias ? bias[0] : 0.f;
  float s1 = 0.f, s2 = 0.f;
  PixWalk pw;
  int64_t p = tid / cv;
  pw.init(p, pstep, g.xh, g.xw);
  // all lanes of a warp run the same number of iterations (shuffles inside): bound by the warp's first pixel
  const int64_t pwarp = (tid - (threadIdx.x & 31)) / cv;
  for (int64_t pb = pwarp; pb < P; pb += pstep, p += pstep, pw.next()) {
    const bool live = p < P;
    float acc = 0.f;
    if (live) {
      const T* yimg = y + (int64_t)pw.img * g.yh * g.yw * ldy + c0;
#pragma unroll
      for (int rh = 0; rh < 3; ++rh) {
        const int qh = pw.h + g.pad - rh;
        const bool okh = qh >= 0 && (g.s == 1 || (qh & 1) == 0);
        const int yh_ = g.s == 1 ? qh : qh >> 1;
        if (!okh || yh_ >= g.yh) continue;
#pragma unroll
        for (int rw = 0; rw < 3; ++rw) {
          const int qw = pw.w + g.pad - rw;
          const bool okw = qw >= 0 && (g.s == 1 || (qw & 1) == 0);
          const int yw_ = g.s == 1 ? qw : qw >> 1;
          if (!okw || yw_ >= g.yw) continue;
          float v[V];
          IO<T, V>::load(yimg + ((int64_t)yh_ * g.yw + yw_) * ldy, v);
#pragma unroll
          for (int e = 0; e < V; ++e) acc = fmaf(v[e], wr[rh * 3 + rw][e], acc);
        }
      }
    }
    for (int off = cv >> 1; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (live && gl == 0) {
      const T o = from_f<T>(acc + b);
      x[p * ldx] = o;
      if (stats) { const float f = to_f(o); s1 += f; s2 = fmaf(f, f, s2); }
    }
  }
  if (stats) {   // one-channel output: plain block reduction
    double a1 = warp_sum((double)s1), a2 = warp_sum((double)s2);
    if ((threadIdx.x & 31) == 0) { sred[threadIdx.x >> 5] = a1; sred[kThreads / 32 + (threadIdx.x >> 5)] = a2; }
    __syncthreads();
    if (threadIdx.x < 2) {
      double t = 0.0;
#pragma unroll
      for (int i = 0; i < kThreads / 32; ++i) t += sred[threadIdx.x * (kThreads / 32) + i];
      atomicAdd(&stats[threadIdx.x], t);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// wgrad: dw[c][t] += sum_p y[p][c] * x[pix(p) * s - pad + t]                  (cx == 1)
// ------------------------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
wgrad_kernel(Geom g, const T* __restrict__ x, int64_t ldx, const T* __restrict__ y, int64_t ldy,
             float* __restrict__ dw) {
  extern __shared__ float smf[];   // (kThreads/32) * cv * 72 floats
  const int cv = g.C / V;   // divides 32
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
  const int c0 = (int)(tid % cv) * V;
  const int64_t P = (int64_t)g.n * g.yh * g.yw;
  const int64_t pstep = nthr / cv;
  float acc[kTaps * V];
#pragma unroll
  for (int i = 0; i < kTaps * V; ++i) acc[i] = 0.f;
  PixWalk pw;
  int64_t p = tid / cv;
  pw.init(p, pstep, g.yh, g.yw);
  for (; p < P; p += pstep, pw.next()) {
    const T* ximg = x + (int64_t)pw.img * g.xh * g.xw * ldx;
    const int ih0 = pw.h * g.s - g.pad, iw0 = pw.w * g.s - g.pad;
    float gy[V];
    IO<T, V>::load(y + p * ldy + c0, gy);
#pragma unroll
    for (int rh = 0; rh < 3; ++rh)
#pragma unroll
      for (int rw = 0; rw < 3; ++rw) {
        const int ih = ih0 + rh, iw = iw0 + rw;
        const bool ok = ih >= 0 && ih < g.xh && iw >= 0 && iw < g.xw;
        const float xv = ok ? to_f(ximg[((int64_t)ih * g.xw + iw) * ldx]) : 0.f;
#pragma unroll
        for (int e = 0; e < V; ++e) acc[(rh * 3 + rw) * V + e] = fmaf(gy[e], xv, acc[(rh * 3 + rw) * V + e]);
      }
  }
  group_reduce<float, kTaps * V>(acc, cv, smf);
  for (int j = threadIdx.x; j < cv * kTaps * V; j += kThreads) {
    const int gq = j / (kTaps * V), i = j - gq * (kTaps * V);
    const int t = i / V, e = i - t * V;
    atomicAdd(&dw[(gq * V + e) * kTaps + t], group_total(smf, cv, kTaps * V, j));
  }
}

static bool geom_ok(const MpganConvGeom* g, Geom* o) {
  if (!g || g->rank != 2) return false;
  if (g->k[1] != 3 || g->k[2] != 3 || g->k[0] != 1) return false;
  if (g->stride[1] != g->stride[2] || (g->stride[1] != 1 && g->stride[1] != 2)) return false;
  if (g->pad[1] != g->pad[2] || g->pad[1] < 0 || g->pad[1] > 1) return false;
  if (g->cx != 1) return false;
  o->n = g->n; o->xh = g->xs[1]; o->xw = g->xs[2]; o->yh = g->ys[1]; o->yw = g->ys[2];
  o->s = g->stride[1]; o->pad = g->pad[1]; o->C = g->cy;
  return o->n > 0 && o->xh > 0 && o->xw > 0 && o->yh > 0 && o->yw > 0 && o->C > 0;
}

static inline int64_t gcd64(int64_t a, int64_t b) { while (b) { int64_t t = a % b; a = b; b = t; } return a; }

// grid with (gridDim * kThreads) % cv == 0 and about `per_thread` pixels per thread, capped at `bps` blocks per SM
static int grid_for(int64_t pixels, int cv, int per_thread, int bps) {
  int64_t b = ceil_div(pixels * cv, (int64_t)kThreads * per_thread);
  int64_t cap = (int64_t)num_sms() * bps;
  const int64_t floor256 = ceil_div(pixels * cv, (int64_t)kThreads * 256);   // never more than 256 pixels per thread
  if (cap < floor256) cap = floor256;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  const int64_t mult = cv / gcd64(cv, kThreads);
  return (int)(ceil_div(b, mult) * mult);
}

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// vector width usable for the C-channel tensor: 8 (C % 8 == 0, aligned), 1 (C == 1), 0 = not covered
static int vec_of(int C, const void* p, int64_t ld) {
  if (C == 1) return 1;
  if (C % 8 == 0 && ld % 8 == 0 && aligned16(p)) return 8;
  return 0;
}

}  // namespace c1f

// Return 0 on success, MPGAN_ERR_* on failure, 1 when the fast path does not cover the call (caller falls back).
int c1f_fprop(const MpganConvGeom* g, int dtype, const void* x, int64_t ldx, const void* w, const float* bias, void* y,
              int64_t ldy, double* stats, cudaStream_t s) {
  using namespace c1f;
  Geom q;
  if (!geom_ok(g, &q)) return 1;
  const int V = vec_of(q.C, y, ldy);
  if (V == 0) return 1;
  const int cv = q.C / V;
  if (cv <= 32 && (32 % cv) != 0) return 1;
  if (cv > 32 && (kThreads % cv) != 0) return 1;
  const int64_t P = (int64_t)q.n * q.yh * q.yw;
  const int grid = grid_for(P, cv, 8, 8);
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    if (V == 8) fprop_kernel<T, 8><<<grid, kThreads, 0, s>>>(q, (const T*)x, ldx, (const T*)w, bias, (T*)y, ldy, stats);
    else fprop_kernel<T, 1><<<grid, kThreads, 0, s>>>(q, (const T*)x, ldx, (const T*)w, bias, (T*)y, ldy, stats);
    MPGAN_CHECK_LAUNCH("c1f_fprop");
    return 0;
  });
}

int c1f_bprop(const MpganConvGeom* g, int dtype, const void* y, int64_t ldy, const void* w, const float* bias, void* x,
              int64_t ldx, double* stats, cudaStream_t s) {
  using namespace c1f;
  Geom q;
  if (!geom_ok(g, &q)) return 1;
  const int V = vec_of(q.C, y, ldy);
  if (V == 0) return 1;
  const int cv = q.C / V;
  if (cv > 32 || (32 % cv) != 0) return 1;
  const int64_t P = (int64_t)q.n * q.xh * q.xw;
  const int grid = grid_for(P, cv, 4, 8);
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    if (V == 8) bprop_kernel<T, 8><<<grid, kThreads, 0, s>>>(q, (const T*)y, ldy, (const T*)w, bias, (T*)x, ldx, stats);
    else bprop_kernel<T, 1><<<grid, kThreads, 0, s>>>(q, (const T*)y, ldy, (const T*)w, bias, (T*)x, ldx, stats);
    MPGAN_CHECK_LAUNCH("c1f_bprop");
    return 0;
  });
}

int c1f_wgrad(const MpganConvGeom* g, int dtype, const void* x, int64_t ldx, const void* y, int64_t ldy, float* dw,
              cudaStream_t s) {
  using namespace c1f;
  Geom q;
  if (!geom_ok(g, &q)) return 1;
  const int V = vec_of(q.C, y, ldy);
  if (V == 0) return 1;
  const int cv = q.C / V;
  if (cv > 8 || (32 % cv) != 0) return 1;
  const int64_t P = (int64_t)q.n * q.yh * q.yw;
  const int grid = grid_for(P, cv, 32, 4);
  const size_t smem = (size_t)(kThreads / 32) * cv * kTaps * V * sizeof(float);
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    if (V == 8) wgrad_kernel<T, 8><<<grid, kThreads, smem, s>>>(q, (const T*)x, ldx, (const T*)y, ldy, dw);
    else wgrad_kernel<T, 1><<<grid, kThreads, smem, s>>>(q, (const T*)x, ldx, (const T*)y, ldy, dw);
    MPGAN_CHECK_LAUNCH("c1f_wgrad");
    return 0;
  });
}

}  // namespace mpgan
