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
  int vec;   // the one-channel tensor is contiguous bf16 with 16-byte aligned rows: window loads are 16-byte vectors
};

struct PixWalk {
  int img, h, w, dn, dh, dw, H, W;
  __device__ __forceinline__ void init(int p, int step, int H_, int W_) {
    H = H_; W = W_;
    w = p % W; p /= W; h = p % H; img = p / H;
    dw = step % W; step /= W; dh = step % H; dn = step / H;
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

// Blackwell issues fp32 FMAs at full rate only in the packed form (FFMA2: two lanes of a 64-bit register pair)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}

// V channels as V/2 packed pairs (V == 1: one pair with a dead upper lane)
template <int V> struct Pairs { static constexpr int N = V / 2; };
template <> struct Pairs<1> { static constexpr int N = 1; };

// ------------------------------------------------------------------------------------------------------------
// fprop: y[p][c] = bias[c] + sum_t x[pix(p) * s - pad + t] * w[c][t]          (w: [C][9], cx == 1)
// All element offsets fit in int32 (checked on the host).
// ------------------------------------------------------------------------------------------------------------
template <typename T, int V, int S>
__global__ void __launch_bounds__(kThreads, sizeof(T) == 4 ? 1 : 2)
fprop_kernel(Geom g, const T* __restrict__ x, int ldx, const T* __restrict__ w, const float* __restrict__ bias,
             T* __restrict__ y, int ldy, double* __restrict__ stats) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  __shared__ double sm[(kThreads / 32) * 32 * 2 * V];   // statistics reduce: up to 32 channel groups of 2V doubles per warp
  constexpr int NP = Pairs<V>::N;
  const int cv = g.C / V;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthr = gridDim.x * blockDim.x;
  const int c0 = (tid % cv) * V;
  const int P = g.n * g.yh * g.yw;
  const int pstep = nthr / cv;
  float2 wr[kTaps][NP], b[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const int e0 = 2 * j, e1 = (2 * j + 1 < V) ? 2 * j + 1 : 2 * j;
    b[j] = make_float2(bias ? bias[c0 + e0] : 0.f, bias ? bias[c0 + e1] : 0.f);
#pragma unroll
    for (int t = 0; t < kTaps; ++t)
      wr[t][j] = make_float2(to_f(w[(c0 + e0) * kTaps + t]), to_f(w[(c0 + e1) * kTaps + t]));
  }
  // bf16 storage: fp32 partials over this thread's (<= 256, see grid_for) pixels, fp64 from the block level on;
  // fp32 storage (the 1e-4 parity mode): the fp32 partials are flushed into fp64 every 4 pixels, because
  // var = E[x^2] - mean^2 cancels badly when |mean| >> std (un-normalised hand-over between cascaded UNets)
  constexpr bool kExact = sizeof(T) == 4;
  float2 s1[NP], s2[NP];
  double d1[kExact ? V : 1], d2[kExact ? V : 1];
#pragma unroll
  for (int j = 0; j < NP; ++j) s1[j] = s2[j] = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < (kExact ? V : 1); ++e) d1[e] = d2[e] = 0.0;
  int it = 0;
  PixWalk pw;
  int p = tid / cv;
  pw.init(p, pstep, g.yh, g.yw);
  const int xrow = g.xw * ldx, ximg = g.xh * xrow;
  for (; p < P; p += pstep, pw.next()) {
    const int ih0 = pw.h * S - g.pad, iw0 = pw.w * S - g.pad;
    const T* xb = x + (pw.img * ximg + ih0 * xrow + iw0 * ldx);
    bool vh[3], vw[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) { vh[r] = (unsigned)(ih0 + r) < (unsigned)g.xh; vw[r] = (unsigned)(iw0 + r) < (unsigned)g.xw; }
    float xv[kTaps];
#pragma unroll
    for (int rh = 0; rh < 3; ++rh)
#pragma unroll
      for (int rw = 0; rw < 3; ++rw)
        xv[rh * 3 + rw] = (vh[rh] && vw[rw]) ? to_f(xb[rh * xrow + rw * ldx]) : 0.f;
    float2 o2[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      float2 a = b[j];
#pragma unroll
      for (int t = 0; t < kTaps; ++t) a = ffma2(make_float2(xv[t], xv[t]), wr[t][j], a);
      o2[j] = a;
    }
    float o[V];
#pragma unroll
    for (int e = 0; e < V; ++e) o[e] = (e & 1) ? o2[e / 2].y : o2[e / 2].x;
    IO<T, V>::store(y + (p * ldy + c0), o);
    if (stats) {
#pragma unroll
      for (int j = 0; j < NP; ++j) {   // statistics of the values as stored
        const float2 f = make_float2(to_f(from_f<T>(o2[j].x)), to_f(from_f<T>(o2[j].y)));
        s1[j].x += f.x; s1[j].y += f.y;
        s2[j] = ffma2(f, f, s2[j]);
      }
      if (kExact && (++it & 3) == 0) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
          d1[kExact ? e : 0] += (double)((e & 1) ? s1[e / 2].y : s1[e / 2].x);
          d2[kExact ? e : 0] += (double)((e & 1) ? s2[e / 2].y : s2[e / 2].x);
        }
#pragma unroll
        for (int j = 0; j < NP; ++j) s1[j] = s2[j] = make_float2(0.f, 0.f);
      }
    }
  }
  if (stats) {
    double a[2 * V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
      a[e] = (double)((e & 1) ? s1[e / 2].y : s1[e / 2].x) + (kExact ? d1[kExact ? e : 0] : 0.0);
      a[V + e] = (double)((e & 1) ? s2[e / 2].y : s2[e / 2].x) + (kExact ? d2[kExact ? e : 0] : 0.0);
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
// Run-based variant of fprop for C % 8 == 0: a thread owns 8 channels and walks RUN consecutive output
// pixels of one row; the 3 x (RUN*S + 2) input window is loaded once per run, so the inner loop is 36 packed FMAs
// per (pixel, 8 channels) plus one 16-byte access -- the per-pixel index arithmetic and tap predicates of the
// pixel-at-a-time kernels (which made them instruction-issue bound at ~5x the FMA floor) are amortised over the run.
// ------------------------------------------------------------------------------------------------------------
template <int S> struct RunOf { static constexpr int RUN = S == 1 ? 8 : 4; static constexpr int WIN = (RUN - 1) * S + 3; };

template <typename T, int S>
__device__ __forceinline__ void load_window(const Geom& g, const T* __restrict__ x, int ldx, int img, int h, int w0,
                                            float (&xv)[3][RunOf<S>::WIN]) {
  constexpr int WIN = RunOf<S>::WIN;
  const int ih0 = h * S - g.pad, iw0 = w0 * S - g.pad;
  const T* xi = x + img * (g.xh * g.xw) * ldx;
#pragma unroll
  for (int rh = 0; rh < 3; ++rh) {
    const int ih = ih0 + rh;
    const bool okh = (unsigned)ih < (unsigned)g.xh;
    const T* xr = xi + (ih * g.xw + iw0) * ldx;
#pragma unroll
    for (int j = 0; j < WIN; ++j)
      xv[rh][j] = (okh && (unsigned)(iw0 + j) < (unsigned)g.xw) ? to_f(xr[j * ldx]) : 0.f;
  }
}

// The same window for a CONTIGUOUS one-channel bf16 image whose rows are 16-byte aligned (xw % 8 == 0): the RUN * S
// columns starting at w0 * S are one aligned 16-byte vector (RUN * S == 8 for both strides), the one or two columns
// beside it are scalar loads -- 2-3 load instructions per window row instead of 9-10 (ncu: 60 % of the samples of the
// pixel-at-a-time kernels sat on their scalar 2-byte loads).  Requires w0 * S + 7 < xw (full run); PAD is 0 or 1.
template <int S, int PAD>
__device__ __forceinline__ void load_window_vec(const Geom& g, const bf16* __restrict__ x, int img, int h, int w0,
                                                float (&xv)[3][RunOf<S>::WIN]) {
  constexpr int WIN = RunOf<S>::WIN;
  static_assert(RunOf<S>::RUN * S == 8, "one 16-byte vector per window row");
  const int ih0 = h * S - PAD, cv0 = w0 * S;
  const bf16* xi = x + (size_t)img * g.xh * g.xw;
#pragma unroll
  for (int rh = 0; rh < 3; ++rh) {
    const int ih = ih0 + rh;
    const bool okh = (unsigned)ih < (unsigned)g.xh;
    const bf16* xr = xi + (size_t)(okh ? ih : 0) * g.xw;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (okh) u = *reinterpret_cast<const uint4*>(xr + cv0);
    const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      xv[rh][PAD + 2 * e] = __uint_as_float(uu[e] << 16);
      xv[rh][PAD + 2 * e + 1] = __uint_as_float(uu[e] & 0xffff0000u);
    }
#pragma unroll
    for (int j = 0; j < WIN; ++j) {
      if (j < PAD || j >= PAD + 8) {
        const int iw = cv0 - PAD + j;
        xv[rh][j] = (okh && (unsigned)iw < (unsigned)g.xw) ? to_f(xr[iw]) : 0.f;
      }
    }
  }
}

template <typename T, int S>
__device__ __forceinline__ void load_window_any(const Geom& g, const T* __restrict__ x, int ldx, int img, int h, int w0,
                                                float (&xv)[3][RunOf<S>::WIN]) {
  if constexpr (sizeof(T) == 2) {
    if (g.vec && w0 * S + 7 < g.xw) {
      if (g.pad == 1) load_window_vec<S, 1>(g, x, img, h, w0, xv);
      else load_window_vec<S, 0>(g, x, img, h, w0, xv);
      return;
    }
  }
  load_window<T, S>(g, x, ldx, img, h, w0, xv);
}

template <typename T, int S>
__global__ void __launch_bounds__(kThreads, 1)
fprop_run_kernel(Geom g, const T* __restrict__ x, int ldx, const T* __restrict__ w, const float* __restrict__ bias,
                 T* __restrict__ y, int ldy, double* __restrict__ stats, const float* __restrict__ slope) {
  pdl_wait();
  pdl_launch();
  constexpr int RUN = RunOf<S>::RUN, WIN = RunOf<S>::WIN;
  __shared__ double sm[(kThreads / 32) * 32 * 16];
  const int cv = g.C / 8;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthr = gridDim.x * blockDim.x;
  const int c0 = (tid % cv) * 8;
  const int rpr = (g.yw + RUN - 1) / RUN;                  // runs per output row
  const int nruns = g.n * g.yh * rpr;
  const int rstep = nthr / cv;
  float2 wr[kTaps][4], b[4];
  bool wvec = false;
  if constexpr (sizeof(T) == 2) wvec = (reinterpret_cast<uintptr_t>(w) & 15) == 0;
  if (wvec) {   // the 8 x 9 weights of the group are 144 contiguous, 16-byte aligned bytes: nine vector loads, not 72 scalar ones
    uint4 u[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) u[i] = __ldg(reinterpret_cast<const uint4*>(w + c0 * kTaps) + i);
    const uint16_t* wl = reinterpret_cast<const uint16_t*>(u);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int t = 0; t < kTaps; ++t)
        wr[t][j] = make_float2(__uint_as_float((uint32_t)wl[(2 * j) * kTaps + t] << 16),
                               __uint_as_float((uint32_t)wl[(2 * j + 1) * kTaps + t] << 16));
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    b[j] = make_float2(bias ? bias[c0 + 2 * j] : 0.f, bias ? bias[c0 + 2 * j + 1] : 0.f);
    if (!wvec) {
#pragma unroll
      for (int t = 0; t < kTaps; ++t)
        wr[t][j] = make_float2(to_f(w[(c0 + 2 * j) * kTaps + t]), to_f(w[(c0 + 2 * j + 1) * kTaps + t]));
    }
  }
  constexpr bool kExact = sizeof(T) == 4;   // fp32 storage: flush the fp32 statistics partials into fp64 every run
  const float slope_a = slope ? __ldg(slope) : 1.f;
  float2 s1[4], s2[4];
  double d1[kExact ? 8 : 1], d2[kExact ? 8 : 1];
#pragma unroll
  for (int j = 0; j < 4; ++j) s1[j] = s2[j] = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < (kExact ? 8 : 1); ++e) d1[e] = d2[e] = 0.0;
  for (int run = tid / cv; run < nruns; run += rstep) {
    const int rw = run % rpr;
    const int rr = run / rpr;
    const int h = rr % g.yh, img = rr / g.yh;
    const int w0 = rw * RUN;
    float xv[3][WIN];
    load_window_any<T, S>(g, x, ldx, img, h, w0, xv);
    T* yrow = y + ((img * g.yh + h) * g.yw + w0) * ldy + c0;
#pragma unroll
    for (int i = 0; i < RUN; ++i) {
      float2 o2[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 a = b[j];
#pragma unroll
        for (int rh = 0; rh < 3; ++rh)
#pragma unroll
          for (int q = 0; q < 3; ++q) a = ffma2(make_float2(xv[rh][i * S + q], xv[rh][i * S + q]), wr[rh * 3 + q][j], a);
        o2[j] = a;
      }
      if (slope) {   // inference fusion: BatchNorm folded into (w, bias) by the caller, PReLU here (uniform branch)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          o2[j].x = o2[j].x > 0.f ? o2[j].x : slope_a * o2[j].x;
          o2[j].y = o2[j].y > 0.f ? o2[j].y : slope_a * o2[j].y;
        }
      }
      if (w0 + i < g.yw) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (e & 1) ? o2[e / 2].y : o2[e / 2].x;
        IO<T, 8>::store(yrow + i * ldy, o);
        if (stats) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {   // statistics of the values as stored
            const float2 f = make_float2(to_f(from_f<T>(o2[j].x)), to_f(from_f<T>(o2[j].y)));
            s1[j].x += f.x; s1[j].y += f.y;
            s2[j] = ffma2(f, f, s2[j]);
          }
        }
      }
    }
    if (kExact && stats) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        d1[kExact ? e : 0] += (double)((e & 1) ? s1[e / 2].y : s1[e / 2].x);
        d2[kExact ? e : 0] += (double)((e & 1) ? s2[e / 2].y : s2[e / 2].x);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) s1[j] = s2[j] = make_float2(0.f, 0.f);
    }
  }
  if (stats) {
    double a[16];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      a[e] = (double)((e & 1) ? s1[e / 2].y : s1[e / 2].x) + (kExact ? d1[kExact ? e : 0] : 0.0);
      a[8 + e] = (double)((e & 1) ? s2[e / 2].y : s2[e / 2].x) + (kExact ? d2[kExact ? e : 0] : 0.0);
    }
    group_reduce<double, 16>(a, cv, sm);
    for (int j = threadIdx.x; j < cv * 16; j += kThreads) {
      const int gq = j / 16, i = j - gq * 16;
      atomicAdd(&stats[(i / 8) * g.C + gq * 8 + (i % 8)], group_total(sm, cv, 16, j));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// bprop: x[q] = bias + sum_t sum_c y[(q + pad - t)/s][c] * w[c][t]            (cx == 1; cv = C/V lanes per pixel)
// ------------------------------------------------------------------------------------------------------------
template <typename T, int V, int S>
__global__ void __launch_bounds__(kThreads, 2)
bprop_kernel(Geom g, const T* __restrict__ y, int ldy, const T* __restrict__ w, const float* __restrict__ bias,
             T* __restrict__ x, int ldx, double* __restrict__ stats) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  __shared__ double sred[2 * (kThreads / 32)];
  constexpr int NP = Pairs<V>::N;
  const int cv = g.C / V;   // divides 32
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthr = gridDim.x * blockDim.x;
  const int gl = tid % cv;
  const int c0 = gl * V;
  const int P = g.n * g.xh * g.xw;
  const int pstep = nthr / cv;
  float2 wr[kTaps][NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const int e0 = 2 * j;
    const bool has1 = 2 * j + 1 < V;
#pragma unroll
    for (int t = 0; t < kTaps; ++t)
      wr[t][j] = make_float2(to_f(w[(c0 + e0) * kTaps + t]), has1 ? to_f(w[(c0 + e0 + 1) * kTaps + t]) : 0.f);
  }
  const float b = bias ? bias[0] : 0.f;
  float s1 = 0.f, s2 = 0.f;
  PixWalk pw;
  int p = tid / cv;
  pw.init(p, pstep, g.xh, g.xw);
  const int yrow = g.yw * ldy, yimg = g.yh * yrow;
  // all lanes of a warp run the same number of iterations (shuffles inside): bound by the warp's first pixel
  const int pwarp = (tid - (int)(threadIdx.x & 31)) / cv;
  for (int pb = pwarp; pb < P; pb += pstep, p += pstep, pw.next()) {
    const bool live = p < P;
    bool vh[3], vw[3];
    int oh[3], ow[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int qh = pw.h + g.pad - r, qw = pw.w + g.pad - r;
      const int yh_ = S == 1 ? qh : qh >> 1, yw_ = S == 1 ? qw : qw >> 1;
      vh[r] = live && qh >= 0 && (S == 1 || (qh & 1) == 0) && yh_ < g.yh;
      vw[r] = qw >= 0 && (S == 1 || (qw & 1) == 0) && yw_ < g.yw;
      oh[r] = yh_ * yrow; ow[r] = yw_ * ldy;
    }
    const T* yb = y + (pw.img * yimg + c0);
    float2 acc2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int rh = 0; rh < 3; ++rh)
#pragma unroll
      for (int rw = 0; rw < 3; ++rw) {
        float v[V];
        if (vh[rh] && vw[rw]) {
          IO<T, V>::load(yb + (oh[rh] + ow[rw]), v);
        } else {
#pragma unroll
          for (int e = 0; e < V; ++e) v[e] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < NP; ++j)
          acc2 = ffma2(make_float2(v[2 * j], (2 * j + 1 < V) ? v[(2 * j + 1 < V) ? 2 * j + 1 : 0] : 0.f), wr[rh * 3 + rw][j], acc2);
      }
    float acc = acc2.x + acc2.y;
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
// bprop, stride 2, pad 1, X = 2*Y (ConvTranspose k3 s2 p1 op1 forward; data gradient of the s2 first convolutions):
// Y-centric form -- the lane group of Y pixel (a,b) produces the 2x2 output block (2a..2a+1, 2b..2b+1):
//   x(2a,  2b)   = Y(a,b).W11
//   x(2a,  2b+1) = Y(a,b+1).W10 + Y(a,b).W12
//   x(2a+1,2b)   = Y(a+1,b).W01 + Y(a,b).W21
//   x(2a+1,2b+1) = Y(a+1,b+1).W00 + Y(a+1,b).W02 + Y(a,b+1).W20 + Y(a,b).W22
// (no multiply by structural zeros, four coalesced 16-byte loads per lane, nine 8-channel dot products)
// ------------------------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(kThreads, 2)
bprop_s2_kernel(Geom g, const T* __restrict__ y, int ldy, const T* __restrict__ w, const float* __restrict__ bias,
                T* __restrict__ x, int ldx, double* __restrict__ stats) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  __shared__ double sred[2 * (kThreads / 32)];
  constexpr int NP = Pairs<V>::N;
  const int cv = g.C / V;   // divides 32
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthr = gridDim.x * blockDim.x;
  const int gl = tid % cv;
  const int c0 = gl * V;
  const int P = g.n * g.yh * g.yw;
  const int pstep = nthr / cv;
  float2 wr[kTaps][NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const int e0 = 2 * j;
    const bool has1 = 2 * j + 1 < V;
#pragma unroll
    for (int t = 0; t < kTaps; ++t)
      wr[t][j] = make_float2(to_f(w[(c0 + e0) * kTaps + t]), has1 ? to_f(w[(c0 + e0 + 1) * kTaps + t]) : 0.f);
  }
  const float b = bias ? bias[0] : 0.f;
  float s1 = 0.f, s2 = 0.f;
  PixWalk pw;
  int p = tid / cv;
  pw.init(p, pstep, g.yh, g.yw);
  const int yrow = g.yw * ldy;
  const int xrow = g.xw * ldx;
  const int pwarp = (tid - (int)(threadIdx.x & 31)) / cv;
  for (int pb = pwarp; pb < P; pb += pstep, p += pstep, pw.next()) {
    const bool live = p < P;
    const bool r1 = live && pw.h + 1 < g.yh, c1 = pw.w + 1 < g.yw;
    const T* yb = y + (live ? (p * ldy + c0) : 0);
    float v[4][V];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const bool ok = q == 0 ? live : (q == 1 ? (live && c1) : (q == 2 ? r1 : (r1 && c1)));
      if (ok) {
        IO<T, V>::load(yb + ((q >> 1) * yrow + (q & 1) * ldy), v[q]);
      } else {
#pragma unroll
        for (int e = 0; e < V; ++e) v[q][e] = 0.f;
      }
    }
    float2 o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) o[q] = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      float2 u[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) u[q] = make_float2(v[q][2 * j], (2 * j + 1 < V) ? v[q][(2 * j + 1 < V) ? 2 * j + 1 : 0] : 0.f);
      o[0] = ffma2(u[0], wr[4][j], o[0]);                                   // W11
      o[1] = ffma2(u[1], wr[3][j], o[1]); o[1] = ffma2(u[0], wr[5][j], o[1]);   // W10, W12
      o[2] = ffma2(u[2], wr[1][j], o[2]); o[2] = ffma2(u[0], wr[7][j], o[2]);   // W01, W21
      o[3] = ffma2(u[3], wr[0][j], o[3]); o[3] = ffma2(u[2], wr[2][j], o[3]);   // W00, W02
      o[3] = ffma2(u[1], wr[6][j], o[3]); o[3] = ffma2(u[0], wr[8][j], o[3]);   // W20, W22
    }
    float r[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      r[q] = o[q].x + o[q].y;
      for (int off = cv >> 1; off >= 1; off >>= 1) r[q] += __shfl_xor_sync(0xffffffffu, r[q], off);
    }
    if (live && gl == 0) {
      T* xb = x + ((pw.img * g.xh + 2 * pw.h) * xrow + 2 * pw.w * ldx);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const T ov = from_f<T>(r[q] + b);
        xb[(q >> 1) * xrow + (q & 1) * ldx] = ov;
        if (stats) { const float f = to_f(ov); s1 += f; s2 = fmaf(f, f, s2); }
      }
    }
  }
  if (stats) {
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
template <typename T, int V, int S>
__global__ void __launch_bounds__(kThreads, 2)
wgrad_kernel(Geom g, const T* __restrict__ x, int ldx, const T* __restrict__ y, int ldy, float* __restrict__ dw) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  extern __shared__ float smf[];   // (kThreads/32) * cv * 9 * V floats
  constexpr int NP = Pairs<V>::N;
  const int cv = g.C / V;   // divides 32
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthr = gridDim.x * blockDim.x;
  const int c0 = (tid % cv) * V;
  const int P = g.n * g.yh * g.yw;
  const int pstep = nthr / cv;
  float2 acc2[kTaps][NP];
#pragma unroll
  for (int t = 0; t < kTaps; ++t)
#pragma unroll
    for (int j = 0; j < NP; ++j) acc2[t][j] = make_float2(0.f, 0.f);
  PixWalk pw;
  int p = tid / cv;
  pw.init(p, pstep, g.yh, g.yw);
  const int xrow = g.xw * ldx, ximg = g.xh * xrow;
  for (; p < P; p += pstep, pw.next()) {
    const int ih0 = pw.h * S - g.pad, iw0 = pw.w * S - g.pad;
    const T* xb = x + (pw.img * ximg + ih0 * xrow + iw0 * ldx);
    bool vh[3], vw[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) { vh[r] = (unsigned)(ih0 + r) < (unsigned)g.xh; vw[r] = (unsigned)(iw0 + r) < (unsigned)g.xw; }
    float gy[V];
    IO<T, V>::load(y + (p * ldy + c0), gy);
    float2 g2[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) g2[j] = make_float2(gy[2 * j], (2 * j + 1 < V) ? gy[(2 * j + 1 < V) ? 2 * j + 1 : 0] : 0.f);
#pragma unroll
    for (int rh = 0; rh < 3; ++rh)
#pragma unroll
      for (int rw = 0; rw < 3; ++rw) {
        const float xv = (vh[rh] && vw[rw]) ? to_f(xb[rh * xrow + rw * ldx]) : 0.f;
#pragma unroll
        for (int j = 0; j < NP; ++j) acc2[rh * 3 + rw][j] = ffma2(make_float2(xv, xv), g2[j], acc2[rh * 3 + rw][j]);
      }
  }
  float acc[kTaps * V];
#pragma unroll
  for (int t = 0; t < kTaps; ++t)
#pragma unroll
    for (int e = 0; e < V; ++e) acc[t * V + e] = (e & 1) ? acc2[t][e / 2].y : acc2[t][e / 2].x;
  group_reduce<float, kTaps * V>(acc, cv, smf);
  for (int j = threadIdx.x; j < cv * kTaps * V; j += kThreads) {
    const int gq = j / (kTaps * V), i = j - gq * (kTaps * V);
    const int t = i / V, e = i - t * V;
    atomicAdd(&dw[(gq * V + e) * kTaps + t], group_total(smf, cv, kTaps * V, j));
  }
}

// Run-based weight gradient (C % 8 == 0): a thread owns 8 channels (72 register accumulators) and walks RUN consecutive
// Y pixels of one row; the 3 x WIN window of the one-channel X is loaded once per run (vector loads when contiguous).
template <typename T, int S>
__global__ void __launch_bounds__(kThreads, 2)
wgrad_run_kernel(Geom g, const T* __restrict__ x, int ldx, const T* __restrict__ y, int ldy, float* __restrict__ dw) {
  pdl_wait();
  pdl_launch();
  extern __shared__ float smf[];   // (kThreads/32) * cv * 72 floats
  constexpr int RUN = RunOf<S>::RUN, WIN = RunOf<S>::WIN;
  const int cv = g.C / 8;   // divides 32
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthr = gridDim.x * blockDim.x;
  const int c0 = (tid % cv) * 8;
  const int rpr = (g.yw + RUN - 1) / RUN;
  const int nruns = g.n * g.yh * rpr;
  const int rstep = nthr / cv;
  float2 acc2[kTaps][4];
#pragma unroll
  for (int t = 0; t < kTaps; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc2[t][j] = make_float2(0.f, 0.f);
  for (int run = tid / cv; run < nruns; run += rstep) {
    const int rw = run % rpr;
    const int rr = run / rpr;
    const int h = rr % g.yh, img = rr / g.yh;
    const int w0 = rw * RUN;
    float xv[3][WIN];
    load_window_any<T, S>(g, x, ldx, img, h, w0, xv);
    const T* yrow = y + ((img * g.yh + h) * g.yw + w0) * ldy + c0;
#pragma unroll
    for (int i = 0; i < RUN; ++i) {
      if (w0 + i < g.yw) {
        float gy[8];
        IO<T, 8>::load(yrow + i * ldy, gy);
#pragma unroll
        for (int rh = 0; rh < 3; ++rh)
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const float v = xv[rh][i * S + q];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              acc2[rh * 3 + q][j] = ffma2(make_float2(v, v), make_float2(gy[2 * j], gy[2 * j + 1]), acc2[rh * 3 + q][j]);
          }
      }
    }
  }
  float acc[kTaps * 8];
#pragma unroll
  for (int t = 0; t < kTaps; ++t)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[t * 8 + e] = (e & 1) ? acc2[t][e / 2].y : acc2[t][e / 2].x;
  group_reduce<float, kTaps * 8>(acc, cv, smf);
  for (int j = threadIdx.x; j < cv * kTaps * 8; j += kThreads) {
    const int gq = j / (kTaps * 8), i = j - gq * (kTaps * 8);
    const int t = i / 8, e = i - t * 8;
    atomicAdd(&dw[(gq * 8 + e) * kTaps + t], group_total(smf, cv, kTaps * 8, j));
  }
}

// Weight gradient of the 1 -> 1 stride-1 layer (the UNet's last convolution): dw[t] += sum_p y[p] * x[p - pad + t] over
// two contiguous one-channel bf16 images.  A thread takes runs of 8 pixels: one 16-byte load of y, the 3 x 10 window of x
// as vector loads, nine accumulators; block reduce, nine atomics per block.
__global__ void __launch_bounds__(kThreads, 4)
wgrad11_run_kernel(Geom g, const bf16* __restrict__ x, const bf16* __restrict__ y, float* __restrict__ dw) {
  pdl_wait();
  pdl_launch();
  __shared__ float sred[(kThreads / 32) * kTaps];
  constexpr int RUN = 8, WIN = 10;
  const int rpr = g.yw / RUN;                       // yw % 8 == 0 (host)
  const int nruns = g.n * g.yh * rpr;
  float acc[kTaps];
#pragma unroll
  for (int t = 0; t < kTaps; ++t) acc[t] = 0.f;
  for (int run = blockIdx.x * blockDim.x + threadIdx.x; run < nruns; run += gridDim.x * blockDim.x) {
    const int rw = run % rpr;
    const int rr = run / rpr;
    const int h = rr % g.yh, img = rr / g.yh;
    const int w0 = rw * RUN;
    float xv[3][WIN];
    if (g.pad == 1) load_window_vec<1, 1>(g, x, img, h, w0, xv);
    else load_window_vec<1, 0>(g, x, img, h, w0, xv);
    const uint4 u = *reinterpret_cast<const uint4*>(y + ((size_t)(img * g.yh + h) * g.yw + w0));
    const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < RUN; ++i) {
      const float gy = (i & 1) ? __uint_as_float(uu[i / 2] & 0xffff0000u) : __uint_as_float(uu[i / 2] << 16);
#pragma unroll
      for (int rh = 0; rh < 3; ++rh)
#pragma unroll
        for (int q = 0; q < 3; ++q) acc[rh * 3 + q] = fmaf(gy, xv[rh][i + q], acc[rh * 3 + q]);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < kTaps; ++t) {
    const float v = warp_sum(acc[t]);
    if (lane == 0) sred[warp * kTaps + t] = v;
  }
  __syncthreads();
  if (threadIdx.x < kTaps) {
    float v = 0.f;
#pragma unroll
    for (int wq = 0; wq < kThreads / 32; ++wq) v += sred[wq * kTaps + threadIdx.x];
    atomicAdd(&dw[threadIdx.x], v);
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
  o->vec = 0;
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
static bool run_enabled() {   // MPGAN_NO_C1RUN=1: the pixel-at-a-time kernels (A/B measurements)
  static int v = -1;
  if (v < 0) { const char* e = getenv("MPGAN_NO_C1RUN"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}

// vector width usable for the C-channel tensor: 8 (C % 8 == 0, aligned), 1 (C == 1), 0 = not covered
static int vec_of(int C, const void* p, int64_t ld) {
  if (C == 1) return 1;
  if (C % 8 == 0 && ld % 8 == 0 && aligned16(p)) return 8;
  return 0;
}

// every element offset of a tensor with `pixels` pixels of stride ld must fit in int32
static inline bool fits32(int64_t pixels, int64_t ld) { return pixels * ld < ((int64_t)1 << 31) - 4096; }

#define C1F_LAUNCH(KERNEL, V_, S_, GRID, SMEM, STREAM, ...)                                        \
  do {                                                                                             \
    if (V_ == 8 && S_ == 1) launch_k(KERNEL<T, 8, 1>, GRID, kThreads, SMEM, STREAM, __VA_ARGS__);    \
    else if (V_ == 8) launch_k(KERNEL<T, 8, 2>, GRID, kThreads, SMEM, STREAM, __VA_ARGS__);          \
    else if (S_ == 1) launch_k(KERNEL<T, 1, 1>, GRID, kThreads, SMEM, STREAM, __VA_ARGS__);          \
    else launch_k(KERNEL<T, 1, 2>, GRID, kThreads, SMEM, STREAM, __VA_ARGS__);                       \
  } while (0)

}  // namespace c1f

// Return 0 on success, MPGAN_ERR_* on failure, 1 when the fast path does not cover the call (caller falls back).
// slope (optional, device scalar): PReLU on (conv + bias) -- the inference fusion; only the run-based kernel has it, so
// with a slope the call returns 1 (not covered) instead of falling back to the pixel-at-a-time kernel.
int c1f_fprop(const MpganConvGeom* g, int dtype, const void* x, int64_t ldx, const void* w, const float* bias, void* y,
              int64_t ldy, double* stats, cudaStream_t s, const float* slope) {
  using namespace c1f;
  Geom q;
  if (!geom_ok(g, &q)) return 1;
  const int V = vec_of(q.C, y, ldy);
  if (V == 0) return 1;
  const int cv = q.C / V;
  if (cv <= 32 && (32 % cv) != 0) return 1;
  if (cv > 32 && (kThreads % cv) != 0) return 1;
  const int64_t P = (int64_t)q.n * q.yh * q.yw;
  if (!fits32(P, ldy) || !fits32((int64_t)q.n * q.xh * q.xw, ldx) || P * cv >= ((int64_t)1 << 30)) return 1;
  q.vec = (dtype == MPGAN_BF16 && ldx == 1 && q.xw % 8 == 0 && aligned16(x) && run_enabled()) ? 1 : 0;
  // run-based kernel: a win for stride 1 always, for stride 2 when the window loads are 16-byte vectors
  if (V == 8 && (q.s == 1 || q.vec) && cv <= 32 && (32 % cv) == 0) {
    const int run = q.s == 1 ? 8 : 4;
    const int64_t nruns = (int64_t)q.n * q.yh * ((q.yw + run - 1) / run);
    const int gridr = q.vec ? grid_for(nruns, cv, 4, 2) : grid_for(nruns, cv, 2, 4);   // fewer, fatter threads: the 72-weight
                                                                                      // prologue and the statistics reduce are per thread
    MPGAN_DISPATCH_DTYPE(dtype, T, {
      if (q.s == 1) launch_k(fprop_run_kernel<T, 1>, gridr, kThreads, 0, s, q, (const T*)x, (int)ldx, (const T*)w, bias, (T*)y, (int)ldy, stats, slope);
      else launch_k(fprop_run_kernel<T, 2>, gridr, kThreads, 0, s, q, (const T*)x, (int)ldx, (const T*)w, bias, (T*)y, (int)ldy, stats, slope);
      MPGAN_CHECK_LAUNCH("c1f_fprop_run");
      return 0;
    });
  }
  if (slope) return 1;
  const int grid = grid_for(P, cv, 8, 8);
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    C1F_LAUNCH(fprop_kernel, V, q.s, grid, 0, s, q, (const T*)x, (int)ldx, (const T*)w, bias, (T*)y,
                                                               (int)ldy, stats);
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
  if (!fits32(P, ldx) || !fits32((int64_t)q.n * q.yh * q.yw, ldy) || P * cv >= ((int64_t)1 << 30)) return 1;
  if (q.s == 2 && q.pad == 1 && q.xh == 2 * q.yh && q.xw == 2 * q.yw) {   // Y-centric 2x2-block form
    const int64_t PY = (int64_t)q.n * q.yh * q.yw;
    const int grid2 = grid_for(PY, cv, 2, 8);
    MPGAN_DISPATCH_DTYPE(dtype, T, {
      if (V == 8) launch_k(bprop_s2_kernel<T, 8>, grid2, kThreads, 0, s, q, (const T*)y, (int)ldy, (const T*)w, bias, (T*)x, (int)ldx, stats);
      else launch_k(bprop_s2_kernel<T, 1>, grid2, kThreads, 0, s, q, (const T*)y, (int)ldy, (const T*)w, bias, (T*)x, (int)ldx, stats);
      MPGAN_CHECK_LAUNCH("c1f_bprop_s2");
      return 0;
    });
  }
  const int grid = grid_for(P, cv, 4, 8);
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    C1F_LAUNCH(bprop_kernel, V, q.s, grid, 0, s, q, (const T*)y, (int)ldy, (const T*)w, bias, (T*)x,
                                                               (int)ldx, stats);
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
  if (!fits32(P, ldy) || !fits32((int64_t)q.n * q.xh * q.xw, ldx) || P * cv >= ((int64_t)1 << 30)) return 1;
  const size_t smem = (size_t)(kThreads / 32) * cv * kTaps * V * sizeof(float);
  q.vec = (dtype == MPGAN_BF16 && ldx == 1 && q.xw % 8 == 0 && aligned16(x) && run_enabled()) ? 1 : 0;
  if (V == 1 && q.C == 1 && q.vec && q.s == 1 && ldy == 1 && q.yw % 8 == 0 && aligned16(y) && q.yw == q.xw + 2 * q.pad - 2 &&
      q.xw >= 16) {
    const int64_t nruns = (int64_t)q.n * q.yh * (q.yw / 8);
    int grid11 = (int)ceil_div(nruns, (int64_t)kThreads * 2);
    if (grid11 > num_sms() * 4) grid11 = num_sms() * 4;
    launch_k(wgrad11_run_kernel, grid11, kThreads, 0, s, q, (const bf16*)x, (const bf16*)y, dw);
    MPGAN_CHECK_LAUNCH("c1f_wgrad11_run");
    return 0;
  }
  if (V == 8 && q.vec) {   // run-based kernel with vector window loads
    const int run = q.s == 1 ? 8 : 4;
    const int64_t nruns = (int64_t)q.n * q.yh * ((q.yw + run - 1) / run);
    const int gridr = grid_for(nruns, cv, 4, 2);
    MPGAN_DISPATCH_DTYPE(dtype, T, {
      if (q.s == 1) launch_k(wgrad_run_kernel<T, 1>, gridr, kThreads, smem, s, q, (const T*)x, (int)ldx, (const T*)y, (int)ldy, dw);
      else launch_k(wgrad_run_kernel<T, 2>, gridr, kThreads, smem, s, q, (const T*)x, (int)ldx, (const T*)y, (int)ldy, dw);
      MPGAN_CHECK_LAUNCH("c1f_wgrad_run");
      return 0;
    });
  }
  const int grid = grid_for(P, cv, 32, 4);
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    C1F_LAUNCH(wgrad_kernel, V, q.s, grid, smem, s, q, (const T*)x, (int)ldx, (const T*)y, (int)ldy, dw);
    MPGAN_CHECK_LAUNCH("c1f_wgrad");
    return 0;
  });
}

}  // namespace mpgan
