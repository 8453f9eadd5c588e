// Batch-norm (training + eval) fused with PReLU / LeakyReLU / Tanh and the residual add: bandwidth kernels.
// Replaces nn.BatchNorm{2,3}d + nn.LeakyReLU / nn.PReLU (+ ResidualUnit's add) and their autograd backward
// (/root/reference/code/GAN/GAN_final.py:167-189; MONAI Convolution "norm"/"act", ResidualUnit.forward).
//
// Layout: channels-last (pixels, C) with pixel stride ld.  Vector path: 8 channels per thread (16 B bf16 /
// 32 B f32) when C, ld are multiples of 8 and the pointers are 16/32-byte aligned; scalar path otherwise (C = 1).
// Per-channel sums are accumulated in fp64 end to end (per thread, shared-memory reduce, one global atomic per
// channel per block): E[x^2]-mean^2 must survive |mean| >> std.
#include <string.h>

#include "common.cuh"

namespace mpgan {

template <typename T, int V> struct Vec;
template <> struct Vec<float, 8> {
  static __device__ __forceinline__ void load(const float* p, float* o) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* o) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(o[4], o[5], o[6], o[7]);
  }
};
template <> struct Vec<bf16, 8> {
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
template <> struct Vec<float, 4> {
  static __device__ __forceinline__ void load(const float* p, float* o) {
    float4 a = *reinterpret_cast<const float4*>(p);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* o) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
  }
};
template <> struct Vec<bf16, 4> {
  static __device__ __forceinline__ void load(const bf16* p, float* o) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 2; ++i) { float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
  }
  static __device__ __forceinline__ void store(bf16* p, const float* o) {
    uint2 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 2; ++i) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
    *reinterpret_cast<uint2*>(p) = u;
  }
};
template <typename T> struct Vec<T, 1> {
  static __device__ __forceinline__ void load(const T* p, float* o) { o[0] = to_f(*p); }
  static __device__ __forceinline__ void store(T* p, const float* o) { *p = from_f<T>(o[0]); }
};

__device__ __forceinline__ float act_fwd(float z, int act, float slope) {
  if (act == MPGAN_ACT_PRELU || act == MPGAN_ACT_LEAKY) return z > 0.f ? z : slope * z;
  if (act == MPGAN_ACT_TANH) return tanhf(z);
  return z;
}
// d act / dz given z
__device__ __forceinline__ float act_grad(float z, int act, float slope) {
  if (act == MPGAN_ACT_PRELU || act == MPGAN_ACT_LEAKY) return z > 0.f ? 1.f : slope;
  if (act == MPGAN_ACT_TANH) { float t = tanhf(z); return 1.f - t * t; }
  return 1.f;
}

constexpr int kThreads = 256;
constexpr int kPixPerBlock = 2048;  // pixel run per block for the reductions
constexpr int kMaxPixPerThread = 512;

// ---------------- statistics: sum, sum of squares ----------------
template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
bn_stats_kernel(const T* __restrict__ x, int64_t ldx, int64_t P, int C, double* __restrict__ stats) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  // fp64 partial sums: var = E[x^2] - mean^2 cancels catastrophically in fp32 when |mean| >> std (the un-normalised
  // hand-over between cascaded UNets produces exactly that)
  __shared__ double s1[kThreads * V], s2[kThreads * V];
  const int cv = C / V;
  const int CL = cv < kThreads ? cv : kThreads;   // channel lanes
  const int PL = kThreads / CL;                   // pixel lanes
  const int cl = threadIdx.x % CL, pl = threadIdx.x / CL;
  const int64_t p0 = (int64_t)blockIdx.x * kPixPerBlock;
  const int64_t p1 = min(P, p0 + kPixPerBlock);
  for (int cb = 0; cb < cv; cb += CL) {
    const int vc = cb + cl;
    double a1[V], a2[V];
#pragma unroll
    for (int i = 0; i < V; ++i) a1[i] = a2[i] = 0.0;
    if (pl < PL && vc < cv) {
      for (int64_t p = p0 + pl; p < p1; p += PL) {
        float v[V];
        Vec<T, V>::load(x + p * ldx + (int64_t)vc * V, v);
#pragma unroll
        for (int i = 0; i < V; ++i) { a1[i] += (double)v[i]; a2[i] = fma((double)v[i], (double)v[i], a2[i]); }
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) { s1[threadIdx.x * V + i] = a1[i]; s2[threadIdx.x * V + i] = a2[i]; }
    __syncthreads();
    // CL*V channel slots; slot j owned by thread j (< kThreads*V/PL ...): loop
    for (int j = threadIdx.x; j < CL * V; j += kThreads) {
      int lane = j / V, e = j % V;
      if (cb + lane < cv) {
        double t1 = 0.0, t2 = 0.0;
        for (int q = 0; q < PL; ++q) { t1 += s1[(q * CL + lane) * V + e]; t2 += s2[(q * CL + lane) * V + e]; }
        int ch = (cb + lane) * V + e;
        atomicAdd(&stats[ch], t1);
        atomicAdd(&stats[C + ch], t2);
      }
    }
    __syncthreads();
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats, int64_t P, int C, const float* gamma,
                                   const float* beta, float eps, float momentum, int training,
                                   float* running_mean, float* running_var, int64_t* nbt, float* mean_out,
                                   float* invstd_out, float* scale, float* shift) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && training && nbt) *nbt += 1;
  if (c >= C) return;
  float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  float mean, invstd;
  if (training) {
    double m = stats[c] / (double)P;
    double var = stats[C + c] / (double)P - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    invstd = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    if (running_var) {
      double unb = P > 1 ? var * ((double)P / (double)(P - 1)) : var;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
  } else {
    mean = running_mean[c];
    invstd = 1.f / sqrtf(running_var[c] + eps);
  }
  if (mean_out) mean_out[c] = mean;
  if (invstd_out) invstd_out[c] = invstd;
  float sc = g * invstd;
  scale[c] = sc;
  shift[c] = b - mean * sc;
}

// The hot kernels below share one mapping: the grid is sized so that (gridDim*blockDim) is a multiple of the
// number of channel vectors per pixel (cv), hence every thread keeps ONE channel group for its whole pixel walk and
// the per-channel constants live in registers; pixels are walked with a 4x (apply: 2x) unroll for memory-level
// parallelism.  Per-channel reductions: per-thread fp64 partials -> warp shuffles across the lanes that share a
// channel group -> one shared-memory slot per (warp, channel) -> one global fp64 atomic per channel per block.

// scale/shift (and the saved mean/invstd) straight from the fp64 statistics: the arithmetic of bn_finalize_kernel
struct BnTrain {
  const double* stats;       // [2C] sum, sum of squares; nullptr = scale/shift given
  const float* gamma;
  const float* beta;
  float eps, momentum;
  float* running_mean;
  float* running_var;
  int64_t* nbt;
  float* mean_out;           // [C] each; written by the first block
  float* invstd_out;
  float* scale_out;
  float* shift_out;
};

__device__ __forceinline__ void bn_train_coeffs(const BnTrain& f, int c, int C, int64_t P, bool writer, float& sc,
                                                float& sh) {
  const double m = f.stats[c] / (double)P;
  double var = f.stats[C + c] / (double)P - m * m;
  if (var < 0.0) var = 0.0;
  const float mean = (float)m;
  const float invstd = (float)(1.0 / sqrt(var + (double)f.eps));
  const float g = f.gamma ? f.gamma[c] : 1.f, b = f.beta ? f.beta[c] : 0.f;
  sc = g * invstd;
  sh = b - mean * sc;
  if (writer) {
    if (f.running_mean) f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * mean;
    if (f.running_var) {
      double unb = P > 1 ? var * ((double)P / (double)(P - 1)) : var;
      f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * (float)unb;
    }
    if (f.mean_out) f.mean_out[c] = mean;
    if (f.invstd_out) f.invstd_out[c] = invstd;
    f.scale_out[c] = sc;
    f.shift_out[c] = sh;
  }
}

// ---------------- apply: y = act(x*scale + shift) (+ res) ----------------
template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
bn_act_apply_kernel(const T* __restrict__ x, int64_t ldx, int64_t P, int C, const float* __restrict__ scale,
                    const float* __restrict__ shift, const BnTrain f, int act, const float* __restrict__ alpha,
                    float leaky, const T* __restrict__ res, int64_t ldres, T* __restrict__ y, int64_t ldy) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  const int cv = C / V;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
  const int c0 = (int)(tid % cv) * V;
  const int64_t pstep = nthr / cv;
  const float slope = (act == MPGAN_ACT_PRELU || (act == MPGAN_ACT_LEAKY && alpha)) ? *alpha : leaky;
  float sc[V], sh[V];
  if (f.stats) {
    // one channel per thread -> shared memory (the fp64 divide / sqrt chain runs once per channel per block)
    extern __shared__ float s_coef[];   // [2*C]
    if (tid == 0 && f.nbt) *f.nbt += 1;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float a, b;
      bn_train_coeffs(f, c, C, P, blockIdx.x == 0, a, b);
      s_coef[c] = a;
      s_coef[C + c] = b;
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < V; ++e) { sc[e] = s_coef[c0 + e]; sh[e] = s_coef[C + c0 + e]; }
  } else {
#pragma unroll
    for (int e = 0; e < V; ++e) { sc[e] = scale ? scale[c0 + e] : 1.f; sh[e] = shift ? shift[c0 + e] : 0.f; }
  }
  int64_t p = tid / cv;
  for (; p + pstep < P; p += 2 * pstep) {
    float v0[V], v1[V], r0[V], r1[V];
    Vec<T, V>::load(x + p * ldx + c0, v0);
    Vec<T, V>::load(x + (p + pstep) * ldx + c0, v1);
    if (res) { Vec<T, V>::load(res + p * ldres + c0, r0); Vec<T, V>::load(res + (p + pstep) * ldres + c0, r1); }
#pragma unroll
    for (int e = 0; e < V; ++e) {
      float z0 = act_fwd(fmaf(v0[e], sc[e], sh[e]), act, slope), z1 = act_fwd(fmaf(v1[e], sc[e], sh[e]), act, slope);
      if (res) { z0 += r0[e]; z1 += r1[e]; }
      v0[e] = z0; v1[e] = z1;
    }
    Vec<T, V>::store(y + p * ldy + c0, v0);
    Vec<T, V>::store(y + (p + pstep) * ldy + c0, v1);
  }
  if (p < P) {
    float v0[V], r0[V];
    Vec<T, V>::load(x + p * ldx + c0, v0);
    if (res) Vec<T, V>::load(res + p * ldres + c0, r0);
#pragma unroll
    for (int e = 0; e < V; ++e) {
      float z0 = act_fwd(fmaf(v0[e], sc[e], sh[e]), act, slope);
      if (res) z0 += r0[e];
      v0[e] = z0;
    }
    Vec<T, V>::store(y + p * ldy + c0, v0);
  }
}

// Block-level reduction of NV per-thread fp64 partials that belong to channel group (tid % cv); the totals of
// value i of group g are added to out[i_hi * C + g*V + i_lo] (i = i_hi*V + i_lo) with one global atomic each.
// sm: kThreads*NV doubles.
template <int V, int NV, typename A, typename OutT>
__device__ __forceinline__ void group_reduce_to_global(A (&a)[NV], int cv, int C, int64_t tid, A* sm,
                                                       OutT* __restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (cv <= 32 && (32 % cv) == 0) {
    for (int off = 16; off >= cv; off >>= 1) {
#pragma unroll
      for (int i = 0; i < NV; ++i) a[i] += __shfl_xor_sync(0xffffffffu, a[i], off);
    }
    if (lane < cv) {
#pragma unroll
      for (int i = 0; i < NV; ++i) sm[(warp * cv + lane) * NV + i] = a[i];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < cv * NV; j += kThreads) {
      const int g = j / NV, i = j - g * NV;
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) t += (double)sm[(w * cv + g) * NV + i];
      atomicAdd(&out[(i / V) * C + g * V + (i % V)], (OutT)t);
    }
  } else if ((kThreads % cv) == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) sm[threadIdx.x * NV + i] = a[i];
    __syncthreads();
    const int reps = kThreads / cv;
    for (int j = threadIdx.x; j < cv * NV; j += kThreads) {
      const int g = j / NV, i = j - g * NV;
      double t = 0.0;
      for (int k = 0; k < reps; ++k) t += (double)sm[(g + k * cv) * NV + i];
      atomicAdd(&out[(i / V) * C + g * V + (i % V)], (OutT)t);
    }
  } else {  // channel group varies with the block: straight to global (odd channel counts; tests only)
    const int g = (int)(tid % cv);
#pragma unroll
    for (int i = 0; i < NV; ++i) atomicAdd(&out[(i / V) * C + g * V + (i % V)], (OutT)a[i]);
  }
}

// ---------------- backward pass 1: reductions ----------------
template <typename T, int V>
__global__ void __launch_bounds__(kThreads, 3)
bn_act_bwd_reduce_kernel(const T* __restrict__ dy, int64_t lddy, const T* __restrict__ x, int64_t ldx, int64_t P,
                         int C, const float* __restrict__ mean, const float* __restrict__ invstd,
                         const float* __restrict__ scale, const float* __restrict__ shift, int act,
                         const float* __restrict__ alpha, float leaky, double* __restrict__ sums) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  // per-thread partials are fp32 (<= kMaxPixPerThread terms each, no cancellation structure in these sums),
  // everything above the block level is fp64
  __shared__ float sm[kThreads * 2 * V];
  __shared__ double sslope[kThreads / 32];
  const int cv = C / V;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
  const int c0 = (int)(tid % cv) * V;
  const int64_t pstep = nthr / cv;
  const float slope = (act == MPGAN_ACT_PRELU || (act == MPGAN_ACT_LEAKY && alpha)) ? *alpha : leaky;
  float sc[V], sh[V], mu[V], is[V];
#pragma unroll
  for (int e = 0; e < V; ++e) {
    sc[e] = scale ? scale[c0 + e] : 1.f; sh[e] = shift ? shift[c0 + e] : 0.f;
    mu[e] = mean ? mean[c0 + e] : 0.f; is[e] = invstd ? invstd[c0 + e] : 1.f;
  }
  float a[2 * V], fs = 0.f;
#pragma unroll
  for (int e = 0; e < 2 * V; ++e) a[e] = 0.f;
  constexpr int U = V == 4 ? 4 : 2;
  for (int64_t p = tid / cv; p < P; p += U * pstep) {
    float g[U][V], xv[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (p + u * pstep < P) {
        Vec<T, V>::load(dy + (p + u * pstep) * lddy + c0, g[u]);
        Vec<T, V>::load(x + (p + u * pstep) * ldx + c0, xv[u]);
      } else {
#pragma unroll
        for (int e = 0; e < V; ++e) { g[u][e] = 0.f; xv[u][e] = 0.f; }
      }
    }
#pragma unroll
    for (int e = 0; e < V; ++e) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float z = fmaf(xv[u][e], sc[e], sh[e]);
        if (act == MPGAN_ACT_PRELU && z <= 0.f) fs = fmaf(g[u][e], z, fs);
        const float gz = g[u][e] * act_grad(z, act, slope);
        a[e] += gz;
        a[V + e] = fmaf(gz, (xv[u][e] - mu[e]) * is[e], a[V + e]);
      }
    }
  }
  group_reduce_to_global<V, 2 * V, float, double>(a, cv, C, tid, sm, sums);
  if (act == MPGAN_ACT_PRELU) {
    double w = warp_sum((double)fs);
    if ((threadIdx.x & 31) == 0) sslope[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
#pragma unroll
      for (int i = 0; i < kThreads / 32; ++i) t += sslope[i];
      atomicAdd(&sums[2 * C], t);
    }
  }
}

// ---------------- backward pass 2: dx, parameter grads, and (optionally) the conv bias grad = sum_p dx ----------------
// dx = k1*gz + k2*(x - mu) + k3 with k1 = scale, k2 = -scale*invstd*mean(g*xhat), k3 = -scale*mean(g)
template <typename T, int V>
__global__ void __launch_bounds__(kThreads, 3)
bn_act_bwd_apply_kernel(const T* __restrict__ dy, int64_t lddy, const T* __restrict__ x, int64_t ldx, int64_t P,
                        int C, const float* __restrict__ mean, const float* __restrict__ invstd,
                        const float* __restrict__ scale, const float* __restrict__ shift, int act,
                        const float* __restrict__ alpha, float leaky, const double* __restrict__ sums,
                        float* dgamma, float* dbeta, float* dalpha, float* dbias, T* __restrict__ dx, int64_t lddx) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  __shared__ float sm[kThreads * V];
  const int cv = C / V;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
  const int c0 = (int)(tid % cv) * V;
  const int64_t pstep = nthr / cv;
  const float slope = (act == MPGAN_ACT_PRELU || (act == MPGAN_ACT_LEAKY && alpha)) ? *alpha : leaky;
  const float invP = 1.f / (float)P;
  if (blockIdx.x == 0) {  // parameter gradients (accumulate), once
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dbeta) dbeta[c] += (float)sums[c];
      if (dgamma) dgamma[c] += (float)sums[C + c];
    }
    if (threadIdx.x == 0 && dalpha && act == MPGAN_ACT_PRELU) *dalpha += (float)sums[2 * C];
  }
  const bool train = mean != nullptr;
  float sc[V], sh[V], mu[V], k2[V], k3[V];
#pragma unroll
  for (int e = 0; e < V; ++e) {
    sc[e] = scale ? scale[c0 + e] : 1.f; sh[e] = shift ? shift[c0 + e] : 0.f;
    mu[e] = train ? mean[c0 + e] : 0.f;
    const float is = train ? invstd[c0 + e] : 1.f;
    const float mg = train ? (float)sums[c0 + e] * invP : 0.f;
    const float mgx = train ? (float)sums[C + c0 + e] * invP : 0.f;
    k2[e] = -sc[e] * is * mgx;
    k3[e] = -sc[e] * mg;
  }
  float bsum[V];
#pragma unroll
  for (int e = 0; e < V; ++e) bsum[e] = 0.f;
  constexpr int U = V == 4 ? 4 : 2;
  for (int64_t p = tid / cv; p < P; p += U * pstep) {
    float g[U][V], xv[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (p + u * pstep < P) {
        Vec<T, V>::load(dy + (p + u * pstep) * lddy + c0, g[u]);
        Vec<T, V>::load(x + (p + u * pstep) * ldx + c0, xv[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (p + u * pstep < P) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const float z = fmaf(xv[u][e], sc[e], sh[e]);
          g[u][e] = fmaf(sc[e], g[u][e] * act_grad(z, act, slope), fmaf(k2[e], xv[u][e] - mu[e], k3[e]));
        }
        Vec<T, V>::store(dx + (p + u * pstep) * lddx + c0, g[u]);
        if (dbias) {  // bias gradient of the producing conv: column sum of dx as stored (rounded to T)
#pragma unroll
          for (int e = 0; e < V; ++e) bsum[e] += to_f(from_f<T>(g[u][e]));
        }
      }
    }
  }
  if (dbias) group_reduce_to_global<V, V, float, float>(bsum, cv, C, tid, sm, dbias);
}

static inline bool vec_ok(const void* p, int64_t ld, int dtype) {
  size_t align = dtype == MPGAN_F32 ? 16 : 16;
  return p == nullptr || (((uintptr_t)p % align) == 0 && (ld % 8) == 0);
}

static inline int64_t gcd64(int64_t a, int64_t b) { while (b) { int64_t t = a % b; a = b; b = t; } return a; }

// grid for the fixed-channel-group mapping: gridDim*kThreads must be a multiple of cv.
// pix_per_thread: pixels each thread should walk (amortises the per-thread prologue / reduction epilogue).
static inline int ew_grid(int64_t pixels, int cv, int blocks_per_sm = 8, int pix_per_thread = 2) {
  int64_t total = pixels * cv;
  int64_t b = ceil_div(total, (int64_t)kThreads * pix_per_thread);
  int64_t cap = (int64_t)num_sms() * blocks_per_sm;
  const int64_t floor_b = ceil_div(total, (int64_t)kThreads * kMaxPixPerThread);   // bounds the fp32 per-thread partials
  if (cap < floor_b) cap = floor_b;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  int64_t mult = cv / gcd64(cv, kThreads);
  return (int)(ceil_div(b, mult) * mult);
}

// bn_stream.cu: bulk-copy streaming variants for contiguous bf16 tensors
bool bn_stream_ok(int dtype, const void* in0, const void* in1, const void* out, int64_t ld0, int64_t ld1, int64_t ldo,
                  int64_t pixels, int32_t C, int act);
int bn_stream_fwd(const void* x, int64_t pixels, int32_t C, const float* scale, const float* shift,
                  const double* stats, const float* gamma, const float* beta, float eps, float momentum,
                  float* running_mean, float* running_var, int64_t* nbt, float* mean, float* invstd, float* scale_out,
                  float* shift_out, int act, const float* alpha, float slope, const void* res, void* y, int64_t ldy,
                  cudaStream_t s);
int bn_stream_reduce(const void* dy, const void* x, int64_t pixels, int32_t C, const float* mean, const float* invstd,
                     const float* scale, const float* shift, int act, const float* alpha, float slope, double* sums,
                     cudaStream_t s);
int bn_stream_bwd(const void* dy, const void* x, int64_t pixels, int32_t C, const float* mean, const float* invstd,
                  const float* scale, const float* shift, int act, const float* alpha, float slope, const double* sums,
                  float* dgamma, float* dbeta, float* dalpha, float* dbias, void* dx, int64_t lddx, cudaStream_t s);

}  // namespace mpgan

using namespace mpgan;

extern "C" int mpgan_bn_stats(int dtype, const void* x, int64_t ldx, int64_t pixels, int32_t c, double* stats,
                              void* stream) {
  MPGAN_REQUIRE(pixels > 0 && c > 0 && ldx >= c, MPGAN_ERR_SHAPE, "bn_stats: bad shape");
  const bool vec = (c % 8 == 0) && vec_ok(x, ldx, dtype);
  const int grid = (int)ceil_div(pixels, kPixPerBlock);
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    if (vec) launch_k(bn_stats_kernel<T, 8>, grid, kThreads, 0, (cudaStream_t)stream, (const T*)x, ldx, pixels, c, stats);
    else launch_k(bn_stats_kernel<T, 1>, grid, kThreads, 0, (cudaStream_t)stream, (const T*)x, ldx, pixels, c, stats);
    MPGAN_CHECK_LAUNCH("bn_stats");
    return 0;
  });
}

extern "C" int mpgan_bn_finalize(const double* stats, int64_t pixels, int32_t c, const float* gamma,
                                 const float* beta, float eps, float momentum, int training, float* running_mean,
                                 float* running_var, int64_t* num_batches_tracked, float* mean, float* invstd,
                                 float* scale, float* shift, void* stream) {
  MPGAN_REQUIRE(c > 0 && scale && shift, MPGAN_ERR_SHAPE, "bn_finalize: bad arguments");
  MPGAN_REQUIRE(training ? (stats != nullptr && pixels > 0) : (running_mean && running_var), MPGAN_ERR_SHAPE,
                "bn_finalize: missing statistics");
  launch_k(bn_finalize_kernel, (c + 127) / 128, 128, 0, (cudaStream_t)stream, 
      stats, pixels, c, gamma, beta, eps, momentum, training, running_mean, running_var, num_batches_tracked, mean,
      invstd, scale, shift);
  MPGAN_CHECK_LAUNCH("bn_finalize");
  return 0;
}

template <typename T>
static int launch_apply(const void* x, int64_t ldx, int64_t pixels, int32_t c, const float* scale, const float* shift,
                        const BnTrain& f, int act, const float* alpha, float leaky_slope, const void* res,
                        int64_t ldres, void* y, int64_t ldy, bool vec, cudaStream_t s) {
  const size_t smem = f.stats ? (size_t)2 * c * sizeof(float) : 0;
  if (vec)
    launch_k(bn_act_apply_kernel<T, 8>, ew_grid(pixels, c / 8, 8, 4), kThreads, smem, s, 
        (const T*)x, ldx, pixels, c, scale, shift, f, act, alpha, leaky_slope, (const T*)res, ldres, (T*)y, ldy);
  else
    launch_k(bn_act_apply_kernel<T, 1>, ew_grid(pixels, c, 8, 4), kThreads, smem, s, 
        (const T*)x, ldx, pixels, c, scale, shift, f, act, alpha, leaky_slope, (const T*)res, ldres, (T*)y, ldy);
  MPGAN_CHECK_LAUNCH("bn_act_apply");
  return 0;
}

extern "C" int mpgan_bn_act_apply(int dtype, const void* x, int64_t ldx, int64_t pixels, int32_t c,
                                  const float* scale, const float* shift, int act, const float* alpha,
                                  float leaky_slope, const void* res, int64_t ldres, void* y, int64_t ldy,
                                  void* stream) {
  MPGAN_REQUIRE(pixels > 0 && c > 0 && ldx >= c && ldy >= c, MPGAN_ERR_SHAPE, "bn_act_apply: bad shape");
  MPGAN_REQUIRE((scale == nullptr) == (shift == nullptr), MPGAN_ERR_SHAPE, "scale/shift must both be given");
  MPGAN_REQUIRE(act != MPGAN_ACT_PRELU || alpha, MPGAN_ERR_SHAPE, "PReLU needs alpha");
  if (bn_stream_ok(dtype, x, res, y, ldx, ldres, ldy, pixels, c, act))   // contiguous bf16: bulk-copy streaming kernel
    return bn_stream_fwd(x, pixels, c, scale, shift, nullptr, nullptr, nullptr, 0.f, 0.f, nullptr, nullptr, nullptr, nullptr,
                         nullptr, nullptr, nullptr, act, alpha, leaky_slope, res, y, ldy, (cudaStream_t)stream);
  const bool vec = (c % 8 == 0) && vec_ok(x, ldx, dtype) && vec_ok(y, ldy, dtype) && vec_ok(res, res ? ldres : 8, dtype);
  BnTrain f;
  memset(&f, 0, sizeof(f));
  MPGAN_DISPATCH_DTYPE(dtype, T, return launch_apply<T>(x, ldx, pixels, c, scale, shift, f, act, alpha, leaky_slope,
                                                        res, ldres, y, ldy, vec, (cudaStream_t)stream));
}

extern "C" int mpgan_bn_train_apply(int dtype, const void* x, int64_t ldx, int64_t pixels, int32_t c,
                                    const double* stats, const float* gamma, const float* beta, float eps,
                                    float momentum, float* running_mean, float* running_var,
                                    int64_t* num_batches_tracked, float* mean, float* invstd, float* scale,
                                    float* shift, int act, const float* alpha, float leaky_slope, const void* res,
                                    int64_t ldres, void* y, int64_t ldy, void* stream) {
  MPGAN_REQUIRE(pixels > 0 && c > 0 && ldx >= c && ldy >= c, MPGAN_ERR_SHAPE, "bn_train_apply: bad shape");
  MPGAN_REQUIRE(stats && scale && shift, MPGAN_ERR_SHAPE, "bn_train_apply: stats, scale and shift are required");
  MPGAN_REQUIRE(act != MPGAN_ACT_PRELU || alpha, MPGAN_ERR_SHAPE, "PReLU needs alpha");
  if (bn_stream_ok(dtype, x, res, y, ldx, ldres, ldy, pixels, c, act))   // contiguous bf16: bulk-copy streaming kernel
    return bn_stream_fwd(x, pixels, c, nullptr, nullptr, stats, gamma, beta, eps, momentum, running_mean, running_var,
                         num_batches_tracked, mean, invstd, scale, shift, act, alpha, leaky_slope, res, y, ldy,
                         (cudaStream_t)stream);
  const bool vec = (c % 8 == 0) && vec_ok(x, ldx, dtype) && vec_ok(y, ldy, dtype) && vec_ok(res, res ? ldres : 8, dtype);
  BnTrain f;
  f.stats = stats; f.gamma = gamma; f.beta = beta; f.eps = eps; f.momentum = momentum;
  f.running_mean = running_mean; f.running_var = running_var; f.nbt = num_batches_tracked;
  f.mean_out = mean; f.invstd_out = invstd; f.scale_out = scale; f.shift_out = shift;
  MPGAN_DISPATCH_DTYPE(dtype, T, return launch_apply<T>(x, ldx, pixels, c, nullptr, nullptr, f, act, alpha,
                                                        leaky_slope, res, ldres, y, ldy, vec, (cudaStream_t)stream));
}

namespace mpgan {

// ------------------------------------------------------------------------------------------------------------
// Tail of every UNet of the generator (MONAI UNet top level, is_top: ConvT(32->1) -> BatchNorm(1) -> PReLU ->
// ResidualUnit(1->1, conv only, identity residual)):   h = prelu(bn(c));  y = conv3x3(h) + bias + h
// on a ONE-channel full-resolution image.  Three launches (BatchNorm apply, one-channel convolution, residual add)
// of 10-17 us each on a 4 MB tensor become one stencil kernel: the BatchNorm coefficients come from the fp64 statistics
// that the ConvTranspose epilogue reduced, each thread transforms its 3 x 6 window of c on the fly (h is rounded to
// bf16 exactly as the unfused path stores it) and produces 4 output pixels; h itself is written only when the
// backward pass needs it.  Block 0 updates the running statistics and the saved mean / invstd / scale / shift.
// ------------------------------------------------------------------------------------------------------------
// one window row of a contiguous one-channel bf16 image: columns w0 - 1 .. w0 + RUN, zero outside the image.  RUN == 8
// (W % 8 == 0, 16-byte aligned image): the eight inner columns are one vector load.
template <int RUN>
__device__ __forceinline__ void tail_load_row(const bf16* __restrict__ img, int y, int w0, int H, int W, float (&v)[RUN + 2]) {
  const bool oky = (unsigned)y < (unsigned)H;
  const bf16* row = img + (int64_t)(oky ? y : 0) * W;
  if constexpr (RUN == 8) {
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (oky) u = *reinterpret_cast<const uint4*>(row + w0);
    const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[1 + 2 * e] = __uint_as_float(uu[e] << 16);
      v[2 + 2 * e] = __uint_as_float(uu[e] & 0xffff0000u);
    }
    v[0] = (oky && w0 > 0) ? to_f(row[w0 - 1]) : 0.f;
    v[9] = (oky && w0 + 8 < W) ? to_f(row[w0 + 8]) : 0.f;
  } else {
#pragma unroll
    for (int j = 0; j < RUN + 2; ++j) {
      const int x = w0 - 1 + j;
      v[j] = (oky && (unsigned)x < (unsigned)W) ? to_f(row[x]) : 0.f;
    }
  }
}

template <int RUN>
__global__ void __launch_bounds__(kThreads)
c1_tail_fwd_kernel(const bf16* __restrict__ c, int n, int H, int W, const BnTrain f, const float* __restrict__ alpha,
                   const bf16* __restrict__ w9, const float* __restrict__ bias, bf16* __restrict__ h_out,
                   bf16* __restrict__ y_out) {
  pdl_wait();
  pdl_launch();
  const int64_t P = (int64_t)n * H * W;
  float sc, sh;
  if (f.stats) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && f.nbt) *f.nbt += 1;
    bn_train_coeffs(f, 0, 1, P, blockIdx.x == 0 && threadIdx.x == 0, sc, sh);
  } else {          // evaluation mode: scale / shift of the running statistics are given (mpgan_bn_finalize)
    sc = f.scale_out[0];
    sh = f.shift_out[0];
  }
  const float slope = *alpha;
  float wt[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) wt[t] = to_f(w9[t]);
  const float b = bias ? bias[0] : 0.f;
  const int wq = (W + RUN - 1) / RUN;               // RUN-pixel runs per row
  const int64_t runs = (int64_t)n * H * wq;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < runs; r += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(r % wq);
    const int64_t row = r / wq;
    const int hh = (int)(row % H), img = (int)(row / H);
    const int w0 = q * RUN;
    const bf16* ci = c + (int64_t)img * H * W;
    float hv[3][RUN + 2];
#pragma unroll
    for (int rh = 0; rh < 3; ++rh) {
      const int y = hh - 1 + rh;
      tail_load_row<RUN>(ci, y, w0, H, W, hv[rh]);
      const bool oky = (unsigned)y < (unsigned)H;
#pragma unroll
      for (int j = 0; j < RUN + 2; ++j) {
        const int x = w0 - 1 + j;
        float v = 0.f;
        if (oky && (unsigned)x < (unsigned)W) {
          const float z = fmaf(hv[rh][j], sc, sh);
          v = to_f(from_f<bf16>(z > 0.f ? z : slope * z));     // h as the unfused path stores it
        }
        hv[rh][j] = v;
      }
    }
    const int64_t o0 = ((int64_t)img * H + hh) * W + w0;
    if constexpr (RUN == 8) {
      uint32_t hp[4], yp[4];
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        float a0 = b, a1 = b;
#pragma unroll
        for (int rh = 0; rh < 3; ++rh)
#pragma unroll
          for (int rw = 0; rw < 3; ++rw) {
            a0 = fmaf(hv[rh][i + rw], wt[rh * 3 + rw], a0);
            a1 = fmaf(hv[rh][i + 1 + rw], wt[rh * 3 + rw], a1);
          }
        const float h0 = hv[1][i + 1], h1 = hv[1][i + 2];
        __nv_bfloat162 hh2 = __floats2bfloat162_rn(h0, h1);
        __nv_bfloat162 yy2 = __floats2bfloat162_rn(to_f(from_f<bf16>(a0)) + h0, to_f(from_f<bf16>(a1)) + h1);
        hp[i / 2] = *reinterpret_cast<uint32_t*>(&hh2);
        yp[i / 2] = *reinterpret_cast<uint32_t*>(&yy2);
      }
      if (h_out) *reinterpret_cast<uint4*>(h_out + o0) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
      *reinterpret_cast<uint4*>(y_out + o0) = make_uint4(yp[0], yp[1], yp[2], yp[3]);
    } else {
#pragma unroll
      for (int i = 0; i < RUN; ++i) {
        if (w0 + i >= W) break;
        float a = b;
#pragma unroll
        for (int rh = 0; rh < 3; ++rh)
#pragma unroll
          for (int rw = 0; rw < 3; ++rw) a = fmaf(hv[rh][i + rw], wt[rh * 3 + rw], a);
        const float hc = hv[1][i + 1];
        if (h_out) h_out[o0 + i] = from_f<bf16>(hc);
        y_out[o0 + i] = from_f<bf16>(to_f(from_f<bf16>(a)) + hc);     // conv result rounded, then the residual add (two kernels before)
      }
    }
  }
}


// Backward twin of c1_tail_fwd_kernel: for the one-channel tail  y = conv3x3(h) + bias + h,  h = prelu(bn(c)):
//   dh = conv3x3^T(dy) + dy        (data gradient of the 1->1 conv plus the identity residual, rounded to bf16 at the same
//                                    two points as the unfused kernels)
// and, in the same pass, the BatchNorm(1) + PReLU backward REDUCTION of the layer below it:
//   sums[0] += sum dz,  sums[1] += sum dz * xhat,  sums[2] += sum_{z<=0} dh * z      with z = c*scale + shift, dz = dh * prelu'(z)
// (replaces c1f::bprop_kernel<1,1> + add_copy + bn_act_bwd_reduce_kernel<1>: three launches on the backward critical path).
template <int RUN>
__global__ void __launch_bounds__(kThreads)
c1_tail_bwd_reduce_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ c, int n, int H, int W,
                          const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ scale,
                          const float* __restrict__ shift, const float* __restrict__ alpha, const bf16* __restrict__ w9,
                          bf16* __restrict__ dh_out, double* __restrict__ sums) {
  pdl_wait();
  pdl_launch();
  __shared__ double red[3][kThreads / 32];
  const float sc = scale[0], sh = shift[0], mu = mean[0], is = invstd[0], slope = alpha[0];
  float wt[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) wt[t] = to_f(w9[8 - t]);      // flipped taps: dh[q] = sum_t w[8 - t] * dy[q - 1 + t]
  const int wq = (W + RUN - 1) / RUN;
  const int64_t runs = (int64_t)n * H * wq;
  float a0 = 0.f, a1 = 0.f, fs = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < runs; r += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(r % wq);
    const int64_t row = r / wq;
    const int hh = (int)(row % H), img = (int)(row / H);
    const int w0 = q * RUN;
    const bf16* di = dy + (int64_t)img * H * W;
    float dv[3][RUN + 2];
#pragma unroll
    for (int rh = 0; rh < 3; ++rh) tail_load_row<RUN>(di, hh - 1 + rh, w0, H, W, dv[rh]);
#pragma unroll
    for (int i = 0; i < RUN; ++i) {
      if (w0 + i >= W) break;
      float acc = 0.f;
#pragma unroll
      for (int rh = 0; rh < 3; ++rh)
#pragma unroll
        for (int rw = 0; rw < 3; ++rw) acc = fmaf(dv[rh][i + rw], wt[rh * 3 + rw], acc);
      const bf16 dhb = from_f<bf16>(to_f(from_f<bf16>(acc)) + dv[1][i + 1]);
      const int64_t o = ((int64_t)img * H + hh) * W + w0 + i;
      dh_out[o] = dhb;
      const float g = to_f(dhb), cv = to_f(c[o]);
      const float z = fmaf(cv, sc, sh);
      if (z <= 0.f) fs = fmaf(g, z, fs);
      const float gz = g * (z > 0.f ? 1.f : slope);
      a0 += gz;
      a1 = fmaf(gz, (cv - mu) * is, a1);
    }
  }
  double r0 = warp_sum((double)a0), r1 = warp_sum((double)a1), r2 = warp_sum((double)fs);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = r0; red[1][threadIdx.x >> 5] = r1; red[2][threadIdx.x >> 5] = r2; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) t += red[threadIdx.x][i];
    atomicAdd(&sums[threadIdx.x], t);
  }
}

}  // namespace mpgan

extern "C" int mpgan_c1_tail_fwd(const void* c_bf16, int32_t n, int32_t h, int32_t w, const double* stats,
                                 const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                                 float* running_var, int64_t* num_batches_tracked, float* mean, float* invstd,
                                 float* scale, float* shift, const float* alpha, const void* w9_bf16, const float* bias,
                                 void* h_out_bf16, void* y_out_bf16, void* stream) {
  // stats == nullptr: evaluation mode, scale / shift are INPUTS and no running statistic is touched
  MPGAN_REQUIRE(c_bf16 && scale && shift && alpha && w9_bf16 && y_out_bf16, MPGAN_ERR_SHAPE, "c1_tail_fwd: null pointer");
  MPGAN_REQUIRE(n > 0 && h > 0 && w > 0, MPGAN_ERR_SHAPE, "c1_tail_fwd: empty tensor");
  BnTrain f;
  f.stats = stats; f.gamma = gamma; f.beta = beta; f.eps = eps; f.momentum = momentum;
  f.running_mean = running_mean; f.running_var = running_var; f.nbt = num_batches_tracked;
  f.mean_out = mean; f.invstd_out = invstd; f.scale_out = scale; f.shift_out = shift;
  // rows of 8k pixels in 16-byte aligned images: 8-pixel runs with vector loads / stores
  const bool vec = w % 8 == 0 && ((uintptr_t)c_bf16 & 15) == 0 && ((uintptr_t)y_out_bf16 & 15) == 0 &&
                   ((uintptr_t)h_out_bf16 & 15) == 0;
  const int64_t runs = (int64_t)n * h * ((w + (vec ? 7 : 3)) / (vec ? 8 : 4));
  int64_t blocks = ceil_div(runs, (int64_t)kThreads);
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (vec)
    launch_k(c1_tail_fwd_kernel<8>, (int)blocks, kThreads, 0, (cudaStream_t)stream, (const bf16*)c_bf16, (int)n, (int)h, (int)w, f,
             alpha, (const bf16*)w9_bf16, bias, (bf16*)h_out_bf16, (bf16*)y_out_bf16);
  else
    launch_k(c1_tail_fwd_kernel<4>, (int)blocks, kThreads, 0, (cudaStream_t)stream, (const bf16*)c_bf16, (int)n, (int)h, (int)w, f,
             alpha, (const bf16*)w9_bf16, bias, (bf16*)h_out_bf16, (bf16*)y_out_bf16);
  MPGAN_CHECK_LAUNCH("c1_tail_fwd_kernel");
  return 0;
}

extern "C" int mpgan_c1_tail_bwd_reduce(const void* dy_bf16, const void* c_bf16, int32_t n, int32_t h, int32_t w,
                                        const float* mean, const float* invstd, const float* scale, const float* shift,
                                        const float* alpha, const void* w9_bf16, void* dh_bf16, double* sums3,
                                        void* stream) {
  MPGAN_REQUIRE(dy_bf16 && c_bf16 && mean && invstd && scale && shift && alpha && w9_bf16 && dh_bf16 && sums3,
                MPGAN_ERR_SHAPE, "c1_tail_bwd_reduce: null pointer");
  MPGAN_REQUIRE(n > 0 && h > 0 && w > 0, MPGAN_ERR_SHAPE, "c1_tail_bwd_reduce: empty tensor");
  const bool vec = w % 8 == 0 && ((uintptr_t)dy_bf16 & 15) == 0;
  const int64_t runs = (int64_t)n * h * ((w + (vec ? 7 : 3)) / (vec ? 8 : 4));
  int64_t blocks = ceil_div(runs, (int64_t)kThreads * 2);
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (vec)
    launch_k(c1_tail_bwd_reduce_kernel<8>, (int)blocks, kThreads, 0, (cudaStream_t)stream, (const bf16*)dy_bf16,
             (const bf16*)c_bf16, (int)n, (int)h, (int)w, mean, invstd, scale, shift, alpha, (const bf16*)w9_bf16,
             (bf16*)dh_bf16, sums3);
  else
    launch_k(c1_tail_bwd_reduce_kernel<4>, (int)blocks, kThreads, 0, (cudaStream_t)stream, (const bf16*)dy_bf16,
             (const bf16*)c_bf16, (int)n, (int)h, (int)w, mean, invstd, scale, shift, alpha, (const bf16*)w9_bf16,
             (bf16*)dh_bf16, sums3);
  MPGAN_CHECK_LAUNCH("c1_tail_bwd_reduce_kernel");
  return 0;
}

extern "C" int mpgan_bn_act_bwd_reduce(int dtype, const void* dy, int64_t lddy, const void* x, int64_t ldx,
                                       int64_t pixels, int32_t c, const float* mean, const float* invstd,
                                       const float* scale, const float* shift, int act, const float* alpha,
                                       float leaky_slope, double* sums, void* stream) {
  MPGAN_REQUIRE(pixels > 0 && c > 0 && ldx >= c && lddy >= c && sums, MPGAN_ERR_SHAPE, "bn_act_bwd_reduce: bad shape");
  if (bn_stream_ok(dtype, dy, x, nullptr, lddy, ldx, 0, pixels, c, act))   // contiguous bf16: bulk-copy streaming kernel
    return bn_stream_reduce(dy, x, pixels, c, mean, invstd, scale, shift, act, alpha, leaky_slope, sums,
                            (cudaStream_t)stream);
  const bool vec = (c % 8 == 0) && vec_ok(x, ldx, dtype) && vec_ok(dy, lddy, dtype);
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    if (vec)  // 4 channels per thread: the per-channel constants fit in registers at 3 blocks per SM
      launch_k(bn_act_bwd_reduce_kernel<T, 4>, ew_grid(pixels, c / 4, 6, 16), kThreads, 0, (cudaStream_t)stream, 
          (const T*)dy, lddy, (const T*)x, ldx, pixels, c, mean, invstd, scale, shift, act, alpha, leaky_slope, sums);
    else
      launch_k(bn_act_bwd_reduce_kernel<T, 1>, ew_grid(pixels, c, 8, 32), kThreads, 0, (cudaStream_t)stream, 
          (const T*)dy, lddy, (const T*)x, ldx, pixels, c, mean, invstd, scale, shift, act, alpha, leaky_slope, sums);
    MPGAN_CHECK_LAUNCH("bn_act_bwd_reduce");
    return 0;
  });
}

extern "C" int mpgan_bn_act_bwd_apply(int dtype, const void* dy, int64_t lddy, const void* x, int64_t ldx,
                                      int64_t pixels, int32_t c, const float* mean, const float* invstd,
                                      const float* scale, const float* shift, int act, const float* alpha,
                                      float leaky_slope, const double* sums, float* dgamma, float* dbeta,
                                      float* dalpha, float* dbias, void* dx, int64_t lddx, void* stream) {
  MPGAN_REQUIRE(pixels > 0 && c > 0 && ldx >= c && lddy >= c && lddx >= c && sums, MPGAN_ERR_SHAPE,
                "bn_act_bwd_apply: bad shape");
  MPGAN_REQUIRE((mean == nullptr) == (invstd == nullptr), MPGAN_ERR_SHAPE, "mean/invstd must both be given");
  if (bn_stream_ok(dtype, dy, x, dx, lddy, ldx, lddx, pixels, c, act))   // contiguous bf16: bulk-copy streaming kernel
    return bn_stream_bwd(dy, x, pixels, c, mean, invstd, scale, shift, act, alpha, leaky_slope, sums, dgamma, dbeta,
                         dalpha, dbias, dx, lddx, (cudaStream_t)stream);
  const bool vec = (c % 8 == 0) && vec_ok(x, ldx, dtype) && vec_ok(dy, lddy, dtype) && vec_ok(dx, lddx, dtype);
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    if (vec)
      launch_k(bn_act_bwd_apply_kernel<T, 4>, ew_grid(pixels, c / 4, 6, 16), kThreads, 0, (cudaStream_t)stream, 
          (const T*)dy, lddy, (const T*)x, ldx, pixels, c, mean, invstd, scale, shift, act, alpha, leaky_slope, sums,
          dgamma, dbeta, dalpha, dbias, (T*)dx, lddx);
    else
      launch_k(bn_act_bwd_apply_kernel<T, 1>, ew_grid(pixels, c, 8, 16), kThreads, 0, (cudaStream_t)stream, 
          (const T*)dy, lddy, (const T*)x, ldx, pixels, c, mean, invstd, scale, shift, act, alpha, leaky_slope, sums,
          dgamma, dbeta, dalpha, dbias, (T*)dx, lddx);
    MPGAN_CHECK_LAUNCH("bn_act_bwd_apply");
    return 0;
  });
}
