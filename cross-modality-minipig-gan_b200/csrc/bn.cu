// Batch-norm (training + eval) fused with PReLU / LeakyReLU / Tanh and the residual add: bandwidth kernels.
// Replaces nn.BatchNorm{2,3}d + nn.LeakyReLU / nn.PReLU (+ ResidualUnit's add) and their autograd backward
// (/root/reference/code/GAN/GAN_final.py:167-189; MONAI Convolution "norm"/"act", ResidualUnit.forward).
//
// Layout: channels-last (pixels, C) with pixel stride ld.  Vector path: 8 channels per thread (16 B bf16 /
// 32 B f32) when C, ld are multiples of 8 and the pointers are 16/32-byte aligned; scalar path otherwise (C = 1).
// Per-channel sums are accumulated in fp64 end to end (per thread, shared-memory reduce, one global atomic per
// channel per block): E[x^2]-mean^2 must survive |mean| >> std.
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

// ---------------- statistics: sum, sum of squares ----------------
template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
bn_stats_kernel(const T* __restrict__ x, int64_t ldx, int64_t P, int C, double* __restrict__ stats) {
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

// ---------------- apply: y = act(x*scale + shift) (+ res) ----------------
template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
bn_act_apply_kernel(const T* __restrict__ x, int64_t ldx, int64_t P, int C, const float* __restrict__ scale,
                    const float* __restrict__ shift, int act, const float* __restrict__ alpha, float leaky,
                    const T* __restrict__ res, int64_t ldres, T* __restrict__ y, int64_t ldy) {
  const int cv = C / V;
  const int64_t total = P * cv;
  const float slope = (act == MPGAN_ACT_PRELU || (act == MPGAN_ACT_LEAKY && alpha)) ? *alpha : leaky;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / cv;
    const int c0 = (int)(i - p * cv) * V;
    float v[V], r[V];
    Vec<T, V>::load(x + p * ldx + c0, v);
    if (res) Vec<T, V>::load(res + p * ldres + c0, r);
#pragma unroll
    for (int e = 0; e < V; ++e) {
      float z = v[e];
      if (scale) z = fmaf(z, scale[c0 + e], shift[c0 + e]);
      z = act_fwd(z, act, slope);
      if (res) z += r[e];
      v[e] = z;
    }
    Vec<T, V>::store(y + p * ldy + c0, v);
  }
}

// ---------------- backward pass 1: reductions ----------------
template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
bn_act_bwd_reduce_kernel(const T* __restrict__ dy, int64_t lddy, const T* __restrict__ x, int64_t ldx, int64_t P,
                         int C, const float* __restrict__ mean, const float* __restrict__ invstd,
                         const float* __restrict__ scale, const float* __restrict__ shift, int act,
                         const float* __restrict__ alpha, float leaky, double* __restrict__ sums) {
  __shared__ double s1[kThreads * V], s2[kThreads * V];
  const int cv = C / V;
  const int CL = cv < kThreads ? cv : kThreads;
  const int PL = kThreads / CL;
  const int cl = threadIdx.x % CL, pl = threadIdx.x / CL;
  const int64_t p0 = (int64_t)blockIdx.x * kPixPerBlock;
  const int64_t p1 = min(P, p0 + kPixPerBlock);
  const float slope = (act == MPGAN_ACT_PRELU || (act == MPGAN_ACT_LEAKY && alpha)) ? *alpha : leaky;
  double aslope = 0.0;
  for (int cb = 0; cb < cv; cb += CL) {
    const int vc = cb + cl;
    double a1[V], a2[V];
#pragma unroll
    for (int i = 0; i < V; ++i) a1[i] = a2[i] = 0.0;
    if (pl < PL && vc < cv) {
      const int c0 = vc * V;
      float sc[V], sh[V], mu[V], is[V];
#pragma unroll
      for (int e = 0; e < V; ++e) {
        sc[e] = scale ? scale[c0 + e] : 1.f; sh[e] = shift ? shift[c0 + e] : 0.f;
        mu[e] = mean ? mean[c0 + e] : 0.f; is[e] = invstd ? invstd[c0 + e] : 1.f;
      }
      for (int64_t p = p0 + pl; p < p1; p += PL) {
        float g[V], xv[V];
        Vec<T, V>::load(dy + p * lddy + c0, g);
        Vec<T, V>::load(x + p * ldx + c0, xv);
#pragma unroll
        for (int e = 0; e < V; ++e) {
          float z = fmaf(xv[e], sc[e], sh[e]);
          if (act == MPGAN_ACT_PRELU && z <= 0.f) aslope = fma((double)g[e], (double)z, aslope);
          float gz = g[e] * act_grad(z, act, slope);
          a1[e] += (double)gz;
          a2[e] = fma((double)gz, (double)((xv[e] - mu[e]) * is[e]), a2[e]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) { s1[threadIdx.x * V + i] = a1[i]; s2[threadIdx.x * V + i] = a2[i]; }
    __syncthreads();
    for (int j = threadIdx.x; j < CL * V; j += kThreads) {
      int lane = j / V, e = j % V;
      if (cb + lane < cv) {
        double t1 = 0.0, t2 = 0.0;
        for (int q = 0; q < PL; ++q) { t1 += s1[(q * CL + lane) * V + e]; t2 += s2[(q * CL + lane) * V + e]; }
        int ch = (cb + lane) * V + e;
        atomicAdd(&sums[ch], t1);
        atomicAdd(&sums[C + ch], t2);
      }
    }
    __syncthreads();
  }
  if (act == MPGAN_ACT_PRELU) {  // block-wide fp64 sum of the slope-gradient partials
    double w = warp_sum(aslope);
    __shared__ double dred[kThreads / 32];
    if ((threadIdx.x & 31) == 0) dred[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int i = 0; i < kThreads / 32; ++i) tot += dred[i];
      atomicAdd(&sums[2 * C], tot);
    }
  }
}

// ---------------- backward pass 2: dx, and parameter grads ----------------
template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
bn_act_bwd_apply_kernel(const T* __restrict__ dy, int64_t lddy, const T* __restrict__ x, int64_t ldx, int64_t P,
                        int C, const float* __restrict__ mean, const float* __restrict__ invstd,
                        const float* __restrict__ scale, const float* __restrict__ shift, int act,
                        const float* __restrict__ alpha, float leaky, const double* __restrict__ sums,
                        float* dgamma, float* dbeta, float* dalpha, T* __restrict__ dx, int64_t lddx) {
  const int cv = C / V;
  const int64_t total = P * cv;
  const float slope = (act == MPGAN_ACT_PRELU || (act == MPGAN_ACT_LEAKY && alpha)) ? *alpha : leaky;
  const float invP = 1.f / (float)P;
  if (blockIdx.x == 0) {  // parameter gradients (accumulate), once
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dbeta) dbeta[c] += (float)sums[c];
      if (dgamma) dgamma[c] += (float)sums[C + c];
    }
    if (threadIdx.x == 0 && dalpha && act == MPGAN_ACT_PRELU) *dalpha += (float)sums[2 * C];
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / cv;
    const int c0 = (int)(i - p * cv) * V;
    float g[V], xv[V];
    Vec<T, V>::load(dy + p * lddy + c0, g);
    Vec<T, V>::load(x + p * ldx + c0, xv);
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const int c = c0 + e;
      float sc = scale ? scale[c] : 1.f, sh = shift ? shift[c] : 0.f;
      float z = fmaf(xv[e], sc, sh);
      float gz = g[e] * act_grad(z, act, slope);
      if (mean) {
        float xh = (xv[e] - mean[c]) * invstd[c];
        float mg = (float)sums[c] * invP, mgx = (float)sums[C + c] * invP;
        gz = sc * (gz - mg - xh * mgx);
      } else {
        gz *= sc;  // eval-mode / affine-only
      }
      g[e] = gz;
    }
    Vec<T, V>::store(dx + p * lddx + c0, g);
  }
}

static inline bool vec_ok(const void* p, int64_t ld, int dtype) {
  size_t align = dtype == MPGAN_F32 ? 16 : 16;
  return p == nullptr || (((uintptr_t)p % align) == 0 && (ld % 8) == 0);
}

static inline int ew_grid(int64_t total) {
  int64_t b = ceil_div(total, kThreads);
  int64_t cap = (int64_t)num_sms() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace mpgan

using namespace mpgan;

extern "C" int mpgan_bn_stats(int dtype, const void* x, int64_t ldx, int64_t pixels, int32_t c, double* stats,
                              void* stream) {
  MPGAN_REQUIRE(pixels > 0 && c > 0 && ldx >= c, MPGAN_ERR_SHAPE, "bn_stats: bad shape");
  const bool vec = (c % 8 == 0) && vec_ok(x, ldx, dtype);
  const int grid = (int)ceil_div(pixels, kPixPerBlock);
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    if (vec) bn_stats_kernel<T, 8><<<grid, kThreads, 0, (cudaStream_t)stream>>>((const T*)x, ldx, pixels, c, stats);
    else bn_stats_kernel<T, 1><<<grid, kThreads, 0, (cudaStream_t)stream>>>((const T*)x, ldx, pixels, c, stats);
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
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      stats, pixels, c, gamma, beta, eps, momentum, training, running_mean, running_var, num_batches_tracked, mean,
      invstd, scale, shift);
  MPGAN_CHECK_LAUNCH("bn_finalize");
  return 0;
}

extern "C" int mpgan_bn_act_apply(int dtype, const void* x, int64_t ldx, int64_t pixels, int32_t c,
                                  const float* scale, const float* shift, int act, const float* alpha,
                                  float leaky_slope, const void* res, int64_t ldres, void* y, int64_t ldy,
                                  void* stream) {
  MPGAN_REQUIRE(pixels > 0 && c > 0 && ldx >= c && ldy >= c, MPGAN_ERR_SHAPE, "bn_act_apply: bad shape");
  MPGAN_REQUIRE((scale == nullptr) == (shift == nullptr), MPGAN_ERR_SHAPE, "scale/shift must both be given");
  MPGAN_REQUIRE(act != MPGAN_ACT_PRELU || alpha, MPGAN_ERR_SHAPE, "PReLU needs alpha");
  const bool vec = (c % 8 == 0) && vec_ok(x, ldx, dtype) && vec_ok(y, ldy, dtype) && vec_ok(res, res ? ldres : 8, dtype);
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    if (vec)
      bn_act_apply_kernel<T, 8><<<ew_grid(pixels * (c / 8)), kThreads, 0, (cudaStream_t)stream>>>(
          (const T*)x, ldx, pixels, c, scale, shift, act, alpha, leaky_slope, (const T*)res, ldres, (T*)y, ldy);
    else
      bn_act_apply_kernel<T, 1><<<ew_grid(pixels * c), kThreads, 0, (cudaStream_t)stream>>>(
          (const T*)x, ldx, pixels, c, scale, shift, act, alpha, leaky_slope, (const T*)res, ldres, (T*)y, ldy);
    MPGAN_CHECK_LAUNCH("bn_act_apply");
    return 0;
  });
}

extern "C" int mpgan_bn_act_bwd_reduce(int dtype, const void* dy, int64_t lddy, const void* x, int64_t ldx,
                                       int64_t pixels, int32_t c, const float* mean, const float* invstd,
                                       const float* scale, const float* shift, int act, const float* alpha,
                                       float leaky_slope, double* sums, void* stream) {
  MPGAN_REQUIRE(pixels > 0 && c > 0 && ldx >= c && lddy >= c && sums, MPGAN_ERR_SHAPE, "bn_act_bwd_reduce: bad shape");
  const bool vec = (c % 8 == 0) && vec_ok(x, ldx, dtype) && vec_ok(dy, lddy, dtype);
  const int grid = (int)ceil_div(pixels, kPixPerBlock);
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    if (vec)
      bn_act_bwd_reduce_kernel<T, 8><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
          (const T*)dy, lddy, (const T*)x, ldx, pixels, c, mean, invstd, scale, shift, act, alpha, leaky_slope, sums);
    else
      bn_act_bwd_reduce_kernel<T, 1><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
          (const T*)dy, lddy, (const T*)x, ldx, pixels, c, mean, invstd, scale, shift, act, alpha, leaky_slope, sums);
    MPGAN_CHECK_LAUNCH("bn_act_bwd_reduce");
    return 0;
  });
}

extern "C" int mpgan_bn_act_bwd_apply(int dtype, const void* dy, int64_t lddy, const void* x, int64_t ldx,
                                      int64_t pixels, int32_t c, const float* mean, const float* invstd,
                                      const float* scale, const float* shift, int act, const float* alpha,
                                      float leaky_slope, const double* sums, float* dgamma, float* dbeta,
                                      float* dalpha, void* dx, int64_t lddx, void* stream) {
  MPGAN_REQUIRE(pixels > 0 && c > 0 && ldx >= c && lddy >= c && lddx >= c && sums, MPGAN_ERR_SHAPE,
                "bn_act_bwd_apply: bad shape");
  MPGAN_REQUIRE((mean == nullptr) == (invstd == nullptr), MPGAN_ERR_SHAPE, "mean/invstd must both be given");
  const bool vec = (c % 8 == 0) && vec_ok(x, ldx, dtype) && vec_ok(dy, lddy, dtype) && vec_ok(dx, lddx, dtype);
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    if (vec)
      bn_act_bwd_apply_kernel<T, 8><<<ew_grid(pixels * (c / 8)), kThreads, 0, (cudaStream_t)stream>>>(
          (const T*)dy, lddy, (const T*)x, ldx, pixels, c, mean, invstd, scale, shift, act, alpha, leaky_slope, sums,
          dgamma, dbeta, dalpha, (T*)dx, lddx);
    else
      bn_act_bwd_apply_kernel<T, 1><<<ew_grid(pixels * c), kThreads, 0, (cudaStream_t)stream>>>(
          (const T*)dy, lddy, (const T*)x, ldx, pixels, c, mean, invstd, scale, shift, act, alpha, leaky_slope, sums,
          dgamma, dbeta, dalpha, (T*)dx, lddx);
    MPGAN_CHECK_LAUNCH("bn_act_bwd_apply");
    return 0;
  });
}
