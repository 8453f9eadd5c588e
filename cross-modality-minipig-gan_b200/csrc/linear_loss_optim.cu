// Linear-after-Flatten (split-K GEMV), Sigmoid+BCE, L1, fused Adam, patch gather / scatter-add.
// Replaces nn.Linear + nn.Sigmoid (/root/reference/code/GAN/GAN_final.py:199-204), F.binary_cross_entropy and
// F.l1_loss (GAN_final.py:244-248), torch.optim.Adam (GAN_final.py:298-308) and the RandSpatialCropSamplesd
// slicing + torch.cat of /root/reference/test_runs/GAN.py:313-337.
#include <algorithm>

#include "common.cuh"

namespace mpgan {

constexpr int kLinK = 4096;  // k-chunk per block

// y[b][j] += sum_k x[b][k] w[j][k]  (+ bias[j] once)
template <typename T>
__global__ void __launch_bounds__(256)
linear_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ y, int64_t K, int J) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  __shared__ float part[8];
  const int b = blockIdx.y;
  const int64_t k0 = (int64_t)blockIdx.x * kLinK;
  const int64_t k1 = min(K, k0 + kLinK);
  const T* xr = x + (int64_t)b * K;
  float xv[kLinK / 256];
#pragma unroll
  for (int i = 0; i < kLinK / 256; ++i) {
    int64_t k = k0 + threadIdx.x + i * 256;
    xv[i] = k < k1 ? to_f(xr[k]) : 0.f;
  }
  for (int j = 0; j < J; ++j) {
    const T* wr = w + (int64_t)j * K;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < kLinK / 256; ++i) {
      int64_t k = k0 + threadIdx.x + i * 256;
      if (k < k1) acc = fmaf(xv[i], to_f(wr[k]), acc);
    }
    float tot = block_sum(acc, part);
    if (threadIdx.x == 0) {
      if (blockIdx.x == 0 && bias) tot += bias[j];
      atomicAdd(&y[(int64_t)b * J + j], tot);
    }
  }
}

// dx[b][k] = sum_j dy[b][j] w[j][k]
template <typename T>
__global__ void linear_dx_kernel(const T* __restrict__ w, const float* __restrict__ dy, T* __restrict__ dx,
                                 int B, int64_t K, int J) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  const int64_t total = (int64_t)B * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = i / K, k = i - b * K;
    float acc = 0.f;
    for (int j = 0; j < J; ++j) acc = fmaf(dy[b * J + j], to_f(w[(int64_t)j * K + k]), acc);
    dx[i] = from_f<T>(acc);
  }
}

// dw[j][k] += sum_b dy[b][j] x[b][k]
template <typename T>
__global__ void linear_dw_kernel(const T* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                                 int B, int64_t K, int J) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  const int64_t total = (int64_t)J * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t j = i / K, k = i - j * K;
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc = fmaf(dy[(int64_t)b * J + j], to_f(x[(int64_t)b * K + k]), acc);
    dw[i] += acc;
  }
}

// Small weight matrix (J * K <= 1024), large batch -- the patch discriminator's Linear(64, 1) over 4 096 patches: the batch
// is split over blocks (one thread per weight element, a run of rows per block, one atomic per element per block) instead
// of 64 threads walking 4 096 rows each (0.8 ms -> a few us).  The bias gradient rides along.
template <typename T>
__global__ void linear_dw_small_kernel(const T* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                                       float* __restrict__ db, int B, int K, int J, int rows) {
  pdl_wait();
  pdl_launch();
  const int i = threadIdx.x;
  const int b0 = blockIdx.x * rows, b1 = min(B, b0 + rows);
  if (i < J * K) {
    const int j = i / K, k = i - j * K;
    float acc = 0.f;
    for (int b = b0; b < b1; ++b) acc = fmaf(dy[(int64_t)b * J + j], to_f(x[(int64_t)b * K + k]), acc);
    if (acc != 0.f) atomicAdd(&dw[i], acc);
  }
  if (db && i < J) {
    float acc = 0.f;
    for (int b = b0; b < b1; ++b) acc += dy[(int64_t)b * J + i];
    if (acc != 0.f) atomicAdd(&db[i], acc);
  }
}

// 8-wide bf16 variants (K % 8 == 0, 16-byte aligned rows): one 16-byte access per operand, the batch / output index in
// blockIdx.y (no 64-bit divisions).  The scalar kernels above cover fp32 and ragged shapes.
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const uint32_t v[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(v[i] << 16); f[2 * i + 1] = __uint_as_float(v[i] & 0xffff0000u); }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint32_t v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    v[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(v[0], v[1], v[2], v[3]);
}

__global__ void __launch_bounds__(256) linear_dx_vec_kernel(const bf16* __restrict__ w, const float* __restrict__ dy,
                                                            bf16* __restrict__ dx, int64_t K8, int J) {
  pdl_wait();
  pdl_launch();
  const int b = blockIdx.y;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < K8; i += (int64_t)gridDim.x * blockDim.x) {
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int j = 0; j < J; ++j) {
      float wv[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(w + (int64_t)j * K8 * 8) + i), wv);
      const float g = dy[b * J + j];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(g, wv[e], acc[e]);
    }
    reinterpret_cast<uint4*>(dx + (int64_t)b * K8 * 8)[i] = pack8(acc);
  }
}

__global__ void __launch_bounds__(256) linear_dw_vec_kernel(const bf16* __restrict__ x, const float* __restrict__ dy,
                                                            float* __restrict__ dw, int B, int64_t K8, int J) {
  pdl_wait();
  pdl_launch();
  const int j = blockIdx.y;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < K8; i += (int64_t)gridDim.x * blockDim.x) {
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll 4
    for (int b = 0; b < B; ++b) {
      float xv[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + (int64_t)b * K8 * 8) + i), xv);
      const float g = dy[(int64_t)b * J + j];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(g, xv[e], acc[e]);
    }
    float4* o = reinterpret_cast<float4*>(dw + ((int64_t)j * K8 + i) * 8);
    float4 a = o[0], c = o[1];
    a.x += acc[0]; a.y += acc[1]; a.z += acc[2]; a.w += acc[3];
    c.x += acc[4]; c.y += acc[5]; c.z += acc[6]; c.w += acc[7];
    o[0] = a; o[1] = c;
  }
}

__global__ void linear_db_kernel(const float* __restrict__ dy, float* __restrict__ db, int B, int J) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= J) return;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) acc += dy[(int64_t)b * J + j];
  db[j] += acc;
}

// ---- sigmoid, BCE on probabilities ----
__global__ void sigmoid_fwd_kernel(const float* __restrict__ z, float* __restrict__ p, int n) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 1.f / (1.f + expf(-z[i]));
}
__global__ void sigmoid_bwd_kernel(const float* __restrict__ dp, const float* __restrict__ p, float* __restrict__ dz,
                                   int n) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dz[i] = dp[i] * p[i] * (1.f - p[i]);
}

__global__ void __launch_bounds__(256)
bce_fwd_kernel(const float* __restrict__ prob, const float* __restrict__ target, float weight,
               float* __restrict__ loss, int n) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  __shared__ float part[8];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float p = prob[i], t = target[i];
    float lp = fmaxf(logf(p), -100.f);
    float lq = fmaxf(log1pf(-p), -100.f);
    acc += (t - 1.f) * lq - t * lp;
  }
  float tot = block_sum(acc, part);
  if (threadIdx.x == 0) *loss += weight * (tot / (float)n);
}

__global__ void bce_bwd_kernel(const float* __restrict__ prob, const float* __restrict__ target, float weight,
                               const float* __restrict__ gscale, float* __restrict__ dprob, int n) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float g = (gscale ? *gscale : 1.f) * weight / (float)n;
  float p = prob[i];
  dprob[i] = g * (p - target[i]) / fmaxf(p * (1.f - p), 1e-12f);
}

// ---- L1 ----
template <typename T>
__global__ void __launch_bounds__(256)
l1_fwd_kernel(const T* __restrict__ a, const T* __restrict__ b, int64_t n, float scale, float* __restrict__ loss) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  __shared__ float part[8];
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc += fabsf(to_f(a[i]) - to_f(b[i]));
  float tot = block_sum(acc, part);
  if (threadIdx.x == 0) atomicAdd(loss, tot * scale);
}

template <typename T>
__global__ void l1_bwd_kernel(const T* __restrict__ a, const T* __restrict__ b, int64_t n, float scale,
                              const float* __restrict__ gscale, T* __restrict__ da, int accumulate) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  const float g = (gscale ? *gscale : 1.f) * scale;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float d = to_f(a[i]) - to_f(b[i]);
    float v = d > 0.f ? g : (d < 0.f ? -g : 0.f);
    if (accumulate) v += to_f(da[i]);
    da[i] = from_f<T>(v);
  }
}

// One pass over (a, b): loss += scale * sum |a - b| (optional) and da (+)= g * sign(a - b) (optional), 16 bytes per thread
// per operand.  The feature-matching ("perceptual") loss of test_runs/GAN.py:288-298 runs this once per activation pair,
// accumulating the gradient straight into the discriminator's backward tensors (2.5 GB of activations per side at batch
// 32 x 128 patches: one read of each instead of forward pass + backward pass + add).
template <typename T> struct Vec16;
template <> struct Vec16<float> { static constexpr int N = 4; };
template <> struct Vec16<bf16> { static constexpr int N = 8; };
template <typename T>
__device__ __forceinline__ void load16(const T* p, float* f);
template <> __device__ __forceinline__ void load16<float>(const float* p, float* f) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
template <> __device__ __forceinline__ void load16<bf16>(const bf16* p, float* f) {
  unpack8(*reinterpret_cast<const uint4*>(p), f);
}
template <typename T>
__device__ __forceinline__ void store16(T* p, const float* f);
template <> __device__ __forceinline__ void store16<float>(float* p, const float* f) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
}
template <> __device__ __forceinline__ void store16<bf16>(bf16* p, const float* f) { *reinterpret_cast<uint4*>(p) = pack8(f); }

template <typename T>
__global__ void __launch_bounds__(256)
l1_fused_kernel(const T* __restrict__ a, const T* __restrict__ b, int64_t n, float scale, const float* __restrict__ gscale,
                float* __restrict__ loss, double* __restrict__ loss64, T* __restrict__ da, int accumulate, int vec) {
  pdl_wait();
  pdl_launch();
  __shared__ float part[8];
  constexpr int V = Vec16<T>::N;
  const float g = (gscale ? *gscale : 1.f) * scale;
  float acc = 0.f;
  const int64_t nv = vec ? n / V : 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    float fa[V], fb[V], fo[V];
    load16<T>(a + i * V, fa);
    load16<T>(b + i * V, fb);
    if (da && accumulate) load16<T>(da + i * V, fo);
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float d = fa[e] - fb[e];
      acc += fabsf(d);
      const float v = d > 0.f ? g : (d < 0.f ? -g : 0.f);
      fo[e] = (da && accumulate) ? fo[e] + v : v;
    }
    if (da) store16<T>(da + i * V, fo);
  }
  for (int64_t i = nv * V + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = to_f(a[i]) - to_f(b[i]);
    acc += fabsf(d);
    if (da) {
      float v = d > 0.f ? g : (d < 0.f ? -g : 0.f);
      if (accumulate) v += to_f(da[i]);
      da[i] = from_f<T>(v);
    }
  }
  if (loss || loss64) {
    const float tot = block_sum(acc, part);
    if (threadIdx.x == 0) {
      if (loss) atomicAdd(loss, tot * scale);
      // fp64 slot: per-block partials of a small term must not be rounded against a running total that already holds the
      // large ones (measured: 1.9e-4 relative on the 16-term feature-matching loss when accumulated in fp32)
      if (loss64) atomicAdd(loss64, (double)tot * (double)scale);
    }
  }
}

// ---- Adam ----
// state[0] = step count (as int bits), state[1] = step_size = lr/(1-beta1^t), state[2] = sqrt(1-beta2^t).
// The step counter lives on the device so a captured CUDA graph advances it on every replay.
__global__ void adam_prep_kernel(float* __restrict__ state, float lr, float beta1, float beta2) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  int step = __float_as_int(state[0]) + 1;
  state[0] = __int_as_float(step);
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  state[1] = (float)((double)lr / bc1);
  state[2] = (float)sqrt(bc2);
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float beta1, float beta2, float eps,
                            const float* __restrict__ state, bf16* __restrict__ shadow) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  const float step_size = state[1], bc2_sqrt = state[2];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    float mi = m[i] + (gi - m[i]) * (1.f - beta1);          // exp_avg.lerp_(grad, 1 - beta1)
    float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    float pi = p[i] - step_size * (mi / denom);
    m[i] = mi; v[i] = vi; p[i] = pi;
    if (shadow) shadow[i] = __float2bfloat16_rn(pi);
  }
}

// ---- patches ----
template <typename T>
__global__ void patch_gather_kernel(const T* __restrict__ vol, int rank, int s0, int s1, int s2, int C,
                                    const int* __restrict__ origins, int num_samples, int roi, T* __restrict__ out,
                                    int64_t total) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  const int r0 = rank == 3 ? roi : 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C); int64_t r = i / C;
    int w = (int)(r % roi); r /= roi;
    int h = (int)(r % roi); r /= roi;
    int d = (int)(r % r0); r /= r0;
    int64_t patch = r;                 // = b*num_samples + s
    int64_t b = patch / num_samples;
    const int* o = origins + patch * rank;
    int od = rank == 3 ? o[0] : 0, oh = o[rank - 2], ow = o[rank - 1];
    int64_t src = ((((b * s0) + od + d) * s1 + oh + h) * s2 + ow + w) * C + c;
    out[i] = vol[src];
  }
}

// deterministic backward: every voxel sums the patches covering it in sample order
template <typename T>
__global__ void patch_scatter_kernel(const T* __restrict__ dpatch, int rank, int s0, int s1, int s2, int C,
                                     const int* __restrict__ origins, int num_samples, int roi,
                                     T* __restrict__ dvol, int64_t total) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  const int r0 = rank == 3 ? roi : 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C); int64_t r = i / C;
    int w = (int)(r % s2); r /= s2;
    int h = (int)(r % s1); r /= s1;
    int d = (int)(r % s0); r /= s0;
    int64_t b = r;
    float acc = to_f(dvol[i]);
    for (int s = 0; s < num_samples; ++s) {
      int64_t patch = b * num_samples + s;
      const int* o = origins + patch * rank;
      int od = rank == 3 ? o[0] : 0, oh = o[rank - 2], ow = o[rank - 1];
      int pd = d - od, ph = h - oh, pw = w - ow;
      if (pd >= 0 && pd < r0 && ph >= 0 && ph < roi && pw >= 0 && pw < roi)
        acc += to_f(dpatch[((((patch * r0) + pd) * roi + ph) * roi + pw) * C + c]);
    }
    dvol[i] = from_f<T>(acc);
  }
}

static inline int ew_grid(int64_t total) {
  int64_t b = ceil_div(total, 256);
  int64_t cap = (int64_t)num_sms() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace mpgan

using namespace mpgan;

extern "C" int mpgan_linear_fwd(int dtype, const void* x, const void* w, const float* bias, float* y,
                                int32_t batch, int64_t k, int32_t j, void* stream) {
  MPGAN_REQUIRE(batch > 0 && k > 0 && j > 0 && batch <= 65535, MPGAN_ERR_SHAPE, "linear_fwd: bad shape");
  dim3 grid((unsigned)ceil_div(k, kLinK), (unsigned)batch);
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    launch_k(linear_fwd_kernel<T>, grid, 256, 0, (cudaStream_t)stream, (const T*)x, (const T*)w, bias, y, k, j);
    MPGAN_CHECK_LAUNCH("linear_fwd");
    return 0;
  });
}

extern "C" int mpgan_linear_bwd(int dtype, const void* x, const void* w, const float* dy, void* dx, float* dw,
                                float* db, int32_t batch, int64_t k, int32_t j, void* stream) {
  MPGAN_REQUIRE(batch > 0 && k > 0 && j > 0 && dy, MPGAN_ERR_SHAPE, "linear_bwd: bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    const bool vec = sizeof(T) == 2 && k % 8 == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)x & 15) == 0 &&
                     ((uintptr_t)dx & 15) == 0 && ((uintptr_t)dw & 15) == 0 && batch <= 65535 && j <= 65535;
    const int vgrid = (int)std::min<int64_t>(ceil_div(k / 8, (int64_t)256), (int64_t)num_sms() * 8);
    if (dx) {
      if (vec) launch_k(linear_dx_vec_kernel, dim3(vgrid, batch), 256, 0, s, (const bf16*)w, dy, (bf16*)dx, k / 8, j);
      else launch_k(linear_dx_kernel<T>, ew_grid((int64_t)batch * k), 256, 0, s, (const T*)w, dy, (T*)dx, batch, k, j);
      MPGAN_CHECK_LAUNCH("linear_dx");
    }
    if (dw && (int64_t)j * k <= 1024 && batch >= 512) {
      const int rows = 64;
      const int threads = (int)(((int64_t)j * k + 31) / 32 * 32);
      launch_k(linear_dw_small_kernel<T>, (batch + rows - 1) / rows, threads, 0, s, (const T*)x, dy, dw, db, (int)batch, (int)k,
               (int)j, rows);
      MPGAN_CHECK_LAUNCH("linear_dw_small");
      return 0;
    }
    if (dw) {
      if (vec) launch_k(linear_dw_vec_kernel, dim3(vgrid, j), 256, 0, s, (const bf16*)x, dy, dw, batch, k / 8, j);
      else launch_k(linear_dw_kernel<T>, ew_grid((int64_t)j * k), 256, 0, s, (const T*)x, dy, dw, batch, k, j);
      MPGAN_CHECK_LAUNCH("linear_dw");
    }
    if (db) {
      launch_k(linear_db_kernel, (j + 63) / 64, 64, 0, s, dy, db, batch, j);
      MPGAN_CHECK_LAUNCH("linear_db");
    }
    return 0;
  });
}

extern "C" int mpgan_sigmoid_fwd(const float* z, float* prob, int32_t n, void* stream) {
  MPGAN_REQUIRE(n > 0 && z && prob, MPGAN_ERR_SHAPE, "sigmoid_fwd: bad arguments");
  launch_k(sigmoid_fwd_kernel, (n + 127) / 128, 128, 0, (cudaStream_t)stream, z, prob, n);
  MPGAN_CHECK_LAUNCH("sigmoid_fwd");
  return 0;
}

extern "C" int mpgan_sigmoid_bwd(const float* dprob, const float* prob, float* dz, int32_t n, void* stream) {
  MPGAN_REQUIRE(n > 0 && dprob && prob && dz, MPGAN_ERR_SHAPE, "sigmoid_bwd: bad arguments");
  launch_k(sigmoid_bwd_kernel, (n + 127) / 128, 128, 0, (cudaStream_t)stream, dprob, prob, dz, n);
  MPGAN_CHECK_LAUNCH("sigmoid_bwd");
  return 0;
}

extern "C" int mpgan_bce_fwd(const float* prob, const float* target, float weight, float* loss, int32_t n,
                             void* stream) {
  MPGAN_REQUIRE(n > 0 && prob && target && loss, MPGAN_ERR_SHAPE, "bce_fwd: bad arguments");
  launch_k(bce_fwd_kernel, 1, 256, 0, (cudaStream_t)stream, prob, target, weight, loss, n);
  MPGAN_CHECK_LAUNCH("bce_fwd");
  return 0;
}

extern "C" int mpgan_bce_bwd(const float* prob, const float* target, float weight, const float* gscale,
                             float* dprob, int32_t n, void* stream) {
  MPGAN_REQUIRE(n > 0 && prob && target && dprob, MPGAN_ERR_SHAPE, "bce_bwd: bad arguments");
  launch_k(bce_bwd_kernel, (n + 127) / 128, 128, 0, (cudaStream_t)stream, prob, target, weight, gscale, dprob, n);
  MPGAN_CHECK_LAUNCH("bce_bwd");
  return 0;
}

extern "C" int mpgan_l1_fwd(int dtype, const void* a, const void* b, int64_t n, float weight, float* loss,
                            void* stream) {
  MPGAN_REQUIRE(n > 0 && a && b && loss, MPGAN_ERR_SHAPE, "l1_fwd: bad arguments");
  int grid = ew_grid(n);
  if (grid > num_sms() * 4) grid = num_sms() * 4;
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    launch_k(l1_fwd_kernel<T>, grid, 256, 0, (cudaStream_t)stream, (const T*)a, (const T*)b, n, weight / (float)n, loss);
    MPGAN_CHECK_LAUNCH("l1_fwd");
    return 0;
  });
}

extern "C" int mpgan_l1_bwd(int dtype, const void* a, const void* b, int64_t n, float weight, const float* gscale,
                            void* da, int accumulate, void* stream) {
  MPGAN_REQUIRE(n > 0 && a && b && da, MPGAN_ERR_SHAPE, "l1_bwd: bad arguments");
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    launch_k(l1_bwd_kernel<T>, ew_grid(n), 256, 0, (cudaStream_t)stream, (const T*)a, (const T*)b, n, weight / (float)n,
                                                                 gscale, (T*)da, accumulate);
    MPGAN_CHECK_LAUNCH("l1_bwd");
    return 0;
  });
}

extern "C" int mpgan_l1_fwd_bwd(int dtype, const void* a, const void* b, int64_t n, float weight, const float* gscale,
                                float* loss, double* loss64, void* da, int accumulate, void* stream) {
  MPGAN_REQUIRE(n > 0 && a && b && (loss || loss64 || da), MPGAN_ERR_SHAPE, "l1_fwd_bwd: bad arguments");
  const int vec = (((uintptr_t)a | (uintptr_t)b | (uintptr_t)da) & 15) == 0 ? 1 : 0;
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    int64_t items = vec ? ceil_div(n, (int64_t)Vec16<T>::N) : n;
    int grid = ew_grid(items);
    if (grid > num_sms() * 8) grid = num_sms() * 8;
    launch_k(l1_fused_kernel<T>, grid, 256, 0, (cudaStream_t)stream, (const T*)a, (const T*)b, n, weight / (float)n, gscale, loss,
             loss64, (T*)da, accumulate, vec);
    MPGAN_CHECK_LAUNCH("l1_fwd_bwd");
    return 0;
  });
}

extern "C" int mpgan_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                               float lr, float beta1, float beta2, float eps, float* state, void* bf16_shadow,
                               void* stream) {
  MPGAN_REQUIRE(n > 0 && state && param && grad && exp_avg && exp_avg_sq, MPGAN_ERR_SHAPE, "adam: bad arguments");
  launch_k(adam_prep_kernel, 1, 1, 0, (cudaStream_t)stream, state, lr, beta1, beta2);
  MPGAN_CHECK_LAUNCH("adam_prep");
  launch_k(adam_kernel, ew_grid(n), 256, 0, (cudaStream_t)stream, param, grad, exp_avg, exp_avg_sq, n, beta1, beta2, eps,
                                                           state, (bf16*)bf16_shadow);
  MPGAN_CHECK_LAUNCH("adam");
  return 0;
}

extern "C" int mpgan_patch_gather(int dtype, const void* vol, int32_t batch, int32_t rank, const int32_t* spatial,
                                  int32_t c, const int32_t* origins, int32_t num_samples, int32_t roi, void* out,
                                  void* stream) {
  MPGAN_REQUIRE((rank == 2 || rank == 3) && batch > 0 && c > 0 && num_samples > 0 && roi > 0 && spatial && origins,
                MPGAN_ERR_SHAPE, "patch_gather: bad arguments");
  int s0 = rank == 3 ? spatial[0] : 1, s1 = spatial[rank - 2], s2 = spatial[rank - 1];
  int64_t total = (int64_t)batch * num_samples * (rank == 3 ? roi : 1) * roi * roi * c;
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    launch_k(patch_gather_kernel<T>, ew_grid(total), 256, 0, (cudaStream_t)stream, (const T*)vol, rank, s0, s1, s2, c,
                                                                           origins, num_samples, roi, (T*)out, total);
    MPGAN_CHECK_LAUNCH("patch_gather");
    return 0;
  });
}

extern "C" int mpgan_patch_scatter_add(int dtype, const void* dpatch, int32_t batch, int32_t rank,
                                       const int32_t* spatial, int32_t c, const int32_t* origins,
                                       int32_t num_samples, int32_t roi, void* dvol, void* stream) {
  MPGAN_REQUIRE((rank == 2 || rank == 3) && batch > 0 && c > 0 && num_samples > 0 && roi > 0 && spatial && origins,
                MPGAN_ERR_SHAPE, "patch_scatter_add: bad arguments");
  int s0 = rank == 3 ? spatial[0] : 1, s1 = spatial[rank - 2], s2 = spatial[rank - 1];
  int64_t total = (int64_t)batch * s0 * s1 * s2 * c;
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    launch_k(patch_scatter_kernel<T>, ew_grid(total), 256, 0, (cudaStream_t)stream, (const T*)dpatch, rank, s0, s1, s2, c,
                                                                            origins, num_samples, roi, (T*)dvol, total);
    MPGAN_CHECK_LAUNCH("patch_scatter_add");
    return 0;
  });
}
