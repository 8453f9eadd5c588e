// Element-wise / layout helpers: residual add + concat copy + casts, tanh, per-channel column sums, weight
// transposes and the NCHW<->NHWC flatten permutation used by the Linear layer after nn.Flatten
// (/root/reference/code/GAN/GAN_final.py:117,199-203; MONAI SkipConnection's torch.cat, ResidualUnit's add).
#include <stdarg.h>

#include "common.cuh"

namespace mpgan {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

template <typename TI, typename TO>
__global__ void add_copy_kernel(const TI* __restrict__ a, int64_t lda, const TI* __restrict__ b, int64_t ldb,
                                TO* __restrict__ y, int64_t ldy, int64_t P, int C) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  const int64_t total = P * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i / C;
    int c = (int)(i - p * C);
    float v = to_f(a[p * lda + c]);
    if (b) v += to_f(b[p * ldb + c]);
    y[p * ldy + c] = from_f<TO>(v);
  }
}

// 8 channels per thread (16-byte bf16 / 2 x 16-byte fp32 accesses); C, lda, ldb, ldy multiples of 8, 16-byte bases
__device__ __forceinline__ void ld8(const float* p, float* o) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
__device__ __forceinline__ void ld8(const bf16* p, float* o) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void st8(float* p, const float* o) {
  *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(o[4], o[5], o[6], o[7]);
}
__device__ __forceinline__ void st8(bf16* p, const float* o) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
add_copy_vec_kernel(const TI* __restrict__ a, int64_t lda, const TI* __restrict__ b, int64_t ldb, TO* __restrict__ y,
                    int64_t ldy, int64_t P, int cv) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  const int64_t total = P * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / cv;
    const int c = (int)(i - p * cv) * 8;
    float v[8], w[8];
    ld8(a + p * lda + c, v);
    if (b) {
      ld8(b + p * ldb + c, w);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += w[e];
    }
    st8(y + p * ldy + c, v);
  }
}

template <typename TI, typename TO>
__global__ void tanh_fwd_kernel(const TI* __restrict__ x, TO* __restrict__ y, int64_t n) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = from_f<TO>(tanhf(to_f(x[i])));
}
template <typename T>
__global__ void tanh_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx, int64_t n) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float t = to_f(y[i]);
    dx[i] = from_f<T>(to_f(dy[i]) * (1.f - t * t));
  }
}

// out[c] += sum_p x[p,c]; block = 256 threads over (pixel lanes x channel lanes)
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, int64_t ldx, int64_t P, int C, float* __restrict__ out, int pix_per_block) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  __shared__ float sm[256];
  const int CL = C < 256 ? C : 256, PL = 256 / CL;
  const int cl = threadIdx.x % CL, pl = threadIdx.x / CL;
  const int64_t p0 = (int64_t)blockIdx.x * pix_per_block, p1 = min(P, p0 + pix_per_block);
  for (int cb = 0; cb < C; cb += CL) {
    float acc = 0.f;
    if (pl < PL && cb + cl < C)
      for (int64_t p = p0 + pl; p < p1; p += PL) acc += to_f(x[p * ldx + cb + cl]);
    sm[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < CL && cb + threadIdx.x < C) {
      float t = 0.f;
      for (int q = 0; q < PL; ++q) t += sm[q * CL + threadIdx.x];
      atomicAdd(&out[cb + threadIdx.x], t);
    }
    __syncthreads();
  }
}

template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ s, TO* __restrict__ d, int64_t n) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    d[i] = from_f<TO>(to_f(s[i]));
}

// [cy][taps][cx] -> [cx][taps][cy]
template <typename TI, typename TO>
__global__ void weight_transpose_kernel(const TI* __restrict__ s, TO* __restrict__ d, int cy, int taps, int cx) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  const int64_t total = (int64_t)cy * taps * cx;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int y = (int)(i % cy);  // destination-major enumeration: i = (x*taps + t)*cy + y
    int64_t r = i / cy;
    int t = (int)(r % taps);
    int xx = (int)(r / taps);
    d[i] = from_f<TO>(to_f(s[((int64_t)y * taps + t) * cx + xx]));
  }
}

// All transposed bf16 weight copies of a network in ONE launch (after every Adam step the generator refreshes 83 of
// them; as separate 2 us launches they were 0.12 ms of the step).  table: n entries of 5 int64 {src, dst, cy, taps, cx}.
__global__ void weight_transpose_batch_kernel(const long long* __restrict__ table) {
  pdl_wait();
  pdl_launch();
  const long long* e = table + 5 * (long long)blockIdx.y;
  const bf16* s = reinterpret_cast<const bf16*>(e[0]);
  bf16* d = reinterpret_cast<bf16*>(e[1]);
  const int cy = (int)e[2], taps = (int)e[3], cx = (int)e[4];
  const int total = cy * taps * cx;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int y = i % cy;
    const int r = i / cy;
    const int t = r % taps, xx = r / taps;
    d[i] = s[(y * taps + t) * cx + xx];
  }
}

// to_cl: dst[r][s][c] = src[r][c][s];  else dst[r][c][s] (+)= src[r][s][c]
template <typename TI, typename TO>
__global__ void permute_flatten_kernel(const TI* __restrict__ src, TO* __restrict__ dst, int rows, int C,
                                       int64_t S, int to_cl, int accumulate) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  const int64_t total = (int64_t)rows * C * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / (C * S);
    int64_t rem = i - r * C * S;
    if (to_cl) {  // i enumerates dst [r][s][c]
      int64_t s = rem / C;
      int c = (int)(rem - s * C);
      dst[i] = from_f<TO>(to_f(src[(r * C + c) * S + s]));
    } else {  // i enumerates src [r][s][c]; write dst [r][c][s]
      int64_t s = rem / C;
      int c = (int)(rem - s * C);
      int64_t o = (r * C + c) * S + s;
      float v = to_f(src[i]);
      if (accumulate) v += to_f(dst[o]);
      dst[o] = from_f<TO>(v);
    }
  }
}

static inline int ew_grid(int64_t total) {
  int64_t b = ceil_div(total, 256);
  int64_t cap = (int64_t)num_sms() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

#define DISPATCH2(dti, dto, TI, TO, ...)                                                           \
  do {                                                                                             \
    if ((dti) == MPGAN_F32 && (dto) == MPGAN_F32) { typedef float TI; typedef float TO; __VA_ARGS__; }          \
    else if ((dti) == MPGAN_F32 && (dto) == MPGAN_BF16) { typedef float TI; typedef mpgan::bf16 TO; __VA_ARGS__; } \
    else if ((dti) == MPGAN_BF16 && (dto) == MPGAN_F32) { typedef mpgan::bf16 TI; typedef float TO; __VA_ARGS__; } \
    else if ((dti) == MPGAN_BF16 && (dto) == MPGAN_BF16) { typedef mpgan::bf16 TI; typedef mpgan::bf16 TO; __VA_ARGS__; } \
    else { mpgan::set_error("bad dtype pair %d/%d", (int)(dti), (int)(dto)); return MPGAN_ERR_UNSUPPORTED; }  \
  } while (0)

}  // namespace mpgan

using namespace mpgan;

extern "C" int mpgan_version(void) { return MPGAN_VERSION; }
extern "C" const char* mpgan_last_error(void) { return g_err; }
extern "C" int mpgan_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return 0; }
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

extern "C" int mpgan_add_copy(int dtype_in, const void* a, int64_t lda, const void* b, int64_t ldb, int dtype_out,
                              void* y, int64_t ldy, int64_t pixels, int32_t c, void* stream) {
  MPGAN_REQUIRE(pixels > 0 && c > 0 && lda >= c && ldy >= c && (!b || ldb >= c), MPGAN_ERR_SHAPE, "add_copy: bad shape");
  const bool vec = c % 8 == 0 && lda % 8 == 0 && ldy % 8 == 0 && (!b || ldb % 8 == 0) && ((uintptr_t)a % 16) == 0 &&
                   ((uintptr_t)y % 16) == 0 && (!b || ((uintptr_t)b % 16) == 0);
  DISPATCH2(dtype_in, dtype_out, TI, TO, {
    if (vec)
      launch_k(add_copy_vec_kernel<TI, TO>, ew_grid(pixels * (c / 8)), 256, 0, (cudaStream_t)stream, 
          (const TI*)a, lda, (const TI*)b, ldb, (TO*)y, ldy, pixels, c / 8);
    else
      launch_k(add_copy_kernel<TI, TO>, ew_grid(pixels * c), 256, 0, (cudaStream_t)stream, (const TI*)a, lda, (const TI*)b,
                                                                                   ldb, (TO*)y, ldy, pixels, c);
    MPGAN_CHECK_LAUNCH("add_copy");
    return 0;
  });
}

extern "C" int mpgan_tanh_fwd(int dtype_in, const void* x, int dtype_out, void* y, int64_t n, void* stream) {
  MPGAN_REQUIRE(n > 0, MPGAN_ERR_SHAPE, "tanh: empty");
  DISPATCH2(dtype_in, dtype_out, TI, TO, {
    launch_k(tanh_fwd_kernel<TI, TO>, ew_grid(n), 256, 0, (cudaStream_t)stream, (const TI*)x, (TO*)y, n);
    MPGAN_CHECK_LAUNCH("tanh_fwd");
    return 0;
  });
}

extern "C" int mpgan_tanh_bwd(int dtype, const void* dy, const void* y, void* dx, int64_t n, void* stream) {
  MPGAN_REQUIRE(n > 0, MPGAN_ERR_SHAPE, "tanh: empty");
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    launch_k(tanh_bwd_kernel<T>, ew_grid(n), 256, 0, (cudaStream_t)stream, (const T*)dy, (const T*)y, (T*)dx, n);
    MPGAN_CHECK_LAUNCH("tanh_bwd");
    return 0;
  });
}

extern "C" int mpgan_colsum(int dtype, const void* x, int64_t ldx, int64_t pixels, int32_t c, float* out,
                            void* stream) {
  MPGAN_REQUIRE(pixels > 0 && c > 0 && ldx >= c, MPGAN_ERR_SHAPE, "colsum: bad shape");
  // enough blocks to cover the machine twice; each block walks >= 64 pixels
  int64_t ppb64 = ceil_div(pixels, (int64_t)num_sms() * 2);
  const int ppb = (int)(ppb64 < 64 ? 64 : (ppb64 > 2048 ? 2048 : ppb64));
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    launch_k(colsum_kernel<T>, (int)ceil_div(pixels, ppb), 256, 0, (cudaStream_t)stream, (const T*)x, ldx, pixels, c, out, ppb);
    MPGAN_CHECK_LAUNCH("colsum");
    return 0;
  });
}

extern "C" int mpgan_cast(int dtype_src, const void* src, int dtype_dst, void* dst, int64_t n, void* stream) {
  MPGAN_REQUIRE(n > 0, MPGAN_ERR_SHAPE, "cast: empty");
  DISPATCH2(dtype_src, dtype_dst, TI, TO, {
    launch_k(cast_kernel<TI, TO>, ew_grid(n), 256, 0, (cudaStream_t)stream, (const TI*)src, (TO*)dst, n);
    MPGAN_CHECK_LAUNCH("cast");
    return 0;
  });
}

extern "C" int mpgan_weight_transpose(int dtype_src, const void* src, int dtype_dst, void* dst, int32_t cy,
                                      int32_t taps, int32_t cx, void* stream) {
  MPGAN_REQUIRE(cy > 0 && taps > 0 && cx > 0, MPGAN_ERR_SHAPE, "weight_transpose: bad shape");
  DISPATCH2(dtype_src, dtype_dst, TI, TO, {
    launch_k(weight_transpose_kernel<TI, TO>, ew_grid((int64_t)cy * taps * cx), 256, 0, (cudaStream_t)stream, 
        (const TI*)src, (TO*)dst, cy, taps, cx);
    MPGAN_CHECK_LAUNCH("weight_transpose");
    return 0;
  });
}

extern "C" int mpgan_weight_transpose_batch(const int64_t* table_dev, int32_t n, int32_t max_elems, void* stream) {
  MPGAN_REQUIRE(table_dev && n > 0 && max_elems > 0, MPGAN_ERR_SHAPE, "weight_transpose_batch: bad arguments");
  MPGAN_REQUIRE(n <= 65535, MPGAN_ERR_UNSUPPORTED, "weight_transpose_batch: more than 65535 tensors");
  int bx = (max_elems + 256 * 8 - 1) / (256 * 8);
  if (bx > 128) bx = 128;
  launch_k(weight_transpose_batch_kernel, dim3(bx, n), 256, 0, (cudaStream_t)stream, (const long long*)table_dev);
  MPGAN_CHECK_LAUNCH("weight_transpose_batch");
  return 0;
}

extern "C" int mpgan_permute_flatten(int dtype_src, const void* src, int dtype_dst, void* dst, int32_t rows,
                                     int32_t c, int64_t spatial, int to_cl, int accumulate, void* stream) {
  MPGAN_REQUIRE(rows > 0 && c > 0 && spatial > 0, MPGAN_ERR_SHAPE, "permute_flatten: bad shape");
  DISPATCH2(dtype_src, dtype_dst, TI, TO, {
    launch_k(permute_flatten_kernel<TI, TO>, ew_grid((int64_t)rows * c * spatial), 256, 0, (cudaStream_t)stream, 
        (const TI*)src, (TO*)dst, rows, c, spatial, to_cl, accumulate);
    MPGAN_CHECK_LAUNCH("permute_flatten");
    return 0;
  });
}
