// HBM-streaming variants of the three hot BatchNorm kernels for LARGE contiguous bf16 tensors (the discriminator's
// 64..256-channel activations: 60-520 MB each, far beyond L2).  The register-staged kernels in bn.cu stall at
// 45-60 % of the HBM rate: the bytes a thread can keep in flight are bounded by its registers.  Here one producer
// thread streams 16 KB chunks with 1-D bulk async copies (cp.async.bulk -> mbarrier complete_tx) into a shared-
// memory ring (up to 6 stages x 2 tensors = 192 KB in flight per SM), 16 consumer warps read the ring with
// conflict-free 16-byte shared loads, do the arithmetic with packed fp32 (FFMA2) instructions and write results
// straight to global memory with coalesced 16-byte stores.  Same math and same reduction tree (fp32 per thread,
// fp64 above the block level) as bn.cu; mpgan_bn_* dispatch here when the tensor qualifies.
// Replaces nn.BatchNorm3d + nn.LeakyReLU and their backward (/root/reference/code/GAN/GAN_final.py:170-189).
#include "common.cuh"
#include "tc_common.cuh"

namespace mpgan {
namespace bns {

using tc::mbar_arrive;
using tc::mbar_expect_tx;
using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_u32;

constexpr int kConsumers = 512;                 // 16 consumer warps
constexpr int kThreadsS = kConsumers + 32;      // + 1 producer warp
constexpr int kChunkBytes = 16384;              // per tensor per stage
constexpr int kMaxStages = 6;

__device__ __forceinline__ void bulk_load(void* smem, const void* gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                   "r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

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
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ uint32_t pack2(float2 v) {
  __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
  return *reinterpret_cast<uint32_t*>(&h);
}
// d act / dz for LeakyReLU / PReLU (slope) or identity (slope == 1)
__device__ __forceinline__ float2 agrad2(float2 z, float slope) {
  return make_float2(z.x > 0.f ? 1.f : slope, z.y > 0.f ? 1.f : slope);
}

struct StreamParams {
  const bf16* in0;     // dy (backward) or x (forward)
  const bf16* in1;     // x (backward); unused forward
  bf16* out;           // dx / y; nullptr for the reduction
  int64_t total;       // elements = pixels * C
  int C;
  int stages;
  int nin;             // input tensors per stage (1 or 2)
  // per-channel constants (fp32, length C each)
  const float* scale; const float* shift; const float* mean; const float* invstd;
  float slope;         // LeakyReLU slope; 1 = no activation
  const float* alpha;  // device slope (PReLU, or LeakyReLU with a learnable slope); overrides `slope`
  int prelu;           // accumulate / apply the PReLU slope gradient
  int has_res;         // forward: in1 is a residual added after the activation
  int64_t ldo;         // output pixel stride in elements (>= C)
  int log2cv;
  float* dalpha;
  int64_t P;
  // reduce mode
  double* sums;        // [2C + 1]
  // backward apply
  const double* sums_in;
  float* dgamma; float* dbeta; float* dbias;
  // forward apply (training): statistics -> coefficients
  const double* stats; const float* gamma; const float* beta; float eps, momentum;
  float* running_mean; float* running_var; long long* nbt;
  float* mean_out; float* invstd_out; float* scale_out; float* shift_out;
};

enum { MODE_FWD = 0, MODE_REDUCE = 1, MODE_BWD = 2 };

template <int MODE>
__global__ void __launch_bounds__(kThreadsS, 1)
bn_stream_kernel(const StreamParams p) {
  pdl_wait();
  pdl_launch();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  uint8_t* ring = smem;                                             // [stages][nin][kChunkBytes]
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + p.stages * p.nin * kChunkBytes);
  uint64_t* empty = full + kMaxStages;
  float* s_coef = reinterpret_cast<float*>(empty + kMaxStages);     // [4][C]
  float* s_red = s_coef + 4 * p.C;                                  // [16 warps][...] reduction scratch
  float* s_slope = s_red + kConsumers * 16;                         // [16 warps]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.C, cv = C >> 3;
  const float slope = p.alpha ? *p.alpha : p.slope;
  const int64_t total_bytes = p.total * 2;
  const int64_t nchunks = (total_bytes + kChunkBytes - 1) / kChunkBytes;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kConsumers / 32); }
    tc::fence_barrier_init();
  }
  // per-channel coefficients -> shared memory: [0] a, [1] b, [2] c, [3] d (meaning depends on MODE)
  if (MODE == MODE_FWD) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float sc, sh;
      if (p.stats) {
        const double m = p.stats[c] / (double)p.P;
        double var = p.stats[C + c] / (double)p.P - m * m;
        if (var < 0.0) var = 0.0;
        const float mean = (float)m;
        const float invstd = (float)(1.0 / sqrt(var + (double)p.eps));
        const float g = p.gamma ? p.gamma[c] : 1.f, b = p.beta ? p.beta[c] : 0.f;
        sc = g * invstd;
        sh = b - mean * sc;
        if (blockIdx.x == 0) {
          if (p.running_mean) p.running_mean[c] = (1.f - p.momentum) * p.running_mean[c] + p.momentum * mean;
          if (p.running_var) {
            double unb = p.P > 1 ? var * ((double)p.P / (double)(p.P - 1)) : var;
            p.running_var[c] = (1.f - p.momentum) * p.running_var[c] + p.momentum * (float)unb;
          }
          if (p.mean_out) p.mean_out[c] = mean;
          if (p.invstd_out) p.invstd_out[c] = invstd;
          p.scale_out[c] = sc;
          p.shift_out[c] = sh;
        }
      } else {
        sc = p.scale ? p.scale[c] : 1.f;
        sh = p.shift ? p.shift[c] : 0.f;
      }
      s_coef[c] = sc; s_coef[C + c] = sh;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.stats && p.nbt) *p.nbt += 1;
  } else if (MODE == MODE_REDUCE) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float sc = p.scale ? p.scale[c] : 1.f, sh = p.shift ? p.shift[c] : 0.f;
      const float mu = p.mean ? p.mean[c] : 0.f, is = p.invstd ? p.invstd[c] : 1.f;
      s_coef[c] = sc; s_coef[C + c] = sh; s_coef[2 * C + c] = is; s_coef[3 * C + c] = -mu * is;   // xhat = x*is - mu*is
    }
  } else {
    const bool train = p.mean != nullptr;
    const float invP = 1.f / (float)p.P;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float sc = p.scale ? p.scale[c] : 1.f, sh = p.shift ? p.shift[c] : 0.f;
      const float mu = train ? p.mean[c] : 0.f, is = train ? p.invstd[c] : 1.f;
      const float mg = train ? (float)p.sums_in[c] * invP : 0.f;
      const float mgx = train ? (float)p.sums_in[C + c] * invP : 0.f;
      const float k2 = -sc * is * mgx;
      s_coef[c] = sc; s_coef[C + c] = sh; s_coef[2 * C + c] = k2; s_coef[3 * C + c] = -sc * mg - k2 * mu;  // k2*x + k3'
      if (blockIdx.x == 0) {
        if (p.dbeta) p.dbeta[c] += (float)p.sums_in[c];
        if (p.dgamma) p.dgamma[c] += (float)p.sums_in[C + c];
      }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.dalpha && p.prelu) *p.dalpha += (float)p.sums_in[2 * C];
  }
  __syncthreads();

  if (warp == kConsumers / 32) {
    if (lane == 0) {  // ================= producer =================
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const int64_t off = ch * kChunkBytes;
        const uint32_t bytes = (uint32_t)min((int64_t)kChunkBytes, total_bytes - off);
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], bytes * p.nin);
        uint8_t* dst = ring + (size_t)stage * p.nin * kChunkBytes;
        bulk_load(dst, reinterpret_cast<const uint8_t*>(p.in0) + off, bytes, &full[stage]);
        if (p.nin == 2) bulk_load(dst + kChunkBytes, reinterpret_cast<const uint8_t*>(p.in1) + off, bytes, &full[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {  // ================= consumers =================
    // item i of a chunk = 16 bytes = 8 channels.  cv = C/8 is a power of two <= 64 (checked on the host), so it
    // divides both the 1024 items of a chunk and the 512 consumer threads: item (chunk, t + r*512) always belongs to
    // channel group t % cv -- the per-channel constants of a thread live in registers for the whole kernel.
    const int t = threadIdx.x;
    const int cg = t % cv;
    float2 ka[4], kb[4], kc[4], kd[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = cg * 8 + 2 * j;
      ka[j] = make_float2(s_coef[c], s_coef[c + 1]);
      kb[j] = make_float2(s_coef[C + c], s_coef[C + c + 1]);
      kc[j] = kd[j] = make_float2(0.f, 0.f);
      if (MODE != MODE_FWD) {
        kc[j] = make_float2(s_coef[2 * C + c], s_coef[2 * C + c + 1]);
        kd[j] = make_float2(s_coef[3 * C + c], s_coef[3 * C + c + 1]);
      }
    }
    float2 acc1[4], acc2[4];
    float fs = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc1[j] = acc2[j] = make_float2(0.f, 0.f);
    int stage = 0;
    uint32_t phase = 0;
    constexpr int kItems = kChunkBytes / 16;   // 1024
    for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
      const int64_t off = ch * kChunkBytes;
      const int bytes = (int)min((int64_t)kChunkBytes, total_bytes - off);
      const int nitems = bytes >> 4;
      mbar_wait(&full[stage], phase);
      const uint8_t* src = ring + (size_t)stage * p.nin * kChunkBytes;
#pragma unroll
      for (int r = 0; r < kItems / kConsumers; ++r) {
        const int i = t + r * kConsumers;
        if (i < nitems) {
          const uint4 u0 = *reinterpret_cast<const uint4*>(src + (size_t)i * 16);
          const uint32_t w0[4] = {u0.x, u0.y, u0.z, u0.w};
          uint32_t o[4];
          if (MODE == MODE_FWD) {
            uint32_t w1[4] = {0u, 0u, 0u, 0u};
            if (p.has_res) {
              const uint4 u1 = *reinterpret_cast<const uint4*>(src + kChunkBytes + (size_t)i * 16);
              w1[0] = u1.x; w1[1] = u1.y; w1[2] = u1.z; w1[3] = u1.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 z = ffma2(unpack2(w0[j]), ka[j], kb[j]);
              z.x = z.x > 0.f ? z.x : z.x * slope;
              z.y = z.y > 0.f ? z.y : z.y * slope;
              if (p.has_res) { const float2 r = unpack2(w1[j]); z.x += r.x; z.y += r.y; }
              o[j] = pack2(z);
            }
          } else {
            const uint4 u1 = *reinterpret_cast<const uint4*>(src + kChunkBytes + (size_t)i * 16);
            const uint32_t w1[4] = {u1.x, u1.y, u1.z, u1.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 g = unpack2(w0[j]), x = unpack2(w1[j]);
              const float2 z = ffma2(x, ka[j], kb[j]);
              const float2 f = agrad2(z, slope);
              const float2 gz = make_float2(g.x * f.x, g.y * f.y);
              if (MODE == MODE_REDUCE) {
                if (p.prelu) {
                  fs = fmaf(g.x, fminf(z.x, 0.f), fs);
                  fs = fmaf(g.y, fminf(z.y, 0.f), fs);
                }
                const float2 xh = ffma2(x, kc[j], kd[j]);
                acc1[j].x += gz.x; acc1[j].y += gz.y;
                acc2[j] = ffma2(gz, xh, acc2[j]);
              } else {
                const float2 dx = ffma2(ka[j], gz, ffma2(kc[j], x, kd[j]));
                o[j] = pack2(dx);
                if (p.dbias) {
                  const float2 rr = unpack2(o[j]);
                  acc1[j].x += rr.x; acc1[j].y += rr.y;
                }
              }
            }
          }
          if (MODE != MODE_REDUCE) {
            bf16* dst;
            if (p.ldo == C) {
              dst = reinterpret_cast<bf16*>(reinterpret_cast<uint8_t*>(p.out) + off + (size_t)i * 16);
            } else {  // channel slice of a wider (concat) buffer
              const int64_t pix = (ch * kItems + i) >> p.log2cv;
              dst = p.out + pix * p.ldo + cg * 8;
            }
            *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    // ---- reductions (thread's channel group is fixed: cv divides 512) ----
    if (MODE == MODE_REDUCE || (MODE == MODE_BWD && p.dbias)) {
      constexpr int NV = MODE == MODE_REDUCE ? 16 : 8;
      float a[NV];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a[2 * j] = acc1[j].x; a[2 * j + 1] = acc1[j].y;
        if (MODE == MODE_REDUCE) { a[8 + 2 * j] = acc2[j].x; a[8 + 2 * j + 1] = acc2[j].y; }
      }
      if (cv <= 32) {
        for (int off = 16; off >= cv; off >>= 1) {
#pragma unroll
          for (int i = 0; i < NV; ++i) a[i] += __shfl_xor_sync(0xffffffffu, a[i], off);
        }
        if (lane < cv) {
#pragma unroll
          for (int i = 0; i < NV; ++i) s_red[(warp * cv + lane) * NV + i] = a[i];
        }
      } else {   // cv == 64: the 32 lanes of a warp hold 32 different groups
#pragma unroll
        for (int i = 0; i < NV; ++i) s_red[t * NV + i] = a[i];
      }
    }
    if (MODE == MODE_REDUCE && p.prelu) {
      const float wsum = warp_sum(fs);
      if (lane == 0) s_slope[warp] = wsum;
    }
  }
  __syncthreads();
  if (MODE == MODE_REDUCE && p.prelu && threadIdx.x == 0) {
    double tsl = 0.0;
    for (int w = 0; w < kConsumers / 32; ++w) tsl += (double)s_slope[w];
    atomicAdd(&p.sums[2 * C], tsl);
  }
  if (MODE == MODE_REDUCE || (MODE == MODE_BWD && p.dbias)) {
    constexpr int NV = MODE == MODE_REDUCE ? 16 : 8;
    const int nw = kConsumers / 32;
    for (int j = threadIdx.x; j < cv * NV; j += blockDim.x) {
      const int g = j / NV, i = j - g * NV;
      double tsum = 0.0;
      if (cv <= 32) {
        for (int w = 0; w < nw; ++w) tsum += (double)s_red[(w * cv + g) * NV + i];
      } else {
        for (int k = 0; k < kConsumers / cv; ++k) tsum += (double)s_red[(g + k * cv) * NV + i];
      }
      const int c = g * 8 + (i & 7);
      if (MODE == MODE_REDUCE) atomicAdd(&p.sums[(i >> 3) * C + c], tsum);
      else atomicAdd(&p.dbias[c], (float)tsum);
    }
  }
}

}  // namespace bns

// Whether the streaming kernels cover a call: bf16, inputs contiguous (ld == C), the output possibly a channel slice
// (ld >= C), 8 | C with C/8 a power of two <= 64 (so a thread's channel group never changes), 16-byte aligned, and
// not tiny.
bool bn_stream_ok(int dtype, const void* in0, const void* in1, const void* out, int64_t ld0, int64_t ld1, int64_t ldo,
                  int64_t pixels, int32_t C, int act) {
  if (dtype != MPGAN_BF16) return false;
  if (C % 8 != 0 || C > 512) return false;
  const int cv = C / 8;
  if ((cv & (cv - 1)) != 0) return false;
  if (ld0 != C || (in1 && ld1 != C)) return false;
  if (out && (ldo < C || ldo % 8 != 0)) return false;
  if (((uintptr_t)in0 & 15) || ((uintptr_t)in1 & 15) || ((uintptr_t)out & 15)) return false;
  if (act == MPGAN_ACT_TANH) return false;
  static int64_t min_bytes = -1;
  if (min_bytes < 0) {
    const char* e = getenv("MPGAN_BN_STREAM_MIN_MB");
    min_bytes = (e ? atoll(e) : 2) << 20;
  }
  return pixels * C * 2 >= min_bytes;
}

static size_t stream_smem(int C, int nin, int* stages) {
  const size_t fixed = 128 + 2 * bns::kMaxStages * 8 + (size_t)4 * C * 4 + (size_t)bns::kConsumers * 16 * 4 + 64 + 256;
  int s = (int)((227 * 1024 - fixed) / ((size_t)nin * bns::kChunkBytes));
  if (s > bns::kMaxStages) s = bns::kMaxStages;
  *stages = s;
  return fixed + (size_t)s * nin * bns::kChunkBytes;
}

template <int MODE>
static int launch_stream(bns::StreamParams& p, cudaStream_t s) {
  int stages;
  size_t smem = stream_smem(p.C, p.nin, &stages);
  {
    // generator-sized tensors (a few 16 KB chunks per CTA): two stages are enough, and the small footprint lets a CTA
    // of the concurrent weight-gradient stream share the SM
    const int64_t per_cta = ceil_div(ceil_div(p.total * 2, (int64_t)bns::kChunkBytes), (int64_t)num_sms());
    if (per_cta <= 16 && stages > 2) {
      smem -= (size_t)(stages - 2) * p.nin * bns::kChunkBytes;
      stages = 2;
    }
  }
  p.stages = stages;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(bns::bn_stream_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    MPGAN_REQUIRE(e == cudaSuccess, MPGAN_ERR_CUDA, "cudaFuncSetAttribute(bn_stream): %s", cudaGetErrorString(e));
    attr_done = true;
  }
  const int64_t nchunks = ceil_div(p.total * 2, (int64_t)bns::kChunkBytes);
  const int grid = (int)(nchunks < num_sms() ? nchunks : num_sms());
  launch_k(bns::bn_stream_kernel<MODE>, grid, bns::kThreadsS, smem, s, p);
  MPGAN_CHECK_LAUNCH("bn_stream_kernel");
  return 0;
}

static int ilog2i(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

int bn_stream_fwd(const void* x, int64_t pixels, int32_t C, const float* scale, const float* shift,
                  const double* stats, const float* gamma, const float* beta, float eps, float momentum,
                  float* running_mean, float* running_var, int64_t* nbt, float* mean, float* invstd, float* scale_out,
                  float* shift_out, int act, const float* alpha, float slope, const void* res, void* y, int64_t ldy,
                  cudaStream_t s) {
  bns::StreamParams p;
  memset(&p, 0, sizeof(p));
  p.in0 = (const bf16*)x; p.in1 = (const bf16*)res; p.out = (bf16*)y; p.total = pixels * C; p.C = C;
  p.nin = res ? 2 : 1; p.has_res = res ? 1 : 0; p.P = pixels; p.ldo = ldy; p.log2cv = ilog2i(C / 8);
  p.scale = scale; p.shift = shift; p.stats = stats; p.gamma = gamma; p.beta = beta; p.eps = eps; p.momentum = momentum;
  p.running_mean = running_mean; p.running_var = running_var; p.nbt = (long long*)nbt;
  p.mean_out = mean; p.invstd_out = invstd; p.scale_out = scale_out; p.shift_out = shift_out;
  p.slope = act == MPGAN_ACT_NONE ? 1.f : slope;
  p.alpha = act == MPGAN_ACT_NONE ? nullptr : alpha;
  return launch_stream<bns::MODE_FWD>(p, s);
}

int bn_stream_reduce(const void* dy, const void* x, int64_t pixels, int32_t C, const float* mean, const float* invstd,
                     const float* scale, const float* shift, int act, const float* alpha, float slope, double* sums,
                     cudaStream_t s) {
  bns::StreamParams p;
  memset(&p, 0, sizeof(p));
  p.in0 = (const bf16*)dy; p.in1 = (const bf16*)x; p.total = pixels * C; p.C = C; p.nin = 2; p.P = pixels;
  p.ldo = C; p.log2cv = ilog2i(C / 8);
  p.mean = mean; p.invstd = invstd; p.scale = scale; p.shift = shift; p.sums = sums;
  p.slope = act == MPGAN_ACT_NONE ? 1.f : slope;
  p.alpha = act == MPGAN_ACT_NONE ? nullptr : alpha;
  p.prelu = act == MPGAN_ACT_PRELU ? 1 : 0;
  return launch_stream<bns::MODE_REDUCE>(p, s);
}

int bn_stream_bwd(const void* dy, const void* x, int64_t pixels, int32_t C, const float* mean, const float* invstd,
                  const float* scale, const float* shift, int act, const float* alpha, float slope, const double* sums,
                  float* dgamma, float* dbeta, float* dalpha, float* dbias, void* dx, int64_t lddx, cudaStream_t s) {
  bns::StreamParams p;
  memset(&p, 0, sizeof(p));
  p.in0 = (const bf16*)dy; p.in1 = (const bf16*)x; p.out = (bf16*)dx; p.total = pixels * C; p.C = C; p.nin = 2;
  p.P = pixels; p.ldo = lddx; p.log2cv = ilog2i(C / 8);
  p.mean = mean; p.invstd = invstd; p.scale = scale; p.shift = shift; p.sums_in = sums;
  p.dgamma = dgamma; p.dbeta = dbeta; p.dalpha = dalpha; p.dbias = dbias;
  p.slope = act == MPGAN_ACT_NONE ? 1.f : slope;
  p.alpha = act == MPGAN_ACT_NONE ? nullptr : alpha;
  p.prelu = act == MPGAN_ACT_PRELU ? 1 : 0;
  return launch_stream<bns::MODE_BWD>(p, s);
}

}  // namespace mpgan
