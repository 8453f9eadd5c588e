// Generic CUDA-core implicit-GEMM convolution (gather form): any channel count, rank 2/3, f32 or bf16 storage,
// fp32 accumulation.  It is the fp32-mode path (parity <= 1e-4 against the oracle), the path for the C=1 edge
// layers and for 3-D, and the correctness anchor for the tcgen05 kernels in conv_tc.cu.
//
// Replaces torch's convolution forward / convolution_backward behind nn.Conv{2,3}d and nn.ConvTranspose{2,3}d
// (/root/reference/code/GAN/GAN_final.py:167-189; MONAI Convolution via GAN_final.py:106-114).
#include "common.cuh"

namespace mpgan {

struct ConvP {
  int rank, n;
  int xs[3], ys[3], k[3], st[3], pd[3];
  int cx, cy, taps;
};

static int make_convp(const MpganConvGeom* g, ConvP* p) {
  MPGAN_REQUIRE(g != nullptr, MPGAN_ERR_SHAPE, "null geometry");
  MPGAN_REQUIRE(g->rank == 2 || g->rank == 3, MPGAN_ERR_SHAPE, "rank must be 2 or 3 (got %d)", g->rank);
  p->rank = g->rank;
  p->n = g->n;
  p->cx = g->cx;
  p->cy = g->cy;
  p->taps = 1;
  for (int i = 0; i < 3; ++i) {
    p->xs[i] = g->xs[i]; p->ys[i] = g->ys[i]; p->k[i] = g->k[i]; p->st[i] = g->stride[i]; p->pd[i] = g->pad[i];
    MPGAN_REQUIRE(p->xs[i] > 0 && p->ys[i] > 0 && p->k[i] > 0 && p->st[i] > 0 && p->pd[i] >= 0, MPGAN_ERR_SHAPE,
                  "bad conv geometry at dim %d", i);
    // y must be a legal output extent of the conv (ConvTranspose output_padding makes xs larger than minimal)
    int64_t ymax = ((int64_t)p->xs[i] + 2 * p->pd[i] - p->k[i]) / p->st[i] + 1;
    MPGAN_REQUIRE(p->ys[i] <= ymax, MPGAN_ERR_SHAPE, "ys[%d]=%d exceeds conv output extent %lld", i, p->ys[i],
                  (long long)ymax);
    p->taps *= p->k[i];
  }
  if (g->rank == 2)
    MPGAN_REQUIRE(p->xs[0] == 1 && p->ys[0] == 1 && p->k[0] == 1 && p->st[0] == 1 && p->pd[0] == 0,
                  MPGAN_ERR_SHAPE, "rank-2 geometry must have unit depth");
  MPGAN_REQUIRE(p->n > 0 && p->cx > 0 && p->cy > 0, MPGAN_ERR_SHAPE, "bad batch/channels");
  return 0;
}

constexpr int BK = 16;

// MODE 0: out grid = Y (channels cy), gather X (channels cx):  xpos = ypos*st - pd + r
// MODE 1: out grid = X (channels cx), gather Y (channels cy):  ypos = (xpos + pd - r)/st when divisible
template <typename T, int BM, int BN, int MODE>
__global__ void __launch_bounds__(256)
conv_gather_kernel(ConvP p, const T* __restrict__ in, int64_t ldi, const T* __restrict__ w,
                   const float* __restrict__ bias, T* __restrict__ out, int64_t ldo) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  constexpr int TX = BN / 4, TY = BM / 4;
  static_assert(TX * TY == 256, "tile/thread mismatch");
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ int pn[BM], pd_[BM], ph[BM], pw[BM];

  const int tid = threadIdx.x;
  const int* os = MODE == 0 ? p.ys : p.xs;   // output spatial
  const int* is = MODE == 0 ? p.xs : p.ys;   // gathered spatial
  const int N = MODE == 0 ? p.cy : p.cx;
  const int C = MODE == 0 ? p.cx : p.cy;
  const int64_t P = (int64_t)p.n * os[0] * os[1] * os[2];
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int K = p.taps * C;

  for (int i = tid; i < BM; i += 256) {
    int64_t m = m0 + i;
    if (m < P) {
      int ww = (int)(m % os[2]); int64_t r = m / os[2];
      int hh = (int)(r % os[1]); r /= os[1];
      int dd = (int)(r % os[0]);
      pn[i] = (int)(r / os[0]); pd_[i] = dd; ph[i] = hh; pw[i] = ww;
    } else {
      pn[i] = -1; pd_[i] = 0; ph[i] = 0; pw[i] = 0;
    }
  }
  __syncthreads();

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ty = tid / TX, tx = tid % TX;
  const int a_kk = tid % BK, a_m = tid / BK;  // A loader: fixed k-lane, BM/16 rows

  for (int k0 = 0; k0 < K; k0 += BK) {
    // ---- A tile (gathered activations) ----
    {
      int k = k0 + a_kk;
      bool kvalid = k < K;
      int t = kvalid ? k / C : 0;
      int c = k - t * C;
      int rw = t % p.k[2]; int t2 = t / p.k[2];
      int rh = t2 % p.k[1]; int rd = t2 / p.k[1];
#pragma unroll
      for (int i = 0; i < BM / 16; ++i) {
        int m = a_m + i * 16;
        float v = 0.f;
        int nimg = pn[m];
        if (kvalid && nimg >= 0) {
          int id, ih, iw;
          bool ok = true;
          if (MODE == 0) {
            id = pd_[m] * p.st[0] - p.pd[0] + rd;
            ih = ph[m] * p.st[1] - p.pd[1] + rh;
            iw = pw[m] * p.st[2] - p.pd[2] + rw;
          } else {
            int qd = pd_[m] + p.pd[0] - rd, qh = ph[m] + p.pd[1] - rh, qw = pw[m] + p.pd[2] - rw;
            ok = qd >= 0 && qh >= 0 && qw >= 0 && (qd % p.st[0] == 0) && (qh % p.st[1] == 0) &&
                 (qw % p.st[2] == 0);
            id = qd / p.st[0]; ih = qh / p.st[1]; iw = qw / p.st[2];
          }
          ok = ok && id >= 0 && id < is[0] && ih >= 0 && ih < is[1] && iw >= 0 && iw < is[2];
          if (ok) {
            int64_t pix = (((int64_t)nimg * is[0] + id) * is[1] + ih) * is[2] + iw;
            v = to_f(in[pix * ldi + c]);
          }
        }
        As[a_kk][m] = v;
      }
    }
    // ---- B tile (weights, OTI [cy][taps][cx]) ----
    for (int idx = tid; idx < BK * BN; idx += 256) {
      int kk = idx / BN, nn = idx % BN;
      int k = k0 + kk, n = n0 + nn;
      float v = 0.f;
      if (k < K && n < N) {
        int t = k / C, c = k - t * C;
        int64_t off = MODE == 0 ? ((int64_t)n * p.taps + t) * p.cx + c : ((int64_t)c * p.taps + t) * p.cx + n;
        v = to_f(w[off]);
      }
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= P) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < N) {
        float v = acc[i][j] + (bias ? bias[n] : 0.f);
        out[m * ldo + n] = from_f<T>(v);
      }
    }
  }
}

// dw[cy][t][cx] += sum over Y pixels  y[p][cy] * x[gather(p,t)][cx]
template <typename T>
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(ConvP p, const T* __restrict__ x, int64_t ldx, const T* __restrict__ y, int64_t ldy,
                  float* __restrict__ dw, int64_t pix_per_block) {
  pdl_wait();      // programmatic dependent launch: everything before this overlaps the previous kernel
  pdl_launch();
  constexpr int BM = 64, BN = 64;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ int cn[2][BK], cd[2][BK], ch[2][BK], cw[2][BK];

  const int tid = threadIdx.x;
  const int64_t P = (int64_t)p.n * p.ys[0] * p.ys[1] * p.ys[2];
  const int J = p.taps * p.cx;
  const int j0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int64_t pbeg = (int64_t)blockIdx.z * pix_per_block;
  const int64_t pend = min(P, pbeg + pix_per_block);
  if (pbeg >= pend) return;

  auto decode = [&](int buf, int64_t pbase) {
    if (tid < BK) {
      int64_t m = pbase + tid;
      if (m < pend) {
        int ww = (int)(m % p.ys[2]); int64_t r = m / p.ys[2];
        int hh = (int)(r % p.ys[1]); r /= p.ys[1];
        int dd = (int)(r % p.ys[0]);
        cn[buf][tid] = (int)(r / p.ys[0]); cd[buf][tid] = dd; ch[buf][tid] = hh; cw[buf][tid] = ww;
      } else {
        cn[buf][tid] = -1; cd[buf][tid] = 0; ch[buf][tid] = 0; cw[buf][tid] = 0;
      }
    }
  };

  // this thread's B column (tap, channel) is fixed for the whole block
  const int b_jj = tid % BN, b_pp = tid / BN;  // 4 pixel rows per pass, 4 passes
  const int j = j0 + b_jj;
  const bool jvalid = j < J;
  const int t = jvalid ? j / p.cx : 0;
  const int c = j - t * p.cx;
  const int rw = t % p.k[2]; const int t2 = t / p.k[2];
  const int rh = t2 % p.k[1]; const int rd = t2 / p.k[1];
  const int a_m = tid % BM, a_pp = tid / BM;

  // fp32 FMA over runs of 64 pixels, runs summed in fp64: weight gradients behind a BatchNorm are sums with heavy
  // cancellation, and a plain fp32 running sum over 1e4 pixels costs 5e-4 of relative accuracy
  float acc[4][4];
  double dacc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) { acc[i][jj] = 0.f; dacc[i][jj] = 0.0; }
  const int ty = tid / 16, tx = tid % 16;

  decode(0, pbeg);
  __syncthreads();
  int buf = 0, run = 0;
  for (int64_t pb = pbeg; pb < pend; pb += BK, buf ^= 1) {
    decode(buf ^ 1, pb + BK);
#pragma unroll
    for (int i = 0; i < BK / 4; ++i) {
      int pp = a_pp + i * 4;
      int64_t m = pb + pp;
      float v = 0.f;
      if (m < pend && m0 + a_m < p.cy) v = to_f(y[m * ldy + m0 + a_m]);
      As[pp][a_m] = v;
    }
#pragma unroll
    for (int i = 0; i < BK / 4; ++i) {
      int pp = b_pp + i * 4;
      float v = 0.f;
      int nimg = cn[buf][pp];
      if (jvalid && nimg >= 0) {
        int id = cd[buf][pp] * p.st[0] - p.pd[0] + rd;
        int ih = ch[buf][pp] * p.st[1] - p.pd[1] + rh;
        int iw = cw[buf][pp] * p.st[2] - p.pd[2] + rw;
        if (id >= 0 && id < p.xs[0] && ih >= 0 && ih < p.xs[1] && iw >= 0 && iw < p.xs[2]) {
          int64_t pix = (((int64_t)nimg * p.xs[0] + id) * p.xs[1] + ih) * p.xs[2] + iw;
          v = to_f(x[pix * ldx + c]);
        }
      }
      Bs[pp][b_jj] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) b[jj] = Bs[kk][tx * 4 + jj];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(a[i], b[jj], acc[i][jj]);
    }
    if (++run == 4) {
      run = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) { dacc[i][jj] += (double)acc[i][jj]; acc[i][jj] = 0.f; }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= p.cy) continue;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      int jo = j0 + tx * 4 + jj;
      if (jo < J) atomicAdd(&dw[(int64_t)m * J + jo], (float)(dacc[i][jj] + (double)acc[i][jj]));
    }
  }
}

template <typename T, int MODE>
static int launch_gather(const ConvP& p, const void* in, int64_t ldi, const void* w, const float* bias, void* out,
                         int64_t ldo, cudaStream_t s) {
  const int* os = MODE == 0 ? p.ys : p.xs;
  const int N = MODE == 0 ? p.cy : p.cx;
  const int64_t P = (int64_t)p.n * os[0] * os[1] * os[2];
  if (N <= 16) {
    dim3 grid((unsigned)ceil_div(P, 256), (unsigned)ceil_div(N, 16));
    launch_k(conv_gather_kernel<T, 256, 16, MODE>, grid, 256, 0, s, p, (const T*)in, ldi, (const T*)w, bias, (T*)out, ldo);
  } else {
    dim3 grid((unsigned)ceil_div(P, 64), (unsigned)ceil_div(N, 64));
    launch_k(conv_gather_kernel<T, 64, 64, MODE>, grid, 256, 0, s, p, (const T*)in, ldi, (const T*)w, bias, (T*)out, ldo);
  }
  MPGAN_CHECK_LAUNCH("conv_gather_kernel");
  return 0;
}

}  // namespace mpgan

using namespace mpgan;

extern "C" int mpgan_conv_fprop(const MpganConvGeom* g, int dtype, const void* x, int64_t ldx, const void* w,
                                const float* bias, void* y, int64_t ldy, void* stream) {
  ConvP p;
  int rc = make_convp(g, &p);
  if (rc) return rc;
  MPGAN_REQUIRE(ldx >= p.cx && ldy >= p.cy, MPGAN_ERR_SHAPE, "pixel stride smaller than channel count");
  MPGAN_DISPATCH_DTYPE(dtype, T, return (launch_gather<T, 0>(p, x, ldx, w, bias, y, ldy, (cudaStream_t)stream)));
}

extern "C" int mpgan_conv_bprop(const MpganConvGeom* g, int dtype, const void* y, int64_t ldy, const void* w,
                                const float* bias, void* x, int64_t ldx, void* stream) {
  ConvP p;
  int rc = make_convp(g, &p);
  if (rc) return rc;
  MPGAN_REQUIRE(ldx >= p.cx && ldy >= p.cy, MPGAN_ERR_SHAPE, "pixel stride smaller than channel count");
  MPGAN_DISPATCH_DTYPE(dtype, T, return (launch_gather<T, 1>(p, y, ldy, w, bias, x, ldx, (cudaStream_t)stream)));
}

extern "C" int mpgan_conv_wgrad(const MpganConvGeom* g, int dtype, const void* x, int64_t ldx, const void* y,
                                int64_t ldy, float* dw, void* stream) {
  ConvP p;
  int rc = make_convp(g, &p);
  if (rc) return rc;
  MPGAN_REQUIRE(ldx >= p.cx && ldy >= p.cy, MPGAN_ERR_SHAPE, "pixel stride smaller than channel count");
  const int64_t P = (int64_t)p.n * p.ys[0] * p.ys[1] * p.ys[2];
  const int J = p.taps * p.cx;
  const int gx = (int)ceil_div(J, 64), gy = (int)ceil_div(p.cy, 64);
  int64_t want_z = ceil_div((int64_t)num_sms() * 4, (int64_t)gx * gy);
  int64_t ppb = ceil_div(P, want_z);
  ppb = ceil_div(ppb < 256 ? 256 : ppb, BK) * BK;
  const int gz = (int)ceil_div(P, ppb);
  dim3 grid(gx, gy, gz);
  MPGAN_DISPATCH_DTYPE(dtype, T, {
    launch_k(conv_wgrad_kernel<T>, grid, 256, 0, (cudaStream_t)stream, p, (const T*)x, ldx, (const T*)y, ldy, dw, ppb);
    MPGAN_CHECK_LAUNCH("conv_wgrad_kernel");
    return 0;
  });
}
