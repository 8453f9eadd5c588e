// tcgen05 forward convolution for ONE input channel (3x3, stride 1 or 2): D's first layer 1->64 (the largest
// activation of the step is its output: 264 MB per pass) and the generator's 1->16 / 1->32 entry layers.
// Included by conv_tc.cu.
//
// These layers are nine multiply-adds per output value -- pure HBM-write streams -- but on CUDA cores the nine scalar
// 2-byte loads and 36 packed FMAs per (pixel, 8 channels) made them instruction-issue bound at ~1.3 TB/s.  Here the
// arithmetic is one tcgen05.mma per 128-pixel tile: four builder warps assemble the im2col tile
// A[128 pixels][16] (9 taps + 7 zeros, K-major, 32-byte rows, SW32) in shared memory, the weights B[N][16] stay
// resident, the accumulator ring lives in TMEM and the usual small-N epilogue (bias, bf16, 16-byte stores,
// BatchNorm statistics in registers) drains it.  Replaces cuDNN behind nn.Conv2d(1, 64, 3) at
// /root/reference/code/GAN/GAN_final.py:167-169 and MONAI's first ResidualUnit convolutions (GAN_final.py:106-114).
#pragma once

namespace mpgan {
namespace tc {

constexpr int kC1Threads = 512;   // warps 0-3 im2col builders, 4 MMA issuer + TMEM owner, 5 TMA producer, 8-15 epilogue
constexpr int kC1XStages = 16;    // ring of raw input windows (one TMA box each)
constexpr int kC1XBytes = 2304;   // >= 33 rows x 64 B (stride 2) / 18 rows x 32 B (stride 1); multiple of 128
constexpr int kC1Stages = 8;
constexpr int kC1StageBytes = 128 * 32;

struct C1mmaParams {
  int nimg, ih, iw, oh, ow, stride, pad;
  int tiles_w, tiles_h, total_tiles;
  int xrow;                 // bytes per row of an input window in shared memory (box width * 2)
  uint32_t xbytes;          // bytes of one window (TMA transaction size)
  int xoff;                 // column of the window's first needed pixel (the TMA start is kept 16-byte aligned)
  int wide;                 // output rows 32-byte aligned: 256-bit stores
  const bf16* w;            // [N][9]
  bf16* out;
  long long out_sn, out_sh, out_sw;
  const float* bias;
  double* stats;
};

template <int N>
__global__ void __launch_bounds__(kC1Threads, 1)
c1mma_fprop_kernel(const __grid_constant__ C1mmaParams P, const __grid_constant__ CUtensorMap tmX) {
  constexpr int NACC = 512 / N > 8 ? 8 : 512 / N;
  constexpr int TMEM_COLS = NACC * N < 32 ? 32 : NACC * N;
  constexpr int CH = N >= 32 ? 32 : 16;
  constexpr int NCH = N / CH;                       // 1 (N = 16, 32) or 2 (N = 64)
  constexpr uint32_t kEmptyArrivals = NCH == 1 ? 4u : 8u;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* a_sm = smem;                                         // [kC1Stages][128 rows x 32 B]
  uint8_t* x_sm = smem + kC1Stages * kC1StageBytes;             // [kC1XStages][kC1XBytes] raw input windows
  uint8_t* w_sm = x_sm + kC1XStages * kC1XBytes;                // [N rows x 32 B]  (<= 2 KB)
  uint64_t* a_full = reinterpret_cast<uint64_t*>(w_sm + 2048);
  uint64_t* a_empty = a_full + kC1Stages;
  uint64_t* tfull = a_empty + kC1Stages;
  uint64_t* tempty = tfull + 8;
  uint64_t* x_full = tempty + 8;
  uint64_t* x_empty = x_full + kC1XStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_empty + kC1XStages);
  float* s_stats = reinterpret_cast<float*>(w_sm + 2048 + 1024);  // [8 epilogue warps][2 * N]
  float* s_bias = s_stats + 8 * 2 * N;                            // [N]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kC1Stages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 8; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kEmptyArrivals); }
    for (int i = 0; i < kC1XStages; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    fence_barrier_init();
    prefetch_tmap(&tmX);
  }
  if (warp == 4) tmem_alloc<TMEM_COLS>(tmem_slot);
  for (int i = threadIdx.x; i < 8 * 2 * N; i += blockDim.x) s_stats[i] = 0.f;
  pdl_wait();
  pdl_launch();
  if (threadIdx.x < N) {   // resident weights: row n = 9 taps + 7 zeros, chunks swizzled like the A rows
    const int n = threadIdx.x;
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0u;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const uint32_t b = (uint32_t)__bfloat16_as_ushort(P.w[n * 9 + t]);
      v[t >> 1] |= (t & 1) ? (b << 16) : b;
    }
    const uint32_t sw = (uint32_t)(n >> 2) & 1u;
    *reinterpret_cast<uint4*>(w_sm + n * 32 + ((0u ^ sw) << 4)) = make_uint4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<uint4*>(w_sm + n * 32 + ((1u ^ sw) << 4)) = make_uint4(v[4], v[5], v[6], v[7]);
    s_bias[n] = P.bias ? __ldg(&P.bias[n]) : 0.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = P.tiles_w * P.tiles_h;

  if (warp == 5) {
    if (elect_one()) {  // ================= TMA producer: the raw (16S+2) x (8S+2) input window of every tile =========
      int xs = 0;
      uint32_t xpar = 0;
      TileWalk<3> tw;
      { const int radix[3] = {P.tiles_w, P.tiles_h, 1 << 30}; tw.init((int)blockIdx.x, (int)gridDim.x, radix); }
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, tw.next()) {
        const int twi = tw.d[0], thi = tw.d[1], img = tw.d[2];
        mbar_wait(&x_empty[xs], xpar ^ 1);
        mbar_expect_tx(&x_full[xs], P.xbytes);
        // the innermost coordinate must stay a multiple of 8 elements (16-byte aligned global address; an odd start
        // raises an illegal-instruction fault), so a padded layer loads from 8 columns further left
        tma_load_3d(x_sm + xs * kC1XBytes, &tmX, &x_full[xs], twi * HT_W * P.stride - P.pad - P.xoff,
                    thi * HT_H * P.stride - P.pad, img);   // out-of-bounds (incl. negative) coordinates are zero filled
        if (++xs == kC1XStages) { xs = 0; xpar ^= 1; }
      }
    }
  } else if (warp < 4) {
    // ================= im2col builders: warp w assembles tiles w, w+4, ... (lane = 4 rows of the tile) ==========
    const int S = P.stride;
    int it = warp;
    for (int tile = blockIdx.x + warp * gridDim.x; tile < P.total_tiles; tile += 4 * gridDim.x, it += 4) {
      const int stage = it % kC1Stages;
      const uint32_t par = (uint32_t)(it / kC1Stages) & 1u;
      const int xs = it % kC1XStages;
      const uint32_t xpar = (uint32_t)(it / kC1XStages) & 1u;
      mbar_wait(&x_full[xs], xpar);
      const uint8_t* xw = x_sm + xs * kC1XBytes;
      uint32_t v[4][9];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int m = lane + 32 * k;
        const uint8_t* p0 = xw + ((m >> 3) * S) * P.xrow + ((m & 7) * S + P.xoff) * 2;
#pragma unroll
        for (int rh = 0; rh < 3; ++rh)
#pragma unroll
          for (int rw = 0; rw < 3; ++rw)
            v[k][rh * 3 + rw] = (uint32_t)*reinterpret_cast<const unsigned short*>(p0 + rh * P.xrow + rw * 2);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&x_empty[xs]);   // window is in registers
      mbar_wait(&a_empty[stage], par ^ 1);
      uint8_t* st = a_sm + stage * kC1StageBytes;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int m = lane + 32 * k;
        const uint32_t sw = (uint32_t)(m >> 2) & 1u;
        uint8_t* rowp = st + m * 32;
        *reinterpret_cast<uint4*>(rowp + ((0u ^ sw) << 4)) = make_uint4(
            v[k][0] | (v[k][1] << 16), v[k][2] | (v[k][3] << 16), v[k][4] | (v[k][5] << 16), v[k][6] | (v[k][7] << 16));
        *reinterpret_cast<uint4*>(rowp + ((1u ^ sw) << 4)) = make_uint4(v[k][8], 0u, 0u, 0u);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tcgen05.mma reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[stage]);
    }
  } else if (warp == 4) {
    if (elect_one()) {  // ================= MMA issuer: one 128 x N x 16 MMA per tile =================
      constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
      const uint64_t adesc0 = make_smem_desc(smem_u32(a_sm), 16, 256, 6);   // SW32, 8-row groups 256 B apart
      const uint64_t bdesc = make_smem_desc(smem_u32(w_sm), 16, 256, 6);
      int stage = 0, acc = 0;
      uint32_t apar = 0, tpar = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], tpar ^ 1);
        mbar_wait(&a_full[stage], apar);
        tc_fence_after();
        umma_bf16(tmem_base + (uint32_t)(acc * N), adesc0 + (uint64_t)(uint32_t)(stage * (kC1StageBytes >> 4)), bdesc,
                  idesc, 0u);
        umma_commit(&a_empty[stage]);
        umma_commit(&tfull[acc]);
        if (++stage == kC1Stages) { stage = 0; apar ^= 1; }
        if (++acc == NACC) { acc = 0; tpar ^= 1; }
      }
    }
  } else if (warp >= 8) {  // ================= epilogue (8 warps) =================
    const int ew = warp - 8;
    const int q = ew & 3;          // TMEM lane quarter (== warp % 4)
    const int half = ew >> 2;
    const int row = q * 32 + lane;
    const int lh = row >> 3, lw = row & 7;
    float* sl = s_stats + ew * 2 * N;
    const int c0 = NCH == 2 ? half * CH : 0;
    unsigned long long s1[CH / 2], s2[CH / 2];   // packed fp32 pairs
#pragma unroll
    for (int j = 0; j < CH / 2; ++j) { s1[j] = 0ull; s2[j] = 0ull; }
    int it = 0;
    TileWalk<3> tw;
    { const int radix[3] = {P.tiles_w, P.tiles_h, 1 << 30}; tw.init((int)blockIdx.x, (int)gridDim.x, radix); }
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it, tw.next()) {
      if (NCH == 1 && (it & 1) != half) continue;
      const int twi = tw.d[0], thi = tw.d[1], img = tw.d[2];
      const int oh = thi * HT_H + lh, ow = twi * HT_W + lw;
      const bool valid = oh < P.oh && ow < P.ow;
      bf16* orow = P.out + (long long)img * P.out_sn + (long long)oh * P.out_sh + (long long)ow * P.out_sw + c0;
      const int acc = it % NACC;
      const uint32_t par = (uint32_t)(it / NACC) & 1u;
      mbar_wait(&tfull[acc], par);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * N + c0);
      uint32_t r[CH];
      if (CH == 32) tmem_ld_32x32(t_addr, r);
      else tmem_ld_32x16(t_addr, r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      epi_chunk_store<CH>(r, s_bias + c0, orow, valid, P.stats != nullptr, s1, s2, nullptr, P.wide != 0);
    }
    if (P.stats) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = j < CH ? ((j & 1) ? unpack_f32x2(s1[j / 2]).y : unpack_f32x2(s1[j / 2]).x) : 0.f;
      const float t1 = warp_transpose_reduce32(v, lane);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = j < CH ? ((j & 1) ? unpack_f32x2(s2[j / 2]).y : unpack_f32x2(s2[j / 2]).x) : 0.f;
      const float t2 = warp_transpose_reduce32(v, lane);
      if (lane < CH) { sl[c0 + lane] = t1; sl[N + c0 + lane] = t2; }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (P.stats) {
    for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) {
      double s = 0.0;
#pragma unroll
      for (int e = 0; e < 8; ++e) s += (double)s_stats[e * 2 * N + i];
      if (s != 0.0) atomicAdd(&P.stats[i], s);
    }
  }
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <int N>
static int launch_c1mma(const C1mmaParams& P, const CUtensorMap& mX, cudaStream_t s) {
  const size_t smem = (size_t)kC1Stages * kC1StageBytes + (size_t)kC1XStages * kC1XBytes + 2048 + 1024 +
                      (size_t)8 * 2 * N * 4 + (size_t)N * 4 + 64;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(c1mma_fprop_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    MPGAN_REQUIRE(e == cudaSuccess, MPGAN_ERR_CUDA, "cudaFuncSetAttribute(c1mma): %s", cudaGetErrorString(e));
    attr_done = true;
  }
  const int grid = P.total_tiles < num_sms() ? P.total_tiles : num_sms();
  launch_k(c1mma_fprop_kernel<N>, grid, kC1Threads, smem, s, P, mX);
  MPGAN_CHECK_LAUNCH("c1mma_fprop_kernel");
  return 0;
}

}  // namespace tc

// Forward 3x3 convolution of a one-channel bf16 image through the tensor cores.  Returns 1 when not covered.
int c1mma_fprop(const MpganConvGeom* g, int dtype, const void* x, int64_t ldx, const void* w, const float* bias, void* y,
                int64_t ldy, double* stats, cudaStream_t s) {
  using namespace tc;
  if (!g || dtype != MPGAN_BF16 || g->rank != 2 || g->cx != 1) return 1;
  if (g->k[1] != 3 || g->k[2] != 3 || g->stride[1] != g->stride[2] || (g->stride[1] != 1 && g->stride[1] != 2)) return 1;
  if (g->pad[1] != g->pad[2] || g->pad[1] < 0 || g->pad[1] > 1) return 1;
  if (g->stride[1] != 1) return 1;   // stride 2 (generator entry layers) stays on the direct kernels: no gain measured
  const int N = g->cy;
  if (!(N == 16 || N == 32 || N == 64)) return 1;
  if (ldy % 8 != 0 || ((uintptr_t)y & 15) || ((uintptr_t)x & 15)) return 1;
  if (ldx != 1 || g->xs[2] % 8 != 0) return 1;   // the raw image is read through a (w, h, n) tensor map: 16-byte row pitch
  C1mmaParams P;
  memset(&P, 0, sizeof(P));
  P.nimg = g->n; P.ih = g->xs[1]; P.iw = g->xs[2]; P.oh = g->ys[1]; P.ow = g->ys[2];
  P.stride = g->stride[1]; P.pad = g->pad[1];
  P.tiles_w = (P.ow + HT_W - 1) / HT_W; P.tiles_h = (P.oh + HT_H - 1) / HT_H;
  P.total_tiles = g->n * P.tiles_w * P.tiles_h;
  P.w = (const bf16*)w;
  P.xoff = P.pad ? 8 - P.pad : 0;
  const int need = (HT_W - 1) * P.stride + 3 + P.xoff;
  const int bw = need <= 16 ? 16 : 32, bh = (HT_H - 1) * P.stride + 3;   // window (power-of-two row bytes)
  P.xrow = bw * 2; P.xbytes = (uint32_t)(bw * 2 * bh);
  if ((int)P.xbytes > kC1XBytes) return 1;
  CUtensorMap mX;
  {
    uint64_t dims[3] = {(uint64_t)P.iw, (uint64_t)P.ih, (uint64_t)g->n};
    uint64_t str[2] = {(uint64_t)P.iw, (uint64_t)P.iw * P.ih};
    uint32_t box[3] = {(uint32_t)bw, (uint32_t)bh, 1u};
    int rc = encode_map(&mX, x, 3, dims, str, box, true);
    if (rc) return rc;
  }
  P.out = (bf16*)y;
  P.out_sn = (long long)P.oh * P.ow * ldy; P.out_sh = (long long)P.ow * ldy; P.out_sw = ldy;
  P.bias = bias; P.stats = stats;
  P.wide = (ldy % 16 == 0 && ((uintptr_t)y & 31) == 0) ? 1 : 0;
  switch (N) {
    case 16: return launch_c1mma<16>(P, mX, s);
    case 32: return launch_c1mma<32>(P, mX, s);
    default: return launch_c1mma<64>(P, mX, s);
  }
}

}  // namespace mpgan
