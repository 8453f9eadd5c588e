// "Halo-resident" tcgen05 engines for the stride-2 3x3 (pad 1) layers of the MONAI UNet -- the down-sampling
// convolutions (Conv k3 s2 p1: forward, and the data gradient of a ConvTranspose) and the up-sampling ones
// (ConvTranspose k3 s2 p1 op1: forward, and the data gradient of a stride-2 convolution).  Included by conv_tc.cu.
//
// Why: tapgemm_kernel feeds these layers with one TMA box of 128 rows x 32/64 bytes per filter tap plus one weight box
// per tap.  ncu (profiles/ncu_r2_small_channel.md): the epilogue warps sit on the accumulator-full barrier 38 % of the
// samples and a 16 -> 32 layer needs ~3 700 cycles per 128-pixel tile = ~2.6 cycles per TMA row -- the TMA row rate is the
// bound, at 20-25 % of the HBM rate.  Here, as in conv_halo.cuh, the weights are resident in shared memory and the
// activations of a tile arrive as ONE halo box per plane; the filter taps are shifted UMMA descriptors into that box.
//
//   MODE 0  (X grid -> Y grid = X / 2):  y[oh, ow] = sum_{kh,kw} x[2 oh - 1 + kh, 2 ow - 1 + kw] . w[kh][kw]
//     Two planes per tile, one per column parity wp of x: plane[wp] = x[2 (h0 - 1) .. 2 (h0 - 1) + 33][wp + 2 (w0 - 1 + j)], a
//     4-D box (C, 9, 34, 1) of the parity view (pixel stride 2 ld).  Output row lh of the tile reads plane row 2 lh + kh + 1
//     (SBO = 18 plane rows), column j = lw + (kw == 0 ? 0 : 1) of plane (kw == 1 ? 0 : 1): still nine MMAs of K = C per
//     tile, fed by 2 boxes instead of 9 + 9.
//   MODE 1  (Y grid -> X grid = 2 Y), "pixel shuffle":  x[2a + i, 2b + j] for the four parity classes (i, j) are four
//     groups of Co accumulator columns of ONE tile over the Y grid; they read the 2 x 2 neighbourhood y[a + da, b + db]:
//       (0,0): Y00.W11     (0,1): Y01.W10 + Y00.W12     (1,0): Y10.W01 + Y00.W21     (1,1): Y11.W00 + Y10.W02 + Y01.W20 + Y00.W22
//     With the column groups ordered (0,0), (0,1), (1,1), (1,0) and the weight blocks ordered to match, that is four MMAs
//     per k-step: Y00 x [W11 W12 W22 W21] (N = 4 Co), Y01 x [W10 W20] (2 Co), Y10 x [W02 W01] (2 Co), Y11 x [W00] (Co).
//     One box (KC, 9, 17, 1) per 64-channel plane.
// Epilogue as in the halo kernel: bias (+ PReLU slope for the inference fusion) (+ residual row) -> bf16 -> 32-byte
// stores, BatchNorm statistics of the stored values in per-thread packed registers.
// Replaces cuDNN behind MONAI's strided Convolution / ResidualUnit residual conv / transposed Convolution
// (/root/reference/code/GAN/GAN_final.py:106-114).
#pragma once

namespace mpgan {
namespace tc {

constexpr int S2_TW = 8, S2_TH = 16;                  // tile on the coarse (Y) grid: M = 128 rows
constexpr int S2_HW = S2_TW + 1, S2_HH = S2_TH + 1;   // 9 x 17 halo on the coarse grid
constexpr int kS2Threads = 384;                       // warp 0 TMA, 1 MMA, 2 TMEM owner, 3 idle, 4-11 epilogue
constexpr int kS2MaxBuf = 8, kS2MaxAcc = 8, kS2Aux = 512;

struct S2Maps { CUtensorMap m[2]; };

struct S2Params {
  int nimg, ch, cw;        // coarse grid = tile grid (mode 0: the output grid; mode 1: the input grid)
  int C;                   // reduced channels
  int tiles_w, tiles_h, total_tiles, nbuf;
  const bf16* w;           // mode 0: [N][9][C];  mode 1: [Co][9][C] (transposed shadow)
  bf16* out;
  long long out_sn, out_sh, out_sw;   // strides of the output grid (mode 1: the fine grid)
  const float* bias;
  double* stats;
  const bf16* res;         // optional tensor on the output grid added before rounding
  long long res_sn, res_sh, res_sw;
  const float* slope;      // optional device scalar: PReLU on (acc + bias) before the residual (inference fusion)
  int wide;                // output rows 32-byte aligned: 256-bit stores
};

__device__ __forceinline__ constexpr int s2_pow2(int v) { int p = 32; while (p < v) p <<= 1; return p; }

template <int MODE, int NP, int KC>   // NP: produced channels (mode 1: per parity class); KC: channels per plane
__global__ void __launch_bounds__(kS2Threads, 1)
halo_s2_kernel(const __grid_constant__ S2Params P, const __grid_constant__ S2Maps tm) {
  constexpr int NT = MODE == 0 ? NP : 4 * NP;                     // TMEM columns of one accumulator tile
  constexpr int NACC = 256 / NT > kS2MaxAcc ? kS2MaxAcc : (256 / NT < 2 ? 2 : 256 / NT);
  constexpr int TMEM_COLS = s2_pow2(NACC * NT);
  static_assert(TMEM_COLS <= 512, "accumulator ring exceeds TMEM");
  constexpr int rowb = KC * 2;                                    // bytes per pixel row of a plane: 128 / 64 / 32
  constexpr int cpr = rowb >> 4;
  constexpr uint32_t swz_mask = (uint32_t)(cpr - 1);
  constexpr int PROWS = MODE == 0 ? 2 * S2_HH * S2_HW : S2_HH * S2_HW;   // 306 / 153 pixel rows per plane
  constexpr int a_bytes = (PROWS * rowb + 1023) & ~1023;
  constexpr uint32_t layout = KC == 64 ? 2u : (KC == 32 ? 4u : 6u);       // SW128 / SW64 / SW32
  constexpr int ksteps = KC >> 4;
  // epilogue work split.  mode 0, NP >= 64: both warp groups take every tile, half of the columns each; otherwise the
  // groups take alternate tiles.  Per-thread statistics registers exist for column spans <= 32 only.
  constexpr bool SPLIT = MODE == 0 && NP >= 64;
  constexpr int SPAN = MODE == 0 ? (SPLIT ? NP / 2 : NP) : NP;
  constexpr int CH = SPAN < 32 ? SPAN : 32;
  constexpr int NCHUNK = SPAN / CH;
  constexpr bool STATS_OK = SPAN <= 32;
  static_assert(CH == 16 || CH == 32, "column chunk");
  static_assert(SPAN % CH == 0, "column span");
  constexpr uint32_t kEmptyArrivals = SPLIT ? 8u : 4u;

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int C = P.C;
  const int nplanes = MODE == 0 ? 2 : C / KC;
  const int w_bytes = 9 * C * NP * 2;
  uint8_t* w_sm = smem;
  uint8_t* a_sm = smem + ((w_bytes + 1023) & ~1023);
  uint8_t* aux = a_sm + P.nbuf * a_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(aux);
  uint64_t* a_empty = a_full + kS2MaxBuf;
  uint64_t* tfull = a_empty + kS2MaxBuf;
  uint64_t* tempty = tfull + kS2MaxAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kS2MaxAcc);
  float* s_stats = reinterpret_cast<float*>(aux + kS2Aux);   // [8 epilogue warps][2 * NP]
  float* s_bias = s_stats + 8 * 2 * NP;                      // [NP]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&tm.m[0]);
    if (MODE == 0) prefetch_tmap(&tm.m[1]);
  }
  if (threadIdx.x == 32) {
    for (int i = 0; i < kS2MaxBuf; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kS2MaxAcc; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kEmptyArrivals); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
  for (int i = threadIdx.x; i < 8 * 2 * NP; i += blockDim.x) s_stats[i] = 0.f;
  pdl_wait();
  pdl_launch();
  // resident weights: global [n][tap][C] -> shared [block][n][KC channels], rows swizzled exactly like a TMA box.
  // mode 0: block = tap (one plane holds all C = KC channels);  mode 1: block = kc * 9 + position of the tap in the
  // MMA-group order {W11 W12 W22 W21 | W10 W20 | W02 W01 | W00}.
  {
    const int cpp = C >> 3;                 // 16-byte chunks per (n, tap) row of the global tensor
    const int total = NP * 9 * cpp;
    for (int g = threadIdx.x; g < total; g += blockDim.x) {
      const int chunk = g % cpp;
      const int r = g / cpp;
      const int tap = r % 9, n = r / 9;
      const int kc = chunk / cpr, cc = chunk - kc * cpr;
      int block;
      if (MODE == 0) {
        block = tap;
      } else {
        const int pos = tap == 4 ? 0 : tap == 5 ? 1 : tap == 8 ? 2 : tap == 7 ? 3 : tap == 3 ? 4 : tap == 6 ? 5 : tap == 2 ? 6
                        : tap == 1 ? 7 : 8;
        block = kc * 9 + pos;
      }
      const uint32_t blk = (uint32_t)(block * NP) * rowb;    // multiple of the swizzle period (NP % 8 == 0)
      uint32_t off = (uint32_t)n * rowb + cc * 16;
      off ^= ((off >> 7) & swz_mask) << 4;
      cp_async16(smem_u32(w_sm + blk + off), P.w + (size_t)g * 8);
    }
    for (int i = threadIdx.x; i < NP; i += blockDim.x) s_bias[i] = P.bias ? __ldg(&P.bias[i]) : 0.f;
    cp_async_wait_all();
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {  // ================= TMA producer: one box per (tile, plane) =================
      int pbuf = 0;
      uint32_t ppar = 0;
      TileWalk<3> tw;
      { const int radix[3] = {P.tiles_w, P.tiles_h, 1 << 30}; tw.init((int)blockIdx.x, (int)gridDim.x, radix); }
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, tw.next()) {
        const int w0 = tw.d[0] * S2_TW, h0 = tw.d[1] * S2_TH, img = tw.d[2];
        for (int p = 0; p < nplanes; ++p) {
          mbar_wait(&a_empty[pbuf], ppar ^ 1);
          mbar_expect_tx(&a_full[pbuf], (uint32_t)(PROWS * rowb));
          if (MODE == 0) tma_load_4d(a_sm + pbuf * a_bytes, &tm.m[p], &a_full[pbuf], 0, w0 - 1, 2 * (h0 - 1), img);
          else tma_load_4d(a_sm + pbuf * a_bytes, &tm.m[0], &a_full[pbuf], p * KC, w0, h0, img);
          if (++pbuf == P.nbuf) { pbuf = 0; ppar ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {  // ================= MMA issuer =================
      constexpr uint32_t sbo_a = (MODE == 0 ? 2 * S2_HW : S2_HW) * rowb, sbo_b = 8 * rowb;
      const uint64_t adesc0 = make_smem_desc(smem_u32(a_sm), 16, sbo_a, layout);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(w_sm), 16, sbo_b, layout);
      int acc = 0, buf = 0;
      uint32_t tpar = 0, apar = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], tpar ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NT);
        if constexpr (MODE == 0) {
          constexpr uint32_t idesc = make_idesc_bf16(128, NP, 0, 0);
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            mbar_wait(&a_full[buf], apar);
            tc_fence_after();
            const uint64_t ad = adesc0 + (uint64_t)(uint32_t)(buf * (a_bytes >> 4));
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
              for (int kwi = 0; kwi < (p == 0 ? 1 : 2); ++kwi) {
                const int kw = p == 0 ? 1 : (kwi == 0 ? 0 : 2);
                const int acol = kw == 0 ? 0 : 1;
#pragma unroll
                for (int ks = 0; ks < ksteps; ++ks) {
                  const uint64_t adesc = ad + (uint64_t)((((kh + 1) * S2_HW + acol) * rowb + ks * 32) >> 4);
                  const uint64_t bdesc = bdesc0 + (uint64_t)((((kh * 3 + kw) * NP) * rowb + ks * 32) >> 4);
                  if (p == 0 && kh == 0 && ks == 0) umma_bf16(d_tmem, adesc, bdesc, idesc, 0u);
                  else umma_bf16(d_tmem, adesc, bdesc, idesc, 1u);
                }
              }
            }
            umma_commit(&a_empty[buf]);
            if (++buf == P.nbuf) { buf = 0; apar ^= 1; }
          }
        } else {
          constexpr uint32_t id4 = make_idesc_bf16(128, 4 * NP, 0, 0), id2 = make_idesc_bf16(128, 2 * NP, 0, 0),
                             id1 = make_idesc_bf16(128, NP, 0, 0);
          for (int kc = 0; kc < nplanes; ++kc) {
            mbar_wait(&a_full[buf], apar);
            tc_fence_after();
            const uint64_t ad = adesc0 + (uint64_t)(uint32_t)(buf * (a_bytes >> 4));
            const uint64_t bd = bdesc0 + (uint64_t)(uint32_t)((kc * 9 * NP * rowb) >> 4);
#pragma unroll
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint32_t k16 = (uint32_t)((ks * 32) >> 4);
              umma_bf16(d_tmem, ad + k16, bd + k16, id4, (kc | ks) != 0 ? 1u : 0u);                                         // Y00
              umma_bf16(d_tmem + NP, ad + (uint32_t)(rowb >> 4) + k16, bd + (uint32_t)((4 * NP * rowb) >> 4) + k16, id2, 1u);   // Y01
              umma_bf16(d_tmem + 2 * NP, ad + (uint32_t)((S2_HW * rowb) >> 4) + k16, bd + (uint32_t)((6 * NP * rowb) >> 4) + k16,
                        id2, 1u);                                                                                           // Y10
              umma_bf16(d_tmem + 2 * NP, ad + (uint32_t)(((S2_HW + 1) * rowb) >> 4) + k16,
                        bd + (uint32_t)((8 * NP * rowb) >> 4) + k16, id1, 1u);                                              // Y11
            }
            umma_commit(&a_empty[buf]);
            if (++buf == P.nbuf) { buf = 0; apar ^= 1; }
          }
        }
        umma_commit(&tfull[acc]);
        if (++acc == NACC) { acc = 0; tpar ^= 1; }
      }
    }
  } else if (warp >= 4) {  // ================= epilogue (8 warps) =================
    const int ew = warp - 4;
    const int q = ew & 3;          // TMEM lane quarter (== warp % 4)
    const int half = ew >> 2;
    const int row = q * 32 + lane;
    const int lh = row >> 3, lw = row & 7;
    float* sl = s_stats + ew * 2 * NP;
    const int cbeg = SPLIT ? half * SPAN : 0;
    unsigned long long s1[CH / 2], s2[CH / 2];   // packed fp32 pairs (used when STATS_OK)
#pragma unroll
    for (int j = 0; j < CH / 2; ++j) { s1[j] = 0ull; s2[j] = 0ull; }
    const bool do_stats = STATS_OK && P.stats != nullptr;
    int it = 0;
    TileWalk<3> tw;
    { const int radix[3] = {P.tiles_w, P.tiles_h, 1 << 30}; tw.init((int)blockIdx.x, (int)gridDim.x, radix); }
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it, tw.next()) {
      if (!SPLIT && (it & 1) != half) continue;
      const int ph = tw.d[1] * S2_TH + lh, pw = tw.d[0] * S2_TW + lw, img = tw.d[2];
      const bool valid = ph < P.ch && pw < P.cw;
      const int acc = it % NACC;
      const uint32_t par = (uint32_t)(it / NACC) & 1u;
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NT);
      if constexpr (MODE == 0) {
        bf16* obase = P.out + (long long)img * P.out_sn + (long long)ph * P.out_sh + (long long)pw * P.out_sw;
        const bf16* rbase = P.res ? P.res + (long long)img * P.res_sn + (long long)ph * P.res_sh + (long long)pw * P.res_sw
                                  : nullptr;
        // the residual row of a one-chunk tile is requested BEFORE the wait for the accumulator: its DRAM latency
        // overlaps the MMAs instead of following them (ncu: 31 % of the halo kernel's samples sat on that load)
        uint32_t rq[CH / 2];
        const bool has_res = rbase != nullptr && valid;
        if (NCHUNK == 1 && has_res) load_res_row<CH>(rbase + cbeg, rq);
        mbar_wait(&tfull[acc], par);
        tc_fence_after();
#pragma unroll 1
        for (int k = 0; k < NCHUNK; ++k) {
          const int c0 = cbeg + k * CH;
          uint32_t r[CH];
          if (CH == 32) tmem_ld_32x32(t_addr + c0, r);
          else tmem_ld_32x16(t_addr + c0, r);
          tmem_ld_wait();
          if (k == NCHUNK - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
          }
          if (NCHUNK != 1 && has_res) load_res_row<CH>(rbase + c0, rq);
          epi_chunk_store_rq<CH>(r, s_bias + c0, obase + c0, valid, do_stats, s1, s2, rq, has_res, P.wide != 0, P.slope);
        }
      } else {
        const bool has_res = P.res != nullptr && valid;
        mbar_wait(&tfull[acc], par);
        tc_fence_after();
#pragma unroll
        for (int cls = 0; cls < 4; ++cls) {       // column groups (0,0), (0,1), (1,1), (1,0)
          const int i = cls >= 2 ? 1 : 0, j = (cls == 1 || cls == 2) ? 1 : 0;
          bf16* orow = P.out + (long long)img * P.out_sn + (long long)(2 * ph + i) * P.out_sh + (long long)(2 * pw + j) * P.out_sw;
          uint32_t rq[CH / 2];
          if (has_res)
            load_res_row<CH>(P.res + (long long)img * P.res_sn + (long long)(2 * ph + i) * P.res_sh +
                                 (long long)(2 * pw + j) * P.res_sw, rq);
          uint32_t r[CH];
          if (CH == 32) tmem_ld_32x32(t_addr + cls * NP, r);
          else tmem_ld_32x16(t_addr + cls * NP, r);
          tmem_ld_wait();
          if (cls == 3) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
          }
          epi_chunk_store_rq<CH>(r, s_bias, orow, valid, do_stats, s1, s2, rq, has_res, P.wide != 0, P.slope);
        }
      }
    }
    if (do_stats) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = j < CH ? ((j & 1) ? unpack_f32x2(s1[j / 2]).y : unpack_f32x2(s1[j / 2]).x) : 0.f;
      const float t1 = warp_transpose_reduce32(v, lane);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = j < CH ? ((j & 1) ? unpack_f32x2(s2[j / 2]).y : unpack_f32x2(s2[j / 2]).x) : 0.f;
      const float t2 = warp_transpose_reduce32(v, lane);
      if (lane < CH) { sl[cbeg + lane] = t1; sl[NP + cbeg + lane] = t2; }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (P.stats) {
    for (int i = threadIdx.x; i < 2 * NP; i += blockDim.x) {
      double s = 0.0;
#pragma unroll
      for (int e = 0; e < 8; ++e) s += (double)s_stats[e * 2 * NP + i];
      if (s != 0.0) atomicAdd(&P.stats[i], s);
    }
  }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <int MODE, int NP, int KC>
static int launch_halo_s2(const S2Params& P, const S2Maps& m, size_t smem, cudaStream_t s) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(halo_s2_kernel<MODE, NP, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    MPGAN_REQUIRE(e == cudaSuccess, MPGAN_ERR_CUDA, "cudaFuncSetAttribute(halo_s2): %s", cudaGetErrorString(e));
    attr_done = true;
  }
  const int grid = P.total_tiles < num_sms() ? P.total_tiles : num_sms();
  launch_k(halo_s2_kernel<MODE, NP, KC>, grid, kS2Threads, smem, s, P, m);
  MPGAN_CHECK_LAUNCH("halo_s2_kernel");
  return 0;
}

static bool halo_s2_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MPGAN_NO_HALO_S2"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}

// Stride-2 3x3 pad-1 layer through the halo-resident engines.  dir 0: X grid (xh, xw) -> Y grid, w = [N = cy][9][C = cx];
// dir 1: Y grid -> X grid = 2 Y, w = transposed shadow [Co = cx][9][C = cy].  Returns 1 when the layer is not covered
// (the caller falls back to the tap-GEMM kernel).
static int halo_s2_run(int dir, int n, int xh, int xw, int yh, int yw, int C, int N, const void* in, int64_t ldi,
                       const void* w, const float* bias, void* out, int64_t ldo, double* stats, const void* res,
                       int64_t ldres, cudaStream_t s, const float* slope) {
  if (!halo_s2_enabled()) return 1;
  if (xh != 2 * yh || xw != 2 * yw) return 1;
  if (ldi % 8 != 0 || ((uintptr_t)in & 15) || ((uintptr_t)w & 15)) return 1;
  if (ldo % 8 != 0 || ((uintptr_t)out & 15)) return 1;
  if (res && (ldres % 8 != 0 || ((uintptr_t)res & 15))) return 1;
  int KC;
  if (dir == 0) {
    if (!(C == 16 || C == 32)) return 1;
    if (!(N == 32 || N == 64 || N == 128 || N == 192)) return 1;
    if (stats && N > 64) return 1;          // statistics registers exist for column spans <= 32 (see the kernel)
    KC = C;
  } else {
    if (!(N == 16 || N == 32)) return 1;
    if (!(C == 32 || (C % 64 == 0 && C <= 512))) return 1;
    KC = C < 64 ? C : 64;
  }
  const int prows = dir == 0 ? 2 * S2_HH * S2_HW : S2_HH * S2_HW;
  const size_t w_bytes = ((size_t)9 * C * N * 2 + 1023) & ~(size_t)1023;
  const size_t a_bytes = ((size_t)prows * KC * 2 + 1023) & ~(size_t)1023;
  const size_t aux = kS2Aux + (size_t)8 * 2 * N * 4 + (size_t)N * 4;
  const size_t budget = 227 * 1024 - 1024;
  const int planes = dir == 0 ? 2 : C / KC;
  const int min_buf = planes + 1 > 3 ? planes + 1 : 3;
  if (w_bytes + (size_t)min_buf * a_bytes + aux > budget) return 1;
  int nbuf = (int)((budget - w_bytes - aux) / a_bytes);
  if (nbuf > kS2MaxBuf) nbuf = kS2MaxBuf;
  const int tiles_w = (yw + S2_TW - 1) / S2_TW, tiles_h = (yh + S2_TH - 1) / S2_TH;
  const int total = n * tiles_w * tiles_h;
  if (total <= 32 * num_sms()) {   // generator-sized layer: leave room for a CTA of the concurrent weight-gradient stream
    const int cap = (int)(((size_t)100 * 1024 > w_bytes + aux ? (size_t)100 * 1024 - w_bytes - aux : 0) / a_bytes);
    const int lim = cap > min_buf ? cap : min_buf;
    if (nbuf > lim) nbuf = lim;
  }
  S2Params P;
  memset(&P, 0, sizeof(P));
  P.nimg = n; P.ch = yh; P.cw = yw; P.C = C;
  P.tiles_w = tiles_w; P.tiles_h = tiles_h; P.total_tiles = total; P.nbuf = nbuf;
  P.w = (const bf16*)w; P.out = (bf16*)out; P.bias = bias; P.stats = stats; P.slope = slope;
  const int oh = dir == 0 ? yh : xh, ow = dir == 0 ? yw : xw;
  P.out_sn = (long long)oh * ow * ldo; P.out_sh = (long long)ow * ldo; P.out_sw = ldo;
  P.res = (const bf16*)res;
  P.res_sn = (long long)oh * ow * ldres; P.res_sh = (long long)ow * ldres; P.res_sw = ldres;
  P.wide = (ldo % 16 == 0 && ((uintptr_t)out & 31) == 0) ? 1 : 0;
  S2Maps maps;
  memset(&maps, 0, sizeof(maps));
  if (dir == 0) {
    for (int wp = 0; wp < 2; ++wp) {   // column-parity views of X: pixel stride 2 ld, all rows
      uint64_t dims[4] = {(uint64_t)C, (uint64_t)(xw / 2), (uint64_t)xh, (uint64_t)n};
      uint64_t str[3] = {(uint64_t)ldi * 2, (uint64_t)ldi * xw, (uint64_t)ldi * xw * xh};
      uint32_t box[4] = {(uint32_t)KC, (uint32_t)S2_HW, (uint32_t)(2 * S2_HH), 1u};
      int rc = encode_map(&maps.m[wp], (const bf16*)in + (int64_t)wp * ldi, 4, dims, str, box);
      if (rc) return rc;
    }
  } else {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)yw, (uint64_t)yh, (uint64_t)n};
    uint64_t str[3] = {(uint64_t)ldi, (uint64_t)ldi * yw, (uint64_t)ldi * yw * yh};
    uint32_t box[4] = {(uint32_t)KC, (uint32_t)S2_HW, (uint32_t)S2_HH, 1u};
    int rc = encode_map(&maps.m[0], in, 4, dims, str, box);
    if (rc) return rc;
    maps.m[1] = maps.m[0];
  }
  const size_t smem = w_bytes + (size_t)nbuf * a_bytes + aux + 1024;
  if (dir == 0) {
#define S2_M0(NN)                                                        \
  if (KC == 16) return launch_halo_s2<0, NN, 16>(P, maps, smem, s);      \
  return launch_halo_s2<0, NN, 32>(P, maps, smem, s);
    switch (N) {
      case 32: S2_M0(32)
      case 64: S2_M0(64)
      case 128: S2_M0(128)
      default: S2_M0(192)
    }
#undef S2_M0
  }
#define S2_M1(NN)                                                        \
  if (KC == 32) return launch_halo_s2<1, NN, 32>(P, maps, smem, s);      \
  return launch_halo_s2<1, NN, 64>(P, maps, smem, s);
  switch (N) {
    case 16: S2_M1(16)
    default: S2_M1(32)
  }
#undef S2_M1
}

}  // namespace tc
}  // namespace mpgan
