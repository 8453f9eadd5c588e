// Weight gradient of the generator's small-channel 3x3 layers (16 / 32 input channels) with ONE halo box of X per tile.
// Included by conv_tc.cu.
//
// Why: wgrad_kernel loads, per 64-pixel tile, the dY tile padded to 128 channel rows (two boxes) and one X tile PER TAP
// (nine boxes of 64 rows x 32 / 64 bytes): 1 400 TMA rows per 128 pixels for 8 KB of useful data.  Warm-cache launch list
// of the generator's backward (profiles/launches_r2_gbwd.md): the 16 / 32-channel weight gradients are 24 % of its device
// time at 0.9 TB/s.  Here a tile is 16 x 8 pixels of the Y grid, the X halo (18 x 10 pixels) arrives as one box and dY as
// one box: 308 rows.  Both operands are MN-major straight from NHWC (pixels are K):
//     D[(tap atom, ci)][co] += sum_pixels  Xview[pixel + tap][ci] * dY[pixel][co]
//   * A = the X halo: the 8 consecutive pixels of a tile row are the 8 K rows of a swizzle atom, the next tile row is the
//     next K group (SBO = one halo row), and the next M atom is LBO = ONE PIXEL further: atoms 0..2 of one MMA are the
//     three taps of a filter row (stride 1) / filter column (stride 2: LBO = one plane row); the remaining atoms of the
//     M = 128 operand read neighbouring pixels and are discarded.
//   * B = the dY tile, N = cy (one atom).
//   Three MMAs per 16 pixels (one per filter row / per (column parity, column)), accumulators (3 x cy columns) resident in
//   TMEM for the whole kernel; fp32 atomics into the flat gradient buffer at the end (contiguous per warp).
// MODE 0: stride 1, pad 0 / 1.   MODE 1: stride 2, pad 1, X = 2 Y: X arrives as two column-parity planes (conv_s2.cuh).
// Replaces cuDNN's wgrad behind MONAI's Convolution / ResidualUnit (/root/reference/code/GAN/GAN_final.py:106-114).
#pragma once

namespace mpgan {
namespace tc {

constexpr int kWhStagesMax = 8;

struct WgHaloParams {
  int nimg, yh, yw, pad;
  int tiles_w, tiles_h, total_tiles;
  float* dw;   // [cy][9][cx] fp32
};

template <int MODE, int CX, int CY>
__global__ void __launch_bounds__(256, 1)
wgrad_halo_kernel(const __grid_constant__ WgHaloParams P, const __grid_constant__ S2Maps tmX,
                  const __grid_constant__ CUtensorMap tmY) {
  constexpr int rowb_x = CX * 2, rowb_y = CY * 2;
  constexpr uint32_t lay_x = CX == 64 ? 2u : (CX == 32 ? 4u : 6u);
  constexpr uint32_t lay_y = CY == 64 ? 2u : (CY == 32 ? 4u : 6u);
  constexpr int NATOM = 128 / CX;                                // M atoms of one MMA (three are taps)
  // plane geometry in pixel rows: K-group pitch (one tile row further), atom pitch (next tap), box rows
  constexpr int KG = MODE == 0 ? HW : 2 * S2_HW;                 // 10 / 18
  constexpr int AT = MODE == 0 ? 1 : S2_HW;                      // 1 / 9
  constexpr int BOXROWS = MODE == 0 ? HPIX : 2 * S2_HH * S2_HW;  // 180 / 306
  constexpr int NPLANE = MODE == 0 ? 1 : 2;
  // furthest pixel row any (discarded) atom touches: last K group + last atom + 8 rows of the group + the variant offset
  constexpr int REACH = MODE == 0 ? (15 + 2) * KG + (NATOM - 1) * AT + 8 : (2 * 15 + 1) * S2_HW + 1 + (NATOM - 1) * AT + 8;
  constexpr int XROWS = REACH > BOXROWS ? REACH : BOXROWS;
  constexpr int XB = (XROWS * rowb_x + 1023) & ~1023;            // one plane incl. the slack the discarded atoms read
  constexpr int YB = 128 * rowb_y;
  constexpr int STAGE = NPLANE * XB + YB;
  constexpr int STAGES = (96 * 1024) / STAGE > kWhStagesMax ? kWhStagesMax : ((96 * 1024) / STAGE < 3 ? 3 : (96 * 1024) / STAGE);
  static_assert(STAGES * STAGE + 256 <= 227 * 1024, "pipeline does not fit");
  constexpr int NCOL = 3 * CY;
  constexpr int TMEM_COLS = NCOL <= 64 ? 64 : (NCOL <= 128 ? 128 : 256);
  static_assert(NCOL <= 256, "accumulators exceed half of TMEM");

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  uint64_t* empty = full + STAGES;
  uint64_t* done = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && elect_one()) {
    prefetch_tmap(&tmX.m[0]);
    if (MODE == 1) prefetch_tmap(&tmX.m[1]);
    prefetch_tmap(&tmY);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
  // the slack rows behind each plane are read by the discarded atoms only, but must not hold NaN-free garbage for the
  // accumulator rows that ARE stored: they are separate rows (M), so no initialisation is needed
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch();

  if (warp == 0) {
    if (elect_one()) {  // ================= TMA producer =================
      int stage = 0;
      uint32_t phase = 0;
      TileWalk<3> tw;
      { const int radix[3] = {P.tiles_w, P.tiles_h, 1 << 30}; tw.init((int)blockIdx.x, (int)gridDim.x, radix); }
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, tw.next()) {
        const int w0 = tw.d[0] * HT_W, h0 = tw.d[1] * HT_H, img = tw.d[2];
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sx = smem + stage * STAGE;
        mbar_expect_tx(&full[stage], (uint32_t)(NPLANE * BOXROWS * rowb_x + YB));
        if (MODE == 0) {
          tma_load_4d(sx, &tmX.m[0], &full[stage], 0, w0 - P.pad, h0 - P.pad, img);
        } else {
          tma_load_4d(sx, &tmX.m[0], &full[stage], 0, w0 - 1, 2 * (h0 - 1), img);
          tma_load_4d(sx + XB, &tmX.m[1], &full[stage], 0, w0 - 1, 2 * (h0 - 1), img);
        }
        tma_load_4d(sx + NPLANE * XB, &tmY, &full[stage], 0, w0, h0, img);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {  // ================= MMA issuer =================
      constexpr uint32_t idesc = make_idesc_bf16(128, CY, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      bool first = true;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sx = smem_u32(smem + stage * STAGE);
        const uint32_t sy = sx + NPLANE * XB;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {   // 16 pixels = two tile rows
          const uint64_t bdesc = make_smem_desc(sy + ks * 16 * rowb_y, 8 * rowb_y, 8 * rowb_y, lay_y);
#pragma unroll
          for (int v = 0; v < 3; ++v) {
            // variant v -> TMEM column block v.  stride 1: filter row dh = v (atoms = dw).  stride 2: filter column
            // kw = v (atoms = kh): kw 0 -> plane 1 col 0, kw 1 -> plane 0 col 1, kw 2 -> plane 1 col 1
            uint32_t a_addr;
            if (MODE == 0) a_addr = sx + (uint32_t)(((2 * ks + v) * KG) * rowb_x);
            else a_addr = sx + (v == 1 ? 0u : (uint32_t)XB) + (uint32_t)((((2 * (2 * ks) + 1) * S2_HW) + (v == 0 ? 0 : 1)) * rowb_x);
            const uint64_t adesc = make_smem_desc(a_addr, AT * rowb_x, KG * rowb_x, lay_x);
            umma_bf16(tmem_base + (uint32_t)(v * CY), adesc, bdesc, idesc, (first && ks == 0) ? 0u : 1u);
          }
        }
        umma_commit(&empty[stage]);
        first = false;
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(done);
    }
  } else if (warp >= 4) {  // ================= epilogue: TMEM -> fp32 atomics =================
    const int q = warp - 4;
    const int m = q * 32 + lane;
    const int atom = m / CX, ci = m - atom * CX;
    mbar_wait(done, 0);
    tc_fence_after();
    if (q * 32 < 3 * CX) {   // warps whose lanes hold tap atoms
      constexpr int CH = CY >= 32 ? 32 : 16;
#pragma unroll 1
      for (int v = 0; v < 3; ++v) {
        const int tap = MODE == 0 ? v * 3 + atom : atom * 3 + v;
#pragma unroll 1
        for (int c0 = 0; c0 < CY; c0 += CH) {
          uint32_t r[CH];
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(v * CY + c0);
          if (CH == 32) tmem_ld_32x32(ta, r);
          else tmem_ld_32x16(ta, r);
          tmem_ld_wait();
          if (atom < 3) {
            float* d = P.dw + (size_t)tap * CX + ci;
#pragma unroll
            for (int j = 0; j < CH; ++j) atomicAdd(d + (size_t)(c0 + j) * 9 * CX, __uint_as_float(r[j]));
          }
        }
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <int MODE, int CX, int CY>
static int launch_wgrad_halo(const WgHaloParams& P, const S2Maps& mX, const CUtensorMap& mY, cudaStream_t s) {
  constexpr int rowb_x = CX * 2, rowb_y = CY * 2;
  constexpr int NATOM = 128 / CX;
  constexpr int KG = MODE == 0 ? HW : 2 * S2_HW, AT = MODE == 0 ? 1 : S2_HW;
  constexpr int BOXROWS = MODE == 0 ? HPIX : 2 * S2_HH * S2_HW;
  constexpr int REACH = MODE == 0 ? (15 + 2) * KG + (NATOM - 1) * AT + 8 : (2 * 15 + 1) * S2_HW + 1 + (NATOM - 1) * AT + 8;
  constexpr int XROWS = REACH > BOXROWS ? REACH : BOXROWS;
  constexpr int XB = (XROWS * rowb_x + 1023) & ~1023;
  constexpr int STAGE = (MODE == 0 ? 1 : 2) * XB + 128 * rowb_y;
  constexpr int STAGES = (96 * 1024) / STAGE > kWhStagesMax ? kWhStagesMax : ((96 * 1024) / STAGE < 3 ? 3 : (96 * 1024) / STAGE);
  constexpr int SMEM = STAGES * STAGE + 256;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_halo_kernel<MODE, CX, CY>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    MPGAN_REQUIRE(e == cudaSuccess, MPGAN_ERR_CUDA, "cudaFuncSetAttribute(wgrad_halo): %s", cudaGetErrorString(e));
    attr_done = true;
  }
  const int grid = P.total_tiles < num_sms() ? P.total_tiles : num_sms();
  launch_k(wgrad_halo_kernel<MODE, CX, CY>, grid, 256, SMEM, s, P, mX, mY);
  MPGAN_CHECK_LAUNCH("wgrad_halo_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// The same idea for 64 input channels (stride 1): D layer 2's weight gradient (64 x 128, 254^2, batch 32) and the
// generator's 64-channel units.  A 128-byte pixel row is one swizzle atom, so an M = 128 operand holds two taps (dw, dw + 1):
// two MMAs per filter row -- atoms (0, 1) and (2, [3: discarded]).  The accumulators of all nine taps (9 x 64 x cy fp32)
// exceed TMEM for cy = 128, so the three filter rows go to three CTAs (NG = 3, as wgrad_kernel's tap groups do) that walk
// the same tiles side by side; a CTA then loads only the 16 halo rows of its filter row.  Operand bytes per 128 pixels:
// 20 KB of X + 2 cy x 128 B of dY, against 3 x 8 KB of X taps + the dY tile per 64 pixels before.
// ---------------------------------------------------------------------------------------------------------
template <int CY, int NG>
__global__ void __launch_bounds__(256, 1)
wgrad_halo64_kernel(const __grid_constant__ WgHaloParams P, const __grid_constant__ CUtensorMap tmX,
                    const __grid_constant__ CUtensorMap tmY) {
  constexpr int CX = 64, rowb_x = 128;
  constexpr int NDH = 3 / NG;                                    // filter rows per CTA
  constexpr int XROWS_BOX = (HT_H + NDH - 1) * HW;               // 160 (one filter row) / 180 pixel rows
  constexpr int REACH = (15 + NDH - 1) * HW + 3 + 8;             // atoms 2, 3 of the second MMA
  constexpr int XROWS = REACH > XROWS_BOX ? REACH : XROWS_BOX;
  constexpr int XB = (XROWS * rowb_x + 1023) & ~1023;
  constexpr int NYB = CY / 64;                                   // 64-channel boxes of dY
  constexpr int YB = NYB * 128 * 128;
  constexpr int STAGE = XB + YB;
  constexpr int STAGES = (220 * 1024) / STAGE > kWhStagesMax ? kWhStagesMax : (220 * 1024) / STAGE;
  static_assert(STAGES >= 3, "pipeline too shallow");
  constexpr int NCOL = NDH * 2 * CY;
  constexpr int TMEM_COLS = NCOL <= 128 ? 128 : (NCOL <= 256 ? 256 : 512);
  static_assert(NCOL <= 512, "accumulators exceed TMEM");

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  uint64_t* empty = full + STAGES;
  uint64_t* done = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && elect_one()) { prefetch_tmap(&tmX); prefetch_tmap(&tmY); }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch();

  const int grp = NG == 1 ? 0 : (int)(blockIdx.x % NG);          // this CTA's filter row (NG == 3)
  const int cta = (int)blockIdx.x / NG, ncta = (int)gridDim.x / NG;

  if (warp == 0) {
    if (elect_one()) {  // ================= TMA producer =================
      int stage = 0;
      uint32_t phase = 0;
      TileWalk<3> tw;
      { const int radix[3] = {P.tiles_w, P.tiles_h, 1 << 30}; tw.init(cta, ncta, radix); }
      for (int tile = cta; tile < P.total_tiles; tile += ncta, tw.next()) {
        const int w0 = tw.d[0] * HT_W, h0 = tw.d[1] * HT_H, img = tw.d[2];
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sx = smem + stage * STAGE;
        mbar_expect_tx(&full[stage], (uint32_t)(XROWS_BOX * rowb_x + YB));
        tma_load_4d(sx, &tmX, &full[stage], 0, w0 - P.pad, h0 - P.pad + grp, img);
#pragma unroll
        for (int b = 0; b < NYB; ++b) tma_load_4d(sx + XB + b * 16384, &tmY, &full[stage], b * 64, w0, h0, img);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {  // ================= MMA issuer =================
      constexpr uint32_t idesc = make_idesc_bf16(128, CY, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      bool first = true;
      for (int tile = cta; tile < P.total_tiles; tile += ncta) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sx = smem_u32(smem + stage * STAGE);
        const uint32_t sy = sx + XB;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t bdesc = make_smem_desc(sy + ks * 16 * 128, 16384, 1024, 2);      // N atoms = the 64-channel boxes
#pragma unroll
          for (int dh = 0; dh < NDH; ++dh) {
#pragma unroll
            for (int v = 0; v < 2; ++v) {   // atoms (dw 0, 1) / (dw 2, discarded)
              const uint32_t a_addr = sx + (uint32_t)(((2 * ks + dh) * HW + 2 * v) * rowb_x);
              const uint64_t adesc = make_smem_desc(a_addr, rowb_x, HW * rowb_x, 2);
              umma_bf16(tmem_base + (uint32_t)((dh * 2 + v) * CY), adesc, bdesc, idesc, (first && ks == 0) ? 0u : 1u);
            }
          }
        }
        umma_commit(&empty[stage]);
        first = false;
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(done);
    }
  } else if (warp >= 4) {  // ================= epilogue: TMEM -> fp32 atomics =================
    const int q = warp - 4;
    const int m = q * 32 + lane;
    const int atom = m >> 6, ci = m & 63;
    if (cta < P.total_tiles) {
      mbar_wait(done, 0);
      tc_fence_after();
#pragma unroll 1
      for (int dh = 0; dh < NDH; ++dh) {
#pragma unroll 1
        for (int v = 0; v < 2; ++v) {
          const int dw = 2 * v + atom;
          const int tap = (NG == 1 ? dh : grp) * 3 + dw;
#pragma unroll 1
          for (int c0 = 0; c0 < CY; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((dh * 2 + v) * CY + c0), r);
            tmem_ld_wait();
            if (dw < 3) {
              float* d = P.dw + (size_t)tap * CX + ci;
#pragma unroll
              for (int j = 0; j < 32; ++j) atomicAdd(d + (size_t)(c0 + j) * 9 * CX, __uint_as_float(r[j]));
            }
          }
        }
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <int CY, int NG>
static int launch_wgrad_halo64(const WgHaloParams& P, const CUtensorMap& mX, const CUtensorMap& mY, cudaStream_t s) {
  constexpr int NDH = 3 / NG;
  constexpr int XROWS_BOX = (HT_H + NDH - 1) * HW;
  constexpr int REACH = (15 + NDH - 1) * HW + 3 + 8;
  constexpr int XROWS = REACH > XROWS_BOX ? REACH : XROWS_BOX;
  constexpr int XB = (XROWS * 128 + 1023) & ~1023;
  constexpr int STAGE = XB + (CY / 64) * 128 * 128;
  constexpr int STAGES = (220 * 1024) / STAGE > kWhStagesMax ? kWhStagesMax : (220 * 1024) / STAGE;
  constexpr int SMEM = STAGES * STAGE + 256;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_halo64_kernel<CY, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    MPGAN_REQUIRE(e == cudaSuccess, MPGAN_ERR_CUDA, "cudaFuncSetAttribute(wgrad_halo64): %s", cudaGetErrorString(e));
    attr_done = true;
  }
  int per = num_sms() / NG;
  if (per > P.total_tiles) per = P.total_tiles;
  launch_k(wgrad_halo64_kernel<CY, NG>, per * NG, 256, SMEM, s, P, mX, mY);
  MPGAN_CHECK_LAUNCH("wgrad_halo64_kernel");
  return 0;
}

static int wgrad_halo64_run(int pad, int n, int xh, int xw, int yh, int yw, int cy, const void* x, int64_t ldx, const void* y,
                            int64_t ldy, float* dw, cudaStream_t s) {
  if (!(cy == 64 || cy == 128)) return 1;
  if (ldx % 8 != 0 || ldy % 8 != 0 || ((uintptr_t)x & 15) || ((uintptr_t)y & 15)) return 1;
  if (pad < 0 || pad > 1 || yh != xh + 2 * pad - 2 || yw != xw + 2 * pad - 2) return 1;
  WgHaloParams P;
  memset(&P, 0, sizeof(P));
  P.nimg = n; P.yh = yh; P.yw = yw; P.pad = pad;
  P.tiles_w = (yw + HT_W - 1) / HT_W; P.tiles_h = (yh + HT_H - 1) / HT_H;
  P.total_tiles = n * P.tiles_w * P.tiles_h;
  P.dw = dw;
  CUtensorMap mX, mY;
  const int ndh = cy == 128 ? 1 : 3;
  {
    uint64_t dims[4] = {64u, (uint64_t)xw, (uint64_t)xh, (uint64_t)n};
    uint64_t str[3] = {(uint64_t)ldx, (uint64_t)ldx * xw, (uint64_t)ldx * xw * xh};
    uint32_t box[4] = {64u, (uint32_t)HW, (uint32_t)(HT_H + ndh - 1), 1u};
    int rc = encode_map(&mX, x, 4, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)cy, (uint64_t)yw, (uint64_t)yh, (uint64_t)n};
    uint64_t str[3] = {(uint64_t)ldy, (uint64_t)ldy * yw, (uint64_t)ldy * yw * yh};
    uint32_t box[4] = {64u, (uint32_t)HT_W, (uint32_t)HT_H, 1u};
    int rc = encode_map(&mY, y, 4, dims, str, box);
    if (rc) return rc;
  }
  if (cy == 128) return launch_wgrad_halo64<128, 3>(P, mX, mY, s);
  return launch_wgrad_halo64<64, 1>(P, mX, mY, s);
}

static bool wgrad_halo_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MPGAN_NO_WGRAD_HALO"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}

// dw[cy][9][cx] += the weight gradient of a rank-2 3x3 layer X(cx) -> Y(cy).  stride 1 (pad 0 / 1) or stride 2 (pad 1,
// X = 2 Y).  Returns 1 when the layer is not covered (the caller runs wgrad_kernel).
static int wgrad_halo_run(int stride, int pad, int n, int xh, int xw, int yh, int yw, int cx, int cy, const void* x,
                          int64_t ldx, const void* y, int64_t ldy, float* dw, cudaStream_t s) {
  if (!wgrad_halo_enabled()) return 1;
  if (cx == 64 && stride == 1) return wgrad_halo64_run(pad, n, xh, xw, yh, yw, cy, x, ldx, y, ldy, dw, s);
  if (!(cx == 16 || cx == 32) || !(cy == 16 || cy == 32 || cy == 64)) return 1;
  if (ldx % 8 != 0 || ldy % 8 != 0 || ((uintptr_t)x & 15) || ((uintptr_t)y & 15)) return 1;
  if (stride == 1) {
    if (pad < 0 || pad > 1 || yh != xh + 2 * pad - 2 || yw != xw + 2 * pad - 2) return 1;
  } else {
    if (stride != 2 || pad != 1 || xh != 2 * yh || xw != 2 * yw) return 1;
  }
  WgHaloParams P;
  memset(&P, 0, sizeof(P));
  P.nimg = n; P.yh = yh; P.yw = yw; P.pad = pad;
  P.tiles_w = (yw + HT_W - 1) / HT_W; P.tiles_h = (yh + HT_H - 1) / HT_H;
  P.total_tiles = n * P.tiles_w * P.tiles_h;
  P.dw = dw;
  S2Maps mX;
  CUtensorMap mY;
  memset(&mX, 0, sizeof(mX));
  int rc;
  if (stride == 1) {
    uint64_t dims[4] = {(uint64_t)cx, (uint64_t)xw, (uint64_t)xh, (uint64_t)n};
    uint64_t str[3] = {(uint64_t)ldx, (uint64_t)ldx * xw, (uint64_t)ldx * xw * xh};
    uint32_t box[4] = {(uint32_t)cx, (uint32_t)HW, (uint32_t)HH, 1u};
    rc = encode_map(&mX.m[0], x, 4, dims, str, box);
    if (rc) return rc;
    mX.m[1] = mX.m[0];
  } else {
    for (int wp = 0; wp < 2; ++wp) {
      uint64_t dims[4] = {(uint64_t)cx, (uint64_t)(xw / 2), (uint64_t)xh, (uint64_t)n};
      uint64_t str[3] = {(uint64_t)ldx * 2, (uint64_t)ldx * xw, (uint64_t)ldx * xw * xh};
      uint32_t box[4] = {(uint32_t)cx, (uint32_t)S2_HW, (uint32_t)(2 * S2_HH), 1u};
      rc = encode_map(&mX.m[wp], (const bf16*)x + (int64_t)wp * ldx, 4, dims, str, box);
      if (rc) return rc;
    }
  }
  {
    uint64_t dims[4] = {(uint64_t)cy, (uint64_t)yw, (uint64_t)yh, (uint64_t)n};
    uint64_t str[3] = {(uint64_t)ldy, (uint64_t)ldy * yw, (uint64_t)ldy * yw * yh};
    uint32_t box[4] = {(uint32_t)cy, (uint32_t)HT_W, (uint32_t)HT_H, 1u};
    rc = encode_map(&mY, y, 4, dims, str, box);
    if (rc) return rc;
  }
#define WGH_CY(MODE_, CX_)                                                          \
  switch (cy) {                                                                     \
    case 16: return launch_wgrad_halo<MODE_, CX_, 16>(P, mX, mY, s);                \
    case 32: return launch_wgrad_halo<MODE_, CX_, 32>(P, mX, mY, s);                \
    default: return launch_wgrad_halo<MODE_, CX_, 64>(P, mX, mY, s);                \
  }
  if (stride == 1) {
    if (cx == 16) { WGH_CY(0, 16) }
    WGH_CY(0, 32)
  }
  if (cx == 16) { WGH_CY(1, 16) }
  WGH_CY(1, 32)
#undef WGH_CY
}

}  // namespace tc
}  // namespace mpgan
