// "Halo-resident" tcgen05 implicit GEMM for stride-1 3x3 convolutions (forward, and the data gradient of a stride-1
// 3x3 convolution, which is the same operation with flipped taps and transposed weights).  Included by conv_tc.cu.
//
// Why a second conv engine: tapgemm_kernel re-reads the activation tile once per filter tap through TMA (9 x 128
// rows per 128 output pixels).  For the generator's 16/32-channel layers that is TMA-row-rate bound (32/64-byte
// rows), and for the discriminator's 64->128 layer it is L2->SM bandwidth bound.  Here every CTA
//   * keeps ALL weights of the layer resident in shared memory (9*C*N bf16 <= 147 KB), loaded once,
//   * loads each activation halo tile ((16+2) x (8+2) pixels, <= 64 channels per plane) ONCE with a single TMA box
//     (out-of-bounds zero fill = padding) into a ring of planes; the box lands as the swizzled canonical K-major
//     UMMA operand [180 halo pixels][min(C,64) channels] (128/64/32-byte rows), and
//   * issues the 9*(C/16) tcgen05.mma of a tile straight from those planes: a filter tap is nothing but a
//     whole-pixel shift of the descriptor start address (rows of the 128-row A operand = 16 tile rows of 8
//     consecutive pixels, SBO = one halo row of 10 pixels).  The swizzle XOR is a function of the absolute
//     shared-memory address on both the TMA-write and the MMA-read side (verified on B200: base_offset = 0 with
//     start addresses that are not multiples of the swizzle period gives exact results), so shifted windows of one
//     plane are legal operands.
// Accumulators: 128 x N fp32 in TMEM, double buffered; 8 epilogue warps (2 per TMEM lane quarter, alternating
// 32-column chunks): bias, bf16 store at a channel offset, BatchNorm statistics via a shared-memory transpose.
// Replaces cuDNN behind nn.Conv2d (k3 s1) at /root/reference/code/GAN/GAN_final.py:173-176 (D layer 2) and the MONAI
// UNet's stride-1 units (GAN_final.py:106-114).
#pragma once

namespace mpgan {
namespace tc {

constexpr int HT_W = 8, HT_H = 16;              // output tile (M = 128 rows: 16 groups of 8 consecutive pixels)
constexpr int HH = HT_H + 2, HW = HT_W + 2;     // halo tile
constexpr int HPIX = HH * HW;                   // 180 pixels
constexpr int kHaloThreads = 384;               // warp 0 TMA, 1 MMA, 2 TMEM owner, 3 idle, 4-11 epilogue
constexpr int kMaxBuf = 8;
constexpr int kHaloAux = 512;                   // barriers + tmem slot
constexpr int kMaxAcc = 8;                      // TMEM accumulator stages
constexpr int kTrW = 17;                        // padded row of the per-warp transpose scratch (bf16x2 words)

struct HaloParams {
  int nimg, oh, ow;
  int C, N, pad, flip;
  int tiles_w, tiles_h, total_tiles;
  int nbuf;
  const bf16* w;       // [N][9][C]
  bf16* out;
  long long out_sn, out_sh, out_sw;
  const float* bias;
  double* stats;
  const bf16* res;     // optional tensor added to the result before rounding (same grid as out)
  long long res_sn, res_sh, res_sw;
  int wide;            // output rows are 32-byte aligned: 256-bit stores
  const float* slope;  // optional device scalar: PReLU applied to (acc + bias) before the residual (inference fusion)
  int n_store;         // 0 = all N channels; 1 = only channel 0 (one-channel output computed with a zero-padded N = 16)
  int nissue;          // MMA-issuing threads (1 or 2): issuer i takes tiles i, i + nissue, ... (independent accumulators)
  uint32_t idesc;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int N, int KC>   // KC = min(C, 64): channels per plane (one swizzle span)
__global__ void __launch_bounds__(kHaloThreads, 1)
halo3x3_kernel(const __grid_constant__ HaloParams P, const __grid_constant__ CUtensorMap tmA) {
  // accumulator ring: as many 128 x N fp32 tiles as fit in 512 TMEM columns (<= 8): small layers are bound by the
  // MMA -> epilogue -> MMA round trip, not by throughput, so the ring must be deep
  // N <= 64 (generator layers): at most half of TMEM and (host side) a capped plane ring, so that a CTA of the concurrent
  // weight-gradient stream can share the SM (a 128-register cap was tried too: it only bought spills)
  constexpr int NACC = N <= 64 ? (256 / N > kMaxAcc ? kMaxAcc : 256 / N) : 512 / N;
  constexpr int TMEM_COLS = NACC * N < 32 ? 32 : NACC * N;
  // No static shared memory in this kernel, so the dynamic window starts at the CTA's (1024-byte aligned) base; using
  // the symbol directly (instead of a manually aligned pointer) lets the compiler emit LDS/STS rather than generic
  // loads and stores for every shared-memory access of the epilogue.
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int C = P.C;
  const int nkc = C / KC;
  constexpr int rowb = KC * 2;                           // bytes per pixel row of a plane: 128 / 64 / 32
  constexpr int cpr = rowb >> 4;                         // 16-byte chunks per row
  constexpr uint32_t swz_mask = (uint32_t)(cpr - 1);     // Swizzle<log2(cpr),4,3>: chunk ^= (offset >> 7) & mask
  const int w_bytes = 9 * C * N * 2;
  constexpr int a_bytes = (HPIX * rowb + 1023) & ~1023;  // one plane, padded to the swizzle period
  uint8_t* w_sm = smem;
  uint8_t* a_sm = smem + ((w_bytes + 1023) & ~1023);
  uint8_t* aux = a_sm + P.nbuf * a_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(aux);
  uint64_t* a_empty = a_full + kMaxBuf;
  uint64_t* tfull = a_empty + kMaxBuf;
  uint64_t* tempty = tfull + kMaxAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kMaxAcc);
  float* s_stats = reinterpret_cast<float*>(aux + kHaloAux);   // [8 epilogue warps][2*N]
  float* s_bias = s_stats + 8 * 2 * N;                         // [N]
  // (no transpose scratch: the statistics of the N = 128 layers are reduced with register shuffles -- the 17 KB it took
  //  are one more activation plane in flight for the layers whose resident weights fill the shared memory, D layer 2)
  // N <= 32: the two epilogue warp groups take alternate tiles (4 arrivals free an accumulator);
  // N >= 64: they take alternate 32-column chunks of every tile (8 arrivals)
  constexpr uint32_t kEmptyArrivals = N <= 32 ? 4u : 8u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) prefetch_tmap(&tmA);
  if (threadIdx.x == 32) {
    for (int i = 0; i < kMaxBuf; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kMaxAcc; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kEmptyArrivals); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
  for (int i = threadIdx.x; i < 8 * 2 * N; i += blockDim.x) s_stats[i] = 0.f;
  pdl_wait();      // PDL: barrier init / TMEM allocation above overlap the previous kernel; global reads start here
  pdl_launch();
  // resident weights: global [n][tap][C] -> smem [tap][kc][n][KC channels], rows swizzled exactly like a TMA box
  {
    const int cpp = C >> 3;
    const int total = N * 9 * cpp;
    if (P.n_store == 4) {
      // Stride-2 ConvTranspose / data gradient into ONE channel (k3 s2 p1, X = 2Y) as a stride-1 "pixel-shuffle"
      // layer: output column n = (i, j) is the parity class of the 2x2 block written by input pixel (a, b), and it
      // reads the 2x2 neighbourhood Y(a + da, b + db) = halo taps (da + 1, db + 1):
      //   (0,0): Y00.W11   (0,1): Y01.W10 + Y00.W12   (1,0): Y10.W01 + Y00.W21   (1,1): Y11.W00 + Y10.W02 + Y01.W20 + Y00.W22
      // P.w is the layer's own [C][9] weight; the [16][9][C] operand is assembled here, zeros elsewhere.
      for (int e = threadIdx.x; e < N * 9 * C; e += blockDim.x) {
        const int c = e % C;
        const int r = e / C;
        const int tap = r % 9, n = r / 9;
        int src = -1;
        if (n == 0) src = tap == 4 ? 4 : -1;
        else if (n == 1) src = tap == 5 ? 3 : (tap == 4 ? 5 : -1);
        else if (n == 2) src = tap == 7 ? 1 : (tap == 4 ? 7 : -1);
        else if (n == 3) src = tap == 8 ? 0 : (tap == 7 ? 2 : (tap == 5 ? 6 : (tap == 4 ? 8 : -1)));
        const bf16 v = src >= 0 ? P.w[c * 9 + src] : from_f<bf16>(0.f);
        const int kc = c / KC, cc = c - kc * KC;
        const uint32_t blk = (uint32_t)((tap * nkc + kc) * N) * rowb;
        uint32_t off = (uint32_t)n * rowb + (uint32_t)(cc >> 3) * 16;
        off ^= ((off >> 7) & swz_mask) << 4;
        *reinterpret_cast<bf16*>(w_sm + blk + off + (cc & 7) * 2) = v;
      }
    } else
    for (int g = threadIdx.x; g < total; g += blockDim.x) {
      const int chunk = g % cpp;
      const int r = g / cpp;
      const int tap = r % 9, n = r / 9;
      const int kc = chunk / cpr, cc = chunk - kc * cpr;
      const uint32_t blk = (uint32_t)((tap * nkc + kc) * N) * rowb;   // multiple of the swizzle period
      uint32_t off = (uint32_t)n * rowb + cc * 16;
      off ^= ((off >> 7) & swz_mask) << 4;
      cp_async16(smem_u32(w_sm + blk + off), P.w + (size_t)g * 8);
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) s_bias[i] = P.bias ? __ldg(&P.bias[P.n_store == 4 ? 0 : i]) : 0.f;
    cp_async_wait_all();
    fence_proxy_async();   // these generic-proxy writes are read by tcgen05.mma (async proxy)
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if constexpr (N >= 128) {
    // Register re-partition (setmaxnreg): the producer / issuer warpgroup gives up registers so that every epilogue thread
    // can keep the BatchNorm statistics of its 64 columns in 128 fp32 registers for the whole kernel -- the per-chunk warp
    // transpose-reductions they replace were 60 % of the 11 500 warp-instructions per tile that made D layer 2's forward
    // epilogue-bound (ncu: 39.5 % tensor-pipe active, the MMA issuer spinning on the accumulator-empty barrier).
    // (the instructions sit at the head of the two role regions below)
  }

  if (warp < 4) {
  if constexpr (N >= 128) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    if (elect_one()) {  // ================= TMA producer: one box per (tile, 64-channel plane) =================
      int pbuf = 0;
      uint32_t ppar = 0;
      TileWalk<3> tw;
      { const int radix[3] = {P.tiles_w, P.tiles_h, 1 << 30}; tw.init((int)blockIdx.x, (int)gridDim.x, radix); }
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, tw.next()) {
        const int twi = tw.d[0], thi = tw.d[1], img = tw.d[2];
        const int h0 = thi * HT_H - P.pad, w0 = twi * HT_W - P.pad;
        for (int kc = 0; kc < nkc; ++kc) {
          mbar_wait(&a_empty[pbuf], ppar ^ 1);
          mbar_expect_tx(&a_full[pbuf], (uint32_t)(HPIX * rowb));
          tma_load_4d(a_sm + pbuf * a_bytes, &tmA, &a_full[pbuf], kc * KC, w0, h0, img);
          if (++pbuf == P.nbuf) { pbuf = 0; ppar ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    const int iss = warp == 1 ? 0 : 1;
    if (iss < P.nissue && elect_one()) {  // ================= MMA issuer(s) =================
      // ONE thread issues every MMA of a tile, so for small N (a 128 x 16 x 16 MMA occupies the tensor pipe for
      // ~8 cycles) its instruction stream -- and the accumulate-into-the-same-TMEM-tile dependency between the taps of a
      // tile -- is the critical path: descriptors are a precomputed 64-bit base plus compile-time offsets (fully
      // unrolled taps / k-steps), ring indices advance without divisions, and with P.nissue == 2 a second thread (warp 3)
      // issues the odd tiles: two independent accumulation chains in flight.
      constexpr uint32_t layout = KC == 64 ? 2u : (KC == 32 ? 4u : 6u);   // SW128 / SW64 / SW32
      constexpr uint32_t sbo_a = HW * rowb, sbo_b = 8 * rowb;
      constexpr int ksteps = KC >> 4;
      constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
      const uint64_t adesc0 = make_smem_desc(smem_u32(a_sm), 16, sbo_a, layout);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(w_sm), 16, sbo_b, layout);
      uint32_t boff[9];                                  // weight block of tap t (taps flipped for the data gradient)
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) boff[tap] = (uint32_t)(((P.flip ? 8 - tap : tap) * nkc * N * rowb) >> 4);
      int acc = 0, buf = 0;
      uint32_t tpar = 0, apar = 0;
      const int nissue = P.nissue;
      auto skip_tile = [&]() {   // advance the ring positions over a tile that the other issuer handles
        for (int kc = 0; kc < nkc; ++kc) { if (++buf == P.nbuf) { buf = 0; apar ^= 1; } }
        if (++acc == NACC) { acc = 0; tpar ^= 1; }
      };
      for (int i = 0; i < iss; ++i) skip_tile();
      for (int tile = blockIdx.x + iss * gridDim.x; tile < P.total_tiles; tile += nissue * gridDim.x) {
        mbar_wait(&tempty[acc], tpar ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * N);
        for (int kc = 0; kc < nkc; ++kc) {
          mbar_wait(&a_full[buf], apar);
          tc_fence_after();
          const uint64_t ad = adesc0 + (uint64_t)(uint32_t)(buf * (a_bytes >> 4));
          const uint64_t bd = bdesc0 + (uint64_t)(uint32_t)((kc * N * rowb) >> 4);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint64_t adesc = ad + (uint64_t)((((tap / 3) * HW + (tap % 3)) * rowb + ks * 32) >> 4);
              const uint64_t bdesc = bd + (uint64_t)(boff[tap] + (uint32_t)((ks * 32) >> 4));
              if (tap == 0 && ks == 0) umma_bf16(d_tmem, adesc, bdesc, idesc, kc != 0 ? 1u : 0u);
              else umma_bf16(d_tmem, adesc, bdesc, idesc, 1u);
            }
          }
          umma_commit(&a_empty[buf]);
          if (++buf == P.nbuf) { buf = 0; apar ^= 1; }
        }
        umma_commit(&tfull[acc]);
        if (++acc == NACC) { acc = 0; tpar ^= 1; }
        for (int i = 1; i < nissue; ++i) skip_tile();
      }
    }
  }
  } else {  // ================= epilogue (8 warps) =================
    if constexpr (N >= 128) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int ew = warp - 4;
    const int q = ew & 3;          // TMEM lane quarter (== warp % 4)
    const int half = ew >> 2;      // alternating column chunks
    const int row = q * 32 + lane;
    const int lh = row >> 3, lw = row & 7;
    float* sl = s_stats + ew * 2 * N;
    constexpr int CH = N >= 32 ? 32 : 16;
    if constexpr (N <= 64) {
      // Small layers: minimal per-tile chain -- one TMEM load, the accumulator is handed back to the MMA warp as soon
      // as the values are in registers, and the BatchNorm statistics are plain per-thread register accumulators
      // (thread = tile row, fixed tile order => reproducible) that are transposed and reduced ONCE, after the last tile.
      constexpr int NCH = N / CH;                 // 1 (N = 16, 32) or 2 (N = 64)
      const int c0 = NCH == 2 ? half * CH : 0;
      unsigned long long s1[CH / 2], s2[CH / 2];   // packed fp32 pairs
#pragma unroll
      for (int j = 0; j < CH / 2; ++j) { s1[j] = 0ull; s2[j] = 0ull; }
      int it = 0;
      TileWalk<3> tw, twa;     // twa runs kResAhead tiles ahead: L2 prefetch of the residual rows (cold 134 MB tensors at inference size)
      constexpr int kResAhead = 4;
      { const int radix[3] = {P.tiles_w, P.tiles_h, 1 << 30}; tw.init((int)blockIdx.x, (int)gridDim.x, radix);
        twa.init((int)blockIdx.x + kResAhead * (int)gridDim.x, (int)gridDim.x, radix); }
      const bool res_pf = P.res != nullptr && P.n_store == 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it, tw.next(), twa.next()) {
        if (NCH == 1 && (it & 1) != half) continue;
        if (res_pf && tile + kResAhead * (int)gridDim.x < P.total_tiles) {
          const int oha = twa.d[1] * HT_H + lh, owa = twa.d[0] * HT_W + lw;
          if (oha < P.oh && owa < P.ow)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(P.res + (long long)twa.d[2] * P.res_sn + (long long)oha * P.res_sh +
                                                          (long long)owa * P.res_sw + c0));
        }
        const int twi = tw.d[0], thi = tw.d[1], img = tw.d[2];
        const int oh = thi * HT_H + lh, ow = twi * HT_W + lw;
        const bool valid = oh < P.oh && ow < P.ow;
        bf16* orow = P.out + (long long)img * P.out_sn + (long long)oh * P.out_sh + (long long)ow * P.out_sw + c0;
        const int acc = it % NACC;
        const uint32_t par = (uint32_t)(it / NACC) & 1u;
        const bf16* rrow = P.res ? P.res + (long long)img * P.res_sn + (long long)oh * P.res_sh + (long long)ow * P.res_sw + c0
                                 : nullptr;
        // the residual row is requested BEFORE the wait for the accumulator: its DRAM latency overlaps the MMAs of the tile
        // instead of following them (ncu, 16 -> 16 inference layer: 31 % of the samples sat on this load)
        uint32_t rq[CH / 2];
        const bool pre_res = rrow != nullptr && valid && P.n_store == 0;
        if (pre_res) load_res_row<CH>(rrow, rq);
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
        if (P.n_store == 4) {   // 2x2 pixel shuffle into a one-channel image of twice the size (+ its BatchNorm statistics)
          if (c0 == 0) {
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = __uint_as_float(r[e]) + s_bias[0];
            __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
            if (valid) {
              bf16* o = P.out + (long long)img * P.out_sn + (long long)(2 * oh) * P.out_sh + 2 * ow;
              *reinterpret_cast<__nv_bfloat162*>(o) = h0;
              *reinterpret_cast<__nv_bfloat162*>(o + P.out_sh) = h1;
              if (P.stats) {
                const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
                const float sa = (f0.x + f0.y) + (f1.x + f1.y);
                const float sb = fmaf(f0.x, f0.x, f0.y * f0.y) + fmaf(f1.x, f1.x, f1.y * f1.y);
                s1[0] = fma_f32x2(pack_f32x2(sa, 0.f), pack_f32x2(1.f, 1.f), s1[0]);
                s2[0] = fma_f32x2(pack_f32x2(sb, 0.f), pack_f32x2(1.f, 1.f), s2[0]);
              }
            }
          }
          continue;
        }
        if (P.n_store == 1) {   // one real output channel (data gradient of a one-input-channel layer)
          if (valid && c0 == 0) {
            float v = __uint_as_float(r[0]) + s_bias[0];
            if (rrow) v += to_f(rrow[0]);
            orow[0] = from_f<bf16>(v);
          }
          continue;
        }
        epi_chunk_store_rq<CH>(r, s_bias + c0, orow, valid, P.stats != nullptr, s1, s2, rq, pre_res, P.wide != 0, P.slope);
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
    } else {
    // N = 128: both warp groups take every tile, two 32-column chunks each (columns half * 32 + {0, 64}); the statistics
    // of those 64 columns live in per-thread packed fp32 registers (thread = tile row, fixed tile order => reproducible)
    // and are transposed / reduced ONCE, after the last tile.
    static_assert(CH == 32, "N = 128 path works on 32-column chunks");
    constexpr int NCK = N / (2 * CH);                  // chunks per thread: 2
    unsigned long long S1[NCK][CH / 2], S2[NCK][CH / 2];
#pragma unroll
    for (int k = 0; k < NCK; ++k)
#pragma unroll
      for (int j = 0; j < CH / 2; ++j) { S1[k][j] = 0ull; S2[k][j] = 0ull; }
    const unsigned long long ones = pack_f32x2(1.f, 1.f);
    const bool has_slope = P.slope != nullptr;
    const float slope_a = has_slope ? __ldg(P.slope) : 1.f;
    const bool do_stats = P.stats != nullptr;
    int it = 0;
    TileWalk<3> tw;
    { const int radix[3] = {P.tiles_w, P.tiles_h, 1 << 30}; tw.init((int)blockIdx.x, (int)gridDim.x, radix); }
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it, tw.next()) {
      const int twi = tw.d[0], thi = tw.d[1], img = tw.d[2];
      const int oh = thi * HT_H + lh, ow = twi * HT_W + lw;
      const bool valid = oh < P.oh && ow < P.ow;
      bf16* orow = P.out + (long long)img * P.out_sn + (long long)oh * P.out_sh + (long long)ow * P.out_sw;
      const int acc = it % NACC;
      const uint32_t par = (uint32_t)(it / NACC) & 1u;
      mbar_wait(&tfull[acc], par);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * N);
      const bf16* res_row = (P.res && valid) ? P.res + (long long)img * P.res_sn + (long long)oh * P.res_sh + (long long)ow * P.res_sw
                                             : nullptr;
#pragma unroll
      for (int k = 0; k < NCK; ++k) {
        const int c0 = half * CH + k * 2 * CH;
        uint32_t r[32];
        tmem_ld_32x32(t_addr + c0, r);
        tmem_ld_wait();
        if (k == NCK - 1) {   // the accumulator is in registers: hand it back before the arithmetic
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        uint32_t packed[CH / 2];
        if (!has_slope && res_row == nullptr) {      // training forward: bias only (uniform branch)
#pragma unroll
          for (int j = 0; j < CH / 4; ++j) {
            const float4 b = *reinterpret_cast<const float4*>(s_bias + c0 + 4 * j);
            const float2 v0 = unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1])), ones,
                                                     pack_f32x2(b.x, b.y)));
            const float2 v1 = unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])), ones,
                                                     pack_f32x2(b.z, b.w)));
            __nv_bfloat162 h0 = __floats2bfloat162_rn(v0.x, v0.y), h1 = __floats2bfloat162_rn(v1.x, v1.y);
            packed[2 * j] = *reinterpret_cast<uint32_t*>(&h0);
            packed[2 * j + 1] = *reinterpret_cast<uint32_t*>(&h1);
          }
        } else {
#pragma unroll
          for (int j = 0; j < CH / 2; ++j) {
            float f0 = __uint_as_float(r[2 * j]), f1 = __uint_as_float(r[2 * j + 1]);
            f0 += s_bias[c0 + 2 * j]; f1 += s_bias[c0 + 2 * j + 1];
            if (has_slope) { f0 = f0 > 0.f ? f0 : slope_a * f0; f1 = f1 > 0.f ? f1 : slope_a * f1; }
            if (res_row) {
              const uint32_t u = *reinterpret_cast<const uint32_t*>(res_row + c0 + 2 * j);
              f0 += __uint_as_float(u << 16); f1 += __uint_as_float(u & 0xffff0000u);
            }
            __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
            packed[j] = *reinterpret_cast<uint32_t*>(&h);
          }
        }
        if (valid) {
          if (P.wide) {
#pragma unroll
            for (int j = 0; j < CH / 16; ++j)
              st_global_256(orow + c0 + j * 16, packed[8 * j], packed[8 * j + 1], packed[8 * j + 2], packed[8 * j + 3],
                            packed[8 * j + 4], packed[8 * j + 5], packed[8 * j + 6], packed[8 * j + 7]);
          } else {
#pragma unroll
            for (int j = 0; j < CH / 8; ++j)
              *reinterpret_cast<uint4*>(orow + c0 + j * 8) =
                  make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          }
          if (do_stats) {  // statistics of the values as stored
#pragma unroll
            for (int j = 0; j < CH / 2; ++j) {
              const unsigned long long f = pack_f32x2(__uint_as_float(packed[j] << 16), __uint_as_float(packed[j] & 0xffff0000u));
              S1[k][j] = fma_f32x2(f, ones, S1[k][j]);
              S2[k][j] = fma_f32x2(f, f, S2[k][j]);
            }
          }
        }
      }
    }
    if (do_stats) {
#pragma unroll
      for (int k = 0; k < NCK; ++k) {
        const int c0 = half * CH + k * 2 * CH;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (j & 1) ? unpack_f32x2(S1[k][j / 2]).y : unpack_f32x2(S1[k][j / 2]).x;
        const float t1 = warp_transpose_reduce32(v, lane);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (j & 1) ? unpack_f32x2(S2[k][j / 2]).y : unpack_f32x2(S2[k][j / 2]).x;
        const float t2 = warp_transpose_reduce32(v, lane);
        sl[c0 + lane] = t1;
        sl[N + c0 + lane] = t2;
      }
    }
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
      if (P.n_store == 4) {   // one real channel: {sum, sum of squares} live in slots 0 and N
        if (i == 0 || i == N) atomicAdd(&P.stats[i == 0 ? 0 : 1], s);
      } else if (s != 0.0) atomicAdd(&P.stats[i], s);
    }
  }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <int N, int KC>
static int launch_halo(const HaloParams& P, const CUtensorMap& mA, size_t smem, cudaStream_t s) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(halo3x3_kernel<N, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    MPGAN_REQUIRE(e == cudaSuccess, MPGAN_ERR_CUDA, "cudaFuncSetAttribute(halo3x3): %s", cudaGetErrorString(e));
    attr_done = true;
  }
  const int grid = P.total_tiles < num_sms() ? P.total_tiles : num_sms();
  launch_k(halo3x3_kernel<N, KC>, grid, kHaloThreads, smem, s, P, mA);
  MPGAN_CHECK_LAUNCH("halo3x3_kernel");
  return 0;
}

// Stride-1 3x3 convolution through the halo-resident kernel.  dir 0: fprop (w = [N=cy][9][C=cx]); dir 1: data
// gradient (w = transposed shadow [N=cx][9][C=cy], taps flipped).  Returns 1 when the layer is not covered.
static int halo3x3_run(int dir, int n, int ih, int iw, int oh, int ow, int C, int N, int pad, const void* in,
                       int64_t ldi, const void* w, const float* bias, void* out, int64_t ldo, double* stats,
                       const void* res, int64_t ldres, cudaStream_t s, int n_store = 0, const float* slope = nullptr) {
  if (!(C == 16 || C == 32 || C == 64 || C == 128)) return 1;
  if (!(N == 16 || N == 32 || N == 64 || N == 128)) return 1;
  if (ldi % 8 != 0 || ((uintptr_t)in & 15) || (n_store != 4 && ((uintptr_t)w & 15))) return 1;
  if (n_store == 0) {
    if (ldo % 8 != 0 || ((uintptr_t)out & 15)) return 1;
    if (res && (ldres % 8 != 0 || ((uintptr_t)res & 15))) return 1;
  } else if (N != 16 || (n_store == 1 && stats)) {
    return 1;
  }
  const int KC = C < 64 ? C : 64;
  const size_t w_bytes = ((size_t)9 * C * N * 2 + 1023) & ~(size_t)1023;
  const size_t a_bytes = ((size_t)HPIX * KC * 2 + 1023) & ~(size_t)1023;   // one <= 64-channel plane
  const size_t aux = kHaloAux + (size_t)8 * 2 * N * 4 + (size_t)N * 4;   // barriers, stats slots, bias
  const size_t budget = 227 * 1024 - 1024;
  if (w_bytes + 2 * a_bytes + aux > budget) return 1;
  int nbuf = (int)((budget - w_bytes - aux) / a_bytes);
  if (nbuf > kMaxBuf) nbuf = kMaxBuf;
  const int total_tiles_ = n * ((ow + HT_W - 1) / HT_W) * ((oh + HT_H - 1) / HT_H);
  if (N <= 64 && total_tiles_ <= 32 * num_sms()) {   // generator-sized layer: leave room for a second CTA on the SM
    const int cap = (int)(((size_t)100 * 1024 > w_bytes + aux ? (size_t)100 * 1024 - w_bytes - aux : 0) / a_bytes);
    const int floor_ = C > 64 ? 2 : 3;
    if (nbuf > (cap > floor_ ? cap : floor_)) nbuf = cap > floor_ ? cap : floor_;
  }
  HaloParams P;
  memset(&P, 0, sizeof(P));
  P.nimg = n; P.oh = oh; P.ow = ow; P.C = C; P.N = N;
  P.pad = dir == 0 ? pad : 2 - pad;
  P.flip = dir;
  P.tiles_w = (ow + HT_W - 1) / HT_W; P.tiles_h = (oh + HT_H - 1) / HT_H;
  P.total_tiles = n * P.tiles_w * P.tiles_h;
  P.nbuf = nbuf;
  P.w = (const bf16*)w; P.out = (bf16*)out;
  P.out_sn = (long long)oh * ow * ldo; P.out_sh = (long long)ow * ldo; P.out_sw = ldo;
  if (n_store == 4) { P.out_sn = 4LL * oh * ow; P.out_sh = 2LL * ow; P.out_sw = 1; }   // one-channel image of twice the size
  P.bias = bias; P.stats = stats;
  P.wide = (n_store == 0 && ldo % 16 == 0 && ((uintptr_t)out & 31) == 0) ? 1 : 0;
  P.n_store = n_store;
  P.slope = slope;
  P.res = (const bf16*)res;
  P.res_sn = (long long)oh * ow * ldres; P.res_sh = (long long)ow * ldres; P.res_sw = ldres;
  P.idesc = make_idesc_bf16(128, N, 0, 0);
  {
    // two MMA issuers for the small layers (issue / accumulate-dependency bound); one for the MMA-throughput-bound N = 128
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("MPGAN_HALO_ISSUERS"); forced = e ? atoi(e) : 0; }
    P.nissue = (N == 16 && C == 16) ? 2 : 1;   // measured: 16->16 @128^2 17.2 -> 13.6 us; 32 / 64 channels: no gain
    if (forced == 1 || forced == 2) P.nissue = forced;
  }
  CUtensorMap mA;
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)iw, (uint64_t)ih, (uint64_t)n};
    uint64_t str[3] = {(uint64_t)ldi, (uint64_t)ldi * iw, (uint64_t)ldi * iw * ih};
    uint32_t box[4] = {(uint32_t)KC, (uint32_t)HW, (uint32_t)HH, 1u};
    int rc = encode_map(&mA, in, 4, dims, str, box);
    if (rc) return rc;
  }
  const size_t smem = w_bytes + nbuf * a_bytes + aux + 1024;
#define HALO_KC(NN)                                                     \
  switch (KC) {                                                         \
    case 16: return launch_halo<NN, 16>(P, mA, smem, s);                \
    case 32: return launch_halo<NN, 32>(P, mA, smem, s);                \
    default: return launch_halo<NN, 64>(P, mA, smem, s);                \
  }
  switch (N) {
    case 16: HALO_KC(16)
    case 32: HALO_KC(32)
    case 64: HALO_KC(64)
    default: HALO_KC(128)
  }
#undef HALO_KC
}

}  // namespace tc
}  // namespace mpgan
