// sm_100a primitives used by the tcgen05 convolution kernels: mbarrier, TMA (cp.async.bulk.tensor), TMEM
// alloc/ld, tcgen05.mma / commit, UMMA shared-memory and instruction descriptors.  Inline PTX only.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpgan {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t.reg .b32 R;\n\t"
      "elect.sync R|P1, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- TMA loads (tile mode), completing on an mbarrier ----
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
          "r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ---- TMEM ----
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 columns of fp32: thread t of the warp receives lane (base_lane + t), columns [col, col+32)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- UMMA ----
// Shared-memory matrix descriptor (sm_100): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48)
// | layout [61,64): 0 none, 2 SW128, 4 SW64, 6 SW32.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, M x N, majors (0 = K-major, 1 = MN-major)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}


// Mixed-radix counter over a persistent CTA's tile sequence tile0, tile0 + step, ...  (digit 0 = fastest index, the last
// digit is unbounded).  The divisions happen once, in init(); next() is ND adds with carry -- a per-tile decode with
// 3-10 runtime divisions (~40 cycles each) on a single producer / epilogue thread was a measurable part of the
// 7-28-tile generator kernels.
template <int ND> struct TileWalk {
  int d[ND], s[ND], r[ND];
  __device__ __forceinline__ void init(int tile0, int step, const int (&radix)[ND]) {
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      r[k] = radix[k];
      if (k < ND - 1) {
        d[k] = tile0 % r[k]; tile0 /= r[k];
        s[k] = step % r[k]; step /= r[k];
      } else {
        d[k] = tile0; s[k] = step;
      }
    }
  }
  __device__ __forceinline__ void next() {
    int carry = 0;
#pragma unroll
    for (int k = 0; k < ND - 1; ++k) {
      const int v = d[k] + s[k] + carry;
      carry = v >= r[k] ? 1 : 0;
      d[k] = v - (carry ? r[k] : 0);
    }
    d[ND - 1] += s[ND - 1] + carry;
  }
};

// ---- small-N epilogue arithmetic (shared by the halo, tap-GEMM and one-channel kernels) ----
// Blackwell issues fp32 FMAs at full rate in the packed form (two lanes of a 64-bit register pair).
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long fma_f32x2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float2 unpack_f32x2(unsigned long long v) {
  float2 f;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(f.x), "=f"(f.y) : "l"(v));
  return f;
}

// 32-byte store (sm_100 STG.256): one full sector per thread instead of two half-sector writes
__device__ __forceinline__ void st_global_256(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e,
                                              uint32_t f, uint32_t g, uint32_t h) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e),
               "r"(f), "r"(g), "r"(h)
               : "memory");
}

// One CH-column chunk of one accumulator row: + bias -> bf16 -> 16-byte stores at `orow`, and (optionally) the
// BatchNorm statistics of the values AS STORED, accumulated in per-thread packed registers s1 / s2 (CH/2 pairs each).
// `rrow` (optional): a bf16 row of the same shape added before rounding (fused residual / gradient accumulation).
// `slope` (optional, device scalar): PReLU / LeakyReLU applied to (acc + bias) BEFORE the residual is added -- the
// inference-mode fusion conv -> folded BatchNorm -> PReLU (+ residual) of MONAI's Convolution / ResidualUnit.
template <int CH>
__device__ __forceinline__ void load_res_row(const bf16* rrow, uint32_t (&rq)[CH / 2]) {
#pragma unroll
  for (int j = 0; j < CH / 8; ++j) {
    const uint4 q = *reinterpret_cast<const uint4*>(rrow + j * 8);
    rq[4 * j] = q.x; rq[4 * j + 1] = q.y; rq[4 * j + 2] = q.z; rq[4 * j + 3] = q.w;
  }
}

// core of the chunk epilogue with the residual row already in registers (`rq`, read only when has_res): callers that
// know the output position before the accumulator is ready request the row early, so that its latency overlaps the MMAs
template <int CH>
__device__ __forceinline__ void epi_chunk_store_rq(const uint32_t (&r)[CH], const float* s_bias_c0, bf16* orow, bool valid,
                                                   bool do_stats, unsigned long long (&s1)[CH / 2],
                                                   unsigned long long (&s2)[CH / 2], const uint32_t (&rq)[CH / 2],
                                                   bool has_res, bool wide = false, const float* slope = nullptr) {
  // All flags are warp-uniform.  The residual is folded in without per-column branches (a missing residual is a row of
  // zeros) and the activation is ONE uniform branch around two straight-line loops: the previous form (a branch per
  // group of four columns) made this epilogue ~250 dependent instructions per 16-column tile, which bounded the
  // 16-channel layers (ncu source view: the MMA issuer spinning on the accumulator-empty barrier).
  const unsigned long long ones = pack_f32x2(1.f, 1.f);
  uint32_t packed[CH / 2];
  uint32_t q[CH / 2];
#pragma unroll
  for (int j = 0; j < CH / 2; ++j) q[j] = has_res ? rq[j] : 0u;
  if (slope != nullptr) {
    const float a = __ldg(slope);
#pragma unroll
    for (int j = 0; j < CH / 4; ++j) {
      const float4 b = *reinterpret_cast<const float4*>(s_bias_c0 + 4 * j);
      float2 v0 = unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1])), ones,
                                         pack_f32x2(b.x, b.y)));
      float2 v1 = unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])), ones,
                                         pack_f32x2(b.z, b.w)));
      v0.x = (v0.x > 0.f ? v0.x : a * v0.x) + __uint_as_float(q[2 * j] << 16);
      v0.y = (v0.y > 0.f ? v0.y : a * v0.y) + __uint_as_float(q[2 * j] & 0xffff0000u);
      v1.x = (v1.x > 0.f ? v1.x : a * v1.x) + __uint_as_float(q[2 * j + 1] << 16);
      v1.y = (v1.y > 0.f ? v1.y : a * v1.y) + __uint_as_float(q[2 * j + 1] & 0xffff0000u);
      __nv_bfloat162 h0 = __floats2bfloat162_rn(v0.x, v0.y);
      __nv_bfloat162 h1 = __floats2bfloat162_rn(v1.x, v1.y);
      packed[2 * j] = *reinterpret_cast<uint32_t*>(&h0);
      packed[2 * j + 1] = *reinterpret_cast<uint32_t*>(&h1);
    }
  } else {
#pragma unroll
    for (int j = 0; j < CH / 4; ++j) {
      float4 b = *reinterpret_cast<const float4*>(s_bias_c0 + 4 * j);
      b.x += __uint_as_float(q[2 * j] << 16); b.y += __uint_as_float(q[2 * j] & 0xffff0000u);
      b.z += __uint_as_float(q[2 * j + 1] << 16); b.w += __uint_as_float(q[2 * j + 1] & 0xffff0000u);
      const float2 v0 = unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1])), ones,
                                               pack_f32x2(b.x, b.y)));
      const float2 v1 = unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])), ones,
                                               pack_f32x2(b.z, b.w)));
      __nv_bfloat162 h0 = __floats2bfloat162_rn(v0.x, v0.y);
      __nv_bfloat162 h1 = __floats2bfloat162_rn(v1.x, v1.y);
      packed[2 * j] = *reinterpret_cast<uint32_t*>(&h0);
      packed[2 * j + 1] = *reinterpret_cast<uint32_t*>(&h1);
    }
  }
  if (valid) {
    if (wide) {   // row address 32-byte aligned
#pragma unroll
      for (int j = 0; j < CH / 16; ++j)
        st_global_256(orow + j * 16, packed[8 * j], packed[8 * j + 1], packed[8 * j + 2], packed[8 * j + 3], packed[8 * j + 4],
                      packed[8 * j + 5], packed[8 * j + 6], packed[8 * j + 7]);
    } else {
#pragma unroll
      for (int j = 0; j < CH / 8; ++j)
        *reinterpret_cast<uint4*>(orow + j * 8) =
            make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
    }
    if (do_stats) {
#pragma unroll
      for (int j = 0; j < CH / 2; ++j) {
        const unsigned long long f =
            pack_f32x2(__uint_as_float(packed[j] << 16), __uint_as_float(packed[j] & 0xffff0000u));
        s1[j] = fma_f32x2(f, ones, s1[j]);
        s2[j] = fma_f32x2(f, f, s2[j]);
      }
    }
  }
}

template <int CH>
__device__ __forceinline__ void epi_chunk_store(const uint32_t (&r)[CH], const float* s_bias_c0, bf16* orow, bool valid,
                                                bool do_stats, unsigned long long (&s1)[CH / 2],
                                                unsigned long long (&s2)[CH / 2], const bf16* rrow = nullptr,
                                                bool wide = false, const float* slope = nullptr) {
  uint32_t rq[CH / 2];
  const bool has_res = rrow != nullptr && valid;
  if (has_res) load_res_row<CH>(rrow, rq);
  epi_chunk_store_rq<CH>(r, s_bias_c0, orow, valid, do_stats, s1, s2, rq, has_res, wide, slope);
}

}  // namespace tc
}  // namespace mpgan
