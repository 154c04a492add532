// tc_common.cuh - thin inline-PTX layer for the sm_100a tensor-core path: mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld), UMMA descriptors.
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix/instruction descriptor" tables
// (cross-checked against CuTe's cute/arch/mma_sm100_desc.hpp in the image).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace b2s {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}

// ---- TMA --------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tile load: coordinates (c0 = innermost element index, c1 = row index)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// same, delivered to the same CTA-relative smem offset (and signalling the mbarrier at the same offset) of every CTA
// of the cluster whose bit is set in cta_mask
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}

// ---- thread-block clusters --------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// all threads of all CTAs of the cluster (also a CTA-wide barrier)
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// 4-D tile load (c0 innermost); out-of-bounds elements (negative or past the extent) arrive as zeros
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// contiguous global -> shared bulk copy (bytes multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -------------------------------------------------------------------------------
// whole warp; writes the TMEM base address (lane<<16 | column) to *smem_slot
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// same with the A operand in tensor memory (lane = row, one 32-bit column = two consecutive K elements)
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// same, arriving on the mbarrier at this CTA-relative offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp receives row (lane base + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM, same shape as tmem_ld32
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
        "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
        "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors -------------------------------------------------------------------
// shared-memory matrix descriptor, 128-byte swizzle, tiles 1024-byte aligned.
//   bits [0,14)  start address >> 4      bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4 bits [46,48) version = 1 (sm_100)   bits [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
// same, no swizzle ("interleave"): 8-row x 16-byte core matrices stored contiguously (128 B);
// sbo = bytes between consecutive 8-row groups, lbo = bytes between the two 16-byte K chunks of one MMA
__device__ __forceinline__ uint64_t smem_desc_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor for kind::f16 with bf16 A/B and fp32 D.
//   [4,6) D fmt (1 = f32)  [7,10) A fmt (1 = bf16)  [10,13) B fmt  [15] A major  [16] B major (1 = MN-major)
//   [17,23) N >> 3         [24,29) M >> 4
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// same with fp16 A/B (format code 0)
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// instruction descriptor of the operand-plane scheme NP (1, 3: bf16 planes; 2: fp16 planes)
template <int NP>
__host__ __device__ constexpr uint32_t idesc_planes(int M, int N, int a_mn_major, int b_mn_major) {
  return NP == 2 ? idesc_f16(M, N, a_mn_major, b_mn_major) : idesc_bf16(M, N, a_mn_major, b_mn_major);
}

// ---- fp32 carried as bf16 planes --------------------------------------------------------------
// x = a0 + a1 + a2 exactly up to 24 significant bits (each a_i a bf16, a_{i+1} = bf16(residual)):
// with six cross products a_i b_j (i + j <= 2) accumulated in fp32 the tensor core reproduces an
// fp32 dot product to fp32 rounding level.  pack_planes2 splits two values and returns, per plane,
// the packed bf16x2 word (first value in the low half).
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// ---- fp32 carried as TWO fp16 planes ("h2", NP == 2) --------------------------------------------
// x = h0 + h1 * 2^-11 with h0 = fp16(x) and h1 = fp16((x - h0) * 2^11): 22 significant bits whatever the magnitude
// of x (the scaled residual is as far from the fp16 subnormal range as x itself), and a product x*w needs only the
// THREE cross terms h0*w0 (main) and h0*w1 + h1*w0 (corrections, accumulated separately and weighted 2^-11 by the
// epilogue; the dropped h1*w1 is 2^-22 of the product) - half the tensor-core work of the bf16x3 scheme at an
// operand error (2^-23 rms) that stays below the fp32 rounding noise of the accumulation itself.  The price is
// fp16's exponent range: |x| must stay below 65504.  Every producer of h2 planes therefore checks the packed words
// for infinities (h2_ovf) and raises a sticky device flag; the matcher re-runs a flagged batch on the bf16x3 path.
// Attention operands (q, k, v, P: a contraction there has ONE accumulator) use the unscaled variant
// h1 = fp16(s*x - h0), h0 = fp16(s*x) with a fixed power-of-two prescale s (H2_ATTN_PRESCALE; P: 2^14, P <= 1).
constexpr float H2_RS = 2048.f;                 // residual scale 2^11
constexpr float H2_IRS = 1.f / 2048.f;
constexpr float H2_ATTN_PRESCALE = 16.f;        // q, k, v are stored as 16 x
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }
// nonzero iff one of the two fp16 halves of w is +-inf or nan
__device__ __forceinline__ uint32_t h2_ovf(uint32_t w) { return ((w & 0x7FFF7FFFu) + 0x04000400u) & 0x80008000u; }
// (a - lo(w), b - hi(w)) for a packed fp16 pair w: two mixed-precision adds (FHADD with a negated .H0 / .H1 operand) - no
// unpacking; exact when w = fp16x2(a, b)
__device__ __forceinline__ void sub_f16x2(float a, float b, uint32_t w, float& ra, float& rb) {
  asm("{\n\t.reg .b16 lo, hi;\n\t.reg .b32 nw;\n\tneg.f16x2 nw, %2;\n\tmov.b32 {lo, hi}, nw;\n\t"
      "add.rn.f32.f16 %0, lo, %3;\n\tadd.rn.f32.f16 %1, hi, %4;\n\t}" : "=f"(ra), "=f"(rb) : "r"(w), "f"(a), "f"(b));
}
// GEMM operand format (scaled residual)
__device__ __forceinline__ void pack_h2(float a, float b, uint32_t& w0, uint32_t& w1) {
  w0 = pack_f16x2(a, b);
  float ra, rb;
  sub_f16x2(a, b, w0, ra, rb);
  w1 = pack_f16x2(ra * H2_RS, rb * H2_RS);
}
// attention operand format (unscaled residual) of values that already carry their prescale
__device__ __forceinline__ void pack_h2_raw(float a, float b, uint32_t& w0, uint32_t& w1) {
  w0 = pack_f16x2(a, b);
  float ra, rb;
  sub_f16x2(a, b, w0, ra, rb);
  w1 = pack_f16x2(ra, rb);
}
// attention operand format (prescaled by s, unscaled residual)
__device__ __forceinline__ void pack_h2_attn(float a, float b, float s, uint32_t& w0, uint32_t& w1) { pack_h2_raw(a * s, b * s, w0, w1); }
template <int NP>
__device__ __forceinline__ void pack_planes2(float a, float b, uint32_t (&w)[NP]) {
  if (NP == 2) { pack_h2(a, b, w[0], w[NP - 1]); return; }
  w[0] = pack_bf16x2(a, b);
  if (NP > 1) {
    a -= __uint_as_float(w[0] << 16); b -= __uint_as_float(w[0] & 0xFFFF0000u);
    w[1] = pack_bf16x2(a, b);
  }
  if (NP > 2) {
    a -= __uint_as_float(w[1] << 16); b -= __uint_as_float(w[1] & 0xFFFF0000u);
    w[2] = pack_bf16x2(a, b);
  }
}
}  // namespace tc
// store 8 consecutive values of one row as NP operand planes (plane p at y + p * plane); NP == 2: returns nonzero when a
// value left the fp16 range (the caller raises the matcher's range flag)
template <int NP>
__device__ __forceinline__ uint32_t store_planes8(__nv_bfloat16* y, size_t plane, const float (&f)[8]) {
  uint32_t w[NP][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t pw[NP];
    tc::pack_planes2<NP>(f[2 * j], f[2 * j + 1], pw);
#pragma unroll
    for (int pl = 0; pl < NP; ++pl) w[pl][j] = pw[pl];
  }
#pragma unroll
  for (int pl = 0; pl < NP; ++pl) *reinterpret_cast<uint4*>(y + pl * plane) = make_uint4(w[pl][0], w[pl][1], w[pl][2], w[pl][3]);
  if (NP != 2) return 0u;
  return tc::h2_ovf(w[0][0]) | tc::h2_ovf(w[0][1]) | tc::h2_ovf(w[0][2]) | tc::h2_ovf(w[0][3]);
}

namespace tc {

// operand-plane pairs (a_i, b_j) of the split product, smallest contributions first
template <int NP> struct PlaneTerms {
  static constexpr int N = NP == 1 ? 1 : NP == 2 ? 3 : 6;
  // NP == 3: (2,0) (1,1) (0,2) (1,0) (0,1) (0,0)      NP == 2: (1,0) (0,1) (0,0)
  __host__ __device__ static constexpr int a(int t) { return NP == 1 ? 0 : NP == 2 ? (t == 0 ? 1 : 0) : (t == 0 ? 2 : (t == 1 || t == 3) ? 1 : 0); }
  __host__ __device__ static constexpr int b(int t) { return NP == 1 ? 0 : NP == 2 ? (t == 1 ? 1 : 0) : (t == 2 ? 2 : (t == 1 || t == 4) ? 1 : 0); }
  // weight of the correction accumulator (every term but the last) in the epilogue's sum
  static constexpr float CORR = NP == 2 ? 1.f / 2048.f : 1.f;
};

}  // namespace tc

// host: 2-D bf16 tensor map, 128-byte swizzle. inner = contiguous extent (elements), box_inner must be 64.
int make_tmap_bf16_2d(CUtensorMap* out, const void* gptr, uint64_t inner, uint64_t rows, uint64_t row_pitch_bytes,
                      uint32_t box_inner, uint32_t box_rows);
// host: 4-D bf16 tensor map without swizzle. dims[0] is the contiguous extent; strides_bytes[i] is the pitch of dims[i+1].
int make_tmap_bf16_4d(CUtensorMap* out, const void* gptr, const uint64_t dims[4], const uint64_t strides_bytes[3],
                      const uint32_t box[4]);

}  // namespace b2s
