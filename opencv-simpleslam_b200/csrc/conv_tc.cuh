// conv_tc.cuh - 3x3 convolution (pad 1) as an implicit GEMM on tcgen05/TMEM, fp32-faithful (operands
// carried as NP planes, tc_common.cuh: NP = 2 two fp16 planes / three cross products with a device-side range
// check - the default; NP = 3 three bf16 planes / six cross products).  ALIKED block1.conv2, block2.conv1/2.
//
// Activations live in HBM as "chunk planes": [plane p][8-channel chunk c][H][W][8] bf16, i.e. one pixel of
// one chunk is 16 bytes and pixels of a row are contiguous.  A CTA produces R output rows x 128 pixels:
//  * one 4-D TMA per plane brings the (R+2) x 130 input window of every chunk into shared memory
//    (the halo - and the zero padding at the image border - comes from TMA's out-of-bounds zero fill);
//  * in shared memory a chunk row is a run of 16-byte pixels, which IS the no-swizzle K-major UMMA operand
//    layout (8 consecutive pixels x 16 bytes = one core matrix): the A operand of tap (ky, kx) is the same
//    tile addressed one row / one pixel further - no im2col is ever built;
//  * per output row: 9 taps x CIN/16 k-steps x 6 plane terms MMAs (M = 128 pixels, N = COUT, K = 16
//    channels) into two TMEM accumulators (main a0*w0 term / correction terms, see gemm_tc.cuh);
//  * epilogue (thread == pixel): bias (BN folded), optional residual, SELU, then any of: fp32 CHW map,
//    chunk planes for the next conv, 2x2-average-pooled chunk planes (block2's input).
// Weights: [plane][k-chunk = tap * CIN/8 + c][COUT][8] bf16 (BN scale folded), one bulk copy per CTA.
#pragma once
#include "tc_common.cuh"

namespace b2s {

template <int CIN, int COUT, int NP = 3, int R_ = 4>
struct ConvTcCfg {
  static constexpr int R = R_, TW = 128, PW = TW + 2;      // R output rows per CTA (5 on the fp16x2 path of <32,32>: 128 CTAs = one wave at half resolution)
  static constexpr int NCH = CIN / 8;
  static constexpr int RS = PW * 16;                       // bytes of one input row of one chunk
  static constexpr int CS = (R + 2) * RS;                  // chunk stride
  static constexpr int A_PLANE = NCH * CS;
  static constexpr int W_PLANE = 9 * CIN * COUT * 2;
  static constexpr int ACC = R * 2 * COUT;                 // accumulator columns: (main, correction) per output row
  static constexpr int TMEM_COLS = ACC <= 32 ? 32 : ACC <= 64 ? 64 : ACC <= 128 ? 128 : ACC <= 256 ? 256 : 512;
  static_assert(ACC <= 512, "accumulators exceed the tensor memory");
  static constexpr int SMEM = NP * A_PLANE + NP * W_PLANE + 128 /*align*/ + 128 /*barriers*/ + COUT * 4;
  static_assert(SMEM <= 227 * 1024, "tile does not fit the shared memory");
  static constexpr int THREADS = 192;                      // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
};

struct ConvTcParams {
  int H, W;                         // this layer's resolution (input == output)
  const __nv_bfloat16* wplanes;     // [3][9*CIN/8][COUT][8]
  const float* bias;                // [COUT]
  const float* residual;            // fp32 CHW [COUT][H][W] (nullable), added before the activation
  float* out_chw;                   // fp32 CHW [COUT][H][W] (nullable)
  __nv_bfloat16* out_planes;        // chunk planes [3][COUT/8][H][W][8] (nullable)
  __nv_bfloat16* out_pooled;        // chunk planes of the 2x2 average [3][COUT/8][H/2][W/2][8] (nullable)
  int* range_flag;                  // NP == 2: set to 1 when a value written as fp16 planes leaves the fp16 range
};

// 8 channels of one pixel -> NP planes (16 bytes each) at d + p * plane; NP == 2: returns nonzero when a value overflowed fp16
template <int NP>
__device__ __forceinline__ uint32_t conv_store_planes8(__nv_bfloat16* d, size_t plane, const float* f) {
  uint32_t w[NP][4];
  uint32_t ov = 0u;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t pw[NP];
    tc::pack_planes2<NP>(f[2 * j], f[2 * j + 1], pw);
#pragma unroll
    for (int pl = 0; pl < NP; ++pl) w[pl][j] = pw[pl];
    if (NP == 2) ov |= tc::h2_ovf(pw[0]);
  }
#pragma unroll
  for (int pl = 0; pl < NP; ++pl) *reinterpret_cast<uint4*>(d + pl * plane) = make_uint4(w[pl][0], w[pl][1], w[pl][2], w[pl][3]);
  return ov;
}

template <int CIN, int COUT, int NP = 3, int R_ = 4>
__global__ void __launch_bounds__(192) k_conv3x3_tc(const __grid_constant__ CUtensorMap mapIn, ConvTcParams p) {
  using Cfg = ConvTcCfg<CIN, COUT, NP, R_>;
  using Terms = tc::PlaneTerms<NP>;
  extern __shared__ uint8_t smem_raw[];
  // aligned by pointer ARITHMETIC on the __shared__ array: an integer round trip would make every staging access a generic LD / ST
  uint8_t* smem = smem_raw + ((128u - (tc::smem_u32(smem_raw) & 127u)) & 127u);
  uint8_t* sA = smem;                                       // [NP planes][NCH][R+2][130][8]
  uint8_t* sW = sA + NP * Cfg::A_PLANE;                     // [NP planes][9*NCH][COUT][8]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + NP * Cfg::W_PLANE);
  uint64_t* w_full = bars;
  uint64_t* in_full = bars + 1;
  uint64_t* row_full = bars + 2;                            // [R]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 + Cfg::R);
  float* s_bias = reinterpret_cast<float*>(bars) + 32;      // 128 bytes behind the barriers

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * Cfg::TW, y0 = blockIdx.y * Cfg::R;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&mapIn);
    tc::mbar_init(w_full, 1); tc::mbar_init(in_full, 1);
    for (int r = 0; r < Cfg::R; ++r) tc::mbar_init(&row_full[r], 1);
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  if (threadIdx.x >= 64 && threadIdx.x < 64 + COUT) s_bias[threadIdx.x - 64] = __ldg(p.bias + threadIdx.x - 64);
  pdl_trigger();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (tc::elect_one()) {
      tc::mbar_expect_tx(w_full, NP * Cfg::W_PLANE);         // constant weights: before the dependency wait
      tc::bulk_load(sW, p.wplanes, NP * Cfg::W_PLANE, w_full);
      pdl_wait();                                             // the input map is written by the previous kernel
      tc::mbar_expect_tx(in_full, NP * Cfg::A_PLANE);
#pragma unroll
      for (int pl = 0; pl < NP; ++pl) tc::tma_load_4d(sA + pl * Cfg::A_PLANE, &mapIn, in_full, 0, x0 - 1, y0 - 1, pl * Cfg::NCH);
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::idesc_planes<NP>(128, COUT, 0, 0);
      const uint32_t a_base = tc::smem_u32(sA), w_base = tc::smem_u32(sW);
      tc::mbar_wait(w_full, 0);
      tc::mbar_wait(in_full, 0);
      tc::tc_fence_after();
#pragma unroll 1
      for (int r = 0; r < Cfg::R; ++r) {
        const uint32_t acc = tmem_base + r * 2 * COUT;      // main at acc, corrections at acc + COUT
        uint32_t used = 0;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t a_tap = a_base + ((r + tap / 3) * Cfg::PW + tap % 3) * 16;
#pragma unroll
          for (int s = 0; s < CIN / 16; ++s) {
#pragma unroll
            for (int t = 0; t < Terms::N; ++t) {
              const uint64_t ad = tc::smem_desc_nosw(a_tap + Terms::a(t) * Cfg::A_PLANE + 2 * s * Cfg::CS, Cfg::CS, 128);
              const uint64_t bd = tc::smem_desc_nosw(w_base + Terms::b(t) * Cfg::W_PLANE + (tap * Cfg::NCH + 2 * s) * COUT * 16,
                                                     COUT * 16, 128);
              const int which = t == Terms::N - 1 ? 0 : 1;
              tc::umma_bf16(acc + which * COUT, ad, bd, idesc, (used >> which) & 1u);
              used |= 1u << which;
            }
          }
        }
        tc::umma_commit(&row_full[r]);
      }
    }
  } else {
    // ===== epilogue: thread == output pixel x0 + px of every row =====
    pdl_wait();                                               // residual / output buffers belong to the stream order
    const int quad = warp & 3;
    const int px = quad * 32 + lane, gx = x0 + px;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const size_t HW = (size_t)p.H * p.W;
    float prev[COUT];                                         // previous (even) row, kept for the 2x2 pooling
    uint32_t ovf = 0u;
#pragma unroll 1
    for (int r = 0; r < Cfg::R; ++r) {
      tc::mbar_wait(&row_full[r], 0);
      tc::tc_fence_after();
      float f[COUT];
#pragma unroll
      for (int c0 = 0; c0 < COUT; c0 += 16) {
        uint32_t a[16], b[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]),
                       "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15])
                     : "r"(tmem_base + lane_addr + r * 2 * COUT + c0) : "memory");
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]), "=r"(b[8]),
                       "=r"(b[9]), "=r"(b[10]), "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15])
                     : "r"(tmem_base + lane_addr + r * 2 * COUT + COUT + c0) : "memory");
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) f[c0 + j] = (NP == 2 ? fmaf(__uint_as_float(b[j]), Terms::CORR, __uint_as_float(a[j])) : __uint_as_float(a[j]) + __uint_as_float(b[j])) + s_bias[c0 + j];
      }
      const int gy = y0 + r;
      const bool inside = gy < p.H && gx < p.W;
      const size_t pix = (size_t)gy * p.W + gx;
      if (p.residual && inside) {
#pragma unroll
        for (int c = 0; c < COUT; ++c) f[c] += p.residual[(size_t)c * HW + pix];
      }
#pragma unroll
      for (int c = 0; c < COUT; ++c) f[c] = selu_f(f[c]);
      if (inside) {
        if (p.out_chw) {
#pragma unroll
          for (int c = 0; c < COUT; ++c) p.out_chw[(size_t)c * HW + pix] = f[c];
        }
        if (p.out_planes) {
#pragma unroll
          for (int ch = 0; ch < COUT / 8; ++ch)
            ovf |= conv_store_planes8<NP>(p.out_planes + ((size_t)ch * HW + pix) * 8, (size_t)(COUT / 8) * HW * 8, f + 8 * ch);
        }
      }
      if (p.out_pooled) {                                     // uniform branch; all lanes shuffle
        if ((r & 1) == 0) {
#pragma unroll
          for (int c = 0; c < COUT; ++c) prev[c] = f[c];
        } else {
          // upstream avg_pool2d order: (((top-left + top-right) + bottom-left) + bottom-right) * 0.25
          float q[COUT];
#pragma unroll
          for (int c = 0; c < COUT; ++c) {
            const float tr = __shfl_down_sync(0xffffffffu, prev[c], 1), br = __shfl_down_sync(0xffffffffu, f[c], 1);
            q[c] = (((prev[c] + tr) + f[c]) + br) * 0.25f;
          }
          const int H2 = p.H >> 1, W2 = p.W >> 1;
          const int qx = gx >> 1, qy = gy >> 1;
          if ((lane & 1) == 0 && qy < H2 && qx < W2) {
            const size_t HW2 = (size_t)H2 * W2, qpix = (size_t)qy * W2 + qx;
#pragma unroll
            for (int ch = 0; ch < COUT / 8; ++ch)
              ovf |= conv_store_planes8<NP>(p.out_pooled + ((size_t)ch * HW2 + qpix) * 8, (size_t)(COUT / 8) * HW2 * 8, q + 8 * ch);
          }
        }
      }
    }
    if (NP == 2 && ovf && p.range_flag) *p.range_flag = 1;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace b2s
