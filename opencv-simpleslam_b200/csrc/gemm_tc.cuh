// gemm_tc.cuh - tcgen05/TMEM GEMM fed by TMA, fp32 accumulation, fused epilogues.
//   C[M,N] = epi( [A1 | A2][M,K] * W[N,K]^T + bias )      (both operands K-major, 128B swizzle)
// Operands are bf16 PLANES (tc_common.cuh): NP = 1 is the plain bf16 path; NP = 3 carries fp32
// values as three bf16 planes and issues the six cross products a_i w_j (i + j <= 2) per k-step,
// which reproduces the fp32 product to fp32 rounding level on the tensor pipe ("fp32 on bf16x3").
// Plane p of an activation buffer lives `plane_rows` rows below plane 0 (same tensor map);
// plane p of a weight lives K columns to the right ([N, NP*K]).
// One 128 x BN output tile per CTA, BK = 64 (one 128-byte swizzle atom per row), TMA->smem ring
// (a stage holds all NP planes of both operands), warp-specialised: warp 0 = TMA producer,
// warp 1 = MMA issuer (one elected lane issues tcgen05.mma, accumulator in TMEM), warps 2..9 =
// epilogue (tcgen05.ld: thread == accumulator row, so row-wise epilogues - rotary pairs, residual,
// plane splitting - are thread-local; two warps per TMEM lane quadrant alternate over the 32-column
// chunks and write through a swizzled staging tile so that global stores are row-contiguous).
// Row space: SEGMENTS of seg_stride rows (lightglue_kernels.cuh: segment 2p + s = image s of pair p of a batch); grid.y
// enumerates (segment, 128-row tile) and the live rows of a segment come from the pair's device-resident state.
#pragma once
#include "tc_common.cuh"

namespace b2s {

// epilogues: planes only | rotary + planes | fp32 only | in-place fp32 residual stream + planes | fp32 + planes
enum TcEpi { TC_EPI_BF16 = 0, TC_EPI_ROTARY_BF16 = 1, TC_EPI_F32 = 2, TC_EPI_RESID_F32_BF16 = 3, TC_EPI_F32_BF16 = 4 };

constexpr int TC_CTRL_INTS = 32;    // == LGC_INTS: ints of device-resident state per pair

struct TcGemmParams {
  int K, K1, N;                     // K total, K1 = columns served by map A1 (K1 == K when single source)
  int seg_stride;                   // rows between segment bases (0: a single segment)
  int tiles_per_seg;                // grid.y = segments * tiles_per_seg
  int seg_rows;                     // static live rows of every segment (used when neither ctrl nor m_dev is set)
  int plane_rows;                   // rows between consecutive operand planes of A1 / A2 (NP > 1)
  const float* bias;                // [N]
  int epi;
  __nv_bfloat16* out_bf16; int ld_bf16;     // TC_EPI_BF16 / ROTARY / RESID (bf16 copy); plane p at + p * out_plane
  size_t out_plane;                         // elements between output planes
  float* out_f32; int ld_f32;               // TC_EPI_F32 (plain) / RESID (in-place residual stream)
  const float* rot_cos; const float* rot_sin; int rot_cols;   // [rows,32] tables
  const int* ctrl;                  // LightGlue device state (nullable), TC_CTRL_INTS per pair: live rows of segment g =
                                    //   ctrl[(g >> 1) * 32 + 2 + (g & 1)]; the pair's tiles exit when it has stopped
  int ctrl_mode;                    // 0/1 transformer layer (skipped after an early exit); 2 assignment head: runs after a stop
                                    //   too, weight rows / bias of layer ctrl[6] (w_layer_rows apart); 3 similarity of pair p =
                                    //   grid segment p: A = segment 2p (M = ctrl[2]), B = the activation rows of segment 2p + 1
                                    //   (N = ctrl[3], planes w_plane_rows apart), outputs out_pair_stride floats apart
  int w_layer_rows, w_plane_rows, w_row0;
  size_t out_pair_stride;
  int out_planes;                   // planes written by the bf16 epilogues (0 = NP; 1 lets a three-plane GEMM feed a bf16 layer)
  float alpha;                      // epilogue: (acc + bias) * alpha (0 means 1)
  const int* m_dev; int m_mult;     // nullable: live rows of segment 0 = *m_dev * m_mult (ALIKED keypoint count)
  int act;                          // 0 none, 1 SELU (after bias / residual)
  const float* residual; int ld_res; // nullable fp32 [rows, ld_res]: added after bias * alpha, before the activation
  float* out_f32_t; int ld_f32_t;    // TC_EPI_F32, nullable: also write the transposed result out_f32_t[col * ld_f32_t + row]
  unsigned long long* ts;            // nullable profiling hook: per CTA 6 globaltimer stamps (b2s_bench_gemm_tc3)
  int attn_fmt;                      // NP == 2: output planes in the attention operand format (tc_common.cuh) instead of the GEMM one
  int* range_flag;                   // NP == 2, nullable: set to 1 when an output value leaves the fp16 range
};

// fp32 weight [N,K] -> np operand planes side by side: out[n][p*K + k] (np = 2: fp16 pair, scaled residual; *bad is set
// when a weight is outside the fp16 range)
static __global__ void k_weight_planes(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int N, int K, int np, int* bad = nullptr) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * K) return;
  const int n = (int)(i / K), k = (int)(i % K);
  float r = w[i];
  if (np == 2) {
    const __half h0 = __float2half_rn(r);
    const __half h1 = __float2half_rn((r - __half2float(h0)) * tc::H2_RS);
    __half* o = reinterpret_cast<__half*>(out);
    o[(size_t)n * 2 * K + k] = h0;
    o[(size_t)n * 2 * K + K + k] = h1;
    if (bad && !(fabsf(r) <= 65504.f)) *bad = 1;
    return;
  }
  for (int p = 0; p < np; ++p) {
    const __nv_bfloat16 b = __float2bfloat16_rn(r);
    out[(size_t)n * np * K + (size_t)p * K + k] = b;
    r -= __bfloat162float(b);
  }
}

template <int BN, int NP>
struct TcGemmCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;       // one plane
  static constexpr int STAGE_BYTES = NP * (A_BYTES + B_BYTES);
  static constexpr int STAGES = NP == 1 ? 3 : (BN >= 96 ? 2 : 3);
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + BN * 4 /*bias*/;
  // warps 0 / 1 = TMA / MMA; EPI_WARPS epilogue warps: four per TMEM lane quadrant, each owning every fourth 32-column
  // chunk (one warp per scheduler cannot hide its own instruction latency: the thread==row epilogue is latency bound)
  static constexpr int EPI_WARPS = 8;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  // The tensor core truncates its fp32 accumulator after every k-step (measured: bias -4.4e-9 * K on positive
  // data), which would cost the NP = 3 path its fp32 fidelity for long K.  So the accumulation is spread over
  // NACC TMEM accumulators that the epilogue adds in fp32 registers (round-to-nearest): accumulator 0 takes the
  // five correction terms (2^-8 of the magnitude, their truncation is negligible), accumulators 1.. take the
  // main a0*w0 term round-robin by k-block.
  static constexpr int NACC = NP == 1 ? 1 : 512 / BN;
  static constexpr int TMEM_COLS = NACC * BN <= 64 ? 64 : NACC * BN <= 128 ? 128 : NACC * BN <= 256 ? 256 : 512;   // power of two
};

// (A cluster-multicast variant sharing the A tile between the column tiles of a row tile was measured in round 1 - no
// gain: the main loop is bound by shared-memory bandwidth, not by the L2 -> SM fill - and removed.)
template <int BN, int NP>
__global__ void __launch_bounds__(TcGemmCfg<BN, NP>::THREADS) k_gemm_tc(const __grid_constant__ CUtensorMap mapA1,
                                                 const __grid_constant__ CUtensorMap mapA2,
                                                 const __grid_constant__ CUtensorMap mapW, TcGemmParams p) {
  using Cfg = TcGemmCfg<BN, NP>;
  using Terms = tc::PlaneTerms<NP>;
  extern __shared__ uint8_t smem_raw[];
  // aligned by pointer ARITHMETIC on the __shared__ array: an integer round trip would make every staging access a generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  // stage s: A planes at s*STAGE_BYTES + p*A_BYTES, W planes behind them
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = bars + Cfg::STAGES;  // [STAGES]
  uint64_t* tmem_full = bars + 2 * Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 1);
  float* s_bias = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_m = blockIdx.y, n0 = blockIdx.x * BN;
  auto stamp = [&](int slot) {
    if (p.ts) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); p.ts[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 6 + slot] = t; }
  };
  if (threadIdx.x == 0) stamp(0);                    // CTA start
  const int gseg = tile_m / p.tiles_per_seg, tile_s = tile_m - gseg * p.tiles_per_seg;   // grid segment, tile inside it
  const int seg = p.ctrl_mode == 3 ? 2 * gseg : gseg;                                  // row segment of the A operand
  const int row0 = seg * p.seg_stride + tile_s * Cfg::BM;                              // global row of the tile
  int rows_live = p.seg_rows - tile_s * Cfg::BM;                                      // live rows in this tile
  const int nkb = p.K / Cfg::BK;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&mapA1); tc::tma_prefetch_desc(&mapA2); tc::tma_prefetch_desc(&mapW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(tmem_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  if (warp >= 2) {                       // bias is a constant weight: staged before the dependency wait
    const int t = threadIdx.x - 64;
    if (t < BN) s_bias[t] = (p.bias && n0 + t < p.N) ? __ldg(p.bias + n0 + t) : 0.f;
  }
  pdl_trigger();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();            // everything above overlaps the previous kernel's tail; operands are touched only below
  if (threadIdx.x == 0) stamp(1);                    // dependencies resolved
  int w_row = p.w_row0, n_live = p.N;
  float* out_f32 = p.out_f32; float* out_f32_t = p.out_f32_t;
  if (p.ctrl) {          // device-resident sizes: pruning shrinks the segments, an early exit empties the layer
    const int* c = p.ctrl + (seg >> 1) * TC_CTRL_INTS;
    const bool sides = c[2] > 0 && c[3] > 0;
    const bool on = sides && (p.ctrl_mode >= 2 || !c[1]);
    rows_live = on ? c[2 + (seg & 1)] - tile_s * Cfg::BM : 0;
    if (p.ctrl_mode == 2) {
      w_row += c[6] * p.w_layer_rows;
      if (warp >= 2) {               // the bias depends on the layer: restage it (epilogue warps only)
        const int t = threadIdx.x - 64;
        if (t < BN) s_bias[t] = __ldg(p.bias + w_row + n0 + t);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * Cfg::EPI_WARPS) : "memory");
      }
    } else if (p.ctrl_mode == 3) {
      n_live = c[3];
      if (n0 >= n_live) rows_live = 0;
      w_row = (seg + 1) * p.seg_stride;
      out_f32 += gseg * p.out_pair_stride;
      if (out_f32_t) out_f32_t += gseg * p.out_pair_stride;
    }
  }
  if (p.m_dev) rows_live = *p.m_dev * p.m_mult - tile_m * Cfg::BM;

  if (rows_live <= 0) {
    // nothing to do for this tile (uniform per CTA)
  } else if (warp == 0) {
    // ===== TMA producer =====
    if (tc::elect_one()) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % Cfg::STAGES, ph = (kb / Cfg::STAGES) & 1;
        tc::mbar_wait(&empty[s], ph ^ 1);
        tc::mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
        uint8_t* st = smem + s * Cfg::STAGE_BYTES;
        const int k0 = kb * Cfg::BK;
#pragma unroll
        for (int pl = 0; pl < NP; ++pl) {
          if (k0 < p.K1) tc::tma_load_2d(st + pl * Cfg::A_BYTES, &mapA1, &full[s], k0, row0 + pl * p.plane_rows);
          else tc::tma_load_2d(st + pl * Cfg::A_BYTES, &mapA2, &full[s], k0 - p.K1, row0 + pl * p.plane_rows);
          if (p.w_plane_rows) tc::tma_load_2d(st + NP * Cfg::A_BYTES + pl * Cfg::B_BYTES, &mapW, &full[s], k0, w_row + n0 + pl * p.w_plane_rows);
          else tc::tma_load_2d(st + NP * Cfg::A_BYTES + pl * Cfg::B_BYTES, &mapW, &full[s], pl * p.K + k0, w_row + n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::idesc_planes<NP>(128, BN, 0, 0);
      uint32_t used = 0;                                  // accumulators already written (first MMA overwrites)
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % Cfg::STAGES, ph = (kb / Cfg::STAGES) & 1;
        tc::mbar_wait(&full[s], ph);
        tc::tc_fence_after();
        if (kb == 0) stamp(2);                       // first operand stage landed
        const uint32_t a0 = tc::smem_u32(smem + s * Cfg::STAGE_BYTES), b0 = a0 + NP * Cfg::A_BYTES;
#pragma unroll
        for (int t = 0; t < Terms::N; ++t) {
          const int acc = (Cfg::NACC == 1) ? 0 : (t == Terms::N - 1 ? 1 + kb % (Cfg::NACC - 1) : 0);
#pragma unroll
          for (int k = 0; k < Cfg::BK / 16; ++k) {
            const uint64_t ad = tc::smem_desc_sw128(a0 + Terms::a(t) * Cfg::A_BYTES + k * 32, 16, 1024);
            const uint64_t bd = tc::smem_desc_sw128(b0 + Terms::b(t) * Cfg::B_BYTES + k * 32, 16, 1024);
            tc::umma_bf16(tmem_base + acc * BN, ad, bd, idesc, (used >> acc) & 1u);
            used |= 1u << acc;
          }
        }
        tc::umma_commit(&empty[s]);                          // frees the stage when these MMAs retire
        if (kb == nkb - 1) tc::umma_commit(tmem_full);
      }
    }
  } else {
    // ===== epilogue: a warp may touch the TMEM lane quadrant (warp % 4); the column chunks are dealt round-robin to the
    //       EPI_WARPS / 4 warps of a quadrant =====
    const int quad = warp & 3, cgrp = (warp - 2) >> 2;   // chunks c0 = cgrp*32, step EPI_WARPS/4*32
    const int r = quad * 32 + lane;                 // accumulator row handled by this thread
    const bool live = r < rows_live;
    const size_t grow = (size_t)(row0 + r);
    tc::mbar_wait(tmem_full, 0);
    tc::tc_fence_after();
    if (threadIdx.x == 64) stamp(3);                 // accumulators complete
    // Output staging (the operand stages are idle now: every MMA has retired and every TMA write has landed).
    // thread == row makes each store instruction touch 32 different lines and the LSU serialises them; instead a warp
    // parks its 32 x 32 chunk in shared memory (16-byte slots XOR-swizzled by row: conflict-free both ways) and writes
    // it back row-contiguously (4 rows x 128 B or 8 rows x 64 B per instruction).
    uint8_t* stg = smem + (warp - 2) * 4096;
    float* sf = reinterpret_cast<float*>(stg);                      // fp32 tile [32 rows][32]
    uint32_t* sp = reinterpret_cast<uint32_t*>(stg);                // or one bf16 plane tile [32 rows][16 words]
    const int rows_q = rows_live - quad * 32;                       // live rows of this warp's quadrant
    // output row of the quadrant: global row, except for the per-pair similarity matrices (row inside the pair's matrix)
    const size_t qrow0 = (size_t)(p.ctrl_mode == 3 ? tile_s * Cfg::BM : row0) + quad * 32;
    uint4* myrow = reinterpret_cast<uint4*>(sf + lane * 32);
    auto store_f32_tile = [&](int gc) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int R = it * 4 + (lane >> 3), sl = lane & 7;
        const uint4 v = reinterpret_cast<const uint4*>(sf + R * 32)[sl ^ (R & 7)];
        if (R < rows_q) *reinterpret_cast<uint4*>(out_f32 + (qrow0 + R) * p.ld_f32 + gc + sl * 4) = v;
      }
    };
    const int np_out = p.out_planes ? p.out_planes : NP;
#pragma unroll 1
    for (int c0 = cgrp * 32; c0 < BN; c0 += 8 * Cfg::EPI_WARPS) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + c0;
      float f[32];
      if (Cfg::NACC == 1) {
        tc::tmem_ld32(taddr, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      } else {
        // main-term accumulators first (1 .. n_main), the correction accumulator last
        const int n_main = nkb < Cfg::NACC - 1 ? nkb : Cfg::NACC - 1;
        tc::tmem_ld32(taddr + BN, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
#pragma unroll 1
        for (int a = 2; a <= n_main; ++a) {
          tc::tmem_ld32(taddr + a * BN, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] += __uint_as_float(v[j]);
        }
        tc::tmem_ld32(taddr, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = NP == 2 ? fmaf(__uint_as_float(v[j]), Terms::CORR, f[j]) : f[j] + __uint_as_float(v[j]);
      }
      const int gc = n0 + c0;
      if (gc >= p.N) continue;                       // uniform per warp
      {
        const float4* b4 = reinterpret_cast<const float4*>(s_bias + c0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bv = b4[q];
          f[4 * q] += bv.x; f[4 * q + 1] += bv.y; f[4 * q + 2] += bv.z; f[4 * q + 3] += bv.w;
        }
      }
      if (p.alpha != 0.f) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] *= p.alpha;
      }
      if (p.residual && live) {
        const float4* rs = reinterpret_cast<const float4*>(p.residual + grow * p.ld_res + gc);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 rv = rs[q];
          f[4 * q] += rv.x; f[4 * q + 1] += rv.y; f[4 * q + 2] += rv.z; f[4 * q + 3] += rv.w;
        }
      }
      if (p.act == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = selu_f(f[j]);
      }
      if (p.epi == TC_EPI_ROTARY_BF16 && gc < p.rot_cols && live) {
        const int f0 = (gc & 63) >> 1;   // 0 or 16: this chunk's 16 (cos,sin) pairs are contiguous
        const float4* c4 = reinterpret_cast<const float4*>(p.rot_cos + grow * 32 + f0);
        const float4* s4 = reinterpret_cast<const float4*>(p.rot_sin + grow * 32 + f0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 cv = __ldg(c4 + q), sv = __ldg(s4 + q);
          const float cs[4] = {cv.x, cv.y, cv.z, cv.w}, sn[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = 8 * q + 2 * e;
            const float a = f[j], b = f[j + 1];
            // torch: t * cos + rotate_half(t) * sin, two rounded products then one add (no contraction)
            f[j] = __fadd_rn(__fmul_rn(a, cs[e]), __fmul_rn(-b, sn[e]));
            f[j + 1] = __fadd_rn(__fmul_rn(b, cs[e]), __fmul_rn(a, sn[e]));
          }
        }
      }
      if (p.epi == TC_EPI_F32 || p.epi == TC_EPI_F32_BF16) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          myrow[q ^ (lane & 7)] = make_uint4(__float_as_uint(f[4 * q]), __float_as_uint(f[4 * q + 1]), __float_as_uint(f[4 * q + 2]), __float_as_uint(f[4 * q + 3]));
        __syncwarp();
        store_f32_tile(gc);
        if (out_f32_t) {
          // transposed copy (similarity^T for the column-wise statistics of the assignment): lane = row of the chunk,
          // so one store instruction writes 32 consecutive floats of a transposed row
          const int ncol = n_live - gc < 32 ? n_live - gc : 32;
          if (lane < rows_q) {
            for (int c = 0; c < ncol; ++c)
              out_f32_t[(size_t)(gc + c) * p.ld_f32_t + qrow0 + lane] = sf[lane * 32 + ((((c >> 2) ^ (lane & 7)) << 2) | (c & 3))];
          }
        }
        __syncwarp();
        if (p.epi == TC_EPI_F32) continue;
      }
      if (p.epi == TC_EPI_RESID_F32_BF16) {
        // x (fp32 residual stream, updated in place): row-contiguous load into the staging tile, add, write back
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int R = it * 4 + (lane >> 3), sl = lane & 7;
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (R < rows_q) v = *reinterpret_cast<const uint4*>(out_f32 + (qrow0 + R) * p.ld_f32 + gc + sl * 4);
          reinterpret_cast<uint4*>(sf + R * 32)[sl ^ (R & 7)] = v;
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint4 xv = myrow[q ^ (lane & 7)];
          f[4 * q] += __uint_as_float(xv.x); f[4 * q + 1] += __uint_as_float(xv.y);
          f[4 * q + 2] += __uint_as_float(xv.z); f[4 * q + 3] += __uint_as_float(xv.w);
          myrow[q ^ (lane & 7)] = make_uint4(__float_as_uint(f[4 * q]), __float_as_uint(f[4 * q + 1]), __float_as_uint(f[4 * q + 2]), __float_as_uint(f[4 * q + 3]));
        }
        __syncwarp();
        store_f32_tile(gc);
        __syncwarp();
      }
      // operand planes, one at a time: plane p = bf16(residual), residual -= plane p  (== tc::pack_planes2);
      // NP == 2: the two fp16 planes (GEMM or attention operand format), with the range check
      uint32_t h1w[NP == 2 ? 16 : 1];
#pragma unroll
      for (int pl = 0; pl < NP; ++pl) {
        if (pl >= np_out) break;
        uint32_t w[16];
        if (NP == 2) {
          if (pl == 0) {
            uint32_t ov = 0u;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (p.attn_fmt) tc::pack_h2_attn(f[2 * j], f[2 * j + 1], tc::H2_ATTN_PRESCALE, w[j], h1w[j]);
              else tc::pack_h2(f[2 * j], f[2 * j + 1], w[j], h1w[j]);
              ov |= tc::h2_ovf(w[j]);
            }
            if (ov && live && p.range_flag) *p.range_flag = 1;
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) w[j] = h1w[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            w[j] = tc::pack_bf16x2(f[2 * j], f[2 * j + 1]);
            if (pl + 1 < NP) { f[2 * j] -= __uint_as_float(w[j] << 16); f[2 * j + 1] -= __uint_as_float(w[j] & 0xFFFF0000u); }
          }
        }
        uint4* prow = reinterpret_cast<uint4*>(sp + lane * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) prow[j ^ ((lane >> 1) & 3)] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int R = it * 8 + (lane >> 2), sl = lane & 3;
          const uint4 v = reinterpret_cast<const uint4*>(sp + R * 16)[sl ^ ((R >> 1) & 3)];
          if (R < rows_q) *reinterpret_cast<uint4*>(p.out_bf16 + pl * p.out_plane + (qrow0 + R) * p.ld_bf16 + gc + sl * 8) = v;
        }
        __syncwarp();
      }
    }
  }
  if (threadIdx.x == 64) stamp(4);                   // epilogue of warp 2 done
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  if (threadIdx.x == 0) stamp(5);                    // CTA end
}

}  // namespace b2s
