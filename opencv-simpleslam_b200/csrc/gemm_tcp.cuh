// gemm_tcp.cuh - PERSISTENT tcgen05/TMEM GEMM: same contract, operands, epilogues and parameter block as gemm_tc.cuh
//   C[M,N] = epi( [A1 | A2][M,K] * W[N,K]^T + bias )      (bf16 planes, NP = 1 or 3, see gemm_tc.cuh)
// but ONE CTA PER SM that walks a static round-robin list of 128 x BN output tiles, so that
//  * the per-CTA prologue (barrier init, TMEM allocation, descriptor prefetch, first TMA round trip) is paid once per
//    launch instead of once per tile (a batch of 8 pairs is 1024 tiles per layer GEMM on 148 SMs), and
//  * the epilogue of tile i (TMEM drain, bias / rotary / residual / plane split, stores) overlaps the main loop of tile
//    i + 1: tensor memory holds TWO accumulator buffers of 256 columns (tmem_full / tmem_empty barriers), the TMA -> MMA
//    shared-memory ring simply keeps running across tiles, and the epilogue warps own a dedicated staging area.
// Warp roles as before: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = epilogue (thread == accumulator row).
// fp32 fidelity (NP = 3): a buffer holds two accumulators - accumulator 0 takes the five correction terms (2^-8 of the
// magnitude), accumulator 1 the main a0*w0 term - which the epilogue adds in fp32 registers (the tensor core truncates its
// fp32 accumulator after every k-step; see gemm_tc.cuh).
#pragma once
#include "gemm_tc.cuh"

namespace b2s {

// EW = 16: the PLANES-ONLY variant for the K = 256 layer GEMMs (QKV with rotary, cross QK / V), whose tile period is set by
// the epilogue (main loop 3.5-5 k cycles, 8 epilogue warps with two 32 x 32 chunks each 7 k): sixteen epilogue warps own one
// chunk each (four resident warps per scheduler hide the tcgen05.ld / staging / store latencies), 2 KB of staging per
// warp, no fp32 output / residual / activation paths (the launcher only selects it for TC_EPI_BF16 / TC_EPI_ROTARY_BF16).
template <int BN, int NP, int EW = 0>
struct TcpGemmCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;       // one plane
  static constexpr int STAGE_BYTES = NP * (A_BYTES + B_BYTES);
#ifndef B2S_TCP_EPI_WARPS_WIDE
#define B2S_TCP_EPI_WARPS_WIDE 8
#endif
#ifndef B2S_TCP_EPI_WARPS_96
#define B2S_TCP_EPI_WARPS_96 B2S_TCP_EPI_WARPS_WIDE
#endif
  static constexpr int EPI_WARPS = EW ? EW : (BN == 64 || NP == 3) ? 8 : (BN == 96 ? B2S_TCP_EPI_WARPS_96 : B2S_TCP_EPI_WARPS_WIDE);    // <= 4 * BN / 32: every epilogue warp must own a chunk
  static constexpr bool PLANES_ONLY = EW == 16;
  static_assert(EW == 0 || (EW == 16 && BN == 128 && NP == 2), "the 16-warp epilogue is the fp16x2, 128-wide, planes-only variant");
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int STG_WARP = PLANES_ONLY ? 2048 : 4096;               // staging tile of one epilogue warp (fp32 32 x 32, or one plane 32 x 32 x 2 B)
  static constexpr int STG_BYTES = EPI_WARPS * STG_WARP;
  static constexpr int BIAS_FLOATS = 32 * EPI_WARPS;                       // one 32-float bias chunk per epilogue warp
  static constexpr int BUDGET = 227 * 1024 - STG_BYTES - 1024 /*align*/ - 256 /*barriers*/ - BIAS_FLOATS * 4;
  static constexpr int STAGES = BUDGET / STAGE_BYTES > 4 ? 4 : BUDGET / STAGE_BYTES;
  static constexpr int SMEM = STAGES * STAGE_BYTES + STG_BYTES + 1024 + 256 + BIAS_FLOATS * 4;
  static constexpr int ACC_COLS = 256;                                     // columns of one accumulator buffer
  // fp32 path: TWO accumulators whatever the tile width - correction terms / main a0*w0 term.  An output element's
  // accumulation sequence (and therefore its bits) then does not depend on BN, and the epilogue drains 2 instead of
  // 256 / BN accumulators: measured 1.50 -> 1.45 ms per match at 8 pairs, GEMM error vs float64 3e-7 (K = 256) / 5e-7
  // (K = 512) of the output scale - still below torch's own fp32 matmul (6e-7) - and every parity test unchanged.
  static constexpr int NACC = NP == 1 ? 1 : 2;
  // fp16x2 path: the a0*w0 and a0*w1 products of a k-step as ONE MMA with the two weight planes (adjacent in the stage)
  // as a 2*BN-row B operand -> [main | correction] accumulator columns, then a1*w0 into the correction columns.  The A
  // plane a0 is read from shared memory once instead of twice per k-step (the SS-form main loop is bound by shared-memory
  // bandwidth: operand reads + TMA writes).
#ifndef B2S_TCP_DIRECT_STORE
#define B2S_TCP_DIRECT_STORE 0
#endif
#ifndef B2S_TCP_STACK
#define B2S_TCP_STACK 1
#endif
  static constexpr bool STACK = NP == 2 && B2S_TCP_STACK != 0;
  static constexpr int MAIN_COL = STACK ? 0 : BN, CORR_COL = STACK ? BN : 0;   // accumulator columns inside a buffer (NACC == 2)
  static_assert(STAGES >= 2, "need a double-buffered operand ring");
  static_assert(NACC * BN <= ACC_COLS, "accumulators of one tile must fit one TMEM buffer");
};

// what a role needs to know about output tile t (all three roles derive it identically; ctrl is constant during the launch)
struct TcpTile { int row0, rows_live, tile_s, n0, w_row, n_live; size_t out_off; };

__device__ __forceinline__ TcpTile tcp_tile(const TcGemmParams& p, int t, int n_tiles, int BN) {
  TcpTile ti;
  const int tile_m = t / n_tiles;
  ti.n0 = (t - tile_m * n_tiles) * BN;
  const int gseg = tile_m / p.tiles_per_seg;
  ti.tile_s = tile_m - gseg * p.tiles_per_seg;
  const int seg = p.ctrl_mode == 3 ? 2 * gseg : gseg;
  ti.row0 = seg * p.seg_stride + ti.tile_s * 128;
  ti.rows_live = p.seg_rows - ti.tile_s * 128;
  ti.w_row = p.w_row0; ti.n_live = p.N; ti.out_off = 0;
  if (p.ctrl) {
    const int* c = p.ctrl + (seg >> 1) * TC_CTRL_INTS;
    const bool sides = c[2] > 0 && c[3] > 0;
    const bool on = sides && (p.ctrl_mode >= 2 || !c[1]);
    ti.rows_live = on ? c[2 + (seg & 1)] - ti.tile_s * 128 : 0;
    if (p.ctrl_mode == 2) {
      ti.w_row += c[6] * p.w_layer_rows;
    } else if (p.ctrl_mode == 3) {
      ti.n_live = c[3];
      if (ti.n0 >= ti.n_live) ti.rows_live = 0;
      ti.w_row = (seg + 1) * p.seg_stride;
      ti.out_off = gseg * p.out_pair_stride;
    }
  }
  if (p.m_dev) ti.rows_live = *p.m_dev * p.m_mult - tile_m * 128;
  return ti;
}

template <int BN, int NP, int EW = 0>
__global__ void __launch_bounds__(TcpGemmCfg<BN, NP, EW>::THREADS, 1) k_gemm_tcp(const __grid_constant__ CUtensorMap mapA1,
                                                     const __grid_constant__ CUtensorMap mapA2,
                                                     const __grid_constant__ CUtensorMap mapW, TcGemmParams p, int m_tiles) {
  using Cfg = TcpGemmCfg<BN, NP, EW>;
  constexpr bool PO = Cfg::PLANES_ONLY;
  using Terms = tc::PlaneTerms<NP>;
  extern __shared__ uint8_t smem_raw[];
  // aligned by pointer ARITHMETIC on the __shared__ array: an integer round trip would make every staging access a generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  // stage s: A planes at s*STAGE_BYTES + p*A_BYTES, W planes behind them; then the staging tiles, then the barriers
  uint8_t* stg_base = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_base + Cfg::STG_BYTES);
  uint64_t* full = bars;                          // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + Cfg::STAGES;           // [STAGES]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * Cfg::STAGES;   // [2]       MMA -> epilogue
  uint64_t* tmem_empty = tmem_full + 2;           // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_bias = reinterpret_cast<float*>(stg_base + Cfg::STG_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // profiling hook (b2s_bench_gemm_*, ts != nullptr): SM-clock stamps of CTA 0, [role][local tile < 16][event < 4];
  // role 0 producer, 1 MMA issuer, 2 / 3 epilogue warps 2 / 6
  const bool tr = p.ts && blockIdx.x == 0;
  auto stamp = [&](int role, int lt, int ev) { if (tr && lt < 16) p.ts[(role * 16 + lt) * 4 + ev] = (unsigned long long)clock64(); };
  const int n_tiles = (p.N + BN - 1) / BN;
  const int total = m_tiles * n_tiles;
  const int nkb = p.K / Cfg::BK;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&mapA1); tc::tma_prefetch_desc(&mapA2); tc::tma_prefetch_desc(&mapW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&tmem_full[b], 1); tc::mbar_init(&tmem_empty[b], Cfg::EPI_WARPS); }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, 512);
  pdl_trigger();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();            // everything above overlaps the previous kernel's tail; operands / device state are touched only below

  if (warp == 0) {
    // ===== TMA producer: the operand ring runs across tiles =====
    if (tc::elect_one()) {
      int it = 0, ltp = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const TcpTile ti = tcp_tile(p, t, n_tiles, BN);
        if (ti.rows_live <= 0) continue;
        stamp(0, ltp, 0);                                    // producer: first request of the tile ...
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % Cfg::STAGES, ph = (it / Cfg::STAGES) & 1;
          tc::mbar_wait(&empty[s], ph ^ 1);
          if (kb == nkb - 1) stamp(0, ltp++, 1);             // ... its last stage slot became free
          tc::mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
          uint8_t* st = smem + s * Cfg::STAGE_BYTES;
          const int k0 = kb * Cfg::BK;
#pragma unroll
          for (int pl = 0; pl < NP; ++pl) {
            if (k0 < p.K1) tc::tma_load_2d(st + pl * Cfg::A_BYTES, &mapA1, &full[s], k0, ti.row0 + pl * p.plane_rows);
            else tc::tma_load_2d(st + pl * Cfg::A_BYTES, &mapA2, &full[s], k0 - p.K1, ti.row0 + pl * p.plane_rows);
            if (p.w_plane_rows) tc::tma_load_2d(st + NP * Cfg::A_BYTES + pl * Cfg::B_BYTES, &mapW, &full[s], k0, ti.w_row + ti.n0 + pl * p.w_plane_rows);
            else tc::tma_load_2d(st + NP * Cfg::A_BYTES + pl * Cfg::B_BYTES, &mapW, &full[s], pl * p.K + k0, ti.w_row + ti.n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: tile lt accumulates into TMEM buffer lt & 1 while the epilogue drains the other one =====
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::idesc_planes<NP>(128, BN, 0, 0);
      int it = 0, lt = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const TcpTile ti = tcp_tile(p, t, n_tiles, BN);
        if (ti.rows_live <= 0) continue;
        const int ab = lt & 1, aph = (lt >> 1) & 1;
        stamp(1, lt, 0);                                    // MMA: tile start
        tc::mbar_wait(&tmem_empty[ab], aph ^ 1);            // the epilogue has drained this buffer (two tiles ago)
        tc::tc_fence_after();
        stamp(1, lt, 1);                                    // accumulator buffer free
        const uint32_t acc_base = tmem_base + ab * Cfg::ACC_COLS;
        uint32_t used = 0;                                  // accumulators already written (first MMA overwrites)
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % Cfg::STAGES, ph = (it / Cfg::STAGES) & 1;
          tc::mbar_wait(&full[s], ph);
          tc::tc_fence_after();
          if (kb == 0) stamp(1, lt, 2);                     // first operand stage of the tile landed
          const uint32_t a0 = tc::smem_u32(smem + s * Cfg::STAGE_BYTES), b0 = a0 + NP * Cfg::A_BYTES;
          if constexpr (Cfg::STACK) {
            constexpr uint32_t idesc2 = tc::idesc_planes<NP>(128, 2 * BN, 0, 0);
#pragma unroll
            for (int k = 0; k < Cfg::BK / 16; ++k) {      // a0 * [w0 ; w1] -> [main | correction]
              const uint64_t ad = tc::smem_desc_sw128(a0 + k * 32, 16, 1024);
              const uint64_t bd = tc::smem_desc_sw128(b0 + k * 32, 16, 1024);
              tc::umma_bf16(acc_base, ad, bd, idesc2, used);
              used = 1u;
            }
#pragma unroll
            for (int k = 0; k < Cfg::BK / 16; ++k) {      // a1 * w0 -> correction
              const uint64_t ad = tc::smem_desc_sw128(a0 + Cfg::A_BYTES + k * 32, 16, 1024);
              const uint64_t bd = tc::smem_desc_sw128(b0 + k * 32, 16, 1024);
              tc::umma_bf16(acc_base + BN, ad, bd, idesc, 1u);
            }
          } else {
#pragma unroll
            for (int tm = 0; tm < Terms::N; ++tm) {
              const int acc = (Cfg::NACC == 1) ? 0 : (tm == Terms::N - 1 ? 1 + kb % (Cfg::NACC - 1) : 0);
#pragma unroll
              for (int k = 0; k < Cfg::BK / 16; ++k) {
                const uint64_t ad = tc::smem_desc_sw128(a0 + Terms::a(tm) * Cfg::A_BYTES + k * 32, 16, 1024);
                const uint64_t bd = tc::smem_desc_sw128(b0 + Terms::b(tm) * Cfg::B_BYTES + k * 32, 16, 1024);
                tc::umma_bf16(acc_base + acc * BN, ad, bd, idesc, (used >> acc) & 1u);
                used |= 1u << acc;
              }
            }
          }
          tc::umma_commit(&empty[s]);                       // frees the stage when these MMAs retire
        }
        tc::umma_commit(&tmem_full[ab]);                    // accumulators of this tile complete
        stamp(1, lt, 3);                                    // all MMAs of the tile issued
        ++lt;
      }
    }
  } else {
    // ===== epilogue: a warp may touch the TMEM lane quadrant (warp % 4); the column chunks are dealt round-robin to the
    //       EPI_WARPS / 4 warps of a quadrant =====
    const int quad = warp & 3, cgrp = (warp - 2) >> 2;   // chunks c0 = cgrp*32, step EPI_WARPS/4*32
    const int r = quad * 32 + lane;                 // accumulator row handled by this thread
    // Output staging: thread == row makes each store instruction touch 32 different lines and the LSU serialises them;
    // instead a warp parks its 32 x 32 chunk in shared memory (16-byte slots XOR-swizzled by row: conflict-free both
    // ways) and writes it back row-contiguously (4 rows x 128 B or 8 rows x 64 B per instruction).
    uint8_t* stg = stg_base + (warp - 2) * Cfg::STG_WARP;
    float* sf = reinterpret_cast<float*>(stg);                      // fp32 tile [32 rows][32]
    uint32_t* sp = reinterpret_cast<uint32_t*>(stg);                // or one bf16 plane tile [32 rows][16 words]
    uint4* myrow = reinterpret_cast<uint4*>(sf + lane * 32);
    const int np_out = p.out_planes ? p.out_planes : NP;
    int lt = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      const TcpTile ti = tcp_tile(p, t, n_tiles, BN);
      if (ti.rows_live <= 0) continue;
      const int ab = lt & 1, aph = (lt >> 1) & 1;
      const bool live = r < ti.rows_live;
      const size_t grow = (size_t)(ti.row0 + r);
      float* out_f32 = p.out_f32 ? p.out_f32 + ti.out_off : nullptr;
      float* out_f32_t = p.out_f32_t ? p.out_f32_t + ti.out_off : nullptr;
      const int rows_q = ti.rows_live - quad * 32;                       // live rows of this warp's quadrant
      // output row of the quadrant: global row, except for the per-pair similarity matrices (row inside the pair's matrix)
      const size_t qrow0 = (size_t)(p.ctrl_mode == 3 ? ti.tile_s * 128 : ti.row0) + quad * 32;
      auto store_f32_tile = [&](int gc) {
#pragma unroll
        for (int i8 = 0; i8 < 8; ++i8) {
          const int R = i8 * 4 + (lane >> 3), sl = lane & 7;
          const uint4 v = reinterpret_cast<const uint4*>(sf + R * 32)[sl ^ (R & 7)];
          if (R < rows_q) *reinterpret_cast<uint4*>(out_f32 + (qrow0 + R) * p.ld_f32 + gc + sl * 4) = v;
        }
      };
      if (lane == 0 && (warp == 2 || warp == 6)) stamp(warp == 2 ? 2 : 3, lt, 0);   // epilogue: waiting for the accumulators
      // the bias of this warp's chunks (lane = column) is requested before the wait for the accumulators: eight broadcast
      // __ldg per chunk inside the thread == row epilogue were measured at 11 % of the K = 256 GEMMs
      float breg[(BN + 8 * Cfg::EPI_WARPS - 1) / (8 * Cfg::EPI_WARPS)];
      if (p.bias) {
        const float* bsrc = p.bias + ti.w_row * (p.ctrl_mode == 2 ? 1 : 0) + ti.n0 + lane;
#pragma unroll
        for (int k = 0; k < (int)(sizeof(breg) / sizeof(float)); ++k) {
          const int c0 = cgrp * 32 + k * 8 * Cfg::EPI_WARPS;
          breg[k] = (c0 < BN && ti.n0 + c0 < p.N) ? __ldg(bsrc + c0) : 0.f;
        }
      }
      // TC_EPI_RESID_F32_BF16: the fp32 residual stream x of this warp's chunk (row-contiguous, staged like the outputs) is
      // requested before the accumulators are waited for, the next chunk's while the current one is packed and stored -
      // the loads sat at the head of the epilogue's dependency chain (ncu source page: 60 % of the FFN2 launch's stall
      // samples on the eight stores that park x in the staging tile)
      uint4 xp[8];
      auto load_x = [&](int gc) {
#pragma unroll
        for (int i8 = 0; i8 < 8; ++i8) {
          const int R = i8 * 4 + (lane >> 3), sl = lane & 7;
          xp[i8] = make_uint4(0u, 0u, 0u, 0u);
          if (R < rows_q) xp[i8] = *reinterpret_cast<const uint4*>(out_f32 + (qrow0 + R) * p.ld_f32 + gc + sl * 4);
        }
      };
      if (!PO && p.epi == TC_EPI_RESID_F32_BF16 && ti.n0 + cgrp * 32 < p.N) load_x(ti.n0 + cgrp * 32);
      tc::mbar_wait(&tmem_full[ab], aph);
      tc::tc_fence_after();
      if (lane == 0 && (warp == 2 || warp == 6)) stamp(warp == 2 ? 2 : 3, lt, 1);   // accumulators complete
      const uint32_t acc_base = tmem_base + ab * Cfg::ACC_COLS;
#pragma unroll 1
      for (int c0 = cgrp * 32; c0 < BN; c0 += 8 * Cfg::EPI_WARPS) {
        uint32_t v[32];
        const uint32_t taddr = acc_base + ((uint32_t)(quad * 32) << 16) + c0;
        float f[32];
        if (Cfg::NACC == 1) {
          tc::tmem_ld32(taddr, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        } else {
          // main-term accumulators first (1 .. n_main), the correction accumulator last
          const int n_main = nkb < Cfg::NACC - 1 ? nkb : Cfg::NACC - 1;
          tc::tmem_ld32(taddr + Cfg::MAIN_COL, v);
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
          tc::tmem_ld32(taddr + Cfg::CORR_COL, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = NP == 2 ? fmaf(__uint_as_float(v[j]), Terms::CORR, f[j]) : f[j] + __uint_as_float(v[j]);
        }
        if (c0 + 8 * Cfg::EPI_WARPS >= BN) {
          // that was this warp's last read of the buffer: hand it back to the MMA issuer before the stores
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&tmem_empty[ab]);
          if (lane == 0 && (warp == 2 || warp == 6)) stamp(warp == 2 ? 2 : 3, lt, 2);   // buffer handed back
        }
        const int gc = ti.n0 + c0;
        if (gc >= p.N) continue;                       // uniform per warp
        if (p.bias) {
          float* sb = s_bias + (warp - 2) * 32;
          constexpr int NB = (int)(sizeof(breg) / sizeof(float));
          static_assert(NB <= 2, "an epilogue warp owns at most two chunks");
          sb[lane] = (NB > 1 && c0 >= 8 * Cfg::EPI_WARPS) ? breg[NB - 1] : breg[0];
          __syncwarp();
          const float4* b4 = reinterpret_cast<const float4*>(sb);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bv = b4[q];
            f[4 * q] += bv.x; f[4 * q + 1] += bv.y; f[4 * q + 2] += bv.z; f[4 * q + 3] += bv.w;
          }
          __syncwarp();
        }
        if (!PO && p.alpha != 0.f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] *= p.alpha;
        }
        if (!PO && p.residual && live) {
          const float4* rs = reinterpret_cast<const float4*>(p.residual + grow * p.ld_res + gc);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 rv = rs[q];
            f[4 * q] += rv.x; f[4 * q + 1] += rv.y; f[4 * q + 2] += rv.z; f[4 * q + 3] += rv.w;
          }
        }
        if (!PO && p.act == 1) {
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
        if (!PO && (p.epi == TC_EPI_F32 || p.epi == TC_EPI_F32_BF16)) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            myrow[q ^ (lane & 7)] = make_uint4(__float_as_uint(f[4 * q]), __float_as_uint(f[4 * q + 1]), __float_as_uint(f[4 * q + 2]), __float_as_uint(f[4 * q + 3]));
          __syncwarp();
          store_f32_tile(gc);
          if (out_f32_t) {
            // transposed copy (similarity^T for the column-wise statistics of the assignment): lane = row of the chunk,
            // so one store instruction writes 32 consecutive floats of a transposed row
            const int ncol = ti.n_live - gc < 32 ? ti.n_live - gc : 32;
            if (lane < rows_q) {
              for (int c = 0; c < ncol; ++c)
                out_f32_t[(size_t)(gc + c) * p.ld_f32_t + qrow0 + lane] = sf[lane * 32 + ((((c >> 2) ^ (lane & 7)) << 2) | (c & 3))];
            }
          }
          __syncwarp();
          if (p.epi == TC_EPI_F32) continue;
        }
        if (!PO && p.epi == TC_EPI_RESID_F32_BF16) {
          // x (fp32 residual stream, updated in place): row-contiguous load into the staging tile, add, write back
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            const int R = i8 * 4 + (lane >> 3), sl = lane & 7;
            reinterpret_cast<uint4*>(sf + R * 32)[sl ^ (R & 7)] = xp[i8];
          }
          {
            const int c1 = c0 + 8 * Cfg::EPI_WARPS;
            if (c1 < BN && ti.n0 + c1 < p.N) load_x(ti.n0 + c1);      // next chunk of this warp, in flight behind this one
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
#if B2S_TCP_DIRECT_STORE
          // planes straight from registers: a thread owns 64 contiguous bytes of its row (no shared-memory round trip -
          // the staging tile competes with the MMA operand reads and the TMA writes for shared-memory bandwidth)
          if (lane < rows_q) {
            uint4* dst = reinterpret_cast<uint4*>(p.out_bf16 + pl * p.out_plane + (qrow0 + lane) * p.ld_bf16 + gc);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
          }
          continue;
#endif
          uint4* prow = reinterpret_cast<uint4*>(sp + lane * 16);
#pragma unroll
          for (int j = 0; j < 4; ++j) prow[j ^ ((lane >> 1) & 3)] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
          __syncwarp();
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const int R = i4 * 8 + (lane >> 2), sl = lane & 3;
            const uint4 v4 = reinterpret_cast<const uint4*>(sp + R * 16)[sl ^ ((R >> 1) & 3)];
            if (R < rows_q) *reinterpret_cast<uint4*>(p.out_bf16 + pl * p.out_plane + (qrow0 + R) * p.ld_bf16 + gc + sl * 8) = v4;
          }
          __syncwarp();
        }
      }
      if (lane == 0 && (warp == 2 || warp == 6)) stamp(warp == 2 ? 2 : 3, lt, 3);     // stores of the tile issued
      ++lt;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace b2s
