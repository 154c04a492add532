// attn_tc3.cuh - fp32-faithful flash attention on tcgen05/TMEM: Q, K, V and the probabilities P
// are carried as three bf16 planes each (tc_common.cuh) and every contraction issues the six cross
// products a_i b_j (i + j <= 2), so S = Q K^T and O = P V are accumulated to fp32 rounding level
// on the tensor pipe; the softmax itself runs in fp32 registers.
//   O = softmax(Q K^T * scale) V,  head_dim 64, 128-query tile per CTA, 64-key tiles.
// Same role split as attn_tc.cuh (two softmax groups on alternating key tiles, TMA warp, MMA warp, lazy
// rescale).  Differences:
//  * P never touches shared memory: the softmax threads write its three planes into TENSOR MEMORY
//    (tcgen05.st, lane = query row, one column = two keys) and the P V MMAs take their A operand from
//    there (TS form).  Shared memory then only feeds the B operands, which is what bounds this kernel
//    (head_dim 64 means N = 64 MMAs: an A operand from shared memory would cost 2x the B traffic);
//  * the tensor core truncates its fp32 accumulator after every k-step, so a tile's six terms are issued
//    smallest first (the main a0*b0 term last: only its 4 k-steps round at full magnitude) and the per-tile
//    P V result is added to the running output in fp32 REGISTERS instead of being accumulated across tiles;
//  * the score MMAs are issued for PAIRS of key tiles (N = 128: tile 2i for group 0 and 2i+1 for group 1 land
//    in adjacent TMEM columns), which halves the re-reads of Q from shared memory; K therefore arrives as
//    128-key tiles (2-slot ring, 3 planes x 16 KB), V as 64-key tiles (3-slot ring, 3 planes x 8 KB).
// TMEM (512 columns allocated): S[buffer][g] 64 | O[g] 64 | P[g] NP x 32, g = 0, 1 (one score buffer with NP = 3, two with
// NP = 2, see A3Cfg).   grid = (ceil(max nq/128), heads, nprob).
#pragma once
#include "attn_tc.cuh"

namespace b2s {

constexpr int A3_BQ = 128, A3_BK = 64, A3_D = 64, A3_SLOTS = 3;
constexpr int A3_QPL = A3_BQ * A3_D * 2;             // 16 KB: one Q plane  [128 x 64] 16-bit
constexpr int A3_KPL = A3_BK * A3_D * 2;             //  8 KB: one K / V plane [64 x 64] 16-bit
constexpr int A3_THREADS = 384;
// NP = 3: fp32 as three bf16 planes (six products per contraction); NP = 2: two fp16 planes in the attention operand
// format of tc_common.cuh (three products; q, k, v prescaled by 16, P by 2^7 - the lazy softmax reference lets P reach 2^8).
template <int NP>
struct A3Cfg {
  static constexpr int SLOT = NP * A3_KPL;             // one V slot
  static constexpr int KSLOTS = NP == 3 ? 2 : 3;       // K slots = pairs of key tiles
  static constexpr int KSLOT = NP * A3_QPL;
  static constexpr int SMEM = NP * A3_QPL + KSLOTS * KSLOT + A3_SLOTS * SLOT + 1024 + 256;
  static constexpr int PCOLS = NP * 32;                // tensor-memory columns of one group's P planes
  // score buffers in tensor memory: S[buffer][group] 64 columns each.  NP = 2 leaves room for TWO buffers (2 x 128 S + 128 O
  // + 128 P = 512 columns): the score pair of key tiles (2i + 2, 2i + 3) is issued while the softmax groups still work on
  // pair i, so no group ever waits for its scores (NP = 3: 128 S + 128 O + 192 P, one buffer)
  // (measured with NP = 2, 2048 x 2048: two buffers + staggered groups 29.4 us per launch vs 29.8 with one - the softmax
  //  groups, not the score MMAs, set the period - so one buffer is used)
  static constexpr int SBUF = 1;
  static_assert(2 * 66 * 128 * 4 <= KSLOTS * KSLOT, "the final merge buffers alias the K ring");
  static_assert(8 * 2048 <= A3_SLOTS * SLOT, "the output staging tiles alias the V ring");
};
constexpr float A3_P_PRESCALE = 128.f, A3_P_PRESCALE_LOG2 = 7.f;   // NP = 2: P is stored as 2^7 P (the exponent gets + 7: no multiply)

struct Attn3Params {
  AttnTcProb prob[2]; int cap;     // problems: see AttnTcParams
  int qcol, kcol, vcol;            // column offsets of Q / K / V (head h adds h*64)
  int plane_rows;                  // rows between operand planes in the q/k/v buffer
  float scale_log2e;
  __nv_bfloat16* out; int ldo; size_t out_plane;   // ctx planes [3][2*cap, 256] bf16
  const int* ctrl; int cross;
  unsigned long long* stats;       // nullable: stats[cross] += nq * nk of every live problem
  int pv_issuers;                  // 1: one thread issues every P V product; 2: one thread per softmax group
  int trace_cta;                   // CTA to trace: x | y << 8 | z << 16
  long long* trace;                // nullable profiling hook (b2s_trace_attn_tc3): clock64 stamps of CTA (0,0,0), [role][tile][event]
  int* range_flag;                 // NP == 2, nullable: set to 1 when an output value leaves the fp16 range
  float out_scale;                 // NP == 2: 1 / (v prescale) applied to the normalised output (0 means 1); P and l carry the same 2^7
};

// P chunk c (32 keys) of one row -> three bf16 planes in REGISTERS (16 packed words each); returns the partial row sum.
// The planes go to tensor memory later (a3_store_p_chunk), once the previous P V of the group has retired: everything
// that does not depend on that MMA (exponentials, plane split) is done while it is still running.
template <bool MASK, int NP>
__device__ __forceinline__ float a3_make_p_chunk(const uint32_t (&v)[32], int c, int limit, float scale, float m_used, uint32_t (&pk)[NP][16]) {
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int t = 0; t < 32; t += 2) {
    float p0 = ex2_approx(fmaf(__uint_as_float(v[t]), scale, -m_used));
    float p1 = ex2_approx(fmaf(__uint_as_float(v[t + 1]), scale, -m_used));
    if (MASK) {
      if (c * 32 + t >= limit) p0 = 0.f;
      if (c * 32 + t + 1 >= limit) p1 = 0.f;
    }
    sum0 += p0; sum1 += p1;
    uint32_t w[NP];
    if (NP == 2) tc::pack_h2_raw(p0, p1, w[0], w[NP - 1]);     // the caller folded the 2^7 prescale into m_used
    else tc::pack_planes2<NP>(p0, p1, w);
#pragma unroll
    for (int pl = 0; pl < NP; ++pl) pk[pl][t >> 1] = w[pl];
  }
  return sum0 + sum1;
}
template <int NP>
__device__ __forceinline__ void a3_store_p_chunk(const uint32_t (&pk)[NP][16], int c, uint32_t p_addr) {
#pragma unroll
  for (int pl = 0; pl < NP; ++pl) tc::tmem_st16(p_addr + pl * 32 + c * 16, pk[pl]);
}

template <int NP>
__global__ void __launch_bounds__(A3_THREADS, 1) k_attn_tc3(const __grid_constant__ CUtensorMap mapQ,
                                                            const __grid_constant__ CUtensorMap mapKV, Attn3Params p) {
  using Cfg = A3Cfg<NP>;
  constexpr int A3_NP = NP, A3_SLOT = Cfg::SLOT, A3_KSLOTS = Cfg::KSLOTS, A3_KSLOT = Cfg::KSLOT;
  extern __shared__ uint8_t smem_raw[];
  // aligned by pointer ARITHMETIC on the __shared__ array: an integer round trip would make every staging access a generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                               // [3 planes]
  uint8_t* sK = sQ + A3_NP * A3_QPL;                // ring [A3_KSLOTS][3 planes][128 keys]; reused for the final merge
  uint8_t* sV = sK + A3_KSLOTS * A3_KSLOT;          // ring [A3_SLOTS][3 planes][64 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + A3_SLOTS * A3_SLOT);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;                 uint64_t* k_empty = k_full + A3_SLOTS;
  uint64_t* v_full = k_empty + A3_SLOTS;       uint64_t* v_empty = v_full + A3_SLOTS;
  uint64_t* s_full = v_empty + A3_SLOTS;       uint64_t* s_free = s_full + 4;      // [buffer][group]
  uint64_t* p_full = s_free + 4;               uint64_t* o_full = p_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const long long t_entry = clock64();
  AttnTcProb pr = p.prob[p.ctrl ? 0 : blockIdx.z];
  const int q0 = blockIdx.x * A3_BQ;
  if (!p.ctrl && q0 >= pr.nq) return;                    // uniform per CTA (static sizes)
  const int h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 8 && lane == 0) { tc::tma_prefetch_desc(&mapQ); tc::tma_prefetch_desc(&mapKV); }
  if (warp == 9 && lane == 0) {
    tc::mbar_init(q_full, 1);
    for (int b = 0; b < A3_SLOTS; ++b) {
      tc::mbar_init(&k_full[b], 1); tc::mbar_init(&k_empty[b], 1);
      tc::mbar_init(&v_full[b], 1); tc::mbar_init(&v_empty[b], 1);
    }
    for (int b = 0; b < 4; ++b) { tc::mbar_init(&s_full[b], 1); tc::mbar_init(&s_free[b], 128); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&p_full[b], 128); tc::mbar_init(&o_full[b], 1); }
    tc::fence_barrier_init();
  }
  if (warp == 10) tc::tmem_alloc(tmem_slot, 512);
  pdl_trigger();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  if (p.ctrl) {          // device-resident sizes (pruning / early exit)
    const int z = blockIdx.z, side = z & 1;
    const int* c = p.ctrl + (z >> 1) * 32;
    const bool on = !c[1] && c[2] > 0 && c[3] > 0;
    pr.q_row = z * p.cap; pr.k_row = (p.cross ? z ^ 1 : z) * p.cap;
    pr.nq = on ? c[2 + side] : 0;
    pr.nk = c[2 + (p.cross ? 1 - side : side)];
  }
  const bool live = q0 < pr.nq;                          // uniform per CTA
  const int nt = live ? (pr.nk + A3_BK - 1) / A3_BK : 0;
  if (p.stats && live && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)
    atomicAdd(&p.stats[p.cross ? 1 : 0], (unsigned long long)pr.nq * (unsigned long long)pr.nk);
  const uint32_t tS = tmem_base, tO = tmem_base + Cfg::SBUF * 128, tP = tO + 128;
  const bool tr = p.trace && (int)(blockIdx.x | (blockIdx.y << 8) | (blockIdx.z << 16)) == p.trace_cta;
  auto stamp = [&](int role, int tile, int ev) { if (tr) p.trace[(role * 64 + tile) * 8 + ev] = clock64(); };
  if (tr && threadIdx.x == 0) p.trace[(0 * 64 + 63) * 8 + 0] = t_entry;   // kernel entry
  if (threadIdx.x == 0) stamp(0, 63, 1);                   // dependencies resolved, sizes known   // S[g] + 64 g, O[g] + 64 g, P[g] + 96 g

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  }
  if (!live) {
    // nothing to do for this query tile
  } else if (warp == 8) {
    if (tc::elect_one()) {
      tc::mbar_expect_tx(q_full, A3_NP * A3_QPL);
#pragma unroll
      for (int pl = 0; pl < A3_NP; ++pl)
        tc::tma_load_2d(sQ + pl * A3_QPL, &mapQ, q_full, p.qcol + h * 64, pl * p.plane_rows + pr.q_row + q0);
      // K arrives as 128-key tiles, one per pair of key tiles, and is requested ONE PAIR AHEAD of the V tiles: its ring
      // slot frees as soon as the score MMAs of pair jp - 1 retire, whereas a V slot only frees after the P V of three
      // tiles earlier - issued in tile order, a K request would queue behind that wait and land too late for the next
      // score pair (measured: ~1500 cycles of the S issuer waiting for K per pair).
      auto load_k = [&](int jp) {
        const int kb = jp % A3_KSLOTS, kph = (jp / A3_KSLOTS) & 1;
        tc::mbar_wait(&k_empty[kb], kph ^ 1);
        tc::mbar_expect_tx(&k_full[kb], A3_KSLOT);
#pragma unroll
        for (int pl = 0; pl < A3_NP; ++pl)
          tc::tma_load_2d(sK + kb * A3_KSLOT + pl * A3_QPL, &mapQ, &k_full[kb], p.kcol + h * 64, pl * p.plane_rows + pr.k_row + jp * 2 * A3_BK);
      };
      const int npairs = (nt + 1) >> 1;
      constexpr int KAHEAD = Cfg::SBUF;                    // pairs requested ahead of the V tiles (<= A3_KSLOTS - 1)
      for (int jp = 0; jp < KAHEAD && jp < npairs; ++jp) load_k(jp);
      for (int j = 0; j < nt; ++j) {
        if ((j & 1) == 0 && (j >> 1) + KAHEAD < npairs) load_k((j >> 1) + KAHEAD);
        const int b = j % A3_SLOTS, ph = (j / A3_SLOTS) & 1;
        tc::mbar_wait(&v_empty[b], ph ^ 1);
        tc::mbar_expect_tx(&v_full[b], A3_SLOT);
#pragma unroll
        for (int pl = 0; pl < A3_NP; ++pl)
          tc::tma_load_2d(sV + b * A3_SLOT + pl * A3_KPL, &mapKV, &v_full[b], p.vcol + h * 64, pl * p.plane_rows + pr.k_row + j * A3_BK);
      }
    }
  } else if (warp == 9) {
    if (tc::elect_one()) {
      using Terms = tc::PlaneTerms<A3_NP>;
      constexpr uint32_t idesc_s = tc::idesc_planes<NP>(128, 128, 0, 0);   // S pair: A = Q (K-major), B = 128 keys (K-major)
      const uint32_t q_addr = tc::smem_u32(sQ);
      // (the barrier that completes LAST is waited on last: a wait on an already-complete mbarrier still costs ~100
      //  cycles, which must not sit between the late event and the first MMA)
      auto wait_k = [&](int jp) { tc::mbar_wait(&k_full[jp % A3_KSLOTS], (jp / A3_KSLOTS) & 1); };
      auto issue_s_pair = [&](int jp) {                              // tiles 2jp (-> S[.][0]) and 2jp + 1 (-> S[.][1]); K pair jp has landed
        const int b = jp % A3_KSLOTS, sb = jp % Cfg::SBUF;
        tc::tc_fence_after();
        stamp(0, jp, 0);                                   // K pair landed, S pair issue starts
        const uint32_t k_addr = tc::smem_u32(sK + b * A3_KSLOT);
#pragma unroll
        for (int t = 0; t < Terms::N; ++t)
#pragma unroll
          for (int k = 0; k < A3_D / 16; ++k) {
            const uint64_t ad = tc::smem_desc_sw128(q_addr + Terms::a(t) * A3_QPL + k * 32, 16, 1024);
            const uint64_t bd = tc::smem_desc_sw128(k_addr + Terms::b(t) * A3_QPL + k * 32, 16, 1024);
            tc::umma_bf16(tS + sb * 128, ad, bd, idesc_s, (t | k) ? 1u : 0u);
          }
        tc::umma_commit(&s_full[2 * sb]);
        tc::umma_commit(&s_full[2 * sb + 1]);
        tc::umma_commit(&k_empty[b]);
        stamp(0, jp, 1);                                   // S pair issued
      };
      tc::mbar_wait(q_full, 0);
      const int npairs = (nt + 1) >> 1;
      for (int jp = 0; jp < npairs; ++jp) {
        wait_k(jp);                           // landed long ago (requested ahead)
        if (jp >= Cfg::SBUF) {                // both groups have the previous tiles of this score buffer in registers
          const int sb = jp % Cfg::SBUF, ph = (jp / Cfg::SBUF - 1) & 1;
          tc::mbar_wait(&s_free[2 * sb], ph);
          tc::mbar_wait(&s_free[2 * sb + 1], ph);
          stamp(0, jp - 1, 6);
        }
        issue_s_pair(jp);
      }
    }
  } else if (warp == 10 || (warp == 11 && p.pv_issuers == 2)) {
    // ===== second (and third) MMA issuer: the P V products.  An issuing thread is blocked for about as long as its MMAs execute
    //       and every commit / barrier round trip costs it a few hundred cycles, so with ONE issuer the tensor pipe idles
    //       between the S pair and the two P V groups of a tile pair (measured: 3440 busy of 4735 cycles per pair).  The
    //       score and the P V MMAs touch disjoint tensor-memory columns and are ordered by mbarriers only (s_free / p_full /
    //       o_full), so they can be issued from two threads whose gaps overlap.  With pv_issuers == 2 the two softmax
    //       groups' P V products come from two threads as well (warp 10: even key tiles, warp 11: odd ones): every
    //       issue batch pays ~450 cycles of barrier round trips before its first MMA (trace: P published -> P V start),
    //       and the pipe idles whenever all issuers are in that phase at once. =====
    if (tc::elect_one()) {
      using Terms = tc::PlaneTerms<A3_NP>;
      constexpr uint32_t idesc_o = tc::idesc_planes<NP>(128, 64, 0, 1);    // PV: A = P (tensor memory), B = V (MN-major)
      const int jstep = p.pv_issuers == 2 ? 2 : 1;
      for (int j = (p.pv_issuers == 2 ? warp - 10 : 0); j < nt; j += jstep) {
        const int b = j % A3_SLOTS, ph = (j / A3_SLOTS) & 1, g = j & 1;
        tc::mbar_wait(&v_full[b], ph);                // V tile j landed (early)
        tc::mbar_wait(&p_full[g], (j >> 1) & 1);      // P of tile j is in tensor memory (the late event: waited on last)
        tc::tc_fence_after();
        stamp(0, j >> 1, 2 + 2 * (j & 1));
        const uint32_t v_addr = tc::smem_u32(sV + b * A3_SLOT);
#pragma unroll
        for (int t = 0; t < Terms::N; ++t)
#pragma unroll
          for (int kk = 0; kk < A3_BK / 16; ++kk) {
            // P plane: 32 columns (64 keys), 16 keys = 8 columns; V plane: [64 keys x 64 d] MN-major, 16 keys = 2048 B
            const uint64_t bd = tc::smem_desc_sw128(v_addr + Terms::b(t) * A3_KPL + kk * 2048, 16, 1024);
            tc::umma_bf16_ts(tO + g * 64, tP + g * Cfg::PCOLS + Terms::a(t) * 32 + kk * 8, bd, idesc_o, (t | kk) ? 1u : 0u);   // fresh per tile
          }
        tc::umma_commit(&o_full[g]);
        tc::umma_commit(&v_empty[b]);
        stamp(0, j >> 1, 3 + 2 * (j & 1));
      }
    }
  } else if (warp < 8) {
    // ===== softmax groups: g = 0 (warps 0..3) takes even key tiles, g = 1 (warps 4..7) odd ones =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int g = warp >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                       // query row of this thread
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    float m_used = -INFINITY, l_run = 0.f;
    float o[A3_D];                                        // running output (fp32 registers)
#pragma unroll
    for (int d = 0; d < A3_D; ++d) o[d] = 0.f;
    const uint32_t s_base = tS + g * 64 + lane_addr, o_addr = tO + g * 64 + lane_addr, p_addr = tP + g * Cfg::PCOLS + lane_addr;
    // o += the P V result of one tile sitting in TMEM (exact fp32 adds)
    auto add_pv = [&]() {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t a[32];
        tc::tmem_ld32(o_addr + 32 * c, a);
        tc::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) o[32 * c + e] += __uint_as_float(a[e]);
      }
    };
    int t = 0;                                            // index among this group's tiles
    for (int j = g; j < nt; j += 2, ++t) {
      const int sb = t % Cfg::SBUF;                        // score buffer of pair t
      tc::mbar_wait(&s_full[2 * sb + g], (t / Cfg::SBUF) & 1);
      tc::tc_fence_after();
      if (quad == 0 && lane == 0) stamp(1 + g, t, 0);       // S tile complete (seen by the group)
      uint32_t s[2][32];
      const uint32_t s_addr = s_base + sb * 128;
      tc::tmem_ld32(s_addr, s[0]); tc::tmem_ld32(s_addr + 32, s[1]);
      tc::tmem_ld_wait();
      tc::tc_fence_before();
      tc::mbar_arrive(&s_free[2 * sb + g]);               // S[sb][g] is in registers: a later QK^T of this group may overwrite it
      const int limit = pr.nk - j * A3_BK;
      const bool full = limit >= A3_BK;
      float mx;
      if (full) mx = fmaxf(atc_chunk_max<false>(s[0], 0, limit), atc_chunk_max<false>(s[1], 1, limit));
      else mx = fmaxf(atc_chunk_max<true>(s[0], 0, limit), atc_chunk_max<true>(s[1], 1, limit));
      const float mxs = mx * p.scale_log2e;
      // reference maximum of this tile (lazy: raised only beyond 2^8) - decided before the exponentials
      float corr = 1.f;
      bool rescale = false;
      if (t == 0) {
        m_used = mxs;
      } else if (__any_sync(0xffffffffu, mxs > m_used + ATC_LAZY)) {
        const float m_new = fmaxf(m_used, mxs);
        corr = ex2_approx(m_used - m_new);                // 1 for rows whose reference did not move
        m_used = m_new;
        rescale = true;
      }
      uint32_t pk0[A3_NP][16], pk1[A3_NP][16];
      float sum;
      // NP == 2: P and its row sum are 2^7 too large (exponent offset); the final normalisation O / l cancels it
      const float m_eff = NP == 2 ? m_used - A3_P_PRESCALE_LOG2 : m_used;
      if (full) sum = a3_make_p_chunk<false, NP>(s[0], 0, limit, p.scale_log2e, m_eff, pk0) + a3_make_p_chunk<false, NP>(s[1], 1, limit, p.scale_log2e, m_eff, pk1);
      else sum = a3_make_p_chunk<true, NP>(s[0], 0, limit, p.scale_log2e, m_eff, pk0) + a3_make_p_chunk<true, NP>(s[1], 1, limit, p.scale_log2e, m_eff, pk1);
      if (t > 0) {
        // PV of the group's previous tile must have retired before P[g] is rewritten; collect its result (it was
        // computed against the previous reference maximum: add first, rescale afterwards)
        if (quad == 0 && lane == 0) stamp(1 + g, t, 1);     // exponentials / planes done
        tc::mbar_wait(&o_full[g], (t - 1) & 1);
        tc::tc_fence_after();
        if (quad == 0 && lane == 0) stamp(1 + g, t, 2);     // previous P V retired
        add_pv();
        if (rescale) {
          l_run *= corr;
#pragma unroll
          for (int d = 0; d < A3_D; ++d) o[d] *= corr;
        }
      }
      l_run += sum;
      a3_store_p_chunk<NP>(pk0, 0, p_addr);
      a3_store_p_chunk<NP>(pk1, 1, p_addr);
      tc::tmem_st_wait();               // P is in tensor memory
      tc::tc_fence_before();            // order our tcgen05.ld / st before the MMAs that follow the arrive
      tc::mbar_arrive(&p_full[g]);
      if (quad == 0 && lane == 0) stamp(1 + g, t, 3);       // P published
    }
    if (quad == 0 && lane == 0) stamp(1 + g, 63, 0);        // key loop done
    // ---- the group's last PV ----
    if (t > 0) {
      tc::mbar_wait(&o_full[g], (t - 1) & 1);
      tc::tc_fence_after();
      add_pv();
    }
    const float m_run = m_used;
    tc::tc_fence_before();
    // ---- merge the two groups' partial softmax states (every MMA and every TMA load has completed) ----
    // Both groups park their state in shared memory ([d][row] floats, the K ring is idle), then group g produces output
    // dims [32 g, 32 g + 32) of every row: merge, split into planes, and - as thread == row would make every store
    // instruction touch 32 lines - write through a swizzled staging tile so that a store covers 8 rows x 64 B.
    float* mrg = reinterpret_cast<float*>(sK) + g * (66 * 128);       // [66][128]: O^T (64 rows), m, l of group g
    asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
    for (int d = 0; d < A3_D; ++d) mrg[d * 128 + r] = o[d];
    mrg[64 * 128 + r] = m_run; mrg[65 * 128 + r] = l_run;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    {
      const float* ma = reinterpret_cast<const float*>(sK);            // group 0's state
      const float* mb = ma + 66 * 128;                                 // group 1's state
      const float m_a = ma[64 * 128 + r], l_a = ma[65 * 128 + r], m_b = mb[64 * 128 + r], l_b = mb[65 * 128 + r];
      const float m = fmaxf(m_a, m_b);
      const float ca = ex2_approx(m_a - m), cb = (m_b == -INFINITY) ? 0.f : ex2_approx(m_b - m);
      float inv = 1.f / (l_a * ca + l_b * cb);
      if (NP == 2 && p.out_scale != 0.f) inv *= p.out_scale;          // power of two: undoes the operand prescales exactly
      const float fa = ca * inv, fb = cb * inv;
      float f[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) { const int d = 32 * g + e; f[e] = ma[d * 128 + r] * fa + mb[d * 128 + r] * fb; }
      uint32_t* sp = reinterpret_cast<uint32_t*>(sV) + warp * 512;     // staging tile of this warp: [32 rows][16 words]
      const int rows_q = pr.nq - q0 - quad * 32;                       // live rows of this warp's quadrant
      __nv_bfloat16* dst = p.out + (size_t)(pr.q_row + q0 + quad * 32) * p.ldo + h * 64 + 32 * g;
      uint32_t h1w[NP == 2 ? 16 : 1];
#pragma unroll
      for (int pl = 0; pl < A3_NP; ++pl) {
        uint32_t w[16];
        if (NP == 2) {
          if (pl == 0) {
            uint32_t ov = 0u;
#pragma unroll
            for (int j = 0; j < 16; ++j) { tc::pack_h2(f[2 * j], f[2 * j + 1], w[j], h1w[j]); ov |= tc::h2_ovf(w[j]); }
            if (ov && lane < rows_q && p.range_flag) *p.range_flag = 1;
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) w[j] = h1w[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            w[j] = tc::pack_bf16x2(f[2 * j], f[2 * j + 1]);
            if (pl + 1 < A3_NP) { f[2 * j] -= __uint_as_float(w[j] << 16); f[2 * j + 1] -= __uint_as_float(w[j] & 0xFFFF0000u); }
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
          if (R < rows_q) *reinterpret_cast<uint4*>(dst + pl * p.out_plane + (size_t)R * p.ldo + sl * 8) = v;
        }
        __syncwarp();
      }
    }
  }
  if (threadIdx.x == 0) stamp(0, 63, 2);                   // output written (thread 0)
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) stamp(0, 63, 3);                   // CTA done
  if (warp == 10) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace b2s
