// attn_tc3.cuh - fp32-faithful flash attention on tcgen05/TMEM: Q, K, V and the probabilities P
// are carried as three bf16 planes each (tc_common.cuh) and every contraction issues the six cross
// products a_i b_j (i + j <= 2), so S = Q K^T and O = P V are accumulated to fp32 rounding level
// on the tensor pipe; the softmax itself runs in fp32 registers.
//   O = softmax(Q K^T * scale) V,  head_dim 64, 128-query tile per CTA, 64-key tiles.
// Same role split as attn_tc.cuh (two softmax groups on alternating key tiles, TMA warp, MMA warp,
// lazy rescale with O kept in TMEM).  Differences: K and V tiles (3 planes x 8 KB) share ONE ring of
// 3 slots filled in exactly the order the MMA warp consumes them (K0 K1 [K2 V0] [K3 V1] ...), P has
// three planes per group.  The tensor core truncates its fp32 accumulator after every k-step, so every
// contraction keeps TWO TMEM accumulators - the main q0*k0 (p0*v0) term and the five correction terms - and
// the per-tile P V result is added to the running output in fp32 REGISTERS (round-to-nearest) instead of
// being accumulated across tiles by the MMA.  TMEM: S 2 groups x (main, corr) x 64 + O likewise = 512 columns.
// grid = (ceil(max nq/128), heads, nprob).
#pragma once
#include "attn_tc.cuh"

namespace b2s {

constexpr int A3_BQ = 128, A3_BK = 64, A3_D = 64, A3_NP = 3, A3_SLOTS = 3;
constexpr int A3_QPL = A3_BQ * A3_D * 2;             // 16 KB: one Q plane  [128 x 64] bf16
constexpr int A3_KPL = A3_BK * A3_D * 2;             //  8 KB: one K / V plane [64 x 64] bf16
constexpr int A3_SLOT = A3_NP * A3_KPL;              // 24 KB
constexpr int A3_PPL = A3_BQ * A3_BK * 2;            // 16 KB: one P plane [128 x 64] bf16
constexpr int A3_SMEM = A3_NP * A3_QPL + A3_SLOTS * A3_SLOT + 2 * A3_NP * A3_PPL + 1024 + 256;
constexpr int A3_THREADS = 384;

struct Attn3Params {
  AttnTcProb prob[2];
  int qcol, kcol, vcol;            // column offsets of Q / K / V (head h adds h*64)
  int plane_rows;                  // rows between operand planes in the q/k/v buffer
  float scale_log2e;
  __nv_bfloat16* out; int ldo; size_t out_plane;   // ctx planes [3][2*cap, 256] bf16
  const int* ctrl; int cross;
};

// P chunk c (32 keys) of one row -> three swizzled K-major planes; returns the partial row sum (fp32 values)
template <bool MASK>
__device__ __forceinline__ float a3_write_p_chunk(const uint32_t (&v)[32], int c, uint32_t prow_addr, int r, int limit, float scale,
                                                  float m_used) {
  float sum0 = 0.f, sum1 = 0.f;
  uint32_t pk[A3_NP][16];
#pragma unroll
  for (int t = 0; t < 32; t += 2) {
    float p0 = ex2_approx(fmaf(__uint_as_float(v[t]), scale, -m_used));
    float p1 = ex2_approx(fmaf(__uint_as_float(v[t + 1]), scale, -m_used));
    if (MASK) {
      if (c * 32 + t >= limit) p0 = 0.f;
      if (c * 32 + t + 1 >= limit) p1 = 0.f;
    }
    sum0 += p0; sum1 += p1;
    uint32_t w[A3_NP];
    tc::pack_planes2<A3_NP>(p0, p1, w);
#pragma unroll
    for (int pl = 0; pl < A3_NP; ++pl) pk[pl][t >> 1] = w[pl];
  }
  // 32 keys = 64 B = four 16-byte chunks: chunk (c * 4 + q) ^ (r & 7) of the 128-byte row
#pragma unroll
  for (int pl = 0; pl < A3_NP; ++pl)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int chunk = ((c * 4 + q) ^ (r & 7));
      st_shared_v4(prow_addr + pl * A3_PPL + chunk * 16, pk[pl][4 * q], pk[pl][4 * q + 1], pk[pl][4 * q + 2], pk[pl][4 * q + 3]);
    }
  return sum0 + sum1;
}

__global__ void __launch_bounds__(A3_THREADS, 1) k_attn_tc3(const __grid_constant__ CUtensorMap mapQ,
                                                            const __grid_constant__ CUtensorMap mapKV, Attn3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                               // [3 planes]
  uint8_t* sR = sQ + A3_NP * A3_QPL;                // ring [A3_SLOTS][3 planes]
  uint8_t* sP = sR + A3_SLOTS * A3_SLOT;            // [2 groups][3 planes]; reused for the final merge
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * A3_NP * A3_PPL);
  uint64_t* q_full = bars;
  uint64_t* r_full = bars + 1;                 uint64_t* r_empty = r_full + A3_SLOTS;
  uint64_t* s_full = r_empty + A3_SLOTS;       uint64_t* s_free = s_full + 2;
  uint64_t* p_full = s_free + 2;               uint64_t* o_full = p_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  AttnTcProb pr = p.prob[blockIdx.z];
  const int q0 = blockIdx.x * A3_BQ;
  if (!p.ctrl && q0 >= pr.nq) return;                    // uniform per CTA (static sizes)
  const int h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 8 && lane == 0) { tc::tma_prefetch_desc(&mapQ); tc::tma_prefetch_desc(&mapKV); }
  if (warp == 9 && lane == 0) {
    tc::mbar_init(q_full, 1);
    for (int b = 0; b < A3_SLOTS; ++b) { tc::mbar_init(&r_full[b], 1); tc::mbar_init(&r_empty[b], 1); }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&s_full[b], 1); tc::mbar_init(&s_free[b], 128);
      tc::mbar_init(&p_full[b], 128); tc::mbar_init(&o_full[b], 1);
    }
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
    const bool on = !p.ctrl[1] && p.ctrl[2] > 0 && p.ctrl[3] > 0;
    pr.nq = on ? p.ctrl[2 + blockIdx.z] : 0;
    pr.nk = p.ctrl[2 + (p.cross ? 1 - blockIdx.z : blockIdx.z)];
  }
  const bool live = q0 < pr.nq;                          // uniform per CTA
  const int nt = live ? (pr.nk + A3_BK - 1) / A3_BK : 0;
  const uint32_t tS = tmem_base, tO = tmem_base + 256;    // S[g] = tS + 128 g + {0 main, 64 corr} ; O[g] likewise

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
      int c = 0;                                           // ring position, in consumption order
      auto load = [&](int col, int j) {
        const int b = c % A3_SLOTS, ph = (c / A3_SLOTS) & 1;
        tc::mbar_wait(&r_empty[b], ph ^ 1);
        tc::mbar_expect_tx(&r_full[b], A3_SLOT);
#pragma unroll
        for (int pl = 0; pl < A3_NP; ++pl)
          tc::tma_load_2d(sR + b * A3_SLOT + pl * A3_KPL, &mapKV, &r_full[b], col + h * 64, pl * p.plane_rows + pr.k_row + j * A3_BK);
        ++c;
      };
      load(p.kcol, 0);
      if (nt > 1) load(p.kcol, 1);
      for (int j = 0; j < nt; ++j) {
        if (j + 2 < nt) load(p.kcol, j + 2);
        load(p.vcol, j);
      }
    }
  } else if (warp == 9) {
    if (tc::elect_one()) {
      using Terms = tc::PlaneTerms<A3_NP>;
      constexpr uint32_t idesc_s = tc::idesc_bf16(128, 64, 0, 0);    // S: A = Q (K-major), B = K tile (K-major)
      constexpr uint32_t idesc_o = tc::idesc_bf16(128, 64, 0, 1);    // PV: A = P (K-major), B = V (MN-major)
      const uint32_t q_addr = tc::smem_u32(sQ);
      int c = 0;
      auto issue_s = [&](int j) {
        const int b = c % A3_SLOTS, ph = (c / A3_SLOTS) & 1, g = j & 1;
        ++c;
        tc::mbar_wait(&r_full[b], ph);
        tc::tc_fence_after();
        const uint32_t k_addr = tc::smem_u32(sR + b * A3_SLOT);
#pragma unroll
        for (int t = 0; t < Terms::N; ++t)
#pragma unroll
          for (int k = 0; k < A3_D / 16; ++k) {
            const uint64_t ad = tc::smem_desc_sw128(q_addr + Terms::a(t) * A3_QPL + k * 32, 16, 1024);
            const uint64_t bd = tc::smem_desc_sw128(k_addr + Terms::b(t) * A3_KPL + k * 32, 16, 1024);
            const bool main_term = t == Terms::N - 1;
            tc::umma_bf16(tS + g * 128 + (main_term ? 0 : 64), ad, bd, idesc_s, main_term ? (k ? 1u : 0u) : ((t | k) ? 1u : 0u));
          }
        tc::umma_commit(&s_full[g]);
        tc::umma_commit(&r_empty[b]);
      };
      auto issue_pv = [&](int j) {
        const int b = c % A3_SLOTS, ph = (c / A3_SLOTS) & 1, g = j & 1;
        ++c;
        tc::mbar_wait(&r_full[b], ph);
        tc::tc_fence_after();
        const uint32_t p_addr = tc::smem_u32(sP + g * A3_NP * A3_PPL), v_addr = tc::smem_u32(sR + b * A3_SLOT);
#pragma unroll
        for (int t = 0; t < Terms::N; ++t)
#pragma unroll
          for (int kk = 0; kk < A3_BK / 16; ++kk) {
            // P plane: [128 x 64] K-major; V plane: [64 keys x 64 d] MN-major, 16 keys = 2 swizzle atoms = 2048 B
            const uint64_t ad = tc::smem_desc_sw128(p_addr + Terms::a(t) * A3_PPL + kk * 32, 16, 1024);
            const uint64_t bd = tc::smem_desc_sw128(v_addr + Terms::b(t) * A3_KPL + kk * 2048, 16, 1024);
            const bool main_term = t == Terms::N - 1;      // per-tile result: both accumulators start afresh
            tc::umma_bf16(tO + g * 128 + (main_term ? 0 : 64), ad, bd, idesc_o, main_term ? (kk ? 1u : 0u) : ((t | kk) ? 1u : 0u));
          }
        tc::umma_commit(&o_full[g]);
        tc::umma_commit(&r_empty[b]);
      };
      tc::mbar_wait(q_full, 0);
      issue_s(0);
      if (nt > 1) issue_s(1);
      for (int j = 0; j < nt; ++j) {
        const int g = j & 1, ph = (j >> 1) & 1;
        if (j + 2 < nt) {                     // the group has S_j in registers: its next score tile can start
          tc::mbar_wait(&s_free[g], ph);
          issue_s(j + 2);
        }
        tc::mbar_wait(&p_full[g], ph);        // P_j written (and O[g] rescaled if the reference max moved)
        issue_pv(j);
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
    const uint32_t prow_addr = tc::smem_u32(sP + g * A3_NP * A3_PPL + r * 128);
    const uint32_t s_addr = tS + g * 128 + lane_addr, o_addr = tO + g * 128 + lane_addr;
    // o += (main + corr) of the PV result sitting in TMEM (exact fp32 adds)
    auto add_pv = [&]() {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t a[32], b[32];
        tc::tmem_ld32(o_addr + 32 * c, a); tc::tmem_ld32(o_addr + 64 + 32 * c, b);
        tc::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) o[32 * c + e] += __uint_as_float(a[e]) + __uint_as_float(b[e]);
      }
    };
    int t = 0;                                            // index among this group's tiles
    for (int j = g; j < nt; j += 2, ++t) {
      tc::mbar_wait(&s_full[g], t & 1);
      tc::tc_fence_after();
      uint32_t s[2][32];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t b[32];
        tc::tmem_ld32(s_addr + 32 * c, s[c]); tc::tmem_ld32(s_addr + 64 + 32 * c, b);
        tc::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) s[c][e] = __float_as_uint(__uint_as_float(s[c][e]) + __uint_as_float(b[e]));
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&s_free[g]);                        // S[g] is in registers: next QK^T of this group may overwrite it
      const int limit = pr.nk - j * A3_BK;
      const bool full = limit >= A3_BK;
      float mx;
      if (full) mx = fmaxf(atc_chunk_max<false>(s[0], 0, limit), atc_chunk_max<false>(s[1], 1, limit));
      else mx = fmaxf(atc_chunk_max<true>(s[0], 0, limit), atc_chunk_max<true>(s[1], 1, limit));
      const float mxs = mx * p.scale_log2e;
      if (t == 0) {
        m_used = mxs;
      } else {
        // PV of the group's previous tile must have retired before P[g] is rewritten; collect its result
        tc::mbar_wait(&o_full[g], (t - 1) & 1);
        tc::tc_fence_after();
        add_pv();
        if (__any_sync(0xffffffffu, mxs > m_used + ATC_LAZY)) {   // lazy: raise the reference max only beyond 2^8
          const float m_new = fmaxf(m_used, mxs);
          const float corr = ex2_approx(m_used - m_new);  // 1 for rows whose reference did not move
          l_run *= corr; m_used = m_new;
#pragma unroll
          for (int d = 0; d < A3_D; ++d) o[d] *= corr;
        }
      }
      float sum;
      if (full) sum = a3_write_p_chunk<false>(s[0], 0, prow_addr, r, limit, p.scale_log2e, m_used) +
                      a3_write_p_chunk<false>(s[1], 1, prow_addr, r, limit, p.scale_log2e, m_used);
      else sum = a3_write_p_chunk<true>(s[0], 0, prow_addr, r, limit, p.scale_log2e, m_used) +
                 a3_write_p_chunk<true>(s[1], 1, prow_addr, r, limit, p.scale_log2e, m_used);
      l_run += sum;
      tc::fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc::tc_fence_before();            // order our tcgen05.ld before the MMAs that follow the arrive
      tc::mbar_arrive(&p_full[g]);
    }
    // ---- the group's last PV ----
    if (t > 0) {
      tc::mbar_wait(&o_full[g], (t - 1) & 1);
      tc::tc_fence_after();
      add_pv();
    }
    const float m_run = m_used;
    tc::tc_fence_before();
    // ---- merge the two groups' partial softmax states (all MMAs that read sP have completed) ----
    float* mrg = reinterpret_cast<float*>(sP);            // [66][128] floats: O^T (64 rows), m, l
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (g == 1) {
#pragma unroll
      for (int d = 0; d < A3_D; ++d) mrg[d * 128 + r] = o[d];
      mrg[64 * 128 + r] = m_run; mrg[65 * 128 + r] = l_run;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (g == 0 && q0 + r < pr.nq) {
      const float m_b = mrg[64 * 128 + r], l_b = mrg[65 * 128 + r];
      const float m = fmaxf(m_run, m_b);
      const float ca = ex2_approx(m_run - m), cb = (m_b == -INFINITY) ? 0.f : ex2_approx(m_b - m);
      const float inv = 1.f / (l_run * ca + l_b * cb);
      const float fa = ca * inv, fb = cb * inv;
      __nv_bfloat16* dst = p.out + (size_t)(pr.q_row + q0 + r) * p.ldo + h * 64;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint32_t w[A3_NP][4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int d = 8 * q + 2 * e;
          uint32_t pw[A3_NP];
          tc::pack_planes2<A3_NP>(o[d] * fa + mrg[d * 128 + r] * fb, o[d + 1] * fa + mrg[(d + 1) * 128 + r] * fb, pw);
#pragma unroll
          for (int pl = 0; pl < A3_NP; ++pl) w[pl][e] = pw[pl];
        }
#pragma unroll
        for (int pl = 0; pl < A3_NP; ++pl)
          reinterpret_cast<uint4*>(dst + pl * p.out_plane)[q] = make_uint4(w[pl][0], w[pl][1], w[pl][2], w[pl][3]);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 10) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace b2s
