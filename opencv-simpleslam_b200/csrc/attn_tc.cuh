// attn_tc.cuh - flash-style attention on tcgen05/TMEM (bf16 operands, fp32 accumulate/softmax).
//   O = softmax(Q K^T * scale) V,  head_dim 64, 128-query tile per CTA, 128-key tiles.
// Warp roles (384 threads = 3 warpgroups): warps 0..3 and 4..7 = two softmax groups, warp 8 = TMA
// producer (Q once, K/V 3-stage rings), warp 9 = MMA issuer; the producer warpgroup gives its
// registers to the softmax warpgroups (setmaxnreg).  Group g owns the key tiles with j & 1 == g
// and its own S / P / O buffers, so two tiles are in flight on the CUDA cores while the tensor
// pipe works on the next S / PV.  Per tile a softmax thread (== query row):
//   1. tcgen05.ld the whole S row (128 fp32) into registers and hands S[g] back at once
//      (s_free) - the MMA warp computes this group's next score tile while the exps run;
//   2. row max; the running reference max m_used is only raised when the row max exceeds it by
//      more than 2^8 (lazy rescale, decided per warp): exps stay <= 256, the final O / l is
//      unchanged in real arithmetic, and the O accumulator - kept in TMEM and accumulated by the
//      PV MMAs themselves - is rescaled (tcgen05.ld / st) only on those rare tiles;
//   3. P = 2^(s*scale - m_used) -> bf16 -> 128B-swizzled K-major smem tile that the PV MMA reads;
//      V is consumed MN-major straight from its row-major [key][d] tile.
// The two groups' partial (m, l, O) are merged once at the end through shared memory.
// grid = (ceil(max nq/128), heads, nprob).  TMEM: S 2x128 + O 2x64 columns (512 allocated).
#pragma once
#include "tc_common.cuh"

namespace b2s {

struct AttnTcProb { int q_row, k_row, nq, nk; };   // row bases inside the [segments * cap, ld] bf16 buffer
// grid.z = problem.  With the LightGlue device state (`ctrl`, 32 ints per pair) problem z is SEGMENT z = 2 * pair + image
// (lightglue_kernels.cuh): queries = segment z, keys = the same segment (self) or the pair's other image (cross), sizes
// from the pair's state.  Without it (unit tests) the problems are the static prob[0..1].
struct AttnTcParams {
  AttnTcProb prob[2]; int cap;
  int qcol, kcol, vcol;            // column offsets of Q / K / V (head h adds h*64)
  float scale_log2e;               // softmax scale * log2(e)
  __nv_bfloat16* out; int ldo;     // ctx [2*cap, 256] bf16, rows aligned with q_row
  const int* ctrl; int cross;      // LightGlue device state (nullable): nq / nk = ctrl[2 + z] (cross: nk = ctrl[3 - z])
  unsigned long long* stats;       // nullable: stats[cross] += nq * nk of every live problem (executed work, for the roofline)
};

constexpr int ATC_BQ = 128, ATC_BK = 128, ATC_D = 64, ATC_KS = 3;
constexpr int ATC_TILE = 128 * 64 * 2;               // 16 KB: one [128 x 64] bf16 tile
constexpr int ATC_SMEM = ATC_TILE /*Q*/ + 2 * ATC_KS * ATC_TILE /*K,V*/ + 2 * 2 * ATC_TILE /*P*/ + 1024 + 256;
constexpr int ATC_THREADS = 384;
constexpr float ATC_LAZY = 8.f;                      // raise the reference max only beyond 2^8

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// P chunk c (32 keys) of one row: 2^(s*scale - m) -> bf16 -> swizzled K-major smem; returns the partial row sum
template <bool MASK>
__device__ __forceinline__ float atc_write_p_chunk(const uint32_t (&v)[32], int c, uint32_t prow_addr, int r, int limit, float scale,
                                                   float m_used) {
  float sum0 = 0.f, sum1 = 0.f;
  uint32_t pk[16];
#pragma unroll
  for (int t = 0; t < 32; t += 2) {
    float p0 = ex2_approx(fmaf(__uint_as_float(v[t]), scale, -m_used));
    float p1 = ex2_approx(fmaf(__uint_as_float(v[t + 1]), scale, -m_used));
    if (MASK) {
      if (c * 32 + t >= limit) p0 = 0.f;
      if (c * 32 + t + 1 >= limit) p1 = 0.f;
    }
    sum0 += p0; sum1 += p1;
    const __nv_bfloat162 hb = __floats2bfloat162_rn(p0, p1);
    pk[t >> 1] = *reinterpret_cast<const uint32_t*>(&hb);
  }
  // 32 keys = 64 B = four 16-byte chunks of block (c >> 1); chunk ((c & 1) * 4 + q) ^ (r & 7)
  const uint32_t blk = prow_addr + (c >> 1) * ATC_TILE;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int chunk = (((c & 1) * 4 + q) ^ (r & 7));
    st_shared_v4(blk + chunk * 16, pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
  }
  return sum0 + sum1;
}

template <bool MASK>
__device__ __forceinline__ float atc_chunk_max(const uint32_t (&v)[32], int c, int limit) {
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < 32; ++t)
    if (!MASK || c * 32 + t < limit) mx = fmaxf(mx, __uint_as_float(v[t]));
  return mx;
}

__global__ void __launch_bounds__(ATC_THREADS, 1) k_attn_tc(const __grid_constant__ CUtensorMap mapQKV, AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // aligned by pointer ARITHMETIC on the __shared__ array: an integer round trip would make every staging access a generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATC_TILE;                 // [ATC_KS]
  uint8_t* sV = sK + ATC_KS * ATC_TILE;        // [ATC_KS]
  uint8_t* sP = sV + ATC_KS * ATC_TILE;        // [2 groups][2 blocks of 64 keys]; reused for the final merge
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * ATC_TILE);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;                 uint64_t* k_empty = k_full + ATC_KS;
  uint64_t* v_full = k_empty + ATC_KS;         uint64_t* v_empty = v_full + ATC_KS;
  uint64_t* s_full = v_empty + ATC_KS;         uint64_t* s_free = s_full + 2;
  uint64_t* p_full = s_free + 2;               uint64_t* o_full = p_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  AttnTcProb pr = p.prob[p.ctrl ? 0 : blockIdx.z];
  const int q0 = blockIdx.x * ATC_BQ;
  if (!p.ctrl && q0 >= pr.nq) return;                    // uniform per CTA (static sizes)
  const int h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 8 && lane == 0) tc::tma_prefetch_desc(&mapQKV);
  if (warp == 9 && lane == 0) {
    tc::mbar_init(q_full, 1);
    for (int b = 0; b < ATC_KS; ++b) {
      tc::mbar_init(&k_full[b], 1); tc::mbar_init(&k_empty[b], 1);
      tc::mbar_init(&v_full[b], 1); tc::mbar_init(&v_empty[b], 1);
    }
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
  pdl_wait();            // everything above overlaps the previous kernel's tail; Q/K/V are touched only below
  if (p.ctrl) {          // device-resident sizes (pruning / early exit)
    const int z = blockIdx.z, side = z & 1;
    const int* c = p.ctrl + (z >> 1) * 32;
    const bool on = !c[1] && c[2] > 0 && c[3] > 0;
    pr.q_row = z * p.cap; pr.k_row = (p.cross ? z ^ 1 : z) * p.cap;
    pr.nq = on ? c[2 + side] : 0;
    pr.nk = c[2 + (p.cross ? 1 - side : side)];
  }
  const bool live = q0 < pr.nq;                          // uniform per CTA
  const int nt = live ? (pr.nk + ATC_BK - 1) / ATC_BK : 0;
  if (p.stats && live && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)
    atomicAdd(&p.stats[p.cross ? 1 : 0], (unsigned long long)pr.nq * (unsigned long long)pr.nk);
  const uint32_t tS = tmem_base, tO = tmem_base + 256;    // S[g] = tS + 128 g ; O[g] = tO + 64 g

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  }
  if (!live) {
    // nothing to do for this query tile
  } else if (warp == 8) {
    if (tc::elect_one()) {
      tc::mbar_expect_tx(q_full, ATC_TILE);
      tc::tma_load_2d(sQ, &mapQKV, q_full, p.qcol + h * 64, pr.q_row + q0);
      for (int j = 0; j < nt; ++j) {
        const int b = j % ATC_KS, ph = (j / ATC_KS) & 1;
        tc::mbar_wait(&k_empty[b], ph ^ 1);
        tc::mbar_expect_tx(&k_full[b], ATC_TILE);
        tc::tma_load_2d(sK + b * ATC_TILE, &mapQKV, &k_full[b], p.kcol + h * 64, pr.k_row + j * ATC_BK);
        tc::mbar_wait(&v_empty[b], ph ^ 1);
        tc::mbar_expect_tx(&v_full[b], ATC_TILE);
        tc::tma_load_2d(sV + b * ATC_TILE, &mapQKV, &v_full[b], p.vcol + h * 64, pr.k_row + j * ATC_BK);
      }
    }
  } else if (warp == 9) {
    if (tc::elect_one()) {
      constexpr uint32_t idesc_s = tc::idesc_bf16(128, 128, 0, 0);   // S: A = Q (K-major), B = K tile (K-major)
      constexpr uint32_t idesc_o = tc::idesc_bf16(128, 64, 0, 1);    // PV: A = P (K-major), B = V (MN-major)
      const uint32_t q_addr = tc::smem_u32(sQ);
      auto issue_s = [&](int j) {
        const int b = j % ATC_KS, ph = (j / ATC_KS) & 1, g = j & 1;
        tc::mbar_wait(&k_full[b], ph);
        tc::tc_fence_after();
        const uint32_t k_addr = tc::smem_u32(sK + b * ATC_TILE);
#pragma unroll
        for (int k = 0; k < ATC_D / 16; ++k) {
          const uint64_t ad = tc::smem_desc_sw128(q_addr + k * 32, 16, 1024);
          const uint64_t bd = tc::smem_desc_sw128(k_addr + k * 32, 16, 1024);
          tc::umma_bf16(tS + g * 128, ad, bd, idesc_s, k ? 1u : 0u);
        }
        tc::umma_commit(&s_full[g]);
        tc::umma_commit(&k_empty[b]);
      };
      auto issue_pv = [&](int j) {
        const int b = j % ATC_KS, ph = (j / ATC_KS) & 1, g = j & 1;
        tc::mbar_wait(&v_full[b], ph);
        tc::tc_fence_after();
        const uint32_t p_addr = tc::smem_u32(sP + g * 2 * ATC_TILE), v_addr = tc::smem_u32(sV + b * ATC_TILE);
#pragma unroll
        for (int kk = 0; kk < ATC_BK / 16; ++kk) {
          // P: two [128 x 64] K-major blocks; V: [128 keys x 64 d], MN-major, 16 keys = 2 swizzle atoms = 2048 B
          const uint64_t ad = tc::smem_desc_sw128(p_addr + (kk >> 2) * ATC_TILE + (kk & 3) * 32, 16, 1024);
          const uint64_t bd = tc::smem_desc_sw128(v_addr + kk * 2048, 16, 1024);
          tc::umma_bf16(tO + g * 64, ad, bd, idesc_o, (j >= 2 || kk) ? 1u : 0u);   // O[g] accumulates over the group's tiles
        }
        tc::umma_commit(&o_full[g]);
        tc::umma_commit(&v_empty[b]);
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
    const uint32_t prow_addr = tc::smem_u32(sP + g * 2 * ATC_TILE + r * 128);
    const uint32_t s_addr = tS + g * 128 + lane_addr, o_addr = tO + g * 64 + lane_addr;
    int t = 0;                                            // index among this group's tiles
    for (int j = g; j < nt; j += 2, ++t) {
      tc::mbar_wait(&s_full[g], t & 1);
      tc::tc_fence_after();
      uint32_t s[4][32];
      tc::tmem_ld32(s_addr, s[0]); tc::tmem_ld32(s_addr + 32, s[1]);
      tc::tmem_ld32(s_addr + 64, s[2]); tc::tmem_ld32(s_addr + 96, s[3]);
      tc::tmem_ld_wait();
      tc::tc_fence_before();
      tc::mbar_arrive(&s_free[g]);                        // S[g] is in registers: next QK^T of this group may overwrite it
      const int limit = pr.nk - j * ATC_BK;
      const bool full = limit >= ATC_BK;
      float mx;
      if (full) mx = fmaxf(fmaxf(atc_chunk_max<false>(s[0], 0, limit), atc_chunk_max<false>(s[1], 1, limit)),
                           fmaxf(atc_chunk_max<false>(s[2], 2, limit), atc_chunk_max<false>(s[3], 3, limit)));
      else mx = fmaxf(fmaxf(atc_chunk_max<true>(s[0], 0, limit), atc_chunk_max<true>(s[1], 1, limit)),
                      fmaxf(atc_chunk_max<true>(s[2], 2, limit), atc_chunk_max<true>(s[3], 3, limit)));
      const float mxs = mx * p.scale_log2e;
      if (t == 0) {
        m_used = mxs;                                     // first PV of the group overwrites O[g]: nothing to rescale
      } else {
        // PV of the group's previous tile must have retired before P[g] is rewritten / O[g] is touched
        tc::mbar_wait(&o_full[g], (t - 1) & 1);
        tc::tc_fence_after();
        if (__any_sync(0xffffffffu, mxs > m_used + ATC_LAZY)) {
          const float m_new = fmaxf(m_used, mxs);
          const float corr = ex2_approx(m_used - m_new);  // 1 for rows whose reference did not move
          l_run *= corr; m_used = m_new;
          uint32_t o[2][32];
          tc::tmem_ld32(o_addr, o[0]); tc::tmem_ld32(o_addr + 32, o[1]);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int e = 0; e < 32; ++e) o[c][e] = __float_as_uint(__uint_as_float(o[c][e]) * corr);
          tc::tmem_st32(o_addr, o[0]); tc::tmem_st32(o_addr + 32, o[1]);
          tc::tmem_st_wait();
        }
      }
      float sum;
      if (full) sum = (atc_write_p_chunk<false>(s[0], 0, prow_addr, r, limit, p.scale_log2e, m_used) +
                       atc_write_p_chunk<false>(s[1], 1, prow_addr, r, limit, p.scale_log2e, m_used)) +
                      (atc_write_p_chunk<false>(s[2], 2, prow_addr, r, limit, p.scale_log2e, m_used) +
                       atc_write_p_chunk<false>(s[3], 3, prow_addr, r, limit, p.scale_log2e, m_used));
      else sum = (atc_write_p_chunk<true>(s[0], 0, prow_addr, r, limit, p.scale_log2e, m_used) +
                  atc_write_p_chunk<true>(s[1], 1, prow_addr, r, limit, p.scale_log2e, m_used)) +
                 (atc_write_p_chunk<true>(s[2], 2, prow_addr, r, limit, p.scale_log2e, m_used) +
                  atc_write_p_chunk<true>(s[3], 3, prow_addr, r, limit, p.scale_log2e, m_used));
      l_run += sum;
      tc::fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc::tc_fence_before();            // order our tcgen05.ld / st before the MMAs that follow the arrive
      tc::mbar_arrive(&p_full[g]);
    }
    // ---- this group's accumulator: O[g] after its last PV ----
    float o[ATC_D];
    if (t > 0) {
      tc::mbar_wait(&o_full[g], (t - 1) & 1);
      tc::tc_fence_after();
      uint32_t v[2][32];
      tc::tmem_ld32(o_addr, v[0]); tc::tmem_ld32(o_addr + 32, v[1]);
      tc::tmem_ld_wait();
#pragma unroll
      for (int d = 0; d < ATC_D; ++d) o[d] = __uint_as_float(v[d >> 5][d & 31]);
    } else {
#pragma unroll
      for (int d = 0; d < ATC_D; ++d) o[d] = 0.f;
    }
    const float m_run = m_used;
    tc::tc_fence_before();
    // ---- merge the two groups' partial softmax states (all MMAs that read sP have completed) ----
    float* mrg = reinterpret_cast<float*>(sP);            // [66][128] floats: O^T (64 rows), m, l
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (g == 1) {
#pragma unroll
      for (int d = 0; d < ATC_D; ++d) mrg[d * 128 + r] = o[d];
      mrg[64 * 128 + r] = m_run; mrg[65 * 128 + r] = l_run;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (g == 0 && q0 + r < pr.nq) {
      const float m_b = mrg[64 * 128 + r], l_b = mrg[65 * 128 + r];
      const float m = fmaxf(m_run, m_b);
      const float ca = ex2_approx(m_run - m), cb = (m_b == -INFINITY) ? 0.f : ex2_approx(m_b - m);
      const float inv = 1.f / (l_run * ca + l_b * cb);
      const float fa = ca * inv, fb = cb * inv;
      uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)(pr.q_row + q0 + r) * p.ldo + h * 64);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float e[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) e[t] = o[8 * q + t] * fa + mrg[(8 * q + t) * 128 + r] * fb;
        __nv_bfloat162 a = __floats2bfloat162_rn(e[0], e[1]), bb = __floats2bfloat162_rn(e[2], e[3]);
        __nv_bfloat162 c = __floats2bfloat162_rn(e[4], e[5]), d = __floats2bfloat162_rn(e[6], e[7]);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&bb);
        u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
        dst[q] = u;
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 10) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace b2s
