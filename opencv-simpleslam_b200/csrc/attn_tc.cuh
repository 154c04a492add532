// attn_tc.cuh - flash-style attention on tcgen05/TMEM (bf16 operands, fp32 accumulate/softmax).
//   O = softmax(Q K^T * scale) V,  head_dim 64, 128-query tile per CTA, 128-key tiles.
// Warp roles (320 threads): warp 0 = TMA producer (Q once, K/V double-buffered), warp 1 = MMA
// issuer (S_j = Q K_j^T into TMEM S[j&1]; PV_j = P_j V_j into TMEM O[j&1]), warps 2..5 and
// 6..9 = two softmax groups: group g owns the key tiles with j & 1 == g (its own S / P / O
// buffers and its own running max / sum / output in registers), so two tiles are in flight on
// the CUDA cores (2 warps per scheduler) while the tensor pipe works on the next S / PV; the
// two partial results are merged once at the end (log-sum-exp merge through shared memory).
// Softmax thread == query row: tcgen05.ld S, P -> bf16 -> 128B-swizzled K-major smem tile that
// the PV MMA reads; V is consumed MN-major straight from its row-major [key][d] tile.
// grid = (ceil(max nq/128), heads, nprob).  TMEM: S 2x128 + O 2x64 columns (512 allocated).
#pragma once
#include "tc_common.cuh"

namespace b2s {

struct AttnTcProb { int q_row, k_row, nq, nk; };   // row bases inside the [2*cap, ld] bf16 buffer
struct AttnTcParams {
  AttnTcProb prob[2];
  int qcol, kcol, vcol;            // column offsets of Q / K / V (head h adds h*64)
  float scale_log2e;               // softmax scale * log2(e)
  __nv_bfloat16* out; int ldo;     // ctx [2*cap, 256] bf16, rows aligned with q_row
};

constexpr int ATC_BQ = 128, ATC_BK = 128, ATC_D = 64;
constexpr int ATC_TILE = 128 * 64 * 2;               // 16 KB: one [128 x 64] bf16 tile
constexpr int ATC_SMEM = ATC_TILE /*Q*/ + 2 * ATC_TILE /*K*/ + 2 * ATC_TILE /*V*/ + 2 * 2 * ATC_TILE /*P*/ + 1024 + 256;
constexpr int ATC_THREADS = 320;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// pass 1 of a key tile: row maximum of the raw scores (keys >= limit masked when MASK)
template <bool MASK>
__device__ __forceinline__ float atc_row_max(uint32_t s_addr, int limit) {
  float mx = -INFINITY;
  uint32_t v[2][32];
  tc::tmem_ld32(s_addr, v[0]);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    tc::tmem_ld_wait();
    if (c < 3) tc::tmem_ld32(s_addr + (c + 1) * 32, v[(c + 1) & 1]);   // next chunk in flight while this one is reduced
#pragma unroll
    for (int t = 0; t < 32; ++t) {
      const float s = __uint_as_float(v[c & 1][t]);
      if (!MASK || c * 32 + t < limit) mx = fmaxf(mx, s);
    }
  }
  return mx;
}

// pass 2: P = 2^(s*scale - m) -> bf16 -> swizzled K-major smem tile; returns the row sum
template <bool MASK>
__device__ __forceinline__ float atc_write_p(uint32_t s_addr, uint32_t prow_addr, int r, int limit, float scale, float m_new) {
  float sum = 0.f;
  uint32_t v[2][32];
  tc::tmem_ld32(s_addr, v[0]);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    tc::tmem_ld_wait();
    if (c < 3) tc::tmem_ld32(s_addr + (c + 1) * 32, v[(c + 1) & 1]);
    uint32_t pk[16];
#pragma unroll
    for (int t = 0; t < 32; t += 2) {
      float p0 = ex2_approx(fmaf(__uint_as_float(v[c & 1][t]), scale, -m_new));
      float p1 = ex2_approx(fmaf(__uint_as_float(v[c & 1][t + 1]), scale, -m_new));
      if (MASK) {
        if (c * 32 + t >= limit) p0 = 0.f;
        if (c * 32 + t + 1 >= limit) p1 = 0.f;
      }
      sum += p0 + p1;
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
  }
  return sum;
}

__global__ void __launch_bounds__(ATC_THREADS, 1) k_attn_tc(const __grid_constant__ CUtensorMap mapQKV, AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATC_TILE;            // [2]
  uint8_t* sV = sK + 2 * ATC_TILE;        // [2]
  uint8_t* sP = sV + 2 * ATC_TILE;        // [2][2 blocks of 64 keys]; reused for the final merge
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * ATC_TILE);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;   uint64_t* k_empty = bars + 3;
  uint64_t* v_full = bars + 5;   uint64_t* v_empty = bars + 7;
  uint64_t* s_full = bars + 9;   uint64_t* p_full = bars + 11;  uint64_t* o_full = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const AttnTcProb pr = p.prob[blockIdx.z];
  const int q0 = blockIdx.x * ATC_BQ;
  if (q0 >= pr.nq) return;                               // uniform per CTA
  const int h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = (pr.nk + ATC_BK - 1) / ATC_BK;

  if (warp == 0 && lane == 0) tc::tma_prefetch_desc(&mapQKV);
  if (warp == 1 && lane == 0) {
    tc::mbar_init(q_full, 1);
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&k_full[b], 1); tc::mbar_init(&k_empty[b], 1);
      tc::mbar_init(&v_full[b], 1); tc::mbar_init(&v_empty[b], 1);
      tc::mbar_init(&s_full[b], 1); tc::mbar_init(&p_full[b], 128); tc::mbar_init(&o_full[b], 1);
    }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + 256;    // S[b] = tS + 128 b ; O[b] = tO + 64 b

  if (warp == 0) {
    if (tc::elect_one()) {
      tc::mbar_expect_tx(q_full, ATC_TILE);
      tc::tma_load_2d(sQ, &mapQKV, q_full, p.qcol + h * 64, pr.q_row + q0);
      for (int j = 0; j < nt; ++j) {
        const int b = j & 1, ph = (j >> 1) & 1;
        tc::mbar_wait(&k_empty[b], ph ^ 1);
        tc::mbar_expect_tx(&k_full[b], ATC_TILE);
        tc::tma_load_2d(sK + b * ATC_TILE, &mapQKV, &k_full[b], p.kcol + h * 64, pr.k_row + j * ATC_BK);
        tc::mbar_wait(&v_empty[b], ph ^ 1);
        tc::mbar_expect_tx(&v_full[b], ATC_TILE);
        tc::tma_load_2d(sV + b * ATC_TILE, &mapQKV, &v_full[b], p.vcol + h * 64, pr.k_row + j * ATC_BK);
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      constexpr uint32_t idesc_s = tc::idesc_bf16(128, 128, 0, 0);   // S: A = Q (K-major), B = K tile (K-major)
      constexpr uint32_t idesc_o = tc::idesc_bf16(128, 64, 0, 1);    // PV: A = P (K-major), B = V (MN-major)
      const uint32_t q_addr = tc::smem_u32(sQ);
      auto issue_s = [&](int j) {
        const int b = j & 1, ph = (j >> 1) & 1;
        tc::mbar_wait(&k_full[b], ph);
        tc::tc_fence_after();
        const uint32_t k_addr = tc::smem_u32(sK + b * ATC_TILE);
#pragma unroll
        for (int k = 0; k < ATC_D / 16; ++k) {
          const uint64_t ad = tc::smem_desc_sw128(q_addr + k * 32, 16, 1024);
          const uint64_t bd = tc::smem_desc_sw128(k_addr + k * 32, 16, 1024);
          tc::umma_bf16(tS + b * 128, ad, bd, idesc_s, k ? 1u : 0u);
        }
        tc::umma_commit(&s_full[b]);
        tc::umma_commit(&k_empty[b]);
      };
      auto issue_pv = [&](int i) {
        const int b = i & 1, ph = (i >> 1) & 1;
        tc::mbar_wait(&v_full[b], ph);
        tc::tc_fence_after();
        const uint32_t p_addr = tc::smem_u32(sP + b * 2 * ATC_TILE), v_addr = tc::smem_u32(sV + b * ATC_TILE);
#pragma unroll
        for (int kk = 0; kk < ATC_BK / 16; ++kk) {
          // P: two [128 x 64] K-major blocks; V: [128 keys x 64 d], MN-major, 16 keys = 2 swizzle atoms = 2048 B
          const uint64_t ad = tc::smem_desc_sw128(p_addr + (kk >> 2) * ATC_TILE + (kk & 3) * 32, 16, 1024);
          const uint64_t bd = tc::smem_desc_sw128(v_addr + kk * 2048, 16, 1024);
          tc::umma_bf16(tO + b * 64, ad, bd, idesc_o, kk ? 1u : 0u);
        }
        tc::umma_commit(&o_full[b]);
        tc::umma_commit(&v_empty[b]);
      };
      tc::mbar_wait(q_full, 0);
      issue_s(0);
      if (nt > 1) issue_s(1);
      for (int j = 0; j < nt; ++j) {
        // p_full(j): group (j&1) has finished reading S[j&1] and has written P[j&1]
        tc::mbar_wait(&p_full[j & 1], (j >> 1) & 1);
        if (j + 2 < nt) issue_s(j + 2);     // next score tile of that group first: it is on the group's critical path
        issue_pv(j);
      }
    }
  } else {
    // ===== softmax groups: g = 0 (warps 2..5) takes even key tiles, g = 1 (warps 6..9) odd ones =====
    const int g = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                       // query row of this thread
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    float o[ATC_D];
#pragma unroll
    for (int d = 0; d < ATC_D; ++d) o[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    const uint32_t prow_addr = tc::smem_u32(sP + g * 2 * ATC_TILE + r * 128);
    const uint32_t s_addr = tS + g * 128 + lane_addr, o_addr = tO + g * 64 + lane_addr;
    float corr_prev = 0.f;
    auto consume = [&](int jprev, float corr) {     // O = O * corr + PV(jprev)
      tc::mbar_wait(&o_full[g], (jprev >> 1) & 1);
      tc::tc_fence_after();
      uint32_t v[2][32];
      tc::tmem_ld32(o_addr, v[0]);
      tc::tmem_ld32(o_addr + 32, v[1]);
      tc::tmem_ld_wait();
#pragma unroll
      for (int t = 0; t < 64; ++t) o[t] = fmaf(o[t], corr, __uint_as_float(v[t >> 5][t & 31]));
    };
    for (int j = g; j < nt; j += 2) {
      tc::mbar_wait(&s_full[g], (j >> 1) & 1);
      tc::tc_fence_after();
      const int limit = pr.nk - j * ATC_BK;
      const bool full = limit >= ATC_BK;
      const float mx = full ? atc_row_max<false>(s_addr, limit) : atc_row_max<true>(s_addr, limit);
      const float m_new = fmaxf(m_run, mx * p.scale_log2e);
      const float corr = ex2_approx(m_run - m_new);
      // PV of this group's previous tile: once it is read back, P[g] and O[g] are free again
      if (j >= 2) consume(j - 2, corr_prev);
      const float sum = full ? atc_write_p<false>(s_addr, prow_addr, r, limit, p.scale_log2e, m_new)
                             : atc_write_p<true>(s_addr, prow_addr, r, limit, p.scale_log2e, m_new);
      l_run = l_run * corr + sum;
      m_run = m_new;
      corr_prev = corr;
      tc::fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc::tc_fence_before();            // order our tcgen05.ld of S / O before the MMAs that overwrite them
      tc::mbar_arrive(&p_full[g]);
    }
    if (nt > g) consume(((nt - 1 - g) & ~1) + g, corr_prev);   // this group's last tile
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
  if (warp == 2) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace b2s
