// gemm_simt.cuh - fp32 FFMA GEMM with fused epilogues for ALIKED's small contractions (block4's deformable-conv
// im2col products, 1x1 aggregations) that are too small for a tensor-core tile.   C[M,N] = epi( [A1 | A2][M,K] * W[N,K]^T )
//   * A may be the K-concatenation of two row-major sources.
//   * grid.z selects one of two row segments.
//   * the live row count can come from device memory (m_dev) so data-dependent sizes
//     (number of detected keypoints) need no host sync.
// Tile: BM x BN x 16, 256 threads, (BM/16)x(BN/16) register micro-tile split in two halves
// per dimension so shared-memory float4 reads are conflict free.
#pragma once
#include <algorithm>
#include "common.cuh"

namespace b2s {

enum Act { ACT_NONE = 0, ACT_SELU = 1, ACT_GELU = 2 };

struct GemmParams {
  const float* A1 = nullptr; int lda1 = 0; int K1 = 0;   // first K segment
  const float* A2 = nullptr; int lda2 = 0;               // second K segment (K - K1 cols)
  const float* W = nullptr;  int ldw = 0;                // [N,K] row-major
  float* C = nullptr;        int ldc = 0;
  int M = 0, N = 0, K = 0;
  // row segments (grid.z): segment z covers rows [seg_base[z], seg_base[z]+seg_rows[z])
  int nseg = 1;
  int seg_base[2] = {0, 0};
  int seg_rows[2] = {0, 0};
  const int* m_dev = nullptr; int m_mult = 1;            // if set: rows = min(M, *m_dev * m_mult)
  // epilogue: v = (acc + bias[n]) * alpha ; v += residual[m,n] ; act ; clamp
  const float* bias = nullptr;
  float alpha = 1.f;
  const float* residual = nullptr; int ldr = 0;
  int act = ACT_NONE;
  float clamp = 0.f;                                     // >0: clamp to [-clamp, clamp]
  // rotary (LightGlue SelfBlock): columns < rot_cols are rotated pairwise with per-row
  // cos/sin tables [rows, 32] (pair index = (col % 64) / 2)
  int rot_cols = 0;
  const float* rot_cos = nullptr;
  const float* rot_sin = nullptr;
  // deterministic split-K for small-M/N problems (single segment only): blockIdx.z = K slice,
  // raw partial sums go to splitk_ws [splitk][M][N] and k_gemm_splitk_epilogue reduces them
  // in a fixed order (no atomics -> bitwise reproducible) before applying the epilogue.
  float* splitk_ws = nullptr; size_t splitk_ws_floats = 0;
  int splitk = 1;
};

template <int BM, int BN>
__global__ void __launch_bounds__(256, (BM >= 128 ? 2 : 3)) k_gemm_simt(GemmParams p) {
  pdl_wait();
  constexpr int BK = 16;
  constexpr int TM = BM / 16, TN = BN / 16;   // micro tile (8 or 4)
  constexpr int HM = TM / 2, HN = TN / 2;     // half tiles
  constexpr int PAD = 4;
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Ws[2][BK][BN + PAD];

  const int z = p.splitk > 1 ? 0 : blockIdx.z;
  int rows = p.nseg > 1 ? p.seg_rows[z] : p.M;
  const int base = p.nseg > 1 ? p.seg_base[z] : 0;
  if (p.m_dev) rows = min(rows, *p.m_dev * p.m_mult);
  const int m0 = blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  if (m0 >= rows) return;

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  // loader mapping: each thread moves float4s along K
  constexpr int A_F4 = BM * BK / 4 / 256;  // float4 per thread for A (2 for BM=128, 1 for 64)
  constexpr int W_F4 = BN * BK / 4 / 256;
  float4 ra[A_F4], rw[W_F4];

  auto load_tiles = [&](int k0) {
    const float* Ab; int lda; int kk;
    if (k0 < p.K1) { Ab = p.A1; lda = p.lda1; kk = k0; }
    else { Ab = p.A2; lda = p.lda2; kk = k0 - p.K1; }
#pragma unroll
    for (int i = 0; i < A_F4; ++i) {
      int f = tid + i * 256;
      int r = f >> 2, kq = f & 3;
      int gr = m0 + r;
      ra[i] = gr < rows ? *reinterpret_cast<const float4*>(Ab + (size_t)(base + gr) * lda + kk + kq * 4)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < W_F4; ++i) {
      int f = tid + i * 256;
      int r = f >> 2, kq = f & 3;
      int gn = n0 + r;
      rw[i] = gn < p.N ? *reinterpret_cast<const float4*>(p.W + (size_t)gn * p.ldw + k0 + kq * 4)
                       : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_F4; ++i) {
      int f = tid + i * 256;
      int r = f >> 2, kq = f & 3;
      As[buf][kq * 4 + 0][r] = ra[i].x; As[buf][kq * 4 + 1][r] = ra[i].y;
      As[buf][kq * 4 + 2][r] = ra[i].z; As[buf][kq * 4 + 3][r] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < W_F4; ++i) {
      int f = tid + i * 256;
      int r = f >> 2, kq = f & 3;
      Ws[buf][kq * 4 + 0][r] = rw[i].x; Ws[buf][kq * 4 + 1][r] = rw[i].y;
      Ws[buf][kq * 4 + 2][r] = rw[i].z; Ws[buf][kq * 4 + 3][r] = rw[i].w;
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  int kt_begin = 0, nk = p.K / BK;
  if (p.splitk > 1) {
    const int per = (nk + p.splitk - 1) / p.splitk;
    kt_begin = blockIdx.z * per;
    nk = min(nk, kt_begin + per);
  }
  if (kt_begin < nk) {
    load_tiles(kt_begin * BK);
    store_tiles(kt_begin & 1);
  }
  __syncthreads();
  for (int kt = kt_begin; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], w[TN];
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < HM; ++i) a[h * HM + i] = As[buf][k][h * (BM / 2) + ty * HM + i];
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < HN; ++j) w[h * HN + j] = Ws[buf][k][h * (BN / 2) + tx * HN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int r = m0 + (i / HM) * (BM / 2) + ty * HM + (i % HM);
    if (r >= rows) continue;
    const size_t grow = (size_t)(base + r);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int c = n0 + (j / HN) * (BN / 2) + tx * HN + (j % HN);
      if (c >= p.N) continue;
      float v = acc[i][j];
      if (p.splitk > 1) {   // raw partial sum; epilogue applied after the ordered reduction
        p.splitk_ws[((size_t)blockIdx.z * p.M + r) * p.N + c] = v;
        continue;
      }
      if (p.bias) v += p.bias[c];
      v *= p.alpha;
      if (c < p.rot_cols) {
        // partner column lives in the same thread (HN is even, columns start even)
        const int jp = j ^ 1;
        float vp = acc[i][jp];
        const int cp = c ^ 1;
        if (p.bias) vp += p.bias[cp];
        vp *= p.alpha;
        const int fi = (c & 63) >> 1;
        const float cs = p.rot_cos[grow * 32 + fi], sn = p.rot_sin[grow * 32 + fi];
        // out[2i] = x[2i]*cos - x[2i+1]*sin ; out[2i+1] = x[2i+1]*cos + x[2i]*sin
        v = (c & 1) ? (v * cs + vp * sn) : (v * cs - vp * sn);
      }
      if (p.residual) v += p.residual[grow * p.ldr + c];
      if (p.act == ACT_SELU) v = selu_f(v);
      else if (p.act == ACT_GELU) v = gelu_erf_f(v);
      if (p.clamp > 0.f) v = fminf(fmaxf(v, -p.clamp), p.clamp);
      p.C[grow * p.ldc + c] = v;
    }
  }
}

// ordered reduction of the split-K partials + the usual epilogue (bias, alpha, residual, act, clamp)
static __global__ void __launch_bounds__(256) k_gemm_splitk_epilogue(GemmParams p) {
  pdl_wait();
  int rows = p.M;
  if (p.m_dev) rows = min(rows, *p.m_dev * p.m_mult);
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)rows * p.N) return;
  const int r = (int)(e / p.N), c = (int)(e % p.N);
  float v = 0.f;
  for (int z = 0; z < p.splitk; ++z) v += p.splitk_ws[((size_t)z * p.M + r) * p.N + c];
  if (p.bias) v += p.bias[c];
  v *= p.alpha;
  if (p.residual) v += p.residual[(size_t)r * p.ldr + c];
  if (p.act == ACT_SELU) v = selu_f(v);
  else if (p.act == ACT_GELU) v = gelu_erf_f(v);
  if (p.clamp > 0.f) v = fminf(fmaxf(v, -p.clamp), p.clamp);
  p.C[(size_t)r * p.ldc + c] = v;
}

// Host launcher: picks the tile that best fills 148 SMs.
inline int gemm_simt(const GemmParams& p, cudaStream_t st, long long* launches, KernelProf* prof = nullptr) {
  if (p.K % 16 != 0 || (p.K1 % 16) != 0) {
    set_error("gemm_simt: K=%d / K1=%d must be multiples of 16", p.K, p.K1);
    return B2S_EINVAL;
  }
  GemmParams q = p;
  if (q.A2 == nullptr) q.K1 = q.K;
  int maxrows = q.M;
  if (q.nseg > 1) maxrows = q.seg_rows[0] > q.seg_rows[1] ? q.seg_rows[0] : q.seg_rows[1];
  if (maxrows <= 0 || q.N <= 0) return 0;
  if (prof) prof->mark(PROF_GEMM, st);
  const long long big = (long long)cdiv(maxrows, 128) * cdiv(q.N, 128) * q.nseg;
  const long long small = (long long)cdiv(maxrows, 64) * cdiv(q.N, 64);
  q.splitk = 1;
  if (q.nseg == 1 && q.splitk_ws && q.rot_cols == 0 && small < 100 && q.K >= 256) {
    int sk = (int)std::min<long long>((148 * 2 + small - 1) / small, q.K / 64);   // >= 4 k-tiles of 16 per slice
    while (sk > 1 && (size_t)sk * q.M * q.N > q.splitk_ws_floats) --sk;
    if (sk > 1) {
      q.splitk = sk;
      dim3 g(cdiv(q.N, 64), cdiv(maxrows, 64), sk);
      launch_k(k_gemm_simt<64, 64>, g, 256, 0, st, q);
      const size_t total = (size_t)maxrows * q.N;
      launch_k(k_gemm_splitk_epilogue, (unsigned)((total + 255) / 256), 256, 0, st, q);
      if (prof) prof->mark(PROF_GEMM, st);
      if (launches) *launches += 2;
      B2S_LAUNCH_CHECK();
      return 0;
    }
  }
  if (big >= 148 && q.N >= 128) {
    dim3 g(cdiv(q.N, 128), cdiv(maxrows, 128), q.nseg);
    launch_k(k_gemm_simt<128, 128>, g, 256, 0, st, q);
  } else {
    dim3 g(cdiv(q.N, 64), cdiv(maxrows, 64), q.nseg);
    launch_k(k_gemm_simt<64, 64>, g, 256, 0, st, q);
  }
  if (prof) prof->mark(PROF_GEMM, st);
  if (launches) ++*launches;
  B2S_LAUNCH_CHECK();
  return 0;
}

}  // namespace b2s
