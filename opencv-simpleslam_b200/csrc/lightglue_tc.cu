// lightglue_tc.cu - bf16 tensor-core path of one LightGlue transformer layer:
// tcgen05/TMEM GEMMs (gemm_tc.cuh) + tcgen05 flash attention (attn_tc.cuh) + small
// bandwidth-bound glue (fp32->bf16, LayerNorm+GELU).  The residual stream stays fp32.
#include "lightglue_tc.cuh"
#include "attn_tc.cuh"
#include "gemm_tc.cuh"

#include <algorithm>
#include <vector>

namespace b2s {

int make_tmap_bf16_2d(CUtensorMap* out, const void* gptr, uint64_t inner, uint64_t rows, uint64_t row_pitch_bytes,
                      uint32_t box_inner, uint32_t box_rows) {
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  // libcuda is resolved at run time through the runtime API so that the library still loads (and its
  // symbols can be checked) on a host without a driver.
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled is not available from the installed driver");
      return B2S_ECUDA;
    }
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr), dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (inner=%llu rows=%llu pitch=%llu box=%ux%u)", (int)r,
              (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)row_pitch_bytes, box_inner, box_rows);
    return B2S_ECUDA;
  }
  return 0;
}

// ---- glue kernels -------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_f32_to_bf16_rows(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int ld,
                                                          int base0, int rows0, int base1, int rows1, const int* __restrict__ ctrl) {
  pdl_wait();
  if (ctrl) { rows0 = ctrl[2]; rows1 = ctrl[3]; }
  // one warp per row of 256
  const int s = blockIdx.y;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= (s ? rows1 : rows0)) return;
  const int lane = threadIdx.x & 31;
  const size_t off = (size_t)((s ? base1 : base0) + row) * ld + lane * 8;
  const float4 a = *reinterpret_cast<const float4*>(x + off), b = *reinterpret_cast<const float4*>(x + off + 4);
  __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
  __nv_bfloat162 h2 = __floats2bfloat162_rn(b.x, b.y), h3 = __floats2bfloat162_rn(b.z, b.w);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
  u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(y + off) = u;
}

// LayerNorm(512) + GELU(erf): fp32 in -> bf16 out.  One warp per row.
__global__ void __launch_bounds__(256) k_ln_gelu_512_bf16(const float* __restrict__ h, __nv_bfloat16* __restrict__ y,
                                                          int base0, int rows0, int base1, int rows1,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          const int* __restrict__ ctrl) {
  pdl_wait();
  if (ctrl) {
    if (ctrl[1] || ctrl[2] <= 0 || ctrl[3] <= 0) return;
    rows0 = ctrl[2]; rows1 = ctrl[3];
  }
  const int s = blockIdx.y;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= (s ? rows1 : rows0)) return;
  const int lane = threadIdx.x & 31;
  const size_t roff = (size_t)((s ? base1 : base0) + row) * 512;
  float4 v[4];
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    v[t] = *reinterpret_cast<const float4*>(h + roff + t * 128 + lane * 4);
    sum += (v[t].x + v[t].y) + (v[t].z + v[t].w);
  }
  const float mean = warp_sum(sum) * (1.f / 512.f);
  float sq = 0.f;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float a = v[t].x - mean, b = v[t].y - mean, c = v[t].z - mean, d = v[t].w - mean;
    sq += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(sq) * (1.f / 512.f) + 1e-5f);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int c = t * 128 + lane * 4;
    const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
    __nv_bfloat162 h0 = __floats2bfloat162_rn(gelu_erf_f((v[t].x - mean) * rstd * g.x + b.x), gelu_erf_f((v[t].y - mean) * rstd * g.y + b.y));
    __nv_bfloat162 h1 = __floats2bfloat162_rn(gelu_erf_f((v[t].z - mean) * rstd * g.z + b.z), gelu_erf_f((v[t].w - mean) * rstd * g.w + b.w));
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(y + roff + c) = u;
  }
}

__global__ void k_f32_to_bf16_flat(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __float2bfloat16_rn(x[i]);
}

// ---- state --------------------------------------------------------------------------------
struct TcLinear {            // bf16 weight [N,K] + its tensor map (box {64, BN})
  __nv_bfloat16* w = nullptr; const float* bias = nullptr; int N = 0, K = 0, BN = 0;
  CUtensorMap map;
};
struct TcLayer {
  TcLinear qkv, wo, w1, w2, cqkv, cwo, cw1, cw2;
  const float *lng, *lnb, *clng, *clnb;
};

struct LgTensorCore {
  DeviceArena warena, wsarena;
  std::vector<TcLayer> L;
  int cap = 0;
  __nv_bfloat16 *xb = nullptr, *qkvb = nullptr, *ctxb = nullptr, *msgb = nullptr, *h1b = nullptr;
  float* h1f = nullptr;
  CUtensorMap m_xb, m_ctxb, m_msgb, m_h1b, m_qkv768, m_qkv512;
  const int* ctrl = nullptr;       // LightGlue device state (sizes / early exit), set per match
  KernelProf* prof = nullptr;
};

static int make_linear(LgTensorCore* tc, const float* w_dev, const float* bias, int N, int K, TcLinear* out) {
  out->N = N; out->K = K; out->bias = bias; out->BN = N <= 256 ? 64 : 128;
  B2S_TRY(tc->warena.alloc(&out->w, (size_t)N * K));
  const size_t n = (size_t)N * K;
  k_f32_to_bf16_flat<<<(unsigned)((n + 255) / 256), 256>>>(w_dev, out->w, n);
  B2S_LAUNCH_CHECK();
  return make_tmap_bf16_2d(&out->map, out->w, K, N, (uint64_t)K * 2, 64, out->BN);
}

// FFN first layer with the attention output projection folded in (exact in real arithmetic):
//   ffn.0([x | out_proj(ctx)]) = x W1x^T + ctx (W1m Wo)^T + (b1 + W1m bo)
// so the block needs no separate out_proj GEMM and `msg` is never rounded to bf16.
static int make_folded_ffn1(LgTensorCore* tc, const float* w1_dev, const float* b1_dev, const float* wo_dev, const float* bo_dev,
                            TcLinear* out) {
  std::vector<float> w1(512 * 512), b1(512), wo(256 * 256), bo(256);
  B2S_CUDA(cudaMemcpy(w1.data(), w1_dev, w1.size() * 4, cudaMemcpyDeviceToHost));
  B2S_CUDA(cudaMemcpy(b1.data(), b1_dev, b1.size() * 4, cudaMemcpyDeviceToHost));
  B2S_CUDA(cudaMemcpy(wo.data(), wo_dev, wo.size() * 4, cudaMemcpyDeviceToHost));
  B2S_CUDA(cudaMemcpy(bo.data(), bo_dev, bo.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<float> wf(512 * 512), bf(512);
  for (int o = 0; o < 512; ++o) {
    const float* w1m = &w1[(size_t)o * 512 + 256];
    double bacc = b1[o];
    for (int j = 0; j < 256; ++j) bacc += (double)w1m[j] * bo[j];
    bf[o] = (float)bacc;
    for (int k = 0; k < 256; ++k) wf[(size_t)o * 512 + k] = w1[(size_t)o * 512 + k];
    double acc[256];
    for (int k = 0; k < 256; ++k) acc[k] = 0.0;
    for (int j = 0; j < 256; ++j) {
      const double a = w1m[j];
      const float* wr = &wo[(size_t)j * 256];
      for (int k = 0; k < 256; ++k) acc[k] += a * wr[k];
    }
    for (int k = 0; k < 256; ++k) wf[(size_t)o * 512 + 256 + k] = (float)acc[k];
  }
  float *wf_dev, *bf_dev;
  B2S_TRY(tc->warena.upload(&wf_dev, wf));
  B2S_TRY(tc->warena.upload(&bf_dev, bf));
  return make_linear(tc, wf_dev, bf_dev, 512, 512, out);
}

int lgtc_create(LgTensorCore** out, size_t n_layers) {
  LgTensorCore* tc = new LgTensorCore();
  tc->L.resize(n_layers);
  cudaFuncSetAttribute(k_gemm_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcGemmCfg<64>::SMEM);
  cudaFuncSetAttribute(k_gemm_tc<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcGemmCfg<128>::SMEM);
  cudaFuncSetAttribute(k_attn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM);
  *out = tc;
  return 0;
}

int lgtc_set_layer(LgTensorCore* tc, int li, const LgTcLayerSrc& s) {
  TcLayer& l = tc->L[li];
  B2S_TRY(make_linear(tc, s.wqkv, s.bqkv, 768, 256, &l.qkv));
  B2S_TRY(make_folded_ffn1(tc, s.w1, s.b1, s.wo, s.bo, &l.w1));
  B2S_TRY(make_linear(tc, s.w2, s.b2, 256, 512, &l.w2));
  B2S_TRY(make_linear(tc, s.cwqkv, s.cbqkv, 512, 256, &l.cqkv));
  B2S_TRY(make_folded_ffn1(tc, s.cw1, s.cb1, s.cwo, s.cbo, &l.cw1));
  B2S_TRY(make_linear(tc, s.cw2, s.cb2, 256, 512, &l.cw2));
  l.lng = s.lng; l.lnb = s.lnb; l.clng = s.clng; l.clnb = s.clnb;
  B2S_CUDA(cudaDeviceSynchronize());
  return 0;
}

int lgtc_alloc_ws(LgTensorCore* tc, int cap) {
  tc->wsarena.release();
  tc->cap = 0;
  if (cap % 128) { set_error("lgtc_alloc_ws: cap %d must be a multiple of 128", cap); return B2S_EINVAL; }
  const size_t R = (size_t)2 * cap;
  B2S_TRY(tc->wsarena.alloc(&tc->xb, R * 256)); B2S_TRY(tc->wsarena.alloc(&tc->qkvb, R * 768));
  B2S_TRY(tc->wsarena.alloc(&tc->ctxb, R * 256)); B2S_TRY(tc->wsarena.alloc(&tc->msgb, R * 256));
  B2S_TRY(tc->wsarena.alloc(&tc->h1b, R * 512)); B2S_TRY(tc->wsarena.alloc(&tc->h1f, R * 512));
  // dead rows between the live counts and the tile boundary are read by TMA: keep them finite
  B2S_CUDA(cudaMemset(tc->xb, 0, R * 256 * 2)); B2S_CUDA(cudaMemset(tc->qkvb, 0, R * 768 * 2));
  B2S_CUDA(cudaMemset(tc->ctxb, 0, R * 256 * 2)); B2S_CUDA(cudaMemset(tc->msgb, 0, R * 256 * 2));
  B2S_CUDA(cudaMemset(tc->h1b, 0, R * 512 * 2)); B2S_CUDA(cudaMemset(tc->h1f, 0, R * 512 * 4));
  B2S_TRY(make_tmap_bf16_2d(&tc->m_xb, tc->xb, 256, R, 512, 64, 128));
  B2S_TRY(make_tmap_bf16_2d(&tc->m_ctxb, tc->ctxb, 256, R, 512, 64, 128));
  B2S_TRY(make_tmap_bf16_2d(&tc->m_msgb, tc->msgb, 256, R, 512, 64, 128));
  B2S_TRY(make_tmap_bf16_2d(&tc->m_h1b, tc->h1b, 512, R, 1024, 64, 128));
  B2S_TRY(make_tmap_bf16_2d(&tc->m_qkv768, tc->qkvb, 768, R, 1536, 64, 128));
  B2S_TRY(make_tmap_bf16_2d(&tc->m_qkv512, tc->qkvb, 512, R, 1024, 64, 128));
  tc->cap = cap;
  return 0;
}

void lgtc_destroy(LgTensorCore* tc) { delete tc; }
void lgtc_set_prof(LgTensorCore* tc, KernelProf* prof) { tc->prof = prof; }

static int tc_gemm(LgTensorCore* tc, cudaStream_t st, const CUtensorMap& a1, const CUtensorMap& a2, int K1, const TcLinear& w,
                   TcGemmParams p, int m, int n, long long* launches) {
  p.K = w.K; p.K1 = K1; p.N = w.N; p.bias = w.bias; p.ctrl = tc->ctrl;
  p.seg_base[0] = 0; p.seg_base[1] = tc->cap; p.seg_rows[0] = m; p.seg_rows[1] = n;
  p.tiles0 = cdiv(m, 128);
  const int tiles = p.tiles0 + cdiv(n, 128);
  if (tiles <= 0) return 0;
  dim3 grid(w.N / w.BN, tiles);
  if (tc->prof) tc->prof->mark(PROF_GEMM, st);
  if (w.BN == 64) launch_k(k_gemm_tc<64>, grid, 192, TcGemmCfg<64>::SMEM, st, a1, a2, w.map, p);
  else launch_k(k_gemm_tc<128>, grid, 192, TcGemmCfg<128>::SMEM, st, a1, a2, w.map, p);
  if (tc->prof) tc->prof->mark(PROF_GEMM, st);
  if (launches) ++*launches;
  B2S_LAUNCH_CHECK();
  return 0;
}

static int tc_attention(LgTensorCore* tc, cudaStream_t st, const CUtensorMap& map, AttnTcParams ap, long long* launches) {
  const int maxq = std::max(ap.prob[0].nq, ap.prob[1].nq);
  if (maxq <= 0) return 0;
  ap.scale_log2e = 0.125f * 1.4426950408889634f;
  ap.out = tc->ctxb; ap.ldo = 256; ap.ctrl = tc->ctrl;
  dim3 grid(cdiv(maxq, ATC_BQ), 4, 2);
  if (tc->prof) tc->prof->mark(PROF_ATTN, st);
  launch_k(k_attn_tc, grid, ATC_THREADS, ATC_SMEM, st, map, ap);
  if (tc->prof) tc->prof->mark(PROF_ATTN, st);
  if (launches) ++*launches;
  B2S_LAUNCH_CHECK();
  return 0;
}

static int tc_ffn(LgTensorCore* tc, cudaStream_t st, float* x, const TcLinear& w1, const float* lng, const float* lnb, const TcLinear& w2,
                  int m, int n, long long* launches) {
  TcGemmParams p = {};
  p.epi = TC_EPI_F32; p.out_f32 = tc->h1f; p.ld_f32 = 512;
  B2S_TRY(tc_gemm(tc, st, tc->m_xb, tc->m_ctxb, 256, w1, p, m, n, launches));          // [x | ctx] W1'^T + b1' (out_proj folded)
  dim3 g(cdiv(std::max(m, n), 8), 2);
  launch_k(k_ln_gelu_512_bf16, g, 256, 0, st, tc->h1f, tc->h1b, 0, m, tc->cap, n, lng, lnb, tc->ctrl);
  if (launches) ++*launches;
  B2S_LAUNCH_CHECK();
  p = TcGemmParams();
  p.epi = TC_EPI_RESID_F32_BF16; p.out_f32 = x; p.ld_f32 = 256; p.out_bf16 = tc->xb; p.ld_bf16 = 256;
  return tc_gemm(tc, st, tc->m_h1b, tc->m_h1b, 512, w2, p, m, n, launches);              // x += h W2^T + b2 ; xb = bf16(x)
}

__nv_bfloat16* lgtc_xb(LgTensorCore* tc) { return tc->xb; }

// One transformer layer.  m, n are upper bounds of the live point counts (they size the grids); the
// live counts and the early-exit flag are read from `ctrl` on the device.  `x` is the fp32 residual
// stream of this layer; its bf16 copy xb is maintained by the FFN epilogues / the pruning gather
// (derive_xb: re-derive it from x first - layer 0, right after the input projection).
int lgtc_layer(LgTensorCore* tc, cudaStream_t st, int li, float* x, const float* cosb, const float* sinb, int cap, int m, int n,
               const int* ctrl, bool derive_xb, long long* launches) {
  if (cap != tc->cap) { set_error("lgtc_layer: workspace capacity mismatch"); return B2S_EINVAL; }
  const TcLayer& l = tc->L[li];
  tc->ctrl = ctrl;
  if (derive_xb) {
    dim3 g(cdiv(std::max(m, n), 8), 2);
    launch_k(k_f32_to_bf16_rows, g, 256, 0, st, x, tc->xb, 256, 0, m, cap, n, ctrl);
    if (launches) ++*launches;
    B2S_LAUNCH_CHECK();
  }
  TcGemmParams p = {};
  // ---- self block ----
  p.epi = TC_EPI_ROTARY_BF16; p.out_bf16 = tc->qkvb; p.ld_bf16 = 768; p.rot_cos = cosb; p.rot_sin = sinb; p.rot_cols = 512;
  B2S_TRY(tc_gemm(tc, st, tc->m_xb, tc->m_xb, 256, l.qkv, p, m, n, launches));
  AttnTcParams ap = {};
  ap.qcol = 0; ap.kcol = 256; ap.vcol = 512; ap.cross = 0;
  ap.prob[0] = {0, 0, m, m}; ap.prob[1] = {cap, cap, n, n};
  B2S_TRY(tc_attention(tc, st, tc->m_qkv768, ap, launches));
  B2S_TRY(tc_ffn(tc, st, x, l.w1, l.lng, l.lnb, l.w2, m, n, launches));
  // ---- cross block ----
  p = TcGemmParams(); p.epi = TC_EPI_BF16; p.out_bf16 = tc->qkvb; p.ld_bf16 = 512;
  B2S_TRY(tc_gemm(tc, st, tc->m_xb, tc->m_xb, 256, l.cqkv, p, m, n, launches));
  ap = AttnTcParams();
  ap.qcol = 0; ap.kcol = 0; ap.vcol = 256; ap.cross = 1;
  ap.prob[0] = {0, cap, m, n}; ap.prob[1] = {cap, 0, n, m};
  B2S_TRY(tc_attention(tc, st, tc->m_qkv512, ap, launches));
  return tc_ffn(tc, st, x, l.cw1, l.clng, l.clnb, l.cw2, m, n, launches);
}

}  // namespace b2s

// ---- unit-test entry points (host buffers) --------------------------------------------------
using namespace b2s;

static std::vector<__nv_bfloat16> to_bf16(const float* x, size_t n) {
  std::vector<__nv_bfloat16> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = __float2bfloat16_rn(x[i]);
  return v;
}

// C[M,N] (fp32) = A[M,K] W[N,K]^T + bias, operands rounded to bf16.  M, N multiples of 64; K of 64.
extern "C" int b2s_test_gemm_tc(const float* A, const float* W, const float* bias, int M, int N, int K, float* C) {
  if (!A || !W || !C || M <= 0 || N % 64 || K % 64) { set_error("b2s_test_gemm_tc: bad shape"); return B2S_EINVAL; }
  DeviceArena ar;
  const int Mp = cdiv(M, 128) * 128;
  __nv_bfloat16 *dA, *dW; float *dB, *dC;
  std::vector<__nv_bfloat16> hA = to_bf16(A, (size_t)M * K), hW = to_bf16(W, (size_t)N * K);
  B2S_TRY(ar.alloc(&dA, (size_t)Mp * K)); B2S_TRY(ar.alloc(&dW, (size_t)N * K)); B2S_TRY(ar.alloc(&dB, (size_t)N)); B2S_TRY(ar.alloc(&dC, (size_t)Mp * N));
  B2S_CUDA(cudaMemset(dA, 0, (size_t)Mp * K * 2));
  B2S_CUDA(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  B2S_CUDA(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
  std::vector<float> zb(N, 0.f);
  B2S_CUDA(cudaMemcpy(dB, bias ? bias : zb.data(), (size_t)N * 4, cudaMemcpyHostToDevice));
  const int BN = N % 128 == 0 && N > 256 ? 128 : 64;
  CUtensorMap ma, mw;
  B2S_TRY(make_tmap_bf16_2d(&ma, dA, K, Mp, (uint64_t)K * 2, 64, 128));
  B2S_TRY(make_tmap_bf16_2d(&mw, dW, K, N, (uint64_t)K * 2, 64, BN));
  cudaFuncSetAttribute(k_gemm_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcGemmCfg<64>::SMEM);
  cudaFuncSetAttribute(k_gemm_tc<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcGemmCfg<128>::SMEM);
  TcGemmParams p = {};
  p.K = K; p.K1 = K; p.N = N; p.bias = dB; p.epi = TC_EPI_F32; p.out_f32 = dC; p.ld_f32 = N;
  p.seg_base[0] = 0; p.seg_rows[0] = M; p.seg_base[1] = 0; p.seg_rows[1] = 0; p.tiles0 = cdiv(M, 128);
  dim3 grid(N / BN, p.tiles0);
  if (BN == 64) k_gemm_tc<64><<<grid, 192, TcGemmCfg<64>::SMEM>>>(ma, ma, mw, p);
  else k_gemm_tc<128><<<grid, 192, TcGemmCfg<128>::SMEM>>>(ma, ma, mw, p);
  B2S_LAUNCH_CHECK();
  B2S_CUDA(cudaDeviceSynchronize());
  B2S_CUDA(cudaMemcpy(C, dC, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
  return 0;
}

// ctx[nq,256] = per-head softmax(q k^T / 8) v, 4 heads of 64; operands rounded to bf16.
extern "C" int b2s_test_attn_tc(const float* q, const float* k, const float* v, int nq, int nk, float* ctx) {
  if (!q || !k || !v || !ctx || nq <= 0 || nk <= 0) { set_error("b2s_test_attn_tc: bad shape"); return B2S_EINVAL; }
  DeviceArena ar;
  const int R = cdiv(std::max(nq, nk), 128) * 128;
  std::vector<__nv_bfloat16> buf((size_t)R * 768, __float2bfloat16_rn(0.f));
  for (int i = 0; i < nq; ++i) for (int c = 0; c < 256; ++c) buf[(size_t)i * 768 + c] = __float2bfloat16_rn(q[(size_t)i * 256 + c]);
  for (int i = 0; i < nk; ++i) for (int c = 0; c < 256; ++c) {
    buf[(size_t)i * 768 + 256 + c] = __float2bfloat16_rn(k[(size_t)i * 256 + c]);
    buf[(size_t)i * 768 + 512 + c] = __float2bfloat16_rn(v[(size_t)i * 256 + c]);
  }
  __nv_bfloat16 *dq, *dctx;
  B2S_TRY(ar.alloc(&dq, buf.size())); B2S_TRY(ar.alloc(&dctx, (size_t)R * 256));
  B2S_CUDA(cudaMemcpy(dq, buf.data(), buf.size() * 2, cudaMemcpyHostToDevice));
  B2S_CUDA(cudaMemset(dctx, 0, (size_t)R * 256 * 2));
  CUtensorMap map;
  B2S_TRY(make_tmap_bf16_2d(&map, dq, 768, R, 1536, 64, 128));
  cudaFuncSetAttribute(k_attn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM);
  AttnTcParams ap = {};
  ap.qcol = 0; ap.kcol = 256; ap.vcol = 512; ap.prob[0] = {0, 0, nq, nk}; ap.prob[1] = {0, 0, 0, 0};
  ap.scale_log2e = 0.125f * 1.4426950408889634f; ap.out = dctx; ap.ldo = 256;
  k_attn_tc<<<dim3(cdiv(nq, 128), 4, 1), ATC_THREADS, ATC_SMEM>>>(map, ap);
  B2S_LAUNCH_CHECK();
  B2S_CUDA(cudaDeviceSynchronize());
  std::vector<__nv_bfloat16> out((size_t)nq * 256);
  B2S_CUDA(cudaMemcpy(out.data(), dctx, out.size() * 2, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < out.size(); ++i) ctx[i] = __bfloat162float(out[i]);
  return 0;
}

// Device-only timing of the attention kernel on random bf16 data: one launch = what a LightGlue
// self block issues (2 problems x 4 heads, nq queries x nk keys each).  ms_out = mean per launch.
extern "C" int b2s_bench_attn_tc(int nq, int nk, int iters, float* ms_out) {
  if (nq <= 0 || nk <= 0 || iters <= 0 || !ms_out) { set_error("b2s_bench_attn_tc: bad argument"); return B2S_EINVAL; }
  DeviceArena ar;
  const int cap = cdiv(std::max(nq, nk), 128) * 128;
  const size_t R = (size_t)2 * cap;
  std::vector<__nv_bfloat16> buf(R * 768);
  uint32_t s = 12345u;
  for (auto& b : buf) { s = s * 1664525u + 1013904223u; b = __float2bfloat16_rn(((s >> 8) & 0xFFFF) / 32768.f - 1.f); }
  __nv_bfloat16 *dq, *dctx;
  B2S_TRY(ar.alloc(&dq, buf.size())); B2S_TRY(ar.alloc(&dctx, R * 256));
  B2S_CUDA(cudaMemcpy(dq, buf.data(), buf.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap map;
  B2S_TRY(make_tmap_bf16_2d(&map, dq, 768, R, 1536, 64, 128));
  cudaFuncSetAttribute(k_attn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM);
  AttnTcParams ap = {};
  ap.qcol = 0; ap.kcol = 256; ap.vcol = 512; ap.prob[0] = {0, 0, nq, nk}; ap.prob[1] = {cap, cap, nq, nk};
  ap.scale_log2e = 0.125f * 1.4426950408889634f; ap.out = dctx; ap.ldo = 256;
  cudaEvent_t e0, e1;
  B2S_CUDA(cudaEventCreate(&e0)); B2S_CUDA(cudaEventCreate(&e1));
  const dim3 grid(cdiv(nq, 128), 4, 2);
  for (int i = 0; i < 3; ++i) k_attn_tc<<<grid, ATC_THREADS, ATC_SMEM>>>(map, ap);
  B2S_CUDA(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) k_attn_tc<<<grid, ATC_THREADS, ATC_SMEM>>>(map, ap);
  B2S_CUDA(cudaEventRecord(e1));
  B2S_LAUNCH_CHECK();
  B2S_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  B2S_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  *ms_out = ms / iters;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return 0;
}
