// aliked.cu - host orchestration of the ALIKED extractor behind b2s_aliked_*.
// Replaces `detector.extract(t0)` (+ `_bgr_to_tensor`) at
// /root/reference/slam/core/features_utils.py:94,219-222; arithmetic spec: SURVEY.md A.1/A.2.
#include "aliked_kernels.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "gemm_tcp.cuh"
#include "conv_tc.cuh"

#include <algorithm>
#include <cstdlib>
#include <cmath>

using namespace b2s;

// output rows per CTA of block2.conv2 on the fp16x2 path (B2S_CONV22_R at build time for A/B measurements)
#ifndef B2S_CONV22_R
#define B2S_CONV22_R 5
#endif
static constexpr int CONV22_R = B2S_CONV22_R;

struct TcWeight {     // fp32 weight [N][K] as three bf16 planes [N][3K] + tensor map (box {64, 64})
  __nv_bfloat16* w = nullptr; CUtensorMap map; int N = 0, K = 0;
};

struct DcnBlockW {   // ResBlock with deformable convs (block3 / block4)
  int cin, cout;
  float *off1_w, *off1_b, *reg1_w, *reg1_b;   // conv1: offset conv [18][9*cin], regular [cout][9*cin] (BN1 folded)
  float *off2_w, *off2_b, *reg2_w, *reg2_b;   // conv2: [18][9*cout], [cout][9*cout] (BN2 folded)
  float *ds_w, *ds_b;                         // downsample [cout][cin] + bias
  TcWeight reg1_tc, reg2_tc;                  // regular convs as tensor-core weights (block3)
};

struct b2s_aliked {
  b2s_aliked_cfg cfg;
  int device = 0, M = 16, n_limit = 0;
  DeviceArena warena, wsarena, hostarena;
  // weights
  float *b1c1_w, *b1c1_b, *b1c2_w, *b1c2_b;            // block1 [cin][9][16]
  float *b2c1_w, *b2c1_b, *b2c2_w, *b2c2_b, *b2ds_w, *b2ds_b;
  DcnBlockW b3, b4;
  float *agg_w[4];                                      // conv1..4 [32][ci]
  float *sh0, *sh2, *sh4, *sh6;                         // score head
  float *so0_w, *so0_b, *so2_w, *so2_b, *sf_w, *aggT;   // desc head
  TcWeight tc_so0, tc_sf, tc_agg;                       // desc-head contractions on the tensor cores (fp32 as bf16x3)
  __nv_bfloat16 *b1c2_pl, *b2c1_pl, *b2c2_pl;           // conv weights as chunk planes (conv_tc.cuh): conv_np planes
  int conv_np = 2;                                       // operand planes of the implicit-GEMM convolutions: 2 (fp16 pair, range-checked) or 3 (bf16)
  int* range_flag = nullptr;                             // device flag: an activation left the fp16 range (conv_np == 2)
  __nv_bfloat16 *t1a_pl, *x1p_pl, *t2a_pl;              // activations as chunk planes: block1.conv1 out, pool2(x1), block2.conv1 out
  CUtensorMap m_t1a, m_x1p, m_t2a; int mapHp = 0, mapWp = 0;
  __nv_bfloat16* col3_pl;                               // block3 deformable im2col as bf16x3 planes [3][P3][<= 576]
  CUtensorMap m_col3a, m_col3b;                         // views with row length 9*32 / 9*64
  // workspace
  int wsHp = 0, wsWp = 0;
  float *img_pad, *resized, *t1a, *x1, *r2, *t2a, *x2, *x3in, *col3, *off3, *t3a, *r3, *x3, *x4in, *col4, *off4, *t4a, *r4, *x4;
  float *x2a, *x3a, *x4a, *p8[3], *s8, *score, *nms, *cand_sc, *sel_sc, *thr, *kp_norm, *disp, *sampled;
  float *toff1, *offs, *descraw;
  __nv_bfloat16 *Apatch, *S, *F;                        // bf16x3 planes: [3][K][1152], [3][K*M][128], [3][K*M][128]
  CUtensorMap m_Apatch, m_S, m_F;
  float* skws = nullptr; size_t skws_floats = 0;        // split-K partial sums
  int *cand_idx, *sel_idx, *dk;
  FeatSrc fsrc = {};                                    // last call's on-demand feature source (SDDH, debug tap)
  int cand_cap = 0;
  // last-call geometry (for debug taps)
  int Hr = 0, Wr = 0, Hp = 0, Wp = 0;
  // host API staging
  size_t himg_bytes = 0; void* himg = nullptr;
  float *hkp = nullptr, *hdesc = nullptr, *hscores = nullptr; int32_t* hn = nullptr;
  long long launches = 0;
  float desc_renorm_eps = 0.f;                           // set per call by b2s_aliked_extract_host_ex
  float batch_renorm_eps = 0.f;                          // b2s_aliked_set_batch_renorm: what b2s_aliked_extract_batch applies
  // split host API (begin / keypoints / finish): pinned staging for the early keypoint copy
  float* pin_kp = nullptr; int32_t* pin_n = nullptr; cudaEvent_t ev_kp = nullptr; bool early_kp = false; bool pending = false;  b2s_remap* undist = nullptr;                           // optional ingest stage (b2s_aliked_set_undistort); borrowed
  // batched extraction (b2s_aliked_extract_batch): extra LANES = complete extractors (own workspace, own stream) created
  // lazily from a copy of the weight blob; frame i of a batch runs on lane i mod n_lanes, the lanes run concurrently
  // CUDA-graph replay of one extraction (batched path): the ~60 launches of a frame are captured once per frame shape
  // over fixed staging buffers (image in, keypoints / descriptors / scores / count out) and replayed with one launch
  DeviceArena grarena;
  cudaGraphExec_t gexec = nullptr; int g_fmt = -1, g_H = 0, g_W = 0, g_stride = 0; float g_eps = 0.f; bool g_failed = false;
  uint8_t* g_img = nullptr; size_t g_img_bytes = 0; float *g_kp = nullptr, *g_desc = nullptr, *g_sc = nullptr; int32_t* g_n = nullptr;
  long long g_launches = 0;
  float* thr0 = nullptr;                                 // device copy of cfg.det_thresh (source of the per-call reset of thr)
  std::vector<uint8_t> blob;
  std::vector<b2s_aliked*> lanes; std::vector<cudaStream_t> lane_st; std::vector<cudaEvent_t> lane_ev; cudaEvent_t ev_fork = nullptr;
};

extern "C" void b2s_aliked_default_cfg(b2s_aliked_cfg* c) {
  c->model = 0; c->max_kp = 2048; c->det_thresh = 0.2f; c->nms_radius = 2; c->resize_long = 1024; c->precision = B2S_FP32;
}

namespace {

// GEMM with the handle's split-K workspace attached (small-M/N problems of the DCN / SDDH stages)
int agemm(b2s_aliked* h, GemmParams& g, cudaStream_t st) {
  g.splitk_ws = h->skws; g.splitk_ws_floats = h->skws_floats;
  return gemm_simt(g, st, &h->launches);
}

struct BnFold { std::vector<float> scale, shift; };

int bn_fold(const WeightBlob& wb, const std::string& name, int c, BnFold* out) {
  const TensorView *g = wb.get(name + ".weight", c), *b = wb.get(name + ".bias", c);
  const TensorView *m = wb.get(name + ".running_mean", c), *v = wb.get(name + ".running_var", c);
  if (!g || !b || !m || !v) return B2S_EINVAL;
  out->scale.resize(c); out->shift.resize(c);
  for (int i = 0; i < c; ++i) {
    const float s = g->data[i] / std::sqrt(v->data[i] + 1e-5f);
    out->scale[i] = s; out->shift[i] = b->data[i] - m->data[i] * s;
  }
  return 0;
}

// [cout][cin][3][3] (*scale[cout]) -> direct-conv layout [cin][9][cout]
int upload_conv_direct(b2s_aliked* h, const WeightBlob& wb, const std::string& name, int cout, int cin, const BnFold& bn, float** w, float** b) {
  const TensorView* t = wb.get(name + ".weight", (size_t)cout * cin * 9);
  if (!t) return B2S_EINVAL;
  std::vector<float> o((size_t)cin * 9 * cout);
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int k = 0; k < 9; ++k) o[((size_t)ci * 9 + k) * cout + co] = t->data[((size_t)co * cin + ci) * 9 + k] * bn.scale[co];
  B2S_TRY(h->warena.upload(w, o));
  return h->warena.upload(b, bn.shift);
}

// [cout][cin][3][3] (*scale) -> GEMM layout [cout][tap*cin + ci]
int upload_conv_gemm(b2s_aliked* h, const WeightBlob& wb, const std::string& name, int cout, int cin, const float* scale, float** w) {
  const TensorView* t = wb.get(name + ".weight", (size_t)cout * cin * 9);
  if (!t) return B2S_EINVAL;
  std::vector<float> o((size_t)cout * 9 * cin);
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int k = 0; k < 9; ++k)
        o[(size_t)co * 9 * cin + (size_t)k * cin + ci] = t->data[((size_t)co * cin + ci) * 9 + k] * (scale ? scale[co] : 1.f);
  return h->warena.upload(w, o);
}

int upload_plain(b2s_aliked* h, const WeightBlob& wb, const std::string& name, size_t numel, float** w) {
  const TensorView* t = wb.get(name, numel);
  if (!t) return B2S_EINVAL;
  std::vector<float> o(t->data, t->data + numel);
  return h->warena.upload(w, o);
}

// fp32 device weight [N][K] -> bf16x3 planes [N][3*Kpad] (K zero-padded to a multiple of 64) + tensor map
int make_tc_weight(b2s_aliked* h, const float* w_dev, int N, int K, TcWeight* out) {
  const int Kpad = cdiv(K, 64) * 64;
  out->N = N; out->K = Kpad;
  const float* src = w_dev;
  if (Kpad != K) {
    float* padded;
    B2S_TRY(h->warena.alloc(&padded, (size_t)N * Kpad));
    B2S_CUDA(cudaMemset(padded, 0, (size_t)N * Kpad * sizeof(float)));
    B2S_CUDA(cudaMemcpy2D(padded, (size_t)Kpad * sizeof(float), w_dev, (size_t)K * sizeof(float), (size_t)K * sizeof(float), N,
                          cudaMemcpyDeviceToDevice));
    src = padded;
  }
  const size_t n = (size_t)N * Kpad;
  B2S_TRY(h->warena.alloc(&out->w, 3 * n));
  k_weight_planes<<<(unsigned)((n + 255) / 256), 256>>>(src, out->w, N, Kpad, 3);
  B2S_LAUNCH_CHECK();
  return make_tmap_bf16_2d(&out->map, out->w, (uint64_t)3 * Kpad, N, (uint64_t)3 * Kpad * 2, 64, 64);
}

int upload_dcn_block(b2s_aliked* h, const WeightBlob& wb, const std::string& blk, int cin, int cout, DcnBlockW* d) {
  d->cin = cin; d->cout = cout;
  BnFold bn1, bn2;
  B2S_TRY(bn_fold(wb, blk + ".bn1", cout, &bn1));
  B2S_TRY(bn_fold(wb, blk + ".bn2", cout, &bn2));
  B2S_TRY(upload_conv_gemm(h, wb, blk + ".conv1.offset_conv", 18, cin, nullptr, &d->off1_w));
  B2S_TRY(upload_plain(h, wb, blk + ".conv1.offset_conv.bias", 18, &d->off1_b));
  B2S_TRY(upload_conv_gemm(h, wb, blk + ".conv1.regular_conv", cout, cin, bn1.scale.data(), &d->reg1_w));
  B2S_TRY(h->warena.upload(&d->reg1_b, bn1.shift));
  B2S_TRY(upload_conv_gemm(h, wb, blk + ".conv2.offset_conv", 18, cout, nullptr, &d->off2_w));
  B2S_TRY(upload_plain(h, wb, blk + ".conv2.offset_conv.bias", 18, &d->off2_b));
  B2S_TRY(upload_conv_gemm(h, wb, blk + ".conv2.regular_conv", cout, cout, bn2.scale.data(), &d->reg2_w));
  B2S_TRY(h->warena.upload(&d->reg2_b, bn2.shift));
  B2S_TRY(upload_plain(h, wb, blk + ".downsample.weight", (size_t)cout * cin, &d->ds_w));
  B2S_TRY(upload_plain(h, wb, blk + ".downsample.bias", cout, &d->ds_b));
  B2S_TRY(make_tc_weight(h, d->reg1_w, cout, 9 * cin, &d->reg1_tc));
  B2S_TRY(make_tc_weight(h, d->reg2_w, cout, 9 * cout, &d->reg2_tc));
  return 0;
}

int load_weights(b2s_aliked* h, const WeightBlob& wb) {
  const int M = h->M;
  BnFold bn;
  B2S_TRY(bn_fold(wb, "block1.bn1", 16, &bn));
  B2S_TRY(upload_conv_direct(h, wb, "block1.conv1", 16, 3, bn, &h->b1c1_w, &h->b1c1_b));
  B2S_TRY(bn_fold(wb, "block1.bn2", 16, &bn));
  B2S_TRY(upload_conv_direct(h, wb, "block1.conv2", 16, 16, bn, &h->b1c2_w, &h->b1c2_b));
  B2S_TRY(bn_fold(wb, "block2.bn1", 32, &bn));
  B2S_TRY(upload_conv_direct(h, wb, "block2.conv1", 32, 16, bn, &h->b2c1_w, &h->b2c1_b));
  B2S_TRY(bn_fold(wb, "block2.bn2", 32, &bn));
  B2S_TRY(upload_conv_direct(h, wb, "block2.conv2", 32, 32, bn, &h->b2c2_w, &h->b2c2_b));
  {
    auto mkconv = [&](const std::string& name, const std::string& bnname, int cout, int cin, __nv_bfloat16** out) -> int {
      BnFold b;
      B2S_TRY(bn_fold(wb, bnname, cout, &b));
      const TensorView* t = wb.get(name + ".weight", (size_t)cout * cin * 9);
      if (!t) return B2S_EINVAL;
      const int nch = cin / 8;
      const size_t plane = (size_t)9 * cin * cout;
      std::vector<__nv_bfloat16> o(3 * plane);
      uint16_t* o16 = reinterpret_cast<uint16_t*>(o.data());
      for (int tap = 0; tap < 9; ++tap)
        for (int cc = 0; cc < nch; ++cc)
          for (int co = 0; co < cout; ++co)
            for (int j = 0; j < 8; ++j) {
              float r = t->data[((size_t)co * cin + cc * 8 + j) * 9 + tap] * b.scale[co];
              const size_t idx = (((size_t)tap * nch + cc) * cout + co) * 8 + j;
              if (h->conv_np == 2) {           // fp16 pair: h0 = fp16(w), h1 = fp16((w - h0) * 2^11)  (tc::pack_h2)
                const __half h0 = __float2half_rn(r);
                const __half h1 = __float2half_rn((r - __half2float(h0)) * 2048.f);
                o16[idx] = *reinterpret_cast<const uint16_t*>(&h0);
                o16[plane + idx] = *reinterpret_cast<const uint16_t*>(&h1);
                continue;
              }
              for (int pl = 0; pl < 3; ++pl) {
                const __nv_bfloat16 q = __float2bfloat16_rn(r);
                o[pl * plane + idx] = q;
                r -= __bfloat162float(q);
              }
            }
      return h->warena.upload(out, o);
    };
    B2S_TRY(mkconv("block1.conv2", "block1.bn2", 16, 16, &h->b1c2_pl));
    B2S_TRY(mkconv("block2.conv1", "block2.bn1", 32, 16, &h->b2c1_pl));
    B2S_TRY(mkconv("block2.conv2", "block2.bn2", 32, 32, &h->b2c2_pl));
  }
  B2S_TRY(upload_plain(h, wb, "block2.downsample.weight", 32 * 16, &h->b2ds_w));
  B2S_TRY(upload_plain(h, wb, "block2.downsample.bias", 32, &h->b2ds_b));
  B2S_TRY(upload_dcn_block(h, wb, "block3", 32, 64, &h->b3));
  B2S_TRY(upload_dcn_block(h, wb, "block4", 64, 128, &h->b4));
  const int ci[4] = {16, 32, 64, 128};
  for (int i = 0; i < 4; ++i) B2S_TRY(upload_plain(h, wb, "conv" + std::to_string(i + 1) + ".weight", (size_t)32 * ci[i], &h->agg_w[i]));
  B2S_TRY(upload_plain(h, wb, "score_head.0.weight", 8 * 128, &h->sh0));
  B2S_TRY(upload_plain(h, wb, "score_head.2.weight", 4 * 8 * 9, &h->sh2));
  B2S_TRY(upload_plain(h, wb, "score_head.4.weight", 4 * 4 * 9, &h->sh4));
  B2S_TRY(upload_plain(h, wb, "score_head.6.weight", 4 * 9, &h->sh6));
  B2S_TRY(upload_conv_gemm(h, wb, "desc_head.offset_conv.0", 2 * M, 128, nullptr, &h->so0_w));
  B2S_TRY(upload_plain(h, wb, "desc_head.offset_conv.0.bias", 2 * M, &h->so0_b));
  B2S_TRY(upload_plain(h, wb, "desc_head.offset_conv.2.weight", (size_t)4 * M * M, &h->so2_w));
  B2S_TRY(upload_plain(h, wb, "desc_head.offset_conv.2.bias", 2 * M, &h->so2_b));
  B2S_TRY(upload_plain(h, wb, "desc_head.sf_conv.weight", 128 * 128, &h->sf_w));
  // agg_weights [M][c][d] -> [d][p*128 + c]   (einsum 'ncp,pcd->nd' as one GEMM over K = M*128)
  const TensorView* ag = wb.get("desc_head.agg_weights", (size_t)M * 128 * 128);
  if (!ag) return B2S_EINVAL;
  std::vector<float> at((size_t)128 * M * 128);
  for (int p = 0; p < M; ++p)
    for (int c = 0; c < 128; ++c)
      for (int d = 0; d < 128; ++d) at[(size_t)d * M * 128 + (size_t)p * 128 + c] = ag->data[((size_t)p * 128 + c) * 128 + d];
  B2S_TRY(h->warena.upload(&h->aggT, at));
  B2S_TRY(make_tc_weight(h, h->so0_w, 2 * M, 1152, &h->tc_so0));
  B2S_TRY(make_tc_weight(h, h->sf_w, 128, 128, &h->tc_sf));
  B2S_TRY(make_tc_weight(h, h->aggT, 128, M * 128, &h->tc_agg));
  B2S_CUDA(cudaDeviceSynchronize());
  return 0;
}

// C = act(A W^T + bias) on the tensor cores, fp32 operands as bf16x3 planes; rows = *n_dev * mult
int tc_gemm(b2s_aliked* h, cudaStream_t st, const CUtensorMap& a, int plane_rows, const TcWeight& w, const float* bias, int act,
            int rows_max, const int32_t* n_dev, int mult, float* out_f32, int ldc, __nv_bfloat16* out_planes, size_t out_plane,
            const float* residual = nullptr) {
  TcGemmParams p = {};
  p.residual = residual; p.ld_res = ldc;
  p.K = w.K; p.K1 = w.K; p.N = w.N; p.bias = bias; p.act = act;
  p.seg_stride = 0; p.seg_rows = rows_max; p.tiles_per_seg = cdiv(rows_max, 128);
  p.plane_rows = plane_rows; p.m_dev = n_dev; p.m_mult = mult;
  if (out_f32) { p.epi = TC_EPI_F32; p.out_f32 = out_f32; p.ld_f32 = ldc; }
  else { p.epi = TC_EPI_BF16; p.out_bf16 = out_planes; p.ld_bf16 = ldc; p.out_plane = out_plane; }
  const int total = cdiv(w.N, 64) * p.tiles_per_seg;
  if (total >= 2 * 148) {
    // many tiles (SDDH sf_conv: 32768 sample rows): the persistent tile scheduler overlaps every tile's epilogue (this
    // K = 128 GEMM is all epilogue) with the next tile's main loop
    launch_k(k_gemm_tcp<64, 3>, dim3(148), TcpGemmCfg<64, 3>::THREADS, TcpGemmCfg<64, 3>::SMEM, st, a, a, w.map, p, p.tiles_per_seg);
  } else {
    launch_k(k_gemm_tc<64, 3>, dim3(cdiv(w.N, 64), p.tiles_per_seg), TcGemmCfg<64, 3>::THREADS, TcGemmCfg<64, 3>::SMEM, st, a, a, w.map, p);
  }
  ++h->launches;
  B2S_LAUNCH_CHECK();
  return 0;
}

int alloc_ws(b2s_aliked* h, int Hp, int Wp) {
  h->wsarena.release();
  h->wsHp = h->wsWp = 0;
  const size_t P = (size_t)Hp * Wp, P2 = P / 4, P3 = P / 64, P4 = P / 1024;
  const size_t K = (size_t)h->n_limit, M = (size_t)h->M;
  DeviceArena& a = h->wsarena;
  B2S_TRY(a.alloc(&h->img_pad, 3 * P)); B2S_TRY(a.alloc(&h->resized, 3 * P));
  B2S_TRY(a.alloc(&h->t1a, (size_t)1)); B2S_TRY(a.alloc(&h->x1, 16 * P));
  B2S_TRY(a.alloc(&h->t1a_pl, 3 * 16 * P)); B2S_TRY(a.alloc(&h->x1p_pl, 3 * 16 * P2)); B2S_TRY(a.alloc(&h->t2a_pl, 3 * 32 * P2));
  B2S_TRY(a.alloc(&h->col3_pl, 3 * 576 * P3));
  B2S_CUDA(cudaMemset(h->col3_pl, 0, 3 * 576 * P3 * 2));
  h->mapHp = h->mapWp = 0;
  B2S_TRY(a.alloc(&h->r2, 32 * P2)); B2S_TRY(a.alloc(&h->t2a, (size_t)1)); B2S_TRY(a.alloc(&h->x2, 32 * P2));
  B2S_TRY(a.alloc(&h->x3in, 32 * P3)); B2S_TRY(a.alloc(&h->col3, 576 * P3)); B2S_TRY(a.alloc(&h->off3, 18 * P3));
  B2S_TRY(a.alloc(&h->t3a, 64 * P3)); B2S_TRY(a.alloc(&h->r3, 64 * P3)); B2S_TRY(a.alloc(&h->x3, 64 * P3));
  B2S_TRY(a.alloc(&h->x4in, 64 * P4)); B2S_TRY(a.alloc(&h->col4, 1152 * P4)); B2S_TRY(a.alloc(&h->off4, 18 * P4));
  B2S_TRY(a.alloc(&h->t4a, 128 * P4)); B2S_TRY(a.alloc(&h->r4, 128 * P4)); B2S_TRY(a.alloc(&h->x4, 128 * P4));
  B2S_TRY(a.alloc(&h->x2a, 32 * P2)); B2S_TRY(a.alloc(&h->x3a, 32 * P3)); B2S_TRY(a.alloc(&h->x4a, 32 * P4));
  B2S_TRY(a.alloc(&h->p8[0], 8 * P2)); B2S_TRY(a.alloc(&h->p8[1], 8 * P3)); B2S_TRY(a.alloc(&h->p8[2], 8 * P4));
  B2S_TRY(a.alloc(&h->s8, 8 * P));
  B2S_TRY(a.alloc(&h->score, P)); B2S_TRY(a.alloc(&h->nms, P));
  B2S_TRY(a.alloc(&h->cand_idx, P)); B2S_TRY(a.alloc(&h->cand_sc, P));
  h->cand_cap = (int)P;
  B2S_TRY(a.alloc(&h->sel_idx, K)); B2S_TRY(a.alloc(&h->sel_sc, K));
  B2S_TRY(a.alloc(&h->dk, (size_t)8)); B2S_TRY(a.alloc(&h->thr, (size_t)1));
  B2S_TRY(a.alloc(&h->kp_norm, 2 * K)); B2S_TRY(a.alloc(&h->disp, K)); B2S_TRY(a.alloc(&h->sampled, K));
  h->skws_floats = std::max<size_t>((size_t)8 * K * 128, (size_t)8 * P3 * 64);
  B2S_TRY(a.alloc(&h->skws, h->skws_floats));
  B2S_TRY(a.alloc(&h->Apatch, 3 * 1152 * K)); B2S_TRY(a.alloc(&h->toff1, 2 * M * K)); B2S_TRY(a.alloc(&h->offs, 2 * M * K));
  B2S_TRY(a.alloc(&h->S, 3 * M * 128 * K)); B2S_TRY(a.alloc(&h->F, 3 * M * 128 * K)); B2S_TRY(a.alloc(&h->descraw, 128 * K));
  // rows past the keypoint count are read by TMA (never stored): keep them finite
  B2S_CUDA(cudaMemset(h->Apatch, 0, 3 * 1152 * K * 2)); B2S_CUDA(cudaMemset(h->S, 0, 3 * M * 128 * K * 2));
  B2S_CUDA(cudaMemset(h->F, 0, 3 * M * 128 * K * 2));
  B2S_TRY(make_tmap_bf16_2d(&h->m_Apatch, h->Apatch, 1152, 3 * K, 1152 * 2, 64, 128));
  B2S_TRY(make_tmap_bf16_2d(&h->m_S, h->S, 128, 3 * K * M, 128 * 2, 64, 128));
  B2S_TRY(make_tmap_bf16_2d(&h->m_F, h->F, M * 128, 3 * K, M * 128 * 2, 64, 128));
  h->wsHp = Hp; h->wsWp = Wp;
  return 0;
}

// one DCN ResBlock on HWC maps; in [P][cin] -> out [P][cout].  Per deformable conv: one fused kernel (offset conv +
// deformable im2col) and one GEMM - on the tensor cores (bf16x3) for block3, on the CUDA cores with split-K for the
// 320-pixel block4 (3 row tiles would leave the tensor-core kernel latency-bound over K = 1152).
template <int CIN, int COUT>
int run_dcn_block(b2s_aliked* h, cudaStream_t st, const DcnBlockW& w, const float* in, int Hh, int Ww, float* col, float* ta,
                  float* res, float* out, bool use_tc) {
  const int P = Hh * Ww;
  const float clampv = (float)std::max(Hh, Ww) / 4.0f;
  auto gemm = [&](const float* A, int K, const float* W, int N, const float* bias, float* C, const float* residual, int act) {
    GemmParams g;
    g.A1 = A; g.lda1 = K; g.K1 = K; g.W = W; g.ldw = K; g.K = K; g.M = P; g.N = N; g.C = C; g.ldc = N;
    g.bias = bias; g.residual = residual; g.ldr = N; g.act = act;
    return agemm(h, g, st);
  };
  const int ppw = P >= 2048 ? 2 : 1;   // pixels per warp: 2 keeps >= 2 CTAs per SM busy at 1/8 resolution (4 was one under-filled wave)
  DcnColParams cp = {};
  cp.H = Hh; cp.W = Ww; cp.clampv = clampv; cp.ppw = ppw;
  if (use_tc) { cp.col_pl = h->col3_pl; cp.plane = (size_t)P * 9 * CIN; } else cp.col_f32 = col;
  // conv1 (+BN1) + SELU
  cp.in = in; cp.w_off = w.off1_w; cp.b_off = w.off1_b;
  launch_k(k_dcn_offcol<CIN>, cdiv(P, 8 * ppw), 256, 18 * 9 * CIN * sizeof(float), st, cp);
  ++h->launches; B2S_LAUNCH_CHECK();
  if (use_tc) B2S_TRY(tc_gemm(h, st, CIN == 32 ? h->m_col3a : h->m_col3b, P, w.reg1_tc, w.reg1_b, 1, P, nullptr, 0, ta, COUT, nullptr, 0));
  else B2S_TRY(gemm(col, 9 * CIN, w.reg1_w, COUT, w.reg1_b, ta, nullptr, ACT_SELU));
  // downsample(x) -> residual
  B2S_TRY(gemm(in, CIN, w.ds_w, COUT, w.ds_b, res, nullptr, ACT_NONE));
  // conv2 (+BN2) + residual + SELU
  cp.in = ta; cp.w_off = w.off2_w; cp.b_off = w.off2_b;
  if (use_tc) cp.plane = (size_t)P * 9 * COUT;
  launch_k(k_dcn_offcol<COUT>, cdiv(P, 8 * ppw), 256, 18 * 9 * COUT * sizeof(float), st, cp);
  ++h->launches; B2S_LAUNCH_CHECK();
  if (use_tc) return tc_gemm(h, st, COUT == 32 ? h->m_col3a : h->m_col3b, P, w.reg2_tc, w.reg2_b, 1, P, nullptr, 0, out, COUT, nullptr, 0, res);
  return gemm(col, 9 * COUT, w.reg2_w, COUT, w.reg2_b, out, res, ACT_SELU);
}

}  // namespace

extern "C" int b2s_aliked_create(const b2s_aliked_cfg* cfg, const void* weights, size_t nbytes, int device, b2s_aliked** out) {
  if (!cfg || !weights || !out) { set_error("b2s_aliked_create: null argument"); return B2S_EINVAL; }
  if (cfg->nms_radius != 2 || (cfg->model != 0 && cfg->model != 1) || cfg->precision != B2S_FP32) {
    set_error("b2s_aliked_create: unsupported cfg (nms_radius must be 2, model 0|1, precision fp32)");
    return B2S_EINVAL;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    set_error("b2s_aliked_create: CUDA device %d not available (%d devices) - there is no CPU fallback", device, ndev);
    return B2S_ENODEV;
  }
  B2S_CUDA(cudaSetDevice(device));
  WeightBlob wb;
  B2S_TRY(wb.parse(weights, nbytes));
  b2s_aliked* h = new b2s_aliked();
  h->cfg = *cfg; h->device = device;
  h->M = cfg->model == 1 ? 32 : 16;
  h->n_limit = cfg->max_kp > 0 ? cfg->max_kp : 20000;
  cudaFuncSetAttribute(k_gemm_tc<64, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcGemmCfg<64, 3>::SMEM);
  cudaFuncSetAttribute(k_gemm_tcp<64, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcpGemmCfg<64, 3>::SMEM);
  cudaFuncSetAttribute(k_dcn_offcol<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 18 * 9 * 64 * (int)sizeof(float));
  cudaFuncSetAttribute(k_dcn_offcol<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 18 * 9 * 128 * (int)sizeof(float));
  cudaFuncSetAttribute(k_conv3x3_tc<16, 16, 3, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvTcCfg<16, 16, 3, 4>::SMEM);
  cudaFuncSetAttribute(k_conv3x3_tc<16, 32, 3, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvTcCfg<16, 32, 3, 4>::SMEM);
  cudaFuncSetAttribute(k_conv3x3_tc<32, 32, 3, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvTcCfg<32, 32, 3, 4>::SMEM);
  cudaFuncSetAttribute(k_conv3x3_tc<16, 16, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvTcCfg<16, 16, 2, 4>::SMEM);
  cudaFuncSetAttribute(k_conv3x3_tc<16, 32, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvTcCfg<16, 32, 2, 4>::SMEM);
  cudaFuncSetAttribute(k_conv3x3_tc<32, 32, 2, CONV22_R>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvTcCfg<32, 32, 2, CONV22_R>::SMEM);
  {
    // implicit-GEMM convolutions on two fp16 planes (default) or three bf16 planes (B2S_ALIKED_CONV_NP=3: for weights whose
    // activations leave the fp16 range - an extraction that does reports B2S_ALIKED_RANGE keypoints and the host raises)
    const char* e = std::getenv("B2S_ALIKED_CONV_NP");
    h->conv_np = (e && e[0] == '3') ? 3 : 2;
    std::vector<int> z(1, 0);
    if (int r0 = h->warena.upload(&h->range_flag, z)) { delete h; return r0; }
  }
  int rc = load_weights(h, wb);
  if (rc) { delete h; return rc; }
  h->blob.assign(static_cast<const uint8_t*>(weights), static_cast<const uint8_t*>(weights) + nbytes);
  {
    std::vector<float> t0(1, h->cfg.det_thresh);
    if ((rc = h->warena.upload(&h->thr0, t0))) { delete h; return rc; }
  }
  *out = h;
  return 0;
}

extern "C" void b2s_aliked_destroy(b2s_aliked* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->gexec) cudaGraphExecDestroy(h->gexec);
  for (b2s_aliked* l : h->lanes) b2s_aliked_destroy(l);
  for (cudaStream_t s : h->lane_st) cudaStreamDestroy(s);
  for (cudaEvent_t e : h->lane_ev) cudaEventDestroy(e);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->pin_kp) cudaFreeHost(h->pin_kp);
  if (h->pin_n) cudaFreeHost(h->pin_n);
  if (h->ev_kp) cudaEventDestroy(h->ev_kp);
  delete h;
}

extern "C" long long b2s_aliked_launch_count(const b2s_aliked* h) {
  if (!h) return 0;
  long long n = h->launches;
  for (const b2s_aliked* l : h->lanes) n += l->launches;
  return n;
}

extern "C" int b2s_aliked_extract(b2s_aliked* h, const void* img, int fmt, int H, int W, int row_stride, void* stream,
                                  float* kpts, float* desc, float* scores, int32_t* n_out) {
  if (!h || !img || !kpts || !desc || !n_out || H < 8 || W < 8 || (fmt != B2S_IMG_BGR_U8_HWC && fmt != B2S_IMG_RGB_F32_CHW)) {
    set_error("b2s_aliked_extract: bad argument (H=%d W=%d fmt=%d)", H, W, fmt);
    return B2S_EINVAL;
  }
  B2S_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  // ---- geometry: kornia resize(side='long') then InputPadder(32) ----
  int Hr = H, Wr = W;
  if (h->cfg.resize_long > 0) {
    const double ar = (double)W / (double)H;
    if (ar >= 1.0) { Hr = (int)((double)h->cfg.resize_long / ar); Wr = h->cfg.resize_long; }
    else { Hr = h->cfg.resize_long; Wr = (int)((double)h->cfg.resize_long * ar); }
  }
  if (Hr < 8 || Wr < 8) { set_error("b2s_aliked_extract: resized image %dx%d too small", Hr, Wr); return B2S_ESIZE; }
  PreParams pp;
  pp.img = img; pp.fmt = fmt; pp.H = H; pp.W = W; pp.stride = row_stride > 0 ? row_stride : 3 * W;
  pp.Hr = Hr; pp.Wr = Wr;
  pp.do_resize = (Hr != H || Wr != W);
  pp.do_blur = 0; pp.ky = pp.kx = 1;
  if (pp.do_resize) {
    const double fy = (double)H / Hr, fx = (double)W / Wr;
    if (std::max(fy, fx) > 1.0) {
      const double sy = std::max((fy - 1.0) / 2.0, 0.001), sx = std::max((fx - 1.0) / 2.0, 0.001);
      int ky = (int)std::max(2.0 * 2 * sy, 3.0), kx = (int)std::max(2.0 * 2 * sx, 3.0);
      if (ky % 2 == 0) ++ky;
      if (kx % 2 == 0) ++kx;
      if (ky > 15 || kx > 15) { set_error("b2s_aliked_extract: downscale factor too large (blur kernel %dx%d > 15)", ky, kx); return B2S_ESIZE; }
      pp.do_blur = 1; pp.ky = ky; pp.kx = kx;
      auto mk = [](int ks, double sigma, float* g) {
        const float denom = (float)(2.0 * sigma * sigma);
        float s = 0.f;
        for (int i = 0; i < ks; ++i) { const float x = (float)(i - ks / 2); g[i] = std::exp(-(x * x) / denom); s += g[i]; }
        for (int i = 0; i < ks; ++i) g[i] /= s;
      };
      mk(ky, sy, pp.gy); mk(kx, sx, pp.gx);
    }
  }
  pp.scale_h = (float)H / (float)Hr; pp.scale_w = (float)W / (float)Wr;
  const int pad_h = (((Hr / 32) + 1) * 32 - Hr) % 32, pad_w = (((Wr / 32) + 1) * 32 - Wr) % 32;
  const int Hp = Hr + pad_h, Wp = Wr + pad_w;
  pp.Hp = Hp; pp.Wp = Wp; pp.pad_t = pad_h / 2; pp.pad_l = pad_w / 2;
  if ((size_t)Hp * Wp > (size_t)h->wsHp * h->wsWp) {
    B2S_CUDA(cudaStreamSynchronize(st));
    B2S_TRY(alloc_ws(h, Hp, Wp));
  }
  h->Hr = Hr; h->Wr = Wr; h->Hp = Hp; h->Wp = Wp;
  pp.out = h->img_pad; pp.resized = h->resized; pp.range_flag = h->range_flag;
  launch_k(k_preprocess, dim3(cdiv(Wp, 256), Hp), 256, 0, st, pp);
  ++h->launches; B2S_LAUNCH_CHECK();

  // ---- block1 (full res) and block2 (1/2 res) ----
  // conv1 (3 -> 16, K = 27) stays on the CUDA cores and emits tensor-core operand planes; the three
  // 16/32-channel convs are implicit GEMMs on tcgen05 (conv_tc.cuh).  pool2 is fused into block1.conv2's
  // epilogue (planes for block2.conv1) and into the 1x1 downsample.
  const int H2 = Hp / 2, W2 = Wp / 2;
  if (h->mapHp != Hp || h->mapWp != Wp) {
    auto mk = [&](CUtensorMap* m, const __nv_bfloat16* ptr, int Hh, int Ww, int nch, int rows) -> int {
      const uint64_t dims[4] = {8, (uint64_t)Ww, (uint64_t)Hh, (uint64_t)h->conv_np * nch};
      const uint64_t str[3] = {16, (uint64_t)Ww * 16, (uint64_t)Hh * Ww * 16};
      const uint32_t box[4] = {8, 130, (uint32_t)rows + 2, (uint32_t)nch};
      return make_tmap_bf16_4d(m, ptr, dims, str, box);
    };
    B2S_TRY(mk(&h->m_t1a, h->t1a_pl, Hp, Wp, 2, 4));
    B2S_TRY(mk(&h->m_x1p, h->x1p_pl, H2, W2, 2, 4));
    B2S_TRY(mk(&h->m_t2a, h->t2a_pl, H2, W2, 4, h->conv_np == 2 ? CONV22_R : 4));
    const uint64_t P3 = (uint64_t)(H2 / 4) * (W2 / 4);
    B2S_TRY(make_tmap_bf16_2d(&h->m_col3a, h->col3_pl, 288, 3 * P3, 288 * 2, 64, 128));
    B2S_TRY(make_tmap_bf16_2d(&h->m_col3b, h->col3_pl, 576, 3 * P3, 576 * 2, 64, 128));
    h->mapHp = Hp; h->mapWp = Wp;
  }
  {
    dim3 g(cdiv(Wp, 32), cdiv(Hp, 16));
    launch_k(k_conv3x3<3, 16, false>, g, dim3(16, 8, 1), 0, st, h->img_pad, Hp, Wp, h->b1c1_w, h->b1c1_b, nullptr, (float*)nullptr, 1, h->t1a_pl,
             h->conv_np, h->range_flag);
    const bool h2 = h->conv_np == 2;
    ConvTcParams cp = {};
    cp.H = Hp; cp.W = Wp; cp.wplanes = h->b1c2_pl; cp.bias = h->b1c2_b; cp.out_chw = h->x1; cp.out_pooled = h->x1p_pl; cp.range_flag = h->range_flag;
    if (h2) launch_k(k_conv3x3_tc<16, 16, 2, 4>, dim3(cdiv(Wp, 128), cdiv(Hp, 4)), 192, ConvTcCfg<16, 16, 2, 4>::SMEM, st, h->m_t1a, cp);
    else launch_k(k_conv3x3_tc<16, 16, 3, 4>, dim3(cdiv(Wp, 128), cdiv(Hp, 4)), 192, ConvTcCfg<16, 16, 3, 4>::SMEM, st, h->m_t1a, cp);
    launch_k(k_pool2_conv1x1<16, 32>, cdiv(H2 * W2, 64), 256, 0, st, h->x1, H2, W2, h->b2ds_w, h->b2ds_b, h->r2);
    cp = ConvTcParams();
    cp.H = H2; cp.W = W2; cp.wplanes = h->b2c1_pl; cp.bias = h->b2c1_b; cp.out_planes = h->t2a_pl; cp.range_flag = h->range_flag;
    if (h2) launch_k(k_conv3x3_tc<16, 32, 2, 4>, dim3(cdiv(W2, 128), cdiv(H2, 4)), 192, ConvTcCfg<16, 32, 2, 4>::SMEM, st, h->m_x1p, cp);
    else launch_k(k_conv3x3_tc<16, 32, 3, 4>, dim3(cdiv(W2, 128), cdiv(H2, 4)), 192, ConvTcCfg<16, 32, 3, 4>::SMEM, st, h->m_x1p, cp);
    cp = ConvTcParams();
    cp.H = H2; cp.W = W2; cp.wplanes = h->b2c2_pl; cp.bias = h->b2c2_b; cp.residual = h->r2; cp.out_chw = h->x2;
    if (h2) launch_k(k_conv3x3_tc<32, 32, 2, CONV22_R>, dim3(cdiv(W2, 128), cdiv(H2, CONV22_R)), 192, ConvTcCfg<32, 32, 2, CONV22_R>::SMEM, st, h->m_t2a, cp);
    else launch_k(k_conv3x3_tc<32, 32, 3, 4>, dim3(cdiv(W2, 128), cdiv(H2, 4)), 192, ConvTcCfg<32, 32, 3, 4>::SMEM, st, h->m_t2a, cp);
    h->launches += 5; B2S_LAUNCH_CHECK();
  }
  // ---- block3 (1/8) and block4 (1/32): DCN via im2col + GEMM, HWC ----
  const int H3 = H2 / 4, W3 = W2 / 4, H4 = H3 / 4, W4 = W3 / 4;
  launch_k(k_pool4_chw_to_hwc, cdiv(H3 * W3, 8), 256, 0, st, h->x2, 32, H3, W3, h->x3in);
  ++h->launches; B2S_LAUNCH_CHECK();
  B2S_TRY((run_dcn_block<32, 64>(h, st, h->b3, h->x3in, H3, W3, h->col3, h->t3a, h->r3, h->x3, true)));
  launch_k(k_pool4_hwc, cdiv(H4 * W4, 8), 256, 0, st, h->x3, 64, H4, W4, h->x4in);
  ++h->launches; B2S_LAUNCH_CHECK();
  B2S_TRY((run_dcn_block<64, 128>(h, st, h->b4, h->x4in, H4, W4, h->col4, h->t4a, h->r4, h->x4, false)));
  // ---- aggregation convs at native resolution ----
  launch_k(k_conv1x1_chw_to_hwc32<32>, cdiv(H2 * W2, 64), 256, 0, st, h->x2, (size_t)H2 * W2, h->agg_w[1], h->x2a);
  ++h->launches; B2S_LAUNCH_CHECK();
  {
    GemmParams g;
    g.A1 = h->x3; g.lda1 = 64; g.K1 = 64; g.W = h->agg_w[2]; g.ldw = 64; g.K = 64; g.M = H3 * W3; g.N = 32; g.C = h->x3a; g.ldc = 32; g.act = ACT_SELU;
    B2S_TRY(agemm(h, g, st));
    g.A1 = h->x4; g.lda1 = 128; g.K1 = 128; g.W = h->agg_w[3]; g.ldw = 128; g.K = 128; g.M = H4 * W4; g.C = h->x4a;
    B2S_TRY(agemm(h, g, st));
  }
  // ---- score head: level projections at native resolution, fused 1x1 stage, 3x3 tail ----
  FeatSrc& fs = h->fsrc;
  {
    fs.x1 = h->x1; fs.Hp = Hp; fs.Wp = Wp;
    fs.xa[0] = h->x2a; fs.xa[1] = h->x3a; fs.xa[2] = h->x4a;
    const int hk[3] = {H2, H3, H4}, wk[3] = {W2, W3, W4};
    Proj8Params pj;
    for (int l = 0; l < 3; ++l) {
      fs.Hk[l] = hk[l]; fs.Wk[l] = wk[l];
      fs.sh[l] = Hp > 1 ? (float)(hk[l] - 1) / (float)(Hp - 1) : 0.f;
      fs.sw[l] = Wp > 1 ? (float)(wk[l] - 1) / (float)(Wp - 1) : 0.f;
      pj.xa[l] = fs.xa[l]; pj.out[l] = h->p8[l]; pj.npix[l] = hk[l] * wk[l];
    }
    fs.W1 = h->agg_w[0]; fs.Hr = Hr; fs.Wr = Wr; fs.pad_t = pp.pad_t; fs.pad_l = pp.pad_l;
    pj.Ws0 = h->sh0;
    launch_k(k_aliked_proj8, cdiv(pj.npix[0] + pj.npix[1] + pj.npix[2], 256), 256, 0, st, pj);
    S8Params s8p;
    s8p.src = fs; s8p.P[0] = h->p8[0]; s8p.P[1] = h->p8[1]; s8p.P[2] = h->p8[2]; s8p.Ws0 = h->sh0; s8p.s8 = h->s8;
    launch_k(k_aliked_s8, dim3(cdiv(Wp, 128), Hp), 128, 0, st, s8p);
    ScoreParams sp;
    sp.s8 = h->s8; sp.Hp = Hp; sp.Wp = Wp; sp.w2 = h->sh2; sp.w4 = h->sh4; sp.w6 = h->sh6;
    sp.score = h->score; sp.Hr = Hr; sp.Wr = Wr; sp.pad_t = pp.pad_t; sp.pad_l = pp.pad_l;
    launch_k(k_aliked_score, dim3(cdiv(Wp, 32), cdiv(Hp, SCORE_TH)), 256, 0, st, sp);
    h->launches += 3; B2S_LAUNCH_CHECK();
  }
  // ---- DKD ----
  {
    B2S_CUDA(cudaMemsetAsync(h->dk, 0, 8 * sizeof(int), st));
    B2S_CUDA(cudaMemcpyAsync(h->thr, h->thr0, sizeof(float), cudaMemcpyDeviceToDevice, st));   // device source: capturable in a CUDA graph
    launch_k(k_dkd_nms, dim3(cdiv(Wr, NMS_T), cdiv(Hr, NMS_T)), 256, 0, st, h->score, Hr, Wr, h->nms, h->thr, h->dk, h->cand_idx, h->cand_sc, h->cand_cap);
    launch_k(k_dkd_fallback, 1, 1024, 0, st, h->score, h->nms, Hr * Wr, h->thr, h->dk, h->cand_idx, h->cand_sc, h->cand_cap);
    launch_k(k_dkd_select, 1, 1024, 0, st, h->cand_sc, h->cand_idx, h->cand_cap, h->n_limit, h->dk);
    launch_k(k_dkd_compact, cdiv(h->cand_cap, 256), 256, 0, st, h->cand_idx, h->cand_sc, h->dk, h->sel_idx, h->sel_sc);
    RefineParams rp;
    rp.dk = h->dk; rp.sel_idx = h->sel_idx; rp.sel_sc = h->sel_sc; rp.score = h->score; rp.H = Hr; rp.W = Wr;
    rp.scale_x = (float)((double)Wr / (double)W); rp.scale_y = (float)((double)Hr / (double)H);
    rp.kp_norm = h->kp_norm; rp.kp_out = kpts; rp.disp = h->disp; rp.sampled = h->sampled; rp.n_out = n_out; rp.range_flag = h->conv_np == 2 ? h->range_flag : nullptr;
    launch_k(k_dkd_refine, cdiv(h->n_limit, RF_KP), 256, 0, st, rp);
    h->launches += 5; B2S_LAUNCH_CHECK();
    // upstream puts DKD's 2nd return value (dispersity) under "keypoint_scores" 
    if (scores) B2S_CUDA(cudaMemcpyAsync(scores, h->disp, (size_t)h->n_limit * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (h->early_kp) {   // split host API: keypoints are final here, the descriptor head still has to run
      B2S_CUDA(cudaMemcpyAsync(h->pin_n, n_out, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      B2S_CUDA(cudaMemcpyAsync(h->pin_kp, kpts, (size_t)h->n_limit * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
      B2S_CUDA(cudaEventRecord(h->ev_kp, st));
    }
  }
  // ---- SDDH: gathers evaluate the feature on demand and emit bf16x3 planes; the three large
  //      contractions (offset conv K=1152, sf_conv, aggregation K=M*128) run on tcgen05 ----
  {
    const int K = h->n_limit, M = h->M;
    const float clampv = (float)std::max(Hr, Wr) / 4.0f;
    launch_k(k_sddh_patch, cdiv(K * 9, 8), 256, 0, st, fs, h->kp_norm, n_out, h->Apatch, (size_t)K * 1152);
    ++h->launches; B2S_LAUNCH_CHECK();
    B2S_TRY(tc_gemm(h, st, h->m_Apatch, K, h->tc_so0, h->so0_b, 1, K, n_out, 1, h->toff1, 2 * M, nullptr, 0));
    GemmParams g;
    g.A1 = h->toff1; g.lda1 = 2 * M; g.K1 = 2 * M; g.W = h->so2_w; g.ldw = 2 * M; g.K = 2 * M; g.M = K; g.N = 2 * M;
    g.C = h->offs; g.ldc = 2 * M; g.bias = h->so2_b; g.clamp = clampv; g.m_dev = n_out; g.m_mult = 1;
    B2S_TRY(agemm(h, g, st));
    launch_k(k_sddh_sample, cdiv(K * M, 8), 256, 0, st, fs, h->kp_norm, h->offs, M, n_out, h->S, (size_t)K * M * 128);
    ++h->launches; B2S_LAUNCH_CHECK();
    B2S_TRY(tc_gemm(h, st, h->m_S, K * M, h->tc_sf, nullptr, 1, K * M, n_out, M, nullptr, 128, h->F, (size_t)K * M * 128));
    B2S_TRY(tc_gemm(h, st, h->m_F, K, h->tc_agg, nullptr, 0, K, n_out, 1, h->descraw, 128, nullptr, 0));
    launch_k(k_desc_normalize, cdiv(K, 8), 256, 0, st, h->descraw, n_out, desc, h->desc_renorm_eps);
    ++h->launches; B2S_LAUNCH_CHECK();
  }
  return 0;
}

// One extraction as a CUDA-graph replay: image -> staging, graph (all kernels of b2s_aliked_extract over the staging
// buffers), results -> the caller's buffers.  Falls back to the eager launch sequence if capture is not possible.
static int aliked_extract_graphed(b2s_aliked* h, const void* img, int fmt, int H, int W, int row_stride, cudaStream_t st,
                                  float* kpts, float* desc, float* scores, int32_t* n_out) {
  static const bool off = [] { const char* e = std::getenv("B2S_ALIKED_GRAPH"); return e && e[0] == '0'; }();
  // the legacy default stream cannot be captured: lane 0 of a caller that works on it stays on the eager launch sequence
  const bool legacy = st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread;
  if (off || legacy || h->g_failed) return b2s_aliked_extract(h, img, fmt, H, W, row_stride, st, kpts, desc, scores, n_out);
  const size_t bytes = fmt == B2S_IMG_BGR_U8_HWC ? (size_t)(row_stride > 0 ? row_stride : 3 * W) * H : (size_t)3 * H * W * sizeof(float);
  if (!h->gexec || h->g_fmt != fmt || h->g_H != H || h->g_W != W || h->g_stride != row_stride || h->g_eps != h->desc_renorm_eps) {
    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
    if (bytes > h->g_img_bytes || !h->g_kp) {
      B2S_CUDA(cudaStreamSynchronize(st));
      h->grarena.release();
      h->g_img_bytes = 0;
      B2S_TRY(h->grarena.alloc(&h->g_img, bytes));
      B2S_TRY(h->grarena.alloc(&h->g_kp, (size_t)h->n_limit * 2));
      B2S_TRY(h->grarena.alloc(&h->g_desc, (size_t)h->n_limit * 128));
      B2S_TRY(h->grarena.alloc(&h->g_sc, (size_t)h->n_limit));
      B2S_TRY(h->grarena.alloc(&h->g_n, (size_t)1));
      h->g_img_bytes = bytes;
    }
    // eager warm-up call: sizes the workspace and the tensor maps (both synchronise / allocate: not capturable)
    B2S_CUDA(cudaMemcpyAsync(h->g_img, img, bytes, cudaMemcpyDeviceToDevice, st));
    B2S_TRY(b2s_aliked_extract(h, h->g_img, fmt, H, W, row_stride, st, h->g_kp, h->g_desc, h->g_sc, h->g_n));
    B2S_CUDA(cudaStreamSynchronize(st));
    const long long l0 = h->launches;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    int rc = 0;
    if (e == cudaSuccess) {
      rc = b2s_aliked_extract(h, h->g_img, fmt, H, W, row_stride, st, h->g_kp, h->g_desc, h->g_sc, h->g_n);
      e = cudaStreamEndCapture(st, &graph);
    }
    if (e == cudaSuccess && rc == 0 && graph) e = cudaGraphInstantiate(&h->gexec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (e != cudaSuccess || rc != 0 || !h->gexec) {
      cudaGetLastError();
      h->gexec = nullptr; h->g_failed = true;          // this driver cannot capture the sequence: stay eager
      return b2s_aliked_extract(h, img, fmt, H, W, row_stride, st, kpts, desc, scores, n_out);
    }
    h->g_launches = h->launches - l0; h->launches = l0;
    h->g_fmt = fmt; h->g_H = H; h->g_W = W; h->g_stride = row_stride; h->g_eps = h->desc_renorm_eps;
  }
  B2S_CUDA(cudaMemcpyAsync(h->g_img, img, bytes, cudaMemcpyDeviceToDevice, st));
  B2S_CUDA(cudaGraphLaunch(h->gexec, st));
  h->launches += h->g_launches;
  const size_t nl = (size_t)h->n_limit;
  B2S_CUDA(cudaMemcpyAsync(kpts, h->g_kp, nl * 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  B2S_CUDA(cudaMemcpyAsync(desc, h->g_desc, nl * 128 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (scores) B2S_CUDA(cudaMemcpyAsync(scores, h->g_sc, nl * sizeof(float), cudaMemcpyDeviceToDevice, st));
  B2S_CUDA(cudaMemcpyAsync(n_out, h->g_n, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  return 0;
}

// Batched extraction of B same-shape frames.  The extractor's ~60 kernels per frame are small (a 320 x 1024 frame does
// not fill 148 SMs) and latency bound, so the frames of a batch are spread over n_lanes complete extractors (own
// workspace, own stream) that run CONCURRENTLY: `stream` forks into the lane streams and joins them again, nothing
// synchronises with the host.  Each lane replays its extraction as a CUDA graph (one launch per frame instead of ~60:
// the host could not issue the launches of four concurrent lanes fast enough).  Outputs are [B, max_kp, .] slabs;
// n_out_dev [B].
extern "C" int b2s_aliked_extract_batch(b2s_aliked* h, const void* const* imgs_dev, int B, int fmt, int H, int W, int row_stride,
                                        int n_lanes, void* stream, float* kpts, float* desc, float* scores, int32_t* n_out) {
  if (!h || !imgs_dev || B < 0 || !kpts || !desc || !n_out) { set_error("b2s_aliked_extract_batch: bad argument"); return B2S_EINVAL; }
  if (B == 0) return 0;
  B2S_CUDA(cudaSetDevice(h->device));
  const int want = std::max(1, std::min(std::min(n_lanes > 0 ? n_lanes : 4, B), 8));
  while ((int)h->lanes.size() < want - 1) {      // lane 0 is this handle on the caller's stream
    b2s_aliked* l = nullptr;
    B2S_TRY(b2s_aliked_create(&h->cfg, h->blob.data(), h->blob.size(), h->device, &l));
    l->desc_renorm_eps = h->desc_renorm_eps;
    h->lanes.push_back(l);
    cudaStream_t s; cudaEvent_t e;
    B2S_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    B2S_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->lane_st.push_back(s); h->lane_ev.push_back(e);
  }
  if (!h->ev_fork) B2S_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nl = (size_t)h->n_limit;
  B2S_CUDA(cudaEventRecord(h->ev_fork, st));
  for (int l = 1; l < want; ++l) B2S_CUDA(cudaStreamWaitEvent(h->lane_st[l - 1], h->ev_fork, 0));
  for (int i = 0; i < B; ++i) {
    const int l = i % want;
    b2s_aliked* lane = l == 0 ? h : h->lanes[l - 1];
    lane->desc_renorm_eps = h->batch_renorm_eps;
    const int rc = aliked_extract_graphed(lane, imgs_dev[i], fmt, H, W, row_stride, l == 0 ? st : h->lane_st[l - 1], kpts + i * nl * 2,
                                          desc + i * nl * 128, scores ? scores + i * nl : nullptr, n_out + i);
    lane->desc_renorm_eps = 0.f;
    B2S_TRY(rc);
  }
  for (int l = 1; l < want; ++l) {
    B2S_CUDA(cudaEventRecord(h->lane_ev[l - 1], h->lane_st[l - 1]));
    B2S_CUDA(cudaStreamWaitEvent(st, h->lane_ev[l - 1], 0));
  }
  return 0;
}

extern "C" int b2s_aliked_extract_host_ex(b2s_aliked* h, const void* img, int fmt, int H, int W, int row_stride,
                                          float* kpts, float* desc, float* scores, int32_t* n_out, float desc_renorm_eps);

// eps > 0: b2s_aliked_extract_batch also applies the reference caller's `des /= (||des||_2 + eps)` (features_utils.py:100,
// eps = 1e-8) in its last kernel, like b2s_aliked_extract_host_ex does for single frames; 0 (default) switches it off
extern "C" int b2s_aliked_set_batch_renorm(b2s_aliked* h, float eps) {
  if (!h || !(eps >= 0.f)) { set_error("b2s_aliked_set_batch_renorm: bad argument"); return B2S_EINVAL; }
  h->batch_renorm_eps = eps;
  return 0;
}

static int aliked_host_staging(b2s_aliked* h, size_t bytes) {
  if (bytes > h->himg_bytes || !h->hkp) {
    h->hostarena.release();
    h->himg_bytes = 0;
    uint8_t* b = nullptr;
    B2S_TRY(h->hostarena.alloc(&b, bytes));
    h->himg = b;
    B2S_TRY(h->hostarena.alloc(&h->hkp, (size_t)h->n_limit * 2));
    B2S_TRY(h->hostarena.alloc(&h->hdesc, (size_t)h->n_limit * 128));
    B2S_TRY(h->hostarena.alloc(&h->hscores, (size_t)h->n_limit));
    B2S_TRY(h->hostarena.alloc(&h->hn, (size_t)1));
    h->himg_bytes = bytes;
  }
  return 0;
}

extern "C" int b2s_aliked_set_undistort(b2s_aliked* h, b2s_remap* r) {
  if (!h) { set_error("b2s_aliked_set_undistort: null handle"); return B2S_EINVAL; }
  if (h->pending) { set_error("b2s_aliked_set_undistort: an extraction is pending"); return B2S_EINVAL; }
  h->undist = r;
  return 0;
}


// n_out of an extraction whose two-plane (fp16) convolutions saw an activation beyond the fp16 range (k_dkd_refine)
static constexpr int32_t ALIKED_RANGE = -2;
static int aliked_range_error() {
  set_error("b2s_aliked: an activation left the fp16 range of the fp16x2 convolutions - re-create the extractor with B2S_ALIKED_CONV_NP=3 (three bf16 planes)");
  return B2S_ERANGE;
}

// host-API ingest: the uploaded frame goes through the attached cv2.remap stage (u8 BGR only) before K0
static int aliked_ingest(b2s_aliked* h, int fmt, int* H, int* W, int* row_stride, cudaStream_t st, const void** img_dev) {
  *img_dev = h->himg;
  if (!h->undist) return 0;
  int dh, dw, sh, sw;
  b2s_remap_dims(h->undist, &dh, &dw, &sh, &sw);
  if (fmt != B2S_IMG_BGR_U8_HWC || *H != sh || *W != sw) {
    set_error("b2s_aliked: the attached undistortion stage expects %dx%d u8 BGR frames, got %dx%d fmt %d", sh, sw, *H, *W, fmt);
    return B2S_EINVAL;
  }
  uint8_t* dst = const_cast<uint8_t*>(b2s_remap_output_dev(h->undist));
  B2S_TRY(b2s_remap_bgr(h->undist, static_cast<const uint8_t*>(h->himg), *row_stride > 0 ? *row_stride : 3 * sw, st, dst, 3 * dw));
  ++h->launches;
  *img_dev = dst; *H = dh; *W = dw; *row_stride = 3 * dw;
  return 0;
}

// Split form of b2s_aliked_extract_host_ex: begin enqueues the whole extraction and returns; keypoints blocks only
// until the detector (DKD) has finished - the descriptor head is still running - so the caller can build its keypoint
// objects meanwhile; finish waits for the descriptors.  One extraction may be pending per handle.
extern "C" int b2s_aliked_extract_host_begin(b2s_aliked* h, const void* img, int fmt, int H, int W, int row_stride, float desc_renorm_eps) {
  if (!h || !img) { set_error("b2s_aliked_extract_host_begin: null argument"); return B2S_EINVAL; }
  if (h->pending) { set_error("b2s_aliked_extract_host_begin: an extraction is already pending on this handle"); return B2S_EINVAL; }
  B2S_CUDA(cudaSetDevice(h->device));
  const size_t bytes = fmt == B2S_IMG_BGR_U8_HWC ? (size_t)(row_stride > 0 ? row_stride : 3 * W) * H : (size_t)3 * H * W * sizeof(float);
  B2S_TRY(aliked_host_staging(h, bytes));
  if (!h->pin_kp) {
    B2S_CUDA(cudaMallocHost((void**)&h->pin_kp, (size_t)h->n_limit * 2 * sizeof(float)));
    B2S_CUDA(cudaMallocHost((void**)&h->pin_n, sizeof(int32_t)));
    B2S_CUDA(cudaEventCreateWithFlags(&h->ev_kp, cudaEventDisableTiming));
  }
  cudaStream_t st = 0;
  B2S_CUDA(cudaMemcpyAsync(h->himg, img, bytes, cudaMemcpyHostToDevice, st));
  const void* src = nullptr;
  B2S_TRY(aliked_ingest(h, fmt, &H, &W, &row_stride, st, &src));
  h->desc_renorm_eps = desc_renorm_eps; h->early_kp = true;
  const int rc = b2s_aliked_extract(h, src, fmt, H, W, row_stride, st, h->hkp, h->hdesc, h->hscores, h->hn);
  h->desc_renorm_eps = 0.f; h->early_kp = false;
  B2S_TRY(rc);
  h->pending = true;
  return 0;
}

extern "C" int b2s_aliked_extract_host_keypoints(b2s_aliked* h, float* kpts, int32_t* n_out) {
  if (!h || !kpts || !n_out || !h->pending) { set_error("b2s_aliked_extract_host_keypoints: no pending extraction"); return B2S_EINVAL; }
  B2S_CUDA(cudaSetDevice(h->device));
  B2S_CUDA(cudaEventSynchronize(h->ev_kp));
  const int32_t n = *h->pin_n;
  *n_out = n > 0 ? n : 0;
  if (n == ALIKED_RANGE) return aliked_range_error();
  if (n > 0) std::memcpy(kpts, h->pin_kp, (size_t)n * 2 * sizeof(float));
  return 0;
}

extern "C" int b2s_aliked_extract_host_finish(b2s_aliked* h, float* desc, float* scores) {
  if (!h || !desc || !h->pending) { set_error("b2s_aliked_extract_host_finish: no pending extraction"); return B2S_EINVAL; }
  B2S_CUDA(cudaSetDevice(h->device));
  h->pending = false;
  cudaStream_t st = 0;
  B2S_CUDA(cudaEventSynchronize(h->ev_kp));
  const int32_t n = *h->pin_n;
  if (n > 0) {
    B2S_CUDA(cudaMemcpyAsync(desc, h->hdesc, (size_t)n * 128 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (scores) B2S_CUDA(cudaMemcpyAsync(scores, h->hscores, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  B2S_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// Device-side copy of the LAST host extraction's results (n rows; valid until the next extraction on this handle) into
// caller-owned device buffers, enqueued on `stream`: lets a host-API caller keep the features it has just received on
// the GPU too, so that a following match of the same frame needs no re-upload (features_utils feature cache).
extern "C" int b2s_aliked_copy_last_features(b2s_aliked* h, float* kpts_dst_dev, float* desc_dst_dev, int n, void* stream) {
  if (!h || !kpts_dst_dev || !desc_dst_dev || n < 0 || n > h->n_limit || !h->hkp) { set_error("b2s_aliked_copy_last_features: bad argument or no host extraction yet"); return B2S_EINVAL; }
  if (h->pending) { set_error("b2s_aliked_copy_last_features: an extraction is pending"); return B2S_EINVAL; }
  B2S_CUDA(cudaSetDevice(h->device));
  if (n == 0) return 0;
  B2S_CUDA(cudaMemcpyAsync(kpts_dst_dev, h->hkp, (size_t)n * 2 * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  B2S_CUDA(cudaMemcpyAsync(desc_dst_dev, h->hdesc, (size_t)n * 128 * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

extern "C" int b2s_aliked_extract_host(b2s_aliked* h, const void* img, int fmt, int H, int W, int row_stride,
                                       float* kpts, float* desc, float* scores, int32_t* n_out) {
  return b2s_aliked_extract_host_ex(h, img, fmt, H, W, row_stride, kpts, desc, scores, n_out, 0.f);
}

extern "C" int b2s_aliked_extract_host_ex(b2s_aliked* h, const void* img, int fmt, int H, int W, int row_stride,
                                          float* kpts, float* desc, float* scores, int32_t* n_out, float desc_renorm_eps) {
  if (!h || !img || !kpts || !desc || !n_out) { set_error("b2s_aliked_extract_host: null argument"); return B2S_EINVAL; }
  B2S_CUDA(cudaSetDevice(h->device));
  const size_t bytes = fmt == B2S_IMG_BGR_U8_HWC ? (size_t)(row_stride > 0 ? row_stride : 3 * W) * H : (size_t)3 * H * W * sizeof(float);
  B2S_TRY(aliked_host_staging(h, bytes));
  cudaStream_t st = 0;
  B2S_CUDA(cudaMemcpyAsync(h->himg, img, bytes, cudaMemcpyHostToDevice, st));
  const void* src = nullptr;
  B2S_TRY(aliked_ingest(h, fmt, &H, &W, &row_stride, st, &src));
  h->desc_renorm_eps = desc_renorm_eps;
  const int rc_ext = b2s_aliked_extract(h, src, fmt, H, W, row_stride, st, h->hkp, h->hdesc, h->hscores, h->hn);
  h->desc_renorm_eps = 0.f;
  B2S_TRY(rc_ext);
  int32_t n = 0;
  B2S_CUDA(cudaMemcpyAsync(&n, h->hn, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  B2S_CUDA(cudaStreamSynchronize(st));
  *n_out = n > 0 ? n : 0;
  if (n == ALIKED_RANGE) return aliked_range_error();
  if (n > 0) {
    B2S_CUDA(cudaMemcpyAsync(kpts, h->hkp, (size_t)n * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
    B2S_CUDA(cudaMemcpyAsync(desc, h->hdesc, (size_t)n * 128 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (scores) B2S_CUDA(cudaMemcpyAsync(scores, h->hscores, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
    B2S_CUDA(cudaStreamSynchronize(st));
  }
  return 0;
}

extern "C" int b2s_aliked_debug_get(b2s_aliked* h, const char* name, float* out, size_t cap, size_t* nout) {
  if (!h || !name || !nout) return B2S_EINVAL;
  B2S_CUDA(cudaSetDevice(h->device));
  B2S_CUDA(cudaDeviceSynchronize());
  const size_t P = (size_t)h->Hp * h->Wp, Pr = (size_t)h->Hr * h->Wr;
  const std::string s(name);
  const float* src = nullptr; size_t cnt = 0;
  if (s == "geometry") {
    const float g[4] = {(float)h->Hr, (float)h->Wr, (float)h->Hp, (float)h->Wp};
    *nout = 4;
    std::memcpy(out, g, std::min<size_t>(cap, 4) * sizeof(float));
    return 0;
  }
  if (s == "padded") { src = h->img_pad; cnt = 3 * P; }            // CHW
  else if (s == "resized") { src = h->resized; cnt = 3 * Pr; }     // CHW
  else if (s == "x1") { src = h->x1; cnt = 16 * P; }               // CHW
  else if (s == "x2") { src = h->x2; cnt = 32 * P / 4; }           // CHW
  else if (s == "x3") { src = h->x3; cnt = 64 * P / 64; }          // HWC
  else if (s == "x4") { src = h->x4; cnt = 128 * P / 1024; }       // HWC
  else if (s == "score_map") { src = h->score; cnt = Pr; }
  else if (s == "nms") { src = h->nms; cnt = Pr; }
  else if (s == "feature_map") {   // HWC; never materialised by the extractor: evaluated here from the last call's maps
    *nout = 128 * Pr;
    if (!h->fsrc.x1) { set_error("debug tensor %s not available before the first extract", name); return B2S_EINVAL; }
    if (cap < 128 * Pr) return 0;
    float* tmp = nullptr;
    B2S_CUDA(cudaMalloc(&tmp, 128 * Pr * sizeof(float)));
    k_aliked_featmap<<<cdiv((int)Pr, 8), 256>>>(h->fsrc, tmp);
    cudaError_t e = cudaMemcpy(out, tmp, 128 * Pr * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(tmp);
    if (e != cudaSuccess) { set_error("feature_map tap: %s", cudaGetErrorString(e)); return B2S_ECUDA; }
    return 0;
  }
  else if (s == "kp_norm") { src = h->kp_norm; cnt = (size_t)h->n_limit * 2; }
  else if (s == "sampled_score") { src = h->sampled; cnt = (size_t)h->n_limit; }
  else if (s == "sddh_offset") { src = h->offs; cnt = (size_t)h->n_limit * 2 * h->M; }
  else if (s == "desc_raw") { src = h->descraw; cnt = (size_t)h->n_limit * 128; }
  else { set_error("unknown debug tensor %s", name); return B2S_EINVAL; }
  if (!src) { set_error("debug tensor %s not available before the first extract", name); return B2S_EINVAL; }
  *nout = cnt;
  const size_t c = std::min(cnt, cap);
  if (c) B2S_CUDA(cudaMemcpy(out, src, c * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}
