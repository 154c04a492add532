// lightglue.cu - host orchestration of the LightGlue matcher behind b2s_lightglue_*.
// Replaces `matcher({...})` at /root/reference/slam/core/features_utils.py:157-161 (the
// arithmetic itself lives in the un-vendored `lightglue` package; spec: SURVEY.md A.3).
#include "gemm_simt.cuh"
#include "lightglue_kernels.cuh"
#include "lightglue_tc.cuh"

#include <algorithm>
#include <cmath>

using namespace b2s;

struct LgLayer {
  float *wqkv, *bqkv, *wo, *bo, *w1, *b1, *lng, *lnb, *w2, *b2;            // self block
  float *cwqkv, *cbqkv, *cwo, *cbo, *cw1, *cb1, *clng, *clnb, *cw2, *cb2;  // cross block
  float *wtok; float btok;                                                 // token_confidence[i]
  float *wmatch; float bmatch; float *wfinal, *bfinal;                     // log_assignment[i]
};

struct b2s_lg {
  b2s_lg_cfg cfg;
  int device = 0;
  DeviceArena warena, wsarena;
  float *in_w = nullptr, *in_b = nullptr, *wr = nullptr;
  std::vector<LgLayer> L;
  std::vector<float> thr;
  // workspace (rows = 2*cap; image0 at row 0, image1 at row cap)
  int cap = 0;
  float *kn = nullptr, *cosb[2] = {nullptr, nullptr}, *sinb[2] = {nullptr, nullptr}, *x[2] = {nullptr, nullptr};
  float *qkv = nullptr, *ctx = nullptr, *msg = nullptr, *h1 = nullptr, *tok = nullptr, *sim = nullptr, *simT = nullptr;
  float *rmax = nullptr, *rlog = nullptr, *cmax = nullptr, *clog = nullptr, *ls = nullptr, *max0 = nullptr;
  int *ind[2] = {nullptr, nullptr}, *keep = nullptr, *srcmap = nullptr, *adapt = nullptr, *m0 = nullptr, *m1 = nullptr, *ctrl = nullptr;
  const float** wfinal_tab = nullptr; const float** bfinal_tab = nullptr; const float** wmatch_tab = nullptr;  // device tables [L]
  float* bmatch_tab = nullptr;
  int *prune_scratch[2] = {nullptr, nullptr};
  int* h_ctrl = nullptr;  // pinned
  // host-API staging
  DeviceArena hostarena;
  int hcap = 0;
  float *hk[2] = {nullptr, nullptr}, *hd[2] = {nullptr, nullptr}, *hms[2] = {nullptr, nullptr}, *hmscores = nullptr;
  int32_t *hmatches = nullptr, *hnm = nullptr, *hm[2] = {nullptr, nullptr}, *hprune[2] = {nullptr, nullptr};
  // debug
  int debug = 0;
  float* dbg_layers = nullptr;  // [n_layers][2*cap][256]
  int dbg_m[16], dbg_n[16], dbg_nlayers = 0, dbg_simm = 0, dbg_simn = 0;
  long long launches = 0;
  KernelProf prof;
  unsigned long long* stats = nullptr;   // device: [0] sum nq*nk over self-attention problems, [1] over cross-attention problems
  LgTensorCore* tc = nullptr;   // bf16 tcgen05 path (precision == B2S_BF16)
};

static int upload_t(b2s_lg* h, const WeightBlob& wb, const std::string& name, size_t numel, float** out) {
  const TensorView* t = wb.get(name, numel);
  if (!t) return B2S_EINVAL;
  std::vector<float> v(t->data, t->data + numel);
  return h->warena.upload(out, v);
}

static int lg_alloc_ws(b2s_lg* h, int cap) {
  cap = (cap + 127) / 128 * 128;
  h->wsarena.release();
  h->cap = 0;
  const size_t R = (size_t)2 * cap;
  B2S_TRY(h->wsarena.alloc(&h->kn, R * 2));
  for (int i = 0; i < 2; ++i) {
    B2S_TRY(h->wsarena.alloc(&h->cosb[i], R * 32));
    B2S_TRY(h->wsarena.alloc(&h->sinb[i], R * 32));
    B2S_TRY(h->wsarena.alloc(&h->x[i], R * 256));
    B2S_TRY(h->wsarena.alloc(&h->ind[i], R));
    B2S_TRY(h->wsarena.alloc(&h->prune_scratch[i], (size_t)cap));
  }
  B2S_TRY(h->wsarena.alloc(&h->qkv, R * 768));
  B2S_TRY(h->wsarena.alloc(&h->ctx, R * 256));
  B2S_TRY(h->wsarena.alloc(&h->msg, R * 256));
  B2S_TRY(h->wsarena.alloc(&h->h1, R * 512));
  B2S_TRY(h->wsarena.alloc(&h->tok, R));
  B2S_TRY(h->wsarena.alloc(&h->sim, (size_t)cap * cap));
  B2S_TRY(h->wsarena.alloc(&h->simT, (size_t)cap * cap));   // transposed similarity (tensor-core path)
  B2S_TRY(h->wsarena.alloc(&h->rmax, (size_t)cap));
  B2S_TRY(h->wsarena.alloc(&h->rlog, (size_t)cap));
  B2S_TRY(h->wsarena.alloc(&h->cmax, (size_t)cap));
  B2S_TRY(h->wsarena.alloc(&h->clog, (size_t)cap));
  B2S_TRY(h->wsarena.alloc(&h->ls, R));
  B2S_TRY(h->wsarena.alloc(&h->max0, (size_t)cap));
  B2S_TRY(h->wsarena.alloc(&h->keep, R));
  B2S_TRY(h->wsarena.alloc(&h->srcmap, R));
  B2S_TRY(h->wsarena.alloc(&h->adapt, (size_t)8 + 2 * LG_MAXBLK));
  B2S_TRY(h->wsarena.alloc(&h->m0, (size_t)cap));
  B2S_TRY(h->wsarena.alloc(&h->m1, (size_t)cap));
  B2S_TRY(h->wsarena.alloc(&h->ctrl, (size_t)LGC_INTS));
  h->dbg_layers = nullptr;
  if (h->debug) B2S_TRY(h->wsarena.alloc(&h->dbg_layers, (size_t)h->cfg.n_layers * R * 256));
  B2S_CUDA(cudaMemset(h->ctrl, 0, LGC_INTS * sizeof(int)));
  if (h->tc) B2S_TRY(lgtc_alloc_ws(h->tc, cap));
  h->cap = cap;
  return 0;
}

extern "C" void b2s_lg_default_cfg(b2s_lg_cfg* c) {
  c->n_layers = 9; c->heads = 4; c->dim = 256; c->in_dim = 128;
  c->depth_conf = 0.95f; c->width_conf = 0.99f; c->filter_thresh = 0.1f;
  c->pruning_min_kpts = -1; c->precision = B2S_FP32; c->max_kp = 2048;
}

extern "C" int b2s_lightglue_create(const b2s_lg_cfg* cfg, const void* weights, size_t nbytes, int device, b2s_lg** out) {
  if (!cfg || !weights || !out) { set_error("b2s_lightglue_create: null argument"); return B2S_EINVAL; }
  if (cfg->precision != B2S_FP32 && cfg->precision != B2S_BF16 && cfg->precision != B2S_FP32_SIMT) {
    set_error("b2s_lightglue_create: unknown precision %d", cfg->precision);
    return B2S_EINVAL;
  }
  if (cfg->dim != 256 || cfg->heads != 4 || cfg->in_dim != 128 || cfg->n_layers < 1 || cfg->n_layers > 16) {
    set_error("b2s_lightglue_create: only dim=256, heads=4, in_dim=128, 1..16 layers are supported");
    return B2S_EINVAL;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    set_error("b2s_lightglue_create: CUDA device %d not available (%d devices) - there is no CPU fallback", device, ndev);
    return B2S_ENODEV;
  }
  B2S_CUDA(cudaSetDevice(device));
  WeightBlob wb;
  B2S_TRY(wb.parse(weights, nbytes));
  b2s_lg* h = new b2s_lg();
  h->cfg = *cfg;
  h->device = device;
  int rc = 0;
  auto fail = [&](int r) { delete h; return r; };
  if ((rc = upload_t(h, wb, "input_proj.weight", 256 * 128, &h->in_w))) return fail(rc);
  if ((rc = upload_t(h, wb, "input_proj.bias", 256, &h->in_b))) return fail(rc);
  if ((rc = upload_t(h, wb, "posenc.Wr.weight", 64, &h->wr))) return fail(rc);
  h->L.resize(cfg->n_layers);
  for (int i = 0; i < cfg->n_layers; ++i) {
    LgLayer& l = h->L[i];
    const std::string p = "transformers." + std::to_string(i);
    // Wqkv: upstream row = head*192 + d*3 + {q,k,v}  ->  ours = {q,k,v}*256 + head*64 + d
    const TensorView* wq = wb.get(p + ".self_attn.Wqkv.weight", 768 * 256);
    const TensorView* bq = wb.get(p + ".self_attn.Wqkv.bias", 768);
    if (!wq || !bq) return fail(B2S_EINVAL);
    std::vector<float> w(768 * 256), b(768);
    for (int hh = 0; hh < 4; ++hh)
      for (int d = 0; d < 64; ++d)
        for (int t = 0; t < 3; ++t) {
          const int src = hh * 192 + d * 3 + t, dst = t * 256 + hh * 64 + d;
          std::memcpy(&w[(size_t)dst * 256], wq->data + (size_t)src * 256, 256 * sizeof(float));
          b[dst] = bq->data[src];
        }
    if ((rc = h->warena.upload(&l.wqkv, w)) || (rc = h->warena.upload(&l.bqkv, b))) return fail(rc);
    if ((rc = upload_t(h, wb, p + ".self_attn.out_proj.weight", 256 * 256, &l.wo))) return fail(rc);
    if ((rc = upload_t(h, wb, p + ".self_attn.out_proj.bias", 256, &l.bo))) return fail(rc);
    if ((rc = upload_t(h, wb, p + ".self_attn.ffn.0.weight", 512 * 512, &l.w1))) return fail(rc);
    if ((rc = upload_t(h, wb, p + ".self_attn.ffn.0.bias", 512, &l.b1))) return fail(rc);
    if ((rc = upload_t(h, wb, p + ".self_attn.ffn.1.weight", 512, &l.lng))) return fail(rc);
    if ((rc = upload_t(h, wb, p + ".self_attn.ffn.1.bias", 512, &l.lnb))) return fail(rc);
    if ((rc = upload_t(h, wb, p + ".self_attn.ffn.3.weight", 256 * 512, &l.w2))) return fail(rc);
    if ((rc = upload_t(h, wb, p + ".self_attn.ffn.3.bias", 256, &l.b2))) return fail(rc);
    // cross: [to_qk ; to_v] stacked -> one [512,256] projection
    const TensorView* wqk = wb.get(p + ".cross_attn.to_qk.weight", 256 * 256);
    const TensorView* bqk = wb.get(p + ".cross_attn.to_qk.bias", 256);
    const TensorView* wv = wb.get(p + ".cross_attn.to_v.weight", 256 * 256);
    const TensorView* bv = wb.get(p + ".cross_attn.to_v.bias", 256);
    if (!wqk || !bqk || !wv || !bv) return fail(B2S_EINVAL);
    std::vector<float> cw(512 * 256), cb(512);
    std::memcpy(cw.data(), wqk->data, 256 * 256 * sizeof(float));
    std::memcpy(cw.data() + 256 * 256, wv->data, 256 * 256 * sizeof(float));
    std::memcpy(cb.data(), bqk->data, 256 * sizeof(float));
    std::memcpy(cb.data() + 256, bv->data, 256 * sizeof(float));
    if ((rc = h->warena.upload(&l.cwqkv, cw)) || (rc = h->warena.upload(&l.cbqkv, cb))) return fail(rc);
    if ((rc = upload_t(h, wb, p + ".cross_attn.to_out.weight", 256 * 256, &l.cwo))) return fail(rc);
    if ((rc = upload_t(h, wb, p + ".cross_attn.to_out.bias", 256, &l.cbo))) return fail(rc);
    if ((rc = upload_t(h, wb, p + ".cross_attn.ffn.0.weight", 512 * 512, &l.cw1))) return fail(rc);
    if ((rc = upload_t(h, wb, p + ".cross_attn.ffn.0.bias", 512, &l.cb1))) return fail(rc);
    if ((rc = upload_t(h, wb, p + ".cross_attn.ffn.1.weight", 512, &l.clng))) return fail(rc);
    if ((rc = upload_t(h, wb, p + ".cross_attn.ffn.1.bias", 512, &l.clnb))) return fail(rc);
    if ((rc = upload_t(h, wb, p + ".cross_attn.ffn.3.weight", 256 * 512, &l.cw2))) return fail(rc);
    if ((rc = upload_t(h, wb, p + ".cross_attn.ffn.3.bias", 256, &l.cb2))) return fail(rc);
    const std::string a = "log_assignment." + std::to_string(i);
    if ((rc = upload_t(h, wb, a + ".matchability.weight", 256, &l.wmatch))) return fail(rc);
    if ((rc = upload_t(h, wb, a + ".final_proj.weight", 256 * 256, &l.wfinal))) return fail(rc);
    if ((rc = upload_t(h, wb, a + ".final_proj.bias", 256, &l.bfinal))) return fail(rc);
    const TensorView* mb = wb.get(a + ".matchability.bias", 1);
    if (!mb) return fail(B2S_EINVAL);
    l.bmatch = mb->data[0];
    l.wtok = nullptr; l.btok = 0.f;
    if (i < cfg->n_layers - 1) {
      const std::string t = "token_confidence." + std::to_string(i) + ".token.0";
      if ((rc = upload_t(h, wb, t + ".weight", 256, &l.wtok))) return fail(rc);
      const TensorView* tb = wb.get(t + ".bias", 1);
      if (!tb) return fail(B2S_EINVAL);
      l.btok = tb->data[0];
    }
    // upstream confidence_threshold(): clip(0.8 + 0.1*exp(-4 i / n_layers), 0, 1) (float64 -> float32)
    double th = 0.8 + 0.1 * std::exp(-4.0 * i / cfg->n_layers);
    h->thr.push_back((float)std::min(1.0, std::max(0.0, th)));
  }
  {
    std::vector<const float*> wf, bf, wm; std::vector<float> bm;
    for (const LgLayer& l : h->L) { wf.push_back(l.wfinal); bf.push_back(l.bfinal); wm.push_back(l.wmatch); bm.push_back(l.bmatch); }
    if ((rc = h->warena.upload(&h->wfinal_tab, wf)) || (rc = h->warena.upload(&h->bfinal_tab, bf)) ||
        (rc = h->warena.upload(&h->wmatch_tab, wm)) || (rc = h->warena.upload(&h->bmatch_tab, bm))) return fail(rc);
  }
  if ((rc = h->warena.alloc(&h->stats, (size_t)4))) return fail(rc);
  cudaMemset(h->stats, 0, 4 * sizeof(unsigned long long));
  if (cudaMallocHost((void**)&h->h_ctrl, LGC_INTS * sizeof(int)) != cudaSuccess) { set_error("cudaMallocHost failed"); return fail(B2S_ENOMEM); }
  cudaFuncSetAttribute(k_attn_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
  if (cfg->precision != B2S_FP32_SIMT) {   // tensor-core paths: bf16 operands, or fp32 carried as three bf16 planes
    if ((rc = lgtc_create(&h->tc, h->L.size(), cfg->precision == B2S_BF16 ? 1 : 3))) return fail(rc);
    lgtc_set_prof(h->tc, &h->prof, h->stats);
    for (size_t i = 0; i < h->L.size(); ++i) {
      const LgLayer& l = h->L[i];
      LgTcLayerSrc s = {l.wqkv, l.bqkv, l.wo, l.bo, l.w1, l.b1, l.lng, l.lnb, l.w2, l.b2,
                        l.cwqkv, l.cbqkv, l.cwo, l.cbo, l.cw1, l.cb1, l.clng, l.clnb, l.cw2, l.cb2};
      if ((rc = lgtc_set_layer(h->tc, (int)i, s))) return fail(rc);
    }
    std::vector<const float*> wf, bf;
    for (const LgLayer& l : h->L) { wf.push_back(l.wfinal); bf.push_back(l.bfinal); }
    if ((rc = lgtc_set_final(h->tc, wf, bf))) return fail(rc);
  }
  if ((rc = lg_alloc_ws(h, cfg->max_kp > 0 ? cfg->max_kp : 2048))) return fail(rc);
  *out = h;
  return 0;
}

extern "C" void b2s_lg_destroy(b2s_lg* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->h_ctrl) cudaFreeHost(h->h_ctrl);
  if (h->tc) lgtc_destroy(h->tc);
  delete h;
}

extern "C" long long b2s_lg_launch_count(const b2s_lg* h) { return h ? h->launches : 0; }

extern "C" int b2s_lg_profile(b2s_lg* h, int on) {
  if (!h) return B2S_EINVAL;
  h->prof.on = on != 0;
  for (int c = 0; c < PROF_NCLASS; ++c) h->prof.used[c] = 0;
  B2S_CUDA(cudaSetDevice(h->device));
  B2S_CUDA(cudaMemset(h->stats, 0, 4 * sizeof(unsigned long long)));
  return 0;
}

extern "C" int b2s_lg_profile_work(b2s_lg* h, double* self_pairs, double* cross_pairs) {
  if (!h || !self_pairs || !cross_pairs) return B2S_EINVAL;
  B2S_CUDA(cudaSetDevice(h->device));
  unsigned long long v[2] = {0, 0};
  B2S_CUDA(cudaDeviceSynchronize());
  B2S_CUDA(cudaMemcpy(v, h->stats, sizeof(v), cudaMemcpyDeviceToHost));   // synchronises with the default stream
  *self_pairs = (double)v[0]; *cross_pairs = (double)v[1];
  return 0;
}

extern "C" int b2s_lg_profile_read(b2s_lg* h, int cls, double* ms, long long* n) {
  if (!h || !ms || !n || cls < 0 || cls >= PROF_NCLASS) return B2S_EINVAL;
  B2S_CUDA(cudaSetDevice(h->device));
  return h->prof.read(cls, ms, n);
}

extern "C" int b2s_lg_set_debug(b2s_lg* h, int on) {
  if (!h) return B2S_EINVAL;
  B2S_CUDA(cudaSetDevice(h->device));
  h->debug = on;
  return lg_alloc_ws(h, h->cap);
}

// one Linear over both images' live rows (m, n = upper bounds; live counts come from h->ctrl)
static int lg_linear(b2s_lg* h, cudaStream_t st, const float* A1, int lda1, int K1, const float* A2, int lda2,
                     const float* W, int K, int N, const float* bias, float* C, int ldc, int m, int n,
                     const float* residual, int ldr, float alpha = 1.f) {
  GemmParams g;
  g.A1 = A1; g.lda1 = lda1; g.K1 = K1; g.A2 = A2; g.lda2 = lda2;
  g.W = W; g.ldw = K; g.C = C; g.ldc = ldc; g.N = N; g.K = K;
  g.nseg = 2; g.seg_base[0] = 0; g.seg_base[1] = h->cap; g.seg_rows[0] = m; g.seg_rows[1] = n;
  g.bias = bias; g.alpha = alpha; g.residual = residual; g.ldr = ldr;
  g.lg_ctrl = h->ctrl; g.lg_mode = 1;
  return gemm_simt(g, st, &h->launches, &h->prof);
}

static int lg_attention(b2s_lg* h, cudaStream_t st, AttnParams ap, int cross, int maxq) {
  if (maxq <= 0) return 0;
  ap.ctrl = h->ctrl; ap.cross = cross; ap.stats = h->prof.on ? h->stats : nullptr;
  dim3 grid(cdiv(maxq, ATT_B), 4, 2);
  h->prof.mark(PROF_ATTN, st);
  launch_k(k_attn_fp32, grid, 256, ATT_SMEM, st, ap);
  h->prof.mark(PROF_ATTN, st);
  ++h->launches;
  B2S_LAUNCH_CHECK();
  return 0;
}

static int lg_ffn(b2s_lg* h, cudaStream_t st, float* x, const float* msg, const float* w1, const float* b1,
                  const float* lng, const float* lnb, const float* w2, const float* b2, int m, int n) {
  // h1 = [x | msg] W1^T + b1 ; LN ; GELU ; x += h1 W2^T + b2
  B2S_TRY(lg_linear(h, st, x, 256, 256, msg, 256, w1, 512, 512, b1, h->h1, 512, m, n, nullptr, 0));
  RowSeg seg = {{0, h->cap}, {m, n}};
  dim3 grid(cdiv(std::max(m, n), 8), 2);
  launch_k(k_ln_gelu_512, grid, 256, 0, st, h->h1, seg, lng, lnb, (const int*)h->ctrl);
  ++h->launches;
  B2S_LAUNCH_CHECK();
  return lg_linear(h, st, h->h1, 512, 512, nullptr, 0, w2, 512, 256, b2, x, 256, m, n, x, 256);
}

static int lg_layer_fp32(b2s_lg* h, cudaStream_t st, int li, int cur, int m, int n) {
  const LgLayer& l = h->L[li];
  float* x = h->x[cur];
  const int cap = h->cap;
  // ---- self block (shared weights, both images) ----
  {
    GemmParams g;
    g.A1 = x; g.lda1 = 256; g.K1 = 256; g.W = l.wqkv; g.ldw = 256; g.C = h->qkv; g.ldc = 768; g.N = 768; g.K = 256;
    g.nseg = 2; g.seg_base[0] = 0; g.seg_base[1] = cap; g.seg_rows[0] = m; g.seg_rows[1] = n;
    g.bias = l.bqkv; g.rot_cols = 512; g.rot_cos = h->cosb[cur]; g.rot_sin = h->sinb[cur];
    g.lg_ctrl = h->ctrl; g.lg_mode = 1;
    B2S_TRY(gemm_simt(g, st, &h->launches, &h->prof));
  }
  AttnParams ap;
  ap.ldq = ap.ldk = ap.ldv = 768; ap.ldo = 256; ap.scale = 0.125f;
  ap.prob[0] = {h->qkv, h->qkv + 256, h->qkv + 512, h->ctx, m, m};
  ap.prob[1] = {h->qkv + (size_t)cap * 768, h->qkv + (size_t)cap * 768 + 256, h->qkv + (size_t)cap * 768 + 512,
                h->ctx + (size_t)cap * 256, n, n};
  B2S_TRY(lg_attention(h, st, ap, 0, std::max(m, n)));
  B2S_TRY(lg_linear(h, st, h->ctx, 256, 256, nullptr, 0, l.wo, 256, 256, l.bo, h->msg, 256, m, n, nullptr, 0));
  B2S_TRY(lg_ffn(h, st, x, h->msg, l.w1, l.b1, l.lng, l.lnb, l.w2, l.b2, m, n));
  // ---- cross block ----
  B2S_TRY(lg_linear(h, st, x, 256, 256, nullptr, 0, l.cwqkv, 256, 512, l.cbqkv, h->qkv, 512, m, n, nullptr, 0));
  ap.ldq = ap.ldk = ap.ldv = 512;
  float* q0 = h->qkv; float* q1 = h->qkv + (size_t)cap * 512;
  ap.prob[0] = {q0, q1, q1 + 256, h->ctx, m, n};
  ap.prob[1] = {q1, q0, q0 + 256, h->ctx + (size_t)cap * 256, n, m};
  B2S_TRY(lg_attention(h, st, ap, 1, std::max(m, n)));
  B2S_TRY(lg_linear(h, st, h->ctx, 256, 256, nullptr, 0, l.cwo, 256, 256, l.cbo, h->msg, 256, m, n, nullptr, 0));
  B2S_TRY(lg_ffn(h, st, x, h->msg, l.cw1, l.cb1, l.clng, l.clnb, l.cw2, l.cb2, m, n));
  return 0;
}

static int fill_i32(b2s_lg* h, cudaStream_t st, int32_t* p, int n, int32_t v) {
  if (!p || n <= 0) return 0;
  launch_k(k_fill_i32, cdiv(n, 256), 256, 0, st, p, n, v);
  ++h->launches;
  B2S_LAUNCH_CHECK();
  return 0;
}
static int fill_f32(b2s_lg* h, cudaStream_t st, float* p, int n, float v) {
  if (!p || n <= 0) return 0;
  launch_k(k_fill_f32, cdiv(n, 256), 256, 0, st, p, n, v);
  ++h->launches;
  B2S_LAUNCH_CHECK();
  return 0;
}

// The whole match is enqueued without a single host synchronisation: the adaptive-depth exit test
// and the adaptive-width pruning (upstream decides both on the host after every layer) are taken
// on the device; kernels of layers behind an exit find the stop flag set and return at once.
extern "C" int b2s_lightglue_match(b2s_lg* h, const float* k0, const float* d0, int m, const float* k1,
                                   const float* d1, int n, const float* size0, const float* size1, void* stream,
                                   int32_t* matches, float* mscores, int32_t* n_matches, int32_t* stop_layer,
                                   int32_t* matches0, int32_t* matches1, float* ms0, float* ms1,
                                   int32_t* prune0, int32_t* prune1) {
  if (!h || !matches || !mscores || !n_matches || m < 0 || n < 0) { set_error("b2s_lightglue_match: bad argument"); return B2S_EINVAL; }
  B2S_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int L = h->cfg.n_layers;
  const bool do_stop = h->cfg.depth_conf > 0.f;
  const bool do_prune = h->cfg.width_conf > 0.f;
  // full-size outputs default to "unmatched"
  B2S_TRY(fill_i32(h, st, matches0, m, -1));
  B2S_TRY(fill_i32(h, st, matches1, n, -1));
  B2S_TRY(fill_f32(h, st, ms0, m, 0.f));
  B2S_TRY(fill_f32(h, st, ms1, n, 0.f));
  if (!do_prune) {  // upstream: prune = n_layers everywhere when pruning is off
    B2S_TRY(fill_i32(h, st, prune0, m, L));
    B2S_TRY(fill_i32(h, st, prune1, n, L));
  }
  if (m == 0 || n == 0) {
    B2S_CUDA(cudaMemsetAsync(n_matches, 0, sizeof(int32_t), st));
    if (do_prune) { B2S_TRY(fill_i32(h, st, prune0, m, 1)); B2S_TRY(fill_i32(h, st, prune1, n, 1)); }
    B2S_TRY(fill_i32(h, st, stop_layer, 1, 1));
    return 0;
  }
  if (std::max(m, n) > h->cap) {
    B2S_CUDA(cudaStreamSynchronize(st));
    B2S_TRY(lg_alloc_ws(h, std::max(m, n)));
  }
  const int cap = h->cap;
  int* pr0 = do_prune ? (prune0 ? prune0 : h->prune_scratch[0]) : nullptr;
  int* pr1 = do_prune ? (prune1 ? prune1 : h->prune_scratch[1]) : nullptr;
  {
    PosencParams pp;
    pp.kp[0] = k0; pp.kp[1] = k1; pp.n[0] = m; pp.n[1] = n; pp.base[0] = 0; pp.base[1] = cap;
    pp.has_size[0] = size0 != nullptr; pp.has_size[1] = size1 != nullptr;
    pp.size[0][0] = size0 ? size0[0] : 0.f; pp.size[0][1] = size0 ? size0[1] : 0.f;
    pp.size[1][0] = size1 ? size1[0] : 0.f; pp.size[1][1] = size1 ? size1[1] : 0.f;
    pp.Wr = h->wr; pp.kn = h->kn; pp.cosb = h->cosb[0]; pp.sinb = h->sinb[0]; pp.ind = h->ind[0];
    pp.prune[0] = pr0; pp.prune[1] = pr1;
    pp.ctrl = h->ctrl; pp.last_init = (do_stop || do_prune) ? 0 : L - 1;
    launch_k(k_lg_posenc, dim3(cdiv(std::max(m, n), 64), 2), 256, 0, st, pp);
    ++h->launches;
    B2S_LAUNCH_CHECK();
  }
  // input projection (two sources -> rows 0.. and cap..)
  for (int s = 0; s < 2; ++s) {
    GemmParams g;
    g.A1 = s ? d1 : d0; g.lda1 = 128; g.K1 = 128; g.W = h->in_w; g.ldw = 128; g.K = 128; g.N = 256;
    g.C = h->x[0] + (size_t)(s ? cap : 0) * 256; g.ldc = 256; g.M = s ? n : m; g.bias = h->in_b;
    B2S_TRY(gemm_simt(g, st, &h->launches));
  }
  h->dbg_nlayers = 0;
  // with pruning enabled layer i lives in ping-pong buffer i & 1 (the gather after layer i moves the
  // survivors across); without it everything stays in buffer 0
  auto buf_of = [&](int i) { return do_prune ? (i & 1) : 0; };
  for (int i = 0; i < L; ++i) {
    const int cur = buf_of(i);
    if (h->tc) B2S_TRY(lgtc_layer(h->tc, st, i, h->x[cur], h->cosb[cur], h->sinb[cur], cap, m, n, h->ctrl, i == 0, &h->launches));
    else B2S_TRY(lg_layer_fp32(h, st, i, cur, m, n));
    if (h->debug && h->dbg_layers) {   // debug only: snapshot the layer output and its live sizes (host sync)
      B2S_CUDA(cudaMemcpyAsync(h->dbg_layers + (size_t)i * 2 * cap * 256, h->x[cur], (size_t)2 * cap * 256 * sizeof(float),
                               cudaMemcpyDeviceToDevice, st));
      B2S_CUDA(cudaMemcpyAsync(h->h_ctrl, h->ctrl, LGC_INTS * sizeof(int), cudaMemcpyDeviceToHost, st));
      B2S_CUDA(cudaStreamSynchronize(st));
      if (!h->h_ctrl[LGC_STOP] && h->h_ctrl[LGC_M] > 0 && h->h_ctrl[LGC_N] > 0) {
        h->dbg_m[i] = h->h_ctrl[LGC_M]; h->dbg_n[i] = h->h_ctrl[LGC_N]; h->dbg_nlayers = i + 1;
      }
    }
    if (i == L - 1 || (!do_stop && !do_prune)) continue;
    const LgLayer& l = h->L[i];
    // upstream: scores > (1 - width_confidence) with a python double; undo the float rounding of the cfg
    const float keep_thr = (float)(1.0 - std::round((double)h->cfg.width_conf * 1e6) / 1e6);
    if (do_prune && cap <= 32 * LG_MAXBLK) {
      // two launches: heads per 32-row block (keep masks, one atomic per CTA), then decision + placement + copy
      const int nblk = cdiv(std::max(m, n), 32);
      HeadBlkParams hp = {};
      hp.x = h->x[cur]; hp.base[0] = 0; hp.base[1] = cap;
      hp.wt = l.wtok; hp.bt = l.btok; hp.wm = l.wmatch; hp.bm = l.bmatch;
      hp.thr = h->thr[i]; hp.keep_thr = keep_thr; hp.use_tok = do_stop;
      hp.tok = h->tok; hp.ctrl = h->ctrl; hp.adapt = h->adapt; hp.layer = i;
      launch_k(k_lg_heads_blk, dim3(nblk, 2), 1024, 0, st, hp);
      const int nxt = cur ^ 1;
      GatherBlkParams gp = {};
      gp.adapt = h->adapt; gp.ctrl = h->ctrl; gp.base[0] = 0; gp.base[1] = cap;
      gp.layer = i; gp.num_points = m + n; gp.do_stop = do_stop ? 1 : 0; gp.pruning_min_kpts = h->cfg.pruning_min_kpts;
      gp.nblk = nblk; gp.depth_conf = h->cfg.depth_conf;
      gp.x_in = h->x[cur]; gp.x_out = h->x[nxt]; gp.cos_in = h->cosb[cur]; gp.cos_out = h->cosb[nxt];
      gp.sin_in = h->sinb[cur]; gp.sin_out = h->sinb[nxt]; gp.ind_in = h->ind[cur]; gp.ind_out = h->ind[nxt];
      gp.prune[0] = pr0; gp.prune[1] = pr1;
      gp.xb_out = h->tc ? lgtc_xb(h->tc) : nullptr;
      gp.xb_planes = h->tc ? lgtc_planes(h->tc) : 0; gp.xb_plane = (size_t)2 * cap * 256;
      launch_k(k_lg_gather_blk, dim3(nblk, 2), 1024, 0, st, gp);
      h->launches += 2;
      B2S_LAUNCH_CHECK();
      continue;
    }
    HeadParams hp = {};
    hp.x = h->x[cur]; hp.seg = {{0, cap}, {m, n}};
    hp.wt = l.wtok; hp.bt = l.btok; hp.wm = l.wmatch; hp.bm = l.bmatch;
    hp.thr = h->thr[i];
    hp.keep_thr = keep_thr;
    hp.use_tok = do_stop; hp.use_match = do_prune;
    hp.tok = h->tok; hp.keep = h->keep; hp.ctrl = h->ctrl; hp.layer = i; hp.ls_pos = nullptr;
    dim3 hg(cdiv(std::max(m, n), 8), 2);
    launch_k(k_lg_heads, hg, 256, 0, st, hp);
    ScanParams sp;
    sp.keep = h->keep; sp.srcmap = h->srcmap; sp.ctrl = h->ctrl; sp.base[0] = 0; sp.base[1] = cap;
    sp.layer = i; sp.num_points = m + n; sp.do_stop = do_stop ? 1 : 0; sp.do_prune = do_prune ? 1 : 0;
    sp.pruning_min_kpts = h->cfg.pruning_min_kpts; sp.depth_conf = h->cfg.depth_conf;
    launch_k(k_lg_prune_scan, 2, 1024, 0, st, sp);
    h->launches += 2;
    B2S_LAUNCH_CHECK();
    if (do_prune) {
      const int nxt = cur ^ 1;
      GatherParams gp;
      gp.srcmap = h->srcmap; gp.ctrl = h->ctrl; gp.base[0] = 0; gp.base[1] = cap;
      gp.x_in = h->x[cur]; gp.x_out = h->x[nxt]; gp.cos_in = h->cosb[cur]; gp.cos_out = h->cosb[nxt];
      gp.sin_in = h->sinb[cur]; gp.sin_out = h->sinb[nxt]; gp.ind_in = h->ind[cur]; gp.ind_out = h->ind[nxt];
      gp.prune[0] = pr0; gp.prune[1] = pr1;
      gp.xb_out = h->tc ? lgtc_xb(h->tc) : nullptr;
      gp.xb_planes = h->tc ? lgtc_planes(h->tc) : 0; gp.xb_plane = (size_t)2 * cap * 256;
      launch_k(k_lg_gather, dim3(cdiv(std::max(m, n), 8), 2), 256, 0, st, gp);
      ++h->launches;
      B2S_LAUNCH_CHECK();
    }
  }
  // ---- assignment with the last executed layer's heads (K14/K15); that layer index, its buffer and the
  //      live sizes are device-side values ----
  float* md = h->qkv;  // [2*cap, 256] (CUDA-core path)
  if (!h->tc) {
    GemmParams g;
    g.A1 = h->x[0]; g.lda1 = 256; g.K1 = 256; g.A1_alt = do_prune ? h->x[1] : nullptr;
    g.W = h->L[L - 1].wfinal; g.ldw = 256; g.K = 256; g.N = 256; g.bias = h->L[L - 1].bfinal; g.alpha = 0.25f;
    g.C = md; g.ldc = 256;
    g.nseg = 2; g.seg_base[0] = 0; g.seg_base[1] = cap; g.seg_rows[0] = m; g.seg_rows[1] = n;
    g.lg_ctrl = h->ctrl; g.lg_mode = 2; g.w_tab = h->wfinal_tab; g.b_tab = h->bfinal_tab;
    B2S_TRY(gemm_simt(g, st, &h->launches, &h->prof));
  }
  {
    HeadParams hp = {};
    hp.x = h->x[0]; hp.x_alt = do_prune ? h->x[1] : nullptr; hp.seg = {{0, cap}, {m, n}};
    hp.wm_tab = h->wmatch_tab; hp.bm_tab = h->bmatch_tab; hp.ctrl = h->ctrl; hp.ls_pos = h->ls;
    dim3 hg(cdiv(std::max(m, n), 8), 2);
    launch_k(k_lg_heads, hg, 256, 0, st, hp);
    ++h->launches;
    B2S_LAUNCH_CHECK();
  }
  const int ld = cap;
  if (h->tc) {
    B2S_TRY(lgtc_assignment(h->tc, st, h->x[0], do_prune ? h->x[1] : nullptr, cap, m, n, h->ctrl, h->sim, h->simT, &h->launches));
  } else {
    GemmParams g;
    g.A1 = md; g.lda1 = 256; g.K1 = 256; g.W = md + (size_t)cap * 256; g.ldw = 256; g.K = 256;
    g.M = m; g.N = n; g.C = h->sim; g.ldc = ld;
    g.lg_ctrl = h->ctrl; g.lg_mode = 3;
    B2S_TRY(gemm_simt(g, st, &h->launches));
  }
  const int* ctrl = h->ctrl;
  if (h->tc) {
    // both directions per launch on sim / sim^T (written by the similarity GEMM's epilogue)
    Lse2Params lp = {h->sim, h->simT, ld, ctrl, h->rmax, h->rlog, h->cmax, h->clog};
    launch_k(k_lg_lse2, dim3(cdiv(std::max(m, n), 8), 2), 256, 0, st, lp);
    Argmax2Params ap = {h->sim, h->simT, ld, ctrl, h->rmax, h->rlog, h->cmax, h->clog, h->ls, h->ls + cap, h->max0, h->m0, h->m1};
    launch_k(k_lg_argmax2, dim3(cdiv(std::max(m, n), 8), 2), 256, 0, st, ap);
    h->launches -= 2;
  } else {
    launch_k(k_lg_row_lse, cdiv(m, 8), 256, 0, st, h->sim, ld, ctrl, h->rmax, h->rlog);
    launch_k(k_lg_col_lse, cdiv(n, 32), 1024, 0, st, h->sim, ld, ctrl, h->cmax, h->clog);
    launch_k(k_lg_row_argmax, cdiv(m, 8), 256, 0, st, h->sim, ld, ctrl, h->rmax, h->rlog, h->cmax, h->clog, h->ls, h->ls + cap, h->max0, h->m0);
    launch_k(k_lg_col_argmax, cdiv(n, 32), 1024, 0, st, h->sim, ld, ctrl, h->rmax, h->rlog, h->cmax, h->clog, h->ls, h->ls + cap, h->m1);
  }
  FilterParams fp = {};
  fp.th = h->cfg.filter_thresh; fp.max0 = h->max0; fp.m0 = h->m0; fp.m1 = h->m1;
  fp.ctrl = h->ctrl; fp.cap = cap; fp.stop_layer = stop_layer;
  fp.ind0 = h->ind[0]; fp.ind1 = h->ind[0] + cap; fp.ind_alt = do_prune ? h->ind[1] : nullptr;
  fp.matches = matches; fp.mscores = mscores; fp.n_matches = n_matches;
  fp.matches0 = matches0; fp.matches1 = matches1; fp.ms0 = ms0; fp.ms1 = ms1;
  launch_k(k_lg_filter, 1, 1024, 0, st, fp);
  h->launches += 5;
  B2S_LAUNCH_CHECK();
  if (h->debug) {
    B2S_CUDA(cudaMemcpyAsync(h->h_ctrl, h->ctrl, LGC_INTS * sizeof(int), cudaMemcpyDeviceToHost, st));
    B2S_CUDA(cudaStreamSynchronize(st));
    h->dbg_simm = h->h_ctrl[LGC_M]; h->dbg_simn = h->h_ctrl[LGC_N];
  }
  return 0;
}

static int lg_host_staging(b2s_lg* h, int need) {
  if (need <= h->hcap) return 0;
  need = (need + 255) / 256 * 256;
  h->hostarena.release();
  h->hcap = 0;
  for (int s = 0; s < 2; ++s) {
    B2S_TRY(h->hostarena.alloc(&h->hk[s], (size_t)need * 2));
    B2S_TRY(h->hostarena.alloc(&h->hd[s], (size_t)need * 128));
    B2S_TRY(h->hostarena.alloc(&h->hms[s], (size_t)need));
    B2S_TRY(h->hostarena.alloc(&h->hm[s], (size_t)need));
    B2S_TRY(h->hostarena.alloc(&h->hprune[s], (size_t)need));
  }
  B2S_TRY(h->hostarena.alloc(&h->hmatches, (size_t)need * 2));
  B2S_TRY(h->hostarena.alloc(&h->hmscores, (size_t)need));
  B2S_TRY(h->hostarena.alloc(&h->hnm, (size_t)4));
  h->hcap = need;
  return 0;
}

extern "C" int b2s_lightglue_match_host(b2s_lg* h, const float* k0, const float* d0, int m, const float* k1,
                                        const float* d1, int n, const float* size0, const float* size1,
                                        int32_t* matches, float* mscores, int32_t* n_matches, int32_t* stop_layer,
                                        int32_t* matches0, int32_t* matches1, float* ms0, float* ms1,
                                        int32_t* prune0, int32_t* prune1) {
  if (!h || !n_matches || m < 0 || n < 0) { set_error("b2s_lightglue_match_host: bad argument"); return B2S_EINVAL; }
  B2S_CUDA(cudaSetDevice(h->device));
  B2S_TRY(lg_host_staging(h, std::max(std::max(m, n), 1)));
  cudaStream_t st = 0;
  if (m > 0) {
    B2S_CUDA(cudaMemcpyAsync(h->hk[0], k0, (size_t)m * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
    B2S_CUDA(cudaMemcpyAsync(h->hd[0], d0, (size_t)m * 128 * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  if (n > 0) {
    B2S_CUDA(cudaMemcpyAsync(h->hk[1], k1, (size_t)n * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
    B2S_CUDA(cudaMemcpyAsync(h->hd[1], d1, (size_t)n * 128 * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  B2S_TRY(b2s_lightglue_match(h, h->hk[0], h->hd[0], m, h->hk[1], h->hd[1], n, size0, size1, st, h->hmatches,
                              h->hmscores, h->hnm, h->hnm + 1, h->hm[0], h->hm[1], h->hms[0], h->hms[1],
                              h->hprune[0], h->hprune[1]));
  B2S_CUDA(cudaMemcpyAsync(h->h_ctrl, h->hnm, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));   // pinned: n_matches, stop
  B2S_CUDA(cudaStreamSynchronize(st));
  const int32_t nm = h->h_ctrl[0];
  *n_matches = nm;
  if (stop_layer) *stop_layer = h->h_ctrl[1];
  if (nm > 0) {
    if (matches) B2S_CUDA(cudaMemcpyAsync(matches, h->hmatches, (size_t)nm * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (mscores) B2S_CUDA(cudaMemcpyAsync(mscores, h->hmscores, (size_t)nm * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  if (matches0 && m) B2S_CUDA(cudaMemcpyAsync(matches0, h->hm[0], (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (matches1 && n) B2S_CUDA(cudaMemcpyAsync(matches1, h->hm[1], (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (ms0 && m) B2S_CUDA(cudaMemcpyAsync(ms0, h->hms[0], (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (ms1 && n) B2S_CUDA(cudaMemcpyAsync(ms1, h->hms[1], (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (prune0 && m) B2S_CUDA(cudaMemcpyAsync(prune0, h->hprune[0], (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (prune1 && n) B2S_CUDA(cudaMemcpyAsync(prune1, h->hprune[1], (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  B2S_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int b2s_lightglue_match_batch(b2s_lg* h, const float* kpts, const float* desc, const int32_t* cu,
                                         int n_frames, const int32_t* pair_i, const int32_t* pair_j, int n_pairs,
                                         void* stream, int stride, int32_t* matches, float* mscores,
                                         int32_t* n_matches) {
  if (!h || !cu || !pair_i || !pair_j || n_pairs < 0) { set_error("b2s_lightglue_match_batch: bad argument"); return B2S_EINVAL; }
  for (int p = 0; p < n_pairs; ++p) {
    const int a = pair_i[p], b = pair_j[p];
    if (a < 0 || b < 0 || a >= n_frames || b >= n_frames) { set_error("pair %d out of range", p); return B2S_EINVAL; }
    const int m = cu[a + 1] - cu[a], n = cu[b + 1] - cu[b];
    if (std::min(m, n) > stride) { set_error("stride %d too small for pair %d", stride, p); return B2S_ESIZE; }
    B2S_TRY(b2s_lightglue_match(h, kpts + (size_t)cu[a] * 2, desc + (size_t)cu[a] * 128, m, kpts + (size_t)cu[b] * 2,
                                desc + (size_t)cu[b] * 128, n, nullptr, nullptr, stream,
                                matches + (size_t)p * stride * 2, mscores + (size_t)p * stride, n_matches + p, nullptr,
                                nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
  }
  return 0;
}

extern "C" int b2s_lg_debug_get(b2s_lg* h, const char* name, float* out, size_t cap_out, size_t* nout) {
  if (!h || !name || !nout) return B2S_EINVAL;
  B2S_CUDA(cudaSetDevice(h->device));
  B2S_CUDA(cudaDeviceSynchronize());
  const std::string s(name);
  const float* src = nullptr; size_t cnt = 0;
  std::vector<float> tmp;
  if (s.rfind("layer", 0) == 0 && h->dbg_layers) {
    // "layer<i>_<side>" -> [rows,256]
    int li = 0, side = 0;
    if (sscanf(name, "layer%d_%d", &li, &side) != 2 || li < 0 || li >= h->dbg_nlayers || side < 0 || side > 1) {
      set_error("debug tensor %s not available", name); return B2S_EINVAL;
    }
    src = h->dbg_layers + ((size_t)li * 2 * h->cap + (side ? h->cap : 0)) * 256;
    cnt = (size_t)(side ? h->dbg_n[li] : h->dbg_m[li]) * 256;
  } else if (s == "sim") {
    // compacted [m,n]
    tmp.resize((size_t)h->dbg_simm * h->dbg_simn);
    if (!tmp.empty())
      B2S_CUDA(cudaMemcpy2D(tmp.data(), (size_t)h->dbg_simn * sizeof(float), h->sim, (size_t)h->cap * sizeof(float),
                            (size_t)h->dbg_simn * sizeof(float), h->dbg_simm, cudaMemcpyDeviceToHost));
    *nout = tmp.size();
    std::memcpy(out, tmp.data(), std::min(cap_out, tmp.size()) * sizeof(float));
    return 0;
  } else {
    set_error("unknown debug tensor %s", name);
    return B2S_EINVAL;
  }
  *nout = cnt;
  const size_t c = std::min(cnt, cap_out);
  if (c) B2S_CUDA(cudaMemcpy(out, src, c * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}
