// lightglue.cu - host orchestration of the LightGlue matcher behind b2s_lightglue_*.
// Replaces `matcher({...})` at /root/reference/slam/core/features_utils.py:157-161 (the
// arithmetic itself lives in the un-vendored `lightglue` package; spec: SURVEY.md A.3).
// One launch sequence serves a BATCH of up to LG_MAXP independent pairs (keyframe-window all-pairs at
// keyframe_utils.py:153 / triangulation_utils.py:131, or consecutive pairs of a stream): every kernel takes the pair
// as a grid dimension, sizes and early-exit / pruning decisions live in per-pair device state.
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
  // workspace: segments of `cap` rows, segment 2p + s = image s of pair p; pcap pairs
  int cap = 0, pcap = 0;
  float *cosb[2] = {nullptr, nullptr}, *sinb[2] = {nullptr, nullptr}, *x[2] = {nullptr, nullptr};
  float *sim = nullptr, *simT = nullptr;
  float *rmax = nullptr, *rlog = nullptr, *cmax = nullptr, *clog = nullptr, *ls = nullptr, *max0 = nullptr;
  int *ind[2] = {nullptr, nullptr}, *prune = nullptr, *adapt = nullptr, *m0 = nullptr, *m1 = nullptr, *ctrl = nullptr;
  const float** wmatch_tab = nullptr; float* bmatch_tab = nullptr;   // device tables [L]
  int* h_ctrl = nullptr;  // pinned
  // host-API staging
  DeviceArena hostarena;
  int hcap = 0;
  float *hk[2] = {nullptr, nullptr}, *hd[2] = {nullptr, nullptr}, *hms[2] = {nullptr, nullptr}, *hmscores = nullptr;
  int32_t *hmatches = nullptr, *hnm = nullptr, *hm[2] = {nullptr, nullptr}, *hprune[2] = {nullptr, nullptr};
  // debug (pair 0 of a batch)
  int debug = 0;
  float* dbg_layers = nullptr;  // [n_layers][2*cap][256]
  int dbg_m[16], dbg_n[16], dbg_nlayers = 0, dbg_simm = 0, dbg_simn = 0;
  long long launches = 0;
  KernelProf prof;
  unsigned long long* stats = nullptr;   // device: [0] sum nq*nk over self-attention problems, [1] over cross-attention problems
  LgTensorCore* tc = nullptr;   // tcgen05 layers: bf16 operands (B2S_BF16), fp32 as two fp16 planes (B2S_FP32) or as three bf16 planes (B2S_FP32X3)
  int planes = 3;               // operand planes of tc: 1, 2 or 3
  // B2S_FP32 only: the weight blob and, created on the first range overflow, a B2S_FP32X3 matcher that re-runs flagged batches
  std::vector<uint8_t> blob;
  b2s_lg* fallback = nullptr;
  long long range_fallbacks = 0;
};
static int planes_of(int precision) { return precision == B2S_BF16 ? 1 : precision == B2S_FP32 ? 2 : 3; }

static int upload_t(b2s_lg* h, const WeightBlob& wb, const std::string& name, size_t numel, float** out) {
  const TensorView* t = wb.get(name, numel);
  if (!t) return B2S_EINVAL;
  std::vector<float> v(t->data, t->data + numel);
  return h->warena.upload(out, v);
}

static size_t lg_ws_bytes(int np, int cap, int pcap, int n_layers, bool debug) {
  const size_t R = (size_t)2 * pcap * cap, PC = (size_t)pcap * cap;
  size_t b = 4 * (2 * (R * 32 * 2 + R * 256 + R) + R /*prune*/ + R /*ls*/ + 2 * PC * cap + 7 * PC + (size_t)pcap * (LG_ADAPT_INTS + LGC_INTS));
  if (debug) b += 4 * (size_t)n_layers * 2 * cap * 256;
  return b + lgtc_ws_bytes(np, cap, pcap);
}

static int lg_alloc_ws(b2s_lg* h, int cap, int pcap) {
  cap = (cap + 127) / 128 * 128;
  pcap = std::max(1, std::min(pcap, LG_MAXP));
  if (cap > 32 * LG_MAXBLK) { set_error("b2s_lightglue: %d keypoints per image exceed the supported %d", cap, 32 * LG_MAXBLK); return B2S_ESIZE; }
  h->wsarena.release();
  h->cap = h->pcap = 0;
  const size_t R = (size_t)2 * pcap * cap, PC = (size_t)pcap * cap;
  for (int i = 0; i < 2; ++i) {
    B2S_TRY(h->wsarena.alloc(&h->cosb[i], R * 32));
    B2S_TRY(h->wsarena.alloc(&h->sinb[i], R * 32));
    B2S_TRY(h->wsarena.alloc(&h->x[i], R * 256));
    B2S_TRY(h->wsarena.alloc(&h->ind[i], R));
  }
  B2S_TRY(h->wsarena.alloc(&h->prune, R));
  B2S_TRY(h->wsarena.alloc(&h->sim, PC * cap));
  B2S_TRY(h->wsarena.alloc(&h->simT, PC * cap));   // transposed similarity
  B2S_TRY(h->wsarena.alloc(&h->rmax, PC));
  B2S_TRY(h->wsarena.alloc(&h->rlog, PC));
  B2S_TRY(h->wsarena.alloc(&h->cmax, PC));
  B2S_TRY(h->wsarena.alloc(&h->clog, PC));
  B2S_TRY(h->wsarena.alloc(&h->ls, R));
  B2S_TRY(h->wsarena.alloc(&h->max0, PC));
  B2S_TRY(h->wsarena.alloc(&h->adapt, (size_t)pcap * LG_ADAPT_INTS));
  B2S_TRY(h->wsarena.alloc(&h->m0, PC));
  B2S_TRY(h->wsarena.alloc(&h->m1, PC));
  B2S_TRY(h->wsarena.alloc(&h->ctrl, (size_t)pcap * LGC_INTS));
  h->dbg_layers = nullptr;
  if (h->debug) B2S_TRY(h->wsarena.alloc(&h->dbg_layers, (size_t)h->cfg.n_layers * 2 * cap * 256));
  B2S_CUDA(cudaMemset(h->ctrl, 0, (size_t)pcap * LGC_INTS * sizeof(int)));
  B2S_TRY(lgtc_alloc_ws(h->tc, cap, pcap));
  h->cap = cap; h->pcap = pcap;
  return 0;
}

extern "C" void b2s_lg_default_cfg(b2s_lg_cfg* c) {
  c->n_layers = 9; c->heads = 4; c->dim = 256; c->in_dim = 128;
  c->depth_conf = 0.95f; c->width_conf = 0.99f; c->filter_thresh = 0.1f;
  c->pruning_min_kpts = -1; c->precision = B2S_FP32; c->max_kp = 2048;
}

extern "C" int b2s_lightglue_create(const b2s_lg_cfg* cfg, const void* weights, size_t nbytes, int device, b2s_lg** out) {
  if (!cfg || !weights || !out) { set_error("b2s_lightglue_create: null argument"); return B2S_EINVAL; }
  if (cfg->precision != B2S_FP32 && cfg->precision != B2S_BF16 && cfg->precision != B2S_FP32X3) {
    set_error("b2s_lightglue_create: unknown precision %d (B2S_FP32 / B2S_FP32X3 = fp32-faithful on tcgen05, B2S_BF16)", cfg->precision);
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
    std::vector<const float*> wm; std::vector<float> bm;
    for (const LgLayer& l : h->L) { wm.push_back(l.wmatch); bm.push_back(l.bmatch); }
    if ((rc = h->warena.upload(&h->wmatch_tab, wm)) || (rc = h->warena.upload(&h->bmatch_tab, bm))) return fail(rc);
  }
  if ((rc = h->warena.alloc(&h->stats, (size_t)4))) return fail(rc);
  cudaMemset(h->stats, 0, 4 * sizeof(unsigned long long));
  if (cudaMallocHost((void**)&h->h_ctrl, LGC_INTS * sizeof(int)) != cudaSuccess) { set_error("cudaMallocHost failed"); return fail(B2S_ENOMEM); }
  // tensor-core layers: bf16 operands, fp32 carried as two fp16 planes, or as three bf16 planes
  auto build_tc = [&](int planes) -> int {
    int r = 0;
    if ((r = lgtc_create(&h->tc, h->L.size(), planes))) return r;
    h->planes = planes;
    lgtc_set_prof(h->tc, &h->prof, h->stats);
    for (size_t i = 0; i < h->L.size(); ++i) {
      const LgLayer& l = h->L[i];
      LgTcLayerSrc s = {l.wqkv, l.bqkv, l.wo, l.bo, l.w1, l.b1, l.lng, l.lnb, l.w2, l.b2,
                        l.cwqkv, l.cbqkv, l.cwo, l.cbo, l.cw1, l.cb1, l.clng, l.clnb, l.cw2, l.cb2};
      if ((r = lgtc_set_layer(h->tc, (int)i, s))) return r;
    }
    std::vector<const float*> wf, bf;
    for (const LgLayer& l : h->L) { wf.push_back(l.wfinal); bf.push_back(l.bfinal); }
    if ((r = lgtc_set_final(h->tc, wf, bf)) || (r = lgtc_set_input(h->tc, h->in_w, h->in_b))) return r;
    return 0;
  };
  if ((rc = build_tc(planes_of(cfg->precision)))) return fail(rc);
  if (h->planes == 2) {
    // a weight outside the fp16 range (never seen with real checkpoints) rules the fp16x2 engine out for this model
    int bad = 0;
    if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpy(&bad, lgtc_range_flag(h->tc), sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) {
      set_error("b2s_lightglue_create: weight range check failed"); return fail(B2S_ECUDA);
    }
    if (bad) {
      lgtc_destroy(h->tc); h->tc = nullptr;
      if ((rc = build_tc(3))) return fail(rc);
    } else {
      h->blob.assign((const uint8_t*)weights, (const uint8_t*)weights + nbytes);
    }
  }
  if ((rc = lg_alloc_ws(h, cfg->max_kp > 0 ? cfg->max_kp : 2048, 1))) return fail(rc);
  *out = h;
  return 0;
}

extern "C" void b2s_lg_destroy(b2s_lg* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->h_ctrl) cudaFreeHost(h->h_ctrl);
  if (h->fallback) b2s_lg_destroy(h->fallback);
  if (h->tc) lgtc_destroy(h->tc);
  delete h;
}

// B2S_FP32 matcher: the bf16x3 matcher that re-runs launch sequences whose values left the fp16 range (created on first
// use from the retained weight blob); *out = nullptr when this matcher needs none (B2S_BF16 / B2S_FP32X3).
extern "C" int b2s_lg_fallback(b2s_lg* h, b2s_lg** out) {
  if (!h || !out) { set_error("b2s_lg_fallback: null argument"); return B2S_EINVAL; }
  *out = nullptr;
  if (h->planes != 2) return 0;
  if (!h->fallback) {
    b2s_lg_cfg c = h->cfg;
    c.precision = B2S_FP32X3; c.max_kp = std::max(h->cap, 128);
    B2S_TRY(b2s_lightglue_create(&c, h->blob.data(), h->blob.size(), h->device, &h->fallback));
  }
  ++h->range_fallbacks;
  *out = h->fallback;
  return 0;
}
extern "C" long long b2s_lg_range_fallbacks(const b2s_lg* h) { return h ? h->range_fallbacks : 0; }
extern "C" int b2s_lg_planes(const b2s_lg* h) { return h ? h->planes : 0; }

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
  B2S_CUDA(cudaDeviceSynchronize());
  h->debug = on;
  return lg_alloc_ws(h, h->cap, h->pcap);
}

extern "C" size_t b2s_lg_workspace_bytes(const b2s_lg_cfg* cfg, int max_kp, int pairs) {
  if (!cfg || max_kp < 1 || pairs < 1) return 0;
  return lg_ws_bytes(planes_of(cfg->precision), (max_kp + 127) / 128 * 128, std::min(pairs, LG_MAXP), cfg->n_layers, false);
}

extern "C" int b2s_lg_max_batch(void) { return LG_MAXP; }

extern "C" int b2s_lg_reserve(b2s_lg* h, int max_kp, int pairs) {
  if (!h || max_kp < 1 || pairs < 1) { set_error("b2s_lg_reserve: bad argument"); return B2S_EINVAL; }
  B2S_CUDA(cudaSetDevice(h->device));
  const int cap = (max_kp + 127) / 128 * 128, pc = std::min(pairs, LG_MAXP);
  if (cap <= h->cap && pc <= h->pcap) return 0;
  B2S_CUDA(cudaDeviceSynchronize());
  return lg_alloc_ws(h, std::max(cap, h->cap), std::max(pc, h->pcap));
}

// ---------------------------------------------------------------------------------------------
// One launch sequence for np <= LG_MAXP pairs.  The whole batch is enqueued without a single host synchronisation: the
// adaptive-depth exit test and the adaptive-width pruning (upstream decides both on the host after every layer) are
// taken on the device per pair; kernels find a stopped pair's flag set and skip its tiles at once.
// ---------------------------------------------------------------------------------------------
static int lg_run(b2s_lg* h, cudaStream_t st, const LgBatchIn& in, const LgBatchOut& out, int np) {
  const int L = h->cfg.n_layers;
  const bool do_stop = h->cfg.depth_conf > 0.f;
  const bool do_prune = h->cfg.width_conf > 0.f;
  int maxm = 0, maxn = 0;
  for (int p = 0; p < np; ++p) { maxm = std::max(maxm, in.pr[p].n[0]); maxn = std::max(maxn, in.pr[p].n[1]); }
  const int maxrows = std::max(maxm, maxn);
  if (maxrows > h->cap || np > h->pcap) {
    B2S_CUDA(cudaStreamSynchronize(st));
    B2S_TRY(lg_alloc_ws(h, std::max(maxrows, h->cap), std::max(np, h->pcap)));
  }
  const int cap = h->cap, nseg = 2 * np;
  const size_t plane_rows = (size_t)2 * h->pcap * cap;
  {
    PosencParams pp = {};
    pp.in = in; pp.cap = cap; pp.Wr = h->wr; pp.cosb = h->cosb[0]; pp.sinb = h->sinb[0]; pp.ind = h->ind[0]; pp.prune = h->prune;
    pp.din = lgtc_din(h->tc); pp.din_plane = plane_rows * 128;
    pp.din_planes = lgtc_faithful_planes(h->tc); pp.range_flag = lgtc_range_flag(h->tc);
    pp.ctrl = h->ctrl; pp.last_init = (do_stop || do_prune) ? 0 : L - 1;
    launch_k(k_lg_posenc, dim3(cdiv(std::max(maxrows, 1), 64), nseg), 256, 0, st, pp);
    ++h->launches;
    B2S_LAUNCH_CHECK();
  }
  h->dbg_nlayers = 0;
  if (maxm > 0 && maxn > 0) {
    B2S_TRY(lgtc_input_proj(h->tc, st, h->x[0], nseg, maxrows, h->ctrl, &h->launches));
    // with pruning enabled layer i lives in ping-pong buffer i & 1 (the gather after layer i moves the
    // survivors across); without it everything stays in buffer 0
    auto buf_of = [&](int i) { return do_prune ? (i & 1) : 0; };
    // upstream: scores > (1 - width_confidence) with a python double; undo the float rounding of the cfg
    const float keep_thr = (float)(1.0 - std::round((double)h->cfg.width_conf * 1e6) / 1e6);
    const int nblk = cdiv(maxrows, 32);
    for (int i = 0; i < L; ++i) {
      const int cur = buf_of(i);
      B2S_TRY(lgtc_layer(h->tc, st, i, h->x[cur], h->cosb[cur], h->sinb[cur], nseg, maxrows, h->ctrl, &h->launches));
      if (h->debug && h->dbg_layers) {   // debug only: snapshot pair 0's layer output and its live sizes (host sync)
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
      // two launches: heads per 32-row block (keep masks, one atomic per CTA), then decision + placement + copy
      HeadBlkParams hp = {};
      hp.x = h->x[cur]; hp.cap = cap;
      hp.wt = l.wtok; hp.bt = l.btok; hp.wm = l.wmatch; hp.bm = l.bmatch;
      hp.thr = h->thr[i]; hp.keep_thr = keep_thr; hp.use_tok = do_stop; hp.use_match = do_prune;
      hp.ctrl = h->ctrl; hp.adapt = h->adapt; hp.layer = i;
      launch_k(k_lg_heads_blk, dim3(nblk, nseg), 1024, 0, st, hp);
      const int nxt = cur ^ 1;
      GatherBlkParams gp = {};
      gp.adapt = h->adapt; gp.ctrl = h->ctrl; gp.cap = cap;
      gp.layer = i; gp.do_stop = do_stop ? 1 : 0; gp.do_prune = do_prune ? 1 : 0; gp.pruning_min_kpts = h->cfg.pruning_min_kpts;
      gp.nblk = nblk; gp.depth_conf = h->cfg.depth_conf;
      gp.x_in = h->x[cur]; gp.x_out = h->x[nxt]; gp.cos_in = h->cosb[cur]; gp.cos_out = h->cosb[nxt];
      gp.sin_in = h->sinb[cur]; gp.sin_out = h->sinb[nxt]; gp.ind_in = h->ind[cur]; gp.ind_out = h->ind[nxt];
      gp.prune = h->prune;
      gp.xb_out = lgtc_xb(h->tc); gp.xb_planes = lgtc_planes(h->tc); gp.xb_plane = plane_rows * 256; gp.range_flag = lgtc_range_flag(h->tc);
      launch_k(k_lg_gather_blk, dim3(nblk, nseg), 1024, 0, st, gp);
      h->launches += 2;
      B2S_LAUNCH_CHECK();
    }
    // ---- assignment with the last executed layer's heads (K14/K15); that layer index, its buffer and the
    //      live sizes are device-side values ----
    {
      FinalPrepParams fp = {};
      fp.x = h->x[0]; fp.x_odd = do_prune ? h->x[1] : nullptr; fp.cap = cap; fp.ctrl = h->ctrl;
      fp.wm_tab = h->wmatch_tab; fp.bm_tab = h->bmatch_tab; fp.tx = lgtc_tx(h->tc); fp.plane = plane_rows * 256; fp.ls = h->ls;
      fp.tx_planes = lgtc_faithful_planes(h->tc); fp.range_flag = lgtc_range_flag(h->tc);
      launch_k(k_lg_final_prep, dim3(cdiv(maxrows, 8), nseg), 256, 0, st, fp);
      ++h->launches;
      B2S_LAUNCH_CHECK();
    }
    B2S_TRY(lgtc_assignment(h->tc, st, np, maxm, maxn, h->ctrl, h->sim, h->simT, &h->launches));
    AssignParams ap = {};
    ap.sim = h->sim; ap.simT = h->simT; ap.ld = cap; ap.pair_stride = (size_t)cap * cap; ap.cap = cap; ap.ctrl = h->ctrl;
    ap.rmax = h->rmax; ap.rlog = h->rlog; ap.cmax = h->cmax; ap.clog = h->clog; ap.ls = h->ls;
    ap.max0 = h->max0; ap.m0 = h->m0; ap.m1 = h->m1;
    launch_k(k_lg_lse2, dim3(cdiv(maxrows, 8), 2 * np), 256, 0, st, ap);
    launch_k(k_lg_argmax2, dim3(cdiv(maxrows, 8), 2 * np), 256, 0, st, ap);
    h->launches += 2;
    B2S_LAUNCH_CHECK();
  }
  FilterParams fp = {};
  fp.out = out; fp.th = h->cfg.filter_thresh; fp.n_layers = L; fp.do_prune = do_prune ? 1 : 0;
  fp.ctrl = h->ctrl; fp.cap = cap; fp.max0 = h->max0; fp.m0 = h->m0; fp.m1 = h->m1;
  fp.ind = h->ind[0]; fp.ind_odd = do_prune ? h->ind[1] : nullptr; fp.prune = h->prune;
  fp.range_flag = h->planes == 2 ? lgtc_range_flag(h->tc) : nullptr;
  launch_k(k_lg_filter, np, 1024, 0, st, fp);
  ++h->launches;
  B2S_LAUNCH_CHECK();
  if (h->debug) {
    B2S_CUDA(cudaMemcpyAsync(h->h_ctrl, h->ctrl, LGC_INTS * sizeof(int), cudaMemcpyDeviceToHost, st));
    B2S_CUDA(cudaStreamSynchronize(st));
    h->dbg_simm = h->h_ctrl[LGC_M]; h->dbg_simn = h->h_ctrl[LGC_N];
  }
  return 0;
}

extern "C" int b2s_lightglue_match(b2s_lg* h, const float* k0, const float* d0, int m, const float* k1,
                                   const float* d1, int n, const float* size0, const float* size1, void* stream,
                                   int32_t* matches, float* mscores, int32_t* n_matches, int32_t* stop_layer,
                                   int32_t* matches0, int32_t* matches1, float* ms0, float* ms1,
                                   int32_t* prune0, int32_t* prune1) {
  if (!h || !matches || !mscores || !n_matches || m < 0 || n < 0) { set_error("b2s_lightglue_match: bad argument"); return B2S_EINVAL; }
  B2S_CUDA(cudaSetDevice(h->device));
  LgBatchIn in = {};
  LgBatchOut out = {};
  LgPairIn& pi = in.pr[0];
  pi.kp[0] = k0; pi.kp[1] = k1; pi.desc[0] = d0; pi.desc[1] = d1; pi.n[0] = m; pi.n[1] = n; pi.n_dev[0] = pi.n_dev[1] = nullptr;
  pi.has_size[0] = size0 != nullptr; pi.has_size[1] = size1 != nullptr;
  if (size0) { pi.size[0][0] = size0[0]; pi.size[0][1] = size0[1]; }
  if (size1) { pi.size[1][0] = size1[0]; pi.size[1][1] = size1[1]; }
  out.pr[0] = {matches, mscores, n_matches, stop_layer, matches0, matches1, ms0, ms1, prune0, prune1, m, n};
  return lg_run(h, (cudaStream_t)stream, in, out, 1);
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
  if (nm == B2S_LG_RANGE) {     // fp16x2 engine: a value left the fp16 range - the bf16x3 engine takes this pair
    b2s_lg* fb = nullptr;
    B2S_TRY(b2s_lg_fallback(h, &fb));
    if (!fb) { set_error("b2s_lightglue_match_host: range overflow without a fallback engine"); return B2S_ECUDA; }
    return b2s_lightglue_match_host(fb, k0, d0, m, k1, d1, n, size0, size1, matches, mscores, n_matches, stop_layer, matches0, matches1,
                                    ms0, ms1, prune0, prune1);
  }
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

// Batched matching: n_pairs pairs over packed varlen features, processed LG_MAXP pairs per launch sequence (the pair is a
// grid dimension of every kernel).  max_batch (0 = LG_MAXP) lowers the number of pairs per sequence, e.g. to keep a
// sequence's working set inside the 126 MB L2.
extern "C" int b2s_lightglue_match_batch_ex(b2s_lg* h, const float* kpts, const float* desc, const int32_t* cu, const int32_t* counts,
                                            const int32_t* counts_dev, int n_frames, const int32_t* pair_i, const int32_t* pair_j, int n_pairs,
                                            void* stream, int stride, int max_batch, int32_t* matches, float* mscores,
                                            int32_t* n_matches, int32_t* stop_layers) {
  if (!h || !cu || !pair_i || !pair_j || n_pairs < 0 || !matches || !mscores || !n_matches) { set_error("b2s_lightglue_match_batch: bad argument"); return B2S_EINVAL; }
  B2S_CUDA(cudaSetDevice(h->device));
  const int B = max_batch > 0 ? std::min(max_batch, LG_MAXP) : LG_MAXP;
  int maxkp = 1;
  auto rows_of = [&](int f) { return counts ? counts[f] : cu[f + 1] - cu[f]; };   // counts: frames need not be packed back to back
  for (int p = 0; p < n_pairs; ++p) {
    const int a = pair_i[p], b = pair_j[p];
    if (a < 0 || b < 0 || a >= n_frames || b >= n_frames) { set_error("pair %d out of range", p); return B2S_EINVAL; }
    const int m = rows_of(a), n = rows_of(b);
    if (m < 0 || n < 0) { set_error("cu is not ascending at pair %d", p); return B2S_EINVAL; }
    if (std::min(m, n) > stride) { set_error("stride %d too small for pair %d", stride, p); return B2S_ESIZE; }
    maxkp = std::max(maxkp, std::max(m, n));
  }
  B2S_TRY(b2s_lg_reserve(h, maxkp, std::min(B, std::max(n_pairs, 1))));
  for (int p0 = 0; p0 < n_pairs; p0 += B) {
    const int np = std::min(B, n_pairs - p0);
    LgBatchIn in = {};
    LgBatchOut out = {};
    for (int q = 0; q < np; ++q) {
      const int p = p0 + q, a = pair_i[p], b = pair_j[p];
      LgPairIn& pi = in.pr[q];
      pi.kp[0] = kpts + (size_t)cu[a] * 2; pi.desc[0] = desc + (size_t)cu[a] * 128; pi.n[0] = rows_of(a);
      pi.kp[1] = kpts + (size_t)cu[b] * 2; pi.desc[1] = desc + (size_t)cu[b] * 128; pi.n[1] = rows_of(b);
      if (counts_dev) { pi.n_dev[0] = counts_dev + a; pi.n_dev[1] = counts_dev + b; }
      out.pr[q] = {matches + (size_t)p * stride * 2, mscores + (size_t)p * stride, n_matches + p, stop_layers ? stop_layers + p : nullptr,
                   nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, pi.n[0], pi.n[1]};
    }
    B2S_TRY(lg_run(h, (cudaStream_t)stream, in, out, np));
  }
  return 0;
}

extern "C" int b2s_lightglue_match_batch(b2s_lg* h, const float* kpts, const float* desc, const int32_t* cu,
                                         int n_frames, const int32_t* pair_i, const int32_t* pair_j, int n_pairs,
                                         void* stream, int stride, int32_t* matches, float* mscores,
                                         int32_t* n_matches) {
  return b2s_lightglue_match_batch_ex(h, kpts, desc, cu, nullptr, nullptr, n_frames, pair_i, pair_j, n_pairs, stream, stride, 0, matches, mscores,
                                      n_matches, nullptr);
}

extern "C" int b2s_lg_debug_get(b2s_lg* h, const char* name, float* out, size_t cap_out, size_t* nout) {
  if (!h || !name || !nout) return B2S_EINVAL;
  B2S_CUDA(cudaSetDevice(h->device));
  B2S_CUDA(cudaDeviceSynchronize());
  const std::string s(name);
  const float* src = nullptr; size_t cnt = 0;
  std::vector<float> tmp;
  if (s.rfind("layer", 0) == 0 && h->dbg_layers) {
    // "layer<i>_<side>" -> [rows,256]  (pair 0)
    int li = 0, side = 0;
    if (sscanf(name, "layer%d_%d", &li, &side) != 2 || li < 0 || li >= h->dbg_nlayers || side < 0 || side > 1) {
      set_error("debug tensor %s not available", name); return B2S_EINVAL;
    }
    src = h->dbg_layers + ((size_t)li * 2 * h->cap + (side ? h->cap : 0)) * 256;
    cnt = (size_t)(side ? h->dbg_n[li] : h->dbg_m[li]) * 256;
  } else if (s == "ctrl") {
    // pair 0's device state after the last match (debug mode): LGC_* ints as floats - decision-margin reports read the
    // per-layer unconfident counts (exit-ratio margins) from it
    tmp.assign(h->h_ctrl, h->h_ctrl + LGC_INTS);
    *nout = tmp.size();
    std::memcpy(out, tmp.data(), std::min(cap_out, tmp.size()) * sizeof(float));
    return 0;
  } else if (s == "sim") {
    // compacted [m,n]  (pair 0)
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
