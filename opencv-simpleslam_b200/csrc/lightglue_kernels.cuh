// lightglue_kernels.cuh - bandwidth/latency-bound stages of the LightGlue matcher, batched over PAIRS:
// keypoint normalisation + learnable Fourier encoding (+ descriptor planes), token-confidence / matchability heads,
// point-pruning compaction, dual log-softmax assignment, mutual-NN filter.  Upstream spec: SURVEY.md Appendix A.3.
//
// Row space of every per-point buffer: SEGMENTS of `cap` rows.  Segment g = 2 * pair + image holds the points of one
// image of one pair at rows [g * cap, g * cap + live); so a batch of P pairs is 2P segments, point pruning shrinks each
// segment on its own, and one launch serves all pairs (blockIdx.y or .z = segment).
#pragma once
#include <cuda_bf16.h>
#include "tc_common.cuh"

namespace b2s {

// Device-resident matcher state, LGC_INTS ints PER PAIR: adaptive depth / width decisions never visit the host.
//   [LGC_STOP] early-exit flag (sticky)   [LGC_M],[LGC_N] live points of image 0 / 1
//   [LGC_CAN0],[LGC_CAN1] this layer's prune-eligibility per side   [LGC_LAST] last executed layer
//   [LGC_NPTS] original m + n   [LGC_UNCONF + i] #(token confidence < thr_i) after layer i (i < 16)   [LGC_M0],[LGC_N0] original m, n
enum { LGC_STOP = 1, LGC_M = 2, LGC_N = 3, LGC_CAN0 = 4, LGC_CAN1 = 5, LGC_LAST = 6, LGC_NPTS = 7, LGC_UNCONF = 8, LGC_M0 = 24, LGC_N0 = 25, LGC_INTS = 32 };
__device__ __forceinline__ bool lg_active(const int* c) { return !c[LGC_STOP] && c[LGC_M] > 0 && c[LGC_N] > 0; }

constexpr int LG_MAXP = 16;      // pairs per launch sequence (larger batches are chunked)
constexpr int LG_MAXBLK = 256;   // 32-row blocks per image (cap <= 8192)
constexpr int LG_ADAPT_INTS = 8 + 2 * LG_MAXBLK;

// per-pair inputs / outputs, passed BY VALUE to the first / last kernel of a batch (no host staging buffer to recycle)
// n[s]: point count of image s, or - when n_dev[s] is set - an upper bound of the device-resident count *n_dev[s] (the
// extractor's n_out): a frame stream then never waits for a count to reach the host
struct LgPairIn { const float* kp[2]; const float* desc[2]; const int* n_dev[2]; int n[2]; int has_size[2]; float size[2][2]; };
struct LgBatchIn { LgPairIn pr[LG_MAXP]; };
struct LgPairOut {
  int32_t* matches; float* mscores; int32_t* n_matches; int32_t* stop_layer;     // compact list (required), executed layers (nullable)
  int32_t* matches0; int32_t* matches1; float* ms0; float* ms1; int32_t* prune0; int32_t* prune1;   // full-size, nullable
  int m, n;                                                                      // original counts (upper bounds with LgPairIn::n_dev)
};
struct LgBatchOut { LgPairOut pr[LG_MAXP]; };

// ---------------------------------------------------------------------------------------
// K9: normalize_keypoints + posenc, the per-pair state reset, and the descriptors as three bf16 planes (operand of the
// input projection on the tensor cores).  grid = (ceil(max n / 64), 2 * pairs), block = 256.
// upstream: size = 1 + max - min (when no image_size); shift = size/2; scale = max(size)/2;
//           kn = (k - shift)/scale ; proj = Wr kn ; emb = (cos proj, sin proj)
// ---------------------------------------------------------------------------------------
struct PosencParams {
  LgBatchIn in;
  int cap;
  const float* Wr;           // [32,2]
  float* cosb; float* sinb;  // [rows,32]
  int* ind;                  // [rows]  identity index map
  int* prune;                // [rows]  per-point prune counters (by original index) -> 1
  __nv_bfloat16* din; size_t din_plane;   // [din_planes][rows,128] descriptor planes
  int din_planes; int* range_flag;        // 3: bf16x3; 2: fp16x2 (out-of-range values raise *range_flag)
  int* ctrl; int last_init;  // device state, reset here (last_init = L-1 when depth/width adaptivity is off)
};

__global__ void __launch_bounds__(256) k_lg_posenc(PosencParams p) {
  pdl_wait();
  const int g = blockIdx.y, pair = g >> 1, s = g & 1;   // blockIdx.x = chunk of 64 points (every CTA re-derives the extent)
  const LgPairIn& in = p.in.pr[pair];
  const int n0 = in.n_dev[0] ? max(0, min(in.n[0], *in.n_dev[0])) : in.n[0];
  const int n1 = in.n_dev[1] ? max(0, min(in.n[1], *in.n_dev[1])) : in.n[1];
  const int n = s ? n1 : n0;
  const float* kp = in.kp[s];
  if (blockIdx.x == 0 && s == 0 && threadIdx.x < LGC_INTS) {
    const int t = threadIdx.x;
    p.ctrl[pair * LGC_INTS + t] = (t == LGC_M || t == LGC_M0) ? n0 : (t == LGC_N || t == LGC_N0) ? n1 : t == LGC_LAST ? p.last_init : t == LGC_NPTS ? n0 + n1 : 0;
  }
  if (blockIdx.x * 64 >= n) return;
  __shared__ float red[4][8];
  __shared__ float sh_shift[2], sh_scale;
  float sx, sy;
  if (in.has_size[s]) {
    sx = in.size[s][0]; sy = in.size[s][1];
  } else {
    float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      float x = kp[2 * i], y = kp[2 * i + 1];
      mnx = fminf(mnx, x); mxx = fmaxf(mxx, x); mny = fminf(mny, y); mxy = fmaxf(mxy, y);
    }
    mnx = -warp_max(-mnx); mny = -warp_max(-mny); mxx = warp_max(mxx); mxy = warp_max(mxy);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[0][w] = mnx; red[1][w] = mny; red[2][w] = mxx; red[3][w] = mxy; }
    __syncthreads();
    mnx = red[0][0]; mny = red[1][0]; mxx = red[2][0]; mxy = red[3][0];
    for (int i = 1; i < 8; ++i) {
      mnx = fminf(mnx, red[0][i]); mny = fminf(mny, red[1][i]);
      mxx = fmaxf(mxx, red[2][i]); mxy = fmaxf(mxy, red[3][i]);
    }
    sx = 1.f + mxx - mnx; sy = 1.f + mxy - mny;
  }
  if (threadIdx.x == 0) {
    sh_shift[0] = sx / 2.f; sh_shift[1] = sy / 2.f; sh_scale = fmaxf(sx, sy) / 2.f;
  }
  __syncthreads();
  const float shx = sh_shift[0], shy = sh_shift[1], sc = sh_scale;
  // warp per point, lane = Fourier frequency: coalesced [row,32] cos/sin writes; lane = 4 descriptor channels
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float w0 = p.Wr[2 * lane], w1 = p.Wr[2 * lane + 1];
  const float* desc = in.desc[s];
  for (int i = blockIdx.x * 64 + warp; i < min(n, blockIdx.x * 64 + 64); i += 8) {
    const int r = g * p.cap + i;
    const float x = (kp[2 * i] - shx) / sc, y = (kp[2 * i + 1] - shy) / sc;
    if (lane == 0) { p.ind[r] = i; p.prune[r] = 1; }
    const float pr = w0 * x + w1 * y;
    p.cosb[(size_t)r * 32 + lane] = cosf(pr);
    p.sinb[(size_t)r * 32 + lane] = sinf(pr);
    const float4 d = *reinterpret_cast<const float4*>(desc + (size_t)i * 128 + lane * 4);
    if (p.din_planes == 2) {
      uint32_t wa[2], wb[2];
      tc::pack_planes2<2>(d.x, d.y, wa);
      tc::pack_planes2<2>(d.z, d.w, wb);
#pragma unroll
      for (int pl = 0; pl < 2; ++pl)
        *reinterpret_cast<uint2*>(p.din + pl * p.din_plane + (size_t)r * 128 + lane * 4) = make_uint2(wa[pl], wb[pl]);
      if ((tc::h2_ovf(wa[0]) | tc::h2_ovf(wb[0])) && p.range_flag) *p.range_flag = 1;
    } else {
      uint32_t wa[3], wb[3];
      tc::pack_planes2<3>(d.x, d.y, wa);
      tc::pack_planes2<3>(d.z, d.w, wb);
#pragma unroll
      for (int pl = 0; pl < 3; ++pl)
        *reinterpret_cast<uint2*>(p.din + pl * p.din_plane + (size_t)r * 128 + lane * 4) = make_uint2(wa[pl], wb[pl]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// K13: adaptive depth / width after layer i in two launches.  A CTA owns 32 consecutive rows of one segment.
//   tok = sigmoid(wt.x + bt); mat = sigmoid(wm.x + bm); keep = (mat > 1 - width_conf) | (tok <= thr)   (upstream
//   get_pruning_mask); ctrl[LGC_UNCONF + i] accumulates #(tok < thr) over both images (upstream check_if_stop)
//   heads : token / matchability heads of its rows -> 32-bit keep mask of the block, ONE atomic per CTA for the
//           unconfident count, and a snapshot of the live state (active, m, n) for the next kernel
//   gather: every CTA takes the (identical) early-exit decision from the final count:
//             stop  <=> 1 - #unconfident / (m + n) > depth_conf, m + n = ORIGINAL counts (pruned points count as confident)
//           then (pruning on) prefix-sums the block popcounts (<= 256 blocks: one warp) to place its rows and copies its
//           kept rows into the other ping-pong buffers; one designated CTA per segment publishes the new size.  A side that
//           is not eligible (count <= pruning_min_kpts) keeps every point.  The kernel never reads a ctrl word it writes.
// adapt (per pair): [0] active  [1] m  [2] n  [8 + s*LG_MAXBLK + b] keep mask of block b of image s
// ---------------------------------------------------------------------------------------
struct HeadBlkParams {
  const float* x; int cap;
  const float* wt; float bt; const float* wm; float bm;
  float thr; float keep_thr; int use_tok, use_match;
  int* ctrl; int* adapt; int layer;
};

__global__ void __launch_bounds__(1024) k_lg_heads_blk(HeadBlkParams p) {   // 32 warps: one row per warp
  pdl_wait();
  const int g = blockIdx.y, pair = g >> 1, s = g & 1, blk = blockIdx.x;
  int* ctrl = p.ctrl + pair * LGC_INTS;
  int* adapt = p.adapt + pair * LG_ADAPT_INTS;
  const bool active = lg_active(ctrl);
  if (blk == 0 && s == 0 && threadIdx.x == 0) { adapt[0] = active ? 1 : 0; adapt[1] = ctrl[LGC_M]; adapt[2] = ctrl[LGC_N]; }
  if (!active) return;
  __shared__ unsigned s_mask;
  __shared__ int s_unconf;
  if (threadIdx.x == 0) { s_mask = 0u; s_unconf = 0; }
  __syncthreads();
  const int rows = ctrl[LGC_M + s];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blk * 32 + warp;
  if (row < rows) {                            // uniform per warp
    const float* x = p.x + (size_t)(g * p.cap + row) * 256;
    const float4 a = *reinterpret_cast<const float4*>(x + lane * 4);
    const float4 b = *reinterpret_cast<const float4*>(x + 128 + lane * 4);
    float dt = 0.f, dm = 0.f;
    if (p.use_tok) {
      const float4 w0 = *reinterpret_cast<const float4*>(p.wt + lane * 4), w1 = *reinterpret_cast<const float4*>(p.wt + 128 + lane * 4);
      dt = a.x * w0.x + a.y * w0.y + a.z * w0.z + a.w * w0.w + b.x * w1.x + b.y * w1.y + b.z * w1.z + b.w * w1.w;
      dt = warp_sum(dt);
    }
    if (p.use_match) {
      const float4 w0 = *reinterpret_cast<const float4*>(p.wm + lane * 4), w1 = *reinterpret_cast<const float4*>(p.wm + 128 + lane * 4);
      dm = a.x * w0.x + a.y * w0.y + a.z * w0.z + a.w * w0.w + b.x * w1.x + b.y * w1.y + b.z * w1.z + b.w * w1.w;
      dm = warp_sum(dm);
    }
    if (lane == 0) {
      bool keep = p.use_match && sigmoid_f(dm + p.bm) > p.keep_thr;
      if (p.use_tok) {
        const float t = sigmoid_f(dt + p.bt);
        if (t < p.thr) atomicAdd(&s_unconf, 1);
        keep = keep || (t <= p.thr);
      }
      if (keep) atomicOr(&s_mask, 1u << warp);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    adapt[8 + s * LG_MAXBLK + blk] = (int)s_mask;
    if (s_unconf) atomicAdd(&ctrl[LGC_UNCONF + p.layer], s_unconf);
  }
}

struct GatherBlkParams {
  const int* adapt; int* ctrl; int cap;
  int layer, do_stop, do_prune, pruning_min_kpts, nblk; float depth_conf;
  const float* x_in; float* x_out; const float* cos_in; float* cos_out; const float* sin_in; float* sin_out;
  const int* ind_in; int* ind_out; int* prune;             // prune [rows] by original index
  __nv_bfloat16* xb_out; int xb_planes; size_t xb_plane;   // operand-plane copy of x (next layer's GEMMs): 1 / 3 bf16 planes, 2 fp16 planes
  int* range_flag;                                         // xb_planes == 2: raised by an out-of-range value
};

__global__ void __launch_bounds__(1024) k_lg_gather_blk(GatherBlkParams p) {   // 32 warps: one row per warp
  pdl_wait();
  const int g = blockIdx.y, pair = g >> 1, s = g & 1, blk = blockIdx.x;
  const int* adapt = p.adapt + pair * LG_ADAPT_INTS;
  int* ctrl = p.ctrl + pair * LGC_INTS;
  if (!adapt[0]) return;
  const bool lead = blk == 0 && s == 0 && threadIdx.x == 0;
  if (p.do_stop) {   // upstream check_if_stop; every CTA takes the same decision from the same final count
    const float ratio = 1.0f - (float)ctrl[LGC_UNCONF + p.layer] / (float)ctrl[LGC_NPTS];
    if (ratio > p.depth_conf) { if (lead) ctrl[LGC_STOP] = 1; return; }
  }
  if (lead) ctrl[LGC_LAST] = p.layer + 1;
  if (!p.do_prune) return;                        // the state stays where it is
  const int n = adapt[1 + s];
  const bool prune = n > p.pruning_min_kpts;      // a side that is not eligible keeps every point
  auto mask_of = [&](int b) -> unsigned {
    const int left = n - b * 32;
    if (left <= 0) return 0u;
    const unsigned live = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
    return prune ? ((unsigned)adapt[8 + s * LG_MAXBLK + b] & live) : live;
  };
  __shared__ int s_before, s_total;
  if (threadIdx.x < 32) {
    int before = 0, total = 0;
    for (int b = threadIdx.x; b < p.nblk; b += 32) {
      const int c = __popc(mask_of(b));
      total += c;
      if (b < blk) before += c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { before += __shfl_xor_sync(0xffffffffu, before, o); total += __shfl_xor_sync(0xffffffffu, total, o); }
    if (threadIdx.x == 0) { s_before = before; s_total = total; }
  }
  __syncthreads();
  if (blk == 0 && threadIdx.x == 0) { ctrl[LGC_M + s] = s_total; ctrl[LGC_CAN0 + s] = prune ? 1 : 0; }
  const unsigned mask = mask_of(blk);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (!((mask >> warp) & 1u)) return;             // uniform per warp
  const int base = g * p.cap;
  const int src = base + blk * 32 + warp;
  const int dst = base + s_before + __popc(mask & ((1u << warp) - 1u));
  const float4* xi = reinterpret_cast<const float4*>(p.x_in + (size_t)src * 256);
  float4* xo = reinterpret_cast<float4*>(p.x_out + (size_t)dst * 256);
  const float4 a = xi[2 * lane], b = xi[2 * lane + 1];
  xo[2 * lane] = a; xo[2 * lane + 1] = b;
  {
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (p.xb_planes == 1) store_planes8<1>(p.xb_out + (size_t)dst * 256 + lane * 8, p.xb_plane, f);
    else if (p.xb_planes == 2) { if (store_planes8<2>(p.xb_out + (size_t)dst * 256 + lane * 8, p.xb_plane, f) && p.range_flag) *p.range_flag = 1; }
    else store_planes8<3>(p.xb_out + (size_t)dst * 256 + lane * 8, p.xb_plane, f);
  }
  p.cos_out[(size_t)dst * 32 + lane] = p.cos_in[(size_t)src * 32 + lane];
  p.sin_out[(size_t)dst * 32 + lane] = p.sin_in[(size_t)src * 32 + lane];
  if (lane == 0) {
    const int orig = p.ind_in[src];
    p.ind_out[dst] = orig;
    if (prune) p.prune[base + orig] += 1;
  }
}

// ---------------------------------------------------------------------------------------
// Assignment head, preparation: the final state x of every live point -> three bf16 planes (operand of final_proj on the
// tensor cores) and ls = logsigmoid(matchability(x)), with the heads of the LAST EXECUTED layer ctrl[LGC_LAST] (its
// state lives in the odd ping-pong buffer when that index is odd and pruning is on).  Runs after an early exit too.
// One warp per row, grid = (ceil(rows/8), segments).
// ---------------------------------------------------------------------------------------
struct FinalPrepParams {
  const float* x; const float* x_odd;     // x_odd nullable (no pruning)
  int cap; const int* ctrl;
  const float* const* wm_tab; const float* bm_tab;   // matchability heads per layer
  __nv_bfloat16* tx; size_t plane;        // [tx_planes][rows,256]
  float* ls;                              // [rows]
  int tx_planes; int* range_flag;         // 3: bf16x3; 2: fp16x2 (+ range check)
};

__global__ void __launch_bounds__(256) k_lg_final_prep(FinalPrepParams p) {
  pdl_wait();
  const int g = blockIdx.y;
  const int* c = p.ctrl + (g >> 1) * LGC_INTS;
  if (c[LGC_M] <= 0 || c[LGC_N] <= 0) return;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= c[LGC_M + (g & 1)]) return;
  const int last = c[LGC_LAST];
  const float* xsrc = (p.x_odd && (last & 1)) ? p.x_odd : p.x;
  const float* wm = p.wm_tab[last];
  const int lane = threadIdx.x & 31;
  const size_t r = (size_t)(g * p.cap + row);
  const float4 a = *reinterpret_cast<const float4*>(xsrc + r * 256 + lane * 8), b = *reinterpret_cast<const float4*>(xsrc + r * 256 + lane * 8 + 4);
  const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  if (p.tx_planes == 2) { if (store_planes8<2>(p.tx + r * 256 + lane * 8, p.plane, f) && p.range_flag) *p.range_flag = 1; }
  else store_planes8<3>(p.tx + r * 256 + lane * 8, p.plane, f);
  const float4 w0 = *reinterpret_cast<const float4*>(wm + lane * 8), w1 = *reinterpret_cast<const float4*>(wm + lane * 8 + 4);
  float dm = a.x * w0.x + a.y * w0.y + a.z * w0.z + a.w * w0.w + b.x * w1.x + b.y * w1.y + b.z * w1.z + b.w * w1.w;
  dm = warp_sum(dm);
  if (lane == 0) p.ls[r] = logsigmoid_f(dm + p.bm_tab[last]);
}

// ---------------------------------------------------------------------------------------
// K14/K15: dual log-softmax statistics and arg-maxima over the materialised similarity.  The similarity GEMM writes sim
// AND sim^T, so the column-wise statistics are row-wise ones of the transposed matrix and both directions run in the same
// launch: blockIdx.y = 2 * pair + direction, one warp per row, 128-bit loads.
// assignment value, in upstream's association order:
//   scores = (log_softmax_row + log_softmax_col) + (logsigmoid(z0) + logsigmoid(z1))
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float assign_val(float s, float rm, float rl, float cm, float cl, float l0, float l1) {
  return (((s - rm) - rl) + ((s - cm) - cl)) + (l0 + l1);
}

struct AssignParams {
  const float* sim; const float* simT; int ld; size_t pair_stride;   // [pairs][cap, cap]
  int cap; const int* ctrl;
  float* rmax; float* rlog; float* cmax; float* clog;                // [pairs, cap]
  const float* ls;                                                   // [rows]: segment 2p = image 0, 2p + 1 = image 1
  float* max0; int* m0; int* m1;                                     // [pairs, cap]
};

__global__ void __launch_bounds__(256) k_lg_lse2(AssignParams p) {
  pdl_wait();
  const int pair = blockIdx.y >> 1, d = blockIdx.y & 1;
  const int* c = p.ctrl + pair * LGC_INTS;
  const int m = c[LGC_M + d], n = c[LGC_M + 1 - d];    // rows / columns in this direction
  if (c[LGC_M] <= 0 || c[LGC_N] <= 0) return;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= m) return;
  const int lane = threadIdx.x & 31;
  const float* r = (d ? p.simT : p.sim) + pair * p.pair_stride + (size_t)i * p.ld;
  const int n4 = n & ~3;
  float mx = -INFINITY;
  for (int j = lane * 4; j < n4; j += 128) {
    const float4 v = *reinterpret_cast<const float4*>(r + j);
    mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  for (int j = n4 + lane; j < n; j += 32) mx = fmaxf(mx, r[j]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane * 4; j < n4; j += 128) {
    const float4 v = *reinterpret_cast<const float4*>(r + j);
    sum += (expf(v.x - mx) + expf(v.y - mx)) + (expf(v.z - mx) + expf(v.w - mx));
  }
  for (int j = n4 + lane; j < n; j += 32) sum += expf(r[j] - mx);
  sum = warp_sum(sum);
  if (lane == 0) { (d ? p.cmax : p.rmax)[pair * p.cap + i] = mx; (d ? p.clog : p.rlog)[pair * p.cap + i] = logf(sum); }
}

__global__ void __launch_bounds__(256) k_lg_argmax2(AssignParams p) {
  pdl_wait();
  const int pair = blockIdx.y >> 1, d = blockIdx.y & 1;
  const int* c = p.ctrl + pair * LGC_INTS;
  const int m = c[LGC_M + d], n = c[LGC_M + 1 - d];
  if (c[LGC_M] <= 0 || c[LGC_N] <= 0) return;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= m) return;
  const int lane = threadIdx.x & 31;
  const float* r = (d ? p.simT : p.sim) + pair * p.pair_stride + (size_t)i * p.ld;
  const float* rmax = p.rmax + pair * p.cap; const float* rlog = p.rlog + pair * p.cap;
  const float* cmax = p.cmax + pair * p.cap; const float* clog = p.clog + pair * p.cap;
  const float* ls0 = p.ls + (size_t)(2 * pair) * p.cap; const float* ls1 = ls0 + p.cap;
  // direction 0: row i of image 0 against columns j of image 1; direction 1: the roles swap but assign_val keeps
  // upstream's operand order (row statistics first)
  const float* omax = d ? rmax : cmax; const float* olog = d ? rlog : clog; const float* ols = d ? ls0 : ls1;
  const float smax = (d ? cmax : rmax)[i], slog = (d ? clog : rlog)[i], sls = (d ? ls1 : ls0)[i];
  float best = -INFINITY; int bj = 0x7fffffff;
  auto consider = [&](float sv, int j) {
    const float v = d ? assign_val(sv, omax[j], olog[j], smax, slog, ols[j], sls) : assign_val(sv, smax, slog, omax[j], olog[j], sls, ols[j]);
    if (v > best) { best = v; bj = j; }
  };
  const int n4 = n & ~3;
  // the other side's statistics are contiguous in j: three 128-bit loads per four scores instead of twelve scalar ones
  auto consider4 = [&](float sv, float om, float ol, float os, int j) {
    const float v = d ? assign_val(sv, om, ol, smax, slog, os, sls) : assign_val(sv, smax, slog, om, ol, sls, os);
    if (v > best) { best = v; bj = j; }
  };
  for (int j = lane * 4; j < n4; j += 128) {
    const float4 v = *reinterpret_cast<const float4*>(r + j);
    const float4 om = *reinterpret_cast<const float4*>(omax + j), ol = *reinterpret_cast<const float4*>(olog + j), os = *reinterpret_cast<const float4*>(ols + j);
    consider4(v.x, om.x, ol.x, os.x, j); consider4(v.y, om.y, ol.y, os.y, j + 1);
    consider4(v.z, om.z, ol.z, os.z, j + 2); consider4(v.w, om.w, ol.w, os.w, j + 3);
  }
  for (int j = n4 + lane; j < n; j += 32) consider(r[j], j);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
    if (ob > best || (ob == best && oj < bj)) { best = ob; bj = oj; }
  }
  if (lane == 0) {
    if (d) p.m1[pair * p.cap + i] = bj == 0x7fffffff ? 0 : bj;
    else { p.max0[pair * p.cap + i] = best; p.m0[pair * p.cap + i] = bj == 0x7fffffff ? 0 : bj; }
  }
}

// K15c: upstream filter_matches + match list + every per-pair output.  One CTA of 1024 threads per pair.
// The full-size outputs are first set to "unmatched" for every ORIGINAL point (pruned points stay that way), then the
// live points scatter their results through the index map.
struct FilterParams {
  LgBatchOut out;
  float th; int n_layers, do_prune;
  const int* ctrl; int cap;
  const float* max0; const int* m0; const int* m1;   // [pairs, cap]
  const int* ind; const int* ind_odd;                // [rows] pruned index -> original index; ind_odd (nullable): odd ping-pong buffer
  const int* prune;                                  // [rows] prune counters by original index
  int* range_flag;   // nullable: [0] sticky "a value left the fp16 range during this launch sequence", [1] CTA arrival counter
};

__global__ void __launch_bounds__(1024) k_lg_filter(FilterParams p) {
  pdl_wait();
  const int pair = blockIdx.x;
  const LgPairOut& o = p.out.pr[pair];
  const int* c = p.ctrl + pair * LGC_INTS;
  const int tid = threadIdx.x;
  // fp16x2 path: a value left the fp16 range somewhere in this launch sequence -> every pair of it reports
  // n_matches = B2S_LG_RANGE (the host re-runs the batch on the bf16x3 engine).  The last CTA to read the flag clears it.
  __shared__ int s_range;
  if (p.range_flag) {
    if (tid == 0) {
      s_range = *reinterpret_cast<volatile int*>(p.range_flag);
      __threadfence();
      if (atomicAdd(p.range_flag + 1, 1) == (int)gridDim.x - 1) { p.range_flag[0] = 0; p.range_flag[1] = 0; }
    }
    __syncthreads();
    if (s_range) {
      if (tid == 0) { *o.n_matches = -2; if (o.stop_layer) *o.stop_layer = 0; }
      return;
    }
  }
  const int om = c[LGC_M0], on = c[LGC_N0];        // original counts (<= o.m, o.n)
  const bool empty_in = om <= 0 || on <= 0;
  // ---- defaults over the original points ----
  for (int i = tid; i < om; i += 1024) {
    if (o.matches0) o.matches0[i] = -1;
    if (o.ms0) o.ms0[i] = 0.f;
    if (o.prune0) o.prune0[i] = !p.do_prune ? p.n_layers : (empty_in ? 1 : p.prune[(size_t)(2 * pair) * p.cap + i]);
  }
  for (int j = tid; j < on; j += 1024) {
    if (o.matches1) o.matches1[j] = -1;
    if (o.ms1) o.ms1[j] = 0.f;
    if (o.prune1) o.prune1[j] = !p.do_prune ? p.n_layers : (empty_in ? 1 : p.prune[(size_t)(2 * pair + 1) * p.cap + j]);
  }
  if (empty_in) {
    if (tid == 0) { *o.n_matches = 0; if (o.stop_layer) *o.stop_layer = 1; }
    return;
  }
  const int m = c[LGC_M], n = c[LGC_N];
  const int last = c[LGC_LAST];
  // upstream: stop = i + 1 of the last layer entered; a side pruned to zero breaks out before the next layer
  if (tid == 0 && o.stop_layer) *o.stop_layer = (m > 0 && n > 0) ? last + 1 : last;
  if (m <= 0 || n <= 0) {
    if (tid == 0) *o.n_matches = 0;
    return;
  }
  __syncthreads();
  const int* ind = (p.ind_odd && (last & 1)) ? p.ind_odd : p.ind;
  const int* ind0 = ind + (size_t)(2 * pair) * p.cap; const int* ind1 = ind0 + p.cap;
  const float* max0 = p.max0 + pair * p.cap; const int* m0 = p.m0 + pair * p.cap; const int* m1 = p.m1 + pair * p.cap;
  __shared__ int wsum[32];
  __shared__ int carry;
  if (tid == 0) carry = 0;
  __syncthreads();
  const int lane = tid & 31, w = tid >> 5;
  for (int i0 = 0; i0 < m; i0 += 1024) {
    const int i = i0 + tid;
    int valid = 0, j = 0; float sc = 0.f;
    if (i < m) {
      j = m0[i];
      const bool mutual = (m1[j] == i);
      sc = mutual ? expf(max0[i]) : 0.f;
      valid = mutual && (sc > p.th);
      const int oi = ind0[i];
      if (o.ms0) o.ms0[oi] = sc;
      if (o.matches0 && valid) o.matches0[oi] = ind1[j];
    }
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    const int inwarp = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) wsum[w] = __popc(bal);
    __syncthreads();
    int off = carry;
    for (int q = 0; q < w; ++q) off += wsum[q];
    if (valid) {
      o.matches[2 * (off + inwarp)] = ind0[i];
      o.matches[2 * (off + inwarp) + 1] = ind1[j];
      o.mscores[off + inwarp] = sc;
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int q = 0; q < 32; ++q) t += wsum[q];
      carry += t;
    }
    __syncthreads();
  }
  if (tid == 0) *o.n_matches = carry;
  // side 1: mscores1 = mutual1 ? mscores0[m1] : 0 ; valid1 = mutual1 & valid0[m1]
  if (o.matches1 || o.ms1) {
    for (int j = tid; j < n; j += 1024) {
      const int i = m1[j];
      const bool mutual1 = (m0[i] == j);
      float sc = 0.f; bool valid1 = false;
      if (mutual1) { sc = expf(max0[i]); valid1 = sc > p.th; }
      const int oj = ind1[j];
      if (o.ms1) o.ms1[oj] = sc;
      if (o.matches1 && valid1) o.matches1[oj] = ind0[i];
    }
  }
}

}  // namespace b2s
