// lightglue_kernels.cuh - bandwidth/latency-bound stages of the LightGlue matcher (fp32):
// keypoint normalisation + learnable Fourier encoding, flash-style attention (FFMA),
// LayerNorm+GELU, token-confidence / matchability heads, point-pruning compaction, dual
// log-softmax assignment, mutual-NN filter.  Upstream spec: SURVEY.md Appendix A.3.
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace b2s {

// Device-resident matcher state (ints): adaptive depth / width decisions never visit the host.
//   [LGC_STOP] early-exit flag (sticky)   [LGC_M],[LGC_N] live points of image 0 / 1
//   [LGC_CAN0],[LGC_CAN1] this layer's prune-eligibility per side   [LGC_LAST] last executed layer
//   [LGC_UNCONF + i] #(token confidence < thr_i) after layer i
enum { LGC_STOP = 1, LGC_M = 2, LGC_N = 3, LGC_CAN0 = 4, LGC_CAN1 = 5, LGC_LAST = 6, LGC_UNCONF = 8, LGC_INTS = 32 };
__device__ __forceinline__ bool lg_active(const int* c) { return !c[LGC_STOP] && c[LGC_M] > 0 && c[LGC_N] > 0; }

// ---------------------------------------------------------------------------------------
// K9: normalize_keypoints + posenc.  grid = (ceil(max n / 64), 2 images), block = 256.
// upstream: size = 1 + max - min (when no image_size); shift = size/2; scale = max(size)/2;
//           kn = (k - shift)/scale ; proj = Wr kn ; emb = (cos proj, sin proj)
// ---------------------------------------------------------------------------------------
struct PosencParams {
  const float* kp[2]; int n[2]; int base[2];
  int has_size[2]; float size[2][2];
  const float* Wr;           // [32,2]
  float* kn;                 // [rows,2]
  float* cosb; float* sinb;  // [rows,32]
  int* ind;                  // [rows]  identity index map
  int* prune[2];             // per image [n] (nullable) -> 1
  int* ctrl; int last_init;  // device state, reset here (last_init = L-1 when depth/width adaptivity is off)
};

__global__ void __launch_bounds__(256) k_lg_posenc(PosencParams p) {
  pdl_wait();
  const int s = blockIdx.y;   // image; blockIdx.x = chunk of 64 points (every CTA re-derives the extent)
  const int n = p.n[s];
  const float* kp = p.kp[s];
  if (blockIdx.x == 0 && s == 0 && threadIdx.x < LGC_INTS) {
    const int t = threadIdx.x;
    p.ctrl[t] = t == LGC_M ? p.n[0] : t == LGC_N ? p.n[1] : t == LGC_LAST ? p.last_init : 0;
  }
  __shared__ float red[4][8];
  __shared__ float sh_shift[2], sh_scale;
  float sx, sy;
  if (p.has_size[s]) {
    sx = p.size[s][0]; sy = p.size[s][1];
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
  // warp per point, lane = Fourier frequency: coalesced [row,32] cos/sin writes
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float w0 = p.Wr[2 * lane], w1 = p.Wr[2 * lane + 1];
  for (int i = blockIdx.x * 64 + warp; i < min(n, blockIdx.x * 64 + 64); i += 8) {
    const int r = p.base[s] + i;
    const float x = (kp[2 * i] - shx) / sc, y = (kp[2 * i + 1] - shy) / sc;
    if (lane == 0) {
      p.kn[2 * r] = x; p.kn[2 * r + 1] = y;
      p.ind[r] = i;
      if (p.prune[s]) p.prune[s][i] = 1;
    }
    const float pr = w0 * x + w1 * y;
    p.cosb[(size_t)r * 32 + lane] = cosf(pr);
    p.sinb[(size_t)r * 32 + lane] = sinf(pr);
  }
}

// ---------------------------------------------------------------------------------------
// K11/K12 attention core (fp32 FFMA, online softmax).  O = softmax(Q K^T * scale) V
// grid = (ceil(max nq / 64), heads, nprob), block = 256, dyn smem = 4 * 64 * 68 * 4 B.
// Thread (ty,tx) of a 16x16 grid owns S[ty*4..+4][tx*4..+4] and O[ty*4..+4][tx*4..+4].
// ---------------------------------------------------------------------------------------
struct AttnProb { const float* Q; const float* K; const float* V; float* O; int nq, nk; };
struct AttnParams { AttnProb prob[2]; int ldq, ldk, ldv, ldo; float scale; const int* ctrl; int cross; unsigned long long* stats; };

constexpr int ATT_B = 64, ATT_D = 64, ATT_LD = 68;
constexpr int ATT_SMEM = 4 * ATT_B * ATT_LD * (int)sizeof(float);

__global__ void __launch_bounds__(256) k_attn_fp32(AttnParams p) {
  pdl_wait();
  extern __shared__ __align__(16) float att_smem[];
  float (*Qs)[ATT_LD] = reinterpret_cast<float (*)[ATT_LD]>(att_smem);                       // [d][q]
  float (*Ks)[ATT_LD] = reinterpret_cast<float (*)[ATT_LD]>(att_smem + ATT_B * ATT_LD);      // [d][k]
  float (*Vs)[ATT_LD] = reinterpret_cast<float (*)[ATT_LD]>(att_smem + 2 * ATT_B * ATT_LD);  // [k][d]
  float (*Ps)[ATT_LD] = reinterpret_cast<float (*)[ATT_LD]>(att_smem + 3 * ATT_B * ATT_LD);  // [k][q]

  AttnProb pr = p.prob[blockIdx.z];
  if (p.ctrl) {             // live sizes come from the device state (pruning / early exit)
    if (!lg_active(p.ctrl)) return;
    pr.nq = p.ctrl[LGC_M + blockIdx.z];
    pr.nk = p.ctrl[LGC_M + (p.cross ? 1 - blockIdx.z : blockIdx.z)];
  }
  const int q0 = blockIdx.x * ATT_B;
  if (q0 >= pr.nq) return;
  if (p.stats && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)
    atomicAdd(&p.stats[p.cross ? 1 : 0], (unsigned long long)pr.nq * (unsigned long long)pr.nk);
  const int hoff = blockIdx.y * ATT_D;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

  // Q tile -> Qs[d][q]
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int f = tid + i * 256, r = f >> 4, dq = f & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < pr.nq) v = *reinterpret_cast<const float4*>(pr.Q + (size_t)(q0 + r) * p.ldq + hoff + dq * 4);
    Qs[dq * 4 + 0][r] = v.x; Qs[dq * 4 + 1][r] = v.y; Qs[dq * 4 + 2][r] = v.z; Qs[dq * 4 + 3][r] = v.w;
  }

  float o[4][4], mrow[4], lrow[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    mrow[i] = -INFINITY; lrow[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }

  for (int k0 = 0; k0 < pr.nk; k0 += ATT_B) {
    __syncthreads();  // previous P.V done (and Q stores visible on the first pass)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int f = tid + i * 256, r = f >> 4, dq = f & 15;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (k0 + r < pr.nk) {
        kv = *reinterpret_cast<const float4*>(pr.K + (size_t)(k0 + r) * p.ldk + hoff + dq * 4);
        vv = *reinterpret_cast<const float4*>(pr.V + (size_t)(k0 + r) * p.ldv + hoff + dq * 4);
      }
      Ks[dq * 4 + 0][r] = kv.x; Ks[dq * 4 + 1][r] = kv.y; Ks[dq * 4 + 2][r] = kv.z; Ks[dq * 4 + 3][r] = kv.w;
      *reinterpret_cast<float4*>(&Vs[r][dq * 4]) = vv;
    }
    __syncthreads();

    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
    for (int d = 0; d < ATT_D; ++d) {
      const float4 a = *reinterpret_cast<const float4*>(&Qs[d][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Ks[d][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(av[i], bv[j], s[i][j]);
    }
    // online softmax
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = (k0 + tx * 4 + j < pr.nk) ? s[i][j] * p.scale : -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float mnew = fmaxf(mrow[i], mx);
      const float corr = expf(mrow[i] - mnew);
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) { s[i][j] = expf(s[i][j] - mnew); sum += s[i][j]; }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
      lrow[i] = lrow[i] * corr + sum;
      mrow[i] = mnew;
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] *= corr;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(&Ps[tx * 4 + j][ty * 4]) = make_float4(s[0][j], s[1][j], s[2][j], s[3][j]);
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < ATT_B; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&Ps[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Vs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = fmaf(av[i], bv[j], o[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = q0 + ty * 4 + i;
    if (r >= pr.nq) continue;
    const float inv = pr.nk > 0 ? 1.f / lrow[i] : 0.f;
    *reinterpret_cast<float4*>(pr.O + (size_t)r * p.ldo + hoff + tx * 4) =
        make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
  }
}

// ---------------------------------------------------------------------------------------
// FFN middle: LayerNorm(512, eps 1e-5, affine) + GELU(erf), in place.  One warp per row.
// ---------------------------------------------------------------------------------------
struct RowSeg { int base[2]; int rows[2]; };

__global__ void __launch_bounds__(256) k_ln_gelu_512(float* h, RowSeg seg, const float* gamma, const float* beta, const int* ctrl) {
  pdl_wait();
  const int s = blockIdx.y;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (ctrl) { if (!lg_active(ctrl)) return; seg.rows[s] = ctrl[LGC_M + s]; }
  if (row >= seg.rows[s]) return;
  const int lane = threadIdx.x & 31;
  float* x = h + (size_t)(seg.base[s] + row) * 512;
  float4 v[4];
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    v[t] = *reinterpret_cast<const float4*>(x + t * 128 + lane * 4);
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
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    const float4 b = *reinterpret_cast<const float4*>(beta + c);
    float4 o;
    o.x = gelu_erf_f((v[t].x - mean) * rstd * g.x + b.x);
    o.y = gelu_erf_f((v[t].y - mean) * rstd * g.y + b.y);
    o.z = gelu_erf_f((v[t].z - mean) * rstd * g.z + b.z);
    o.w = gelu_erf_f((v[t].w - mean) * rstd * g.w + b.w);
    *reinterpret_cast<float4*>(x + c) = o;
  }
}

// ---------------------------------------------------------------------------------------
// K13: per-row heads.  tok = sigmoid(wt.x + bt); mat = sigmoid(wm.x + bm).  One warp/row.
// ctrl[0] accumulates #(tok < thr) over both images (upstream check_if_stop).
// keep[row] = (mat > 1 - width_conf) | (tok <= thr)        (upstream get_pruning_mask)
// Also used (mode z) to produce logsigmoid(z) for the assignment.
// ---------------------------------------------------------------------------------------
struct HeadParams {
  const float* x; RowSeg seg;
  const float* wt; float bt; const float* wm; float bm;
  float thr; float keep_thr; int use_tok; int use_match;
  float* tok; int* keep; int* ctrl; int layer;
  float* ls_pos;   // if set (assignment mode): write logsigmoid(wm.x+bm) and nothing else; the
                   // matchability head of layer ctrl[LGC_LAST] comes from wm_tab / bm_tab and the
                   // state from x_alt when that layer lives in the odd ping-pong buffer
  const float* const* wm_tab; const float* bm_tab; const float* x_alt;
};

__global__ void __launch_bounds__(256) k_lg_heads(HeadParams p) {
  pdl_wait();
  const int s = blockIdx.y;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (p.ls_pos) {
    if (p.ctrl[LGC_M] <= 0 || p.ctrl[LGC_N] <= 0) return;
    const int last = p.ctrl[LGC_LAST];
    p.wm = p.wm_tab[last]; p.bm = p.bm_tab[last];
    if (p.x_alt && (last & 1)) p.x = p.x_alt;
  } else if (!lg_active(p.ctrl)) {
    return;
  }
  p.seg.rows[s] = p.ctrl[LGC_M + s];
  if (row >= p.seg.rows[s]) return;
  const int lane = threadIdx.x & 31;
  const int r = p.seg.base[s] + row;
  const float* x = p.x + (size_t)r * 256;
  const float4 a = *reinterpret_cast<const float4*>(x + lane * 4);
  const float4 b = *reinterpret_cast<const float4*>(x + 128 + lane * 4);
  float dt = 0.f, dm = 0.f;
  if (p.use_tok) {
    const float4 w0 = *reinterpret_cast<const float4*>(p.wt + lane * 4);
    const float4 w1 = *reinterpret_cast<const float4*>(p.wt + 128 + lane * 4);
    dt = a.x * w0.x + a.y * w0.y + a.z * w0.z + a.w * w0.w + b.x * w1.x + b.y * w1.y + b.z * w1.z + b.w * w1.w;
    dt = warp_sum(dt);
  }
  if (p.use_match || p.ls_pos) {
    const float4 w0 = *reinterpret_cast<const float4*>(p.wm + lane * 4);
    const float4 w1 = *reinterpret_cast<const float4*>(p.wm + 128 + lane * 4);
    dm = a.x * w0.x + a.y * w0.y + a.z * w0.z + a.w * w0.w + b.x * w1.x + b.y * w1.y + b.z * w1.z + b.w * w1.w;
    dm = warp_sum(dm);
  }
  if (lane != 0) return;
  if (p.ls_pos) { p.ls_pos[r] = logsigmoid_f(dm + p.bm); return; }
  bool keep = false;
  if (p.use_match) keep = sigmoid_f(dm + p.bm) > p.keep_thr;
  if (p.use_tok) {
    const float t = sigmoid_f(dt + p.bt);
    p.tok[r] = t;
    if (t < p.thr) atomicAdd(&p.ctrl[LGC_UNCONF + p.layer], 1);
    keep = keep || (t <= p.thr);
  }
  p.keep[r] = keep ? 1 : 0;
}

// After layer i: early-exit test (upstream check_if_stop) and the order-preserving compaction map of
// the points that survive pruning (upstream get_pruning_mask).  grid = 2 (one CTA per image), block = 1024.
//   stop  <=> 1 - #unconfident / (m + n) > depth_conf, m + n = ORIGINAL counts (pruned points count as confident)
//   srcmap[base + dst] = src.  A side that is not eligible (count <= pruning_min_kpts) keeps every point.
// On stop nothing else is written: the state stays in this layer's buffer and LGC_LAST stays i.
struct ScanParams {
  const int* keep; int* srcmap; int* ctrl; int base[2];
  int layer, num_points, do_stop, do_prune, pruning_min_kpts; float depth_conf;
};

__global__ void __launch_bounds__(1024) k_lg_prune_scan(ScanParams p) {
  pdl_wait();
  const int s = blockIdx.x;
  __shared__ int sh_go, sh_n;
  if (threadIdx.x == 0) {
    int go = lg_active(p.ctrl) ? 1 : 0;
    if (go && p.do_stop) {
      const float ratio = 1.0f - (float)p.ctrl[LGC_UNCONF + p.layer] / (float)p.num_points;
      if (ratio > p.depth_conf) { go = 0; if (s == 0) p.ctrl[LGC_STOP] = 1; }
    }
    sh_go = go; sh_n = p.ctrl[LGC_M + s];
  }
  __syncthreads();
  if (!sh_go) return;
  if (s == 0 && threadIdx.x == 0) p.ctrl[LGC_LAST] = p.layer + 1;
  if (!p.do_prune) return;
  const int n = sh_n, base = p.base[s];
  const bool prune = n > p.pruning_min_kpts;
  __shared__ int wsum[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    const int k = (i < n) ? (prune ? p.keep[base + i] : 1) : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, k);
    const int inwarp = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) wsum[w] = __popc(bal);
    __syncthreads();
    int off = carry;
    for (int j = 0; j < w; ++j) off += wsum[j];
    if (k) p.srcmap[base + off + inwarp] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int j = 0; j < 32; ++j) t += wsum[j];
      carry += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { p.ctrl[LGC_M + s] = carry; p.ctrl[LGC_CAN0 + s] = prune ? 1 : 0; }
}

// ---------------------------------------------------------------------------------------
// Adaptive depth / width in two launches (pruning enabled): k_lg_heads_blk + k_lg_gather_blk replace
// k_lg_heads + k_lg_prune_scan + k_lg_gather.  A CTA owns 32 consecutive rows of one image.
//   heads : token / matchability heads of its rows -> 32-bit keep mask of the block, ONE atomic per CTA for the
//           unconfident count, and a snapshot of the live state (active, m, n) for the next kernel
//   gather: every CTA takes the (identical) early-exit decision from the final count, prefix-sums the block
//           popcounts (<= 256 blocks: one warp) to place its rows, and copies its kept rows; one designated CTA
//           per image publishes the new size.  The kernel never reads a ctrl word it writes (snapshot instead).
// adapt: [0] active  [1] m  [2] n  [8 + s*LG_MAXBLK + b] keep mask of block b of image s
// ---------------------------------------------------------------------------------------
constexpr int LG_MAXBLK = 256;   // 32-row blocks per image (cap <= 8192)
struct HeadBlkParams {
  const float* x; int base[2];
  const float* wt; float bt; const float* wm; float bm;
  float thr; float keep_thr; int use_tok;
  float* tok; int* ctrl; int* adapt; int layer;
};

__global__ void __launch_bounds__(1024) k_lg_heads_blk(HeadBlkParams p) {   // 32 warps: one row per warp
  pdl_wait();
  const int s = blockIdx.y, blk = blockIdx.x;
  const bool active = lg_active(p.ctrl);
  if (blk == 0 && s == 0 && threadIdx.x == 0) { p.adapt[0] = active ? 1 : 0; p.adapt[1] = p.ctrl[LGC_M]; p.adapt[2] = p.ctrl[LGC_N]; }
  if (!active) return;
  __shared__ unsigned s_mask;
  __shared__ int s_unconf;
  if (threadIdx.x == 0) { s_mask = 0u; s_unconf = 0; }
  __syncthreads();
  const int rows = p.ctrl[LGC_M + s];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4 wt0 = *reinterpret_cast<const float4*>(p.wt + lane * 4), wt1 = *reinterpret_cast<const float4*>(p.wt + 128 + lane * 4);
  const float4 wm0 = *reinterpret_cast<const float4*>(p.wm + lane * 4), wm1 = *reinterpret_cast<const float4*>(p.wm + 128 + lane * 4);
  {
    const int rb = warp, row = blk * 32 + rb;
    if (row < rows) {                            // uniform per warp
    const int r = p.base[s] + row;
    const float* x = p.x + (size_t)r * 256;
    const float4 a = *reinterpret_cast<const float4*>(x + lane * 4);
    const float4 b = *reinterpret_cast<const float4*>(x + 128 + lane * 4);
    float dt = 0.f;
    if (p.use_tok) {
      dt = a.x * wt0.x + a.y * wt0.y + a.z * wt0.z + a.w * wt0.w + b.x * wt1.x + b.y * wt1.y + b.z * wt1.z + b.w * wt1.w;
      dt = warp_sum(dt);
    }
    float dm = a.x * wm0.x + a.y * wm0.y + a.z * wm0.z + a.w * wm0.w + b.x * wm1.x + b.y * wm1.y + b.z * wm1.z + b.w * wm1.w;
    dm = warp_sum(dm);
    if (lane == 0) {
      bool keep = sigmoid_f(dm + p.bm) > p.keep_thr;
      if (p.use_tok) {
        const float t = sigmoid_f(dt + p.bt);
        p.tok[r] = t;
        if (t < p.thr) atomicAdd(&s_unconf, 1);
        keep = keep || (t <= p.thr);
      }
      if (keep) atomicOr(&s_mask, 1u << rb);
    }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    p.adapt[8 + s * LG_MAXBLK + blk] = (int)s_mask;
    if (s_unconf) atomicAdd(&p.ctrl[LGC_UNCONF + p.layer], s_unconf);
  }
}

struct GatherBlkParams {
  const int* adapt; int* ctrl; int base[2];
  int layer, num_points, do_stop, pruning_min_kpts, nblk; float depth_conf;
  const float* x_in; float* x_out; const float* cos_in; float* cos_out; const float* sin_in; float* sin_out;
  const int* ind_in; int* ind_out; int* prune[2];
  __nv_bfloat16* xb_out; int xb_planes; size_t xb_plane;   // bf16 plane copy of x (tensor-core paths)
};

__global__ void __launch_bounds__(1024) k_lg_gather_blk(GatherBlkParams p) {   // 32 warps: one row per warp
  pdl_wait();
  if (!p.adapt[0]) return;
  const int s = blockIdx.y, blk = blockIdx.x;
  const bool lead = blk == 0 && s == 0 && threadIdx.x == 0;
  if (p.do_stop) {   // upstream check_if_stop; every CTA takes the same decision from the same final count
    const float ratio = 1.0f - (float)p.ctrl[LGC_UNCONF + p.layer] / (float)p.num_points;
    if (ratio > p.depth_conf) { if (lead) p.ctrl[LGC_STOP] = 1; return; }
  }
  if (lead) p.ctrl[LGC_LAST] = p.layer + 1;
  const int n = p.adapt[1 + s];
  const bool prune = n > p.pruning_min_kpts;      // a side that is not eligible keeps every point
  auto mask_of = [&](int b) -> unsigned {
    const int left = n - b * 32;
    if (left <= 0) return 0u;
    const unsigned live = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
    return prune ? ((unsigned)p.adapt[8 + s * LG_MAXBLK + b] & live) : live;
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
  if (blk == 0 && threadIdx.x == 0) { p.ctrl[LGC_M + s] = s_total; p.ctrl[LGC_CAN0 + s] = prune ? 1 : 0; }
  const unsigned mask = mask_of(blk);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const int rb = warp;
    if (!((mask >> rb) & 1u)) return;             // uniform per warp
    const int src = p.base[s] + blk * 32 + rb;
    const int dst = p.base[s] + s_before + __popc(mask & ((1u << rb) - 1u));
    const float4* xi = reinterpret_cast<const float4*>(p.x_in + (size_t)src * 256);
    float4* xo = reinterpret_cast<float4*>(p.x_out + (size_t)dst * 256);
    const float4 a = xi[2 * lane], b = xi[2 * lane + 1];
    xo[2 * lane] = a; xo[2 * lane + 1] = b;
    if (p.xb_out) {
      float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      for (int pl = 0; pl < p.xb_planes; ++pl) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __nv_bfloat162 hb = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
          w[j] = *reinterpret_cast<const uint32_t*>(&hb);
          f[2 * j] -= __uint_as_float(w[j] << 16); f[2 * j + 1] -= __uint_as_float(w[j] & 0xFFFF0000u);
        }
        *reinterpret_cast<uint4*>(p.xb_out + pl * p.xb_plane + (size_t)dst * 256 + lane * 8) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    p.cos_out[(size_t)dst * 32 + lane] = p.cos_in[(size_t)src * 32 + lane];
    p.sin_out[(size_t)dst * 32 + lane] = p.sin_in[(size_t)src * 32 + lane];
    if (lane == 0) {
      const int orig = p.ind_in[src];
      p.ind_out[dst] = orig;
      if (p.prune[s] && prune) p.prune[s][orig] += 1;
    }
  }
}

// gather the surviving rows into the other ping-pong buffers (layer i lives in buffer i & 1 when
// pruning is enabled).  grid = (ceil(maxrows/8), 2), block 256 (warp per row).  Optionally also
// writes the bf16 copy of the residual stream that the tensor-core path feeds to TMA.
struct GatherParams {
  const int* srcmap; const int* ctrl; int base[2];
  const float* x_in; float* x_out; const float* cos_in; float* cos_out; const float* sin_in; float* sin_out;
  const int* ind_in; int* ind_out; int* prune[2];
  __nv_bfloat16* xb_out; int xb_planes; size_t xb_plane;   // bf16 plane copy of x (tensor-core paths)
};

__global__ void __launch_bounds__(256) k_lg_gather(GatherParams p) {
  pdl_wait();
  if (p.ctrl[LGC_STOP]) return;          // exit fired in this layer's scan: keep the unpruned state where it is
  const int s = blockIdx.y;
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= p.ctrl[LGC_M + s]) return;
  const int lane = threadIdx.x & 31;
  const int src = p.base[s] + p.srcmap[p.base[s] + j], dst = p.base[s] + j;
  const float4* xi = reinterpret_cast<const float4*>(p.x_in + (size_t)src * 256);
  float4* xo = reinterpret_cast<float4*>(p.x_out + (size_t)dst * 256);
  const float4 a = xi[2 * lane], b = xi[2 * lane + 1];
  xo[2 * lane] = a; xo[2 * lane + 1] = b;
  if (p.xb_out) {
    float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    for (int pl = 0; pl < p.xb_planes; ++pl) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 hb = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
        w[j] = *reinterpret_cast<const uint32_t*>(&hb);
        f[2 * j] -= __uint_as_float(w[j] << 16); f[2 * j + 1] -= __uint_as_float(w[j] & 0xFFFF0000u);
      }
      *reinterpret_cast<uint4*>(p.xb_out + pl * p.xb_plane + (size_t)dst * 256 + lane * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  p.cos_out[(size_t)dst * 32 + lane] = p.cos_in[(size_t)src * 32 + lane];
  p.sin_out[(size_t)dst * 32 + lane] = p.sin_in[(size_t)src * 32 + lane];
  if (lane == 0) {
    const int orig = p.ind_in[src];
    p.ind_out[dst] = orig;
    if (p.prune[s] && p.ctrl[LGC_CAN0 + s]) p.prune[s][orig] += 1;
  }
}

// ---------------------------------------------------------------------------------------
// K14: dual log-softmax statistics over the materialised similarity sim[m,n] (ld).
// row pass: one warp per row -> rmax[i], rlog[i] = log(sum exp(sim - rmax)).
// col pass: CTA = 32 columns x 32 row-lanes -> cmax[j], clog[j].
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_lg_row_lse(const float* sim, int ld, const int* ctrl, float* rmax, float* rlog) {
  pdl_wait();
  const int m = ctrl[LGC_M], n = ctrl[LGC_N];
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= m) return;
  const int lane = threadIdx.x & 31;
  const float* r = sim + (size_t)i * ld;
  float mx = -INFINITY;
  for (int j = lane; j < n; j += 32) mx = fmaxf(mx, r[j]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < n; j += 32) sum += expf(r[j] - mx);
  sum = warp_sum(sum);
  if (lane == 0) { rmax[i] = mx; rlog[i] = logf(sum); }
}

__global__ void __launch_bounds__(1024) k_lg_col_lse(const float* sim, int ld, const int* ctrl, float* cmax, float* clog) {
  pdl_wait();
  const int m = ctrl[LGC_M], n = ctrl[LGC_N];
  __shared__ float smx[32][33], ssm[32][33];
  const int tx = threadIdx.x & 31, tyy = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  float mx = -INFINITY, sum = 0.f;
  if (j < n) {
    for (int i = tyy; i < m; i += 32) mx = fmaxf(mx, sim[(size_t)i * ld + j]);
  }
  smx[tyy][tx] = mx;
  __syncthreads();
  float cm = -INFINITY;
  for (int t = 0; t < 32; ++t) cm = fmaxf(cm, smx[t][tx]);
  if (j < n) {
    for (int i = tyy; i < m; i += 32) sum += expf(sim[(size_t)i * ld + j] - cm);
  }
  ssm[tyy][tx] = sum;
  __syncthreads();
  if (tyy == 0 && j < n) {
    float t = 0.f;
    for (int q = 0; q < 32; ++q) t += ssm[q][tx];
    cmax[j] = cm; clog[j] = logf(t);
  }
}

// assignment value, in upstream's association order:
//   scores = (log_softmax_row + log_softmax_col) + (logsigmoid(z0) + logsigmoid(z1))
__device__ __forceinline__ float assign_val(float s, float rm, float rl, float cm, float cl, float l0, float l1) {
  return (((s - rm) - rl) + ((s - cm) - cl)) + (l0 + l1);
}

// K15a: row-wise max/argmax (first index on ties).  One warp per row.
__global__ void __launch_bounds__(256) k_lg_row_argmax(const float* sim, int ld, const int* ctrl, const float* rmax,
                                                       const float* rlog, const float* cmax, const float* clog,
                                                       const float* ls0, const float* ls1, float* max0, int* m0) {
  pdl_wait();
  const int m = ctrl[LGC_M], n = ctrl[LGC_N];
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= m) return;
  const int lane = threadIdx.x & 31;
  const float* r = sim + (size_t)i * ld;
  const float rm = rmax[i], rl = rlog[i], l0 = ls0[i];
  float best = -INFINITY; int bj = 0x7fffffff;
  for (int j = lane; j < n; j += 32) {
    const float v = assign_val(r[j], rm, rl, cmax[j], clog[j], l0, ls1[j]);
    if (v > best) { best = v; bj = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
    if (ob > best || (ob == best && oj < bj)) { best = ob; bj = oj; }
  }
  if (lane == 0) { max0[i] = best; m0[i] = bj == 0x7fffffff ? 0 : bj; }
}

// K15b: column-wise argmax.  CTA = 32 columns x 32 row-lanes.
__global__ void __launch_bounds__(1024) k_lg_col_argmax(const float* sim, int ld, const int* ctrl, const float* rmax,
                                                        const float* rlog, const float* cmax, const float* clog,
                                                        const float* ls0, const float* ls1, int* m1) {
  pdl_wait();
  const int m = ctrl[LGC_M], n = ctrl[LGC_N];
  __shared__ float sv[32][33];
  __shared__ int si[32][33];
  const int tx = threadIdx.x & 31, tyy = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  float best = -INFINITY; int bi = 0x7fffffff;
  if (j < n) {
    const float cm = cmax[j], cl = clog[j], l1 = ls1[j];
    for (int i = tyy; i < m; i += 32) {
      const float v = assign_val(sim[(size_t)i * ld + j], rmax[i], rlog[i], cm, cl, ls0[i], l1);
      if (v > best) { best = v; bi = i; }
    }
  }
  sv[tyy][tx] = best; si[tyy][tx] = bi;
  __syncthreads();
  if (tyy == 0 && j < n) {
    for (int t = 1; t < 32; ++t) {
      const float ob = sv[t][tx]; const int oi = si[t][tx];
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    m1[j] = bi == 0x7fffffff ? 0 : bi;
  }
}

// ---------------------------------------------------------------------------------------
// K14/K15 on the tensor-core path: the similarity GEMM also writes sim^T, so the column-wise statistics are row-wise
// ones of the transposed matrix and both directions run in the same launch (blockIdx.y = direction), one warp per row,
// 128-bit loads.  Same per-element arithmetic as the four kernels above.
// ---------------------------------------------------------------------------------------
struct Lse2Params { const float* sim; const float* simT; int ld; const int* ctrl; float* rmax; float* rlog; float* cmax; float* clog; };

__global__ void __launch_bounds__(256) k_lg_lse2(Lse2Params p) {
  pdl_wait();
  const int d = blockIdx.y;
  const int m = p.ctrl[LGC_M + d], n = p.ctrl[LGC_M + 1 - d];    // rows / columns in this direction
  if (p.ctrl[LGC_M] <= 0 || p.ctrl[LGC_N] <= 0) return;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= m) return;
  const int lane = threadIdx.x & 31;
  const float* r = (d ? p.simT : p.sim) + (size_t)i * p.ld;
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
  if (lane == 0) { (d ? p.cmax : p.rmax)[i] = mx; (d ? p.clog : p.rlog)[i] = logf(sum); }
}

struct Argmax2Params {
  const float* sim; const float* simT; int ld; const int* ctrl;
  const float* rmax; const float* rlog; const float* cmax; const float* clog; const float* ls0; const float* ls1;
  float* max0; int* m0; int* m1;
};

__global__ void __launch_bounds__(256) k_lg_argmax2(Argmax2Params p) {
  pdl_wait();
  const int d = blockIdx.y;
  const int m = p.ctrl[LGC_M + d], n = p.ctrl[LGC_M + 1 - d];
  if (p.ctrl[LGC_M] <= 0 || p.ctrl[LGC_N] <= 0) return;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= m) return;
  const int lane = threadIdx.x & 31;
  const float* r = (d ? p.simT : p.sim) + (size_t)i * p.ld;
  // direction 0: row i of image 0 against columns j of image 1; direction 1: the roles swap but assign_val keeps
  // upstream's operand order (row statistics first)
  const float* omax = d ? p.rmax : p.cmax; const float* olog = d ? p.rlog : p.clog; const float* ols = d ? p.ls0 : p.ls1;
  const float smax = (d ? p.cmax : p.rmax)[i], slog = (d ? p.clog : p.rlog)[i], sls = (d ? p.ls1 : p.ls0)[i];
  float best = -INFINITY; int bj = 0x7fffffff;
  auto consider = [&](float sv, int j) {
    const float v = d ? assign_val(sv, omax[j], olog[j], smax, slog, ols[j], sls) : assign_val(sv, smax, slog, omax[j], olog[j], sls, ols[j]);
    if (v > best) { best = v; bj = j; }
  };
  const int n4 = n & ~3;
  for (int j = lane * 4; j < n4; j += 128) {
    const float4 v = *reinterpret_cast<const float4*>(r + j);
    consider(v.x, j); consider(v.y, j + 1); consider(v.z, j + 2); consider(v.w, j + 3);
  }
  for (int j = n4 + lane; j < n; j += 32) consider(r[j], j);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
    if (ob > best || (ob == best && oj < bj)) { best = ob; bj = oj; }
  }
  if (lane == 0) {
    if (d) p.m1[i] = bj == 0x7fffffff ? 0 : bj;
    else { p.max0[i] = best; p.m0[i] = bj == 0x7fffffff ? 0 : bj; }
  }
}

// K15c: upstream filter_matches + match list.  Single CTA of 1024 threads.
struct FilterParams {
  int m, n; float th;                        // m, n are read from ctrl on the device
  const int* ctrl; int cap; int32_t* stop_layer;   // stop_layer (device, nullable) = executed layers
  const float* max0; const int* m0; const int* m1;
  const int* ind0; const int* ind1;          // pruned-index -> original index (image 1 at + cap)
  const int* ind_alt;                        // same for the odd ping-pong buffer (nullptr: no pruning)
  int32_t* matches; float* mscores; int32_t* n_matches;   // compact outputs
  int32_t* matches0; int32_t* matches1; float* ms0; float* ms1;  // full-size (nullable)
};

__global__ void __launch_bounds__(1024) k_lg_filter(FilterParams p) {
  pdl_wait();
  p.m = p.ctrl[LGC_M]; p.n = p.ctrl[LGC_N];
  {
    const int last = p.ctrl[LGC_LAST];
    if (p.ind_alt && (last & 1)) { p.ind0 = p.ind_alt; p.ind1 = p.ind_alt + p.cap; }
    // upstream: stop = i + 1 of the last layer entered; a side pruned to zero breaks out before the next layer
    if (threadIdx.x == 0 && p.stop_layer) *p.stop_layer = (p.m > 0 && p.n > 0) ? last + 1 : last;
    if (p.m <= 0 || p.n <= 0) {
      if (threadIdx.x == 0) *p.n_matches = 0;
      return;
    }
  }
  __shared__ int wsum[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i0 = 0; i0 < p.m; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    int valid = 0, j = 0; float sc = 0.f;
    if (i < p.m) {
      j = p.m0[i];
      const bool mutual = (p.m1[j] == i);
      sc = mutual ? expf(p.max0[i]) : 0.f;
      valid = mutual && (sc > p.th);
      const int oi = p.ind0[i];
      if (p.ms0) p.ms0[oi] = sc;
      if (p.matches0) p.matches0[oi] = valid ? p.ind1[j] : -1;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    const int inwarp = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) wsum[w] = __popc(bal);
    __syncthreads();
    int off = carry;
    for (int q = 0; q < w; ++q) off += wsum[q];
    if (valid) {
      p.matches[2 * (off + inwarp)] = p.ind0[i];
      p.matches[2 * (off + inwarp) + 1] = p.ind1[j];
      p.mscores[off + inwarp] = sc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int q = 0; q < 32; ++q) t += wsum[q];
      carry += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *p.n_matches = carry;
  // side 1: mscores1 = mutual1 ? mscores0[m1] : 0 ; valid1 = mutual1 & valid0[m1]
  if (p.matches1 || p.ms1) {
    for (int j = threadIdx.x; j < p.n; j += 1024) {
      const int i = p.m1[j];
      const bool mutual1 = (p.m0[i] == j);
      float sc = 0.f; bool valid1 = false;
      if (mutual1) {
        const bool mutual0 = true;  // m0[i]==j and m1[j]==i
        sc = mutual0 ? expf(p.max0[i]) : 0.f;
        valid1 = sc > p.th;
      }
      const int oj = p.ind1[j];
      if (p.ms1) p.ms1[oj] = sc;
      if (p.matches1) p.matches1[oj] = valid1 ? p.ind0[i] : -1;
    }
  }
}

// fill helpers
__global__ void k_fill_i32(int32_t* p, int n, int32_t v) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void k_fill_f32(float* p, int n, float v) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace b2s
