// lightglue_tc.cuh - tcgen05/TMEM path of the LightGlue layers (interface): bf16 operands (planes = 1),
// fp32 carried as three bf16 planes (planes = 3, fp32-faithful, any range) or as two fp16 planes (planes = 2,
// fp32-faithful at half the tensor-core work, values must stay inside the fp16 range - checked on the device).  Everything is batched over row SEGMENTS
// (lightglue_kernels.cuh): segment 2p + s = image s of pair p, `cap` rows apart.
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"
#include <vector>

namespace b2s {

struct LgTensorCore;
struct LgTcLayerSrc {   // fp32 device weights of one layer (self block then cross block)
  const float *wqkv, *bqkv, *wo, *bo, *w1, *b1, *lng, *lnb, *w2, *b2;
  const float *cwqkv, *cbqkv, *cwo, *cbo, *cw1, *cb1, *clng, *clnb, *cw2, *cb2;
};

int lgtc_create(LgTensorCore** out, size_t n_layers, int planes);
int lgtc_set_layer(LgTensorCore* tc, int layer, const LgTcLayerSrc& src);
int lgtc_set_input(LgTensorCore* tc, const float* w, const float* b);   // input_proj [256,128] + [256]
int lgtc_set_final(LgTensorCore* tc, const std::vector<const float*>& w, const std::vector<const float*>& b);
size_t lgtc_ws_bytes(int planes, int cap, int pairs);
int lgtc_alloc_ws(LgTensorCore* tc, int cap, int pairs);
// x = input_proj(descriptor planes din) for nseg segments: fp32 residual stream + its operand planes
int lgtc_input_proj(LgTensorCore* tc, cudaStream_t st, float* x, int nseg, int maxrows, const int* ctrl, long long* launches);
// one full transformer layer (self + cross) in place on x [segments*cap,256] fp32 master copy; maxrows bounds the
// live counts (grid sizes), the live counts / early-exit flags come from the per-pair device state `ctrl`
int lgtc_layer(LgTensorCore* tc, cudaStream_t st, int layer, float* x, const float* cosb, const float* sinb, int nseg,
               int maxrows, const int* ctrl, long long* launches);
// assignment head: md = final_proj_last(tx)/4 (both images), per pair sim = md0 md1^T and sim^T, always fp32-faithful (bf16x3)
int lgtc_assignment(LgTensorCore* tc, cudaStream_t st, int npairs, int maxm, int maxn, const int* ctrl, float* sim, float* simT,
                    long long* launches);
__nv_bfloat16* lgtc_xb(LgTensorCore* tc);    // bf16 plane copy of the residual stream (the pruning gather refreshes it)
__nv_bfloat16* lgtc_din(LgTensorCore* tc);   // [3][rows,128] descriptor planes (written by k_lg_posenc)
__nv_bfloat16* lgtc_tx(LgTensorCore* tc);    // [3][rows,256] planes of the final state (written by k_lg_final_prep)
int lgtc_planes(LgTensorCore* tc);
int lgtc_faithful_planes(LgTensorCore* tc);   // planes of din / tx / md: 3, or 2 on the fp16x2 path
int* lgtc_range_flag(LgTensorCore* tc);       // device [2]: sticky fp16 range flag, arrival counter (k_lg_filter)
void lgtc_destroy(LgTensorCore* tc);
void lgtc_set_prof(LgTensorCore* tc, KernelProf* prof, unsigned long long* stats);

}  // namespace b2s
