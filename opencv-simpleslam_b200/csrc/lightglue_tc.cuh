// lightglue_tc.cuh - tcgen05/TMEM path of the LightGlue layer (interface): bf16 operands (planes = 1)
// or fp32 carried as three bf16 planes (planes = 3, fp32-faithful).
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
int lgtc_set_final(LgTensorCore* tc, const std::vector<const float*>& w, const std::vector<const float*>& b);
int lgtc_alloc_ws(LgTensorCore* tc, int cap);
// assignment head: md = final_proj_last(x)/4 (both images), sim = md0 md1^T, always fp32-faithful (bf16x3)
int lgtc_assignment(LgTensorCore* tc, cudaStream_t st, const float* x0, const float* x1, int cap, int m, int n, const int* ctrl,
                    float* sim, float* simT, long long* launches);
// one full transformer layer (self + cross) in place on x [2*cap,256] fp32 master copy; m, n bound the
// live counts (grid sizes), the live counts / early-exit flag come from the device state `ctrl`
int lgtc_layer(LgTensorCore* tc, cudaStream_t st, int layer, float* x, const float* cosb, const float* sinb, int cap,
               int m, int n, const int* ctrl, bool derive_xb, long long* launches);
__nv_bfloat16* lgtc_xb(LgTensorCore* tc);   // bf16 plane copy of the residual stream (the pruning gather refreshes it)
int lgtc_planes(LgTensorCore* tc);
void lgtc_destroy(LgTensorCore* tc);
void lgtc_set_prof(LgTensorCore* tc, KernelProf* prof, unsigned long long* stats);

}  // namespace b2s
