// lightglue_tc.cu - bf16 tcgen05/TMEM path of the LightGlue layer (not built yet: the
// entry points report an error so that nothing silently falls back to another path).
#include "lightglue_tc.cuh"

namespace b2s {
struct LgTensorCore { int dummy; };
int lgtc_create(LgTensorCore**, size_t) { set_error("bf16 tensor-core path is not available in this build"); return B2S_EINVAL; }
int lgtc_set_layer(LgTensorCore*, int, const LgTcLayerSrc&) { return B2S_EINVAL; }
int lgtc_alloc_ws(LgTensorCore*, int) { return B2S_EINVAL; }
int lgtc_layer(LgTensorCore*, cudaStream_t, int, float*, const float*, const float*, int, int, int, long long*) { return B2S_EINVAL; }
void lgtc_destroy(LgTensorCore*) {}
}  // namespace b2s
