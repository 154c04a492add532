// common.cuh - shared host/device helpers for libb200slam (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/b200slam.h"

namespace b2s {

void set_error(const char* fmt, ...);

#define B2S_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      b2s::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return B2S_ECUDA;                                                                  \
    }                                                                                    \
  } while (0)

#define B2S_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != 0) return _r;    \
  } while (0)

#define B2S_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      b2s::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return B2S_ECUDA;                                                                  \
    }                                                                                    \
  } while (0)

// ---- weight blob ("B2SW", see weights.py::pack_state) -----------------------------------
struct TensorView {
  const float* data = nullptr;  // host pointer into the blob
  int ndim = 0;
  int dims[4] = {1, 1, 1, 1};
  size_t numel() const { return (size_t)dims[0] * dims[1] * dims[2] * dims[3]; }
};

struct WeightBlob {
  std::map<std::string, TensorView> t;
  int parse(const void* blob, size_t nbytes);
  // returns nullptr (and sets the error) when missing or when numel mismatches (expect>0)
  const TensorView* get(const std::string& name, size_t expect_numel = 0) const;
};

// ---- device memory owned by a handle ----------------------------------------------------
struct DeviceArena {
  std::vector<void*> ptrs;
  ~DeviceArena() { release(); }
  void release() {
    for (void* p : ptrs) cudaFree(p);
    ptrs.clear();
  }
  template <typename T>
  int alloc(T** out, size_t count) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (count ? count : 1) * sizeof(T));
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu bytes) -> %s", count * sizeof(T), cudaGetErrorString(e));
      return B2S_ENOMEM;
    }
    ptrs.push_back(p);
    *out = (T*)p;
    return 0;
  }
  template <typename T>
  int upload(T** out, const std::vector<T>& host) {
    B2S_TRY(alloc(out, host.size()));
    B2S_CUDA(cudaMemcpy(*out, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
  }
};

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// Every kernel of the library is launched with programmatic stream serialization: it may become
// resident (and run its prologue: barrier init, TMEM allocation, descriptor prefetch, weight
// staging) while the previous kernel of the stream is still draining.  Device side: pdl_trigger()
// lets the NEXT kernel start launching, pdl_wait() blocks until the PREVIOUS kernel has completed
// and its memory is visible - it must precede the first access to any buffer another kernel writes
// or reads (activations, workspaces, counters); constant weights may be read before it.
// (Measured: an early pdl_trigger() in the small kernels makes the pair slower - parked CTAs of the next kernel take
// issue slots from the running one - so only the tensor-core kernels, whose prologue is worth overlapping, trigger early.)
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();   // B2S_NO_PDL=1 in the environment turns the launch attribute off (debugging)

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface through B2S_LAUNCH_CHECK()
}
// same with a thread-block cluster of cluster_x CTAs along grid.x (grid.x must be a multiple of it)
template <typename... KArgs, typename... Args>
inline void launch_k_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster_x; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ---- per-kernel-class CUDA-event timing (bench.py's roofline.achieved is measured with it) --
enum ProfClass { PROF_ATTN = 0, PROF_GEMM = 1, PROF_NCLASS = 2 };
struct KernelProf {
  bool on = false;
  std::vector<cudaEvent_t> ev[PROF_NCLASS];
  size_t used[PROF_NCLASS] = {0, 0};
  ~KernelProf() {
    for (auto& v : ev) for (cudaEvent_t e : v) cudaEventDestroy(e);
  }
  void mark(int cls, cudaStream_t st) {   // call before and after a launch
    if (!on) return;
    if (used[cls] == ev[cls].size()) { cudaEvent_t e; cudaEventCreate(&e); ev[cls].push_back(e); }
    cudaEventRecord(ev[cls][used[cls]++], st);
  }
  // sums (stop - start) over the recorded pairs, then forgets them
  int read(int cls, double* ms, long long* n) {
    double tot = 0.0; long long cnt = 0;
    for (size_t i = 0; i + 1 < used[cls]; i += 2) {
      float t = 0.f;
      cudaError_t e = cudaEventSynchronize(ev[cls][i + 1]);
      if (e == cudaSuccess) e = cudaEventElapsedTime(&t, ev[cls][i], ev[cls][i + 1]);
      if (e != cudaSuccess) { set_error("profile read: %s", cudaGetErrorString(e)); return B2S_ECUDA; }
      tot += t; ++cnt;
    }
    used[cls] = 0;
    *ms = tot; *n = cnt;
    return 0;
  }
};

// ---- device helpers -----------------------------------------------------------------------
__device__ __forceinline__ float selu_f(float x) {
  // torch SELU: scale * (max(0,x) + min(0, alpha*(exp(x)-1)))
  const float alpha = 1.6732632423543772848170429916717f;
  const float scale = 1.0507009873554804934193349852946f;
  return x > 0.f ? scale * x : scale * (alpha * expm1f(x));
}
__device__ __forceinline__ float gelu_erf_f(float x) {
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float logsigmoid_f(float x) {
  // torch: min(0,x) - log1p(exp(-|x|))
  return fminf(0.f, x) - log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace b2s
