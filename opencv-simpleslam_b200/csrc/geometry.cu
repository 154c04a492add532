// geometry.cu - host orchestration behind b2s_fm_* (fundamental-matrix RANSAC, replaces the cv2.findFundamentalMat
// call of `filter_matches_ransac`, /root/reference/slam/core/features_utils.py:185-200) and b2s_remap_* (cv2.remap of
// the undistortion maps, /root/reference/slam/monocular/main_revamped.py:313-324).
#include "geometry_kernels.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>

using namespace b2s;

struct b2s_fm {
  int device = 0, max_pts = 0, max_hyp = 0;
  DeviceArena arena;
  float4* xy = nullptr; double* norm = nullptr; double* models = nullptr; int32_t *nmodels = nullptr, *counts = nullptr;
  // host API staging (device copies of the caller's host arrays + pinned result block)
  float *d_pts1 = nullptr, *d_pts2 = nullptr; uint8_t* d_mask = nullptr; double* d_F = nullptr; int32_t* d_res = nullptr;
  uint8_t* h_pin = nullptr;    // pinned: [max_pts] mask | 9 doubles | 2 int32
  int32_t* d_subsets = nullptr;   // cv2-identical path: [max_hyp][7] sample indices
  uint8_t* h_cv = nullptr;        // pinned: pts1 | pts2 ([max_pts][2] f32 each) | subsets [max_hyp][7] | counts [max_hyp][3] | nmodels [max_hyp]
  unsigned sign1 = 0, sign2 = 0;  // completion vectors of OpenCV's SVD (see k_fmcv_models)
  cudaStream_t stream = nullptr;
  long long launches = 0;
};


// ---------------------------------------------------------------------------------------------------------------
// cv2-identical RANSAC, host part: OpenCV's sample stream and the sequential bookkeeping of its loop
// (modules/calib3d/src/ptsetreg.cpp, fundam.cpp; restated in oracle/cv_ransac.py and pinned there against cv2 itself)
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct CvRng {   // cv::RNG: multiply-with-carry
  uint64_t state;
  explicit CvRng(uint64_t s) : state(s ? s : 0xffffffffull) {}
  unsigned next() { state = (uint64_t)(unsigned)state * 4164903690u + (unsigned)(state >> 32); return (unsigned)state; }
  int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

// JacobiSVDImpl_ completes Vt with vectors whose components are +-1/m, sign = (rng.next() & 256) of RNG(0x12345678)
void fmcv_signs(unsigned* s1, unsigned* s2) {
  CvRng rng(0x12345678ull);
  unsigned a = 0, b = 0;
  for (int k = 0; k < 9; ++k) a |= (rng.next() & 256u) ? 1u << k : 0u;
  for (int k = 0; k < 9; ++k) b |= (rng.next() & 256u) ? 1u << k : 0u;
  *s1 = a; *s2 = b;
}

// fundam.cpp haveCollinearPoints on the 7 sampled points (only the last one is tested; float32 differences)
bool fmcv_collinear(const float* pts, const int* idx) {
  const int i = 6;
  const float xi = pts[2 * idx[i]], yi = pts[2 * idx[i] + 1];
  for (int j = 0; j < i; ++j) {
    const double dx1 = pts[2 * idx[j]] - xi, dy1 = pts[2 * idx[j] + 1] - yi;
    for (int k = 0; k < j; ++k) {
      const double dx2 = pts[2 * idx[k]] - xi, dy2 = pts[2 * idx[k] + 1] - yi;
      if (std::fabs(dx2 * dy1 - dy2 * dx1) <= (double)FLT_EPSILON * (std::fabs(dx1) + std::fabs(dy1) + std::fabs(dx2) + std::fabs(dy2))) return true;
    }
  }
  return false;
}

// the subsets of the next `count` iterations of the loop (getSubset with 10000 attempts), continuing the call's RNG
// stream; returns how many were drawn before getSubset failed
int fmcv_subsets(CvRng& rng, const float* pts1, const float* pts2, int n, int count, int32_t* out) {
  for (int it = 0; it < count; ++it) {
    int* idx = out + 7 * it;
    bool found = false;
    for (int attempt = 0; attempt < 10000 && !found; ++attempt) {
      for (int i = 0; i < 7; ++i) {
        int c;
        for (;;) {
          c = rng.uniform(0, n);
          bool dup = false;
          for (int j = 0; j < i; ++j) dup |= idx[j] == c;
          if (!dup) break;
        }
        idx[i] = c;
      }
      found = !fmcv_collinear(pts1, idx) && !fmcv_collinear(pts2, idx);
    }
    if (!found) return it;
  }
  return count;
}

int fmcv_update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = std::max(p, 0.); p = std::min(p, 1.);
  ep = std::max(ep, 0.); ep = std::min(ep, 1.);
  double num = std::max(1. - p, DBL_MIN);
  double denom = 1. - std::pow(1. - ep, model_points);
  if (denom < DBL_MIN) return 0;
  num = std::log(num);
  denom = std::log(denom);
  return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)std::lrint(num / denom);
}
}  // namespace

extern "C" int b2s_fm_create(int device, int max_points, int max_hypotheses, b2s_fm** out) {
  if (!out || max_points < 8 || max_hypotheses < 1) { set_error("b2s_fm_create: bad arguments"); return B2S_EINVAL; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    set_error("b2s_fm_create: no CUDA device %d (there is no CPU fallback)", device);
    return B2S_ENODEV;
  }
  B2S_CUDA(cudaSetDevice(device));
  b2s_fm* h = new b2s_fm();
  h->device = device; h->max_pts = max_points; h->max_hyp = max_hypotheses;
  int rc = 0;
  auto A = [&](int r) { if (rc == 0) rc = r; };
  A(h->arena.alloc(&h->xy, (size_t)max_points));
  A(h->arena.alloc(&h->norm, 6));
  A(h->arena.alloc(&h->models, (size_t)max_hypotheses * 27));
  A(h->arena.alloc(&h->nmodels, (size_t)max_hypotheses));
  A(h->arena.alloc(&h->counts, (size_t)max_hypotheses * 3));
  A(h->arena.alloc(&h->d_pts1, (size_t)max_points * 2));
  A(h->arena.alloc(&h->d_pts2, (size_t)max_points * 2));
  A(h->arena.alloc(&h->d_mask, (size_t)max_points));
  A(h->arena.alloc(&h->d_F, 9));
  A(h->arena.alloc(&h->d_res, 2));
  A(h->arena.alloc(&h->d_subsets, (size_t)max_hypotheses * 7));
  if (rc == 0 && cudaMallocHost((void**)&h->h_pin, (size_t)max_points + 128) != cudaSuccess) { set_error("b2s_fm_create: pinned alloc failed"); rc = B2S_ENOMEM; }
  if (rc == 0 && cudaMallocHost((void**)&h->h_cv, (size_t)max_points * 16 + (size_t)max_hypotheses * 11 * sizeof(int32_t)) != cudaSuccess) { set_error("b2s_fm_create: pinned alloc failed"); rc = B2S_ENOMEM; }
  fmcv_signs(&h->sign1, &h->sign2);
  if (rc == 0 && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("b2s_fm_create: stream"); rc = B2S_ECUDA; }
  if (rc != 0) { if (h->h_pin) cudaFreeHost(h->h_pin); if (h->h_cv) cudaFreeHost(h->h_cv); delete h; return rc; }
  *out = h;
  return 0;
}

extern "C" void b2s_fm_destroy(b2s_fm* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  if (h->h_cv) cudaFreeHost(h->h_cv);
  delete h;
}

extern "C" long long b2s_fm_launch_count(const b2s_fm* h) { return h ? h->launches : 0; }

extern "C" int b2s_fm_ransac(b2s_fm* h, const float* pts1_dev, const float* pts2_dev, const int32_t* pairs_dev, int n,
                             float thresh, int n_hyp, uint64_t seed, void* stream, uint8_t* mask_dev, double* F_dev,
                             int32_t* result_dev) {
  if (!h || !pts1_dev || !pts2_dev || !mask_dev || !F_dev || !result_dev) { set_error("b2s_fm_ransac: null argument"); return B2S_EINVAL; }
  if (n < 7) { set_error("b2s_fm_ransac: need at least 7 correspondences, got %d", n); return B2S_EINVAL; }
  if (n > h->max_pts || n_hyp > h->max_hyp || n_hyp < 1) { set_error("b2s_fm_ransac: n=%d / n_hyp=%d exceed the handle (%d / %d)", n, n_hyp, h->max_pts, h->max_hyp); return B2S_ESIZE; }
  B2S_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  FmParams p;
  p.pts1 = pts1_dev; p.pts2 = pts2_dev; p.pairs = pairs_dev; p.n = n; p.n_hyp = n_hyp; p.seed = seed;
  p.thresh2 = (double)thresh * (double)thresh;
  p.xy = h->xy; p.norm = h->norm; p.models = h->models; p.nmodels = h->nmodels; p.counts = h->counts;
  p.mask = mask_dev; p.F = F_dev; p.result = result_dev;
  launch_k(k_fm_prepare, dim3(1), dim3(256), 0, st, p);
  launch_k(k_fm_hypotheses, dim3(cdiv(n_hyp, 64)), dim3(64), 0, st, p);
  launch_k(k_fm_score, dim3(n_hyp), dim3(128), 0, st, p);
  launch_k(k_fm_select, dim3(1), dim3(256), 0, st, p);
  h->launches += 4;
  B2S_LAUNCH_CHECK();
  return 0;
}

extern "C" int b2s_fm_ransac_host(b2s_fm* h, const float* pts1, const float* pts2, int n, float thresh, int n_hyp,
                                  uint64_t seed, uint8_t* mask, double* F, int32_t* n_inliers, int32_t* model_index) {
  if (!h || !pts1 || !pts2 || !mask) { set_error("b2s_fm_ransac_host: null argument"); return B2S_EINVAL; }
  if (n > h->max_pts) { set_error("b2s_fm_ransac_host: n=%d exceeds the handle (%d)", n, h->max_pts); return B2S_ESIZE; }
  B2S_CUDA(cudaSetDevice(h->device));
  B2S_CUDA(cudaMemcpyAsync(h->d_pts1, pts1, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
  B2S_CUDA(cudaMemcpyAsync(h->d_pts2, pts2, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
  B2S_TRY(b2s_fm_ransac(h, h->d_pts1, h->d_pts2, nullptr, n, thresh, n_hyp, seed, h->stream, h->d_mask, h->d_F, h->d_res));
  uint8_t* pm = h->h_pin;                                           // pinned: mask | 9 doubles | 2 int32
  const size_t off = ((size_t)h->max_pts + 15) & ~(size_t)15;
  double* hF = reinterpret_cast<double*>(pm + off);
  int32_t* hres = reinterpret_cast<int32_t*>(pm + off + 9 * sizeof(double));
  B2S_CUDA(cudaMemcpyAsync(pm, h->d_mask, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  B2S_CUDA(cudaMemcpyAsync(hF, h->d_F, 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  B2S_CUDA(cudaMemcpyAsync(hres, h->d_res, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  B2S_CUDA(cudaStreamSynchronize(h->stream));
  std::memcpy(mask, pm, (size_t)n);
  if (F) std::memcpy(F, hF, 9 * sizeof(double));
  if (n_inliers) *n_inliers = hres[0];
  if (model_index) *model_index = hres[1];
  return 0;
}


// cv2.findFundamentalMat(pts1, pts2, FM_RANSAC, thresh, confidence[, max_iters]) for n >= 15: same mask, same F
extern "C" int b2s_fm_cv_ransac_host(b2s_fm* h, const float* pts1, const float* pts2, int n, double thresh, double confidence,
                                     int max_iters, uint8_t* mask, double* F, int32_t* n_inliers, int32_t* info4) {
  if (!h || !pts1 || !pts2 || !mask) { set_error("b2s_fm_cv_ransac_host: null argument"); return B2S_EINVAL; }
  if (n < 15) { set_error("b2s_fm_cv_ransac_host: OpenCV runs RANSAC from 15 correspondences on (LMedS below), got %d", n); return B2S_EINVAL; }
  if (max_iters < 1) max_iters = 1;
  if (n > h->max_pts || max_iters > h->max_hyp) { set_error("b2s_fm_cv_ransac_host: n=%d / max_iters=%d exceed the handle (%d / %d)", n, max_iters, h->max_pts, h->max_hyp); return B2S_ESIZE; }
  if (thresh <= 0) thresh = 3;
  if (confidence < DBL_EPSILON || confidence > 1 - DBL_EPSILON) confidence = 0.99;
  B2S_CUDA(cudaSetDevice(h->device));
  float* hp1 = reinterpret_cast<float*>(h->h_cv);
  float* hp2 = hp1 + (size_t)h->max_pts * 2;
  int32_t* hsub = reinterpret_cast<int32_t*>(hp2 + (size_t)h->max_pts * 2);
  int32_t* hcnt = hsub + (size_t)h->max_hyp * 7;
  int32_t* hnm = hcnt + (size_t)h->max_hyp * 3;
  std::memcpy(hp1, pts1, (size_t)n * 8);
  std::memcpy(hp2, pts2, (size_t)n * 8);
  B2S_CUDA(cudaMemcpyAsync(h->d_pts1, hp1, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
  B2S_CUDA(cudaMemcpyAsync(h->d_pts2, hp2, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
  if (n_inliers) *n_inliers = 0;
  if (info4) { info4[0] = -1; info4[1] = -1; info4[2] = max_iters; info4[3] = 0; }
  FmCvParams p = {};
  p.m1 = h->d_pts1; p.m2 = h->d_pts2; p.n = n; p.sign1 = h->sign1; p.sign2 = h->sign2;
  p.thresh2 = (float)(thresh * thresh);
  p.mask = h->d_mask; p.F = h->d_F;
  // RANSACPointSetRegistrator::run in WAVES of iterations (64, 256, the rest): with a high inlier ratio the adaptive
  // bound drops to a handful of iterations after the first good model and the later waves are never drawn or solved.
  // The subsets of a wave are drawn on the host while the previous copies / kernels are in flight.
  CvRng rng((uint64_t)-1);
  int niters = max_iters, max_good = 0, win_it = -1, win_k = -1, drawn = 0;
  bool exhausted = false;
  const int wave_end[3] = {std::min(64, max_iters), std::min(320, max_iters), max_iters};
  for (int w = 0; w < 3 && drawn < niters && !exhausted; ++w) {
    const int lo = drawn, want = std::min(wave_end[w], niters) - lo;
    if (want <= 0) continue;
    const int got = fmcv_subsets(rng, pts1, pts2, n, want, hsub + (size_t)lo * 7);
    exhausted = got < want;
    drawn += got;
    if (got == 0) break;
    B2S_CUDA(cudaMemcpyAsync(h->d_subsets + (size_t)lo * 7, hsub + (size_t)lo * 7, (size_t)got * 7 * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    p.n_sub = got; p.subsets = h->d_subsets + (size_t)lo * 7;
    p.models = h->models + (size_t)lo * 27; p.nmodels = h->nmodels + lo; p.counts = h->counts + (size_t)lo * 3;
    launch_k(k_fmcv_models, dim3(cdiv(got, 32)), dim3(32), 0, h->stream, p);
    launch_k(k_fmcv_count, dim3(got), dim3(128), 0, h->stream, p);
    h->launches += 2;
    B2S_CUDA(cudaMemcpyAsync(hcnt + (size_t)lo * 3, h->counts + (size_t)lo * 3, (size_t)got * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    B2S_CUDA(cudaStreamSynchronize(h->stream));
    B2S_LAUNCH_CHECK();
    // the sequential part of the loop over this wave's counts
    for (int it = lo; it < lo + got && it < niters; ++it) {
      for (int k = 0; k < 3; ++k) {
        const int good = hcnt[3 * it + k];
        if (good < 0) break;
        if (good > std::max(max_good, 6)) {
          max_good = good; win_it = it; win_k = k;
          niters = fmcv_update_num_iters(confidence, (double)(n - good) / n, 7, niters);
        }
      }
    }
  }
  (void)hnm;
  if (info4) info4[3] = drawn;
  if (drawn == 0) {            // getSubset failed in iteration 0: cv2 returns no model
    B2S_CUDA(cudaStreamSynchronize(h->stream));
    std::memset(mask, 0, (size_t)n);
    return 0;
  }
  if (info4) { info4[0] = win_it; info4[1] = win_k; info4[2] = niters; }
  if (max_good <= 0) { std::memset(mask, 0, (size_t)n); return 0; }
  p.winner = win_it * 3 + win_k; p.models = h->models;
  launch_k(k_fmcv_mask, dim3(cdiv(n, 256)), dim3(256), 0, h->stream, p);
  h->launches += 1;
  uint8_t* pm = h->h_pin;
  const size_t off = ((size_t)h->max_pts + 15) & ~(size_t)15;
  double* hF = reinterpret_cast<double*>(pm + off);
  B2S_CUDA(cudaMemcpyAsync(pm, h->d_mask, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  B2S_CUDA(cudaMemcpyAsync(hF, h->d_F, 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  B2S_CUDA(cudaStreamSynchronize(h->stream));
  B2S_LAUNCH_CHECK();
  std::memcpy(mask, pm, (size_t)n);
  if (F) std::memcpy(F, hF, 9 * sizeof(double));
  if (n_inliers) *n_inliers = max_good;
  return 0;
}

// test hook: the de-normalised candidate models of hypothesis `hyp` of the most recent call (host copy)
extern "C" int b2s_fm_debug_models(b2s_fm* h, int hyp, double* models27, int32_t* n_models, int32_t* counts3) {
  if (!h || hyp < 0 || hyp >= h->max_hyp) { set_error("b2s_fm_debug_models: bad hypothesis index"); return B2S_EINVAL; }
  B2S_CUDA(cudaSetDevice(h->device));
  B2S_CUDA(cudaDeviceSynchronize());
  if (models27) B2S_CUDA(cudaMemcpy(models27, h->models + (size_t)hyp * 27, 27 * sizeof(double), cudaMemcpyDeviceToHost));
  if (n_models) B2S_CUDA(cudaMemcpy(n_models, h->nmodels + hyp, sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (counts3) B2S_CUDA(cudaMemcpy(counts3, h->counts + (size_t)hyp * 3, 3 * sizeof(int32_t), cudaMemcpyDeviceToHost));
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// remap
// ---------------------------------------------------------------------------------------------------------------
struct b2s_remap {
  int device = 0, sH = 0, sW = 0, dH = 0, dW = 0;
  DeviceArena arena;
  float *mapx = nullptr, *mapy = nullptr; int16_t* wtab = nullptr;
  uint8_t *src = nullptr, *dst = nullptr;     // staging for the host API / output frame (dstride = 3*dW)
  cudaStream_t stream = nullptr;
  long long launches = 0;
};

namespace {
// OpenCV's initInterTab2D(INTER_LINEAR, fixpt): 32x32 sub-pixel positions, four float weights each, rounded to 15-bit
// integers whose sum is forced to 32768 by adjusting the largest (sum too big) or smallest (too small) entry.
void build_weight_table(std::vector<int16_t>& t) {
  t.resize(1024 * 4);
  for (int i = 0; i < 32; ++i)
    for (int j = 0; j < 32; ++j) {
      const float vy[2] = {1.f - (float)i * (1.f / 32.f), (float)i * (1.f / 32.f)};
      const float vx[2] = {1.f - (float)j * (1.f / 32.f), (float)j * (1.f / 32.f)};
      int it[4], sum = 0;
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          const float v = vy[a] * vx[b];
          long r = std::lrintf(v * 32768.f);
          if (r > 32767) r = 32767;
          if (r < -32768) r = -32768;
          it[a * 2 + b] = (int)r; sum += (int)r;
        }
      if (sum != 32768) {
        const int d = 32768 - sum;
        int k = 0;
        for (int q = 1; q < 4; ++q)
          if (d < 0 ? it[q] > it[k] : it[q] < it[k]) k = q;
        it[k] += d;
      }
      for (int q = 0; q < 4; ++q) t[(i * 32 + j) * 4 + q] = (int16_t)it[q];
    }
}
}  // namespace

extern "C" int b2s_remap_create(int device, const float* mapx_host, const float* mapy_host, int dst_h, int dst_w, int src_h,
                                int src_w, b2s_remap** out) {
  if (!out || !mapx_host || !mapy_host || dst_h < 1 || dst_w < 1 || src_h < 1 || src_w < 1) { set_error("b2s_remap_create: bad arguments"); return B2S_EINVAL; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    set_error("b2s_remap_create: no CUDA device %d (there is no CPU fallback)", device);
    return B2S_ENODEV;
  }
  B2S_CUDA(cudaSetDevice(device));
  b2s_remap* h = new b2s_remap();
  h->device = device; h->sH = src_h; h->sW = src_w; h->dH = dst_h; h->dW = dst_w;
  std::vector<int16_t> tab;
  build_weight_table(tab);
  int rc = 0;
  auto A = [&](int r) { if (rc == 0) rc = r; };
  A(h->arena.alloc(&h->mapx, (size_t)dst_h * dst_w));
  A(h->arena.alloc(&h->mapy, (size_t)dst_h * dst_w));
  A(h->arena.upload(&h->wtab, tab));
  A(h->arena.alloc(&h->src, (size_t)src_h * src_w * 3));
  A(h->arena.alloc(&h->dst, (size_t)dst_h * dst_w * 3));
  if (rc == 0 && cudaMemcpy(h->mapx, mapx_host, (size_t)dst_h * dst_w * 4, cudaMemcpyHostToDevice) != cudaSuccess) rc = B2S_ECUDA;
  if (rc == 0 && cudaMemcpy(h->mapy, mapy_host, (size_t)dst_h * dst_w * 4, cudaMemcpyHostToDevice) != cudaSuccess) rc = B2S_ECUDA;
  if (rc == 0 && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) rc = B2S_ECUDA;
  if (rc != 0) { if (rc == B2S_ECUDA) set_error("b2s_remap_create: %s", cudaGetErrorString(cudaGetLastError())); delete h; return rc; }
  *out = h;
  return 0;
}

extern "C" void b2s_remap_destroy(b2s_remap* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

extern "C" void b2s_remap_dims(const b2s_remap* h, int* dst_h, int* dst_w, int* src_h, int* src_w) {
  if (dst_h) *dst_h = h ? h->dH : 0;
  if (dst_w) *dst_w = h ? h->dW : 0;
  if (src_h) *src_h = h ? h->sH : 0;
  if (src_w) *src_w = h ? h->sW : 0;
}

extern "C" long long b2s_remap_launch_count(const b2s_remap* h) { return h ? h->launches : 0; }

extern "C" int b2s_remap_bgr(b2s_remap* h, const uint8_t* src_dev, int src_stride, void* stream, uint8_t* dst_dev, int dst_stride) {
  if (!h || !src_dev || !dst_dev) { set_error("b2s_remap_bgr: null argument"); return B2S_EINVAL; }
  if (src_stride < 3 * h->sW || dst_stride < 3 * h->dW) { set_error("b2s_remap_bgr: stride smaller than a row"); return B2S_EINVAL; }
  B2S_CUDA(cudaSetDevice(h->device));
  RemapParams p;
  p.src = src_dev; p.sH = h->sH; p.sW = h->sW; p.sstride = src_stride;
  p.mapx = h->mapx; p.mapy = h->mapy; p.dH = h->dH; p.dW = h->dW; p.wtab = h->wtab; p.dst = dst_dev; p.dstride = dst_stride;
  launch_k(k_remap_bgr_u8, dim3(cdiv(h->dW, 256), h->dH), dim3(256), 0, (cudaStream_t)stream, p);
  h->launches += 1;
  B2S_LAUNCH_CHECK();
  return 0;
}

extern "C" int b2s_remap_bgr_host(b2s_remap* h, const uint8_t* src, int src_stride, uint8_t* dst, int dst_stride) {
  if (!h || !src || !dst) { set_error("b2s_remap_bgr_host: null argument"); return B2S_EINVAL; }
  if (src_stride < 3 * h->sW || dst_stride < 3 * h->dW) { set_error("b2s_remap_bgr_host: stride smaller than a row"); return B2S_EINVAL; }
  B2S_CUDA(cudaSetDevice(h->device));
  B2S_CUDA(cudaMemcpy2DAsync(h->src, (size_t)3 * h->sW, src, (size_t)src_stride, (size_t)3 * h->sW, h->sH, cudaMemcpyHostToDevice, h->stream));
  B2S_TRY(b2s_remap_bgr(h, h->src, 3 * h->sW, h->stream, h->dst, 3 * h->dW));
  B2S_CUDA(cudaMemcpy2DAsync(dst, (size_t)dst_stride, h->dst, (size_t)3 * h->dW, (size_t)3 * h->dW, h->dH, cudaMemcpyDeviceToHost, h->stream));
  B2S_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

// the handle's own undistorted device frame ([dst_h][dst_w][3] u8, packed) - what b2s_aliked_extract consumes when the
// frame never has to come back to the host
extern "C" const uint8_t* b2s_remap_output_dev(const b2s_remap* h) { return h ? h->dst : nullptr; }

// ---------------------------------------------------------------------------------------------------------------
// reproject_and_match_2d3d
// ---------------------------------------------------------------------------------------------------------------
struct b2s_reproj {
  int device = 0, max_pts = 0, max_kps = 0, cap = 0;
  DeviceArena arena;
  int32_t *cand_kp = nullptr, *count = nullptr, *pos = nullptr, *flags = nullptr; float *cand_d = nullptr, *uv = nullptr;
  long long launches = 0;
};

// k_reproj_assign keeps one owner word per keypoint in shared memory (227 KB per CTA on sm_100)
static constexpr int REPROJ_MAX_KPS = 49152;

extern "C" int b2s_reproj_create(int device, int max_points, int max_kps, int cand_cap, b2s_reproj** out) {
  if (!out || max_points < 1 || max_kps < 1 || cand_cap < 1 || cand_cap > REPROJ_MAXCAP || max_kps > REPROJ_MAX_KPS) {
    set_error("b2s_reproj_create: bad arguments (cand_cap <= %d, max_kps <= %d)", REPROJ_MAXCAP, REPROJ_MAX_KPS);
    return B2S_EINVAL;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    set_error("b2s_reproj_create: no CUDA device %d (there is no CPU fallback)", device);
    return B2S_ENODEV;
  }
  B2S_CUDA(cudaSetDevice(device));
  b2s_reproj* h = new b2s_reproj();
  h->device = device; h->max_pts = max_points; h->max_kps = max_kps; h->cap = cand_cap;
  int rc = 0;
  auto A = [&](int r) { if (rc == 0) rc = r; };
  A(h->arena.alloc(&h->cand_kp, (size_t)max_points * cand_cap));
  A(h->arena.alloc(&h->cand_d, (size_t)max_points * cand_cap));
  A(h->arena.alloc(&h->count, (size_t)max_points));
  A(h->arena.alloc(&h->pos, (size_t)max_points));
  A(h->arena.alloc(&h->uv, (size_t)max_points * 2));
  A(h->arena.alloc(&h->flags, 1));
  if (rc == 0 && max_kps * sizeof(int) > 48 * 1024 &&
      cudaFuncSetAttribute(k_reproj_assign, cudaFuncAttributeMaxDynamicSharedMemorySize, max_kps * (int)sizeof(int)) != cudaSuccess) {
    set_error("b2s_reproj_create: %s", cudaGetErrorString(cudaGetLastError()));
    rc = B2S_ECUDA;
  }
  if (rc != 0) { delete h; return rc; }
  *out = h;
  return 0;
}

extern "C" void b2s_reproj_destroy(b2s_reproj* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  delete h;
}

extern "C" long long b2s_reproj_launch_count(const b2s_reproj* h) { return h ? h->launches : 0; }

extern "C" int b2s_reproj_match(b2s_reproj* h, const double* Xw_dev, const float* mp_desc_dev, const int32_t* mp_nobs_dev,
                                const int32_t* mp_row_dev, int P, int max_obs, const double* K_host, const double* Tcw_host, const float* kps_dev,
                                const float* des_dev, int N, int img_w, int img_h, double radius_px, double max_dist,
                                void* stream, int32_t* kp_of_point_dev, float* uv_dev, int32_t* flags_dev) {
  if (!h || !Xw_dev || !mp_desc_dev || !mp_nobs_dev || !K_host || !Tcw_host || !kps_dev || !des_dev || !kp_of_point_dev) {
    set_error("b2s_reproj_match: null argument");
    return B2S_EINVAL;
  }
  if (P < 1 || N < 1 || max_obs < 1) { set_error("b2s_reproj_match: bad sizes (P=%d N=%d max_obs=%d)", P, N, max_obs); return B2S_EINVAL; }
  if (P > h->max_pts || N > h->max_kps) { set_error("b2s_reproj_match: P=%d / N=%d exceed the handle (%d / %d)", P, N, h->max_pts, h->max_kps); return B2S_ESIZE; }
  B2S_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  ReprojParams p;
  p.Xw = Xw_dev; p.mp_desc = mp_desc_dev; p.mp_nobs = mp_nobs_dev; p.mp_row = mp_row_dev; p.P = P; p.max_obs = max_obs;
  for (int i = 0; i < 9; ++i) p.K[i] = K_host[i];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) p.R[3 * r + c] = Tcw_host[4 * r + c];
    p.t[r] = Tcw_host[4 * r + 3];
  }
  p.kps = kps_dev; p.des = des_dev; p.N = N; p.cap = h->cap;
  p.img_w = (float)img_w; p.img_h = (float)img_h;
  p.radius2 = radius_px * radius_px; p.thr = max_dist;
  p.cand_kp = h->cand_kp; p.cand_d = h->cand_d; p.count = h->count; p.pos = h->pos;
  p.uv = uv_dev ? uv_dev : h->uv; p.out_kp = kp_of_point_dev; p.flags = flags_dev ? flags_dev : h->flags;
  B2S_CUDA(cudaMemsetAsync(p.flags, 0, sizeof(int32_t), st));
  launch_k(k_reproj_candidates, dim3(cdiv(P, REPROJ_WARPS)), dim3(32 * REPROJ_WARPS), 0, st, p);
  launch_k(k_reproj_assign, dim3(1), dim3(1024), (size_t)N * sizeof(int), st, p);
  h->launches += 2;
  B2S_LAUNCH_CHECK();
  return 0;
}
