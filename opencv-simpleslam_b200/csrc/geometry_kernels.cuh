// geometry_kernels.cuh - the stages that sit directly behind the matcher in SimpleSLAM's tracking loop:
//   * fundamental-matrix RANSAC over a match list (`filter_matches_ransac`, features_utils.py:185-200, which calls
//     cv2.findFundamentalMat(pts1, pts2, FM_RANSAC, thresh, 0.99)): all hypotheses are generated and scored in
//     parallel (one thread per minimal 7-point sample, one CTA per hypothesis for the consensus count);
//   * frame ingest: cv2.remap(img, mapx, mapy, INTER_LINEAR) on u8 BGR frames (main_revamped.py:323-324), bit-exact
//     with OpenCV's fixed-point bilinear remap;
//   * projected-landmark window matching (`reproject_and_match_2d3d`, pnp_utils.py:224-304): projection, radius gate
//     and descriptor distances for every (map point, keypoint) candidate.
// All arithmetic that decides an index set is fp64 (RANSAC, projection) or integer (remap).
#pragma once
#include "common.cuh"

namespace b2s {

// ---------------------------------------------------------------------------------------------------------------
// RANSAC
// ---------------------------------------------------------------------------------------------------------------
// splitmix64: counter-based generator, one independent stream per hypothesis (state = seed ^ golden*(h+1))
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t& s) {
  s += 0x9E3779B97F4A7C15ull;
  uint64_t z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__host__ __device__ __forceinline__ double det3(const double* m) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

// real roots of c3 x^3 + c2 x^2 + c1 x + c0 (Numerical-Recipes form, then two Newton steps on the original polynomial)
__host__ __device__ inline int solve_cubic(double c3, double c2, double c1, double c0, double* x) {
  const double big = fmax(fmax(fabs(c3), fabs(c2)), fmax(fabs(c1), fabs(c0)));
  int n = 0;
  if (big == 0.0) return 0;
  if (fabs(c3) <= 1e-12 * big) {
    if (fabs(c2) <= 1e-12 * big) {
      if (fabs(c1) <= 1e-12 * big) return 0;
      x[0] = -c0 / c1;
      return 1;
    }
    const double disc = c1 * c1 - 4.0 * c2 * c0;
    if (disc < 0.0) return 0;
    const double sq = sqrt(disc);
    const double q = -0.5 * (c1 + (c1 >= 0.0 ? sq : -sq));
    x[n++] = q / c2;
    if (q != 0.0) x[n++] = c0 / q;
    return n;
  }
  const double a = c2 / c3, b = c1 / c3, c = c0 / c3;
  const double Q = (a * a - 3.0 * b) / 9.0, R = (2.0 * a * a * a - 9.0 * a * b + 27.0 * c) / 54.0;
  const double Q3 = Q * Q * Q;
  if (R * R < Q3) {
    const double sq = sqrt(Q);
    double ct = R / (sq * sq * sq);
    ct = fmin(1.0, fmax(-1.0, ct));
    const double th = acos(ct);
    x[0] = -2.0 * sq * cos(th / 3.0) - a / 3.0;
    x[1] = -2.0 * sq * cos((th + 6.283185307179586476925287) / 3.0) - a / 3.0;
    x[2] = -2.0 * sq * cos((th - 6.283185307179586476925287) / 3.0) - a / 3.0;
    n = 3;
  } else {
    const double A = -(R >= 0.0 ? 1.0 : -1.0) * cbrt(fabs(R) + sqrt(R * R - Q3));
    const double B = A != 0.0 ? Q / A : 0.0;
    x[0] = (A + B) - a / 3.0;
    n = 1;
  }
  for (int i = 0; i < n; ++i)
    for (int it = 0; it < 2; ++it) {
      const double v = x[i];
      const double f = ((c3 * v + c2) * v + c1) * v + c0, d = (3.0 * c3 * v + 2.0 * c2) * v + c1;
      if (d != 0.0) x[i] = v - f / d;
    }
  return n;
}

// 7-point algorithm on Hartley-normalised correspondences q[i] = (x1, y1, x2, y2): every F with x2^T F x1 = 0 for
// the seven pairs and det F = 0.  Writes up to three row-major 3x3 matrices (normalised coordinates) into Fout and
// returns their number (0 when the sample is degenerate).
__host__ __device__ inline int fm7_solve(const double (*q)[4], double* Fout) {
  double A[7][9];
  for (int i = 0; i < 7; ++i) {
    const double x1 = q[i][0], y1 = q[i][1], x2 = q[i][2], y2 = q[i][3];
    A[i][0] = x2 * x1; A[i][1] = x2 * y1; A[i][2] = x2;
    A[i][3] = y2 * x1; A[i][4] = y2 * y1; A[i][5] = y2;
    A[i][6] = x1;      A[i][7] = y1;      A[i][8] = 1.0;
  }
  int perm[9];
  for (int j = 0; j < 9; ++j) perm[j] = j;
  // Gauss-Jordan with complete pivoting -> [I7 | B] in permuted column order
  for (int r = 0; r < 7; ++r) {
    int pi = r, pj = r;
    double best = -1.0;
    for (int i = r; i < 7; ++i)
      for (int j = r; j < 9; ++j) {
        const double v = fabs(A[i][j]);
        if (v > best) { best = v; pi = i; pj = j; }
      }
    if (!(best > 1e-10)) return 0;
    if (pi != r)
      for (int j = 0; j < 9; ++j) { const double t = A[r][j]; A[r][j] = A[pi][j]; A[pi][j] = t; }
    if (pj != r) {
      for (int i = 0; i < 7; ++i) { const double t = A[i][r]; A[i][r] = A[i][pj]; A[i][pj] = t; }
      const int t = perm[r]; perm[r] = perm[pj]; perm[pj] = t;
    }
    const double inv = 1.0 / A[r][r];
    for (int j = r; j < 9; ++j) A[r][j] *= inv;
    for (int i = 0; i < 7; ++i) {
      if (i == r) continue;
      const double f = A[i][r];
      if (f == 0.0) continue;
      for (int j = r; j < 9; ++j) A[i][j] -= f * A[r][j];
    }
  }
  double f1[9], f2[9];
  for (int i = 0; i < 7; ++i) { f1[perm[i]] = -A[i][7]; f2[perm[i]] = -A[i][8]; }
  f1[perm[7]] = 1.0; f1[perm[8]] = 0.0;
  f2[perm[7]] = 0.0; f2[perm[8]] = 1.0;
  // F(l) = G + l D with G = f2, D = f1 - f2;  det F(l) = c0 + c1 l + c2 l^2 + c3 l^3
  double G[9], D[9], T[9];
  for (int j = 0; j < 9; ++j) { G[j] = f2[j]; D[j] = f1[j] - f2[j]; }
  const double c0 = det3(G), c3 = det3(D);
  double c1 = 0.0, c2 = 0.0;
  for (int r = 0; r < 3; ++r) {
    for (int j = 0; j < 9; ++j) T[j] = (j / 3 == r) ? D[j] : G[j];
    c1 += det3(T);
    for (int j = 0; j < 9; ++j) T[j] = (j / 3 == r) ? G[j] : D[j];
    c2 += det3(T);
  }
  double roots[3];
  const int nr = solve_cubic(c3, c2, c1, c0, roots);
  int nm = 0;
  for (int k = 0; k < nr; ++k) {
    const double l = roots[k];
    if (!(fabs(l) < 1e12)) continue;
    double nrm = 0.0;
    for (int j = 0; j < 9; ++j) { T[j] = G[j] + l * D[j]; nrm += T[j] * T[j]; }
    if (!(nrm > 0.0)) continue;
    for (int j = 0; j < 9; ++j) Fout[nm * 9 + j] = T[j];
    ++nm;
  }
  return nm;
}

// symmetric epipolar error of OpenCV's FM RANSAC: max(d(x2, F x1)^2, d(x1, F^T x2)^2)
__host__ __device__ __forceinline__ double fm_error(const double* F, double x1, double y1, double x2, double y2) {
  const double a = F[0] * x1 + F[1] * y1 + F[2], b = F[3] * x1 + F[4] * y1 + F[5], c = F[6] * x1 + F[7] * y1 + F[8];
  const double d2 = x2 * a + y2 * b + c;
  const double s2 = 1.0 / (a * a + b * b);
  const double a1 = F[0] * x2 + F[3] * y2 + F[6], b1 = F[1] * x2 + F[4] * y2 + F[7], c1 = F[2] * x2 + F[5] * y2 + F[8];
  const double d1 = x1 * a1 + y1 * b1 + c1;
  const double s1 = 1.0 / (a1 * a1 + b1 * b1);
  return fmax(d1 * d1 * s1, d2 * d2 * s2);
}

#if defined(__CUDACC__)
struct FmParams {
  const float* pts1; const float* pts2;   // [*,2] pixel coordinates
  const int32_t* pairs;                   // nullable [n,2]: row i uses pts1[pairs[i][0]], pts2[pairs[i][1]]
  int n, n_hyp;
  uint64_t seed;
  double thresh2;
  float4* xy;        // [n] gathered (x1, y1, x2, y2)
  double* norm;      // [6]  (cx1, cy1, s1, cx2, cy2, s2)
  double* models;    // [n_hyp][3][9] pixel-coordinate F, unit Frobenius norm
  int32_t* nmodels;  // [n_hyp]
  int32_t* counts;   // [n_hyp*3]
  uint8_t* mask;     // [n]
  double* F;         // [9]
  int32_t* result;   // [2]: inlier count, flat index (hypothesis*3 + root) of the winning model (-1: none)
};

// gather the correspondences and compute the two Hartley normalisations (centroid, sqrt2 / mean distance). One CTA.
__global__ void __launch_bounds__(256) k_fm_prepare(FmParams p) {
  pdl_wait();
  __shared__ double red[256][4];
  const int tid = threadIdx.x;
  double s[4] = {0, 0, 0, 0};
  for (int i = tid; i < p.n; i += 256) {
    const int i1 = p.pairs ? p.pairs[2 * i] : i, i2 = p.pairs ? p.pairs[2 * i + 1] : i;
    const float4 v = make_float4(p.pts1[2 * i1], p.pts1[2 * i1 + 1], p.pts2[2 * i2], p.pts2[2 * i2 + 1]);
    p.xy[i] = v;
    s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
  }
  for (int k = 0; k < 4; ++k) red[tid][k] = s[k];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) for (int k = 0; k < 4; ++k) red[tid][k] += red[tid + o][k];
    __syncthreads();
  }
  const double inv_n = 1.0 / (double)p.n;
  const double c[4] = {red[0][0] * inv_n, red[0][1] * inv_n, red[0][2] * inv_n, red[0][3] * inv_n};
  __syncthreads();
  double d1 = 0.0, d2 = 0.0;
  for (int i = tid; i < p.n; i += 256) {
    const float4 v = p.xy[i];
    const double ax = v.x - c[0], ay = v.y - c[1], bx = v.z - c[2], by = v.w - c[3];
    d1 += sqrt(ax * ax + ay * ay); d2 += sqrt(bx * bx + by * by);
  }
  red[tid][0] = d1; red[tid][1] = d2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) { red[tid][0] += red[tid + o][0]; red[tid][1] += red[tid + o][1]; }
    __syncthreads();
  }
  if (tid == 0) {
    const double m1 = red[0][0] * inv_n, m2 = red[0][1] * inv_n;
    p.norm[0] = c[0]; p.norm[1] = c[1]; p.norm[2] = m1 > 1e-12 ? 1.4142135623730951 / m1 : 1.0;
    p.norm[3] = c[2]; p.norm[4] = c[3]; p.norm[5] = m2 > 1e-12 ? 1.4142135623730951 / m2 : 1.0;
  }
}

// one thread per hypothesis: draw 7 distinct correspondences, solve, de-normalise (F = T2^T F^ T1), scale to |F|_F = 1
__global__ void __launch_bounds__(64) k_fm_hypotheses(FmParams p) {
  pdl_wait();
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= p.n_hyp) return;
  uint64_t st = p.seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(h + 1));
  int idx[7];
  for (int k = 0; k < 7; ++k) {
    for (;;) {
      const int c = (int)(splitmix64(st) % (uint64_t)p.n);
      bool dup = false;
      for (int j = 0; j < k; ++j) dup |= idx[j] == c;
      if (!dup) { idx[k] = c; break; }
    }
  }
  const double cx1 = p.norm[0], cy1 = p.norm[1], s1 = p.norm[2], cx2 = p.norm[3], cy2 = p.norm[4], s2 = p.norm[5];
  double q[7][4];
  for (int k = 0; k < 7; ++k) {
    const float4 v = p.xy[idx[k]];
    q[k][0] = (v.x - cx1) * s1; q[k][1] = (v.y - cy1) * s1; q[k][2] = (v.z - cx2) * s2; q[k][3] = (v.w - cy2) * s2;
  }
  double Fn[27];
  const int nm = fm7_solve(q, Fn);
  int kept = 0;
  for (int m = 0; m < nm; ++m) {
    const double* f = Fn + 9 * m;
    // T = [[s,0,-s cx],[0,s,-s cy],[0,0,1]]:  M = F^ T1, then F = T2^T M
    double M[9], F[9];
    for (int r = 0; r < 3; ++r) {
      M[3 * r + 0] = f[3 * r + 0] * s1;
      M[3 * r + 1] = f[3 * r + 1] * s1;
      M[3 * r + 2] = f[3 * r + 2] - s1 * (f[3 * r + 0] * cx1 + f[3 * r + 1] * cy1);
    }
    for (int c = 0; c < 3; ++c) {
      F[0 + c] = s2 * M[0 + c];
      F[3 + c] = s2 * M[3 + c];
      F[6 + c] = M[6 + c] - s2 * (cx2 * M[0 + c] + cy2 * M[3 + c]);
    }
    double nrm = 0.0;
    for (int j = 0; j < 9; ++j) nrm += F[j] * F[j];
    if (!(nrm > 0.0) || !(nrm < 1e300)) continue;
    const double inv = 1.0 / sqrt(nrm);
    double* out = p.models + ((size_t)h * 3 + kept) * 9;
    for (int j = 0; j < 9; ++j) out[j] = F[j] * inv;
    ++kept;
  }
  p.nmodels[h] = kept;
}

// one CTA per hypothesis: consensus count of each of its (<= 3) models over all correspondences
__global__ void __launch_bounds__(128) k_fm_score(FmParams p) {
  pdl_wait();
  const int h = blockIdx.x, tid = threadIdx.x;
  const int nm = p.nmodels[h];
  __shared__ double Fs[27];
  __shared__ int wsum[4][3];
  if (tid < 9 * nm) Fs[tid] = p.models[(size_t)h * 27 + tid];
  __syncthreads();
  int cnt[3] = {0, 0, 0};
  if (nm > 0) {
    for (int i = tid; i < p.n; i += 128) {
      const float4 v = p.xy[i];
      for (int m = 0; m < nm; ++m) cnt[m] += fm_error(Fs + 9 * m, v.x, v.y, v.z, v.w) <= p.thresh2 ? 1 : 0;
    }
  }
#pragma unroll
  for (int m = 0; m < 3; ++m) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt[m] += __shfl_xor_sync(0xffffffffu, cnt[m], o);
    if ((tid & 31) == 0) wsum[tid >> 5][m] = cnt[m];
  }
  __syncthreads();
  if (tid < 3) p.counts[h * 3 + tid] = tid < nm ? wsum[0][tid] + wsum[1][tid] + wsum[2][tid] + wsum[3][tid] : -1;
}

// winner = largest consensus, lowest flat index on ties; then its inlier mask. One CTA.
__global__ void __launch_bounds__(256) k_fm_select(FmParams p) {
  pdl_wait();
  __shared__ int bc[256], bi[256];
  __shared__ double Fs[9];
  const int tid = threadIdx.x, tot = p.n_hyp * 3;
  int best = -1, besti = 0x7fffffff;
  for (int i = tid; i < tot; i += 256) {
    const int c = p.counts[i];
    if (c > best) { best = c; besti = i; }       // ascending i within a thread: first maximum kept
  }
  bc[tid] = best; bi[tid] = besti;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) {
      const int c = bc[tid + o], i = bi[tid + o];
      if (c > bc[tid] || (c == bc[tid] && i < bi[tid])) { bc[tid] = c; bi[tid] = i; }
    }
    __syncthreads();
  }
  const int wc = bc[0], wi = bi[0];
  if (wc < 7) {     // no model at all (every sample degenerate)
    for (int i = tid; i < p.n; i += 256) p.mask[i] = 0;
    if (tid < 9) p.F[tid] = 0.0;
    if (tid == 0) { p.result[0] = 0; p.result[1] = -1; }
    return;
  }
  if (tid < 9) Fs[tid] = p.models[(size_t)wi * 9 + tid];
  __syncthreads();
  for (int i = tid; i < p.n; i += 256) {
    const float4 v = p.xy[i];
    p.mask[i] = fm_error(Fs, v.x, v.y, v.z, v.w) <= p.thresh2 ? 1 : 0;
  }
  if (tid < 9) p.F[tid] = Fs[tid];
  if (tid == 0) { p.result[0] = wc; p.result[1] = wi; }
}

// ---------------------------------------------------------------------------------------------------------------
// cv2-identical RANSAC: OpenCV's own sequential loop (ptsetreg.cpp RANSACPointSetRegistrator::run), parallelised.
// The sample stream of cv::RNG((uint64)-1) does not depend on the models, so the host draws every iteration's subset up
// front (geometry.cu: cv RNG + getSubset + haveCollinearPoints), the device solves and scores all of them at once, and
// the host replays the loop's only sequential part (strictly-greater update, RANSACUpdateNumIters) over the counts.
//   * k_fmcv_models: fundam.cpp run7Point per subset, fp64.  The two null-space rows of Vt that SVDecomp(FULL_UV) hands
//     to run7Point are NOT arbitrary: JacobiSVDImpl_ completes the basis from fixed +-1/9 sign vectors (RNG(0x12345678))
//     projected off the row space, so they depend on the row space only (oracle/cv_ransac.py null_space_basis, checked
//     against cv2.SVDecomp) - which also fixes solveCubic's root order, i.e. the order in which cv2 tries the models.
//   * k_fmcv_count: FMEstimatorCallback::computeError as float32 <= (float)(thresh*thresh), same operation order.
// ---------------------------------------------------------------------------------------------------------------
// cv::solveCubic for c[0] x^3 + c[1] x^2 + c[2] x + c[3], OpenCV's root order
__host__ __device__ inline int cv_solve_cubic(const double* c, double* x) {
  double a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3];
  int n = 0;
  x[0] = x[1] = x[2] = 0.0;
  if (a0 == 0.0) {
    if (a1 == 0.0) {
      if (a2 == 0.0) n = a3 == 0.0 ? -1 : 0;
      else { x[0] = -a3 / a2; n = 1; }
    } else {
      double d = a2 * a2 - 4 * a1 * a3;
      if (d >= 0) {
        d = sqrt(d);
        const double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
        if (fabs(q1) > fabs(q2)) { x[0] = q1 / a1; x[1] = a3 / q1; }
        else { x[0] = q2 / a1; x[1] = a3 / q2; }
        n = d > 0 ? 2 : 1;
      }
    }
  } else {
    a0 = 1. / a0; a1 *= a0; a2 *= a0; a3 *= a0;
    const double Q = (a1 * a1 - 3 * a2) * (1. / 9);
    const double R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
    const double Qcubed = Q * Q * Q;
    double d = (a1 * a1 * (a2 * a2 - 4 * a1 * a3) + 2 * a2 * (9 * a1 * a3 - 2 * a2 * a2) - 27 * a3 * a3) * (1. / 108);
    if (d > 0) {
      const double theta = acos(R / sqrt(Qcubed));
      const double t0 = -2 * sqrt(Q), t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
      x[0] = t0 * cos(t1) - t2;
      x[1] = t0 * cos(t1 + (2. * 3.1415926535897932384626433832795 / 3)) - t2;
      x[2] = t0 * cos(t1 + (4. * 3.1415926535897932384626433832795 / 3)) - t2;
      n = 3;
    } else if (d == 0) {
      if (R >= 0) { x[0] = -2 * pow(R, 1. / 3) - a1 / 3; x[1] = pow(R, 1. / 3) - a1 / 3; }
      else { x[0] = 2 * pow(-R, 1. / 3) - a1 / 3; x[1] = -pow(-R, 1. / 3) - a1 / 3; }
      n = x[0] == x[1] ? 1 : 2;
      x[1] = x[0] == x[1] ? 0 : x[1];
    } else {
      d = sqrt(-d);
      double e = pow(d + fabs(R), 1. / 3);
      if (R > 0) e = -e;
      x[0] = (e + Q / e) - a1 * (1. / 3);
      n = 1;
    }
  }
  return n;
}

// v -= (q . v) q for the nq orthonormal rows of Q, two passes; returns |v|
__host__ __device__ inline double cv_project_off(double* v, const double (*Q)[9], int nq, const double* extra) {
  for (int pass = 0; pass < 2; ++pass) {
    for (int j = 0; j < nq + (extra ? 1 : 0); ++j) {
      const double* q = j < nq ? Q[j] : extra;
      double d = 0.0;
      for (int k = 0; k < 9; ++k) d += q[k] * v[k];
      for (int k = 0; k < 9; ++k) v[k] -= d * q[k];
    }
  }
  double nn = 0.0;
  for (int k = 0; k < 9; ++k) nn += v[k] * v[k];
  return sqrt(nn);
}

// fundam.cpp run7Point on seven float32 correspondences; signs: bit k of sign1 / sign2 set = +1/9 in component k of the
// completion vector of Vt row 7 / 8.  Writes up to three F (row-major, F[8] = 1 when |F[8]| > FLT_EPSILON), returns n.
__host__ __device__ inline int cv_run7point(const float (*m1)[2], const float (*m2)[2], unsigned sign1, unsigned sign2, double* Fout) {
  double c1x = 0, c1y = 0, c2x = 0, c2y = 0;
  for (int i = 0; i < 7; ++i) { c1x += (double)m1[i][0]; c1y += (double)m1[i][1]; c2x += (double)m2[i][0]; c2y += (double)m2[i][1]; }
  const double t = 1. / 7;
  c1x *= t; c1y *= t; c2x *= t; c2y *= t;
  double scale1 = 0, scale2 = 0;
  for (int i = 0; i < 7; ++i) {
    const double ax = m1[i][0] - c1x, ay = m1[i][1] - c1y, bx = m2[i][0] - c2x, by = m2[i][1] - c2y;
    scale1 += sqrt(ax * ax + ay * ay);
    scale2 += sqrt(bx * bx + by * by);
  }
  scale1 *= t; scale2 *= t;
  if (scale1 < 1.1920928955078125e-7 || scale2 < 1.1920928955078125e-7) return 0;
  scale1 = sqrt(2.) / scale1;
  scale2 = sqrt(2.) / scale2;
  double Q[7][9];
  int nq = 0;
  for (int i = 0; i < 7; ++i) {
    const double x0 = (m1[i][0] - c1x) * scale1, y0 = (m1[i][1] - c1y) * scale1;
    const double x1 = (m2[i][0] - c2x) * scale2, y1 = (m2[i][1] - c2y) * scale2;
    double* a = Q[nq];
    a[0] = x1 * x0; a[1] = x1 * y0; a[2] = x1; a[3] = y1 * x0; a[4] = y1 * y0; a[5] = y1; a[6] = x0; a[7] = y0; a[8] = 1;
    const double nv = cv_project_off(a, Q, nq, nullptr);
    if (nv > 0) {
      const double inv = 1. / nv;
      for (int k = 0; k < 9; ++k) a[k] *= inv;
      ++nq;
    }
  }
  double f1[9], f2[9];
  for (int k = 0; k < 9; ++k) { f1[k] = (sign1 >> k) & 1u ? 1. / 9 : -1. / 9; f2[k] = (sign2 >> k) & 1u ? 1. / 9 : -1. / 9; }
  double nv = cv_project_off(f1, Q, nq, nullptr);
  for (int k = 0; k < 9; ++k) f1[k] /= nv;
  nv = cv_project_off(f2, Q, nq, f1);
  for (int k = 0; k < 9; ++k) f2[k] /= nv;
  for (int k = 0; k < 9; ++k) f1[k] -= f2[k];
  double c[4], r[3];
  double t0 = f2[4] * f2[8] - f2[5] * f2[7], t1 = f2[3] * f2[8] - f2[5] * f2[6], t2 = f2[3] * f2[7] - f2[4] * f2[6];
  c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
  c[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) + f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) -
         f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) + f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
         f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
  t0 = f1[4] * f1[8] - f1[5] * f1[7]; t1 = f1[3] * f1[8] - f1[5] * f1[6]; t2 = f1[3] * f1[7] - f1[4] * f1[6];
  c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
  c[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) + f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) -
         f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) + f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
         f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
  const int n = cv_solve_cubic(c, r);
  if (n < 1 || n > 3) return 0;
  for (int k = 0; k < n; ++k) {
    double lambda = r[k], mu = 1.;
    const double s = f1[8] * r[k] + f2[8];
    double fm[9];
    if (fabs(s) > 2.220446049250313e-16) { mu = 1. / s; lambda *= mu; fm[8] = 1.; }
    else fm[8] = 0.;
    for (int i = 0; i < 8; ++i) fm[i] = f1[i] * lambda + f2[i] * mu;
    // F = T2^T fm T1,  T = [[s,0,-s cx],[0,s,-s cy],[0,0,1]]
    double M[9], F[9];
    for (int rr = 0; rr < 3; ++rr) {
      M[3 * rr + 0] = fm[3 * rr + 0] * scale1;
      M[3 * rr + 1] = fm[3 * rr + 1] * scale1;
      M[3 * rr + 2] = fm[3 * rr + 0] * (-scale1 * c1x) + fm[3 * rr + 1] * (-scale1 * c1y) + fm[3 * rr + 2];
    }
    for (int cc = 0; cc < 3; ++cc) {
      F[0 + cc] = scale2 * M[0 + cc];
      F[3 + cc] = scale2 * M[3 + cc];
      F[6 + cc] = (-scale2 * c2x) * M[0 + cc] + (-scale2 * c2y) * M[3 + cc] + M[6 + cc];
    }
    if (fabs(F[8]) > 1.1920928955078125e-7) {
      const double inv = 1. / F[8];
      for (int i = 0; i < 9; ++i) F[i] *= inv;
    }
    for (int i = 0; i < 9; ++i) Fout[9 * k + i] = F[i];
  }
  return n;
}

// FMEstimatorCallback::computeError (float32 result): the source's operation order with every product and sum rounded
// separately (OpenCV's baseline x86-64 build has no fused multiply-add), so the float32 errors compared with the
// threshold are the ones cv2 computes for the same F
__device__ __forceinline__ double cv_dot3(double a, double x, double b, double y, double c) {
  return __dadd_rn(__dadd_rn(__dmul_rn(a, x), __dmul_rn(b, y)), c);
}
__device__ __forceinline__ float cv_fm_error(const double* F, float x1f, float y1f, float x2f, float y2f) {
  const double x1 = x1f, y1 = y1f, x2 = x2f, y2 = y2f;
  double a = cv_dot3(F[0], x1, F[1], y1, F[2]);
  double b = cv_dot3(F[3], x1, F[4], y1, F[5]);
  double c = cv_dot3(F[6], x1, F[7], y1, F[8]);
  const double s2 = __ddiv_rn(1., __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)));
  const double d2 = cv_dot3(x2, a, y2, b, c);
  a = cv_dot3(F[0], x2, F[3], y2, F[6]);
  b = cv_dot3(F[1], x2, F[4], y2, F[7]);
  c = cv_dot3(F[2], x2, F[5], y2, F[8]);
  const double s1 = __ddiv_rn(1., __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)));
  const double d1 = cv_dot3(x1, a, y1, b, c);
  const double e1 = __dmul_rn(__dmul_rn(d1, d1), s1), e2 = __dmul_rn(__dmul_rn(d2, d2), s2);
  return (float)(e1 < e2 ? e2 : e1);   // std::max(e1, e2), NaN handling included
}

#if defined(__CUDACC__)
struct FmCvParams {
  const float* m1; const float* m2;      // [n][2] float32 pixel coordinates
  int n, n_sub;
  const int32_t* subsets;                // [n_sub][7]
  unsigned sign1, sign2;
  float thresh2;                         // (float)(thresh * thresh)
  double* models;                        // [n_sub][3][9]
  int32_t* nmodels;                      // [n_sub]
  int32_t* counts;                       // [n_sub*3]  (-1: no such model)
  int winner;                            // k_fmcv_mask: flat model index
  uint8_t* mask; double* F;
};

__global__ void __launch_bounds__(32) k_fmcv_models(FmCvParams p) {
  pdl_wait();
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= p.n_sub) return;
  float m1[7][2], m2[7][2];
  for (int k = 0; k < 7; ++k) {
    const int i = p.subsets[h * 7 + k];
    m1[k][0] = p.m1[2 * i]; m1[k][1] = p.m1[2 * i + 1];
    m2[k][0] = p.m2[2 * i]; m2[k][1] = p.m2[2 * i + 1];
  }
  double Fm[27];
  const int nm = cv_run7point(m1, m2, p.sign1, p.sign2, Fm);
  for (int j = 0; j < 9 * nm; ++j) p.models[(size_t)h * 27 + j] = Fm[j];
  p.nmodels[h] = nm;
}

__global__ void __launch_bounds__(128) k_fmcv_count(FmCvParams p) {
  pdl_wait();
  const int h = blockIdx.x, tid = threadIdx.x;
  const int nm = p.nmodels[h];
  __shared__ double Fs[27];
  __shared__ int wsum[4][3];
  if (tid < 9 * nm) Fs[tid] = p.models[(size_t)h * 27 + tid];
  __syncthreads();
  int cnt[3] = {0, 0, 0};
  if (nm > 0) {
    for (int i = tid; i < p.n; i += 128) {
      const float2 a = reinterpret_cast<const float2*>(p.m1)[i], b = reinterpret_cast<const float2*>(p.m2)[i];
      for (int m = 0; m < nm; ++m) cnt[m] += cv_fm_error(Fs + 9 * m, a.x, a.y, b.x, b.y) <= p.thresh2 ? 1 : 0;
    }
  }
#pragma unroll
  for (int m = 0; m < 3; ++m) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt[m] += __shfl_xor_sync(0xffffffffu, cnt[m], o);
    if ((tid & 31) == 0) wsum[tid >> 5][m] = cnt[m];
  }
  __syncthreads();
  if (tid < 3) p.counts[h * 3 + tid] = tid < nm ? wsum[0][tid] + wsum[1][tid] + wsum[2][tid] + wsum[3][tid] : -1;
}

__global__ void __launch_bounds__(256) k_fmcv_mask(FmCvParams p) {
  pdl_wait();
  __shared__ double Fs[9];
  if (threadIdx.x < 9) Fs[threadIdx.x] = p.models[(size_t)p.winner * 9 + threadIdx.x];
  __syncthreads();
  for (int i = blockIdx.x * 256 + threadIdx.x; i < p.n; i += gridDim.x * 256) {
    const float2 a = reinterpret_cast<const float2*>(p.m1)[i], b = reinterpret_cast<const float2*>(p.m2)[i];
    p.mask[i] = cv_fm_error(Fs, a.x, a.y, b.x, b.y) <= p.thresh2 ? 1 : 0;
  }
  if (blockIdx.x == 0 && threadIdx.x < 9) p.F[threadIdx.x] = Fs[threadIdx.x];
}
#endif

// ---------------------------------------------------------------------------------------------------------------
// cv2.remap(src u8 C3, mapx f32, mapy f32, INTER_LINEAR, BORDER_CONSTANT 0): OpenCV converts the maps to fixed point
// with 5 fractional bits (saturate_cast<int>(v*32), i.e. round-half-even), looks the four bilinear weights up in a
// 32x32 table of 15-bit integers that sum to 32768, and rounds (sum + 16384) >> 15.  The table is passed in (built on
// the host exactly as OpenCV builds it), the kernel is integer arithmetic only.  One thread per destination pixel.
// ---------------------------------------------------------------------------------------------------------------
struct RemapParams {
  const uint8_t* src; int sH, sW, sstride;
  const float* mapx; const float* mapy; int dH, dW;
  const int16_t* wtab;     // [1024][4]
  uint8_t* dst; int dstride;
};

__global__ void __launch_bounds__(256) k_remap_bgr_u8(RemapParams p) {
  pdl_wait();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= p.dW) return;
  const float mx = p.mapx[(size_t)y * p.dW + x], my = p.mapy[(size_t)y * p.dW + x];
  // cvRound(v * 32): round to nearest even, saturated to int
  const float fx = mx * 32.f, fy = my * 32.f;
  const int sx = fx >= 2147483520.f ? 0x7fffffff : (fx <= -2147483648.f ? (int)0x80000000 : __float2int_rn(fx));
  const int sy = fy >= 2147483520.f ? 0x7fffffff : (fy <= -2147483648.f ? (int)0x80000000 : __float2int_rn(fy));
  // OpenCV stores the integer part as short (saturating) and the 10 fractional bits separately
  const int ix = max(-32768, min(32767, sx >> 5)), iy = max(-32768, min(32767, sy >> 5));
  const int16_t* w = p.wtab + (((sy & 31) << 5) | (sx & 31)) * 4;
  const int w00 = w[0], w01 = w[1], w10 = w[2], w11 = w[3];
  const bool x0 = (unsigned)ix < (unsigned)p.sW, x1 = (unsigned)(ix + 1) < (unsigned)p.sW;
  const bool y0 = (unsigned)iy < (unsigned)p.sH, y1 = (unsigned)(iy + 1) < (unsigned)p.sH;
  const uint8_t* r0 = p.src + (size_t)(y0 ? iy : 0) * p.sstride;
  const uint8_t* r1 = p.src + (size_t)(y1 ? iy + 1 : 0) * p.sstride;
  uint8_t* d = p.dst + (size_t)y * p.dstride + 3 * x;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int p00 = (y0 && x0) ? r0[3 * ix + c] : 0, p01 = (y0 && x1) ? r0[3 * (ix + 1) + c] : 0;
    const int p10 = (y1 && x0) ? r1[3 * ix + c] : 0, p11 = (y1 && x1) ? r1[3 * (ix + 1) + c] : 0;
    const int v = (p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11 + (1 << 14)) >> 15;
    d[c] = (uint8_t)max(0, min(255, v));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// reproject_and_match_2d3d (pnp_utils.py:224-304): for every map point, project with the predicted pose, collect the
// keypoints within radius_px of the projection, and score each by the smallest descriptor distance to the point's
// last (<= 6) observations.  The reference then walks the points in order and gives each its best not-yet-used
// keypoint if that distance is within the threshold; k_reproj_assign reproduces that order-dependent greedy result
// with a parallel fixpoint (a claim is lost only to a smaller point index, exactly the points processed earlier).
// ---------------------------------------------------------------------------------------------------------------
struct ReprojParams {
  const double* Xw;          // [P,3] world positions
  const float* mp_desc;      // [P,max_obs,128] descriptors of the most recent observations (oldest first)
  const int32_t* mp_nobs;    // [P] usable descriptors (0: the reference skips the point)
  const int32_t* mp_row;     // nullable [P]: descriptors / counts of point i live in table row mp_row[i] (else row i)
  int P, max_obs;
  double K[9], R[9], t[3];   // intrinsics, T_cw rotation / translation
  const float* kps;          // [N,2]
  const float* des;          // [N,128]
  int N, cap;
  float img_w, img_h;
  double radius2, thr;
  int32_t* cand_kp;          // [P,cap] keypoint indices, ascending distance
  float* cand_d;             // [P,cap]
  int32_t* count;            // [P]
  int32_t* pos;              // [P] scratch of the assignment fixpoint
  float* uv;                 // [P,2] projections (float32 like the reference's uv_all), (-1,-1) behind the camera
  int32_t* out_kp;           // [P] assigned keypoint or -1
  int32_t* flags;            // [0] != 0: a point whose candidate list was truncated to `cap` ran out of candidates (result unreliable)
};

// A point keeps at most `cap` candidates: only keypoints that pass the descriptor gate (distance <= thr) are stored - the
// others can never be chosen - and when more than `cap` pass, the `cap` nearest in descriptor distance survive.
constexpr int REPROJ_WARPS = 8, REPROJ_MAXCAP = 256, REPROJ_TRUNC = 1 << 30;

__global__ void __launch_bounds__(32 * REPROJ_WARPS) k_reproj_candidates(ReprojParams p) {
  pdl_wait();
  __shared__ int s_kp[REPROJ_WARPS][REPROJ_MAXCAP];
  __shared__ float s_d[REPROJ_WARPS][REPROJ_MAXCAP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pt = blockIdx.x * REPROJ_WARPS + warp;
  if (pt >= p.P) return;
  const double X = p.Xw[3 * pt], Y = p.Xw[3 * pt + 1], Z = p.Xw[3 * pt + 2];
  const double xc = X * p.R[0] + Y * p.R[1] + Z * p.R[2] + p.t[0];
  const double yc = X * p.R[3] + Y * p.R[4] + Z * p.R[5] + p.t[1];
  const double zc = X * p.R[6] + Y * p.R[7] + Z * p.R[8] + p.t[2];
  float u = -1.f, v = -1.f;
  if (zc > 1e-8) {
    const double xn = xc / zc, yn = yc / zc, zn = zc / zc;
    u = (float)(p.K[0] * xn + p.K[1] * yn + p.K[2] * zn);
    v = (float)(p.K[3] * xn + p.K[4] * yn + p.K[5] * zn);
  }
  if (lane == 0) { p.uv[2 * pt] = u; p.uv[2 * pt + 1] = v; }
  const int row = p.mp_row ? p.mp_row[pt] : pt;
  const int nobs = p.mp_nobs[row];
  const bool live = zc > 0.0 && u >= 0.f && u < p.img_w && v >= 0.f && v < p.img_h && nobs > 0;
  if (!live) { if (lane == 0) p.count[pt] = 0; return; }
  const float* mpd = p.mp_desc + (size_t)row * p.max_obs * 128;
  int cnt = 0;
  bool truncated = false;
  for (int base = 0; base < p.N; base += 32) {
    const int i = base + lane;
    bool in = false;
    if (i < p.N) {
      const double dx = (double)p.kps[2 * i] - (double)u, dy = (double)p.kps[2 * i + 1] - (double)v;
      in = dx * dx + dy * dy <= p.radius2;
    }
    unsigned m = __ballot_sync(0xffffffffu, in);
    while (m) {
      const int j = base + __ffs(m) - 1;
      m &= m - 1;
      const float4 b = reinterpret_cast<const float4*>(p.des + (size_t)j * 128)[lane];
      float best = INFINITY;
      for (int o = 0; o < nobs; ++o) {
        const float4 a = reinterpret_cast<const float4*>(mpd + (size_t)o * 128)[lane];
        const float e0 = a.x - b.x, e1 = a.y - b.y, e2 = a.z - b.z, e3 = a.w - b.w;
        const float d = sqrtf(warp_sum(fmaf(e3, e3, fmaf(e2, e2, fmaf(e1, e1, e0 * e0)))));
        best = fminf(best, d);
      }
      if (!((double)best <= p.thr)) continue;           // fails the descriptor gate: never assignable (uniform per warp)
      if (cnt < p.cap) {
        if (lane == 0) { s_kp[warp][cnt] = j; s_d[warp][cnt] = best; }
        ++cnt;
      } else {
        // list full: the new candidate replaces the current worst (largest distance, then largest index) if it is better
        truncated = true;
        __syncwarp();
        float wd = -1.f; int wq = -1, wk = -1;
        for (int q = lane; q < p.cap; q += 32) {
          const float dq = s_d[warp][q]; const int kq = s_kp[warp][q];
          if (dq > wd || (dq == wd && kq > wk)) { wd = dq; wq = q; wk = kq; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float od = __shfl_xor_sync(0xffffffffu, wd, o);
          const int oq = __shfl_xor_sync(0xffffffffu, wq, o), ok = __shfl_xor_sync(0xffffffffu, wk, o);
          if (od > wd || (od == wd && ok > wk)) { wd = od; wq = oq; wk = ok; }
        }
        if (lane == 0 && (best < wd || (best == wd && j < wk))) { s_kp[warp][wq] = j; s_d[warp][wq] = best; }
        __syncwarp();
      }
    }
  }
  __syncwarp();
  // rank sort by (distance, keypoint index)
  for (int e = lane; e < cnt; e += 32) {
    const float d = s_d[warp][e];
    const int ke = s_kp[warp][e];
    int rank = 0;
    for (int q = 0; q < cnt; ++q) {
      const float dq = s_d[warp][q];
      rank += (dq < d || (dq == d && s_kp[warp][q] < ke)) ? 1 : 0;
    }
    p.cand_kp[(size_t)pt * p.cap + rank] = ke;
    p.cand_d[(size_t)pt * p.cap + rank] = d;
  }
  if (lane == 0) p.count[pt] = cnt | (truncated ? REPROJ_TRUNC : 0);
}

// one CTA; dynamic smem: owner[N]
__global__ void __launch_bounds__(1024) k_reproj_assign(ReprojParams p) {
  pdl_wait();
  extern __shared__ int owner[];
  __shared__ int changed;
  const int tid = threadIdx.x;
  for (int q = tid; q < p.P; q += 1024) p.pos[q] = 0;
  for (;;) {   // every round but the last advances at least one point; total advances <= sum(count)
    for (int i = tid; i < p.N; i += 1024) owner[i] = 0x7fffffff;
    if (tid == 0) changed = 0;
    __syncthreads();
    for (int q = tid; q < p.P; q += 1024) {
      const int ps = p.pos[q];
      if (ps < (p.count[q] & (REPROJ_TRUNC - 1))) atomicMin(&owner[p.cand_kp[(size_t)q * p.cap + ps]], q);
    }
    __syncthreads();
    for (int q = tid; q < p.P; q += 1024) {
      const int ps = p.pos[q];
      if (ps < (p.count[q] & (REPROJ_TRUNC - 1)) && owner[p.cand_kp[(size_t)q * p.cap + ps]] != q) {
        p.pos[q] = ps + 1;
        changed = 1;
      }
    }
    __syncthreads();
    const int c = changed;
    __syncthreads();
    if (!c) break;
  }
  for (int q = tid; q < p.P; q += 1024) {
    const int ps = p.pos[q], cnt = p.count[q] & (REPROJ_TRUNC - 1);
    const bool ok = ps < cnt;                    // every stored candidate passed the descriptor gate
    p.out_kp[q] = ok ? p.cand_kp[(size_t)q * p.cap + ps] : -1;
    if (!ok && (p.count[q] & REPROJ_TRUNC)) atomicOr(p.flags, 1);   // lost all `cap` nearest candidates: the dropped ones might have matched
  }
}
#endif  // __CUDACC__

}  // namespace b2s
