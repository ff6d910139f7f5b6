// MAGSAC++ geometric verification on the GPU.
// Replaces the OpenCV calls made by imcui/ui/utils.py:352-372 (cv2.findHomography / cv2.findFundamentalMat with
// method=cv2.USAC_MAGSAC) for batches of correspondence sets.  OpenCV's source is not part of the reference tree, so
// this restates the published MAGSAC++ algorithm (Barath et al., CVPR 2020): minimal-sample hypotheses scored with the
// sigma-consensus++ loss (noise scale marginalised up to sigma_max = threshold / 3.64, 4 degrees of freedom), adaptive
// termination from the inlier ratio at `threshold`, iteratively re-weighted least-squares polishing with the MAGSAC++
// weights, final inlier mask = residual <= threshold.  Parity with cv2 is statistical (different sampling sequence):
// tests compare inlier-mask F1, inlier counts and the model error on cv2's inliers.
//
// One CTA per correspondence set: 256 threads = 256 independent hypotheses per round; every thread solves its own
// minimal problem in registers (fp64) and scores it against all correspondences, which sit in shared memory
// (broadcast reads).  No global traffic inside the loop.
#include <math_constants.h>

#include "../../include/imw_b200.h"
#include <cooperative_groups.h>

#include "common.cuh"

namespace {

constexpr int MS_THREADS = 256;
constexpr float MS_K = 3.64f;  // 0.99 quantile of the chi distribution with 4 degrees of freedom

__device__ __forceinline__ unsigned pcg_hash(unsigned v) {
  unsigned s = v * 747796405u + 2891336453u;
  unsigned w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
  return (w >> 22u) ^ w;
}

// upper incomplete gamma functions for half-integer orders (closed forms)
__device__ __forceinline__ float ugamma_1p5(float x) { return 0.88622692545f * erfcf(sqrtf(x)) + sqrtf(x) * __expf(-x); }
__device__ __forceinline__ float ugamma_2p5(float x) { return 1.5f * ugamma_1p5(x) + x * sqrtf(x) * __expf(-x); }

struct MagsacConsts {
  float sigma_max, thr2_max, two_sigma2, gamma_k, norm, loss_out, thr2_inl;
};
__device__ __forceinline__ MagsacConsts make_consts(float threshold) {
  MagsacConsts c;
  c.sigma_max = threshold / MS_K;
  c.thr2_max = threshold * threshold;               // (k sigma_max)^2
  c.two_sigma2 = 2.f * c.sigma_max * c.sigma_max;
  c.gamma_k = ugamma_1p5(MS_K * MS_K / 2.f);
  c.norm = 1.f / c.sigma_max;                       // C(n) 2^((n-1)/2) is a common factor: dropped
  c.thr2_inl = threshold * threshold;
  // loss at the truncation radius: sigma^2/2 * gamma_lower(2.5, k^2/2)
  const float x = MS_K * MS_K / 2.f;
  c.loss_out = c.norm * (c.sigma_max * c.sigma_max / 2.f) * (1.32934038818f - ugamma_2p5(x));
  return c;
}
// MAGSAC++ loss of a squared residual (eq. 11 of the paper, n = 4)
__device__ __forceinline__ float magsac_loss(float r2, const MagsacConsts& c) {
  if (r2 >= c.thr2_max) return c.loss_out;
  const float x = r2 / c.two_sigma2;
  return c.norm * ((c.sigma_max * c.sigma_max / 2.f) * (1.32934038818f - ugamma_2p5(x)) + (r2 / 4.f) * (ugamma_1p5(x) - c.gamma_k));
}
__device__ __forceinline__ float magsac_weight(float r2, const MagsacConsts& c) {
  if (r2 >= c.thr2_max) return 0.f;
  return c.norm * (ugamma_1p5(r2 / c.two_sigma2) - c.gamma_k);
}

// squared residuals in pixels
__device__ __forceinline__ float resid_h(const float* M, float x0, float y0, float x1, float y1) {
  float w = M[6] * x0 + M[7] * y0 + M[8];
  float iw = 1.f / w;
  float dx = (M[0] * x0 + M[1] * y0 + M[2]) * iw - x1, dy = (M[3] * x0 + M[4] * y0 + M[5]) * iw - y1;
  return dx * dx + dy * dy;
}
__device__ __forceinline__ float resid_f(const float* M, float x0, float y0, float x1, float y1) {  // Sampson distance^2
  float a = M[0] * x0 + M[1] * y0 + M[2], b = M[3] * x0 + M[4] * y0 + M[5], c = M[6] * x0 + M[7] * y0 + M[8];
  float at = M[0] * x1 + M[3] * y1 + M[6], bt = M[1] * x1 + M[4] * y1 + M[7];
  float e = x1 * a + y1 * b + c;
  return e * e / (a * a + b * b + at * at + bt * bt);
}

// ---- minimal solvers (fp64, one thread) -----------------------------------------------------------------------------
// homography from 4 correspondences: 8x8 system with h33 = 1, Gaussian elimination with partial pivoting
__device__ bool solve_h4(const double (*p)[4], double* H) {
  double A[8][9];
  for (int i = 0; i < 4; i++) {
    const double x = p[i][0], y = p[i][1], u = p[i][2], v = p[i][3];
    double* r0 = A[2 * i]; double* r1 = A[2 * i + 1];
    r0[0] = x; r0[1] = y; r0[2] = 1; r0[3] = 0; r0[4] = 0; r0[5] = 0; r0[6] = -u * x; r0[7] = -u * y; r0[8] = u;
    r1[0] = 0; r1[1] = 0; r1[2] = 0; r1[3] = x; r1[4] = y; r1[5] = 1; r1[6] = -v * x; r1[7] = -v * y; r1[8] = v;
  }
  for (int c = 0; c < 8; c++) {
    int piv = c; double best = fabs(A[c][c]);
    for (int r = c + 1; r < 8; r++) if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); piv = r; }
    if (best < 1e-12) return false;
    if (piv != c) for (int k = c; k < 9; k++) { double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
    const double inv = 1.0 / A[c][c];
    for (int r = c + 1; r < 8; r++) {
      const double f = A[r][c] * inv;
      for (int k = c; k < 9; k++) A[r][k] -= f * A[c][k];
    }
  }
  for (int c = 7; c >= 0; c--) {
    double s = A[c][8];
    for (int k = c + 1; k < 8; k++) s -= A[c][k] * H[k];
    H[c] = s / A[c][c];
  }
  H[8] = 1.0;
  return true;
}

__device__ __forceinline__ double det3(const double* M) {
  return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}
// fundamental matrices from 7 correspondences: null space of the 7x9 system + det(a F1 + (1-a) F2) = 0. Up to 3 solutions.
__device__ int solve_f7(const double (*p)[4], double (*F)[9]) {
  double A[7][9];
  for (int i = 0; i < 7; i++) {
    const double x = p[i][0], y = p[i][1], u = p[i][2], v = p[i][3];
    double* r = A[i];
    r[0] = u * x; r[1] = u * y; r[2] = u; r[3] = v * x; r[4] = v * y; r[5] = v; r[6] = x; r[7] = y; r[8] = 1;
  }
  // Gauss-Jordan, pivots in columns 0..6
  for (int c = 0; c < 7; c++) {
    int piv = c; double best = fabs(A[c][c]);
    for (int r = c + 1; r < 7; r++) if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); piv = r; }
    if (best < 1e-10) return 0;
    if (piv != c) for (int k = 0; k < 9; k++) { double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
    const double inv = 1.0 / A[c][c];
    for (int k = 0; k < 9; k++) A[c][k] *= inv;
    for (int r = 0; r < 7; r++)
      if (r != c) {
        const double f = A[r][c];
        for (int k = 0; k < 9; k++) A[r][k] -= f * A[c][k];
      }
  }
  double F1[9], F2[9], Dm[9];
  for (int i = 0; i < 7; i++) { F1[i] = -A[i][7]; F2[i] = -A[i][8]; }
  F1[7] = 1; F1[8] = 0; F2[7] = 0; F2[8] = 1;
  for (int i = 0; i < 9; i++) Dm[i] = F1[i] - F2[i];
  // cubic det(F2 + a D) = c3 a^3 + c2 a^2 + c1 a + c0 from four evaluations
  double T[9];
  auto det_at = [&](double a) { for (int i = 0; i < 9; i++) T[i] = F2[i] + a * Dm[i]; return det3(T); };
  const double d0 = det_at(0), d1 = det_at(1), dm1 = det_at(-1), d2 = det_at(2);
  // d1 = c3 + c2 + c1 + c0 ; dm1 = -c3 + c2 - c1 + c0 ; d2 = 8c3 + 4c2 + 2c1 + c0
  const double c0 = d0, c2 = (d1 + dm1) / 2 - d0;
  const double s1 = (d1 - dm1) / 2;                       // c3 + c1
  const double c3b = (d2 - c0 - 4 * c2 - 2 * s1) / 6.0;   // 8c3 + 2c1 - 2(c3 + c1) = 6 c3
  const double c1 = s1 - c3b;
  double roots[3]; int nr = 0;
  if (fabs(c3b) < 1e-14) {
    if (fabs(c2) > 1e-14) {
      double disc = c1 * c1 - 4 * c2 * c0;
      if (disc >= 0) { double sq = sqrt(disc); roots[nr++] = (-c1 + sq) / (2 * c2); roots[nr++] = (-c1 - sq) / (2 * c2); }
    } else if (fabs(c1) > 1e-14) roots[nr++] = -c0 / c1;
  } else {
    const double a = c2 / c3b, b = c1 / c3b, c = c0 / c3b;
    const double Q = (a * a - 3 * b) / 9, R = (2 * a * a * a - 9 * a * b + 27 * c) / 54;
    if (R * R < Q * Q * Q) {
      const double th = acos(R / sqrt(Q * Q * Q)), sq = -2 * sqrt(Q);
      roots[nr++] = sq * cos(th / 3) - a / 3;
      roots[nr++] = sq * cos((th + 2 * CUDART_PI) / 3) - a / 3;
      roots[nr++] = sq * cos((th - 2 * CUDART_PI) / 3) - a / 3;
    } else {
      const double Aa = -copysign(cbrt(fabs(R) + sqrt(R * R - Q * Q * Q)), R);
      const double Bb = (Aa != 0) ? Q / Aa : 0;
      roots[nr++] = (Aa + Bb) - a / 3;
    }
  }
  for (int k = 0; k < nr; k++)
    for (int i = 0; i < 9; i++) F[k][i] = F2[i] + roots[k] * Dm[i];
  return nr;
}

// ---- 9x9 symmetric eigenproblem (cyclic Jacobi, fp64): eigenvector of the smallest eigenvalue -----------------------
__device__ void smallest_eigvec9(double (*A)[9], double* vec) {
  double V[9][9];
  for (int i = 0; i < 9; i++) for (int j = 0; j < 9; j++) V[i][j] = (i == j);
  for (int sweep = 0; sweep < 12; sweep++) {
    double off = 0;
    for (int i = 0; i < 9; i++) for (int j = i + 1; j < 9; j++) off += A[i][j] * A[i][j];
    if (off < 1e-26) break;
    for (int p = 0; p < 9; p++)
      for (int q = p + 1; q < 9; q++) {
        if (fabs(A[p][q]) < 1e-300) continue;
        const double th = (A[q][q] - A[p][p]) / (2 * A[p][q]);
        const double t = copysign(1.0, th) / (fabs(th) + sqrt(th * th + 1)), c = 1 / sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < 9; k++) { double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
        for (int k = 0; k < 9; k++) { double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
        for (int k = 0; k < 9; k++) { double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
      }
  }
  int best = 0;
  for (int i = 1; i < 9; i++) if (A[i][i] < A[best][best]) best = i;
  for (int k = 0; k < 9; k++) vec[k] = V[k][best];
}
// nearest rank-2 matrix (zero the smallest singular value): F <- F - (F v)(v^T), v = right singular vector of sigma_min
__device__ void enforce_rank2(double* F) {
  double G[3][3];  // F^T F
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) G[i][j] = F[i] * F[j] + F[3 + i] * F[3 + j] + F[6 + i] * F[6 + j];
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 10; sweep++)
    for (int p = 0; p < 3; p++)
      for (int q = p + 1; q < 3; q++) {
        if (fabs(G[p][q]) < 1e-300) continue;
        const double th = (G[q][q] - G[p][p]) / (2 * G[p][q]);
        const double t = copysign(1.0, th) / (fabs(th) + sqrt(th * th + 1)), c = 1 / sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < 3; k++) { double a = G[k][p], b = G[k][q]; G[k][p] = c * a - s * b; G[k][q] = s * a + c * b; }
        for (int k = 0; k < 3; k++) { double a = G[p][k], b = G[q][k]; G[p][k] = c * a - s * b; G[q][k] = s * a + c * b; }
        for (int k = 0; k < 3; k++) { double a = V[k][p], b = V[k][q]; V[k][p] = c * a - s * b; V[k][q] = s * a + c * b; }
      }
  int m = 0;
  for (int i = 1; i < 3; i++) if (G[i][i] < G[m][m]) m = i;
  const double v[3] = {V[0][m], V[1][m], V[2][m]};
  for (int r = 0; r < 3; r++) {
    const double fv = F[3 * r] * v[0] + F[3 * r + 1] * v[1] + F[3 * r + 2] * v[2];
    for (int c = 0; c < 3; c++) F[3 * r + c] -= fv * v[c];
  }
}

struct Norm { float mx0, my0, s0, mx1, my1, s1; };  // Hartley normalisation: p' = (p - m) * s

__device__ void denorm_h(const double* Hn, const Norm& n, double* H) {  // H = T1^-1 Hn T0
  // T0 = [s0 0 -s0 mx0; 0 s0 -s0 my0; 0 0 1], T1^-1 = [1/s1 0 mx1; 0 1/s1 my1; 0 0 1]
  double A[9];
  for (int r = 0; r < 3; r++) {
    A[3 * r] = Hn[3 * r] * n.s0; A[3 * r + 1] = Hn[3 * r + 1] * n.s0;
    A[3 * r + 2] = Hn[3 * r + 2] - Hn[3 * r] * n.s0 * n.mx0 - Hn[3 * r + 1] * n.s0 * n.my0;
  }
  for (int c = 0; c < 3; c++) {
    H[c] = A[c] / n.s1 + n.mx1 * A[6 + c];
    H[3 + c] = A[3 + c] / n.s1 + n.my1 * A[6 + c];
    H[6 + c] = A[6 + c];
  }
}
__device__ void denorm_f(const double* Fn, const Norm& n, double* F) {  // F = T1^T Fn T0
  double A[9];
  for (int r = 0; r < 3; r++) {
    A[3 * r] = Fn[3 * r] * n.s0; A[3 * r + 1] = Fn[3 * r + 1] * n.s0;
    A[3 * r + 2] = Fn[3 * r + 2] - Fn[3 * r] * n.s0 * n.mx0 - Fn[3 * r + 1] * n.s0 * n.my0;
  }
  for (int c = 0; c < 3; c++) {
    F[c] = n.s1 * A[c];
    F[3 + c] = n.s1 * A[3 + c];
    F[6 + c] = -n.s1 * n.mx1 * A[c] - n.s1 * n.my1 * A[3 + c] + A[6 + c];
  }
}

template <typename T>
__device__ T block_sum(T v, T* scratch) {  // scratch: MS_THREADS/32 entries
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (threadIdx.x % 32 == 0) scratch[threadIdx.x / 32] = v;
  __syncthreads();
  T s = 0;
  for (int w = 0; w < MS_THREADS / 32; w++) s += scratch[w];
  return s;
}

// model_type 0 = homography (4-point), 1 = fundamental matrix (7-point)
// G = CTAs per correspondence set.  G > 1: the G CTAs of a set form a thread-block CLUSTER; each draws its own 256 hypotheses
// per round and the round's winner is agreed on through distributed shared memory (every CTA reads the G candidate scores,
// copies the winning model and carries the same best_score / required-iterations state, so the CTAs leave the loop together).
// Few sets (the single-pair API call: F and H of one pair) then still fill the GPU: 8 x 256 hypotheses per round and set.
template <int G>
__global__ void __launch_bounds__(MS_THREADS) magsac_kernel(const float* __restrict__ pts0, const float* __restrict__ pts1,
                                                            const int* __restrict__ counts, int cap, int model_type,
                                                            float threshold, float confidence, int max_iters, unsigned seed,
                                                            double* __restrict__ models, unsigned char* __restrict__ masks,
                                                            int* __restrict__ n_inliers, int* __restrict__ n_iters) {
  extern __shared__ __align__(16) float ms_smem[];
  namespace cg = cooperative_groups;
  const int set = blockIdx.x / G, crank = blockIdx.x % G, tid = threadIdx.x, K = counts[set];
  __shared__ float s_cl_score;        // this CTA's best candidate of the round, read by the other CTAs of the cluster
  __shared__ float s_cl_model[9];
  float* X0 = ms_smem; float* Y0 = X0 + cap; float* X1 = Y0 + cap; float* Y1 = X1 + cap;
  __shared__ double s_red[MS_THREADS / 32];
  __shared__ float s_best_score[MS_THREADS / 32];
  __shared__ int s_best_tid[MS_THREADS / 32];
  __shared__ float s_model[9];
  __shared__ double s_model_d[9];
  __shared__ double s_AtA[9][9];
  __shared__ Norm s_norm;
  __shared__ int s_stop;
  const int S = model_type == 0 ? 4 : 7;
  double* out_model = models + (long long)set * 9;
  unsigned char* out_mask = masks + (long long)set * cap;
  if (K < S + (model_type == 0 ? 0 : 1)) {  // cv2: not enough points -> no model (all CTAs of the cluster leave together)
    if (crank == 0) {
      for (int i = tid; i < cap; i += MS_THREADS) out_mask[i] = 0;
      if (tid < 9) out_model[tid] = 0.0;
      if (tid == 0) { n_inliers[set] = 0; n_iters[set] = 0; }
    }
    return;
  }
  for (int i = tid; i < K; i += MS_THREADS) {
    X0[i] = pts0[((long long)set * cap + i) * 2]; Y0[i] = pts0[((long long)set * cap + i) * 2 + 1];
    X1[i] = pts1[((long long)set * cap + i) * 2]; Y1[i] = pts1[((long long)set * cap + i) * 2 + 1];
  }
  __syncthreads();
  {  // Hartley normalisation parameters
    double sx0 = 0, sy0 = 0, sx1 = 0, sy1 = 0;
    for (int i = tid; i < K; i += MS_THREADS) { sx0 += X0[i]; sy0 += Y0[i]; sx1 += X1[i]; sy1 += Y1[i]; }
    const double mx0 = block_sum(sx0, s_red) / K, my0 = block_sum(sy0, s_red) / K, mx1 = block_sum(sx1, s_red) / K, my1 = block_sum(sy1, s_red) / K;
    double d0 = 0, d1 = 0;
    for (int i = tid; i < K; i += MS_THREADS) {
      d0 += sqrt((X0[i] - mx0) * (X0[i] - mx0) + (Y0[i] - my0) * (Y0[i] - my0));
      d1 += sqrt((X1[i] - mx1) * (X1[i] - mx1) + (Y1[i] - my1) * (Y1[i] - my1));
    }
    const double md0 = block_sum(d0, s_red) / K, md1 = block_sum(d1, s_red) / K;
    if (tid == 0) s_norm = Norm{(float)mx0, (float)my0, (float)(1.41421356 / fmax(md0, 1e-9)), (float)mx1, (float)my1, (float)(1.41421356 / fmax(md1, 1e-9))};
    __syncthreads();
  }
  const Norm nm = s_norm;
  const MagsacConsts mc = make_consts(threshold);
  auto resid = [&](const float* M, int i) { return model_type == 0 ? resid_h(M, X0[i], Y0[i], X1[i], Y1[i]) : resid_f(M, X0[i], Y0[i], X1[i], Y1[i]); };
  float best_score = CUDART_INF_F;
  auto score_model = [&](const float* M) {  // full pass over the correspondences by ONE thread (broadcast smem reads)
    float s = 0.f;
    for (int i = 0; i < K; i++) {
      s += magsac_loss(resid(M, i), mc);
      if ((i & 63) == 63 && s > best_score) return CUDART_INF_F;  // cannot beat the incumbent any more
    }
    return s;
  };

  int iters = 0, required = max_iters, best_inl = 0;
  const float log_fail = logf(fmaxf(1.f - confidence, 1e-12f));
  for (int round = 0; iters < required && iters < max_iters; round++) {
    // ---- one hypothesis per thread
    double smp[7][4];
    unsigned st = pcg_hash(seed ^ (set * 9781u + (round * G + crank) * 6271u + tid * 0x9E3779B9u));
    int idx[7];
    for (int s = 0; s < S; s++) {
      bool dup;
      do {
        st = pcg_hash(st);
        idx[s] = (int)(st % (unsigned)K);
        dup = false;
        for (int t = 0; t < s; t++) dup |= (idx[t] == idx[s]);
      } while (dup);
      smp[s][0] = (X0[idx[s]] - nm.mx0) * nm.s0; smp[s][1] = (Y0[idx[s]] - nm.my0) * nm.s0;
      smp[s][2] = (X1[idx[s]] - nm.mx1) * nm.s1; smp[s][3] = (Y1[idx[s]] - nm.my1) * nm.s1;
    }
    float my_score = CUDART_INF_F, my_model[9];
    if (model_type == 0) {
      double Hn[9], H[9];
      if (solve_h4(smp, Hn)) {
        denorm_h(Hn, nm, H);
        float M[9];
        for (int k = 0; k < 9; k++) M[k] = (float)(H[k] / (fabs(H[8]) > 1e-12 ? H[8] : 1.0));
        my_score = score_model(M);
        for (int k = 0; k < 9; k++) my_model[k] = M[k];
      }
    } else {
      double Fn[3][9], Fd[9];
      const int ns = solve_f7(smp, Fn);
      for (int c = 0; c < ns; c++) {
        denorm_f(Fn[c], nm, Fd);
        double nrm = 0;
        for (int k = 0; k < 9; k++) nrm += Fd[k] * Fd[k];
        nrm = 1.0 / sqrt(fmax(nrm, 1e-300));
        float M[9];
        for (int k = 0; k < 9; k++) M[k] = (float)(Fd[k] * nrm);
        const float sc = score_model(M);
        if (sc < my_score) { my_score = sc; for (int k = 0; k < 9; k++) my_model[k] = M[k]; }
      }
    }
    if (!(my_score == my_score)) my_score = CUDART_INF_F;  // NaN guard
    // ---- block arg-min
    float bs = my_score; int bt = tid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float s2 = __shfl_xor_sync(0xffffffffu, bs, o); int t2 = __shfl_xor_sync(0xffffffffu, bt, o);
      if (s2 < bs || (s2 == bs && t2 < bt)) { bs = s2; bt = t2; }
    }
    if (tid % 32 == 0) { s_best_score[tid / 32] = bs; s_best_tid[tid / 32] = bt; }
    __syncthreads();
    bs = s_best_score[0]; bt = s_best_tid[0];
    for (int w = 1; w < MS_THREADS / 32; w++) if (s_best_score[w] < bs) { bs = s_best_score[w]; bt = s_best_tid[w]; }
    iters += MS_THREADS * G;
    bool improved;
    if (G == 1) {
      improved = bs < best_score;
      if (improved) {
        best_score = bs;
        if (tid == bt) for (int k = 0; k < 9; k++) s_model[k] = my_model[k];
      }
      __syncthreads();
    } else {
      // publish this CTA's candidate, then every CTA of the cluster picks the same winner (lowest score, lowest rank on ties)
      if (tid == bt) { s_cl_score = bs; for (int k = 0; k < 9; k++) s_cl_model[k] = my_model[k]; }
      cg::cluster_group cluster = cg::this_cluster();
      cluster.sync();
      float ws = CUDART_INF_F; int wr = 0;
      for (int r = 0; r < G; r++) {
        const float sr = *cluster.map_shared_rank(&s_cl_score, r);
        if (sr < ws) { ws = sr; wr = r; }
      }
      improved = ws < best_score;
      if (improved) {
        best_score = ws;
        if (tid < 9) s_model[tid] = cluster.map_shared_rank(s_cl_model, wr)[tid];
      }
      cluster.sync();     // nobody overwrites its candidate while a neighbour still reads it; also orders s_model for the block
    }
    if (improved) {  // inlier count of the new best model -> adaptive termination
      int cnt = 0;
      for (int i = tid; i < K; i += MS_THREADS) cnt += resid(s_model, i) <= mc.thr2_inl ? 1 : 0;
      best_inl = (int)block_sum((double)cnt, s_red);
      const float w = fminf((float)best_inl / (float)K, 0.9999f);
      const float denom = logf(fmaxf(1.f - powf(w, (float)S), 1e-12f));
      required = (best_inl <= S) ? max_iters : (int)fminf((float)max_iters, ceilf(log_fail / denom));
    }
    __syncthreads();
  }

  // ---- iteratively re-weighted least squares with the MAGSAC++ weights (sigma-consensus++ polishing)
  if (best_score < CUDART_INF_F) {
    for (int it = 0; it < 4; it++) {
      // A^T W A of the normalised DLT system, accumulated row by row of the upper triangle
      for (int r = 0; r < 9; r++) {
        double acc[9];
        for (int c = 0; c < 9; c++) acc[c] = 0;
        for (int i = tid; i < K; i += MS_THREADS) {
          const float wgt = magsac_weight(resid(s_model, i), mc);
          if (wgt <= 0.f) continue;
          const double x = (X0[i] - nm.mx0) * nm.s0, y = (Y0[i] - nm.my0) * nm.s0, u = (X1[i] - nm.mx1) * nm.s1, v = (Y1[i] - nm.my1) * nm.s1;
          if (model_type == 0) {
            const double a0[9] = {x, y, 1, 0, 0, 0, -u * x, -u * y, -u}, a1[9] = {0, 0, 0, x, y, 1, -v * x, -v * y, -v};
            for (int c = r; c < 9; c++) acc[c] += wgt * (a0[r] * a0[c] + a1[r] * a1[c]);
          } else {
            const double a[9] = {u * x, u * y, u, v * x, v * y, v, x, y, 1};
            for (int c = r; c < 9; c++) acc[c] += wgt * a[r] * a[c];
          }
        }
        for (int c = r; c < 9; c++) {
          const double s = block_sum(acc[c], s_red);
          if (tid == 0) { s_AtA[r][c] = s; s_AtA[c][r] = s; }
        }
      }
      __syncthreads();
      if (tid == 0) {
        double A[9][9], vec[9], Md[9];
        for (int i = 0; i < 9; i++) for (int j = 0; j < 9; j++) A[i][j] = s_AtA[i][j];
        smallest_eigvec9(A, vec);
        if (model_type == 0) {
          denorm_h(vec, nm, Md);
          const double d = fabs(Md[8]) > 1e-12 ? Md[8] : 1.0;
          for (int k = 0; k < 9; k++) Md[k] /= d;
        } else {
          enforce_rank2(vec);
          denorm_f(vec, nm, Md);
          double nrm = 0;
          for (int k = 0; k < 9; k++) nrm += Md[k] * Md[k];
          nrm = 1.0 / sqrt(fmax(nrm, 1e-300));
          for (int k = 0; k < 9; k++) Md[k] *= nrm;
        }
        for (int k = 0; k < 9; k++) s_model_d[k] = Md[k];
      }
      __syncthreads();
      // accept the refit only if it lowers the MAGSAC++ loss
      float cand[9];
      for (int k = 0; k < 9; k++) cand[k] = (float)s_model_d[k];
      double part = 0;
      for (int i = tid; i < K; i += MS_THREADS) part += magsac_loss(resid(cand, i), mc);
      const float cs = (float)block_sum(part, s_red);
      if (tid == 0) s_stop = !(cs < best_score);
      __syncthreads();
      if (s_stop) break;
      best_score = cs;
      if (tid < 9) s_model[tid] = cand[tid];
      __syncthreads();
    }
  }
  // ---- outputs: inlier mask at `threshold`, model scaled like OpenCV (last element 1 when possible); CTA 0 of the cluster writes
  if (crank != 0) return;
  int cnt = 0;
  for (int i = tid; i < cap; i += MS_THREADS) {
    const bool in = (i < K) && best_score < CUDART_INF_F && resid(s_model, i) <= mc.thr2_inl;
    out_mask[i] = in ? 1 : 0;
    cnt += in;
  }
  const int total = (int)block_sum((double)cnt, s_red);
  if (tid == 0) {
    n_inliers[set] = total; n_iters[set] = iters;
    const double d = (best_score < CUDART_INF_F && fabsf(s_model[8]) > 1e-12f) ? (double)s_model[8] : 1.0;
    for (int k = 0; k < 9; k++) out_model[k] = best_score < CUDART_INF_F ? (double)s_model[k] / d : 0.0;
  }
}
}  // namespace

extern "C" int imw_magsac(int n_sets, int cap, const float* pts0, const float* pts1, const int* counts, int model_type,
                          float threshold, float confidence, int max_iters, unsigned seed, double* models, unsigned char* masks,
                          int* n_inliers, int* n_iters, cudaStream_t st) {
  IMW_REQUIRE(n_sets > 0 && cap > 0 && cap <= 12288, "imw_magsac: cap must be in (0, 12288] (got %d)", cap);
  IMW_REQUIRE(model_type == 0 || model_type == 1, "imw_magsac: model_type 0 (homography) or 1 (fundamental)");
  IMW_REQUIRE(threshold > 0.f && confidence > 0.f && confidence < 1.f && max_iters > 0, "imw_magsac: bad threshold/confidence/max_iters");
  const size_t smem = (size_t)4 * cap * sizeof(float);
  // CTAs per set: fill the SMs when there are few sets (cluster of up to 8 CTAs = portable cluster size)
  const int sms = imw_num_sms();
  const int G = n_sets * 8 <= sms ? 8 : (n_sets * 4 <= sms ? 4 : (n_sets * 2 <= sms ? 2 : 1));
  if (G == 1) {
    IMW_CHECK_CUDA(cudaFuncSetAttribute(magsac_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    magsac_kernel<1><<<n_sets, MS_THREADS, smem, st>>>(pts0, pts1, counts, cap, model_type, threshold, confidence, max_iters, seed, models,
                                                       masks, n_inliers, n_iters);
    IMW_CHECK_LAUNCH();
    return IMW_OK;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_sets * G)); cfg.blockDim = dim3(MS_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
#define MS_LAUNCH(GG)                                                                                                             \
  do {                                                                                                                            \
    IMW_CHECK_CUDA(cudaFuncSetAttribute(magsac_kernel<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
    IMW_CHECK_CUDA(cudaLaunchKernelEx(&cfg, magsac_kernel<GG>, pts0, pts1, counts, cap, model_type, threshold, confidence, max_iters, \
                                      seed, models, masks, n_inliers, n_iters));                                                  \
  } while (0)
  if (G == 8) MS_LAUNCH(8); else if (G == 4) MS_LAUNCH(4); else MS_LAUNCH(2);
#undef MS_LAUNCH
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}
