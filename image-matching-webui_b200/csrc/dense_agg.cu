// Dense-match aggregation on the GPU: the per-pair array work of imcui/hloc/match_dense.py:37-121
//   imw_quantize_keypoints  `to_cpts` (:37-40): keypoints -> grid-cell indices + the cell's coordinate
//                           np.round(np.round((k + 0.5) / ps) * ps - 0.5, 2), in the reference's fp32 arithmetic
//   imw_nearest_point       `assign_keypoints(update=False)` (:52-59): nearest stored keypoint within max_error
//                           (the reference builds a scipy KDTree per call; K x M brute force is a few microseconds here)
//   imw_unique_matches      `kpids_to_matches0` (:99-121): drop unassigned matches, resolve n-to-1 conflicts by keeping,
//                           per keypoint id on either side, the best-scoring match (`get_unique_matches`: intersection of
//                           the two arg-max sets), scatter to matches0 / fp16-rounded scores0.
// The keypoint-id dictionary that persists across the pairs of an image (cp_to_id, :68-83) stays on the host
// (hloc/dense_aggregate.py): it is a sequential first-come numbering by construction.
// Index / byte work: one thread per correspondence, 64-bit atomicMax as the per-id arg-max (score bits | inverted index).
#include <cuda_fp16.h>
#include <math.h>

#include "../../include/imw_b200.h"
#include "common.cuh"

namespace {

// cells [n][cap][2] int32 (rint((k + 0.5) / ps)), coords [n][cap][2] fp32 (the tuple the reference uses as dictionary key)
__global__ void da_quantize_kernel(const float* __restrict__ kpts, const int* __restrict__ counts, float ps, int* __restrict__ cells,
                                   float* __restrict__ coords, int cap) {
  const int z = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (counts ? counts[z] : cap)) return;
  const long long o = ((long long)z * cap + i) * 2;
#pragma unroll
  for (int d = 0; d < 2; d++) {
    const float k = kpts[o + d];
    if (ps > 0.f) {
      const float q = rintf(__fdiv_rn(__fadd_rn(k, 0.5f), ps));                 // np.round: half to even
      const float v = __fsub_rn(__fmul_rn(q, ps), 0.5f);
      cells[o + d] = (int)q;
      coords[o + d] = __fdiv_rn(rintf(__fmul_rn(v, 100.f)), 100.f);             // np.round(v, 2)
    } else {                                                                     // ps == 0: the raw keypoint is the key
      cells[o + d] = __float_as_int(k);
      coords[o + d] = k;
    }
  }
}

// nearest of M stored points for each of K query points; ids[i] = -1 beyond max_error (KDTree.query + threshold, :56-58)
__global__ void __launch_bounds__(256) da_nearest_kernel(const float* __restrict__ q, int K, const float* __restrict__ pts, int M, float max_error,
                                                         int* __restrict__ ids) {
  extern __shared__ float s_pts[];   // tiles of 1024 stored points
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float x = i < K ? q[2 * i] : 0.f, y = i < K ? q[2 * i + 1] : 0.f;
  float best = INFINITY; int bj = -1;
  for (int m0 = 0; m0 < M; m0 += 1024) {
    const int n = min(1024, M - m0);
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * n; t += blockDim.x) s_pts[t] = pts[2 * m0 + t];
    __syncthreads();
    for (int j = 0; j < n; j++) {
      const float dx = s_pts[2 * j] - x, dy = s_pts[2 * j + 1] - y;
      const float d2 = dx * dx + dy * dy;
      if (d2 < best) { best = d2; bj = m0 + j; }
    }
  }
  if (i < K) ids[i] = (bj >= 0 && sqrtf(best) <= max_error) ? bj : -1;
}

__device__ __forceinline__ unsigned long long pack_best(float score, int idx) {
  unsigned u = __float_as_uint(score);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);                 // monotone map of fp32 to u32
  return ((unsigned long long)u << 32) | (unsigned)(0x7fffffff - idx);   // equal scores: the smaller index wins
}

__global__ void da_best_kernel(const int* __restrict__ ids0, const int* __restrict__ ids1, const float* __restrict__ scores,
                               const int* __restrict__ counts, unsigned long long* __restrict__ best0, unsigned long long* __restrict__ best1,
                               int cap, int id_cap) {
  const int p = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= counts[p]) return;
  const int a = ids0[(long long)p * cap + i], b = ids1[(long long)p * cap + i];
  if (a < 0 || b < 0 || a >= id_cap || b >= id_cap) return;
  const unsigned long long v = pack_best(scores[(long long)p * cap + i], i);
  atomicMax(best0 + (long long)p * id_cap + a, v);
  atomicMax(best1 + (long long)p * id_cap + b, v);
}

__global__ void da_scatter_kernel(const int* __restrict__ ids0, const int* __restrict__ ids1, const float* __restrict__ scores,
                                  const int* __restrict__ counts, const unsigned long long* __restrict__ best0,
                                  const unsigned long long* __restrict__ best1, int* __restrict__ matches0, __half* __restrict__ scores0,
                                  int* __restrict__ n_kps0, int cap, int id_cap) {
  const int p = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= counts[p]) return;
  const int a = ids0[(long long)p * cap + i], b = ids1[(long long)p * cap + i];
  if (a < 0 || b < 0 || a >= id_cap || b >= id_cap) return;
  const unsigned long long v = pack_best(scores[(long long)p * cap + i], i);
  if (best0[(long long)p * id_cap + a] == v && best1[(long long)p * id_cap + b] == v) {   // arg-max on both sides
    matches0[(long long)p * id_cap + a] = b;
    scores0[(long long)p * id_cap + a] = __float2half_rn(scores[(long long)p * cap + i]);
    atomicMax(n_kps0 + p, a + 1);
  }
}

__global__ void da_init_kernel(unsigned long long* __restrict__ best0, unsigned long long* __restrict__ best1, int* __restrict__ matches0,
                               __half* __restrict__ scores0, int* __restrict__ n_kps0, long long n, int P) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { best0[i] = 0ull; best1[i] = 0ull; matches0[i] = -1; scores0[i] = __float2half_rn(0.f); }
  if (i < P) n_kps0[i] = 0;
}

}  // namespace

extern "C" int imw_quantize_keypoints(int n_sets, int cap, const float* keypoints, const int* counts, float cell_size, int* cells,
                                      float* coords, cudaStream_t st) {
  IMW_REQUIRE(n_sets > 0 && cap > 0 && keypoints && cells && coords && cell_size >= 0.f, "imw_quantize_keypoints: bad arguments");
  da_quantize_kernel<<<dim3(ceil_div(cap, 256), n_sets), 256, 0, st>>>(keypoints, counts, cell_size, cells, coords, cap);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}

extern "C" int imw_nearest_point(int n_query, const float* query, int n_points, const float* points, float max_error, int* ids,
                                 cudaStream_t st) {
  IMW_REQUIRE(n_query > 0 && n_points > 0 && query && points && ids, "imw_nearest_point: bad arguments");
  da_nearest_kernel<<<ceil_div(n_query, 256), 256, 2 * 1024 * sizeof(float), st>>>(query, n_query, points, n_points, max_error, ids);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}

extern "C" size_t imw_unique_matches_workspace_bytes(int n_pairs, int id_cap) { return (size_t)2 * n_pairs * id_cap * sizeof(unsigned long long) + 512; }

extern "C" int imw_unique_matches(int n_pairs, int cap, int id_cap, const int* ids0, const int* ids1, const float* scores, const int* counts,
                                  int* matches0, void* scores0_f16, int* n_kps0, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  IMW_REQUIRE(n_pairs > 0 && cap > 0 && id_cap > 0 && ids0 && ids1 && scores && counts && matches0 && scores0_f16 && n_kps0,
              "imw_unique_matches: bad arguments");
  Workspace ws(workspace, workspace_bytes);
  unsigned long long* best0 = ws.take<unsigned long long>((size_t)n_pairs * id_cap);
  unsigned long long* best1 = ws.take<unsigned long long>((size_t)n_pairs * id_cap);
  if (ws.overflow) { imw_set_error("imw_unique_matches: workspace too small (%zu < %zu)", workspace_bytes, ws.off); return IMW_ERR_WORKSPACE; }
  const long long n = (long long)n_pairs * id_cap;
  da_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(best0, best1, matches0, (__half*)scores0_f16, n_kps0, n, n_pairs);
  IMW_CHECK_LAUNCH();
  const dim3 grid(ceil_div(cap, 256), n_pairs);
  da_best_kernel<<<grid, 256, 0, st>>>(ids0, ids1, scores, counts, best0, best1, cap, id_cap);
  IMW_CHECK_LAUNCH();
  da_scatter_kernel<<<grid, 256, 0, st>>>(ids0, ids1, scores, counts, best0, best1, matches0, (__half*)scores0_f16, n_kps0, cap, id_cap);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}
