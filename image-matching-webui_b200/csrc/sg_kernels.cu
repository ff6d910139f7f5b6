// SuperGlue matcher on the GPU.
// Follows third_party/SuperGluePretrainedNetwork/models/superglue.py:65-283 for batches of independent pairs:
// keypoint encoder MLP (BatchNorm folded by the host), 18 alternating self/cross attentional-propagation layers
// (same tcgen05 3xTF32 GEMM / flash-attention kernels as LightGlue), final projection, dustbin-augmented
// log-space Sinkhorn on the materialised score matrix, mutual arg-max + threshold.
#include <math_constants.h>

#include "../../include/imw_b200.h"
#include "common.cuh"
#include "gemm_simt.cuh"
#include "lg_internal.h"
#include "tc_attn.cuh"
#include "tc_gemm.cuh"
#include "token_epilogues.cuh"

namespace {

// ---- keypoint encoder input: [(k - center) / scaling, score, 0...] (superglue.py:65-72,82-84) -----------
__global__ void sg_kenc_input_kernel(const float* __restrict__ kpts, const float* __restrict__ scores,
                                     const int* __restrict__ counts, const int* __restrict__ image_wh, float* __restrict__ kin,
                                     int cap) {
  const int z = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= counts[z]) return;
  const float w = (float)image_wh[2 * z], h = (float)image_wh[2 * z + 1];
  const float scaling = __fmul_rn(fmaxf(w, h), 0.7f);
  float* o = kin + ((long long)z * cap + i) * 16;
  o[0] = __fdiv_rn(__fsub_rn(kpts[((long long)z * cap + i) * 2], w / 2), scaling);
  o[1] = __fdiv_rn(__fsub_rn(kpts[((long long)z * cap + i) * 2 + 1], h / 2), scaling);
  o[2] = scores[(long long)z * cap + i];
#pragma unroll
  for (int k = 3; k < 16; k++) o[k] = 0.f;
}

__global__ void __launch_bounds__(256) sg_init_tokens_kernel(const float* __restrict__ desc, float* __restrict__ xm,
                                                             int* __restrict__ matches, float* __restrict__ mscores,
                                                             const int* __restrict__ counts, int cap) {
  const int z = blockIdx.y, row = blockIdx.x * 4 + threadIdx.x / 64, t = threadIdx.x % 64;
  if (row >= cap) return;
  if (t == 0) { matches[(long long)z * cap + row] = -1; mscores[(long long)z * cap + row] = 0.f; }
  if (row < counts[z])
    reinterpret_cast<float4*>(xm + ((long long)z * cap + row) * 512)[t] = reinterpret_cast<const float4*>(desc + ((long long)z * cap + row) * D)[t];
}

__global__ void sg_state_kernel(const int* __restrict__ counts, int* __restrict__ empty, int* __restrict__ mcnt, int* __restrict__ ncnt, int P) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  int m = counts[2 * p], n = counts[2 * p + 1];
  empty[p] = (m == 0 || n == 0) ? 1 : 0;
  mcnt[p] = m; ncnt[p] = n;
}

// ---- log-space Sinkhorn with dustbins (superglue.py:143-172) ----------------------------------------------
// couplings C(i,j) = S[i][j] (i<m, j<n), alpha on the dustbin row/column.  u: [P][cap+1], v: [P][cap+1].
struct SkArgs {
  const float* S; float* u; float* v; const int* counts; const int* empty; int cap; float alpha;
};
__device__ __forceinline__ float sk_log_mu(int i, int m, int n) {  // log_mu (rows) : norm, last = log(n) + norm
  const float norm = -logf((float)(m + n));
  return i < m ? norm : logf((float)n) + norm;
}

// u[i] = log_mu[i] - logsumexp_j(C(i,j) + v[j]); one warp per row (i = m is the dustbin row)
__global__ void __launch_bounds__(256) sk_row_kernel(SkArgs a) {
  const int p = blockIdx.y, i = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (a.empty[p]) return;
  const int m = a.counts[2 * p], n = a.counts[2 * p + 1];
  if (i > m) return;
  const float* Srow = a.S + ((long long)p * a.cap + i) * a.cap;
  const float* v = a.v + (long long)p * (a.cap + 1);
  float mx = -CUDART_INF_F, sm = 0.f;
  for (int j = lane; j <= n; j += 32) {
    float c = (i < m && j < n) ? Srow[j] : a.alpha;
    float x = c + v[j];
    if (x > mx) { sm = sm * __expf(mx - x) + 1.f; mx = x; } else sm += __expf(x - mx);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float m2 = __shfl_xor_sync(0xffffffffu, mx, o), s2 = __shfl_xor_sync(0xffffffffu, sm, o);
    float mm = fmaxf(mx, m2);
    sm = (mm == -CUDART_INF_F) ? 0.f : sm * __expf(mx - mm) + s2 * __expf(m2 - mm);
    mx = mm;
  }
  if (lane == 0) a.u[(long long)p * (a.cap + 1) + i] = sk_log_mu(i, m, n) - (mx + logf(sm));
}

// v[j] = log_nu[j] - logsumexp_i(C(i,j) + u[i]); CTA = 32 columns, 8 warps stride the rows (coalesced 128 B reads)
__global__ void __launch_bounds__(256) sk_col_kernel(SkArgs a) {
  const int p = blockIdx.y, j = blockIdx.x * 32 + threadIdx.x % 32, w = threadIdx.x / 32;
  if (a.empty[p]) return;
  const int m = a.counts[2 * p], n = a.counts[2 * p + 1];
  __shared__ float s_m[8][32], s_s[8][32];
  const float* Sp = a.S + (long long)p * a.cap * a.cap;
  const float* u = a.u + (long long)p * (a.cap + 1);
  float mx = -CUDART_INF_F, sm = 0.f;
  if (j <= n) {
    for (int i = w; i <= m; i += 8) {
      float c = (i < m && j < n) ? Sp[(long long)i * a.cap + j] : a.alpha;
      float x = c + u[i];
      if (x > mx) { sm = sm * __expf(mx - x) + 1.f; mx = x; } else sm += __expf(x - mx);
    }
  }
  s_m[w][threadIdx.x % 32] = mx; s_s[w][threadIdx.x % 32] = sm;
  __syncthreads();
  if (w == 0 && j <= n) {
    for (int k = 1; k < 8; k++) {
      float m2 = s_m[k][threadIdx.x], s2 = s_s[k][threadIdx.x], mm = fmaxf(mx, m2);
      sm = (mm == -CUDART_INF_F) ? 0.f : sm * __expf(mx - mm) + s2 * __expf(m2 - mm);
      mx = mm;
    }
    a.v[(long long)p * (a.cap + 1) + j] = sk_log_mu(j, n, m) - (mx + logf(sm));  // log_nu: roles of m, n swapped
  }
}

// Z = ((C + u) + v) - norm on [:m, :n]; row max / arg-max (first index) per image-0 keypoint, column max per image-1
__global__ void __launch_bounds__(256) sk_rowmax_kernel(SkArgs a, float* __restrict__ best_v, int* __restrict__ best_j) {
  const int p = blockIdx.y, i = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (a.empty[p]) return;
  const int m = a.counts[2 * p], n = a.counts[2 * p + 1];
  if (i >= m) return;
  const float norm = -logf((float)(m + n));
  const float* Srow = a.S + ((long long)p * a.cap + i) * a.cap;
  const float* v = a.v + (long long)p * (a.cap + 1);
  const float ui = a.u[(long long)p * (a.cap + 1) + i];
  float bv = -CUDART_INF_F; int bj = 0x7fffffff;
  for (int j = lane; j < n; j += 32) {
    float z = __fsub_rn(__fadd_rn(__fadd_rn(Srow[j], ui), v[j]), norm);
    if (z > bv || (z == bv && j < bj)) { bv = z; bj = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float v2 = __shfl_xor_sync(0xffffffffu, bv, o); int j2 = __shfl_xor_sync(0xffffffffu, bj, o);
    if (v2 > bv || (v2 == bv && j2 < bj)) { bv = v2; bj = j2; }
  }
  if (lane == 0) { best_v[(long long)(2 * p) * a.cap + i] = bv; best_j[(long long)(2 * p) * a.cap + i] = bj; }
}
__global__ void __launch_bounds__(256) sk_colmax_kernel(SkArgs a, float* __restrict__ best_v, int* __restrict__ best_j) {
  const int p = blockIdx.y, j = blockIdx.x * 32 + threadIdx.x % 32, w = threadIdx.x / 32;
  if (a.empty[p]) return;
  const int m = a.counts[2 * p], n = a.counts[2 * p + 1];
  __shared__ float s_v[8][32]; __shared__ int s_i[8][32];
  const float norm = -logf((float)(m + n));
  const float* Sp = a.S + (long long)p * a.cap * a.cap;
  const float* u = a.u + (long long)p * (a.cap + 1);
  float bv = -CUDART_INF_F; int bi = 0x7fffffff;
  if (j < n) {
    const float vj = a.v[(long long)p * (a.cap + 1) + j];
    for (int i = w; i < m; i += 8) {
      float z = __fsub_rn(__fadd_rn(__fadd_rn(Sp[(long long)i * a.cap + j], u[i]), vj), norm);
      if (z > bv || (z == bv && i < bi)) { bv = z; bi = i; }
    }
  }
  s_v[w][threadIdx.x % 32] = bv; s_i[w][threadIdx.x % 32] = bi;
  __syncthreads();
  if (w == 0 && j < n) {
    for (int k = 1; k < 8; k++) {
      float v2 = s_v[k][threadIdx.x]; int i2 = s_i[k][threadIdx.x];
      if (v2 > bv || (v2 == bv && i2 < bi)) { bv = v2; bi = i2; }
    }
    best_v[(long long)(2 * p + 1) * a.cap + j] = bv; best_j[(long long)(2 * p + 1) * a.cap + j] = bi;
  }
}

// mutual check, exp, threshold (superglue.py:266-276)
__global__ void __launch_bounds__(256) sg_match_kernel(const float* __restrict__ best_v, const int* __restrict__ best_j,
                                                       const int* __restrict__ counts, const int* __restrict__ empty,
                                                       int* __restrict__ matches, float* __restrict__ mscores, int cap, float th) {
  const int z = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x, zo = z ^ 1;
  if (empty[z >> 1] || i >= counts[z]) return;
  const long long io = (long long)z * cap + i;
  const int jr = best_j[io];
  const bool j_ok = (unsigned)jr < (unsigned)counts[zo];   // NaN scores leave the arg-max at its init value: unmatched
  const int j = j_ok ? jr : 0;
  const long long jo = (long long)zo * cap + j;
  const bool mutual = j_ok && best_j[jo] == i;
  const float sc0 = (z & 1) ? __expf(best_v[jo]) : __expf(best_v[io]);
  const float ms = mutual ? sc0 : 0.f;
  matches[io] = (mutual && ms > th) ? j : -1;
  mscores[io] = ms;
}

struct SGBuffers {
  float *xm, *q, *k, *v, *ctx, *h, *t0, *t1, *kin, *md, *S, *u, *vv, *best_v;
  int *best_j, *empty, *mcnt, *ncnt;
};
size_t sg_carve(Workspace& ws, SGBuffers& b, int P, int cap) {
  const size_t S = 2 * (size_t)P, T = S * cap;
  b.xm = ws.take<float>(T * 512);
  b.q = ws.take<float>(2 * T * D); b.k = ws.take<float>(2 * T * D); b.v = ws.take<float>(2 * T * D);
  b.ctx = ws.take<float>(T * D); b.h = ws.take<float>(T * 512);
  b.t0 = ws.take<float>(T * D); b.t1 = ws.take<float>(T * D); b.kin = ws.take<float>(T * 16); b.md = ws.take<float>(T * D);
  b.S = ws.take<float>((size_t)P * cap * cap);
  b.u = ws.take<float>((size_t)P * (cap + 1)); b.vv = ws.take<float>((size_t)P * (cap + 1));
  b.best_v = ws.take<float>(T); b.best_j = ws.take<int>(T);
  b.empty = ws.take<int>(P); b.mcnt = ws.take<int>(P); b.ncnt = ws.take<int>(P);
  return ws.off;
}
}  // namespace

extern "C" size_t imw_superglue_workspace_bytes(int n_pairs, int cap) {
  Workspace ws(nullptr, 0);
  SGBuffers b;
  return sg_carve(ws, b, n_pairs, cap) + 256;
}

extern "C" int imw_superglue_forward(const imw_sg_weights* W, const imw_sg_conf* conf, int n_pairs, int cap, const float* kpts,
                                     const float* scores, const float* desc, const int* counts, const int* image_wh,
                                     int* matches, float* mscores, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  IMW_REQUIRE(W && conf && n_pairs > 0 && cap > 0 && cap % 4 == 0, "imw_superglue_forward: bad arguments");
  IMW_REQUIRE(W->n_layers >= 1 && W->n_layers <= IMW_SG_MAX_LAYERS, "imw_superglue_forward: n_layers %d", W->n_layers);
  const int P = n_pairs, S = 2 * P;
  const int use_tc = conf->use_tensor_cores;
  IMW_REQUIRE(!use_tc || cap % 128 == 0, "imw_superglue_forward: use_tensor_cores needs cap %% 128 == 0 (got %d)", cap);
  Workspace ws(workspace, workspace_bytes);
  SGBuffers b;
  sg_carve(ws, b, P, cap);
  if (ws.overflow) { imw_set_error("imw_superglue_forward: workspace too small (%zu < %zu)", workspace_bytes, ws.off); return IMW_ERR_WORKSPACE; }
  const long long sXM = (long long)cap * 512;

  sg_state_kernel<<<ceil_div(P, 128), 128, 0, st>>>(counts, b.empty, b.mcnt, b.ncnt, P);
  IMW_CHECK_LAUNCH_T("sg_state_kernel");
  sg_init_tokens_kernel<<<dim3(ceil_div(cap, 4), S), 256, 0, st>>>(desc, b.xm, matches, mscores, counts, cap);
  IMW_CHECK_LAUNCH_T("sg_init_tokens_kernel");
  // ---- keypoint encoder (superglue.py:75-84): 3 -> 32 -> 64 -> 128 -> 256 -> 256, fp32 CUDA cores (tiny)
  sg_kenc_input_kernel<<<dim3(ceil_div(cap, 256), S), 256, 0, st>>>(kpts, scores, counts, image_wh, b.kin, cap);
  IMW_CHECK_LAUNCH_T("sg_kenc_input_kernel");
  {
    const int dims[6] = {16, 32, 64, 128, 256, 256};
    const float* in = b.kin;
    float* bufs[2] = {b.t0, b.t1};
    for (int l = 0; l < 5; l++) {
      GemmArgs g{};
      g.A = in; g.strideA = (long long)cap * dims[l]; g.lda = dims[l]; g.W = W->kenc_w[l]; g.ldw = dims[l];
      g.M = cap; g.N = dims[l + 1]; g.K = dims[l]; g.Mdyn = counts; g.skip = b.empty; g.skip_shift = 1;
      if (l < 4) {
        float* out = bufs[l & 1];
        IMW_CHECK_CUDA(launch_gemm(g, S, EpiBias{out, (long long)cap * dims[l + 1], dims[l + 1], W->kenc_b[l], 1}, st));
        in = out;
      } else {  // desc = desc + kenc(...) (superglue.py:246-247)
        IMW_CHECK_CUDA(launch_gemm(g, S, EpiStore{b.xm, 512, sXM, W->kenc_b[l], 1}, st));
      }
    }
  }
  auto linear = [&](const float* A, int lda, const float* Wt, int N, int K, auto epi) -> int {
    if (use_tc) {
      TcGemmArgs t{};
      t.K = K; t.N = N; t.tiles_per_slot = cap / 128; t.counts = counts; t.skip = b.empty; t.skip_shift = 1;
      if (use_tc == 2) return launch_tc_gemm<128, 1>(A, (long long)S * cap, lda, Wt, N, t, epi, st);
      return launch_tc_gemm<128, 3>(A, (long long)S * cap, lda, Wt, N, t, epi, st);
    }
    GemmArgs g{};
    g.A = A; g.strideA = (long long)cap * lda; g.lda = lda; g.W = Wt; g.ldw = K; g.M = cap; g.N = N; g.K = K;
    g.Mdyn = counts; g.skip = b.empty; g.skip_shift = 1;
    IMW_CHECK_CUDA(launch_gemm(g, S, epi, st));
    return IMW_OK;
  };
  const long long plane = use_tc ? (long long)S * cap * D : 0;
  if (use_tc) IMW_CHECK_CUDA(cudaMemsetAsync(b.v, 0, sizeof(plane_t) * 2 * plane, st));   // V^T tail columns must be finite
  // ---- attentional GNN (superglue.py:112-140)
  for (int i = 0; i < W->n_layers; i++) {
    const imw_sg_layer& ly = W->layers[i];
    if (int e = linear(b.xm, 512, ly.qkv_w, 3 * D, D, EpiQKVRotary{b.q, b.k, b.v, ly.qkv_b, nullptr, cap, plane})) return e;
    if (use_tc) {
      TcAttnArgs a{b.ctx, counts, b.empty, cap, S, 0.125f, ly.is_cross, (long long)S * HEADS * cap, (long long)S * HEADS * HD};
      if (int e = launch_tc_attn(b.q, b.k, b.v, a, st)) return e;
    } else {
      if (int e = imw_attention_simt(b.q, b.k, b.v, b.ctx, counts, b.empty, cap, S, 0.125f, ly.is_cross, st)) return e;
    }
    if (int e = linear(b.ctx, D, ly.merge_w, D, D, EpiStore{b.xm + D, 512, sXM, ly.merge_b, 0})) return e;
    EpiStore e0{b.h, 512, sXM, ly.mlp0_b, 0};
    e0.relu = 1;
    if (int e = linear(b.xm, 512, ly.mlp0_w, 512, 512, e0)) return e;
    if (int e = linear(b.h, 512, ly.mlp1_w, D, 512, EpiStore{b.xm, 512, sXM, ly.mlp1_b, 1})) return e;
  }
  // ---- final projection + score matrix / sqrt(256) (superglue.py:252-258)
  if (int e = linear(b.xm, 512, W->final_w, D, D, EpiStore{b.md, D, (long long)cap * D, W->final_b, 0})) return e;
  {
    EpiStore es{b.S, cap, (long long)cap * cap / 2, nullptr, 0};
    es.scale = 0.0625f;
    if (use_tc && cap % 128 == 0) {
      TcGemmArgs t{};
      t.K = D; t.N = cap; t.tiles_per_slot = cap / 128; t.counts = counts; t.skip = b.empty; t.skip_shift = 1; t.pair_product = 1;
      int e = (use_tc == 2) ? launch_tc_gemm<128, 1>(b.md, (long long)S * cap, D, b.md, (long long)S * cap, t, es, st)
                            : launch_tc_gemm<128, 3>(b.md, (long long)S * cap, D, b.md, (long long)S * cap, t, es, st);
      if (e) return e;
    } else {
      GemmArgs g{};
      g.A = b.md; g.strideA = 2LL * cap * D; g.lda = D; g.W = b.md + (long long)cap * D; g.strideW = 2LL * cap * D; g.ldw = D;
      g.M = cap; g.N = cap; g.K = D; g.Mdyn = b.mcnt; g.Ndyn = b.ncnt; g.skip = b.empty; g.skip_shift = 0;
      es.strideOut = (long long)cap * cap;
      IMW_CHECK_CUDA(launch_gemm(g, P, es, st));
    }
  }
  // ---- optimal transport (superglue.py:152-172) and matches (:266-276)
  SkArgs sk{b.S, b.u, b.vv, counts, b.empty, cap, W->bin_score};
  IMW_CHECK_CUDA(cudaMemsetAsync(b.vv, 0, sizeof(float) * (size_t)P * (cap + 1), st));
  for (int it = 0; it < conf->sinkhorn_iterations; it++) {
    sk_row_kernel<<<dim3(ceil_div(cap + 1, 8), P), 256, 0, st>>>(sk);
    IMW_CHECK_LAUNCH_T("sk_row_kernel");
    sk_col_kernel<<<dim3(ceil_div(cap + 1, 32), P), 256, 0, st>>>(sk);
    IMW_CHECK_LAUNCH_T("sk_col_kernel");
  }
  if (conf->sinkhorn_iterations == 0) IMW_CHECK_CUDA(cudaMemsetAsync(b.u, 0, sizeof(float) * (size_t)P * (cap + 1), st));
  sk_rowmax_kernel<<<dim3(ceil_div(cap, 8), P), 256, 0, st>>>(sk, b.best_v, b.best_j);
  IMW_CHECK_LAUNCH_T("sk_rowmax_kernel");
  sk_colmax_kernel<<<dim3(ceil_div(cap, 32), P), 256, 0, st>>>(sk, b.best_v, b.best_j);
  IMW_CHECK_LAUNCH_T("sk_colmax_kernel");
  sg_match_kernel<<<dim3(ceil_div(cap, 256), S), 256, 0, st>>>(b.best_v, b.best_j, counts, b.empty, matches, mscores, cap,
                                                               conf->match_threshold);
  IMW_CHECK_LAUNCH_T("sg_match_kernel");
  return IMW_OK;
}
