// LightGlue matcher on the GPU, fp32 CUDA-core baseline path.
// Follows third_party/LightGlue/lightglue/lightglue.py:31-661 for batches of independent pairs
// with ragged keypoint counts, per-pair early exit and per-image point pruning decided on the
// device (no host synchronisation inside the layer loop).
//
// Slots: image `side` of pair p is slot 2p+side.  Every per-token buffer is [slots][cap][...].
#include <math_constants.h>

#include "../../include/imw_b200.h"
#include "common.cuh"
#include "gemm_simt.cuh"
#include "lg_internal.h"
#include "simreduce.cuh"
#include "tc_gemm.cuh"
#include "tc_simreduce.cuh"
#include "tc_attn.cuh"

#include "token_epilogues.cuh"

namespace {

// ---- keypoint normalisation + learnable Fourier positional encoding -----------------------------
// lightglue.py:31-43 (size=None: 1 + max - min of the keypoints) and :68-81.
// enc[slot][tok][0..31] = cos, [32..63] = sin.
// M = 2: (x, y); M = 4: (x, y, scale, orientation) for the add_scale_ori features (sift / doghardnet, lightglue.py:500-506)
__global__ void __launch_bounds__(256) posenc_kernel(const float* __restrict__ kpts, const int* __restrict__ counts,
                                                     const float* __restrict__ Wr, float* __restrict__ enc, int cap, int M,
                                                     const float* __restrict__ scales, const float* __restrict__ oris) {
  const int z = blockIdx.x, tid = threadIdx.x, n = counts[z];
  if (n == 0) return;
  const float* kp = kpts + (long long)z * cap * 2;
  float mnx = CUDART_INF_F, mny = CUDART_INF_F, mxx = -CUDART_INF_F, mxy = -CUDART_INF_F;
  for (int i = tid; i < n; i += 256) {
    float x = kp[2 * i], y = kp[2 * i + 1];
    mnx = fminf(mnx, x); mxx = fmaxf(mxx, x); mny = fminf(mny, y); mxy = fmaxf(mxy, y);
  }
  __shared__ float red[4][8];
  mnx = -warp_max(-mnx); mny = -warp_max(-mny); mxx = warp_max(mxx); mxy = warp_max(mxy);
  if (tid % 32 == 0) { red[0][tid / 32] = mnx; red[1][tid / 32] = mny; red[2][tid / 32] = mxx; red[3][tid / 32] = mxy; }
  __syncthreads();
  mnx = red[0][0]; mny = red[1][0]; mxx = red[2][0]; mxy = red[3][0];
  for (int w = 1; w < 8; w++) { mnx = fminf(mnx, red[0][w]); mny = fminf(mny, red[1][w]); mxx = fmaxf(mxx, red[2][w]); mxy = fmaxf(mxy, red[3][w]); }
  const float sx = __fsub_rn(__fadd_rn(1.f, mxx), mnx), sy = __fsub_rn(__fadd_rn(1.f, mxy), mny);
  const float shx = sx / 2, shy = sy / 2, scale = fmaxf(sx, sy) / 2;
  for (int idx = tid; idx < n * NF; idx += 256) {
    int i = idx / NF, f = idx % NF;
    float x = __fdiv_rn(__fsub_rn(kp[2 * i], shx), scale), y = __fdiv_rn(__fsub_rn(kp[2 * i + 1], shy), scale);
    float p = fmaf(y, Wr[M * f + 1], __fmul_rn(x, Wr[M * f]));
    if (M == 4) p = fmaf(oris[(long long)z * cap + i], Wr[M * f + 3], fmaf(scales[(long long)z * cap + i], Wr[M * f + 2], p));
    float* e = enc + ((long long)z * cap + i) * 64;
    e[f] = cosf(p);
    e[32 + f] = sinf(p);
  }
}

// ---- LayerNorm(512) + exact GELU (lightglue.py:152-157): in place, or -- planes != NULL -- written as the split-fp16 operand
// planes [2][slots * cap][512] of the next linear (split_planes.cuh; plane_elems apart), so that its GEMM needs no splitter warps
__global__ void __launch_bounds__(256) ln_gelu_kernel(float* __restrict__ h, const int* __restrict__ counts,
                                                      const int* __restrict__ skip, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, int cap, plane_t* __restrict__ planes,
                                                      long long plane_elems) {
  const int z = blockIdx.y, row = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (skip[z >> 1] || row >= counts[z]) return;
  float* p = h + ((long long)z * cap + row) * 512;
  float v[16];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    float4 t = *reinterpret_cast<const float4*>(p + q * 128 + lane * 4);
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; i++) s += v[i];
  const float mean = warp_sum(s) * (1.f / 512.f);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 16; i++) { float d = v[i] - mean; ss += d * d; }
  const float rstd = 1.f / sqrtf(warp_sum(ss) * (1.f / 512.f) + 1e-5f);
#pragma unroll
  for (int q = 0; q < 4; q++) {
    int c = q * 128 + lane * 4;
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float y = (v[4 * q + i] - mean) * rstd * gamma[c + i] + beta[c + i];
      o[i] = 0.5f * y * (1.f + erff(y * 0.70710678118654752440f));
    }
    if (planes) {
      __align__(8) __half2 hi[2], lo[2];
      split2x2(o[0], o[1], hi[0], lo[0]); split2x2(o[2], o[3], hi[1], lo[1]);
      plane_t* pp = planes + ((long long)z * cap + row) * 512 + c;
      *reinterpret_cast<uint2*>(pp) = *reinterpret_cast<const uint2*>(hi);
      *reinterpret_cast<uint2*>(pp + plane_elems) = *reinterpret_cast<const uint2*>(lo);
    } else {
      *reinterpret_cast<float4*>(p + c) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ---- attention (flash-style, fp32) ------------------------------------------------------------------
// ctx[z][row][head*64 + d] = softmax_j(scale * q_row . k_j) v_j ; self: k,v of slot z, cross: of z^1.
constexpr int AT = 64, ATP = AT + 4;
__global__ void __launch_bounds__(256) attn_kernel(const float* __restrict__ Q, const float* __restrict__ Kb,
                                                   const float* __restrict__ Vb, float* __restrict__ ctx,
                                                   const int* __restrict__ counts, const int* __restrict__ skip, int cap,
                                                   float scale, int cross) {
  const int z = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * AT;
  if (skip[z >> 1]) return;
  const int nq = counts[z], zk = cross ? (z ^ 1) : z, nk = counts[zk];
  if (q0 >= nq) return;
  extern __shared__ __align__(16) float at_smem[];
  float* Qt = at_smem;            // [HD][ATP]  Qt[d][row]
  float* Kt = Qt + HD * ATP;      // [HD][ATP]  Kt[d][col]
  float* Vs = Kt + HD * ATP;      // [AT][ATP]  Vs[j][d]
  float* Pt = Vs + AT * ATP;      // [AT][ATP]  Pt[j][row]
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const float* qb = Q + (((long long)z * HEADS + head) * cap) * HD;
  const float* kb = Kb + (((long long)zk * HEADS + head) * cap) * HD;
  const float* vb = Vb + (((long long)zk * HEADS + head) * cap) * HD;
  for (int idx = tid; idx < AT * (HD / 4); idx += 256) {
    int r = idx / (HD / 4), qd = idx % (HD / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < nq) v = *reinterpret_cast<const float4*>(qb + (long long)(q0 + r) * HD + qd * 4);
    Qt[(qd * 4 + 0) * ATP + r] = v.x; Qt[(qd * 4 + 1) * ATP + r] = v.y; Qt[(qd * 4 + 2) * ATP + r] = v.z; Qt[(qd * 4 + 3) * ATP + r] = v.w;
  }
  float m[4], l[4], o[4][4];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    m[r] = -CUDART_INF_F; l[r] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; c++) o[r][c] = 0.f;
  }
  for (int k0 = 0; k0 < nk; k0 += AT) {
    __syncthreads();
    for (int idx = tid; idx < AT * (HD / 4); idx += 256) {
      int r = idx / (HD / 4), qd = idx % (HD / 4);
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (k0 + r < nk) {
        kv = *reinterpret_cast<const float4*>(kb + (long long)(k0 + r) * HD + qd * 4);
        vv = *reinterpret_cast<const float4*>(vb + (long long)(k0 + r) * HD + qd * 4);
      }
      Kt[(qd * 4 + 0) * ATP + r] = kv.x; Kt[(qd * 4 + 1) * ATP + r] = kv.y; Kt[(qd * 4 + 2) * ATP + r] = kv.z; Kt[(qd * 4 + 3) * ATP + r] = kv.w;
      *reinterpret_cast<float4*>(Vs + r * ATP + qd * 4) = vv;
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int c = 0; c < 4; c++) s[r][c] = 0.f;
#pragma unroll 8
    for (int d = 0; d < HD; d++) {
      float4 a = *reinterpret_cast<const float4*>(Qt + d * ATP + ty * 4);
      float4 b = *reinterpret_cast<const float4*>(Kt + d * ATP + tx * 4);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) s[r][c] = fmaf(av[r], bv[c], s[r][c]);
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
      float mx = -CUDART_INF_F;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        s[r][c] = (k0 + tx * 4 + c < nk) ? s[r][c] * scale : -CUDART_INF_F;
        mx = fmaxf(mx, s[r][c]);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float mn = fmaxf(m[r], mx);
      const float alpha = expf(m[r] - mn);
      float ps = 0.f;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        float p = expf(s[r][c] - mn);
        ps += p;
        Pt[(tx * 4 + c) * ATP + ty * 4 + r] = p;
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
      l[r] = l[r] * alpha + ps;
      m[r] = mn;
#pragma unroll
      for (int c = 0; c < 4; c++) o[r][c] *= alpha;
    }
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < AT; j++) {
      float4 a = *reinterpret_cast<const float4*>(Pt + j * ATP + ty * 4);
      float4 b = *reinterpret_cast<const float4*>(Vs + j * ATP + tx * 4);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) o[r][c] = fmaf(av[r], bv[c], o[r][c]);
    }
  }
#pragma unroll
  for (int r = 0; r < 4; r++) {
    int row = q0 + ty * 4 + r;
    if (row < nq) {
      float inv = (nk > 0) ? 1.f / l[r] : 0.f;  // empty key set -> zeros (lightglue.py:113-114)
      *reinterpret_cast<float4*>(ctx + ((long long)z * cap + row) * D + head * HD + tx * 4) =
          make_float4(o[r][0] * inv, o[r][1] * inv, o[r][2] * inv, o[r][3] * inv);
    }
  }
}

// ---- token confidence / matchability: sigmoid(w.x + b) per token ------------------------------------
// out[z][row]; optionally counts (value < thr) per pair into cnt[pair] (check_if_stop, lightglue.py:650-661)
__global__ void __launch_bounds__(256) token_logit_kernel(const float* __restrict__ xm, int ldx, const int* __restrict__ counts,
                                                          const int* __restrict__ skip, const float* __restrict__ w_all,
                                                          const float* __restrict__ b_all, const int* __restrict__ layer_sel,
                                                          int layer, float* __restrict__ out, int apply_sigmoid, float thr,
                                                          int* __restrict__ cnt, int cap) {
  const int z = blockIdx.y, row = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (skip[z >> 1] || row >= counts[z]) return;
  const int L = layer_sel ? layer_sel[z >> 1] - 1 : layer;
  const float* w = w_all + (long long)L * D;
  const float* x = xm + ((long long)z * cap + row) * ldx;
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 2; q++) {
    float4 a = *reinterpret_cast<const float4*>(x + q * 128 + lane * 4);
    float4 b = *reinterpret_cast<const float4*>(w + q * 128 + lane * 4);
    s += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
  }
  s = warp_sum(s) + b_all[L];
  float v = apply_sigmoid ? 1.f / (1.f + expf(-s)) : s;
  if (lane == 0) {
    out[(long long)z * cap + row] = v;
    if (cnt && v < thr) atomicAdd(cnt + (z >> 1), 1);
  }
}

// early-exit decision per pair (lightglue.py:552-555,650-661); resets the counter.
__global__ void exit_kernel(int* __restrict__ cnt, int* __restrict__ done, int* __restrict__ stop,
                            const int* __restrict__ counts0, int P, float depth_conf, int layer) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  if (!done[p]) {
    float num = (float)(counts0[2 * p] + counts0[2 * p + 1]);
    float ratio = __fsub_rn(1.0f, __fdiv_rn((float)cnt[p], num));
    if (ratio > depth_conf) { done[p] = 1; stop[p] = layer + 1; }
  }
  cnt[p] = 0;
}

// ---- point pruning (lightglue.py:556-571, 641-648): ordered compaction of x, enc, ind -------------
__global__ void __launch_bounds__(1024) prune_kernel(const float* __restrict__ xm_in, float* __restrict__ xm_out,
                                                     const float* __restrict__ enc_in, float* __restrict__ enc_out,
                                                     const int* __restrict__ ind_in, int* __restrict__ ind_out,
                                                     const float* __restrict__ conf, const float* __restrict__ mscore,
                                                     int* __restrict__ counts, const int* __restrict__ done,
                                                     int* __restrict__ prune, int cap, int pruning_th, float width_conf,
                                                     float conf_thr, int have_conf, const plane_t* __restrict__ xp_in,
                                                     plane_t* __restrict__ xp_out, long long plane_elems) {
  const int z = blockIdx.x, tid = threadIdx.x, lane = tid % 32, wid = tid / 32;
  const int n = counts[z];
  // pairs that already stopped (incl. at this layer: the reference breaks before pruning) are
  // copied through unchanged so that every pair lives in the same ping-pong buffer.
  const bool do_prune = !done[z >> 1] && n > pruning_th;
  __shared__ int s_w[32];
  __shared__ int s_base;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 1024) {
    int i = i0 + tid;
    bool keep = false;
    if (i < n) {
      keep = true;
      if (do_prune) {
        keep = mscore[(long long)z * cap + i] > (1.f - width_conf);
        if (have_conf) keep = keep || (conf[(long long)z * cap + i] <= conf_thr);
      }
    }
    unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_w[wid] = __popc(bal);
    __syncthreads();
    int woff = 0, tot = 0;
    for (int w = 0; w < 32; w++) { int c = s_w[w]; if (w < wid) woff += c; tot += c; }
    int dst = s_base + woff + __popc(bal & ((1u << lane) - 1u));
    if (keep) {
      int orig = ind_in[(long long)z * cap + i];
      ind_out[(long long)z * cap + dst] = orig;
      if (do_prune) prune[(long long)z * cap + orig] += 1;
      // keep one slot of work per thread: copy the 256 x-floats and 64 enc-floats of this token
      const float4* sx = reinterpret_cast<const float4*>(xm_in + ((long long)z * cap + i) * 512);
      float4* dx = reinterpret_cast<float4*>(xm_out + ((long long)z * cap + dst) * 512);
#pragma unroll 8
      for (int q = 0; q < 64; q++) dx[q] = sx[q];
      if (xp_in) {   // ... and its operand planes (256 halves = 32 x 16 B per plane)
#pragma unroll
        for (int pl = 0; pl < 2; pl++) {
          const uint4* sp = reinterpret_cast<const uint4*>(xp_in + pl * plane_elems + ((long long)z * cap + i) * 512);
          uint4* dp = reinterpret_cast<uint4*>(xp_out + pl * plane_elems + ((long long)z * cap + dst) * 512);
#pragma unroll 8
          for (int q = 0; q < 32; q++) dp[q] = sp[q];
        }
      }
      const float4* se = reinterpret_cast<const float4*>(enc_in + ((long long)z * cap + i) * 64);
      float4* de = reinterpret_cast<float4*>(enc_out + ((long long)z * cap + dst) * 64);
#pragma unroll
      for (int q = 0; q < 16; q++) de[q] = se[q];
    }
    __syncthreads();
    if (tid == 0) s_base += tot;
    __syncthreads();
  }
  if (tid == 0) counts[z] = s_base;
}

// ---- assignment: double log-softmax statistics and mutual arg-max -----------------------------------
// pass 1: per-row (max, log-sum-exp) of sim = <md_i, md_j> against the other image
struct OpRowLSE {
  using State = MaxSumState;
  float* rmax; float* rlse; int cap;
  __device__ void init(State& s) const { s.m = -CUDART_INF_F; s.s = 0.f; }
  __device__ void accum(State& s, float v, int, int, int, int) const { lse_accum(s, v); }
  __device__ void accum32(State& s, const float (&v)[32], int, int, int jn, int, int) const { lse_accum32(s, v, jn, 1.f); }
  __device__ State shfl_xor(const State& s, int o) const {
    State t; t.m = __shfl_xor_sync(0xffffffffu, s.m, o); t.s = __shfl_xor_sync(0xffffffffu, s.s, o); return t;
  }
  __device__ void merge(State& a, const State& b) const { lse_merge(a, b); }
  __device__ void store(const State& s, int own, int i) const {
    rmax[(long long)own * cap + i] = s.m;
    rlse[(long long)own * cap + i] = logf(s.s);
  }
};
// pass 2: scores[i,j] = (log_softmax_row + log_softmax_col) + (logsigmoid(z0_i) + logsigmoid(z1_j))
// (lightglue.py:265-277), arg-max over the other image.  Image-0 terms come first in every sum so
// that both passes produce bit-identical scores[i,j].
struct OpAssignArgmax {
  using State = ArgMaxState;
  const float *rmax, *rlse, *lsz; float* best_v; int* best_j; int cap;
  __device__ void init(State& s) const { s.v = -CUDART_INF_F; s.j = 0x7fffffff; }
  __device__ void accum(State& s, float v, int i, int j, int own, int other) const {
    long long io = (long long)own * cap + i, jo = (long long)other * cap + j;
    float t_own = __fsub_rn(__fsub_rn(v, rmax[io]), rlse[io]);
    float t_oth = __fsub_rn(__fsub_rn(v, rmax[jo]), rlse[jo]);
    float sc;
    if ((own & 1) == 0) sc = __fadd_rn(__fadd_rn(t_own, t_oth), __fadd_rn(lsz[io], lsz[jo]));
    else sc = __fadd_rn(__fadd_rn(t_oth, t_own), __fadd_rn(lsz[jo], lsz[io]));
    argmax_accum(s, sc, j);
  }
  __device__ State shfl_xor(const State& s, int o) const {
    State t; t.v = __shfl_xor_sync(0xffffffffu, s.v, o); t.j = __shfl_xor_sync(0xffffffffu, s.j, o); return t;
  }
  __device__ void merge(State& a, const State& b) const { argmax_accum(a, b.v, b.j); }
  __device__ void store(const State& s, int own, int i) const {
    best_v[(long long)own * cap + i] = s.v;
    best_j[(long long)own * cap + i] = s.j;
  }
};

__global__ void logsigmoid_kernel(float* __restrict__ zl, const int* __restrict__ counts, const int* __restrict__ skip, int cap) {
  const int z = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (skip[z >> 1] || i >= counts[z]) return;
  float x = zl[(long long)z * cap + i];
  // F.logsigmoid(x) = min(x,0) - log1p(exp(-|x|))
  zl[(long long)z * cap + i] = fminf(x, 0.f) - log1pf(expf(-fabsf(x)));
}

// filter_matches (lightglue.py:302-318) + scatter back through the pruning indices (:610-619)
__global__ void __launch_bounds__(256) match_kernel(const float* __restrict__ best_v, const int* __restrict__ best_j,
                                                    const int* __restrict__ counts, const int* __restrict__ ind,
                                                    const int* __restrict__ empty, int* __restrict__ matches,
                                                    float* __restrict__ mscores, int cap, float th) {
  const int z = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x, zo = z ^ 1;
  if (empty[z >> 1] || i >= counts[z]) return;
  const long long io = (long long)z * cap + i;
  const int jr = best_j[io];
  const bool j_ok = (unsigned)jr < (unsigned)counts[zo];   // NaN rows leave the arg-max at its init value: unmatched
  const int j = j_ok ? jr : 0;
  const long long jo = (long long)zo * cap + j;
  const bool mutual = j_ok && best_j[jo] == i;
  // mscores0 = exp(max0) where mutual; mscores1[j] = mscores0[m1[j]] where mutual
  const float sc0 = (z & 1) ? expf(best_v[jo]) : expf(best_v[io]);
  const float ms = mutual ? sc0 : 0.f;
  const bool valid = mutual && (ms > th);
  const int oi = ind[io];
  matches[(long long)z * cap + oi] = valid ? ind[jo] : -1;
  mscores[(long long)z * cap + oi] = ms;
}

__global__ void init_state_kernel(const int* __restrict__ counts_in, int* __restrict__ counts, int* __restrict__ counts0,
                                  int* __restrict__ done, int* __restrict__ empty, int* __restrict__ stop, int* __restrict__ cnt,
                                  int P, int L) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  int m = counts_in[2 * p], n = counts_in[2 * p + 1];
  counts[2 * p] = counts0[2 * p] = m;
  counts[2 * p + 1] = counts0[2 * p + 1] = n;
  bool e = (m == 0 || n == 0);
  done[p] = e ? 1 : 0;
  empty[p] = e ? 1 : 0;
  stop[p] = e ? 1 : L;  // "no keypoints" breaks at i=0 -> stop = 1 (lightglue.py:540-541,573-593)
  cnt[p] = 0;
}

// x <- descriptors (input_dim == 256: Identity input_proj), ind <- arange, outputs <- -1 / 0 / prune init
__global__ void __launch_bounds__(256) init_tokens_kernel(const float* __restrict__ desc, float* __restrict__ xm,
                                                          int* __restrict__ ind, int* __restrict__ matches,
                                                          float* __restrict__ mscores, int* __restrict__ prune,
                                                          const int* __restrict__ counts, int cap, int prune_init,
                                                          plane_t* __restrict__ xp, long long plane_elems) {
  const int z = blockIdx.y, row = blockIdx.x * 4 + threadIdx.x / 64, t = threadIdx.x % 64;
  if (row >= cap) return;
  if (t == 0) {
    ind[(long long)z * cap + row] = row;
    matches[(long long)z * cap + row] = -1;
    mscores[(long long)z * cap + row] = 0.f;
    prune[(long long)z * cap + row] = prune_init;
  }
  if (desc && row < counts[z]) {
    const float4 v = reinterpret_cast<const float4*>(desc + ((long long)z * cap + row) * D)[t];
    reinterpret_cast<float4*>(xm + ((long long)z * cap + row) * 512)[t] = v;
    if (xp) {   // the same values as operand planes of the first projection
      __align__(8) __half2 hi[2], lo[2];
      split2x2(v.x, v.y, hi[0], lo[0]); split2x2(v.z, v.w, hi[1], lo[1]);
      plane_t* pp = xp + ((long long)z * cap + row) * 512 + 4 * t;
      *reinterpret_cast<uint2*>(pp) = *reinterpret_cast<const uint2*>(hi);
      *reinterpret_cast<uint2*>(pp + plane_elems) = *reinterpret_cast<const uint2*>(lo);
    }
  }
}

}  // namespace

int imw_attention_simt(const float* q, const float* k, const float* v, float* ctx, const int* counts, const int* skip, int cap,
                       int slots, float scale, int cross, cudaStream_t st) {
  const size_t at_smem = (size_t)4 * 64 * ATP * sizeof(float);
  IMW_CHECK_CUDA(cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)at_smem));
  attn_kernel<<<dim3(ceil_div(cap, AT), HEADS, slots), 256, at_smem, st>>>(q, k, v, ctx, counts, skip, cap, scale, cross);
  IMW_CHECK_LAUNCH_T("attn_kernel");
  return IMW_OK;
}

// =====================================================================================================
struct LGBuffers {
  float *xm[2], *enc[2]; int* ind[2];
  float *q, *k, *v, *ctx, *h, *conf, *mscore, *md, *zl, *rmax, *rlse, *best_v;
  plane_t* hp;     // LayerNorm + GELU output as split-fp16 planes [2][slots * cap][512] (tensor-core path)
  plane_t* xp[2];  // the token state [x | message] as split-fp16 planes, written next to xm by every producer of xm (tensor-core path)
  int *best_j, *counts, *counts0, *done, *empty, *cnt;
};

static size_t lg_carve(Workspace& ws, LGBuffers& b, int P, int cap) {
  const size_t S = 2 * (size_t)P, T = S * cap;
  for (int i = 0; i < 2; i++) { b.xm[i] = ws.take<float>(T * 512); b.enc[i] = ws.take<float>(T * 64); b.ind[i] = ws.take<int>(T); }
  // x2: hi/lo planes for the tcgen05 attention operands (the CUDA-core path uses the first plane only)
  b.q = ws.take<float>(2 * T * D); b.k = ws.take<float>(2 * T * D); b.v = ws.take<float>(2 * T * D);
  b.ctx = ws.take<float>(T * D); b.h = ws.take<float>(T * 512); b.hp = ws.take<plane_t>(2 * T * 512);
  for (int i = 0; i < 2; i++) b.xp[i] = ws.take<plane_t>(2 * T * 512);
  b.conf = ws.take<float>(T); b.mscore = ws.take<float>(T); b.md = ws.take<float>(T * D); b.zl = ws.take<float>(T);
  b.rmax = ws.take<float>(T); b.rlse = ws.take<float>(T); b.best_v = ws.take<float>(T); b.best_j = ws.take<int>(T);
  b.counts = ws.take<int>(S); b.counts0 = ws.take<int>(S); b.done = ws.take<int>(P); b.empty = ws.take<int>(P); b.cnt = ws.take<int>(P);
  return ws.off;
}

extern "C" size_t imw_lightglue_workspace_bytes(int n_pairs, int cap) {
  Workspace ws(nullptr, 0);
  LGBuffers b;
  return lg_carve(ws, b, n_pairs, cap) + 256;
}

static float lg_conf_threshold(int i, int L) {  // lightglue.py:636-639
  double t = 0.8 + 0.1 * exp(-4.0 * i / L);
  return (float)(t < 0 ? 0 : (t > 1 ? 1 : t));
}

extern "C" int imw_lightglue_forward(const imw_lg_weights* W, const imw_lg_conf* conf, int n_pairs, int cap,
                                     const float* kpts, const float* desc, const int* counts_in, int* matches,
                                     float* mscores, int* stop, int* prune, void* workspace, size_t workspace_bytes,
                                     cudaStream_t st) {
  return imw_lightglue_forward_so(W, conf, n_pairs, cap, kpts, nullptr, nullptr, desc, counts_in, matches, mscores, stop, prune, workspace,
                                  workspace_bytes, st);
}

extern "C" int imw_lightglue_forward_so(const imw_lg_weights* W, const imw_lg_conf* conf, int n_pairs, int cap,
                                        const float* kpts, const float* scales, const float* oris, const float* desc,
                                        const int* counts_in, int* matches, float* mscores, int* stop, int* prune, void* workspace,
                                        size_t workspace_bytes, cudaStream_t st) {
  IMW_REQUIRE(W && conf && n_pairs > 0 && cap > 0, "imw_lightglue_forward: bad arguments");
  const int pe_dim = W->posenc_dim == 4 ? 4 : 2;
  IMW_REQUIRE(pe_dim == 2 || (scales && oris), "imw_lightglue_forward: add_scale_ori weights (posenc_dim 4) need the scales / oris inputs");
  IMW_REQUIRE(W->n_layers >= 1 && W->n_layers <= IMW_LG_MAX_LAYERS, "imw_lightglue_forward: n_layers %d", W->n_layers);
  IMW_REQUIRE(W->input_dim == 256 || (W->input_dim == 128 && W->input_proj_w && W->input_proj_b),
              "imw_lightglue_forward: input_dim must be 256 (identity) or 128 with input_proj weights (got %d)", W->input_dim);
  IMW_REQUIRE(cap % 4 == 0, "imw_lightglue_forward: cap %% 4");
  const int P = n_pairs, S = 2 * P, L = W->n_layers;
  Workspace ws(workspace, workspace_bytes);
  LGBuffers b;
  lg_carve(ws, b, P, cap);
  if (ws.overflow) { imw_set_error("imw_lightglue_forward: workspace too small (%zu < %zu)", workspace_bytes, ws.off); return IMW_ERR_WORKSPACE; }

  const bool do_stop = conf->depth_confidence > 0.f;
  const bool do_prune = conf->width_confidence > 0.f && cap > conf->pruning_min_kpts;  // counts <= cap
  const bool prune_semantics = conf->width_confidence > 0.f;
  int cur = 0;

  init_state_kernel<<<ceil_div(P, 128), 128, 0, st>>>(counts_in, b.counts, b.counts0, b.done, b.empty, stop, b.cnt, P, L);
  IMW_CHECK_LAUNCH_T("init_state_kernel");
  // prune output: 1 (+1 per surviving pruning step) when pruning is enabled, n_layers otherwise (:617-619)
  // tensor-core path with host-packed weight planes: every producer of the token state / the MLP hidden layer also writes it as
  // split-fp16 operand planes, and the linears take those by TMA (no splitter warps; the split is done once, not per N tile)
  const bool planes_on = conf->use_tensor_cores == 1 && W->has_lo_planes == 2;
  const long long xp_elems = (long long)S * cap * 512;
  init_tokens_kernel<<<dim3(ceil_div(cap, 4), S), 256, 0, st>>>(W->input_dim == D ? desc : nullptr, b.xm[0], b.ind[0], matches, mscores, prune, b.counts, cap,
                                                                prune_semantics ? 1 : L, planes_on ? b.xp[0] : nullptr, xp_elems);
  IMW_CHECK_LAUNCH_T("init_tokens_kernel");
  posenc_kernel<<<S, 256, 0, st>>>(kpts, b.counts, W->posenc_wr, b.enc[0], cap, pe_dim, scales, oris);
  IMW_CHECK_LAUNCH_T("posenc_kernel");

  const size_t at_smem = (size_t)4 * 64 * ATP * sizeof(float);
  IMW_CHECK_CUDA(cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)at_smem));
  const long long sXM = (long long)cap * 512, sD = (long long)cap * D;
  const dim3 rows8(ceil_div(cap, 8), S);

  const int use_tc = conf->use_tensor_cores;
  IMW_REQUIRE(!use_tc || cap % 128 == 0, "imw_lightglue_forward: use_tensor_cores needs cap %% 128 == 0 (got %d)", cap);
  // Y = X W^T (+ functor epilogue) over all slots: tcgen05 TF32 tiles or the exact-fp32 CUDA-core kernel
  auto linear = [&](const float* A, int lda, const float* Wt, long long w_rows, int N, int K, auto epi, const int* skip,
                    const int* wsel, const plane_t* a_planes = nullptr) -> int {
    if (use_tc) {
      TcGemmArgs t{};
      if (a_planes) { t.a_planes = a_planes; t.a_plane_rows = (long long)S * cap; t.a_plane_ld = lda; }
      t.K = K; t.N = N; t.tiles_per_slot = cap / 128; t.counts = b.counts; t.skip = skip; t.skip_shift = 1;
      t.wsel_minus1 = wsel; t.wsel_shift = 1; t.wsel_rows = N;
      if (W->has_lo_planes == 2) { t.w_planes = Wt + (size_t)w_rows * K; t.w_plane_rows = w_rows; }   // [W fp32 ; split-fp16 planes]
      else t.wlo_rows = W->has_lo_planes ? w_rows : 0;     // host pre-split weights: [W ; W - trunc_tf32(W)]
      if (use_tc == 2) return launch_tc_gemm<128, 1>(A, (long long)S * cap, lda, Wt, w_rows, t, epi, st);
      return launch_tc_gemm<128, 3>(A, (long long)S * cap, lda, Wt, w_rows, t, epi, st);
    }
    GemmArgs g{};
    g.A = A; g.strideA = (long long)cap * lda; g.lda = lda; g.W = Wt; g.strideW = 0; g.ldw = K; g.M = cap; g.N = N; g.K = K;
    g.Mdyn = b.counts; g.skip = skip; g.skip_shift = 1; g.wsel_minus1 = wsel; g.wsel_shift = 1; g.strideWsel = (long long)N * K;
    IMW_CHECK_CUDA(launch_gemm(g, S, epi, st));
    return IMW_OK;
  };
  // attention: tcgen05 flash attention on hi/lo planes (tensor-core modes) or the fp32 CUDA-core kernel
  const long long plane = use_tc ? (long long)S * cap * D : 0;
  if (use_tc) IMW_CHECK_CUDA(cudaMemsetAsync(b.v, 0, sizeof(plane_t) * 2 * plane, st));  // V^T tail columns must be finite
  auto attention = [&](const float* q, const float* k, const float* v, float scale, int cross) -> int {
    if (use_tc) {
      TcAttnArgs a{b.ctx, b.counts, b.done, cap, S, scale, cross, (long long)S * HEADS * cap, (long long)S * HEADS * HD};
      return launch_tc_attn(q, k, v, a, st);
    }
    attn_kernel<<<dim3(ceil_div(cap, AT), HEADS, S), 256, at_smem, st>>>(q, k, v, b.ctx, b.counts, b.done, cap, scale, cross);
    IMW_CHECK_LAUNCH_T("attn_kernel");
    return IMW_OK;
  };
  auto gemm = [&](const float* A, long long sA, int lda, const float* Wt, int N, int K) {
    GemmArgs g{};
    g.A = A; g.strideA = sA; g.lda = lda; g.W = Wt; g.strideW = 0; g.ldw = K; g.M = cap; g.N = N; g.K = K;
    g.Mdyn = b.counts; g.Ndyn = nullptr; g.skip = b.done; g.skip_shift = 1;
    return g;
  };
  // EpiStore into the token state (+ its planes when the plane path is on): col0 = 0 (x, residual) or D (message)
  auto store_xm = [&](int col0, const float* bias, int residual) {
    EpiStore e{b.xm[cur] + col0, 512, sXM, bias, residual};
    if (planes_on) { e.planes = b.xp[cur] + col0; e.plane_elems = xp_elems; }
    return e;
  };
  auto xplanes = [&]() -> const plane_t* { return planes_on ? b.xp[cur] : nullptr; };
  auto ffn = [&](const imw_lg_block& blk) -> int {
    float* xm = b.xm[cur];
    if (int e = linear(xm, 512, blk.ffn0_w, 512, 512, 512, EpiStore{b.h, 512, sXM, blk.ffn0_b, 0}, b.done, nullptr, xplanes())) return e;
    // with host-packed weight planes the second linear takes its activations as planes written right here (no splitter warps)
    const plane_t* hp = planes_on ? b.hp : nullptr;
    ln_gelu_kernel<<<rows8, 256, 0, st>>>(b.h, b.counts, b.done, blk.ln_g, blk.ln_b, cap, const_cast<plane_t*>(hp), (long long)S * cap * 512);
    IMW_CHECK_LAUNCH_T("ln_gelu_kernel");
    if (int e = linear(b.h, 512, blk.ffn3_w, D, D, 512, store_xm(0, blk.ffn3_b, 1), b.done, nullptr, hp)) return e;
    return IMW_OK;
  };

  if (W->input_dim != D) {  // x = input_proj(desc) (lightglue.py:519-520)
    if (int e = linear(desc, W->input_dim, W->input_proj_w, D, D, W->input_dim, store_xm(0, W->input_proj_b, 0), b.done, nullptr)) return e;
  }
  for (int i = 0; i < L; i++) {
    const imw_lg_layer& ly = W->layers[i];
    float* xm = b.xm[cur];
    float* enc = b.enc[cur];
    // ---- self attention (lightglue.py:159-172)
    if (int e = linear(xm, 512, ly.self_blk.qkv_w, 3 * D, 3 * D, D, EpiQKVRotary{b.q, b.k, b.v, ly.self_blk.qkv_b, enc, cap, plane}, b.done, nullptr, xplanes())) return e;
    if (int e = attention(b.q, b.k, b.v, 0.125f, 0)) return e;
    if (int e = linear(b.ctx, D, ly.self_blk.out_w, D, D, D, store_xm(D, ly.self_blk.out_b, 0), b.done, nullptr)) return e;
    if (int e = ffn(ly.self_blk)) return e;
    // ---- cross attention (lightglue.py:199-230)
    if (int e = linear(xm, 512, ly.cross_blk.qkv_w, 2 * D, 2 * D, D, EpiCrossQKV{b.q, b.v, ly.cross_blk.qkv_b, cap, 0.35355339059327373f, plane}, b.done, nullptr, xplanes())) return e;
    if (int e = attention(b.q, b.q, b.v, 1.0f, 1)) return e;
    if (int e = linear(b.ctx, D, ly.cross_blk.out_w, D, D, D, store_xm(D, ly.cross_blk.out_b, 0), b.done, nullptr)) return e;
    if (int e = ffn(ly.cross_blk)) return e;
    if (i == L - 1) break;
    // ---- early stop / pruning (lightglue.py:549-571)
    const float thr = lg_conf_threshold(i, L);
    if (do_stop) {
      token_logit_kernel<<<rows8, 256, 0, st>>>(xm, 512, b.counts, b.done, W->token_w, W->token_b, nullptr, i, b.conf, 1, thr, b.cnt, cap);
      IMW_CHECK_LAUNCH_T("token_logit_kernel");
      exit_kernel<<<ceil_div(P, 128), 128, 0, st>>>(b.cnt, b.done, stop, b.counts0, P, conf->depth_confidence, i);
      IMW_CHECK_LAUNCH_T("exit_kernel");
    }
    if (do_prune) {
      token_logit_kernel<<<rows8, 256, 0, st>>>(xm, 512, b.counts, b.done, W->match_w, W->match_b, nullptr, i, b.mscore, 1, 0.f, nullptr, cap);
      IMW_CHECK_LAUNCH_T("token_logit_kernel");
      prune_kernel<<<S, 1024, 0, st>>>(b.xm[cur], b.xm[cur ^ 1], b.enc[cur], b.enc[cur ^ 1], b.ind[cur], b.ind[cur ^ 1], b.conf,
                                       b.mscore, b.counts, b.done, prune, cap, conf->pruning_min_kpts, conf->width_confidence, thr,
                                       do_stop ? 1 : 0, planes_on ? b.xp[cur] : nullptr, planes_on ? b.xp[cur ^ 1] : nullptr, xp_elems);
      IMW_CHECK_LAUNCH_T("prune_kernel");
      cur ^= 1;
    }
  }
  // ---- assignment with the weights of the layer each pair stopped at (lightglue.py:595-597)
  {
    float* xm = b.xm[cur];
    // the assignment (mutual arg-max on near-equal scores) stays on the exact-fp32 path
    GemmArgs g = gemm(xm, sXM, 512, W->final_w, D, D);
    g.skip = b.empty; g.wsel_minus1 = stop; g.wsel_shift = 1; g.strideWsel = (long long)D * D;
    IMW_CHECK_CUDA(launch_gemm(g, S, EpiFinalProj{b.md, cap, W->final_b, stop}, st));
    token_logit_kernel<<<rows8, 256, 0, st>>>(xm, 512, b.counts, b.empty, W->match_w, W->match_b, stop, 0, b.zl, 0, 0.f, nullptr, cap);
    IMW_CHECK_LAUNCH_T("token_logit_kernel");
    logsigmoid_kernel<<<dim3(ceil_div(cap, 256), S), 256, 0, st>>>(b.zl, b.counts, b.empty, cap);
    IMW_CHECK_LAUNCH_T("logsigmoid_kernel");
    SimArgs sa{b.md, cap, D, D, b.counts, b.empty};
    if (use_tc && tc_simreduce_ok(sa)) {  // split-fp16 similarity tiles in TMEM, same reduction functors
      plane_t* planes = reinterpret_cast<plane_t*>(b.q);   // q/k/v are dead after the last layer: [2][S * cap][D] fp16 fits in q
      if (int e = tc_simreduce_split(sa, S, planes, st)) return e;
      if (int e = launch_tc_simreduce(sa, S, planes, OpRowLSE{b.rmax, b.rlse, cap}, st)) return e;
      if (int e = launch_tc_simreduce(sa, S, planes, OpAssignArgmax{b.rmax, b.rlse, b.zl, b.best_v, b.best_j, cap}, st)) return e;
    } else {
      IMW_CHECK_CUDA(launch_simreduce(sa, S, OpRowLSE{b.rmax, b.rlse, cap}, st));
      IMW_CHECK_CUDA(launch_simreduce(sa, S, OpAssignArgmax{b.rmax, b.rlse, b.zl, b.best_v, b.best_j, cap}, st));
    }
    match_kernel<<<dim3(ceil_div(cap, 256), S), 256, 0, st>>>(b.best_v, b.best_j, b.counts, b.ind[cur], b.empty, matches, mscores,
                                                              cap, conf->filter_threshold);
    IMW_CHECK_LAUNCH_T("match_kernel");
  }
  return IMW_OK;
}

// =====================================================================================================
// unit-test hook: attention on standard-layout q/k/v [slots][4][cap][64] through the CUDA-core kernel
// (use_tc = 0) or the tcgen05 kernel (use_tc = 1: operands are split to hi/lo planes and V transposed here).
namespace {
__global__ void attn_prep_planes_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                        float* __restrict__ qp, float* __restrict__ kp, float* __restrict__ vtp, long long n,
                                        int cap) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  store_planes(qp, n, i, q[i]);
  store_planes(kp, n, i, k[i]);
  const int d = (int)(i % HD);
  const long long row = (i / HD) % cap, zh = i / ((long long)HD * cap);
  store_planes(vtp, n, (zh * HD + d) * cap + row, v[i]);
}
}  // namespace

extern "C" int imw_debug_attention(const float* q, const float* k, const float* v, const int* counts, int slots, int cap,
                                   float scale, int cross, int use_tc, float* ctx, void* scratch, size_t scratch_bytes,
                                   cudaStream_t st) {
  IMW_REQUIRE(slots % 2 == 0 && cap % 128 == 0, "imw_debug_attention: slots even, cap %% 128");
  Workspace ws(scratch, scratch_bytes);
  const long long n = (long long)slots * HEADS * cap * HD;
  int* skip = ws.take<int>(slots / 2);
  float* qp = ws.take<float>(2 * n); float* kp = ws.take<float>(2 * n); float* vtp = ws.take<float>(2 * n);
  if (ws.overflow) { imw_set_error("imw_debug_attention: scratch too small (%zu needed)", ws.off); return IMW_ERR_WORKSPACE; }
  IMW_CHECK_CUDA(cudaMemsetAsync(skip, 0, sizeof(int) * (slots / 2), st));
  if (!use_tc) {
    const size_t at_smem = (size_t)4 * 64 * ATP * sizeof(float);
    IMW_CHECK_CUDA(cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)at_smem));
    attn_kernel<<<dim3(ceil_div(cap, AT), HEADS, slots), 256, at_smem, st>>>(q, k, v, ctx, counts, skip, cap, scale, cross);
    IMW_CHECK_LAUNCH_T("attn_kernel");
    return IMW_OK;
  }
  attn_prep_planes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(q, k, v, qp, kp, vtp, n, cap);
  IMW_CHECK_LAUNCH_T("attn_prep_planes_kernel");
  TcAttnArgs a{ctx, counts, skip, cap, slots, scale, cross, (long long)slots * HEADS * cap, (long long)slots * HEADS * HD};
  return launch_tc_attn(qp, kp, vtp, a, st);
}
