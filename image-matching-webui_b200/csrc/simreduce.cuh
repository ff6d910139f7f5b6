// Streaming similarity reductions: for every row i of set X (own) against all rows j of set Y
// (other) compute s_ij = <x_i, y_j> tile by tile in shared memory and fold f(s_ij, i, j) into a
// per-row state -- the N x M matrix is never written to HBM.
//
// Used by: LightGlue assignment (double log-softmax statistics + mutual arg-max,
// lightglue.py:265-318), hloc NearestNeighbor (nearest_neighbor.py:6-66) and DualSoftMax
// (dual_softmax.py:8-36).  Calling it twice with the roles of X and Y swapped gives row and
// column reductions from bit-identical s_ij (the k-loop order is the same in both calls and
// fma(a,b,c) == fma(b,a,c)), which the equality-based mutual checks of the reference rely on.
#pragma once
#include "common.cuh"

struct SimArgs {
  const float* X;        // [slots][cap][ld]
  int cap, ld, K;        // K % 4 == 0, K <= 256
  const int* counts;     // [slots] valid rows per slot
  const int* skip;       // optional, indexed by pair (slot >> 1)
};

constexpr int SR_T = 64;        // rows per CTA and columns per tile
constexpr int SR_THREADS = 256; // 16 x 16 threads, 4x4 outputs each

// Op interface:
//   State                        per-thread-per-row running state
//   init(State&)
//   accum(State&, float s, int i, int j, int own_slot, int other_slot)
//   merge(State&, const State&)  combine partial states (must be order-insensitive incl. ties)
//   store(const State&, int own_slot, int i)
template <class Op>
__global__ void __launch_bounds__(SR_THREADS) simreduce_kernel(SimArgs a, Op op) {
  const int own = blockIdx.y, other = own ^ 1;
  if (a.skip && a.skip[own >> 1]) return;
  const int n = a.counts[own], m = a.counts[other];
  const int i0 = blockIdx.x * SR_T;
  if (i0 >= n) return;
  extern __shared__ __align__(16) float sr_smem[];
  const int K = a.K;
  float* Xt = sr_smem;                   // [K][SR_T+4]
  float* Yt = sr_smem + K * (SR_T + 4);  // [K][SR_T+4]
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const float* Xo = a.X + (long long)own * a.cap * a.ld;
  const float* Yo = a.X + (long long)other * a.cap * a.ld;

  const int kq = K / 4;
  for (int idx = tid; idx < SR_T * kq; idx += SR_THREADS) {
    int r = idx / kq, q = idx % kq;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i0 + r < n) v = *reinterpret_cast<const float4*>(Xo + (long long)(i0 + r) * a.ld + q * 4);
    Xt[(q * 4 + 0) * (SR_T + 4) + r] = v.x; Xt[(q * 4 + 1) * (SR_T + 4) + r] = v.y;
    Xt[(q * 4 + 2) * (SR_T + 4) + r] = v.z; Xt[(q * 4 + 3) * (SR_T + 4) + r] = v.w;
  }
  typename Op::State st[4];
#pragma unroll
  for (int r = 0; r < 4; r++) op.init(st[r]);

  for (int j0 = 0; j0 < m; j0 += SR_T) {
    __syncthreads();
    for (int idx = tid; idx < SR_T * kq; idx += SR_THREADS) {
      int r = idx / kq, q = idx % kq;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + r < m) v = *reinterpret_cast<const float4*>(Yo + (long long)(j0 + r) * a.ld + q * 4);
      Yt[(q * 4 + 0) * (SR_T + 4) + r] = v.x; Yt[(q * 4 + 1) * (SR_T + 4) + r] = v.y;
      Yt[(q * 4 + 2) * (SR_T + 4) + r] = v.z; Yt[(q * 4 + 3) * (SR_T + 4) + r] = v.w;
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int c = 0; c < 4; c++) s[r][c] = 0.f;
#pragma unroll 4
    for (int k = 0; k < K; k++) {
      float4 x = *reinterpret_cast<const float4*>(Xt + k * (SR_T + 4) + ty * 4);
      float4 y = *reinterpret_cast<const float4*>(Yt + k * (SR_T + 4) + tx * 4);
      float xv[4] = {x.x, x.y, x.z, x.w}, yv[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) s[r][c] = fmaf(xv[r], yv[c], s[r][c]);
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
      int i = i0 + ty * 4 + r;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        int j = j0 + tx * 4 + c;
        if (i < n && j < m) op.accum(st[r], s[r][c], i, j, own, other);
      }
    }
  }
  // merge the 16 column-group partials of each row (lanes tx = 0..15 of a half-warp)
#pragma unroll
  for (int r = 0; r < 4; r++) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      typename Op::State t = op.shfl_xor(st[r], o);
      op.merge(st[r], t);
    }
    int i = i0 + ty * 4 + r;
    if (tx == 0 && i < n) op.store(st[r], own, i);
  }
}

template <class Op>
static inline cudaError_t launch_simreduce(const SimArgs& a, int slots, Op op, cudaStream_t st) {
  size_t smem = (size_t)2 * a.K * (SR_T + 4) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(simreduce_kernel<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid(ceil_div(a.cap, SR_T), slots);
  if (grid.x == 0 || grid.y == 0) return cudaSuccess;
  simreduce_kernel<Op><<<grid, SR_THREADS, smem, st>>>(a, op);
  IMW_COUNT_LAUNCH(st);
  return cudaGetLastError();
}

// ---- reusable states ------------------------------------------------------------------------------
struct MaxSumState { float m, s; };   // running max and sum of exp(v - m)
struct ArgMaxState { float v; int j; };

__device__ __forceinline__ void lse_accum(MaxSumState& st, float v) {
  if (v > st.m) { st.s = st.s * expf(st.m - v) + 1.f; st.m = v; }
  else st.s += expf(v - st.m);
}
__device__ __forceinline__ void lse_merge(MaxSumState& a, const MaxSumState& b) {
  float m = fmaxf(a.m, b.m);
  if (m == -INFINITY) return;
  a.s = a.s * expf(a.m - m) + b.s * expf(b.m - m);
  a.m = m;
}
// chunked form for the tensor-core kernel: fold 32 (scaled) values at once -- one rescale per chunk, no per-element branch
__device__ __forceinline__ void lse_accum32(MaxSumState& st, const float (&v)[32], int jn, float scale) {
  float x[32], cm = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; j++) { x[j] = j < jn ? v[j] * scale : -INFINITY; cm = fmaxf(cm, x[j]); }
  if (cm > st.m) { st.s *= expf(st.m - cm); st.m = cm; }   // st.s == 0 while st.m == -inf
  float a = 0.f;
#pragma unroll
  for (int j = 0; j < 32; j++) a += expf(x[j] - st.m);
  st.s += a;
}
// first-index arg-max (torch.max(dim) semantics: lowest index among equal maxima)
__device__ __forceinline__ void argmax_accum(ArgMaxState& st, float v, int j) {
  if (v > st.v || (v == st.v && j < st.j)) { st.v = v; st.j = j; }
}

// ---- dual-softmax confidence  P = softmax(x, dim 1) * softmax(x, dim 2)  (hloc dual_softmax.py:23, LoFTR coarse_matching.py:115-119)
// pass 1: per-row softmax statistics of scale * sim along the other image: max, sum of exp, and max + log(sum)
struct OpSoftmaxStats {
  using State = MaxSumState;
  float *rmax, *rsum, *rlog; int cap; float scale;
  __device__ void init(State& s) const { s.m = -INFINITY; s.s = 0.f; }
  __device__ void accum(State& s, float v, int, int, int, int) const { lse_accum(s, v * scale); }
  __device__ void accum32(State& s, const float (&v)[32], int, int, int jn, int, int) const { lse_accum32(s, v, jn, scale); }
  __device__ State shfl_xor(const State& s, int o) const {
    State t; t.m = __shfl_xor_sync(0xffffffffu, s.m, o); t.s = __shfl_xor_sync(0xffffffffu, s.s, o); return t;
  }
  __device__ void merge(State& a, const State& b) const { lse_merge(a, b); }
  __device__ void store(const State& s, int own, int i) const {
    const long long o = (long long)own * cap + i;
    rmax[o] = s.m; rsum[o] = s.s; rlog[o] = s.m + logf(s.s);
  }
};
// pass 2: arg-max over the other image of P[i,j] = exp(x - m_i) / s_i * exp(x - m_j) / s_j.  log P = 2x - c_i - c_j with
// c = max + log(sum), so along a row the winner maximises y_j = 2 x_ij - c_j: candidates are screened with one FMA per element
// and the exact fp32 product (same operations, same order as the reference: exp, divide, multiply) is evaluated only for
// elements within a safety margin of the best y seen so far -- a handful per row instead of all of them (round 1 evaluated two
// exp + two divides + four gathers for every element that passed a looser bound: 4x the cost of the statistics pass).
// The margin (1e-3 + 4e-6 |y| in log units) exceeds every fp32 rounding in y and in P by orders of magnitude, so the exact
// arg-max (first index among equal P) is always among the evaluated candidates.
struct OpDualSoftmaxArgmax {
  struct State { float v; int j; float ylim; };
  const float *rmax, *rsum, *rlog; float* best_v; int* best_j; int cap; float scale;
  __device__ void init(State& s) const { s.v = -INFINITY; s.j = 0x7fffffff; s.ylim = -INFINITY; }
  __device__ __forceinline__ void exact(State& s, float x, float y, int i, int j, int own, int other) const {
    const long long io = (long long)own * cap + i, jo = (long long)other * cap + j;
    const long long i0 = (own & 1) ? jo : io, i1 = (own & 1) ? io : jo;   // image-0 / image-1 statistics
    // softmax over image-0 positions (statistics kept per image-1 column) times softmax over image-1 positions
    const float p1 = __fdiv_rn(expf(x - rmax[i1]), rsum[i1]), p2 = __fdiv_rn(expf(x - rmax[i0]), rsum[i0]);
    const float p = __fmul_rn(p1, p2);
    if (p > s.v || (p == s.v && j < s.j)) { s.v = p; s.j = j; }
    s.ylim = fmaxf(s.ylim, y - (1e-3f + 4e-6f * fabsf(y)));
  }
  __device__ void accum(State& s, float v, int i, int j, int own, int other) const {
    const float x = v * scale, y = 2.f * x - rlog[(long long)other * cap + j];
    if (y >= s.ylim) exact(s, x, y, i, j, own, other);
  }
  __device__ void accum32(State& s, const float (&v)[32], int i, int j0, int jn, int own, int other) const {
    // the same 32 columns for every row of the warp: eight 16-byte broadcast loads (j0 % 32 == 0 and cap % 128 == 0 on the
    // tensor-core path: aligned, and in bounds even when jn < 32)
    const float4* cl4 = reinterpret_cast<const float4*>(rlog + (long long)other * cap + j0);
    float cl[32];
#pragma unroll
    for (int q = 0; q < 8; q++) { const float4 t = __ldg(cl4 + q); cl[4 * q] = t.x; cl[4 * q + 1] = t.y; cl[4 * q + 2] = t.z; cl[4 * q + 3] = t.w; }
    float y[32], ym = -INFINITY;
    const float s2 = 2.f * scale;
#pragma unroll
    for (int j = 0; j < 32; j++) { y[j] = j < jn ? fmaf(s2, v[j], -cl[j]) : -INFINITY; ym = fmaxf(ym, y[j]); }
    if (ym < s.ylim) return;
#pragma unroll
    for (int j = 0; j < 32; j++)
      if (y[j] >= s.ylim && j < jn) exact(s, v[j] * scale, y[j], i, j0 + j, own, other);
  }
  __device__ State shfl_xor(const State& s, int o) const {
    State t; t.v = __shfl_xor_sync(0xffffffffu, s.v, o); t.j = __shfl_xor_sync(0xffffffffu, s.j, o); t.ylim = s.ylim; return t;
  }
  __device__ void merge(State& a, const State& b) const { if (b.v > a.v || (b.v == a.v && b.j < a.j)) { a.v = b.v; a.j = b.j; } }
  __device__ void store(const State& s, int own, int i) const { best_v[(long long)own * cap + i] = s.v; best_j[(long long)own * cap + i] = s.j; }
};
