// FP32 CUDA-core GEMM  C[M,N] = A[M,K] * W[N,K]^T  with a functor epilogue.
//
// This is the exact-fp32 baseline path (and the fallback for shapes the tcgen05 path does not
// take).  Batched over blockIdx.z with per-batch strides and optional device-side row/column
// counts and skip flags, so ragged keypoint sets and early-exited pairs cost nothing and need no
// host synchronisation.
#pragma once
#include "common.cuh"

struct GemmArgs {
  const float* A;       // [z][M][lda]
  long long strideA;    // elements between batches
  int lda;
  const float* W;       // [z][N][ldw]   (row n = output feature n)
  long long strideW;
  int ldw;
  int M, N, K;          // static upper bounds; K % 16 == 0
  const int* Mdyn;      // optional: rows valid for batch z  (<= M)
  const int* Ndyn;      // optional: cols valid for batch z  (<= N)
  const int* skip;      // optional: skip[z >> skip_shift] != 0 -> batch z does nothing
  int skip_shift;
  const int* wsel_minus1;  // optional: W += (wsel_minus1[z >> wsel_shift] - 1) * strideWsel (per-pair layer pick)
  int wsel_shift;
  long long strideWsel;
};

constexpr int GEMM_TM = 128, GEMM_TN = 64, GEMM_BK = 16, GEMM_THREADS = 256;

// Epi::operator()(z, row, col, float4 acc, ncols_valid_from_col) handles columns col..col+3.
template <class Epi>
__global__ void __launch_bounds__(GEMM_THREADS, 2) gemm_nt_kernel(GemmArgs g, Epi epi) {
  const int z = blockIdx.z;
  if (g.skip && g.skip[z >> g.skip_shift]) return;
  const int M = g.Mdyn ? g.Mdyn[z] : g.M;
  const int N = g.Ndyn ? g.Ndyn[z] : g.N;
  const int m0 = blockIdx.y * GEMM_TM, n0 = blockIdx.x * GEMM_TN;
  if (m0 >= M || n0 >= N) return;
  const float* __restrict__ A = g.A + (long long)z * g.strideA;
  const float* __restrict__ W = g.W + (long long)z * g.strideW +
                                (g.wsel_minus1 ? (long long)(g.wsel_minus1[z >> g.wsel_shift] - 1) * g.strideWsel : 0);

  __shared__ __align__(16) float As[GEMM_BK][GEMM_TM + 4];
  __shared__ __align__(16) float Ws[GEMM_BK][GEMM_TN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % 16;  // column group: cols tx*4 .. +3
  const int ty = tid / 16;  // row group: rows ty*8 .. +7
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  // loader mapping: A tile 128 rows x 4 float4 -> 512 float4, 2 per thread; W tile 64 x 4 -> 1 per thread
  const int lrow = tid / 4, lk = (tid % 4) * 4;
  for (int k0 = 0; k0 < g.K; k0 += GEMM_BK) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      int r = lrow + h * 64;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < M) v = *reinterpret_cast<const float4*>(A + (long long)(m0 + r) * g.lda + k0 + lk);
      As[lk + 0][r] = v.x; As[lk + 1][r] = v.y; As[lk + 2][r] = v.z; As[lk + 3][r] = v.w;
    }
    {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + lrow < N) v = *reinterpret_cast<const float4*>(W + (long long)(n0 + lrow) * g.ldw + k0 + lk);
      Ws[lk + 0][lrow] = v.x; Ws[lk + 1][lrow] = v.y; Ws[lk + 2][lrow] = v.z; Ws[lk + 3][lrow] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GEMM_BK; k++) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int col = n0 + tx * 4;
  if (col >= N) return;
  const int nvalid = min(4, N - col);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    int row = m0 + ty * 8 + i;
    if (row < M) epi(z, row, col, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]), nvalid);
  }
}

template <class Epi>
static inline cudaError_t launch_gemm(const GemmArgs& g, int batches, Epi epi, cudaStream_t st) {
  dim3 grid(ceil_div(g.N, GEMM_TN), ceil_div(g.M, GEMM_TM), batches);
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return cudaSuccess;
  gemm_nt_kernel<Epi><<<grid, GEMM_THREADS, 0, st>>>(g, epi);
  IMW_COUNT_LAUNCH(st);
  return cudaGetLastError();
}

// ---- common epilogues ---------------------------------------------------------------------------
// out[z][row][col] = acc + bias[col]   (optionally ReLU), arbitrary N.
struct EpiBias {
  float* out; long long strideOut; int ldo; const float* bias; int relu;
  __device__ void operator()(int z, int row, int col, float4 v, int nvalid) const {
    float r[4] = {v.x, v.y, v.z, v.w};
    float* o = out + (long long)z * strideOut + (long long)row * ldo + col;
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (j < nvalid) {
        float x = r[j] + (bias ? bias[col + j] : 0.f);
        if (relu) x = fmaxf(x, 0.f);
        o[j] = x;
      }
  }
  // tcgen05 epilogue form: one element per lane, lanes along consecutive columns of one row
  __device__ float2 prefetch(int, int, int) const { return make_float2(0.f, 0.f); }
  __device__ void elem(int z, int row, int col, float v, float2) const {
    float x = v + (bias ? bias[col] : 0.f);
    if (relu) x = fmaxf(x, 0.f);
    out[(long long)z * strideOut + (long long)row * ldo + col] = x;
  }
  __device__ bool rowwise(int, int, bool, int, const float (&)[32]) const { return false; }
};
