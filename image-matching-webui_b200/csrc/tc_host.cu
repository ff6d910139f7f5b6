// Host side of the tcgen05/TMA kernels: tensor-map construction (driver entry point fetched at run time,
// no link-time libcuda dependency) and debug entry points used by the unit tests.
#include <cudaTypedefs.h>

#include "../../include/imw_b200.h"
#include "common.cuh"
#include "gemm_simt.cuh"
#include "tc_gemm.cuh"

PFN_encodeTiled tc_get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

int tc_make_map_2d_f32(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_cols,
                       uint32_t box_rows) {
  PFN_encodeTiled fn = tc_get_encode_fn();
  if (!fn) { imw_set_error("cuTensorMapEncodeTiled not available"); return IMW_ERR_CUDA; }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * sizeof(float)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { imw_set_error("cuTensorMapEncodeTiled failed: %d (rows %llu cols %llu ld %llu)", (int)r,
                                         (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems); return IMW_ERR_CUDA; }
  return IMW_OK;
}

// 2-D row-major fp16 [rows][cols] tensor (dense rows), box = {box_cols (inner, 64 = one 128-byte swizzled row), box_rows}
int tc_make_map_2d_f16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows) {
  return tc_make_map_2d_f16_ld(map, base, rows, cols, cols, box_cols, box_rows);
}
int tc_make_map_2d_f16_ld(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_cols, uint32_t box_rows) {
  PFN_encodeTiled fn = tc_get_encode_fn();
  if (!fn) { imw_set_error("cuTensorMapEncodeTiled not available"); return IMW_ERR_CUDA; }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { imw_set_error("cuTensorMapEncodeTiled(f16) failed: %d (rows %llu cols %llu)", (int)r,
                                         (unsigned long long)rows, (unsigned long long)cols); return IMW_ERR_CUDA; }
  return IMW_OK;
}

// out[M][N] = A[M][K] W[N][K]^T + bias  on the tcgen05 path (unit test hook; M % 128 == 0, N % 128 == 0, K % 32 == 0)
extern "C" int imw_debug_gemm_tf32(const float* A, const float* W, const float* bias, float* out, int M, int N, int K,
                                   int split, cudaStream_t st) {
  IMW_REQUIRE(M % 128 == 0 && N % 128 == 0 && K % 32 == 0, "imw_debug_gemm_tf32: M%%128, N%%128, K%%32");
  TcGemmArgs g{};
  g.K = K; g.N = N; g.tiles_per_slot = M / 128;
  if (split == 4) g.wlo_rows = N;   // W is [2N][K]: the weights followed by their pre-computed lo plane
  if (split >= 5) { g.w_planes = W + (size_t)N * K; g.w_plane_rows = N; }   // W [N][K] fp32 followed by its split-fp16 planes [2][N][K]
  if (split == 6) { g.a_planes = A; g.a_plane_rows = M; g.a_plane_ld = K; }   // A holds its split-fp16 planes [2][M][K] (producer-written operand)
  if (split >= 3) return launch_tc_gemm<128, 3>(A, M, K, W, N, g, EpiBias{out, 0, N, bias, 0}, st);
  return launch_tc_gemm<128, 1>(A, M, K, W, N, g, EpiBias{out, 0, N, bias, 0}, st);
}

// same product on the CUDA-core path (reference for the unit test)
extern "C" int imw_debug_gemm_fp32(const float* A, const float* W, const float* bias, float* out, int M, int N, int K,
                                   cudaStream_t st) {
  GemmArgs g{};
  g.A = A; g.lda = K; g.W = W; g.ldw = K; g.M = M; g.N = N; g.K = K;
  IMW_CHECK_CUDA(launch_gemm(g, 1, EpiBias{out, 0, N, bias, 0}, st));
  return IMW_OK;
}

// timing ablations of the split-fp16 GEMM (tc_gemm.cuh: tc_gemm_ablate): 0 = off
extern "C" int imw_debug_set_gemm_ablate(int mode) {
  if (mode >= 0) tc_gemm_ablate() = mode;
  return tc_gemm_ablate();
}
