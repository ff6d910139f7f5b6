// Internal launchers shared between the matcher translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

// fp32 CUDA-core attention on standard-layout q/k/v [slots][4][cap][64] -> ctx [slots][cap][256]
int imw_attention_simt(const float* q, const float* k, const float* v, float* ctx, const int* counts, const int* skip, int cap,
                       int slots, float scale, int cross, cudaStream_t st);
