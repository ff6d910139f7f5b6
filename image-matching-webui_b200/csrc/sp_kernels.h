// Internal launchers for the SuperPoint stages (not part of the C ABI; see include/imw_b200.h).
#pragma once
#include <cuda_runtime.h>

int sp_conv3x3(const float* in, const float* w, const float* bias, float* out, int B, int H, int W, int Cin, int Cout,
               int relu, int pool, cudaStream_t st);
// out (fp32 NHWC) or, when out_planes != NULL, three bf16 planes [3][B][H][W][64]
int sp_conv3x3_c1(const float* img, const float* w, const float* bias, float* out, void* out_planes, int B, int H, int W,
                  cudaStream_t st);

// logits [B][h][w][ld >= 65] (NHWC, 65 used) -> dense scores [B][8h][8w]
int sp_softmax_d2s(const float* logits, float* dense, int B, int h, int w, cudaStream_t st, int ld = 65);
// dense [B][H][W] -> nms [B][H][W] (scores kept at surviving maxima, 0 elsewhere)
int sp_nms(const float* dense, float* nms, int B, int H, int W, int radius, cudaStream_t st);
// nms -> keypoints/scores per image (row-major order, or descending score when more than max_kpts pass)
// keys: scratch [B][key_cap] u64.  kpts [B][cap][2] float (x,y), scores [B][cap], counts [B].
int sp_select(const float* nms, unsigned long long* keys, int key_cap, float* kpts, float* scores, int* counts,
              int B, int H, int W, float threshold, int border, int max_kpts, int cap, cudaStream_t st,
              const float* thr_img = nullptr /* optional per-image thresholds [B] on the device */);
size_t sp_select_key_cap(int H, int W);
// in-place L2 normalisation of rows of 256 ([cells][256])
int sp_l2norm_rows(float* x, long long rows, int C, cudaStream_t st);
// bilinear sample of dense descriptors [B][h][w][256] at kpts, then L2 norm -> desc [B][cap][256]
// fix_sampling: hloc/extractors/superpoint.py:16-30 (align_corners=False at (k + 0.5) / 8 - 0.5) instead of superpoint.py:80-92
int sp_sample_desc(const float* dense_desc, const float* kpts, const int* counts, float* desc, int B, int h, int w,
                   int cap, int C, int fix_sampling, cudaStream_t st);

// tcgen05 split-precision conv (tc_conv.cu): activations / weights as three bf16 planes
int tc_conv3x3(const void* in_planes, const void* w_planes, const float* bias, void* out, int B, int H, int W, int Cin,
               int Cout, int relu, int pool, int out_fp32, cudaStream_t st);
int tc_conv1ab_fused(const float* img, const float* w1a, const float* b1a, const void* w1b_planes, const float* b1b, void* out, int B,
                     int H, int W, int pool, cudaStream_t st);
int tc_split_planes(const float* in, void* out_planes, size_t n, cudaStream_t st);
int tc_merge_planes(const void* in_planes, float* out, size_t n, cudaStream_t st);
int tc_conv_general(const void* in_planes, const void* w_planes, const float* bias, const void* res_planes, void* out, int B,
                    int H, int W, int Cin, int Cout, int ksize, int stride, int act, int out_fp32, cudaStream_t st);
