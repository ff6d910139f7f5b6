// ALIKED (aliked-n16) keypoint extractor on the GPU.
// Follows third_party/LightGlue/lightglue/aliked.py:
//   :709-740 extract_dense_map (replicate pad to x32, ConvBlock / ResBlocks with SELU, avg-pools 2/4/4, four 1x1 heads,
//            bilinear x2/x8/x32 upsampling (align_corners), concat -> score head + L2-normalised 128-channel map)
//   :291-349 DeformableConv2d == torchvision.ops.deform_conv2d (3x3, one offset group, no mask)
//   :94-261  DKD (simple_nms r, border, threshold / top-k, 5x5 soft-argmax at T = 0.1, bilinear score)
//   :479-609 SDDH (3x3 patch -> 16 offsets -> bilinear samples -> 1x1 conv + SELU -> per-position aggregation -> L2)
// Everything is exact fp32 on CUDA cores: the network has 16..128 channels and ~5.5 GFLOP per 480x640 image, i.e. it is
// bound by HBM traffic of the full-resolution maps, not by contraction throughput.  Activations are NHWC; BatchNorm is
// folded into the weights by the host (ops.aliked_pack_weights); weights are [tap][Cin][Cout].
#include <math_constants.h>

#include "../../include/imw_b200.h"
#include "common.cuh"
#include "sp_kernels.h"

namespace {

__device__ __forceinline__ float ak_selu(float x) {
  const float scale = 1.0507009873554804934193349852946f, alpha = 1.6732632423543772848170429916717f;
  return x > 0.f ? scale * x : (alpha * scale) * expm1f(x);
}
__device__ __forceinline__ float ak_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- input: planar [B][C][H][W] (C = 1 or 3) -> NHWC4 [B][Hp][Wp][4], replicate padding (InputPadder :264-288),
// gray -> RGB by repetition (grayscale_to_rgb, :759-760) -------------------------------------------------------------------
__global__ void __launch_bounds__(256) ak_pad_image_kernel(const float* __restrict__ img, float* __restrict__ out, int B, int C, int H,
                                                           int W, int Hp, int Wp, int pt, int pl) {
  const long long n = (long long)B * Hp * Wp, i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int x = (int)(i % Wp), y = (int)((i / Wp) % Hp), b = (int)(i / ((long long)Wp * Hp));
  const int sy = min(max(y - pt, 0), H - 1), sx = min(max(x - pl, 0), W - 1);
  const float* p = img + ((long long)b * C * H + sy) * W + sx;
  float4 v;
  v.x = p[0];
  v.y = C == 3 ? p[(long long)H * W] : v.x;
  v.z = C == 3 ? p[2ll * H * W] : v.x;
  v.w = 0.f;
  reinterpret_cast<float4*>(out)[i] = v;
}

// ---- k x k average pooling, NHWC ----------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256) ak_avgpool_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int Ho, int Wo, int C) {
  const long long n = (long long)B * Ho * Wo * C, i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % C);
  long long t = i / C;
  const int x = (int)(t % Wo); t /= Wo;
  const int y = (int)(t % Ho), b = (int)(t / Ho);
  const float* p = in + (((long long)b * Ho * K + (long long)y * K) * (Wo * K) + (long long)x * K) * C + c;
  float s = 0.f;
#pragma unroll
  for (int dy = 0; dy < K; dy++)
#pragma unroll
    for (int dx = 0; dx < K; dx++) s += p[((long long)dy * Wo * K + dx) * C];
  out[i] = s * (1.f / (K * K));
}

// ---- generic small-channel 3x3 convolution (zero pad 1), NHWC, Cout <= 32 ---------------------------------------------------
struct AkConvArgs {
  const float* in; int B, H, W, CIN;          // input map
  const float* w; const float* bias;          // [9][CIN][CO], [CO] (nullable)
  const float* res; int CRES;                 // optional 1x1 shortcut input [B][H][W][CRES]
  const float* ds_w; const float* ds_b;       // [CRES][CO], [CO]
  float* out; int out_ld; int cout;           // stores the first `cout` channels
  int act; float maxoff;                      // 0 none, 1 SELU, 2 clamp(+-maxoff), 3 sigmoid
  int oy0, ox0, Ho, Wo;                       // output window (crop): out[b][y - oy0][x - ox0]
};

template <int CO, int CC>
__global__ void __launch_bounds__(256) ak_conv3x3_kernel(const AkConvArgs a) {
  __shared__ __align__(16) float s_in[CC][10][35];
  __shared__ __align__(16) float s_w[9][CC][CO];
  const int tiles_x = (a.W + 31) / 32;
  const int x0 = (blockIdx.x % tiles_x) * 32, y0 = (blockIdx.x / tiles_x) * 8, b = blockIdx.y, tid = threadIdx.x;
  const int tx = tid % 32, ty = tid / 32;
  float acc[CO];
#pragma unroll
  for (int i = 0; i < CO; i++) acc[i] = 0.f;
  const float* inb = a.in + (long long)b * a.H * a.W * a.CIN;
  for (int c0 = 0; c0 < a.CIN; c0 += CC) {
    __syncthreads();
    for (int i = tid; i < 10 * 34 * CC; i += 256) {
      const int c = i % CC, xx = (i / CC) % 34, yy = i / (CC * 34);
      const int gy = y0 - 1 + yy, gx = x0 - 1 + xx;
      s_in[c][yy][xx] = (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W && c0 + c < a.CIN) ? inb[((long long)gy * a.W + gx) * a.CIN + c0 + c] : 0.f;
    }
    for (int i = tid; i < 9 * CC * CO; i += 256) {
      const int co = i % CO, c = (i / CO) % CC, t = i / (CO * CC);
      s_w[t][c][co] = (c0 + c < a.CIN) ? a.w[((long long)t * a.CIN + c0 + c) * CO + co] : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int c = 0; c < CC; c++)
#pragma unroll
      for (int t = 0; t < 9; t++) {
        const float v = s_in[c][ty + t / 3][tx + t % 3];
        if constexpr (CO % 4 == 0) {
#pragma unroll
          for (int q = 0; q < CO / 4; q++) {
            const float4 w4 = *reinterpret_cast<const float4*>(&s_w[t][c][4 * q]);
            acc[4 * q] = fmaf(v, w4.x, acc[4 * q]); acc[4 * q + 1] = fmaf(v, w4.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(v, w4.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(v, w4.w, acc[4 * q + 3]);
          }
        } else {
#pragma unroll
          for (int q = 0; q < CO; q++) acc[q] = fmaf(v, s_w[t][c][q], acc[q]);
        }
      }
  }
  const int y = y0 + ty, x = x0 + tx;
  if (y >= a.H || x >= a.W) return;
  if (a.bias) {
#pragma unroll
    for (int q = 0; q < CO; q++) acc[q] += a.bias[q];
  }
  if (a.res) {
    const float* r = a.res + (((long long)b * a.H + y) * a.W + x) * a.CRES;
    for (int c = 0; c < a.CRES; c++) {
      const float v = r[c];
#pragma unroll
      for (int q = 0; q < CO; q++) acc[q] = fmaf(v, __ldg(a.ds_w + c * CO + q), acc[q]);
    }
#pragma unroll
    for (int q = 0; q < CO; q++) acc[q] += a.ds_b[q];
  }
  const int oy = y - a.oy0, ox = x - a.ox0;
  if (oy < 0 || oy >= a.Ho || ox < 0 || ox >= a.Wo) return;
  float* o = a.out + (((long long)b * a.Ho + oy) * a.Wo + ox) * a.out_ld;
#pragma unroll
  for (int q = 0; q < CO; q++) {
    float v = acc[q];
    if (a.act == 1) v = ak_selu(v);
    else if (a.act == 2) v = fminf(fmaxf(v, -a.maxoff), a.maxoff);
    else if (a.act == 3) v = 1.f / (1.f + expf(-v));
    if (q < a.cout) o[q] = v;
  }
}

// ---- deformable 3x3 convolution (torchvision deform_conv2d semantics) + folded BN [+ 1x1 shortcut] + SELU ----------------------
// CTA = 16 consecutive pixels of one image: phase 1 gathers the 9 bilinear samples of every input channel into shared
// memory, phase 2 contracts them with the weights (thread -> one output channel x PXT pixels).
struct AkDcnArgs {
  const float* in; int B, H, W, CIN;
  const float* off; int off_ld;               // [B][H][W][off_ld], (dy, dx) pairs for the 9 taps
  const float* w; const float* bias;          // [9][CIN][COUT]
  const float* res; int CRES; const float* ds_w; const float* ds_b;
  float* out;
};

template <int COUT>
__global__ void __launch_bounds__(256) ak_dcn_kernel(const AkDcnArgs a) {
  extern __shared__ __align__(16) float s_col[];  // [16][9][CIN]
  constexpr int G = 256 / COUT, PXT = 16 / G;
  const int b = blockIdx.y, p0 = blockIdx.x * 16, tid = threadIdx.x, HW = a.H * a.W, CIN = a.CIN;
  const float* inb = a.in + (long long)b * HW * CIN;
  const int c4n = CIN / 4;
  for (int i = tid; i < 16 * 9 * c4n; i += 256) {
    const int c4 = i % c4n, t = (i / c4n) % 9, px = i / (c4n * 9);
    const int p = p0 + px;
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p < HW) {
      const int y = p / a.W, x = p % a.W;
      const float* o = a.off + ((long long)b * HW + p) * a.off_ld + 2 * t;
      const float h = (float)(y - 1 + t / 3) + o[0], w = (float)(x - 1 + t % 3) + o[1];
      if (h > -1.f && h < (float)a.H && w > -1.f && w < (float)a.W) {
        const float hf = floorf(h), wf = floorf(w);
        const int hl = (int)hf, wl = (int)wf, hh = hl + 1, wh = wl + 1;
        const float lh = h - hf, lw = w - wf, uh = 1.f - lh, uw = 1.f - lw;
        auto ld = [&](int yy, int xx) -> float4 {
          if (yy < 0 || yy > a.H - 1 || xx < 0 || xx > a.W - 1) return make_float4(0.f, 0.f, 0.f, 0.f);
          return *reinterpret_cast<const float4*>(inb + ((long long)yy * a.W + xx) * CIN + 4 * c4);
        };
        const float4 v1 = ld(hl, wl), v2 = ld(hl, wh), v3 = ld(hh, wl), v4 = ld(hh, wh);
        const float w1 = uh * uw, w2 = uh * lw, w3 = lh * uw, w4 = lh * lw;
        val.x = w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x;
        val.y = w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y;
        val.z = w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z;
        val.w = w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w;
      }
    }
    *reinterpret_cast<float4*>(s_col + ((long long)px * 9 + t) * CIN + 4 * c4) = val;
  }
  __syncthreads();
  const int co = tid % COUT, g = tid / COUT;
  float acc[PXT];
#pragma unroll
  for (int j = 0; j < PXT; j++) acc[j] = 0.f;
  for (int k = 0; k < 9 * CIN; k += 4) {
    const float w0 = __ldg(a.w + (long long)k * COUT + co), w1 = __ldg(a.w + (long long)(k + 1) * COUT + co);
    const float w2 = __ldg(a.w + (long long)(k + 2) * COUT + co), w3 = __ldg(a.w + (long long)(k + 3) * COUT + co);
#pragma unroll
    for (int j = 0; j < PXT; j++) {
      const float4 c = *reinterpret_cast<const float4*>(s_col + (long long)(g * PXT + j) * 9 * CIN + k);
      acc[j] = fmaf(c.x, w0, acc[j]); acc[j] = fmaf(c.y, w1, acc[j]); acc[j] = fmaf(c.z, w2, acc[j]); acc[j] = fmaf(c.w, w3, acc[j]);
    }
  }
  const float bias = a.bias[co];
#pragma unroll
  for (int j = 0; j < PXT; j++) {
    const int p = p0 + g * PXT + j;
    if (p >= HW) continue;
    float v = acc[j] + bias;
    if (a.res) {
      const float* r = a.res + ((long long)b * HW + p) * a.CRES;
      for (int c = 0; c < a.CRES; c++) v = fmaf(r[c], __ldg(a.ds_w + c * COUT + co), v);
      v += a.ds_b[co];
    }
    a.out[((long long)b * HW + p) * COUT + co] = ak_selu(v);
  }
}

// ---- 1x1 convolution to 32 channels + SELU (the four aggregation heads, :723-726) -----------------------------------------------
__global__ void __launch_bounds__(256) ak_conv1x1_selu_kernel(const float* __restrict__ in, const float* __restrict__ w /*[CIN][32]*/,
                                                              float* __restrict__ out, long long npix, int CIN) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= npix * 32) return;
  const int co = (int)(i % 32);
  const float* p = in + (i / 32) * CIN;
  float acc = 0.f;
  for (int c = 0; c < CIN; c++) acc = fmaf(p[c], __ldg(w + c * 32 + co), acc);
  out[i] = ak_selu(acc);
}

// bilinear sample of a [h][w][32] map at output pixel (y, x) of an s-times larger map, align_corners=True
// (upsample_bilinear2d: src = dst * (in - 1) / (out - 1))
__device__ __forceinline__ float ak_upsample(const float* __restrict__ m, int h, int w, int Ho, int Wo, int y, int x, int lane) {
  const float ry = Ho > 1 ? (float)(h - 1) / (float)(Ho - 1) : 0.f, rx = Wo > 1 ? (float)(w - 1) / (float)(Wo - 1) : 0.f;
  const float fy = ry * (float)y, fx = rx * (float)x;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly = fy - (float)y0, lx = fx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
  const float v00 = m[((long long)y0 * w + x0) * 32 + lane], v01 = m[((long long)y0 * w + x1) * 32 + lane];
  const float v10 = m[((long long)y1 * w + x0) * 32 + lane], v11 = m[((long long)y1 * w + x1) * 32 + lane];
  return hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
}

// ---- feature aggregation (:723-735): warp per pixel; lane l owns channel l of each of the four 32-channel groups.
// Writes the L2-normalised 128-channel map on the unpadded window and SELU(score_head.0 (x1234)) on the padded map. ------
__global__ void __launch_bounds__(256, 4) ak_fuse_kernel(const float* __restrict__ x1, const float* __restrict__ w1 /*[16][32]*/,
                                                         const float* __restrict__ a2, const float* __restrict__ a3, const float* __restrict__ a4,
                                                         const float* __restrict__ s0w /*[128][8]*/, float* __restrict__ feat,
                                                         float* __restrict__ s0, int B, int Hp, int Wp, int H, int W, int pt, int pl) {
  __shared__ __align__(16) float s_w0[128 * 8];
  for (int i = threadIdx.x; i < 128 * 8; i += 256) s_w0[i] = s0w[i];
  __syncthreads();
  const int lane = threadIdx.x % 32;
  float w1r[16];
#pragma unroll
  for (int c = 0; c < 16; c++) w1r[c] = w1[c * 32 + lane];
  const long long npix = (long long)B * Hp * Wp;
  for (long long p = (long long)blockIdx.x * 8 + threadIdx.x / 32; p < npix; p += (long long)gridDim.x * 8) {
    const int x = (int)(p % Wp), y = (int)((p / Wp) % Hp), b = (int)(p / ((long long)Wp * Hp));
    const float xv = lane < 16 ? x1[p * 16 + lane] : 0.f;
    float v[4];
    v[1] = ak_upsample(a2 + (long long)b * (Hp / 2) * (Wp / 2) * 32, Hp / 2, Wp / 2, Hp, Wp, y, x, lane);
    v[2] = ak_upsample(a3 + (long long)b * (Hp / 8) * (Wp / 8) * 32, Hp / 8, Wp / 8, Hp, Wp, y, x, lane);
    v[3] = ak_upsample(a4 + (long long)b * (Hp / 32) * (Wp / 32) * 32, Hp / 32, Wp / 32, Hp, Wp, y, x, lane);
    v[0] = 0.f;
#pragma unroll
    for (int c = 0; c < 16; c++) v[0] = fmaf(__shfl_sync(0xffffffffu, xv, c), w1r[c], v[0]);
    v[0] = ak_selu(v[0]);
    // score_head.0: 8 partial dot products per lane, then a transposing butterfly (9 shuffles instead of 8 x 5):
    // after the xor-16 / 8 / 4 exchanges lane l holds output o = bit4*4 + bit3*2 + bit2 summed over its 8-lane class
    float part[8];
#pragma unroll
    for (int o = 0; o < 8; o++) part[o] = 0.f;
#pragma unroll
    for (int g = 0; g < 4; g++) {
      const float4 wa = *reinterpret_cast<const float4*>(&s_w0[(g * 32 + lane) * 8]), wb = *reinterpret_cast<const float4*>(&s_w0[(g * 32 + lane) * 8 + 4]);
      part[0] = fmaf(v[g], wa.x, part[0]); part[1] = fmaf(v[g], wa.y, part[1]); part[2] = fmaf(v[g], wa.z, part[2]); part[3] = fmaf(v[g], wa.w, part[3]);
      part[4] = fmaf(v[g], wb.x, part[4]); part[5] = fmaf(v[g], wb.y, part[5]); part[6] = fmaf(v[g], wb.z, part[6]); part[7] = fmaf(v[g], wb.w, part[7]);
    }
    float k4[4], k2[2], k1;
    {
      const bool hi = lane & 16;
#pragma unroll
      for (int q = 0; q < 4; q++) { const float send = hi ? part[q] : part[q + 4], keep = hi ? part[q + 4] : part[q]; k4[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16); }
    }
    {
      const bool hi = lane & 8;
#pragma unroll
      for (int q = 0; q < 2; q++) { const float send = hi ? k4[q] : k4[q + 2], keep = hi ? k4[q + 2] : k4[q]; k2[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8); }
    }
    {
      const bool hi = lane & 4;
      const float send = hi ? k2[0] : k2[1], keep = hi ? k2[1] : k2[0];
      k1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    k1 += __shfl_xor_sync(0xffffffffu, k1, 2);
    k1 += __shfl_xor_sync(0xffffffffu, k1, 1);
    if ((lane & 3) == 0) s0[p * 8 + (((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1))] = ak_selu(k1);
    const int uy = y - pt, ux = x - pl;
    if (uy >= 0 && uy < H && ux >= 0 && ux < W) {
      const float ss = ak_warp_sum(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
      const float d = fmaxf(sqrtf(ss), 1e-12f);
      float* o = feat + (((long long)b * H + uy) * W + ux) * 128;
#pragma unroll
      for (int g = 0; g < 4; g++) o[g * 32 + lane] = __fdiv_rn(v[g], d);
    }
  }
}

// ---- DKD threshold (:158-166): th = scores_th, or the image mean when nothing passes / scores_th <= 0 ---------------------------
__global__ void __launch_bounds__(1024) ak_threshold_kernel(const float* __restrict__ score, const float* __restrict__ nms, float* __restrict__ thr_out,
                                                            int H, int W, int border, float thr, int use_mean) {
  __shared__ float s_sum[32];
  __shared__ int s_any;
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) s_any = 0;
  __syncthreads();
  const float* sc = score + (long long)b * H * W;
  const float* nm = nms + (long long)b * H * W;
  float sum = 0.f;
  int any = 0;
  for (int i = tid; i < H * W; i += 1024) {
    sum += sc[i];
    const int y = i / W, x = i % W;
    any |= (nm[i] > thr && y >= border && y < H - border && x >= border && x < W - border);
  }
  sum = ak_warp_sum(sum);
  if (tid % 32 == 0) s_sum[tid / 32] = sum;
  if (any) s_any = 1;
  __syncthreads();
  if (tid < 32) {
    float t = ak_warp_sum(s_sum[tid]);
    if (tid == 0) thr_out[b] = (use_mean || !s_any) ? t / (float)(H * W) : thr;
  }
}

// ---- DKD refinement (:182-222): soft-argmax in the (2r+1)^2 window, normalised keypoint, bilinear score ------------------------
__global__ void __launch_bounds__(256) ak_dkd_refine_kernel(const float* __restrict__ score, const float* __restrict__ kint /*[B][cap][2]*/,
                                                            const int* __restrict__ counts, float* __restrict__ kp_norm,
                                                            float* __restrict__ kp_pix, float* __restrict__ kscore, int H, int W,
                                                            int cap, int radius) {
  const int b = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
  if (i >= counts[b]) return;
  const float* sc = score + (long long)b * H * W;
  const long long o = (long long)b * cap + i;
  const int x = (int)kint[o * 2], y = (int)kint[o * 2 + 1];
  float mx = -CUDART_INF_F;
  for (int dy = -radius; dy <= radius; dy++)
    for (int dx = -radius; dx <= radius; dx++) {
      const int yy = y + dy, xx = x + dx;
      const float v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? sc[yy * W + xx] : 0.f;  // nn.Unfold zero padding
      mx = fmaxf(mx, v);
    }
  float se = 0.f, sx = 0.f, sy = 0.f;
  for (int dy = -radius; dy <= radius; dy++)
    for (int dx = -radius; dx <= radius; dx++) {
      const int yy = y + dy, xx = x + dx;
      const float v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? sc[yy * W + xx] : 0.f;
      const float e = expf(__fdiv_rn(v - mx, 0.1f));
      se += e; sx = fmaf(e, (float)dx, sx); sy = fmaf(e, (float)dy, sy);
    }
  const float kx = __fsub_rn(__fmul_rn(__fdiv_rn((float)x + __fdiv_rn(sx, se), (float)(W - 1)), 2.f), 1.f);
  const float ky = __fsub_rn(__fmul_rn(__fdiv_rn((float)y + __fdiv_rn(sy, se), (float)(H - 1)), 2.f), 1.f);
  kp_norm[o * 2] = kx; kp_norm[o * 2 + 1] = ky;
  kp_pix[o * 2] = __fdiv_rn(__fmul_rn((float)(W - 1), __fadd_rn(kx, 1.f)), 2.f);      // wh * (k + 1) / 2 (:771)
  kp_pix[o * 2 + 1] = __fdiv_rn(__fmul_rn((float)(H - 1), __fadd_rn(ky, 1.f)), 2.f);
  // grid_sample(bilinear, align_corners=True, zeros)
  const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(kx, 1.f), 2.f), (float)(W - 1)), iy = __fmul_rn(__fdiv_rn(__fadd_rn(ky, 1.f), 2.f), (float)(H - 1));
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix, wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
  auto at = [&](int yy, int xx) { return (yy >= 0 && yy < H && xx >= 0 && xx < W) ? sc[yy * W + xx] : 0.f; };
  kscore[o] = at(y0, x0) * (wx0 * wy0) + at(y0, x0 + 1) * (wx1 * wy0) + at(y0 + 1, x0) * (wx0 * wy1) + at(y0 + 1, x0 + 1) * (wx1 * wy1);
}

// ---- SDDH (:536-609): CTA = 16 keypoints of one image, 256 threads ----------------------------------------------------------------
constexpr int SD_TK = 16, SD_M = 16, SD_C = 128;
struct AkSddhArgs {
  const float* feat; int H, W;                 // [B][H][W][128]
  const float* kp_norm; const int* counts; int cap;
  const float* off0_w; const float* off0_b;    // [9][128][32], [32]
  const float* off2_w; const float* off2_b;    // [32][32] (in, out), [32]
  const float* sf_w;                           // [128 c][128 d]
  const float* agg;                            // [16][128 c][128 d]
  float* desc;                                 // [B][cap][128]
};

__global__ void __launch_bounds__(256) ak_sddh_kernel(const AkSddhArgs a) {
  extern __shared__ __align__(16) float s_dyn[];
  float* s_f = s_dyn;                                   // [16 k][16 p][128 c]   (first used as the 3x3 patches [16][9][128])
  float* s_2 = s_f + SD_TK * SD_M * SD_C;               // [16 k][128]
  float* s_o1 = s_2 + SD_TK * SD_C;                     // [16][32]
  float* s_off = s_o1 + SD_TK * 32;                     // [16][32]
  float* s_kwh = s_off + SD_TK * 32;                    // [16][2]
  const int b = blockIdx.y, k0 = blockIdx.x * SD_TK, tid = threadIdx.x, lane = tid % 32, wid = tid / 32;
  const int n = a.counts[b];
  if (k0 >= n) return;
  const int H = a.H, W = a.W;
  const float* fb = a.feat + (long long)b * H * W * SD_C;
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  const float max_offset = (float)max(H, W) / 4.0f;
  if (tid < SD_TK) {
    const int k = min(k0 + tid, n - 1);
    const float* kp = a.kp_norm + ((long long)b * a.cap + k) * 2;
    s_kwh[tid * 2] = __fmul_rn(__fadd_rn(__fdiv_rn(kp[0], 2.f), 0.5f), wm1);   // (k / 2 + 0.5) * wh (:552)
    s_kwh[tid * 2 + 1] = __fmul_rn(__fadd_rn(__fdiv_rn(kp[1], 2.f), 0.5f), hm1);
  }
  __syncthreads();
  // 3x3 patches at the truncated keypoint (get_patches :49-65: corner = long(k_long - K/2 + 1), clamped to [0, size-1-K])
  for (int i = tid; i < SD_TK * 9 * (SD_C / 4); i += 256) {
    const int c4 = i % (SD_C / 4), t = (i / (SD_C / 4)) % 9, k = i / (9 * SD_C / 4);
    const int kx = (int)s_kwh[k * 2], ky = (int)s_kwh[k * 2 + 1];
    const int cx = min(max((int)((float)kx - 0.5f), 0), W - 1 - 3), cy = min(max((int)((float)ky - 0.5f), 0), H - 1 - 3);
    reinterpret_cast<float4*>(s_f)[(k * 9 + t) * (SD_C / 4) + c4] =
        *reinterpret_cast<const float4*>(fb + ((long long)(cy + t / 3) * W + cx + t % 3) * SD_C + 4 * c4);
  }
  __syncthreads();
  // offset_conv.0 (3x3 valid, 128 -> 32) + SELU: thread -> output channel o of keypoints kk, kk + 8
  {
    const int o = tid % 32, kk = tid / 32;
    float a0 = 0.f, a1 = 0.f;
    const float* p0 = s_f + (long long)kk * 9 * SD_C;
    const float* p1 = s_f + (long long)(kk + 8) * 9 * SD_C;
    for (int j = 0; j < 9 * SD_C; j++) {
      const float w = __ldg(a.off0_w + j * 32 + o);
      a0 = fmaf(p0[j], w, a0); a1 = fmaf(p1[j], w, a1);
    }
    s_o1[kk * 32 + o] = ak_selu(a0 + a.off0_b[o]);
    s_o1[(kk + 8) * 32 + o] = ak_selu(a1 + a.off0_b[o]);
  }
  __syncthreads();
  {
    const int o = tid % 32, kk = tid / 32;
    float a0 = a.off2_b[o], a1 = a0;
    for (int j = 0; j < 32; j++) {
      const float w = __ldg(a.off2_w + j * 32 + o);
      a0 = fmaf(s_o1[kk * 32 + j], w, a0); a1 = fmaf(s_o1[(kk + 8) * 32 + j], w, a1);
    }
    s_off[kk * 32 + o] = fminf(fmaxf(a0, -max_offset), max_offset);
    s_off[(kk + 8) * 32 + o] = fminf(fmaxf(a1, -max_offset), max_offset);
  }
  __syncthreads();
  // M bilinear samples per keypoint: offset channels are [2][M] (x block then y block, :571); warp -> (k, p) pairs
  for (int q = wid; q < SD_TK * SD_M; q += 8) {
    const int k = q / SD_M, p = q % SD_M;
    const float px = __fadd_rn(s_kwh[k * 2], s_off[k * 32 + p]), py = __fadd_rn(s_kwh[k * 2 + 1], s_off[k * 32 + SD_M + p]);
    // pos = 2 * pos / wh - 1, then grid_sample un-normalises: ((g + 1) / 2) * (size - 1)
    const float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, px), wm1), 1.f), gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, py), hm1), 1.f);
    const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), wm1), iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), hm1);
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix, wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    auto add = [&](int yy, int xx, float wt) {
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
        const float4 v = *reinterpret_cast<const float4*>(fb + ((long long)yy * W + xx) * SD_C + lane * 4);
        acc.x = fmaf(v.x, wt, acc.x); acc.y = fmaf(v.y, wt, acc.y); acc.z = fmaf(v.z, wt, acc.z); acc.w = fmaf(v.w, wt, acc.w);
      }
    };
    add(y0, x0, wx0 * wy0); add(y0, x0 + 1, wx1 * wy0); add(y0 + 1, x0, wx0 * wy1); add(y0 + 1, x0 + 1, wx1 * wy1);
    // (written after the patches were consumed: the barrier above orders it)
    *reinterpret_cast<float4*>(s_f + ((long long)(k * SD_M + p)) * SD_C + lane * 4) = acc;
  }
  __syncthreads();
  // per position p: sf_conv (1x1) + SELU, then the aggregation with agg_weights[p]
  const int d = tid % SD_C, kh = tid / SD_C;   // thread -> channel d of keypoints kh*8 .. kh*8+7
  float dacc[8];
#pragma unroll
  for (int j = 0; j < 8; j++) dacc[j] = 0.f;
  for (int p = 0; p < SD_M; p++) {
    float f2[8];
#pragma unroll
    for (int j = 0; j < 8; j++) f2[j] = 0.f;
    for (int c = 0; c < SD_C; c += 4) {
      const float w0 = __ldg(a.sf_w + (c + 0) * SD_C + d), w1 = __ldg(a.sf_w + (c + 1) * SD_C + d);
      const float w2 = __ldg(a.sf_w + (c + 2) * SD_C + d), w3 = __ldg(a.sf_w + (c + 3) * SD_C + d);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const float4 v = *reinterpret_cast<const float4*>(s_f + ((long long)((kh * 8 + j) * SD_M + p)) * SD_C + c);
        f2[j] = fmaf(v.x, w0, f2[j]); f2[j] = fmaf(v.y, w1, f2[j]); f2[j] = fmaf(v.z, w2, f2[j]); f2[j] = fmaf(v.w, w3, f2[j]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; j++) s_2[(kh * 8 + j) * SD_C + d] = ak_selu(f2[j]);
    __syncthreads();
    const float* ag = a.agg + (long long)p * SD_C * SD_C;
    for (int c = 0; c < SD_C; c += 4) {
      const float w0 = __ldg(ag + (c + 0) * SD_C + d), w1 = __ldg(ag + (c + 1) * SD_C + d);
      const float w2 = __ldg(ag + (c + 2) * SD_C + d), w3 = __ldg(ag + (c + 3) * SD_C + d);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const float4 v = *reinterpret_cast<const float4*>(s_2 + (kh * 8 + j) * SD_C + c);
        dacc[j] = fmaf(v.x, w0, dacc[j]); dacc[j] = fmaf(v.y, w1, dacc[j]); dacc[j] = fmaf(v.z, w2, dacc[j]); dacc[j] = fmaf(v.w, w3, dacc[j]);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; j++) s_2[(kh * 8 + j) * SD_C + d] = dacc[j];
  __syncthreads();
  for (int k = wid; k < SD_TK; k += 8) {
    if (k0 + k >= n) continue;
    const float4 v = *reinterpret_cast<const float4*>(s_2 + k * SD_C + lane * 4);
    const float ss = ak_warp_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w);
    const float dn = fmaxf(sqrtf(ss), 1e-12f);
    *reinterpret_cast<float4*>(a.desc + ((long long)b * a.cap + k0 + k) * SD_C + lane * 4) =
        make_float4(__fdiv_rn(v.x, dn), __fdiv_rn(v.y, dn), __fdiv_rn(v.z, dn), __fdiv_rn(v.w, dn));
  }
}

// ---- host-side launch helpers -------------------------------------------------------------------------------------------------
template <int CO, int CC>
int ak_conv(const AkConvArgs& a, cudaStream_t st) {
  dim3 grid(ceil_div(a.W, 32) * ceil_div(a.H, 8), a.B);
  ak_conv3x3_kernel<CO, CC><<<grid, 256, 0, st>>>(a);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}

AkConvArgs ak_args(const float* in, int B, int H, int W, int CIN, const float* w, const float* bias, float* out, int out_ld, int cout, int act) {
  AkConvArgs a{};
  a.in = in; a.B = B; a.H = H; a.W = W; a.CIN = CIN; a.w = w; a.bias = bias; a.out = out; a.out_ld = out_ld; a.cout = cout; a.act = act;
  a.oy0 = 0; a.ox0 = 0; a.Ho = H; a.Wo = W;
  return a;
}

template <int COUT>
int ak_dcn(const AkDcnArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)16 * 9 * a.CIN * sizeof(float);
  IMW_CHECK_CUDA(cudaFuncSetAttribute(ak_dcn_kernel<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ak_dcn_kernel<COUT><<<dim3(ceil_div(a.H * a.W, 16), a.B), 256, smem, st>>>(a);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}

struct AkBuffers {
  float *img4, *x1, *mid, *p1, *t2, *x2, *p2, *off, *t3, *x3, *p3, *t4, *x4, *a2, *a3, *a4, *feat, *s0, *s1, *s2, *score, *nms, *thr, *kint,
      *kp_norm;
  unsigned long long* keys;
  int* sel_counts;
};

size_t ak_carve(Workspace& ws, AkBuffers& b, int SB, int H, int W, int Hp, int Wp, int cap) {
  const size_t P = (size_t)SB * Hp * Wp;
  b.img4 = ws.take<float>(P * 4); b.x1 = ws.take<float>(P * 16); b.mid = ws.take<float>(P * 16);
  b.p1 = ws.take<float>(P / 4 * 16); b.t2 = ws.take<float>(P / 4 * 32); b.x2 = ws.take<float>(P / 4 * 32);
  b.p2 = ws.take<float>(P / 64 * 32); b.off = ws.take<float>(P / 64 * 24); b.t3 = ws.take<float>(P / 64 * 64); b.x3 = ws.take<float>(P / 64 * 64);
  b.p3 = ws.take<float>(P / 1024 * 64); b.t4 = ws.take<float>(P / 1024 * 128); b.x4 = ws.take<float>(P / 1024 * 128);
  b.a2 = ws.take<float>(P / 4 * 32); b.a3 = ws.take<float>(P / 64 * 32); b.a4 = ws.take<float>(P / 1024 * 32);
  b.feat = ws.take<float>((size_t)SB * H * W * 128);
  b.s0 = ws.take<float>(P * 8); b.s1 = ws.take<float>(P * 4); b.s2 = ws.take<float>(P * 4);
  b.score = ws.take<float>((size_t)SB * H * W); b.nms = ws.take<float>((size_t)SB * H * W);
  b.thr = ws.take<float>(SB);
  b.keys = ws.take<unsigned long long>((size_t)SB * sp_select_key_cap(H, W));
  b.kint = ws.take<float>((size_t)SB * cap * 2); b.kp_norm = ws.take<float>((size_t)SB * cap * 2);
  b.sel_counts = ws.take<int>(2 * SB);
  return ws.off;
}

constexpr int AK_SUB_BATCH = 8;

}  // namespace

extern "C" size_t imw_aliked_workspace_bytes(int n_images, int height, int width, int cap) {
  if (n_images <= 0 || height <= 0 || width <= 0 || cap <= 0) return 0;
  const int Hp = (height + 31) / 32 * 32, Wp = (width + 31) / 32 * 32;
  Workspace ws(nullptr, 0);
  AkBuffers b;
  return ak_carve(ws, b, n_images < AK_SUB_BATCH ? n_images : AK_SUB_BATCH, height, width, Hp, Wp, cap) + 4096;
}

#define RUN(x)                 \
  do {                         \
    int _rc = (x);             \
    if (_rc != IMW_OK) return _rc; \
  } while (0)

extern "C" int imw_aliked_forward(const imw_aliked_weights* Wt, const imw_aliked_conf* conf, int n_images, int channels, int height,
                                  int width, const float* images, int cap, float* keypoints, float* scores, float* descriptors,
                                  int* counts, float* dbg_score_map, float* dbg_feature_map, void* workspace, size_t workspace_bytes,
                                  cudaStream_t st) {
  IMW_REQUIRE(Wt && conf && images && keypoints && scores && descriptors && counts, "imw_aliked_forward: null argument");
  IMW_REQUIRE(n_images > 0 && (channels == 1 || channels == 3), "imw_aliked_forward: images must be [B][1|3][H][W] (channels = %d)", channels);
  IMW_REQUIRE(height >= 32 && width >= 32 && cap > 0, "imw_aliked_forward: image at least 32x32, cap > 0");
  IMW_REQUIRE(conf->nms_radius >= 1 && conf->nms_radius <= 8, "imw_aliked_forward: nms_radius in [1, 8]");
  // DKD(top_k = max_num_keypoints when detection_threshold <= 0, :676-678); top_k <= 0 and scores_th <= 0 -> mean threshold (:165-167)
  const bool topk_mode = !(conf->detection_threshold > 0.f) && conf->max_num_keypoints > 0;
  const bool mean_mode = !(conf->detection_threshold > 0.f) && !topk_mode;
  const int H = height, W = width, Hp = (H + 31) / 32 * 32, Wp = (W + 31) / 32 * 32;
  const int ph = Hp - H, pw = Wp - W, pt = ph / 2, pl = pw / 2;   // InputPadder: [pw/2, pw - pw/2, ph/2, ph - ph/2]
  const int n_limit = conf->max_num_keypoints > 0 ? conf->max_num_keypoints : 20000;
  Workspace ws(workspace, workspace_bytes);
  AkBuffers b;
  const int SBmax = n_images < AK_SUB_BATCH ? n_images : AK_SUB_BATCH;
  ak_carve(ws, b, SBmax, H, W, Hp, Wp, cap);
  if (ws.overflow || !workspace) { imw_set_error("imw_aliked_forward: workspace too small (%zu < %zu)", workspace_bytes, ws.off); return IMW_ERR_WORKSPACE; }
  const imw_aliked_weights& w = *Wt;
  IMW_CHECK_CUDA(cudaFuncSetAttribute(ak_sddh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)((SD_TK * SD_M * SD_C + SD_TK * SD_C + 3 * SD_TK * 32) * sizeof(float))));
  for (int i0 = 0; i0 < n_images; i0 += AK_SUB_BATCH) {
    const int SB = (n_images - i0) < AK_SUB_BATCH ? (n_images - i0) : AK_SUB_BATCH;
    const long long P = (long long)SB * Hp * Wp;
    ak_pad_image_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(images + (long long)i0 * channels * H * W, b.img4, SB, channels, H, W, Hp, Wp, pt, pl);
    IMW_CHECK_LAUNCH_T("ak_pad_image_kernel");
    // block1 (ConvBlock 3 -> 16 -> 16)
    RUN((ak_conv<16, 4>(ak_args(b.img4, SB, Hp, Wp, 4, w.b1c1_w, w.b1c1_b, b.mid, 16, 16, 1), st)));
    RUN((ak_conv<16, 16>(ak_args(b.mid, SB, Hp, Wp, 16, w.b1c2_w, w.b1c2_b, b.x1, 16, 16, 1), st)));
    // block2 (ResBlock 16 -> 32 at 1/2)
    const int H2 = Hp / 2, W2 = Wp / 2, H8 = Hp / 8, W8 = Wp / 8, H32 = Hp / 32, W32 = Wp / 32;
    ak_avgpool_kernel<2><<<(unsigned)((P / 4 * 16 + 255) / 256), 256, 0, st>>>(b.x1, b.p1, SB, H2, W2, 16);
    IMW_CHECK_LAUNCH_T("ak_avgpool_kernel<2>");
    RUN((ak_conv<32, 16>(ak_args(b.p1, SB, H2, W2, 16, w.b2c1_w, w.b2c1_b, b.t2, 32, 32, 1), st)));
    {
      AkConvArgs a = ak_args(b.t2, SB, H2, W2, 32, w.b2c2_w, w.b2c2_b, b.x2, 32, 32, 1);
      a.res = b.p1; a.CRES = 16; a.ds_w = w.b2ds_w; a.ds_b = w.b2ds_b;
      RUN((ak_conv<32, 16>(a, st)));
    }
    // block3 (deformable ResBlock 32 -> 64 at 1/8)
    ak_avgpool_kernel<4><<<(unsigned)((P / 64 * 32 + 255) / 256), 256, 0, st>>>(b.x2, b.p2, SB, H8, W8, 32);
    IMW_CHECK_LAUNCH_T("ak_avgpool_kernel<4>");
    {
      AkConvArgs o = ak_args(b.p2, SB, H8, W8, 32, w.b3o1_w, w.b3o1_b, b.off, 24, 18, 2);
      o.maxoff = (float)(H8 > W8 ? H8 : W8) / 4.0f;
      RUN((ak_conv<24, 16>(o, st)));
      AkDcnArgs d{b.p2, SB, H8, W8, 32, b.off, 24, w.b3c1_w, w.b3c1_b, nullptr, 0, nullptr, nullptr, b.t3};
      RUN(ak_dcn<64>(d, st));
      o = ak_args(b.t3, SB, H8, W8, 64, w.b3o2_w, w.b3o2_b, b.off, 24, 18, 2);
      o.maxoff = (float)(H8 > W8 ? H8 : W8) / 4.0f;
      RUN((ak_conv<24, 16>(o, st)));
      AkDcnArgs d2{b.t3, SB, H8, W8, 64, b.off, 24, w.b3c2_w, w.b3c2_b, b.p2, 32, w.b3ds_w, w.b3ds_b, b.x3};
      RUN(ak_dcn<64>(d2, st));
    }
    // block4 (deformable ResBlock 64 -> 128 at 1/32)
    ak_avgpool_kernel<4><<<(unsigned)((P / 1024 * 64 + 255) / 256), 256, 0, st>>>(b.x3, b.p3, SB, H32, W32, 64);
    IMW_CHECK_LAUNCH_T("ak_avgpool_kernel<4>");
    {
      AkConvArgs o = ak_args(b.p3, SB, H32, W32, 64, w.b4o1_w, w.b4o1_b, b.off, 24, 18, 2);
      o.maxoff = (float)(H32 > W32 ? H32 : W32) / 4.0f;
      RUN((ak_conv<24, 16>(o, st)));
      AkDcnArgs d{b.p3, SB, H32, W32, 64, b.off, 24, w.b4c1_w, w.b4c1_b, nullptr, 0, nullptr, nullptr, b.t4};
      RUN(ak_dcn<128>(d, st));
      o = ak_args(b.t4, SB, H32, W32, 128, w.b4o2_w, w.b4o2_b, b.off, 24, 18, 2);
      o.maxoff = (float)(H32 > W32 ? H32 : W32) / 4.0f;
      RUN((ak_conv<24, 16>(o, st)));
      AkDcnArgs d2{b.t4, SB, H32, W32, 128, b.off, 24, w.b4c2_w, w.b4c2_b, b.p3, 64, w.b4ds_w, w.b4ds_b, b.x4};
      RUN(ak_dcn<128>(d2, st));
    }
    // aggregation heads + fused upsample / concat / normalise / score_head.0
    ak_conv1x1_selu_kernel<<<(unsigned)((P / 4 * 32 + 255) / 256), 256, 0, st>>>(b.x2, w.conv2_w, b.a2, P / 4, 32);
    IMW_CHECK_LAUNCH_T("ak_conv1x1_selu_kernel");
    ak_conv1x1_selu_kernel<<<(unsigned)((P / 64 * 32 + 255) / 256), 256, 0, st>>>(b.x3, w.conv3_w, b.a3, P / 64, 64);
    IMW_CHECK_LAUNCH_T("ak_conv1x1_selu_kernel");
    ak_conv1x1_selu_kernel<<<(unsigned)((P / 1024 * 32 + 255) / 256), 256, 0, st>>>(b.x4, w.conv4_w, b.a4, P / 1024, 128);
    IMW_CHECK_LAUNCH_T("ak_conv1x1_selu_kernel");
    ak_fuse_kernel<<<148 * 16, 256, 0, st>>>(b.x1, w.conv1_w, b.a2, b.a3, b.a4, w.s0_w, b.feat, b.s0, SB, Hp, Wp, H, W, pt, pl);
    IMW_CHECK_LAUNCH_T("ak_fuse_kernel");
    // score head tail (3x3 8 -> 4, 4 -> 4, 4 -> 1 + sigmoid, cropped to the unpadded window)
    RUN((ak_conv<4, 8>(ak_args(b.s0, SB, Hp, Wp, 8, w.s2_w, nullptr, b.s1, 4, 4, 1), st)));
    RUN((ak_conv<4, 4>(ak_args(b.s1, SB, Hp, Wp, 4, w.s4_w, nullptr, b.s2, 4, 4, 1), st)));
    {
      AkConvArgs a = ak_args(b.s2, SB, Hp, Wp, 4, w.s6_w, nullptr, b.score, 1, 1, 3);
      a.oy0 = pt; a.ox0 = pl; a.Ho = H; a.Wo = W;
      RUN((ak_conv<1, 4>(a, st)));
    }
    if (dbg_score_map) IMW_CHECK_CUDA(cudaMemcpyAsync(dbg_score_map + (long long)i0 * H * W, b.score, sizeof(float) * (size_t)SB * H * W, cudaMemcpyDeviceToDevice, st));
    if (dbg_feature_map) IMW_CHECK_CUDA(cudaMemcpyAsync(dbg_feature_map + (long long)i0 * H * W * 128, b.feat, sizeof(float) * (size_t)SB * H * W * 128, cudaMemcpyDeviceToDevice, st));
    // DKD
    RUN(sp_nms(b.score, b.nms, SB, H, W, conf->nms_radius, st));
    const float thr = topk_mode ? 0.f : conf->detection_threshold;
    if (!topk_mode) {
      ak_threshold_kernel<<<SB, 1024, 0, st>>>(b.score, b.nms, b.thr, H, W, conf->nms_radius, thr, mean_mode ? 1 : 0);
      IMW_CHECK_LAUNCH_T("ak_threshold_kernel");
    }
    RUN(sp_select(b.nms, b.keys, (int)sp_select_key_cap(H, W), b.kint, scores + (long long)i0 * cap, b.sel_counts, SB, H, W, thr,
                  conf->nms_radius, topk_mode ? conf->max_num_keypoints : n_limit, cap, st, topk_mode ? nullptr : b.thr));
    IMW_CHECK_CUDA(cudaMemcpyAsync(counts + i0, b.sel_counts, sizeof(int) * SB, cudaMemcpyDeviceToDevice, st));
    IMW_CHECK_CUDA(cudaMemcpyAsync(counts + n_images + i0, b.sel_counts + SB, sizeof(int) * SB, cudaMemcpyDeviceToDevice, st));
    ak_dkd_refine_kernel<<<dim3(ceil_div(cap, 256), SB), 256, 0, st>>>(b.score, b.kint, b.sel_counts, b.kp_norm, keypoints + (long long)i0 * cap * 2,
                                                                       scores + (long long)i0 * cap, H, W, cap, conf->nms_radius);
    IMW_CHECK_LAUNCH_T("ak_dkd_refine_kernel");
    // SDDH
    AkSddhArgs sa{b.feat, H, W, b.kp_norm, b.sel_counts, cap, w.sd_off0_w, w.sd_off0_b, w.sd_off2_w, w.sd_off2_b, w.sd_sf_w, w.sd_agg,
                  descriptors + (long long)i0 * cap * 128};
    ak_sddh_kernel<<<dim3(ceil_div(cap, SD_TK), SB), 256, (SD_TK * SD_M * SD_C + SD_TK * SD_C + 3 * SD_TK * 32) * sizeof(float), st>>>(sa);
    IMW_CHECK_LAUNCH_T("ak_sddh_kernel");
  }
  return IMW_OK;
}
