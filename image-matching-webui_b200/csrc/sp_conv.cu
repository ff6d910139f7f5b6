// SuperPoint VGG encoder / head convolutions, exact-fp32 CUDA-core path.
// Follows third_party/SuperGluePretrainedNetwork/models/superpoint.py:152-166,194-196
// (3x3 conv, zero pad 1, bias, ReLU, optional fused 2x2/2 max-pool).
//
// Layout: activations NHWC fp32 ([B][H][W][C]); weights re-laid out by the host to
// [tap = ky*3+kx][Cin][Cout] so the Cout slice of one (tap, cin) is contiguous.
#include "split_planes.cuh"

#include "common.cuh"
#include "sp_kernels.h"

namespace {

constexpr int TILE = 16;        // output tile 16x16 pixels per CTA
constexpr int CH = 16;          // input-channel chunk staged in shared memory
constexpr int IN_ROW = 20;      // padded row stride of the (TILE+2)-wide halo tile (16 B aligned rows)
constexpr int IN_PLANE = 18 * IN_ROW;
constexpr int COUT_T = 64;      // output channels per CTA

// grid: (tiles_x*tiles_y, Cout/64, B), block 256.
// thread -> 8 output channels (cg) x a 2x4 pixel patch (py,px): the 2x2 pool windows are thread-local.
__global__ void __launch_bounds__(256, 2)
conv3x3_nhwc_kernel(const float* __restrict__ in, const float* __restrict__ wgt, const float* __restrict__ bias,
                    float* __restrict__ out, int H, int W, int Cin, int Cout, int relu, int pool) {
  extern __shared__ __align__(16) float smem[];
  float* s_in = smem;                    // [CH][18][IN_ROW]
  float* s_w = smem + CH * IN_PLANE;     // [9][CH][COUT_T]

  const int tiles_x = (W + TILE - 1) / TILE;
  const int tx0 = (blockIdx.x % tiles_x) * TILE, ty0 = (blockIdx.x / tiles_x) * TILE;
  const int co0 = blockIdx.y * COUT_T;
  const int b = blockIdx.z;
  const int tid = threadIdx.x;
  const int cg = tid % 8, pg = tid / 8;
  const int py = pg / 4, px = pg % 4;

  float acc[2][4][8];
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int k = 0; k < 8; k++) acc[i][j][k] = 0.f;

  const float* in_b = in + (long long)b * H * W * Cin;
  for (int c0 = 0; c0 < Cin; c0 += CH) {
    // stage the halo tile, transposed to [c][y][x]
    for (int i = tid; i < 18 * 18 * (CH / 4); i += 256) {
      int q = i % (CH / 4), p = i / (CH / 4);
      int yy = p / 18, xx = p % 18;
      int gy = ty0 + yy - 1, gx = tx0 + xx - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy >= 0 && gy < H && gx >= 0 && gx < W)
        v = *reinterpret_cast<const float4*>(in_b + ((long long)gy * W + gx) * Cin + c0 + q * 4);
      float* d = s_in + (q * 4) * IN_PLANE + yy * IN_ROW + xx;
      d[0] = v.x; d[IN_PLANE] = v.y; d[2 * IN_PLANE] = v.z; d[3 * IN_PLANE] = v.w;
    }
    for (int i = tid; i < 9 * CH * (COUT_T / 4); i += 256) {
      int q = i % (COUT_T / 4), r = i / (COUT_T / 4);  // r = tap*CH + c
      int tap = r / CH, c = r % CH;
      float4 v = *reinterpret_cast<const float4*>(wgt + ((long long)tap * Cin + c0 + c) * Cout + co0 + q * 4);
      *reinterpret_cast<float4*>(s_w + r * COUT_T + q * 4) = v;
    }
    __syncthreads();
#pragma unroll 2
    for (int c = 0; c < CH; c++) {
      float v[4][6];
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const float* p = s_in + c * IN_PLANE + (2 * py + r) * IN_ROW + 4 * px;
        float4 a = *reinterpret_cast<const float4*>(p);
        float2 e = *reinterpret_cast<const float2*>(p + 4);
        v[r][0] = a.x; v[r][1] = a.y; v[r][2] = a.z; v[r][3] = a.w; v[r][4] = e.x; v[r][5] = e.y;
      }
#pragma unroll
      for (int dy = 0; dy < 3; dy++)
#pragma unroll
        for (int dx = 0; dx < 3; dx++) {
          const float* wp = s_w + ((dy * 3 + dx) * CH + c) * COUT_T + cg * 8;
          float4 w0 = *reinterpret_cast<const float4*>(wp);
          float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
          float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int oy = 0; oy < 2; oy++)
#pragma unroll
            for (int ox = 0; ox < 4; ox++)
#pragma unroll
              for (int k = 0; k < 8; k++) acc[oy][ox][k] = fmaf(v[oy + dy][ox + dx], wv[k], acc[oy][ox][k]);
        }
    }
    __syncthreads();
  }

  float bv[8];
#pragma unroll
  for (int k = 0; k < 8; k++) bv[k] = bias[co0 + cg * 8 + k];
#pragma unroll
  for (int oy = 0; oy < 2; oy++)
#pragma unroll
    for (int ox = 0; ox < 4; ox++)
#pragma unroll
      for (int k = 0; k < 8; k++) {
        float x = acc[oy][ox][k] + bv[k];
        acc[oy][ox][k] = relu ? fmaxf(x, 0.f) : x;
      }

  if (!pool) {
#pragma unroll
    for (int oy = 0; oy < 2; oy++)
#pragma unroll
      for (int ox = 0; ox < 4; ox++) {
        int gy = ty0 + 2 * py + oy, gx = tx0 + 4 * px + ox;
        if (gy < H && gx < W) {
          float* o = out + (((long long)b * H + gy) * W + gx) * Cout + co0 + cg * 8;
          *reinterpret_cast<float4*>(o) = make_float4(acc[oy][ox][0], acc[oy][ox][1], acc[oy][ox][2], acc[oy][ox][3]);
          *reinterpret_cast<float4*>(o + 4) = make_float4(acc[oy][ox][4], acc[oy][ox][5], acc[oy][ox][6], acc[oy][ox][7]);
        }
      }
  } else {
    const int Ho = H / 2, Wo = W / 2;
#pragma unroll
    for (int oxp = 0; oxp < 2; oxp++) {
      int gy = (ty0 + 2 * py) / 2, gx = (tx0 + 4 * px) / 2 + oxp;
      if (gy < Ho && gx < Wo) {
        float m[8];
#pragma unroll
        for (int k = 0; k < 8; k++)
          m[k] = fmaxf(fmaxf(acc[0][2 * oxp][k], acc[0][2 * oxp + 1][k]), fmaxf(acc[1][2 * oxp][k], acc[1][2 * oxp + 1][k]));
        float* o = out + (((long long)b * Ho + gy) * Wo + gx) * Cout + co0 + cg * 8;
        *reinterpret_cast<float4*>(o) = make_float4(m[0], m[1], m[2], m[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(m[4], m[5], m[6], m[7]);
      }
    }
  }
}

// conv1a: Cin = 1.  image [B][H][W] -> [B][H][W][64], bias + ReLU.  HBM-write bound.
// thread -> 4 output channels (t%16) of pixel (t/16); a warp writes 2 pixels x 256 B contiguous.
__global__ void __launch_bounds__(256)
conv3x3_c1_kernel(const float* __restrict__ img, const float* __restrict__ wgt /*[9][1][64]*/,
                  const float* __restrict__ bias, float* __restrict__ out, plane_t* __restrict__ out_planes,
                  size_t plane_stride, int H, int W) {
  __shared__ float s_in[18][19];
  const int tiles_x = (W + TILE - 1) / TILE;
  const int tx0 = (blockIdx.x % tiles_x) * TILE, ty0 = (blockIdx.x / tiles_x) * TILE;
  const int b = blockIdx.z, tid = threadIdx.x;
  const float* im = img + (long long)b * H * W;
  for (int i = tid; i < 18 * 18; i += 256) {
    int yy = i / 18, xx = i % 18, gy = ty0 + yy - 1, gx = tx0 + xx - 1;
    s_in[yy][xx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? im[(long long)gy * W + gx] : 0.f;
  }
  const int cq = tid % 16;
  float w[9][4], bv[4];
#pragma unroll
  for (int t = 0; t < 9; t++)
#pragma unroll
    for (int k = 0; k < 4; k++) w[t][k] = wgt[t * 64 + cq * 4 + k];
#pragma unroll
  for (int k = 0; k < 4; k++) bv[k] = bias[cq * 4 + k];
  __syncthreads();
  for (int p = tid / 16; p < TILE * TILE; p += 16) {
    int yy = p / TILE, xx = p % TILE, gy = ty0 + yy, gx = tx0 + xx;
    if (gy >= H || gx >= W) continue;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int dy = 0; dy < 3; dy++)
#pragma unroll
      for (int dx = 0; dx < 3; dx++) {
        float v = s_in[yy + dy][xx + dx];
#pragma unroll
        for (int k = 0; k < 4; k++) a[k] = fmaf(v, w[dy * 3 + dx][k], a[k]);
      }
    float4 o = make_float4(fmaxf(a[0] + bv[0], 0.f), fmaxf(a[1] + bv[1], 0.f), fmaxf(a[2] + bv[2], 0.f), fmaxf(a[3] + bv[3], 0.f));
    const size_t off = (((size_t)b * H + gy) * W + gx) * 64 + cq * 4;
    if (out_planes) {  // two fp16 planes x = hi + lo * 2^-11 for the tcgen05 split-precision convs (split_planes.cuh)
      float ov[4] = {o.x, o.y, o.z, o.w};
      __align__(8) plane_t p[NP][4];
#pragma unroll
      for (int k = 0; k < 4; k++) split2(ov[k], p[0][k], p[1][k]);
#pragma unroll
      for (int q = 0; q < NP; q++) *reinterpret_cast<uint2*>(out_planes + q * plane_stride + off) = *reinterpret_cast<const uint2*>(p[q]);
    } else {
      *reinterpret_cast<float4*>(out + off) = o;
    }
  }
}

}  // namespace

int sp_conv3x3(const float* in, const float* w, const float* bias, float* out, int B, int H, int W, int Cin, int Cout,
               int relu, int pool, cudaStream_t st) {
  IMW_REQUIRE(Cin % CH == 0 && Cout % COUT_T == 0, "sp_conv3x3: Cin %% 16 / Cout %% 64 (got %d,%d)", Cin, Cout);
  IMW_REQUIRE(!pool || (H % 2 == 0 && W % 2 == 0), "sp_conv3x3: pooled conv needs even H,W");
  size_t smem = (size_t)(CH * IN_PLANE + 9 * CH * COUT_T) * sizeof(float);
  IMW_SMEM_ATTR_ONCE(conv3x3_nhwc_kernel, smem);
  dim3 grid(ceil_div(W, TILE) * ceil_div(H, TILE), Cout / COUT_T, B);
  conv3x3_nhwc_kernel<<<grid, 256, smem, st>>>(in, w, bias, out, H, W, Cin, Cout, relu, pool);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}

int sp_conv3x3_c1(const float* img, const float* w, const float* bias, float* out, void* out_planes, int B, int H, int W,
                  cudaStream_t st) {
  dim3 grid(ceil_div(W, TILE) * ceil_div(H, TILE), 1, B);
  conv3x3_c1_kernel<<<grid, 256, 0, st>>>(img, w, bias, out, (plane_t*)out_planes, (size_t)B * H * W * 64, H, W);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}
