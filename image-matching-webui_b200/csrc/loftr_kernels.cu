// LoFTR dense matcher on the GPU.
// hloc's `loftr` matcher calls kornia.feature.LoFTR (un-vendored); the in-tree source it ports is
// third_party/SE2LoFTR/src/loftr/*, followed here:
//   backbone/resnet_fpn.py:43-118      ResNet-FPN 1/8 + 1/2 (BatchNorm folded by the host) -> tcgen05 bf16x3 convs
//   utils/position_encoding.py:6-42    sine encoding (table built by the host with the reference formula)
//   loftr_module/transformer.py:7-101  8 coarse / 2 fine encoder layers, linear attention (linear_attention.py:14-47)
//   utils/coarse_matching.py:108-250   dual-softmax confidence, threshold, border, mutual max -- streamed, the
//                                      L x S confidence matrix (1.07 GB per pair at 1024^2) is never materialised
//   loftr_module/fine_preprocess.py:29-59, utils/fine_matching.py:18-77  5x5 windows gathered on demand (no F.unfold)
#include "split_planes.cuh"
#include <math_constants.h>

#include "../../include/imw_b200.h"
#include "common.cuh"
#include "gemm_simt.cuh"
#include "simreduce.cuh"
#include "sp_kernels.h"
#include "tc_gemm.cuh"
#include "tc_simreduce.cuh"

namespace {

constexpr int CD = 256, FD = 128, NH = 8;

__device__ __forceinline__ float lf_merge(const plane_t* p, size_t plane, size_t i) { return merge2(p[i], p[plane + i]); }

// ---- conv1: 7x7 stride 2 pad 3, 1 -> 128 channels, BN folded, ReLU (resnet_fpn.py:60-62,102) -> bf16 planes ----------
__global__ void __launch_bounds__(256) lf_conv1_kernel(const float* __restrict__ img, const float* __restrict__ wgt /*[49][128]*/,
                                                       const float* __restrict__ bias, plane_t* __restrict__ out, int B, size_t img_stride,
                                                       int H, int W) {
  __shared__ __align__(16) float s_in[37][38];
  __shared__ __align__(16) float s_w[49][128];
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int tiles_x = (Wo + 15) / 16;
  const int ox0 = (blockIdx.x % tiles_x) * 16, oy0 = (blockIdx.x / tiles_x) * 16, b = blockIdx.z, tid = threadIdx.x;
  const float* im = img + (size_t)b * img_stride;   // images of one side of the pairs may be interleaved with the other side's
  for (int i = tid; i < 37 * 37; i += 256) {
    int yy = i / 37, xx = i % 37, gy = oy0 * 2 - 3 + yy, gx = ox0 * 2 - 3 + xx;
    s_in[yy][xx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? im[(size_t)gy * W + gx] : 0.f;
  }
  for (int i = tid; i < 49 * 128; i += 256) s_w[i / 128][i % 128] = wgt[i];
  __syncthreads();
  const int cq = tid % 32, pg = tid / 32;  // 4 channels per thread, 8 pixels in flight
  const size_t plane = (size_t)B * Ho * Wo * 128;
  for (int p = pg; p < 256; p += 8) {
    const int yy = p / 16, xx = p % 16, oy = oy0 + yy, ox = ox0 + xx;
    if (oy >= Ho || ox >= Wo) continue;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ky = 0; ky < 7; ky++)
#pragma unroll
      for (int kx = 0; kx < 7; kx++) {
        const float v = s_in[2 * yy + ky][2 * xx + kx];
        const float4 w4 = *reinterpret_cast<const float4*>(&s_w[ky * 7 + kx][cq * 4]);
        a[0] = fmaf(v, w4.x, a[0]); a[1] = fmaf(v, w4.y, a[1]); a[2] = fmaf(v, w4.z, a[2]); a[3] = fmaf(v, w4.w, a[3]);
      }
    const size_t off = (((size_t)b * Ho + oy) * Wo + ox) * 128 + cq * 4;
    __align__(8) plane_t q[NP][4];
#pragma unroll
    for (int k = 0; k < 4; k++) split2(fmaxf(a[k] + bias[cq * 4 + k], 0.f), q[0][k], q[1][k]);
#pragma unroll
    for (int s = 0; s < NP; s++) *reinterpret_cast<uint2*>(out + s * plane + off) = *reinterpret_cast<const uint2*>(q[s]);
  }
}

// ---- FPN: out = a + bilinear_2x(b) (align_corners=True), resnet_fpn.py:110-115 -- bf16 planes in / out -------------------
__global__ void __launch_bounds__(256) lf_upsample_add_kernel(const plane_t* __restrict__ a, const plane_t* __restrict__ bsrc,
                                                              plane_t* __restrict__ out, int B, int Ho, int Wo, int C) {
  // one thread = 8 consecutive channels of one output pixel (16-byte loads / stores per plane)
  const size_t n = (size_t)B * Ho * Wo * C, i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= n) return;
  const int c = (int)(i % C);
  size_t t = i / C;
  const int x = (int)(t % Wo); t /= Wo;
  const int y = (int)(t % Ho), b = (int)(t / Ho);
  const int hs = Ho / 2, ws = Wo / 2;
  // area_pixel_compute_source_index with align_corners: src = dst * (in-1)/(out-1)
  const float sy = (Ho > 1) ? (float)(hs - 1) / (float)(Ho - 1) : 0.f, sx = (Wo > 1) ? (float)(ws - 1) / (float)(Wo - 1) : 0.f;
  const float fy = sy * y, fx = sx * x;
  const int y0 = (int)fy, x0 = (int)fx, y1 = min(y0 + 1, hs - 1), x1 = min(x0 + 1, ws - 1);
  const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
  const size_t ps = (size_t)B * hs * ws * C;
  auto load8 = [](const plane_t* base, size_t plane, size_t off, float (&v)[8]) {   // hi + lo * 2^-11 of 8 channels
    const uint4 q0 = *reinterpret_cast<const uint4*>(base + off), q1 = *reinterpret_cast<const uint4*>(base + plane + off);
    const plane_t *e0 = reinterpret_cast<const plane_t*>(&q0), *e1 = reinterpret_cast<const plane_t*>(&q1);
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = merge2(e0[k], e1[k]);
  };
  float s00[8], s01[8], s10[8], s11[8], av[8];
  load8(bsrc, ps, (((size_t)b * hs + y0) * ws + x0) * C + c, s00);
  load8(bsrc, ps, (((size_t)b * hs + y0) * ws + x1) * C + c, s01);
  load8(bsrc, ps, (((size_t)b * hs + y1) * ws + x0) * C + c, s10);
  load8(bsrc, ps, (((size_t)b * hs + y1) * ws + x1) * C + c, s11);
  load8(a, n, i, av);
  __align__(16) plane_t o0[8], o1[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const float up = hy * (hx * s00[k] + lx * s01[k]) + ly * (hx * s10[k] + lx * s11[k]);
    split2(av[k] + up, o0[k], o1[k]);
  }
  *reinterpret_cast<uint4*>(out + i) = *reinterpret_cast<const uint4*>(o0);
  *reinterpret_cast<uint4*>(out + n + i) = *reinterpret_cast<const uint4*>(o1);
}

// bf16 planes -> fp32 (fine feature map used by the window gather)
__global__ void lf_planes_to_f32_kernel(const plane_t* __restrict__ in, float* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = lf_merge(in, n, i);
}

// tokens: xm[z][l][0:256] = coarse_feat[z][l][:] + pe[l][:]   (loftr.py:58-59)
// slot z = 2 * pair + side; the two sides may have different sizes (loftr.py:48-56): fc0 [P][L0][256], fc1 [P][L1][256]
__global__ void lf_tokens_kernel(const float* __restrict__ fc0, const float* __restrict__ fc1, const float* __restrict__ pe0,
                                 const float* __restrict__ pe1, float* __restrict__ xm, int L0, int L1, int cap) {
  const int z = blockIdx.y, side = z & 1, L = side ? L1 : L0;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)L * CD) return;
  const int l = (int)(i / CD), c = (int)(i % CD);
  const float* fc = (side ? fc1 : fc0) + (size_t)(z >> 1) * L * CD;
  xm[((size_t)z * cap + l) * 512 + c] = fc[(size_t)l * CD + c] + (side ? pe1 : pe0)[i];
}

// ---- linear attention (linear_attention.py:31-47) ------------------------------------------------------------------------
// projection epilogue: q -> elu+1, k -> elu+1, v -> v / S   (three [rows][dm] buffers)
struct EpiLinAttnQKV {
  float *q, *k, *v; int dm; long long slot_stride; float kv_len;   // values are divided by the source length (:38-39)
  const int* kv_counts = nullptr;                                  // per-slot token count (a slot's values serve as ITS OWN source rows)
  __device__ __forceinline__ float fmap(float x) const { return x > 0.f ? x + 1.f : expf(x); }  // elu(x) + 1
  __device__ __forceinline__ float conv(int which, float x, int z) const {
    return which < 2 ? fmap(x) : __fdiv_rn(x, kv_counts ? (float)kv_counts[z] : kv_len);
  }
  __device__ void operator()(int z, int row, int col, float4 a, int) const {
    const int which = col / dm, c = col % dm;
    float* dst = (which == 0 ? q : which == 1 ? k : v) + z * slot_stride + (long long)row * dm + c;
    *reinterpret_cast<float4*>(dst) = make_float4(conv(which, a.x, z), conv(which, a.y, z), conv(which, a.z, z), conv(which, a.w, z));
  }
  __device__ float2 prefetch(int, int, int) const { return make_float2(0.f, 0.f); }
  __device__ void elem(int z, int row, int col, float a, float2) const {
    const int which = col / dm, c = col % dm;
    (which == 0 ? q : which == 1 ? k : v)[z * slot_stride + (long long)row * dm + c] = conv(which, a, z);
  }
  // pair form (col, dm even)
  __device__ float4 pair_prefetch(int z, int, int) const { return make_float4(kv_counts ? (float)kv_counts[z] : kv_len, 0.f, 0.f, 0.f); }
  __device__ void pair(int z, int row, int col, float a0, float a1, float4 pre) const {
    const int which = col / dm, c = col - which * dm;
    float* dst = (which == 0 ? q : which == 1 ? k : v) + z * slot_stride + (long long)row * dm + c;
    const float2 o = which < 2 ? make_float2(fmap(a0), fmap(a1)) : make_float2(__fdiv_rn(a0, pre.x), __fdiv_rn(a1, pre.x));
    *reinterpret_cast<float2*>(dst) = o;
  }
  // quad form (col % 4 == 0, dm % 4 == 0; tc_gemm.cuh): one 16-byte store per four columns
  __device__ bool quad_ok() const {
    return !(reinterpret_cast<uintptr_t>(q) & 15) && !(reinterpret_cast<uintptr_t>(k) & 15) && !(reinterpret_cast<uintptr_t>(v) & 15) &&
           !(dm & 3) && !(slot_stride & 3);
  }
  __device__ float4 quad_col(int) const { return make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ float4 quad_prefetch(int z, int, int) const { return make_float4(kv_counts ? (float)kv_counts[z] : kv_len, 0.f, 0.f, 0.f); }
  __device__ void quad(int z, int row, int col, float4 a, float4, float4 pre) const {
    const int which = col / dm, c = col - which * dm;
    float* dst = (which == 0 ? q : which == 1 ? k : v) + z * slot_stride + (long long)row * dm + c;
    const float4 o = which < 2 ? make_float4(fmap(a.x), fmap(a.y), fmap(a.z), fmap(a.w))
                               : make_float4(__fdiv_rn(a.x, pre.x), __fdiv_rn(a.y, pre.x), __fdiv_rn(a.z, pre.x), __fdiv_rn(a.w, pre.x));
    *reinterpret_cast<float4*>(dst) = o;
  }
  __device__ bool rowwise(int, int, bool, int, const float (&)[32]) const { return false; }
};

// KV[g][h][d][v] = sum_s K[s][h][d] V[s][h][v],  Ksum[g][h][d] = sum_s K[s][h][d]   (one CTA per group and head)
template <int DH>
__global__ void __launch_bounds__(DH * DH > 256 ? 1024 : 256)
la_kv_kernel(const float* __restrict__ K, const float* __restrict__ V, float* __restrict__ KV, float* __restrict__ Ksum,
             const int* __restrict__ counts, int n_static, long long group_stride, int dm, const int* __restrict__ skip, int group_xor) {
  // gridDim.z > 1: the token range is cut into gridDim.z chunks and KV / Ksum are PARTIAL sums laid out
  // [g][h][split][...] (summed in a fixed order by la_kv_reduce_kernel: deterministic)
  const int g = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
  if (skip && skip[g]) return;
  const int src = g ^ group_xor;                // cross layers: keys / values of the paired group
  const int n_all = counts ? counts[src] : n_static;
  const int chunk = (n_all + gridDim.z - 1) / gridDim.z;
  const int s_begin = blockIdx.z * chunk, n = min(n_all, s_begin + chunk);
  constexpr int NT = DH * DH > 256 ? 1024 : 256;
  constexpr int TS = 64;
  __shared__ float sK[TS][DH + 1], sV[TS][DH + 1];
  const int d = tid / DH, v = tid % DH;
  const bool active = tid < DH * DH;
  float acc = 0.f, ks = 0.f;
  const float* Kb = K + src * group_stride + h * DH;
  const float* Vb = V + src * group_stride + h * DH;
  for (int s0 = s_begin; s0 < n; s0 += TS) {
    __syncthreads();
    for (int i = tid; i < TS * DH; i += NT) {
      const int s = i / DH, c = i % DH;
      const bool ok = s0 + s < n;
      sK[s][c] = ok ? Kb[(long long)(s0 + s) * dm + c] : 0.f;
      sV[s][c] = ok ? Vb[(long long)(s0 + s) * dm + c] : 0.f;
    }
    __syncthreads();
    if (active) {
#pragma unroll 8
      for (int s = 0; s < TS; s++) { acc = fmaf(sK[s][d], sV[s][v], acc); if (v == 0) ks += sK[s][d]; }
    }
  }
  if (active) {
    const long long gh = ((long long)g * gridDim.y + h) * gridDim.z + blockIdx.z;
    KV[(gh * DH + d) * DH + v] = acc;
    if (v == 0) Ksum[gh * DH + d] = ks;
  }
}

// KV[gh][e] = sum_split part[gh][split][e] (fixed order)
__global__ void __launch_bounds__(256) la_kv_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, long long n_gh, int elems,
                                                           int nsplit, const int* __restrict__ skip, int gh_per_group) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n_gh * elems) return;
  const long long gh = i / elems;
  if (skip && skip[gh / gh_per_group]) return;
  const int e = (int)(i % elems);
  float s = 0.f;
  for (int k = 0; k < nsplit; k++) s += part[(gh * nsplit + k) * elems + e];
  out[i] = s;
}

// msg[l][h*DH + v] = (sum_d Q[l][h][d] KV[h][d][v]) / (sum_d Q[l][h][d] Ksum[h][d] + eps) * S
template <int DH>
__global__ void __launch_bounds__(256) la_apply_kernel(const float* __restrict__ Q, const float* __restrict__ KV,
                                                       const float* __restrict__ Ksum, float* __restrict__ out, const int* __restrict__ counts,
                                                       int n_static, long long group_stride, int dm, int out_ld, long long out_group_stride,
                                                       const int* __restrict__ skip, int group_xor) {
  const int g = blockIdx.y, tid = threadIdx.x;
  if (skip && skip[g]) return;
  const int n = counts ? counts[g] : n_static, nsrc = counts ? counts[g ^ group_xor] : n_static;
  constexpr int H = 8, TPB = 256 / DH;  // tokens per block
  __shared__ float sKV[H][DH][DH + 1], sKs[H][DH], sQ[TPB][H * DH];
  for (int i = tid; i < H * DH * DH; i += 256) sKV[i / (DH * DH)][(i / DH) % DH][i % DH] = KV[(long long)g * H * DH * DH + i];
  for (int i = tid; i < H * DH; i += 256) sKs[i / DH][i % DH] = Ksum[(long long)g * H * DH + i];
  const int t = tid / DH, v = tid % DH;
  for (int l0 = blockIdx.x * TPB; l0 < n; l0 += gridDim.x * TPB) {
    __syncthreads();
    for (int i = tid; i < TPB * H * DH; i += 256) {
      const int tt = i / (H * DH), c = i % (H * DH);
      sQ[tt][c] = (l0 + tt < n) ? Q[g * group_stride + (long long)(l0 + tt) * dm + c] : 0.f;
    }
    __syncthreads();
    const int l = l0 + t;
    if (l < n) {
#pragma unroll
      for (int h = 0; h < H; h++) {
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int d = 0; d < DH; d++) { const float q = sQ[t][h * DH + d]; num = fmaf(q, sKV[h][d][v], num); den = fmaf(q, sKs[h][d], den); }
        out[g * out_group_stride + (long long)l * out_ld + h * DH + v] = num * (1.f / (den + 1e-6f)) * (float)nsrc;
      }
    }
  }
}

// out[row] = LayerNorm(in[row]) (+ residual x[row])   rows of DM, warp per row (transformer.py:52,56-58)
template <int DM>
__global__ void __launch_bounds__(256) lf_layernorm_kernel(const float* __restrict__ in, int ld_in, long long in_group_stride,
                                                           float* __restrict__ out, int ld_out, long long out_group_stride,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const int* __restrict__ counts, int n_static, const int* __restrict__ skip,
                                                           int residual) {
  const int g = blockIdx.y, row = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (skip && skip[g]) return;
  if (row >= (counts ? counts[g] : n_static)) return;
  const float* p = in + g * in_group_stride + (long long)row * ld_in;
  float v[DM / 32];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < DM / 32; i++) { v[i] = p[lane + 32 * i]; s += v[i]; }
  const float mean = warp_sum(s) * (1.f / DM);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < DM / 32; i++) { const float d = v[i] - mean; ss += d * d; }
  const float rstd = 1.f / sqrtf(warp_sum(ss) * (1.f / DM) + 1e-5f);
  float* o = out + g * out_group_stride + (long long)row * ld_out;
#pragma unroll
  for (int i = 0; i < DM / 32; i++) {
    const int c = lane + 32 * i;
    const float y = (v[i] - mean) * rstd * gamma[c] + beta[c];
    o[c] = residual ? o[c] + y : y;
  }
}

// plain store epilogue with optional bias / ReLU (the encoder linears are bias-free)
struct EpiPlain {
  float* out; int ldo; long long strideOut; int relu; const float* bias;
  __device__ void operator()(int z, int row, int col, float4 a, int) const {
    float r[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float x = r[j] + (bias ? bias[col + j] : 0.f);
      r[j] = relu ? fmaxf(x, 0.f) : x;
    }
    *reinterpret_cast<float4*>(out + z * strideOut + (long long)row * ldo + col) = make_float4(r[0], r[1], r[2], r[3]);
  }
  __device__ float2 prefetch(int, int, int) const { return make_float2(0.f, 0.f); }
  __device__ void elem(int z, int row, int col, float a, float2) const {
    const float x = a + (bias ? bias[col] : 0.f);
    out[z * strideOut + (long long)row * ldo + col] = relu ? fmaxf(x, 0.f) : x;
  }
  // pair form (col, ldo, strideOut even; `out` 8-byte aligned)
  __device__ float4 pair_prefetch(int, int, int col) const {
    const float2 b = bias ? *reinterpret_cast<const float2*>(bias + col) : make_float2(0.f, 0.f);
    return make_float4(b.x, b.y, 0.f, 0.f);
  }
  __device__ void pair(int z, int row, int col, float a0, float a1, float4 pre) const {
    float x0 = a0 + pre.x, x1 = a1 + pre.y;
    if (relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
    *reinterpret_cast<float2*>(out + z * strideOut + (long long)row * ldo + col) = make_float2(x0, x1);
  }
  // quad form (col % 4 == 0; rows of `out` 16-byte aligned): the bias is fetched once per block
  __device__ bool quad_ok() const { return !(reinterpret_cast<uintptr_t>(out) & 15) && !(ldo & 3) && !(strideOut & 3); }
  __device__ float4 quad_col(int col) const {
    return bias ? make_float4(bias[col], bias[col + 1], bias[col + 2], bias[col + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __device__ float4 quad_prefetch(int, int, int) const { return make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ void quad(int z, int row, int col, float4 a, float4 b, float4) const {
    float4 x = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    if (relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
    *reinterpret_cast<float4*>(out + z * strideOut + (long long)row * ldo + col) = x;
  }
  __device__ bool rowwise(int, int, bool, int, const float (&)[32]) const { return false; }
};
// merge_feat on cat[window feature, coarse context] (fine_preprocess.py:54-57): the context half is the same for the 25
// window positions of a match and is added from a per-match vector
struct EpiMergeFeat {
  float* out; const float* ctx2; long long gstride_out, gstride_ctx;
  __device__ void operator()(int z, int row, int col, float4 a, int) const {
    const float* c = ctx2 + z * gstride_ctx + (long long)(row / 25) * 256 + 128 + col;
    *reinterpret_cast<float4*>(out + z * gstride_out + (long long)row * 256 + col) = make_float4(a.x + c[0], a.y + c[1], a.z + c[2], a.w + c[3]);
  }
};
__global__ void lf_parity_kernel(int* skip_even, int* skip_odd, int S) {
  int z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z < S) { skip_even[z] = (z & 1) == 0; skip_odd[z] = (z & 1) == 1; }
}

// ---- coarse matching ---------------------------------------------------------------------------------------------------
// conf = softmax(sim, 1) * softmax(sim, 2), sim = <f0/16, f1/16> / 0.1  (coarse_matching.py:108-119): two streaming passes
// threshold + border + mutual max, ordered compaction by i (coarse_matching.py:176-195,241-250); one CTA per pair
__global__ void __launch_bounds__(1024) lf_coarse_select_kernel(const float* __restrict__ best_v, const int* __restrict__ best_j,
                                                                int L, int L1, int cap, int hc, int wc, int hc1, int wc1, float thr, int border,
                                                                int* __restrict__ i_ids, int* __restrict__ j_ids,
                                                                float* __restrict__ mconf, int* __restrict__ mcount, int mcap) {
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid % 32, wid = tid / 32;
  __shared__ int s_w[32];
  __shared__ int s_base;
  if (tid == 0) s_base = 0;
  __syncthreads();
  auto inner = [&](int idx, int hh, int ww) { int y = idx / ww, x = idx % ww; return y >= border && y < hh - border && x >= border && x < ww - border; };
  for (int i0 = 0; i0 < L; i0 += 1024) {
    const int i = i0 + tid;
    bool ok = false; int j = 0; float v = 0.f;
    if (i < L) {
      v = best_v[(long long)(2 * p) * cap + i];
      j = best_j[(long long)(2 * p) * cap + i];
      ok = v > thr && (unsigned)j < (unsigned)L1 && inner(i, hc, wc) && inner(j, hc1, wc1) && best_j[(long long)(2 * p + 1) * cap + j] == i;   // NaN rows: j = init
    }
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_w[wid] = __popc(bal);
    __syncthreads();
    int woff = 0, tot = 0;
    for (int w = 0; w < 32; w++) { const int c = s_w[w]; if (w < wid) woff += c; tot += c; }
    const int dst = s_base + woff + __popc(bal & ((1u << lane) - 1u));
    if (ok && dst < mcap) { i_ids[(long long)p * mcap + dst] = i; j_ids[(long long)p * mcap + dst] = j; mconf[(long long)p * mcap + dst] = v; }
    __syncthreads();
    if (tid == 0) s_base += tot;
    __syncthreads();
  }
  if (tid == 0) mcount[p] = min(s_base, mcap);
}

// ---- fine level -----------------------------------------------------------------------------------------------------------
// 5x5 windows of the 1/2-resolution map around the matched coarse cells (F.unfold k=5, stride=4, pad=2 restricted to the
// matches), fine_preprocess.py:40-50.  U[side][m][ww][c], side 0 <- feat_f0 at i_ids, side 1 <- feat_f1 at j_ids.
__global__ void __launch_bounds__(128) lf_gather_windows_kernel(const float* __restrict__ ff0, const float* __restrict__ ff1,
                                                                const int* __restrict__ i_ids, const int* __restrict__ j_ids,
                                                                const int* __restrict__ mcount, float* __restrict__ U, int hf0, int wf0, int wc0,
                                                                int hf1, int wf1, int wc1, int stride, int mcap, int P) {
  const int m = blockIdx.x, side = blockIdx.y, p = blockIdx.z, c = threadIdx.x;
  if (m >= mcount[p]) return;
  const int id = side == 0 ? i_ids[(long long)p * mcap + m] : j_ids[(long long)p * mcap + m];
  const int hf = side ? hf1 : hf0, wf = side ? wf1 : wf0, wc = side ? wc1 : wc0;
  const int cy = id / wc, cx = id % wc;
  const float* f = (side ? ff1 : ff0) + (size_t)p * hf * wf * FD;
  float* o = U + ((((size_t)side * P + p) * mcap + m) * 25) * FD;
  for (int ww = 0; ww < 25; ww++) {
    const int y = cy * stride - 2 + ww / 5, x = cx * stride - 2 + ww % 5;
    o[ww * FD + c] = (y >= 0 && y < hf && x >= 0 && x < wf) ? f[((size_t)y * wf + x) * FD + c] : 0.f;
  }
}
// coarse features of the matched cells: G[side][p][m][256]
__global__ void __launch_bounds__(256) lf_gather_coarse_kernel(const float* __restrict__ xm, const int* __restrict__ i_ids,
                                                               const int* __restrict__ j_ids, const int* __restrict__ mcount,
                                                               float* __restrict__ G, int cap, int mcap, int P) {
  const int m = blockIdx.x, side = blockIdx.y, p = blockIdx.z, c = threadIdx.x;
  if (m >= mcount[p]) return;
  const int id = side == 0 ? i_ids[(long long)p * mcap + m] : j_ids[(long long)p * mcap + m];
  G[(((size_t)side * P + p) * mcap + m) * CD + c] = xm[((size_t)(2 * p + side) * cap + id) * 512 + c];
}
// X[row][0:128] = U[row] Wm[:, :128]^T (already in T) + ctx[row / 25] + bias : done as epilogue `addvec`; here the final
// sub-pixel expectation (fine_matching.py:46-49,66-77): one warp per match
__global__ void __launch_bounds__(256) lf_fine_match_kernel(const float* __restrict__ X /*[2][P][mcap][25][256]*/, const int* __restrict__ i_ids,
                                                            const int* __restrict__ j_ids, const int* __restrict__ mcount,
                                                            float* __restrict__ kpts0, float* __restrict__ kpts1, int wc, int wc1, float scale_c,
                                                            float scale_f, int mcap, int P) {
  const int p = blockIdx.y, m = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (m >= mcount[p]) return;
  const float* f0 = X + ((((size_t)0 * P + p) * mcap + m) * 25 + 12) * 256;
  const float* f1 = X + (((size_t)1 * P + p) * mcap + m) * 25 * 256;
  float sim = -INFINITY;
  if (lane < 25) {
    float s = 0.f;
    for (int c = 0; c < FD; c++) s = fmaf(f0[c], f1[lane * 256 + c], s);
    sim = s * 0.08838834764831845f;  // 1 / sqrt(128)
  }
  const float mx = warp_max(sim);
  const float e = lane < 25 ? expf(sim - mx) : 0.f;
  const float h = __fdiv_rn(e, warp_sum(e));
  const float gx = lane < 25 ? -1.f + 0.5f * (lane % 5) : 0.f, gy = lane < 25 ? -1.f + 0.5f * (lane / 5) : 0.f;
  const float ex = warp_sum(h * gx), ey = warp_sum(h * gy);
  if (lane == 0) {
    const int i = i_ids[(long long)p * mcap + m], j = j_ids[(long long)p * mcap + m];
    float* k0 = kpts0 + ((long long)p * mcap + m) * 2;
    float* k1 = kpts1 + ((long long)p * mcap + m) * 2;
    k0[0] = (i % wc) * scale_c; k0[1] = (i / wc) * scale_c;
    k1[0] = (j % wc1) * scale_c + ex * 2.f * scale_f; k1[1] = (j / wc1) * scale_c + ey * 2.f * scale_f;
  }
}

__global__ void lf_fill_counts_kernel(int* p, int n, int v_even, int v_odd) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (i & 1) ? v_odd : v_even;
}
__global__ void lf_scale_counts_kernel(const int* mcount, int* rows25, int P) {  // rows of the fine token matrices
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * P) rows25[i] = mcount[i % P] * 25;
}

constexpr int LF_KV_SPLIT = 32;   // token chunks per (slot, head) in the coarse K^T V reduction

struct LFBuffers {
  void *pa, *pb, *pc, *pd, *pe2;   // backbone plane scratch
  float *fc, *fc1, *ff, *ff1, *xm, *q, *k, *v, *msg, *tmp, *h, *kv, *ksum, *kv_part, *ksum_part, *rmax, *rsum, *rlog, *best_v;
  int *best_j, *skip_even, *skip_odd, *cntL, *rows25, *fcnt;
  float *U, *G, *ctx, *fx, *fq, *fk, *fv, *fmsg, *ftmp, *fh, *fkv, *fksum;
};
}  // namespace

// The backbone runs one SIDE of the pairs at a time (P images of one size): its plane scratch is sized for P images of the
// larger side -- half of what a 2P-image pass needs -- and the two sides may have different sizes.
size_t lf_carve(Workspace& ws, LFBuffers& b, int P, int H0, int W0, int H1, int W1, int cap, int mcap) {
  const size_t S = 2 * (size_t)P;
  auto px = [](int H, int W, int d) { return (size_t)(H / d) * (W / d); };
  const size_t a2 = px(H0, W0, 2) > px(H1, W1, 2) ? px(H0, W0, 2) : px(H1, W1, 2), a4 = px(H0, W0, 4) > px(H1, W1, 4) ? px(H0, W0, 4) : px(H1, W1, 4);
  const size_t T = S * cap;
  const size_t FT = 2 * (size_t)P * mcap * 25;  // fine tokens
  // ---- alive across phases: backbone outputs, coarse-matching results, small tables
  b.fc = ws.take<float>((size_t)P * px(H0, W0, 8) * CD); b.fc1 = ws.take<float>((size_t)P * px(H1, W1, 8) * CD);
  b.ff = ws.take<float>((size_t)P * px(H0, W0, 2) * FD); b.ff1 = ws.take<float>((size_t)P * px(H1, W1, 2) * FD);
  b.kv = ws.take<float>(S * NH * 32 * 32); b.ksum = ws.take<float>(S * NH * 32);
  b.kv_part = ws.take<float>(S * NH * LF_KV_SPLIT * 32 * 32); b.ksum_part = ws.take<float>(S * NH * LF_KV_SPLIT * 32);
  b.rmax = ws.take<float>(T); b.rsum = ws.take<float>(T); b.rlog = ws.take<float>(T); b.best_v = ws.take<float>(T); b.best_j = ws.take<int>(T);
  b.skip_even = ws.take<int>(S); b.skip_odd = ws.take<int>(S); b.cntL = ws.take<int>(S); b.rows25 = ws.take<int>(S); b.fcnt = ws.take<int>(S * mcap);
  b.xm = ws.take<float>(T * 512);               // coarse features: written by the tokeniser, read until the fine stage gathers its context
  // ---- three phases share the rest (the stream orders them): (1) backbone plane scratch, (2) coarse transformer, (3) fine stage.
  // Round 1 kept all three side by side: 4 GB per 1024 x 1024 pair; 2.5 GB after the per-side backbone passes; 1.4 GB aliased.
  const size_t base = ws.off;
  const size_t big = (size_t)P * a2 * 256 * NP;  // fp16 elements of the largest plane set (1/2 res, 256 padded channels)
  b.pa = ws.take<plane_t>(big); b.pb = ws.take<plane_t>(big); b.pc = ws.take<plane_t>(big);
  b.pd = ws.take<plane_t>((size_t)P * a4 * 256 * NP); b.pe2 = ws.take<plane_t>((size_t)P * a4 * 256 * NP);
  size_t high = ws.off;
  ws.off = base;                                // (2)
  b.q = ws.take<float>(T * CD); b.k = ws.take<float>(T * CD); b.v = ws.take<float>(T * CD);
  b.msg = ws.take<float>(T * CD); b.tmp = ws.take<float>(T * CD); b.h = ws.take<float>(T * 512);
  if (ws.off > high) high = ws.off;
  ws.off = base;                                // (3)
  b.U = ws.take<float>(FT * FD); b.G = ws.take<float>(2 * (size_t)P * mcap * CD); b.ctx = ws.take<float>(2 * (size_t)P * mcap * 256);
  b.fx = ws.take<float>(FT * 256); b.fq = ws.take<float>(FT * FD); b.fk = ws.take<float>(FT * FD); b.fv = ws.take<float>(FT * FD);
  b.fmsg = ws.take<float>(FT * FD); b.ftmp = ws.take<float>(FT * FD); b.fh = ws.take<float>(FT * 256);
  b.fkv = ws.take<float>(2 * (size_t)P * mcap * NH * 16 * 16); b.fksum = ws.take<float>(2 * (size_t)P * mcap * NH * 16);
  if (ws.off > high) high = ws.off;
  ws.off = high;
  if (ws.base && ws.off > ws.size) ws.overflow = true;
  return ws.off;
}

extern "C" size_t imw_loftr_workspace_bytes_hw(int n_pairs, int height0, int width0, int height1, int width1, int max_matches) {
  Workspace ws(nullptr, 0);
  LFBuffers b;
  const int L0 = (height0 / 8) * (width0 / 8), L1 = (height1 / 8) * (width1 / 8), L = L0 > L1 ? L0 : L1, cap = (L + 127) / 128 * 128;
  return lf_carve(ws, b, n_pairs, height0, width0, height1, width1, cap, max_matches) + 256;
}
extern "C" size_t imw_loftr_workspace_bytes(int n_pairs, int height, int width, int max_matches) {
  return imw_loftr_workspace_bytes_hw(n_pairs, height, width, height, width, max_matches);
}

// images [2P][H][W] fp32: slot 2p = the image whose cells index the ROWS of the confidence matrix ("image0" of the LoFTR
// module; hloc passes its image1 there, hloc/matchers/loftr.py:43-51).  H, W multiples of 8.
// Outputs per pair: keypoints0/1 [P][max_matches][2] (in i order), confidence [P][max_matches], counts [P].
namespace {
struct LFGeom { int H, W, h2, w2, h4, w4, hc, wc, L; };
LFGeom lf_geom(int H, int W) { return LFGeom{H, W, H / 2, W / 2, H / 4, W / 4, H / 8, W / 8, (H / 8) * (W / 8)}; }
}  // namespace

// Pairs whose two images have DIFFERENT sizes (the LoFTR module runs the backbone per image then, loftr.py:48-56; hloc's
// `minima_loftr` / `loftr_aachen`-style confs do not force a common size): images0 [P] frames of height0 x width0 (the image whose
// cells index the ROWS of the confidence matrix), images1 [P] frames of height1 x width1; `stride*` = floats between consecutive
// frames of a side (so that both the interleaved [2P][H][W] layout and two separate arrays fit).  weights->pos_enc is the
// position encoding of side 0 ([hc0*wc0][256]), pos_enc1 that of side 1.
extern "C" int imw_loftr_forward_hw(const imw_loftr_weights* W, const imw_loftr_conf* conf, int n_pairs, int height0, int width0,
                                    int height1, int width1, const float* images0, long long stride0, const float* images1,
                                    long long stride1, const float* pos_enc1, int max_matches, float* keypoints0, float* keypoints1,
                                    float* confidence, int* counts, float* dbg_feat_c, float* dbg_backbone_c, void* workspace,
                                    size_t workspace_bytes, cudaStream_t st) {
  IMW_REQUIRE(W && conf && n_pairs > 0 && height0 % 8 == 0 && width0 % 8 == 0 && height1 % 8 == 0 && width1 % 8 == 0,
              "imw_loftr_forward: H, W must be multiples of 8");
  IMW_REQUIRE(max_matches > 0, "imw_loftr_forward: max_matches must be positive");
  IMW_REQUIRE((long long)n_pairs * max_matches <= 65535, "imw_loftr_forward: n_pairs * max_matches must not exceed 65535 (got %d x %d); "
              "split the batch (ops.loftr_forward does)", n_pairs, max_matches);
  const int P = n_pairs, S = 2 * P;
  const LFGeom g0 = lf_geom(height0, width0), g1 = lf_geom(height1, width1);
  const int L0 = g0.L, L1 = g1.L, Lmax = L0 > L1 ? L0 : L1, cap = (Lmax + 127) / 128 * 128, mcap = max_matches;
  const int H = height0, hc = g0.hc, wc = g0.wc, h2 = g0.h2;    // side-0 names used by the scale factors below (8 and 2 on both sides)
  Workspace ws(workspace, workspace_bytes);
  LFBuffers b;
  lf_carve(ws, b, P, height0, width0, height1, width1, cap, mcap);
  if (ws.overflow) { imw_set_error("imw_loftr_forward: workspace too small (%zu < %zu)", workspace_bytes, ws.off); return IMW_ERR_WORKSPACE; }
  int rc;
#define RUN(x) do { rc = (x); if (rc) return rc; } while (0)
  // ---------------- backbone (resnet_fpn.py:100-118), one side at a time; 196-channel tensors are zero-padded to 256 channels ------
  const imw_loftr_backbone& bb = W->backbone;
  auto backbone = [&](const LFGeom& g, const float* images, long long img_stride, float* fc_out, float* ff_out) -> int {
    const int nimg = P, h2 = g.h2, w2 = g.w2, h4 = g.h4, w4 = g.w4, hc = g.hc, wc = g.wc;
    {
      dim3 grid(ceil_div(w2, 16) * ceil_div(h2, 16), 1, nimg);
      lf_conv1_kernel<<<grid, 256, 0, st>>>(images, bb.conv1_w, bb.conv1_b, (plane_t*)b.pa, nimg, (size_t)img_stride, g.H, g.W);
      IMW_CHECK_LAUNCH_T("lf_conv1_kernel");
    }
    auto conv = [&](const void* in, const imw_loftr_conv& c, const void* res, void* out, int Hin, int Win, int act, int f32) {
      return tc_conv_general(in, c.w, c.b, res, out, nimg, Hin, Win, c.cin, c.cout, c.ksize, c.stride, act, f32, st);
    };
    // layer1 (1/2): two BasicBlocks 128 -> 128
    RUN(conv(b.pa, bb.l1[0], nullptr, b.pb, h2, w2, 1, 0));  RUN(conv(b.pb, bb.l1[1], b.pa, b.pc, h2, w2, 1, 0));   // x = relu(x + y)
    RUN(conv(b.pc, bb.l1[2], nullptr, b.pb, h2, w2, 1, 0));  RUN(conv(b.pb, bb.l1[3], b.pc, b.pa, h2, w2, 1, 0));   // x1 = pa
    // layer2 (1/4): 128 -> 196(256)
    void *x2a = b.pd, *x2t = b.pe2;
    RUN(conv(b.pa, bb.l2_down, nullptr, x2a, h2, w2, 0, 0));                                                        // downsample(x1)
    RUN(conv(b.pa, bb.l2[0], nullptr, x2t, h2, w2, 1, 0));   RUN(conv(x2t, bb.l2[1], x2a, b.pb, h4, w4, 1, 0));      // block 0 -> pb (1/4)
    RUN(conv(b.pb, bb.l2[2], nullptr, x2t, h4, w4, 1, 0));   RUN(conv(x2t, bb.l2[3], b.pb, x2a, h4, w4, 1, 0));      // x2 = x2a (pd)
    // layer3 (1/8): 196(256) -> 256 ; reuse pb / pe2 / pc halves as scratch (all large enough)
    void *x3d = b.pb, *x3t = b.pe2, *x3 = b.pc;
    RUN(conv(x2a, bb.l3_down, nullptr, x3d, h4, w4, 0, 0));
    RUN(conv(x2a, bb.l3[0], nullptr, x3t, h4, w4, 1, 0));
    RUN(conv(x3t, bb.l3[1], x3d, x3, hc, wc, 1, 0));       // block 0 second conv reads x3t (1/8) with residual x3d -> x3
    RUN(conv(x3, bb.l3[2], nullptr, x3t, hc, wc, 1, 0));   RUN(conv(x3t, bb.l3[3], x3, x3d, hc, wc, 1, 0));          // x3 = x3d (pb)
    // FPN
    void* x3_out = b.pc;
    RUN(conv(x3d, bb.l3_out, nullptr, fc_out, hc, wc, 0, 1));               // fp32 coarse feature map [nimg][hc][wc][256]
    RUN(conv(x3d, bb.l3_out, nullptr, x3_out, hc, wc, 0, 0));               // same as planes for the upsampling path
    RUN(conv(x2a, bb.l2_out, nullptr, b.pe2, h4, w4, 0, 0));                // layer2_outconv(x2) -> pe2
    {
      const size_t n = (size_t)nimg * h4 * w4 * 256;
      lf_upsample_add_kernel<<<(unsigned)((n / 8 + 255) / 256), 256, 0, st>>>((const plane_t*)b.pe2, (const plane_t*)x3_out,
                                                                         (plane_t*)b.pb, nimg, h4, w4, 256);
      IMW_CHECK_LAUNCH_T("lf_upsample_add_kernel");
    }
    RUN(conv(b.pb, bb.l2_out2[0], nullptr, b.pe2, h4, w4, 2, 0));  RUN(conv(b.pe2, bb.l2_out2[1], nullptr, x2a, h4, w4, 0, 0));  // x2_out -> pd
    RUN(conv(b.pa, bb.l1_out, nullptr, b.pb, h2, w2, 0, 0));                // layer1_outconv(x1) -> pb (256 padded)
    {
      const size_t n = (size_t)nimg * h2 * w2 * 256;
      lf_upsample_add_kernel<<<(unsigned)((n / 8 + 255) / 256), 256, 0, st>>>((const plane_t*)b.pb, (const plane_t*)x2a,
                                                                         (plane_t*)b.pc, nimg, h2, w2, 256);
      IMW_CHECK_LAUNCH_T("lf_upsample_add_kernel");
    }
    RUN(conv(b.pc, bb.l1_out2[0], nullptr, b.pb, h2, w2, 2, 0));   RUN(conv(b.pb, bb.l1_out2[1], nullptr, ff_out, h2, w2, 0, 1));  // fine map fp32
    return IMW_OK;
  };
  RUN(backbone(g0, images0, stride0, b.fc, b.ff));
  RUN(backbone(g1, images1, stride1, b.fc1, b.ff1));
  if (dbg_backbone_c) {   // [2P][Lmax][256] in slot order (side-major maps -> strided copies)
    IMW_CHECK_CUDA(cudaMemcpy2DAsync(dbg_backbone_c, sizeof(float) * 2 * (size_t)Lmax * CD, b.fc, sizeof(float) * (size_t)L0 * CD,
                                     sizeof(float) * (size_t)L0 * CD, P, cudaMemcpyDeviceToDevice, st));
    IMW_CHECK_CUDA(cudaMemcpy2DAsync(dbg_backbone_c + (size_t)Lmax * CD, sizeof(float) * 2 * (size_t)Lmax * CD, b.fc1, sizeof(float) * (size_t)L1 * CD,
                                     sizeof(float) * (size_t)L1 * CD, P, cudaMemcpyDeviceToDevice, st));
  }

  // ---------------- coarse transformer -----------------------------------------------------------------------------------
  lf_tokens_kernel<<<dim3((unsigned)(((size_t)Lmax * CD + 255) / 256), S), 256, 0, st>>>(b.fc, b.fc1, W->pos_enc, pos_enc1 ? pos_enc1 : W->pos_enc,
                                                                                         b.xm, L0, L1, cap);
  IMW_CHECK_LAUNCH_T("lf_tokens_kernel");
  lf_fill_counts_kernel<<<ceil_div(S, 256), 256, 0, st>>>(b.cntL, S, L0, L1);
  IMW_CHECK_LAUNCH_T("lf_fill_counts_kernel");
  // per-slot skip masks for the sequential cross layers: skip_even[z] = (z even), skip_odd[z] = (z odd)
  lf_parity_kernel<<<ceil_div(S, 256), 256, 0, st>>>(b.skip_even, b.skip_odd, S);
  IMW_CHECK_LAUNCH_T("lf_parity_kernel");
  const int use_tc = conf->use_tensor_cores;
  const int kv_split = LF_KV_SPLIT;
  auto tok_linear = [&](const float* A, int lda, const float* Wt, int N, int K, auto epi, const int* skip) -> int {
    if (use_tc) {
      TcGemmArgs t{};
      t.K = K; t.N = N; t.tiles_per_slot = cap / 128; t.counts = b.cntL; t.skip = skip; t.skip_shift = 0;
      if (W->has_f16_planes) { t.w_planes = Wt + (size_t)N * K; t.w_plane_rows = N; }   // [W fp32 ; split-fp16 planes]
      if (use_tc == 2) return launch_tc_gemm<128, 1>(A, (long long)S * cap, lda, Wt, N, t, epi, st);
      return launch_tc_gemm<128, 3>(A, (long long)S * cap, lda, Wt, N, t, epi, st);
    }
    GemmArgs g{};
    g.A = A; g.strideA = (long long)cap * lda; g.lda = lda; g.W = Wt; g.ldw = K; g.M = cap; g.N = N; g.K = K; g.Mdyn = b.cntL;
    g.skip = skip; g.skip_shift = 0;
    IMW_CHECK_CUDA(launch_gemm(g, S, epi, st));
    return IMW_OK;
  };
  // one encoder layer applied to the slots NOT skipped by `skip_q`; keys / values come from slot z ^ kv_xor
  auto encoder = [&](const imw_loftr_layer& ly, const int* skip_q, int kv_xor) -> int {
    // projections for every slot (q of the updated slots, k/v of their sources)
    if (int e = tok_linear(b.xm, 512, ly.qkv_w, 3 * CD, CD, EpiLinAttnQKV{b.q, b.k, b.v, CD, (long long)cap * CD, 0.f, b.cntL}, nullptr)) return e;
    la_kv_kernel<32><<<dim3(S, NH, kv_split), 1024, 0, st>>>(b.k, b.v, b.kv_part, b.ksum_part, b.cntL, 0, (long long)cap * CD, CD, skip_q, kv_xor);
    IMW_CHECK_LAUNCH_T("la_kv_kernel<32>");
    la_kv_reduce_kernel<<<ceil_div(S * NH * 32 * 32, 256), 256, 0, st>>>(b.kv_part, b.kv, (long long)S * NH, 32 * 32, kv_split, skip_q, NH);
    IMW_CHECK_LAUNCH_T("la_kv_reduce_kernel");
    la_kv_reduce_kernel<<<ceil_div(S * NH * 32, 256), 256, 0, st>>>(b.ksum_part, b.ksum, (long long)S * NH, 32, kv_split, skip_q, NH);
    IMW_CHECK_LAUNCH_T("la_kv_reduce_kernel");
    la_apply_kernel<32><<<dim3(ceil_div(Lmax, 8 * 16), S), 256, 0, st>>>(b.q, b.kv, b.ksum, b.msg, b.cntL, 0, (long long)cap * CD, CD, CD,
                                                                    (long long)cap * CD, skip_q, kv_xor);
    IMW_CHECK_LAUNCH_T("la_apply_kernel<32>");
    if (int e = tok_linear(b.msg, CD, ly.merge_w, CD, CD, EpiPlain{b.tmp, CD, (long long)cap * CD, 0, nullptr}, skip_q)) return e;
    lf_layernorm_kernel<CD><<<dim3(ceil_div(Lmax, 8), S), 256, 0, st>>>(b.tmp, CD, (long long)cap * CD, b.xm + CD, 512, (long long)cap * 512,
                                                                   ly.norm1_g, ly.norm1_b, b.cntL, 0, skip_q, 0);
    IMW_CHECK_LAUNCH();
    if (int e = tok_linear(b.xm, 512, ly.mlp0_w, 512, 512, EpiPlain{b.h, 512, (long long)cap * 512, 1, nullptr}, skip_q)) return e;
    if (int e = tok_linear(b.h, 512, ly.mlp2_w, CD, 512, EpiPlain{b.tmp, CD, (long long)cap * CD, 0, nullptr}, skip_q)) return e;
    lf_layernorm_kernel<CD><<<dim3(ceil_div(Lmax, 8), S), 256, 0, st>>>(b.tmp, CD, (long long)cap * CD, b.xm, 512, (long long)cap * 512,
                                                                   ly.norm2_g, ly.norm2_b, b.cntL, 0, skip_q, 1);
    IMW_CHECK_LAUNCH();
    return IMW_OK;
  };
  for (int i = 0; i < W->n_coarse; i++) {
    const imw_loftr_layer& ly = W->coarse[i];
    if (!ly.is_cross) { RUN(encoder(ly, nullptr, 0)); }
    else { RUN(encoder(ly, b.skip_odd, 1)); RUN(encoder(ly, b.skip_even, 1)); }  // feat0 first, feat1 sees the new feat0
  }
  if (dbg_feat_c) {
    IMW_CHECK_CUDA(cudaMemcpy2DAsync(dbg_feat_c, CD * sizeof(float), b.xm, 512 * sizeof(float), CD * sizeof(float), (size_t)S * cap,
                                     cudaMemcpyDeviceToDevice, st));
  }
  // ---------------- coarse matching --------------------------------------------------------------------------------------
  {
    SimArgs sa{b.xm, cap, 512, CD, b.cntL, nullptr};
    const float scale = 1.f / (16.f * 16.f * conf->temperature);
    if (conf->use_tensor_cores && tc_simreduce_ok(sa)) {
      plane_t* planes = reinterpret_cast<plane_t*>(b.q);   // q is dead after the coarse transformer: [2][S * cap][CD] fp16 fits
      RUN(tc_simreduce_split(sa, S, planes, st));
      RUN(launch_tc_simreduce(sa, S, planes, OpSoftmaxStats{b.rmax, b.rsum, b.rlog, cap, scale}, st));
      RUN(launch_tc_simreduce(sa, S, planes, OpDualSoftmaxArgmax{b.rmax, b.rsum, b.rlog, b.best_v, b.best_j, cap, scale}, st));
    } else {
      IMW_CHECK_CUDA(launch_simreduce(sa, S, OpSoftmaxStats{b.rmax, b.rsum, b.rlog, cap, scale}, st));
      IMW_CHECK_CUDA(launch_simreduce(sa, S, OpDualSoftmaxArgmax{b.rmax, b.rsum, b.rlog, b.best_v, b.best_j, cap, scale}, st));
    }
    int* i_ids = b.fcnt;               // [P][mcap]
    int* j_ids = b.fcnt + (size_t)P * mcap;
    lf_coarse_select_kernel<<<P, 1024, 0, st>>>(b.best_v, b.best_j, L0, L1, cap, g0.hc, g0.wc, g1.hc, g1.wc, conf->match_threshold, conf->border_rm, i_ids, j_ids,
                                                confidence, counts, mcap);
    IMW_CHECK_LAUNCH_T("lf_coarse_select_kernel");
    // ---------------- fine level ---------------------------------------------------------------------------------------
    const int stride_f = h2 / hc;
    lf_gather_windows_kernel<<<dim3(mcap, 2, P), FD, 0, st>>>(b.ff, b.ff1, i_ids, j_ids, counts, b.U, g0.h2, g0.w2, g0.wc, g1.h2, g1.w2, g1.wc,
                                                              stride_f, mcap, P);
    IMW_CHECK_LAUNCH_T("lf_gather_windows_kernel");
    lf_gather_coarse_kernel<<<dim3(mcap, 2, P), CD, 0, st>>>(b.xm, i_ids, j_ids, counts, b.G, cap, mcap, P);
    IMW_CHECK_LAUNCH_T("lf_gather_coarse_kernel");
    lf_scale_counts_kernel<<<ceil_div(2 * P, 128), 128, 0, st>>>(counts, b.rows25, P);
    IMW_CHECK_LAUNCH_T("lf_scale_counts_kernel");
    const int G2 = 2 * P;  // (side, pair) groups of the fine stage: g = side * P + p
    // ctx = down_proj(coarse feats) (fine_preprocess.py:51-53), then its share of merge_feat: Wm[:, 128:] ctx + b.
    // All mcap rows are computed (rows beyond the match count are never read back).
    {
      GemmArgs g{};
      g.A = b.G; g.strideA = (long long)mcap * CD; g.lda = CD; g.W = W->down_proj_w; g.ldw = CD; g.M = mcap; g.N = FD; g.K = CD;
      IMW_CHECK_CUDA(launch_gemm(g, G2, EpiPlain{b.ctx, 256, (long long)mcap * 256, 0, W->down_proj_b}, st));
      GemmArgs g2{};
      g2.A = b.ctx; g2.strideA = (long long)mcap * 256; g2.lda = 256; g2.W = W->merge_feat_w + FD; g2.ldw = 256; g2.M = mcap; g2.N = FD; g2.K = FD;
      IMW_CHECK_CUDA(launch_gemm(g2, G2, EpiPlain{b.ctx + FD, 256, (long long)mcap * 256, 0, W->merge_feat_b}, st));
    }
    {  // fx[:, 0:128] = U Wm[:, :128]^T + ctx2[row / 25]
      GemmArgs g{};
      g.A = b.U; g.strideA = (long long)mcap * 25 * FD; g.lda = FD; g.W = W->merge_feat_w; g.ldw = 256; g.M = mcap * 25; g.N = FD; g.K = FD;
      g.Mdyn = b.rows25;
      IMW_CHECK_CUDA(launch_gemm(g, G2, EpiMergeFeat{b.fx, b.ctx, (long long)mcap * 25 * 256, (long long)mcap * 256}, st));
    }
    // fine transformer: groups of 25 tokens; group index = (side * P + p) * mcap + m; rows per side-pair block = rows25
    const long long fgs = (long long)mcap * 25;   // tokens per (side, pair)
    auto fine_linear = [&](const float* A, int lda, const float* Wt, int N, int K, auto epi, int side_lo, int side_hi) -> int {
      GemmArgs g{};
      g.A = A + (long long)side_lo * P * fgs * lda; g.strideA = fgs * lda; g.lda = lda; g.W = Wt; g.ldw = K; g.M = mcap * 25; g.N = N; g.K = K;
      g.Mdyn = b.rows25 + side_lo * P;
      IMW_CHECK_CUDA(launch_gemm(g, (side_hi - side_lo) * P, epi, st));
      return IMW_OK;
    };
    auto fine_encoder = [&](const imw_loftr_layer& ly, int q_lo, int q_hi, int cross) -> int {
      // q/k/v of all tokens of both sides
      if (int e = fine_linear(b.fx, 256, ly.qkv_w, 3 * FD, FD, EpiLinAttnQKV{b.fq, b.fk, b.fv, FD, fgs * FD, 25.f}, 0, 2)) return e;
      for (int side = q_lo; side < q_hi; side++) {
        const int src = cross ? 1 - side : side;
        // one CTA per (match, head): treat every match as a group of 25 tokens
        la_kv_kernel<16><<<dim3((unsigned)(P * mcap), NH), 256, 0, st>>>(b.fk + (long long)src * P * fgs * FD, b.fv + (long long)src * P * fgs * FD,
                                                                       b.fkv + (long long)side * P * mcap * NH * 256, b.fksum + (long long)side * P * mcap * NH * 16,
                                                                       nullptr, 25, 25LL * FD, FD, nullptr, 0);
        IMW_CHECK_LAUNCH_T("la_kv_kernel<16>");
        la_apply_kernel<16><<<dim3(2, (unsigned)(P * mcap)), 256, 0, st>>>(b.fq + (long long)side * P * fgs * FD, b.fkv + (long long)side * P * mcap * NH * 256,
                                                                        b.fksum + (long long)side * P * mcap * NH * 16, b.fmsg + (long long)side * P * fgs * FD,
                                                                        nullptr, 25, 25LL * FD, FD, FD, 25LL * FD, nullptr, 0);
        IMW_CHECK_LAUNCH_T("la_apply_kernel<16>");
      }
      if (int e = fine_linear(b.fmsg, FD, ly.merge_w, FD, FD, EpiPlain{b.ftmp + (long long)q_lo * P * fgs * FD, FD, fgs * FD, 0, nullptr}, q_lo, q_hi)) return e;
      lf_layernorm_kernel<FD><<<dim3(ceil_div(mcap * 25, 8), (q_hi - q_lo) * P), 256, 0, st>>>(
          b.ftmp + (long long)q_lo * P * fgs * FD, FD, fgs * FD, b.fx + (long long)q_lo * P * fgs * 256 + FD, 256, fgs * 256, ly.norm1_g, ly.norm1_b,
          b.rows25 + q_lo * P, 0, nullptr, 0);
      IMW_CHECK_LAUNCH();
      if (int e = fine_linear(b.fx, 256, ly.mlp0_w, 256, 256, EpiPlain{b.fh + (long long)q_lo * P * fgs * 256, 256, fgs * 256, 1, nullptr}, q_lo, q_hi)) return e;
      if (int e = fine_linear(b.fh, 256, ly.mlp2_w, FD, 256, EpiPlain{b.ftmp + (long long)q_lo * P * fgs * FD, FD, fgs * FD, 0, nullptr}, q_lo, q_hi)) return e;
      lf_layernorm_kernel<FD><<<dim3(ceil_div(mcap * 25, 8), (q_hi - q_lo) * P), 256, 0, st>>>(
          b.ftmp + (long long)q_lo * P * fgs * FD, FD, fgs * FD, b.fx + (long long)q_lo * P * fgs * 256, 256, fgs * 256, ly.norm2_g, ly.norm2_b,
          b.rows25 + q_lo * P, 0, nullptr, 1);
      IMW_CHECK_LAUNCH();
      return IMW_OK;
    };
    for (int i = 0; i < W->n_fine; i++) {
      const imw_loftr_layer& ly = W->fine[i];
      if (!ly.is_cross) { RUN(fine_encoder(ly, 0, 2, 0)); }
      else { RUN(fine_encoder(ly, 0, 1, 1)); RUN(fine_encoder(ly, 1, 2, 1)); }
    }
    lf_fine_match_kernel<<<dim3(ceil_div(mcap, 8), P), 256, 0, st>>>(b.fx, i_ids, j_ids, counts, keypoints0, keypoints1, g0.wc, g1.wc, (float)H / hc,
                                                                    (float)H / h2, mcap, P);
    IMW_CHECK_LAUNCH_T("lf_fine_match_kernel");
  }
#undef RUN
  return IMW_OK;
}

// images [2P][H][W] fp32 (slot 2p+side, both sides of one size): the interleaved layout is side 0 / side 1 with a 2 H W stride
extern "C" int imw_loftr_forward(const imw_loftr_weights* W, const imw_loftr_conf* conf, int n_pairs, int height, int width,
                                 const float* images, int max_matches, float* keypoints0, float* keypoints1, float* confidence,
                                 int* counts, float* dbg_feat_c, float* dbg_backbone_c, void* workspace, size_t workspace_bytes,
                                 cudaStream_t st) {
  const long long hw = (long long)height * width;
  return imw_loftr_forward_hw(W, conf, n_pairs, height, width, height, width, images, 2 * hw, images ? images + hw : nullptr, 2 * hw, nullptr,
                              max_matches, keypoints0, keypoints1, confidence, counts, dbg_feat_c, dbg_backbone_c, workspace,
                              workspace_bytes, st);
}
