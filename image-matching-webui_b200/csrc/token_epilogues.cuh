// Functor epilogues of the token GEMMs shared by the LightGlue and SuperGlue matchers (CUDA-core operator()
// form: 4 columns per call; tcgen05 form: elem()/prefetch()/rowwise(), see tc_gemm.cuh).
#pragma once
#include "common.cuh"
#include "split_planes.cuh"

namespace {
constexpr int D = 256, HEADS = 4, HD = 64, NF = 32;

// ---- GEMM epilogues ---------------------------------------------------------------------------------
// Self-attention projection: columns [q | k | v] x [head][dim]; rotary on q,k (lightglue.py:58-65,
// 165-169).  Output buffers [slots][HEADS][cap][HD].
// two adjacent elements (off even): one 4-byte store per plane
__device__ __forceinline__ void store_planes2(float* base, long long plane, long long off, float x0, float x1) {
  __half2 hi, lo;
  split2x2(x0, x1, hi, lo);
  plane_t* p = reinterpret_cast<plane_t*>(base);
  *reinterpret_cast<__half2*>(p + off) = hi;
  *reinterpret_cast<__half2*>(p + plane + off) = lo;
}
// tcgen05 attention operands: two fp16 planes per tensor (split_planes.cuh), `plane` elements apart
__device__ __forceinline__ void store_planes(float* base, long long plane, long long off, float x) {
  plane_t hi, lo;
  split2(x, hi, lo);
  plane_t* p = reinterpret_cast<plane_t*>(base);
  p[off] = hi; p[plane + off] = lo;
}

struct EpiQKVRotary {
  float *q, *k, *v; const float* bias; const float* enc; int cap;
  long long plane;   // tcgen05 attention layout: elements per hi/lo plane (0 = CUDA-core layout, no planes)
  __device__ void operator()(int z, int row, int col, float4 a, int) const {
    float r[4] = {a.x + bias[col], a.y + bias[col + 1], a.z + bias[col + 2], a.w + bias[col + 3]};
    int which = col / D, c = col % D, head = c / HD, d = c % HD;
    float* dst = (which == 0 ? q : which == 1 ? k : v) + (((long long)z * HEADS + head) * cap + row) * HD + d;
    if (which < 2 && enc) {
      const float* e = enc + ((long long)z * cap + row) * 64;
      float c0 = e[d / 2], c1 = e[d / 2 + 1], s0 = e[32 + d / 2], s1 = e[32 + d / 2 + 1];
      float o0 = __fadd_rn(__fmul_rn(r[0], c0), __fmul_rn(-r[1], s0));
      float o1 = __fadd_rn(__fmul_rn(r[1], c0), __fmul_rn(r[0], s0));
      float o2 = __fadd_rn(__fmul_rn(r[2], c1), __fmul_rn(-r[3], s1));
      float o3 = __fadd_rn(__fmul_rn(r[3], c1), __fmul_rn(r[2], s1));
      r[0] = o0; r[1] = o1; r[2] = o2; r[3] = o3;
    }
    *reinterpret_cast<float4*>(dst) = make_float4(r[0], r[1], r[2], r[3]);
  }
  // tcgen05 epilogue form (lanes along columns; the rotary partner of dim d is the neighbouring lane d^1)
  __device__ float2 prefetch(int z, int row, int col) const {  // (cos, sin) of this lane's rotary frequency
    if (col >= 2 * D || !enc) return make_float2(1.f, 0.f);
    const float* e = enc + ((long long)z * cap + row) * 64 + (col % HD) / 2;
    return make_float2(e[0], e[32]);
  }
  __device__ void elem(int z, int row, int col, float a, float2 cs_sn) const {
    float r = a + bias[col];
    int which = col / D, c = col % D, head = c / HD, d = c % HD;
    float partner = __shfl_xor_sync(0xffffffffu, r, 1);
    if (which < 2) {
      const float cs = cs_sn.x, sn = cs_sn.y;
      r = (d & 1) ? __fadd_rn(__fmul_rn(r, cs), __fmul_rn(partner, sn)) : __fadd_rn(__fmul_rn(r, cs), __fmul_rn(-partner, sn));
    }
    float* dst = (which == 0 ? q : which == 1 ? k : v);
    const long long off = (((long long)z * HEADS + head) * cap + row) * HD + d;
    if (plane) store_planes(dst, plane, off, r);
    else dst[off] = r;
  }
  // pair form (col even): the rotary pair (d, d + 1) lives in one lane; (cos, sin, bias_d, bias_d+1) prefetched
  __device__ float4 pair_prefetch(int z, int row, int col) const {
    float2 cs = make_float2(1.f, 0.f);
    if (col < 2 * D && enc) {
      const float* e = enc + ((long long)z * cap + row) * 64 + (col % HD) / 2;
      cs = make_float2(e[0], e[32]);
    }
    const float2 b = *reinterpret_cast<const float2*>(bias + col);
    return make_float4(cs.x, cs.y, b.x, b.y);
  }
  __device__ void pair(int z, int row, int col, float a0, float a1, float4 pre) const {
    float r0 = a0 + pre.z, r1 = a1 + pre.w;
    const int which = col / D, c = col % D, head = c / HD, d = c % HD;
    if (which < 2) {   // same operations as elem(): even dim r cs - partner sn, odd dim r cs + partner sn
      const float o0 = __fadd_rn(__fmul_rn(r0, pre.x), __fmul_rn(-r1, pre.y));
      const float o1 = __fadd_rn(__fmul_rn(r1, pre.x), __fmul_rn(r0, pre.y));
      r0 = o0; r1 = o1;
    }
    float* dst = (which == 0 ? q : which == 1 ? k : v);
    const long long off = (((long long)z * HEADS + head) * cap + row) * HD + d;
    if (plane) store_planes2(dst, plane, off, r0, r1);
    else *reinterpret_cast<float2*>(dst + off) = make_float2(r0, r1);
  }
  // quad form (col % 4 == 0): two rotary pairs per lane; (cos_0, sin_0, cos_1, sin_1) prefetched, the bias once per block.  Same
  // operations per element as pair().
  __device__ bool quad_ok() const {
    return !(reinterpret_cast<uintptr_t>(q) & 15) && !(reinterpret_cast<uintptr_t>(k) & 15) && !(reinterpret_cast<uintptr_t>(v) & 15) &&
           !(reinterpret_cast<uintptr_t>(enc) & 7);
  }
  __device__ float4 quad_col(int col) const { return make_float4(bias[col], bias[col + 1], bias[col + 2], bias[col + 3]); }
  __device__ float4 quad_prefetch(int z, int row, int col) const {
    if (col >= 2 * D || !enc) return make_float4(1.f, 0.f, 1.f, 0.f);
    const float* e = enc + ((long long)z * cap + row) * 64 + (col % HD) / 2;
    const float2 cs = *reinterpret_cast<const float2*>(e), sn = *reinterpret_cast<const float2*>(e + 32);
    return make_float4(cs.x, sn.x, cs.y, sn.y);
  }
  __device__ void quad(int z, int row, int col, float4 a, float4 b, float4 pre) const {
    float r0 = a.x + b.x, r1 = a.y + b.y, r2 = a.z + b.z, r3 = a.w + b.w;
    const int which = col / D, c = col % D, head = c / HD, d = c % HD;
    if (which < 2) {
      const float o0 = __fadd_rn(__fmul_rn(r0, pre.x), __fmul_rn(-r1, pre.y));
      const float o1 = __fadd_rn(__fmul_rn(r1, pre.x), __fmul_rn(r0, pre.y));
      const float o2 = __fadd_rn(__fmul_rn(r2, pre.z), __fmul_rn(-r3, pre.w));
      const float o3 = __fadd_rn(__fmul_rn(r3, pre.z), __fmul_rn(r2, pre.w));
      r0 = o0; r1 = o1; r2 = o2; r3 = o3;
    }
    float* dst = (which == 0 ? q : which == 1 ? k : v);
    const long long off = (((long long)z * HEADS + head) * cap + row) * HD + d;
    if (plane) {
      __align__(8) __half2 hi[2], lo[2];
      split2x2(r0, r1, hi[0], lo[0]); split2x2(r2, r3, hi[1], lo[1]);
      plane_t* p = reinterpret_cast<plane_t*>(dst);
      *reinterpret_cast<uint2*>(p + off) = *reinterpret_cast<const uint2*>(hi);
      *reinterpret_cast<uint2*>(p + plane + off) = *reinterpret_cast<const uint2*>(lo);
    } else {
      *reinterpret_cast<float4*>(dst + off) = make_float4(r0, r1, r2, r3);
    }
  }
  // tcgen05 attention wants V transposed ([head][d][token], tokens contiguous): taken straight from the
  // thread-per-row TMEM layout (lanes = consecutive tokens -> coalesced), before the epilogue transpose.
  __device__ bool rowwise(int z, int row, bool valid, int col0, const float (&a)[32]) const {
    if (!plane || col0 < 2 * D) return false;
    const int c = col0 - 2 * D, head = c / HD, d0 = c % HD;
    if (valid) {
#pragma unroll
      for (int j = 0; j < 32; j++)
        store_planes(v, plane, (((long long)z * HEADS + head) * HD + d0 + j) * cap + row, a[j] + bias[col0 + j]);
    }
    return true;
  }
};

// Cross-attention projection: columns [qk | v] x [head][dim]; qk scaled by dim_head^-0.25
// (lightglue.py:216: each side multiplied by scale**0.5).
struct EpiCrossQKV {
  float *qk, *v; const float* bias; int cap; float qk_scale;
  long long plane;
  __device__ void operator()(int z, int row, int col, float4 a, int) const {
    float r[4] = {a.x + bias[col], a.y + bias[col + 1], a.z + bias[col + 2], a.w + bias[col + 3]};
    int which = col / D, c = col % D, head = c / HD, d = c % HD;
    float* dst = (which == 0 ? qk : v) + (((long long)z * HEADS + head) * cap + row) * HD + d;
    if (which == 0) { r[0] *= qk_scale; r[1] *= qk_scale; r[2] *= qk_scale; r[3] *= qk_scale; }
    *reinterpret_cast<float4*>(dst) = make_float4(r[0], r[1], r[2], r[3]);
  }
  __device__ float2 prefetch(int, int, int) const { return make_float2(0.f, 0.f); }
  __device__ void elem(int z, int row, int col, float a, float2) const {
    float r = a + bias[col];
    int which = col / D, c = col % D, head = c / HD, d = c % HD;
    if (which == 0) r *= qk_scale;
    float* dst = (which == 0 ? qk : v);
    const long long off = (((long long)z * HEADS + head) * cap + row) * HD + d;
    if (plane) store_planes(dst, plane, off, r);
    else dst[off] = r;
  }
  __device__ float4 pair_prefetch(int, int, int col) const {
    const float2 b = *reinterpret_cast<const float2*>(bias + col);
    return make_float4(b.x, b.y, 0.f, 0.f);
  }
  __device__ void pair(int z, int row, int col, float a0, float a1, float4 pre) const {
    float r0 = a0 + pre.x, r1 = a1 + pre.y;
    const int which = col / D, c = col % D, head = c / HD, d = c % HD;
    if (which == 0) { r0 *= qk_scale; r1 *= qk_scale; }
    float* dst = (which == 0 ? qk : v);
    const long long off = (((long long)z * HEADS + head) * cap + row) * HD + d;
    if (plane) store_planes2(dst, plane, off, r0, r1);
    else *reinterpret_cast<float2*>(dst + off) = make_float2(r0, r1);
  }
  // quad form (col % 4 == 0), see EpiQKVRotary
  __device__ bool quad_ok() const { return !(reinterpret_cast<uintptr_t>(qk) & 15) && !(reinterpret_cast<uintptr_t>(v) & 15); }
  __device__ float4 quad_col(int col) const { return make_float4(bias[col], bias[col + 1], bias[col + 2], bias[col + 3]); }
  __device__ float4 quad_prefetch(int, int, int) const { return make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ void quad(int z, int row, int col, float4 a, float4 b, float4) const {
    float r0 = a.x + b.x, r1 = a.y + b.y, r2 = a.z + b.z, r3 = a.w + b.w;
    const int which = col / D, c = col % D, head = c / HD, d = c % HD;
    if (which == 0) { r0 *= qk_scale; r1 *= qk_scale; r2 *= qk_scale; r3 *= qk_scale; }
    float* dst = (which == 0 ? qk : v);
    const long long off = (((long long)z * HEADS + head) * cap + row) * HD + d;
    if (plane) {
      __align__(8) __half2 hi[2], lo[2];
      split2x2(r0, r1, hi[0], lo[0]); split2x2(r2, r3, hi[1], lo[1]);
      plane_t* p = reinterpret_cast<plane_t*>(dst);
      *reinterpret_cast<uint2*>(p + off) = *reinterpret_cast<const uint2*>(hi);
      *reinterpret_cast<uint2*>(p + plane + off) = *reinterpret_cast<const uint2*>(lo);
    } else {
      *reinterpret_cast<float4*>(dst + off) = make_float4(r0, r1, r2, r3);
    }
  }
  __device__ bool rowwise(int z, int row, bool valid, int col0, const float (&a)[32]) const {
    if (!plane || col0 < D) return false;
    const int c = col0 - D, head = c / HD, d0 = c % HD;
    if (valid) {
#pragma unroll
      for (int j = 0; j < 32; j++)
        store_planes(v, plane, (((long long)z * HEADS + head) * HD + d0 + j) * cap + row, a[j] + bias[col0 + j]);
    }
    return true;
  }
};

// out[z][row][col] (+)= acc + bias  -- plain / residual variants, N % 4 == 0
struct EpiStore {
  float* out; int ldo; long long strideOut; const float* bias; int residual;
  int relu = 0;        // SuperGlue MLP: ReLU after the (BatchNorm-folded) first layer
  float scale = 1.f;   // applied to acc before the bias (SuperGlue score matrix: 1/sqrt(256))
  plane_t* planes = nullptr;    // optional second output: the same values as split-fp16 planes (same [z][row][col] geometry as `out`,
  long long plane_elems = 0;    //   lo plane plane_elems elements behind hi): the operand of the next linear, written by its producer
  __device__ void operator()(int z, int row, int col, float4 a, int) const {
    float4* o = reinterpret_cast<float4*>(out + z * strideOut + (long long)row * ldo + col);
    float4 r = make_float4(a.x * scale + (bias ? bias[col] : 0.f), a.y * scale + (bias ? bias[col + 1] : 0.f),
                           a.z * scale + (bias ? bias[col + 2] : 0.f), a.w * scale + (bias ? bias[col + 3] : 0.f));
    if (relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
    if (residual) { float4 x = *o; r.x += x.x; r.y += x.y; r.z += x.z; r.w += x.w; }
    *o = r;
  }
  __device__ float2 prefetch(int z, int row, int col) const {
    return make_float2(residual ? out[z * strideOut + (long long)row * ldo + col] : 0.f, 0.f);
  }
  __device__ void elem(int z, int row, int col, float a, float2 res) const {
    float r = a * scale + (bias ? bias[col] : 0.f);
    if (relu) r = fmaxf(r, 0.f);
    const long long off = z * strideOut + (long long)row * ldo + col;
    out[off] = r + res.x;
    if (planes) split2(r + res.x, planes[off], planes[plane_elems + off]);
  }
  // pair form (col, ldo, strideOut even; `out` 8-byte aligned): (residual_0, residual_1, bias_0, bias_1) prefetched
  __device__ float4 pair_prefetch(int z, int row, int col) const {
    float2 r = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
    if (residual) r = *reinterpret_cast<const float2*>(out + z * strideOut + (long long)row * ldo + col);
    if (bias) b = *reinterpret_cast<const float2*>(bias + col);
    return make_float4(r.x, r.y, b.x, b.y);
  }
  __device__ void pair(int z, int row, int col, float a0, float a1, float4 pre) const {
    float r0 = a0 * scale + pre.z, r1 = a1 * scale + pre.w;
    if (relu) { r0 = fmaxf(r0, 0.f); r1 = fmaxf(r1, 0.f); }
    const long long off = z * strideOut + (long long)row * ldo + col;
    *reinterpret_cast<float2*>(out + off) = make_float2(r0 + pre.x, r1 + pre.y);
    if (planes) {
      __half2 hi, lo;
      split2x2(r0 + pre.x, r1 + pre.y, hi, lo);
      *reinterpret_cast<__half2*>(planes + off) = hi;
      *reinterpret_cast<__half2*>(planes + plane_elems + off) = lo;
    }
  }
  // quad form (col % 4 == 0; rows of `out` 16-byte aligned -- quad_ok()): four adjacent columns per lane, eight column quads x four
  // rows per warp instruction.  The per-column constants (bias) are fetched once per 32 x 32 block instead of once per row, and one
  // 16-byte load / store (plus one 8-byte store per plane) replaces two of the pair form: ~40 % fewer instructions per block, and the
  // epilogue warps' issue slots are what bounds this GEMM (tc_gemm.cuh).
  __device__ bool quad_ok() const {
    return !(reinterpret_cast<uintptr_t>(out) & 15) && !(ldo & 3) && !(strideOut & 3) && !(reinterpret_cast<uintptr_t>(planes) & 7);
  }
  __device__ float4 quad_col(int col) const {
    return bias ? make_float4(bias[col], bias[col + 1], bias[col + 2], bias[col + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __device__ float4 quad_prefetch(int z, int row, int col) const {
    return residual ? *reinterpret_cast<const float4*>(out + z * strideOut + (long long)row * ldo + col) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __device__ void quad(int z, int row, int col, float4 a, float4 b, float4 res) const {
    float4 r = make_float4(a.x * scale + b.x, a.y * scale + b.y, a.z * scale + b.z, a.w * scale + b.w);
    if (relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
    r.x += res.x; r.y += res.y; r.z += res.z; r.w += res.w;
    const long long off = z * strideOut + (long long)row * ldo + col;
    *reinterpret_cast<float4*>(out + off) = r;
    if (planes) {
      __align__(8) __half2 hi[2], lo[2];
      split2x2(r.x, r.y, hi[0], lo[0]); split2x2(r.z, r.w, hi[1], lo[1]);
      *reinterpret_cast<uint2*>(planes + off) = *reinterpret_cast<const uint2*>(hi);
      *reinterpret_cast<uint2*>(planes + plane_elems + off) = *reinterpret_cast<const uint2*>(lo);
    }
  }
  __device__ bool rowwise(int, int, bool, int, const float (&)[32]) const { return false; }
};

// final_proj of the layer the pair stopped at; output divided by d^0.25 = 4 (lightglue.py:288-290)
struct EpiFinalProj {
  float* out; int cap; const float* bias_all; const int* stop;
  __device__ void operator()(int z, int row, int col, float4 a, int) const {
    const float* b = bias_all + (stop[z >> 1] - 1) * D;
    float4 r = make_float4((a.x + b[col]) * 0.25f, (a.y + b[col + 1]) * 0.25f, (a.z + b[col + 2]) * 0.25f, (a.w + b[col + 3]) * 0.25f);
    *reinterpret_cast<float4*>(out + ((long long)z * cap + row) * D + col) = r;
  }
  __device__ float2 prefetch(int, int, int) const { return make_float2(0.f, 0.f); }
  __device__ void elem(int z, int row, int col, float a, float2) const {
    out[((long long)z * cap + row) * D + col] = (a + bias_all[(stop[z >> 1] - 1) * D + col]) * 0.25f;
  }
  __device__ bool rowwise(int, int, bool, int, const float (&)[32]) const { return false; }
};

}  // namespace
