// Split-precision operand format of the tcgen05 convolutions: every fp32 value x travels as TWO fp16 planes
//     hi = fp16(x)                 (11 significant bits)
//     lo = fp16((x - hi) * 2^11)   (the next 11 bits, pre-scaled so that they sit in fp16's normal range)
// x = hi + lo * 2^-11 to 2^-22 relative (|x| < 65504; absolute floor 2^-36 for values in fp16's subnormal range).
// A product a*b is assembled from THREE partial products  a_hi b_hi + (a_hi b_lo + a_lo b_hi) * 2^-11  (the dropped
// a_lo b_lo term is <= 2^-22 |a||b|): the scaled cross terms accumulate in their own TMEM accumulator and are folded in with
// one multiply in the epilogue.  Versus three bf16 planes / six products this halves the tensor-core work, the shared-memory
// operand reads and the activation bytes (4 B per element, like fp32), at an operand error (measured on SuperPoint: 2-4e-6
// on the dense score map) of the same size as the reference's own fp32 accumulation noise (1.2-2e-6) -- keypoint sets stay
// identical (tools/split_precision_probe.py).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

typedef __half plane_t;
constexpr int NP = 2;                       // planes per tensor
constexpr float PLANE_LO_SCALE = 2048.f;    // 2^11
constexpr float PLANE_LO_INV = 1.f / 2048.f;

__device__ __forceinline__ void split2(float x, plane_t& hi, plane_t& lo) {
  x = fminf(fmaxf(x, -65504.f), 65504.f);   // fp16 range (never reached by the networks on the path; keeps inf/NaN out)
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * PLANE_LO_SCALE);
}
// two values at once: packed saturating converts (F2FP.SATFINITE: the fp16 range clamp costs nothing).  Same results as split2 on
// each value for |x| <= 65504; beyond it both planes saturate (finite either way; never reached by the networks on the path).
__device__ __forceinline__ uint32_t pack_f16x2_sat(float x0, float x1) {   // (lo half = x0, hi half = x1)
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0));
  return r;
}
__device__ __forceinline__ void split2x2(float x0, float x1, __half2& hi, __half2& lo) {
  const uint32_t h = pack_f16x2_sat(x0, x1);
  hi = *reinterpret_cast<const __half2*>(&h);
  const float2 hf = __half22float2(hi);
  const uint32_t l = pack_f16x2_sat((x0 - hf.x) * PLANE_LO_SCALE, (x1 - hf.y) * PLANE_LO_SCALE);
  lo = *reinterpret_cast<const __half2*>(&l);
}
__device__ __forceinline__ float merge2(plane_t hi, plane_t lo) { return fmaf(__half2float(lo), PLANE_LO_INV, __half2float(hi)); }
