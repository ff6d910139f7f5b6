// SuperPoint detector / descriptor post-processing (HBM-bound integer/compare work).
// Follows third_party/SuperGluePretrainedNetwork/models/superpoint.py:
//   :167-170 softmax-65 + drop dustbin + 8x8 depth-to-space      -> sp_softmax_d2s
//   :47-62   simple_nms (3 rounds, exact float equality)         -> sp_nms
//   :174-191 threshold, row-major nonzero, border, top-k         -> sp_select
//   :196     channel L2 norm of the dense descriptor map         -> sp_l2norm_rows
//   :80-92   bilinear sampling (align_corners=True) + L2 norm    -> sp_sample_desc
#include <math_constants.h>

#include "common.cuh"
#include "sp_kernels.h"

namespace {

// ---- softmax over 65 channels + depth-to-space -----------------------------------------------
// one warp per 8x8 cell; lane l owns channels l, l+32 (and lane 0 the dustbin, channel 64).
__global__ void __launch_bounds__(256) softmax_d2s_kernel(const float* __restrict__ logits, float* __restrict__ dense,
                                                          int B, int h, int w, int ld) {
  const long long cell = (long long)blockIdx.x * 8 + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  const long long ncell = (long long)B * h * w;
  if (cell >= ncell) return;
  const float* p = logits + cell * ld;
  float a = p[lane], b = p[lane + 32];
  float d = (lane == 0) ? p[64] : -CUDART_INF_F;
  float m = warp_max(fmaxf(fmaxf(a, b), d));
  float ea = expf(a - m), eb = expf(b - m);
  float ed = (lane == 0) ? expf(d - m) : 0.f;
  float s = warp_sum(ea + eb + ed);
  const int bi = (int)(cell / ((long long)h * w));
  const int rem = (int)(cell % ((long long)h * w));
  const int cy = rem / w, cx = rem % w;
  const int W = w * 8;
  float* o = dense + ((long long)bi * h * 8 + cy * 8) * W + cx * 8;
  // channel c -> pixel (8cy + c/8, 8cx + c%8)
  o[(lane / 8) * W + (lane % 8)] = __fdiv_rn(ea, s);
  o[(lane / 8 + 4) * W + (lane % 8)] = __fdiv_rn(eb, s);
}

// ---- simple_nms ---------------------------------------------------------------------------------
// CTA = 32x32 output pixels + halo of 5r (three max-pools and two dilations of radius r chained).
constexpr int NMS_T_SMALL = 32, NMS_T_LARGE = 64;   // output tile edge: 64 while (64 + 10 r)^2 cells fit in shared memory (r <= 5)

// RAD > 0: compile-time radius (fully unrolled 2r+1-tap windows); RAD = 0: run-time radius `r_dyn`
template <int RAD>
__global__ void __launch_bounds__(512) nms_kernel(const float* __restrict__ dense, float* __restrict__ out, int H, int W,
                                                  int r_dyn, int NMS_T) {
  const int r = RAD > 0 ? RAD : r_dyn;
  extern __shared__ __align__(16) unsigned char nms_smem[];
  const int halo = 5 * r, R = NMS_T + 2 * halo, RR = R * R;
  float* S = reinterpret_cast<float*>(nms_smem);  // scores (-inf outside the image)
  float* A = S + RR;                              // row-pooled temp
  float* SS = A + RR;                             // suppressed scores
  unsigned char* mask = reinterpret_cast<unsigned char*>(SS + RR);
  unsigned char* supp = mask + RR;
  unsigned char* tmpb = supp + RR;

  const int tiles_x = (W + NMS_T - 1) / NMS_T;
  const int x0 = (blockIdx.x % tiles_x) * NMS_T - halo, y0 = (blockIdx.x / tiles_x) * NMS_T - halo;
  const float* img = dense + (long long)blockIdx.z * H * W;
  const int lane = threadIdx.x % 32, wrp = threadIdx.x / 32, nwarps = blockDim.x / 32;
  // rows are distributed over the warps, columns over the lanes: no integer division in the inner loops
#define NMS_FOR_EACH(yy, xx) for (int yy = wrp; yy < R; yy += nwarps) for (int xx = lane; xx < R; xx += 32)

  NMS_FOR_EACH(yy, xx) {
    int gy = y0 + yy, gx = x0 + xx;
    S[yy * R + xx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? img[(long long)gy * W + gx] : -CUDART_INF_F;
  }
  __syncthreads();
  // max_mask = scores == max_pool(scores)   (separable: rows then columns)
  NMS_FOR_EACH(yy, xx) {
    // windows that would leave the staged region are skipped: those cells lie in the halo margin that the
    // 5r budget already writes off, and fixed-size windows unroll completely for a compile-time radius
    if (xx < r || xx >= R - r) continue;
    float m = -CUDART_INF_F;
#pragma unroll
    for (int d = -r; d <= r; d++) m = fmaxf(m, S[yy * R + xx + d]);
    A[yy * R + xx] = m;
  }
  __syncthreads();
  NMS_FOR_EACH(yy, xx) {
    if (yy < r || yy >= R - r) { mask[yy * R + xx] = 0; continue; }
    float m = -CUDART_INF_F;
#pragma unroll
    for (int d = -r; d <= r; d++) m = fmaxf(m, A[(yy + d) * R + xx]);
    float s = S[yy * R + xx];
    mask[yy * R + xx] = (s == m) && (s != -CUDART_INF_F);
  }
  __syncthreads();
  for (int it = 0; it < 2; it++) {
    // supp = max_pool(max_mask) > 0
    NMS_FOR_EACH(yy, xx) {
      if (xx < r || xx >= R - r) { tmpb[yy * R + xx] = 0; continue; }
      unsigned char m = 0;
#pragma unroll
      for (int d = -r; d <= r; d++) m |= mask[yy * R + xx + d];
      tmpb[yy * R + xx] = m;
    }
    __syncthreads();
    NMS_FOR_EACH(yy, xx) {
      unsigned char m = 0;
      if (yy >= r && yy < R - r) {
#pragma unroll
        for (int d = -r; d <= r; d++) m |= tmpb[(yy + d) * R + xx];
      }
      supp[yy * R + xx] = m;
      // supp_scores = where(supp, 0, scores); stays -inf outside the image (max_pool2d padding)
      float s = S[yy * R + xx];
      SS[yy * R + xx] = (s == -CUDART_INF_F) ? -CUDART_INF_F : (m ? 0.f : s);
    }
    __syncthreads();
    NMS_FOR_EACH(yy, xx) {
      if (xx < r || xx >= R - r) continue;
      float m = -CUDART_INF_F;
#pragma unroll
      for (int d = -r; d <= r; d++) m = fmaxf(m, SS[yy * R + xx + d]);
      A[yy * R + xx] = m;
    }
    __syncthreads();
    NMS_FOR_EACH(yy, xx) {
      if (yy < r || yy >= R - r) continue;
      float m = -CUDART_INF_F;
#pragma unroll
      for (int d = -r; d <= r; d++) m = fmaxf(m, A[(yy + d) * R + xx]);
      float s = SS[yy * R + xx];
      bool new_max = (s == m) && (s != -CUDART_INF_F);
      if (new_max && !supp[yy * R + xx]) mask[yy * R + xx] = 1;
    }
    __syncthreads();
  }
#undef NMS_FOR_EACH
  float* o = out + (long long)blockIdx.z * H * W;
  for (int yy = wrp; yy < NMS_T; yy += nwarps)
    for (int xx = lane; xx < NMS_T; xx += 32) {
      int gy = y0 + halo + yy, gx = x0 + halo + xx;
      if (gy < H && gx < W) {
        int si = (yy + halo) * R + xx + halo;
        o[(long long)gy * W + gx] = mask[si] ? S[si] : 0.f;
      }
    }
}

// ---- simple_nms, restructured (radius 1..5) --------------------------------------------------------------------------------
// Same arithmetic as nms_kernel (exact float equality, -inf outside the image, results valid 5r inside the staged region), but
//   * max-pools are register-blocked: the row pass produces 4 outputs from three (five at r = 5) 16-byte shared-memory loads, the
//     column pass slides a 2r+1 window down 8 rows per thread (8 + 2r loads for 8 outputs) -- the old kernel read 2r+1 values
//     per output per pass, ten passes over the tile: 2.3 ms per 128 images = 150 GB/s of the 6.5 TB/s HBM peak, bound by
//     shared-memory wavefronts;
//   * masks are bit rows (one ballot per 32 pixels): the two dilations are shifts and ORs on a few hundred words;
//   * supp_scores = where(supp, 0, scores) is never materialised: both passes rebuild it from the scores and the supp bits.
// Layout: S / A float [R][RS] with 8 pad columns on either side (float4-aligned windows), bit rows [R][NW + 2] with one zero pad
// word on either side.
constexpr int NF_T = 64, NF_THREADS = 512;
template <int RAD> struct NmsFast {
  static constexpr int R = NF_T + 10 * RAD;                 // staged rows / columns
  static constexpr int RP = (R + 31) / 32 * 32;             // columns rounded up to whole bit words
  static constexpr int NW = RP / 32, BW = NW + 2;           // words per bit row (+ pads)
  static constexpr int RS = RP + 16;                        // float row stride (8 pad columns each side)
  static constexpr size_t smem = (size_t)2 * R * RS * sizeof(float) + (size_t)4 * R * BW * sizeof(unsigned);
};

template <int RAD>
__global__ void __launch_bounds__(NF_THREADS) nms_fast_kernel(const float* __restrict__ dense, float* __restrict__ out, int H, int W) {
  using C = NmsFast<RAD>;
  constexpr int R = C::R, RP = C::RP, NW = C::NW, BW = C::BW, RS = C::RS, r = RAD, halo = 5 * RAD;
  extern __shared__ __align__(16) unsigned char nms_smem[];
  float* S = reinterpret_cast<float*>(nms_smem);            // scores, -inf outside the image / in the pads
  float* A = S + R * RS;                                    // row-pooled temp
  unsigned* M = reinterpret_cast<unsigned*>(A + R * RS);    // max_mask bits
  unsigned* SUP = M + R * BW;                               // supp bits
  unsigned* TMP = SUP + R * BW;                             // row-dilated mask
  unsigned* NEWM = TMP + R * BW;                            // unused words stay zero (pads)

  const int tiles_x = (W + NF_T - 1) / NF_T;
  const int x0 = (blockIdx.x % tiles_x) * NF_T - halo, y0 = (blockIdx.x / tiles_x) * NF_T - halo;
  const float* img = dense + (long long)blockIdx.z * H * W;
  const int tid = threadIdx.x, lane = tid % 32, wrp = tid / 32;
  constexpr int NWARPS = NF_THREADS / 32;
  const float NINF = -CUDART_INF_F;

  for (int i = tid; i < R * RS; i += NF_THREADS) {
    const int yy = i / RS, xx = i % RS - 8;
    const int gy = y0 + yy, gx = x0 + xx;
    S[i] = (xx >= 0 && xx < R && gy >= 0 && gy < H && gx >= 0 && gx < W) ? img[(long long)gy * W + gx] : NINF;
  }
  for (int i = tid; i < 4 * R * BW; i += NF_THREADS) M[i] = 0u;
  __syncthreads();

  // value of pixel (yy, xx) in the current score map: scores (SUPP = false) or where(supp, 0, scores)
  // row pass: A = max over [x - r, x + r]; four outputs per item from aligned 16-byte loads
  auto row_pass = [&](bool use_supp) {
    constexpr int NV = RAD <= 4 ? 3 : 5, LEFT = RAD <= 4 ? 4 : 8;        // float4 loads, first column relative to x
    for (int it = tid; it < R * (RP / 4); it += NF_THREADS) {
      const int yy = it / (RP / 4), x = (it % (RP / 4)) * 4;
      float v[4 * NV];
      const float4* src = reinterpret_cast<const float4*>(S + yy * RS + 8 + x - LEFT);
#pragma unroll
      for (int q = 0; q < NV; q++) { const float4 t = src[q]; v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w; }
      if (use_supp) {
        const int bp = x - LEFT + 32;                                      // bit position in the padded bit row
        const unsigned* wr = SUP + yy * BW + (bp >> 5);
        const unsigned long long bits = (((unsigned long long)wr[1] << 32) | wr[0]) >> (bp & 31);
#pragma unroll
        for (int k = 0; k < 4 * NV; k++)
          if ((bits >> k) & 1ull) v[k] = (v[k] == NINF) ? NINF : 0.f;
      }
      float o[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        float m = v[LEFT + i - r];
#pragma unroll
        for (int d = -r + 1; d <= r; d++) m = fmaxf(m, v[LEFT + i + d]);
        o[i] = m;
      }
      *reinterpret_cast<float4*>(A + yy * RS + 8 + x) = make_float4(o[0], o[1], o[2], o[3]);
    }
  };
  // column pass: m = max over rows [y - r, y + r] of A; bit = (value == m) && value != -inf.  One warp = 32 columns x 8 rows.
  auto col_pass = [&](bool use_supp) {
    constexpr int RUN = 8, NRUN = (R + RUN - 1) / RUN;
    for (int task = wrp; task < NW * NRUN; task += NWARPS) {
      const int cb = task % NW, yb = (task / NW) * RUN;
      const int xx = cb * 32 + lane;
      float win[RUN + 2 * RAD];
#pragma unroll
      for (int k = 0; k < RUN + 2 * RAD; k++) {
        const int yy = yb - r + k;
        win[k] = (yy >= 0 && yy < R) ? A[yy * RS + 8 + xx] : NINF;
      }
#pragma unroll
      for (int k = 0; k < RUN; k++) {
        const int yy = yb + k;
        if (yy >= R) break;                                   // warp-uniform
        float m = win[k];
#pragma unroll
        for (int d = 1; d <= 2 * RAD; d++) m = fmaxf(m, win[k + d]);
        float c = S[yy * RS + 8 + xx];
        const unsigned sup = use_supp ? SUP[yy * BW + 1 + cb] : 0u;
        if ((sup >> lane) & 1u) c = (c == NINF) ? NINF : 0.f;
        const unsigned bal = __ballot_sync(0xffffffffu, (c == m) && (c != NINF) && xx < R);
        if (lane == 0) {
          if (use_supp) M[yy * BW + 1 + cb] |= bal & ~sup;    // max_mask | (new_max_mask & ~supp_mask)
          else M[yy * BW + 1 + cb] = bal;
        }
      }
    }
  };
  // supp = max_pool(max_mask) > 0: dilation by r along the row (shifts across word boundaries), then along the column
  auto dilate = [&]() {
    for (int i = tid; i < R * NW; i += NF_THREADS) {
      const int yy = i / NW, w = 1 + i % NW;
      const unsigned lo = M[yy * BW + w - 1], mid = M[yy * BW + w], hi = M[yy * BW + w + 1];
      unsigned d = mid;
#pragma unroll
      for (int k = 1; k <= RAD; k++) d |= (mid << k) | (lo >> (32 - k)) | (mid >> k) | (hi << (32 - k));
      TMP[yy * BW + w] = d;
    }
    __syncthreads();
    for (int i = tid; i < R * NW; i += NF_THREADS) {
      const int yy = i / NW, w = 1 + i % NW;
      unsigned d = 0u;
#pragma unroll
      for (int k = -RAD; k <= RAD; k++)
        if (yy + k >= 0 && yy + k < R) d |= TMP[(yy + k) * BW + w];
      SUP[yy * BW + w] = d;
    }
    __syncthreads();
  };

  row_pass(false);
  __syncthreads();
  col_pass(false);
  __syncthreads();
  for (int it = 0; it < 2; it++) {
    dilate();
    row_pass(true);
    __syncthreads();
    col_pass(true);
    __syncthreads();
  }
  float* o = out + (long long)blockIdx.z * H * W;
  for (int i = tid; i < NF_T * NF_T; i += NF_THREADS) {
    const int yy = i / NF_T + halo, xx = i % NF_T + halo;
    const int gy = y0 + yy, gx = x0 + xx;
    if (gy < H && gx < W) {
      const bool keep = (M[yy * BW + 1 + (xx >> 5)] >> (xx & 31)) & 1u;
      o[(long long)gy * W + gx] = keep ? S[yy * RS + 8 + xx] : 0.f;
    }
  }
}

// ---- threshold / border / ordered compaction / top-k -------------------------------------------
// key = score bits (positive float => order-preserving) << 32 | (~pixel index): sorting keys in
// descending order yields descending score with ascending row-major index on ties.
__device__ __forceinline__ unsigned long long make_key(float s, unsigned idx) {
  return ((unsigned long long)__float_as_uint(s) << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
}

__device__ void bitonic_sort_desc(unsigned long long* a, int n_pad, int tid, int nthreads) {
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n_pad; i += nthreads) {
        int l = i ^ j;
        if (l > i) {
          unsigned long long x = a[i], y = a[l];
          bool desc = ((i & k) == 0);
          if (desc ? (x < y) : (x > y)) { a[i] = y; a[l] = x; }
        }
      }
      __syncthreads();
    }
  }
}

constexpr int SEL_THREADS = 1024;
constexpr int SEL_SMEM_KEYS = 16384;  // 128 KB of shared memory for the in-CTA sort

__global__ void __launch_bounds__(SEL_THREADS) select_kernel(const float* __restrict__ nms, unsigned long long* __restrict__ keys_all,
                                                             int key_cap, float* __restrict__ kpts, float* __restrict__ scores,
                                                             int* __restrict__ counts, int H, int W, float thr_uniform, int border,
                                                             int max_kpts, int cap, const float* __restrict__ thr_img) {
  extern __shared__ __align__(16) unsigned long long s_keys[];
  __shared__ int s_scan[SEL_THREADS / 32];
  __shared__ int s_total;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float thr = thr_img ? thr_img[b] : thr_uniform;   // per-image threshold (ALIKED's mean fallback)
  const float* img = nms + (long long)b * H * W;
  unsigned long long* keys = keys_all + (long long)b * key_cap;
  const int npix = H * W;
  const int lane = tid % 32, wid = tid / 32;
  // each warp scans a contiguous range of pixels 32 at a time (coalesced); row-major order is kept
  // through ballot ranks inside a step and a running offset across steps.
  const int per_warp = ((npix + (SEL_THREADS / 32) - 1) / (SEL_THREADS / 32) + 31) / 32 * 32;
  const int p0 = wid * per_warp, p1 = min(p0 + per_warp, npix);

  auto pass = [&](int p, float v) -> bool {
    int y = p / W, x = p % W;
    return (v > thr) && (y >= border) && (y < H - border) && (x >= border) && (x < W - border);
  };
  int cnt = 0;
  for (int p = p0 + lane; p < p0 + per_warp; p += 32) {
    bool ok = (p < p1) && pass(p, img[p]);
    cnt += __popc(__ballot_sync(0xffffffffu, ok));
  }
  if (lane == 0) s_scan[wid] = cnt;
  __syncthreads();
  if (wid == 0) {
    int v = s_scan[lane];
    int winc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    s_scan[lane] = winc - v;
    if (lane == 31) s_total = winc;
  }
  __syncthreads();
  int off = s_scan[wid];
  const int n = s_total;
  for (int p = p0 + lane; p < p0 + per_warp; p += 32) {
    float v = (p < p1) ? img[p] : 0.f;
    bool ok = (p < p1) && pass(p, v);
    unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (ok) keys[off + __popc(bal & ((1u << lane) - 1u))] = make_key(v, (unsigned)p);
    off += __popc(bal);
  }
  __syncthreads();  // global writes by this CTA are visible to this CTA after the barrier

  int n_out;
  unsigned long long* src = keys;
  if (max_kpts >= 0 && n > max_kpts) {
    int n_pad = 1;
    while (n_pad < n) n_pad <<= 1;
    if (n_pad <= SEL_SMEM_KEYS) {
      for (int i = tid; i < n_pad; i += SEL_THREADS) s_keys[i] = (i < n) ? keys[i] : 0ull;
      __syncthreads();
      bitonic_sort_desc(s_keys, n_pad, tid, SEL_THREADS);
      src = s_keys;
    } else {
      for (int i = n + tid; i < n_pad && i < key_cap; i += SEL_THREADS) keys[i] = 0ull;
      __syncthreads();
      // key_cap is sized to the next power of two of H*W, so n_pad <= key_cap always holds
      bitonic_sort_desc(keys, n_pad, tid, SEL_THREADS);
    }
    n_out = max_kpts;
  } else {
    n_out = n;
  }
  if (tid == 0) {
    counts[b] = min(n_out, cap);      // keypoints written
    counts[gridDim.x + b] = n_out;    // keypoints the reference would return (overflow check)
  }
  n_out = min(n_out, cap);
  for (int i = tid; i < n_out; i += SEL_THREADS) {
    unsigned long long k = src[i];
    unsigned idx = 0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull);
    float s = __uint_as_float((unsigned)(k >> 32));
    kpts[((long long)b * cap + i) * 2 + 0] = (float)(idx % W);
    kpts[((long long)b * cap + i) * 2 + 1] = (float)(idx / W);
    scores[(long long)b * cap + i] = s;
  }
}

// ---- L2 normalisation of rows -------------------------------------------------------------------
__global__ void __launch_bounds__(256) l2norm_rows_kernel(float* __restrict__ x, long long rows, int C) {
  long long row = (long long)blockIdx.x * 8 + threadIdx.x / 32;
  if (row >= rows) return;
  int lane = threadIdx.x % 32;
  float* p = x + row * C;
  float ss = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    float4 v = *reinterpret_cast<const float4*>(p + c);
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  ss = warp_sum(ss);
  float d = fmaxf(sqrtf(ss), 1e-12f);  // F.normalize eps
  for (int c = lane * 4; c < C; c += 128) {
    float4 v = *reinterpret_cast<const float4*>(p + c);
    v.x = __fdiv_rn(v.x, d); v.y = __fdiv_rn(v.y, d); v.z = __fdiv_rn(v.z, d); v.w = __fdiv_rn(v.w, d);
    *reinterpret_cast<float4*>(p + c) = v;
  }
}

// ---- descriptor sampling ---------------------------------------------------------------------------
// warp per keypoint; C = 256 -> 8 channels per lane.
__global__ void __launch_bounds__(256) sample_desc_kernel(const float* __restrict__ dd, const float* __restrict__ kpts,
                                                          const int* __restrict__ counts, float* __restrict__ desc, int h,
                                                          int w, int cap, int C, int fix_sampling) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 8 + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (i >= counts[b]) return;
  const float kx = kpts[((long long)b * cap + i) * 2], ky = kpts[((long long)b * cap + i) * 2 + 1];
  const float s = 8.f;
  // superpoint.py:83-86: (k - s/2 + 0.5) / (w*s - s/2 - 0.5) * 2 - 1, then grid_sample unnormalise
  float gx = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(__fsub_rn(kx, s / 2), 0.5f), (w * s - s / 2 - 0.5f)), 2.f), 1.f);
  float gy = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(__fsub_rn(ky, s / 2), 0.5f), (h * s - s / 2 - 0.5f)), 2.f), 1.f);
  float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), (float)(w - 1));
  float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), (float)(h - 1));
  if (fix_sampling) {
    // hloc/extractors/superpoint.py:19-27: (k + 0.5) / ([w, h] * s) * 2 - 1, grid_sample(align_corners=False):
    // unnormalise ((g + 1) * size - 1) / 2
    gx = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(kx, 0.5f), (w * s)), 2.f), 1.f);
    gy = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(ky, 0.5f), (h * s)), 2.f), 1.f);
    ix = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)w), 1.f), 2.f);
    iy = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)h), 1.f), 2.f);
  }
  float fx = floorf(ix), fy = floorf(iy);
  int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  float wx1 = ix - fx, wx0 = (fx + 1.f) - ix, wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
  float w_nw = wx0 * wy0, w_ne = wx1 * wy0, w_sw = wx0 * wy1, w_se = wx1 * wy1;
  const float* base = dd + (long long)b * h * w * C;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; k++) acc[k] = 0.f;
  auto add = [&](int yy, int xx, float wt) {
    if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
      const float* p = base + ((long long)yy * w + xx) * C + lane * 8;
      float4 a = *reinterpret_cast<const float4*>(p), c = *reinterpret_cast<const float4*>(p + 4);
      acc[0] += a.x * wt; acc[1] += a.y * wt; acc[2] += a.z * wt; acc[3] += a.w * wt;
      acc[4] += c.x * wt; acc[5] += c.y * wt; acc[6] += c.z * wt; acc[7] += c.w * wt;
    }
  };
  add(y0, x0, w_nw); add(y0, x1, w_ne); add(y1, x0, w_sw); add(y1, x1, w_se);
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < 8; k++) ss += acc[k] * acc[k];
  ss = warp_sum(ss);
  float d = fmaxf(sqrtf(ss), 1e-12f);
  float* o = desc + ((long long)b * cap + i) * C + lane * 8;
  *reinterpret_cast<float4*>(o) = make_float4(__fdiv_rn(acc[0], d), __fdiv_rn(acc[1], d), __fdiv_rn(acc[2], d), __fdiv_rn(acc[3], d));
  *reinterpret_cast<float4*>(o + 4) = make_float4(__fdiv_rn(acc[4], d), __fdiv_rn(acc[5], d), __fdiv_rn(acc[6], d), __fdiv_rn(acc[7], d));
}

}  // namespace

int sp_softmax_d2s(const float* logits, float* dense, int B, int h, int w, cudaStream_t st, int ld) {
  long long ncell = (long long)B * h * w;
  softmax_d2s_kernel<<<(unsigned)((ncell + 7) / 8), 256, 0, st>>>(logits, dense, B, h, w, ld);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}

int sp_nms(const float* dense, float* nms, int B, int H, int W, int radius, cudaStream_t st) {
  IMW_REQUIRE(radius >= 0 && radius <= 8, "sp_nms: nms_radius must be in [0,8] (got %d)", radius);
  if (radius >= 1 && radius <= 5) {   // register-blocked / bit-mask kernel
    dim3 fgrid(ceil_div(W, NF_T) * ceil_div(H, NF_T), 1, B);
    auto flaunch = [&](auto kern, size_t fsmem) -> cudaError_t {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
      if (e != cudaSuccess) return e;
      kern<<<fgrid, NF_THREADS, fsmem, st>>>(dense, nms, H, W);
      return cudaSuccess;
    };
    switch (radius) {
      case 1: IMW_CHECK_CUDA(flaunch(nms_fast_kernel<1>, NmsFast<1>::smem)); break;
      case 2: IMW_CHECK_CUDA(flaunch(nms_fast_kernel<2>, NmsFast<2>::smem)); break;
      case 3: IMW_CHECK_CUDA(flaunch(nms_fast_kernel<3>, NmsFast<3>::smem)); break;
      case 4: IMW_CHECK_CUDA(flaunch(nms_fast_kernel<4>, NmsFast<4>::smem)); break;
      default: IMW_CHECK_CUDA(flaunch(nms_fast_kernel<5>, NmsFast<5>::smem)); break;
    }
    IMW_CHECK_LAUNCH();
    return IMW_OK;
  }
  const int NMS_T = radius <= 5 ? NMS_T_LARGE : NMS_T_SMALL;   // halo redundancy (T + 10 r)^2 / T^2: 2.6x instead of 5x at r = 4
  int R = NMS_T + 10 * radius;
  size_t smem = (size_t)R * R * (3 * sizeof(float) + 3);
  dim3 grid(ceil_div(W, NMS_T) * ceil_div(H, NMS_T), 1, B);
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, NMS_T == NMS_T_LARGE ? 512 : 256, smem, st>>>(dense, nms, H, W, radius, NMS_T);
    return cudaSuccess;
  };
  if (radius == 2) IMW_CHECK_CUDA(launch(nms_kernel<2>));
  else if (radius == 3) IMW_CHECK_CUDA(launch(nms_kernel<3>));
  else if (radius == 4) IMW_CHECK_CUDA(launch(nms_kernel<4>));
  else IMW_CHECK_CUDA(launch(nms_kernel<0>));
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}

size_t sp_select_key_cap(int H, int W) {
  size_t n = 1;
  while (n < (size_t)H * W) n <<= 1;
  return n;
}

int sp_select(const float* nms, unsigned long long* keys, int key_cap, float* kpts, float* scores, int* counts, int B,
              int H, int W, float threshold, int border, int max_kpts, int cap, cudaStream_t st, const float* thr_img) {
  size_t smem = (size_t)SEL_SMEM_KEYS * sizeof(unsigned long long);
  IMW_CHECK_CUDA(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  select_kernel<<<B, SEL_THREADS, smem, st>>>(nms, keys, key_cap, kpts, scores, counts, H, W, threshold, border, max_kpts, cap, thr_img);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}

int sp_l2norm_rows(float* x, long long rows, int C, cudaStream_t st) {
  IMW_REQUIRE(C % 4 == 0, "sp_l2norm_rows: C %% 4");
  l2norm_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, rows, C);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}

int sp_sample_desc(const float* dense_desc, const float* kpts, const int* counts, float* desc, int B, int h, int w, int cap,
                   int C, int fix_sampling, cudaStream_t st) {
  IMW_REQUIRE(C == 256, "sp_sample_desc: descriptor_dim must be 256");
  dim3 grid(ceil_div(cap, 8), B);
  sample_desc_kernel<<<grid, 256, 0, st>>>(dense_desc, kpts, counts, desc, h, w, cap, C, fix_sampling);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}
