// GPU pre-/post-processing of the hloc drivers (rows a1 / a5 / a10):
//   imw_preprocess      replaces imcui/hloc/extract_features.py:120-162 (and match_dense.py:588-640, same calls):
//                       cv2.cvtColor(RGB2GRAY) on uint8, astype(float32), cv2.resize INTER_AREA (resize_max, then
//                       force_resize; INTER_LINEAR when up-sampling, :30-31), / 255, torchvision antialias resize to a
//                       multiple of dfactor -- decoded uint8 frames in HBM -> the fp32 [B,C,H,W] tensor the extractor takes.
//   imw_gather_matches  replaces imcui/hloc/match_features.py:236-257: valid = matches0 > -1, gather matched keypoints,
//                       rescale (k + 0.5) * s - 0.5 -- so that ONE compact D2H per batch replaces per-pair .cpu() + NumPy.
//
// The arithmetic of the reference lives in OpenCV / ATen; the kernels follow those routines operation by operation
// (same tables, same summation order, no FMA contraction where the library has none, FMA where it has one), restated
// and pinned in oracle/preprocess.py.  All of it is HBM-bound byte work: one thread per output pixel, coalesced along x.
#include <math.h>

#include "../../include/imw_b200.h"
#include "common.cuh"

namespace {

constexpr int PP_MAX_TAPS = 64;   // taps per axis of one output pixel (scale factor <= 62)

// ---- uint8 frames -> fp32 planes -----------------------------------------------------------------------------------
// cv2.cvtColor(RGB2GRAY) on uint8: 15-bit fixed point, round half up (color_rgb.simd.hpp: R2Y 9798, G2Y 19235, B2Y 3735).
__device__ __forceinline__ float load_px(const uint8_t* __restrict__ img, int W, int C, int gray, int c, int y, int x) {
  const uint8_t* p = img + ((long long)y * W + x) * C;
  if (C == 3 && gray) return (float)((p[0] * 9798 + p[1] * 19235 + p[2] * 3735 + (1 << 14)) >> 15);
  return (float)p[c];
}

__global__ void __launch_bounds__(256) pp_convert_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int H, int W, int C,
                                                         int gray, int Co, int div255) {
  const int b = blockIdx.z, y = blockIdx.y, x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  const uint8_t* img = in + (long long)b * H * W * C;
  for (int c = 0; c < Co; c++) {
    float v = load_px(img, W, C, gray, c, y, x);
    if (div255) v = __fdiv_rn(v, 255.0f);
    out[(((long long)b * Co + c) * H + y) * W + x] = v;
  }
}

// ---- cv2 INTER_AREA (float32), general scale: computeResizeAreaTab -------------------------------------------------------
// Per destination index: first source index, tap count, taps' alphas (the taps are consecutive source cells).
struct AreaTab { int* start; int* count; float* alpha; };   // alpha [n][PP_MAX_TAPS]

__global__ void pp_area_tab_kernel(AreaTab t, int ssize, int dsize) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x;
  if (dx >= dsize) return;
  const double scale = 1.0 / ((double)dsize / (double)ssize);
  const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
  const double cell = fmin(scale, (double)ssize - fsx1);
  int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
  sx2 = min(sx2, ssize - 1);
  sx1 = min(sx1, sx2);
  float* a = t.alpha + (long long)dx * PP_MAX_TAPS;
  int n = 0, first = sx1;
  if (sx1 - fsx1 > 1e-3) { first = sx1 - 1; a[n++] = (float)((sx1 - fsx1) / cell); }
  for (int sx = sx1; sx < sx2 && n < PP_MAX_TAPS; sx++) a[n++] = (float)(1.0 / cell);
  if (fsx2 - sx2 > 1e-3 && n < PP_MAX_TAPS) a[n++] = (float)(fmin(fmin(fsx2 - sx2, 1.0), cell) / cell);
  t.start[dx] = first;
  t.count[dx] = n;
}

// Source of a resize stage: fp32 planes [B][Co][Hs][Ws] or the uint8 frames themselves (conversion fused into the read).
struct Src {
  const float* f32; const uint8_t* u8; int Hs, Ws, C, gray, Co;
  __device__ __forceinline__ float at(int b, int c, int y, int x) const {
    if (u8) return load_px(u8 + (long long)b * Hs * Ws * C, Ws, C, gray, c, y, x);
    return f32[(((long long)b * Co + c) * Hs + y) * Ws + x];
  }
};

// ResizeArea_Invoker: per source row: buf[dx] = sum_k S[si_k] * alpha_k (in tap order, separate multiply and add);
// first row of a destination row: sum = beta * buf, further rows: sum += beta * buf.
__global__ void __launch_bounds__(256) pp_area_kernel(Src s, AreaTab xt, AreaTab yt, float* __restrict__ out, int Hd, int Wd, int div255) {
  const int b = blockIdx.z / s.Co, c = blockIdx.z % s.Co, dy = blockIdx.y, dx = blockIdx.x * blockDim.x + threadIdx.x;
  if (dx >= Wd) return;
  const int x0 = xt.start[dx], nx = xt.count[dx], y0 = yt.start[dy], ny = yt.count[dy];
  const float* ax = xt.alpha + (long long)dx * PP_MAX_TAPS;
  const float* ay = yt.alpha + (long long)dy * PP_MAX_TAPS;
  float sum = 0.f;
  for (int j = 0; j < ny; j++) {
    float buf = 0.f;
    for (int k = 0; k < nx; k++) buf = __fadd_rn(buf, __fmul_rn(s.at(b, c, y0 + j, x0 + k), ax[k]));
    const float t = __fmul_rn(ay[j], buf);
    sum = (j == 0) ? t : __fadd_rn(sum, t);
  }
  if (div255) sum = __fdiv_rn(sum, 255.0f);
  out[(((long long)b * s.Co + c) * Hd + dy) * Wd + dx] = sum;
}

// ResizeAreaFast_Invoker (integer scales): block sum in row-major order, unrolled by four as the library's scalar loop
// (sum += ((S0 + S1) + S2) + S3), times float(1 / area).  The 2x2 single-channel case goes through the library's SIMD path
// ((a + b) + (c + d)) * 0.25 for whole 4-pixel vectors (universal intrinsics at the SSE baseline); the row tail and the
// 3-channel call run the scalar loop.
__global__ void __launch_bounds__(256) pp_area_fast_kernel(Src s, float* __restrict__ out, int Hd, int Wd, int sx, int sy, int div255) {
  const int b = blockIdx.z / s.Co, c = blockIdx.z % s.Co, dy = blockIdx.y, dx = blockIdx.x * blockDim.x + threadIdx.x;
  if (dx >= Wd) return;
  float r;
  if (sx == 2 && sy == 2 && s.Co == 1 && dx < (Wd & ~3)) {
    const float a0 = s.at(b, c, 2 * dy, 2 * dx), a1 = s.at(b, c, 2 * dy, 2 * dx + 1);
    const float b0 = s.at(b, c, 2 * dy + 1, 2 * dx), b1 = s.at(b, c, 2 * dy + 1, 2 * dx + 1);
    r = __fmul_rn(__fadd_rn(__fadd_rn(a0, a1), __fadd_rn(b0, b1)), 0.25f);
  } else {
    const int area = sx * sy;
    float sum = 0.f;
    int k = 0;
    auto px = [&](int kk) { return s.at(b, c, dy * sy + kk / sx, dx * sx + kk % sx); };
    for (; k <= area - 4; k += 4) sum = __fadd_rn(sum, __fadd_rn(__fadd_rn(__fadd_rn(px(k), px(k + 1)), px(k + 2)), px(k + 3)));
    for (; k < area; k++) sum = __fadd_rn(sum, px(k));
    r = __fmul_rn(sum, (float)(1.0 / area));
  }
  if (div255) r = __fdiv_rn(r, 255.0f);
  out[(((long long)b * s.Co + c) * Hd + dy) * Wd + dx] = r;
}

// ---- cv2 INTER_LINEAR (float32), OpenCV's own arithmetic (HResizeLinear then VResizeLinear, no FMA) ----------------------
// NB: pip builds of OpenCV route this call through Intel IPP, whose arithmetic is not published: up-sampling parity is
// ~4e-6 relative, not bit-exact (oracle/preprocess.py header).  The reference only takes this branch when up-sampling.
__device__ __forceinline__ void linear_coef(int d, int ssize, int dsize, int& s0, float& f) {
  const double scale = 1.0 / ((double)dsize / (double)ssize);
  float fx = (float)((d + 0.5) * scale - 0.5);
  int sx = (int)floorf(fx);
  fx = __fsub_rn(fx, (float)sx);
  if (sx < 0) { sx = 0; fx = 0.f; }
  if (sx >= ssize - 1) { sx = ssize - 1; fx = 0.f; }
  s0 = sx; f = fx;
}
__global__ void __launch_bounds__(256) pp_linear_kernel(Src s, float* __restrict__ out, int Hd, int Wd, int div255) {
  const int b = blockIdx.z / s.Co, c = blockIdx.z % s.Co, dy = blockIdx.y, dx = blockIdx.x * blockDim.x + threadIdx.x;
  if (dx >= Wd) return;
  int sx, sy; float fx, fy;
  linear_coef(dx, s.Ws, Wd, sx, fx);
  linear_coef(dy, s.Hs, Hd, sy, fy);
  const int sx1 = min(sx + 1, s.Ws - 1), sy1 = min(sy + 1, s.Hs - 1);
  const float a0 = __fsub_rn(1.f, fx), b0 = __fsub_rn(1.f, fy);
  const float r0 = __fadd_rn(__fmul_rn(s.at(b, c, sy, sx), a0), __fmul_rn(s.at(b, c, sy, sx1), fx));
  const float r1 = __fadd_rn(__fmul_rn(s.at(b, c, sy1, sx), a0), __fmul_rn(s.at(b, c, sy1, sx1), fx));
  float r = __fadd_rn(__fmul_rn(r0, b0), __fmul_rn(r1, fy));
  if (div255) r = __fdiv_rn(r, 255.0f);
  out[(((long long)b * s.Co + c) * Hd + dy) * Wd + dx] = r;
}

// ---- torchvision F.resize(antialias=True) = ATen _upsample_bilinear2d_aa (CPU, float) -----------------------------------
// _compute_indices_weights_aa in scalar_t = float; one 1-D pass per axis (width first), t = src0 * w0 then t = fma(src_j, w_j, t).
__global__ void pp_aa_tab_kernel(AreaTab t, int in_size, int out_size) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out_size) return;
  const float scale = __fdiv_rn((float)in_size, (float)out_size);
  const float support = scale >= 1.f ? scale : 1.f;
  const float invscale = scale >= 1.f ? __fdiv_rn(1.f, scale) : 1.f;
  const float center = __fmul_rn(scale, (float)i + 0.5f);
  const int xmin = max((int)__fadd_rn(__fsub_rn(center, support), 0.5f), 0);
  const int xsize = min(min((int)__fadd_rn(__fadd_rn(center, support), 0.5f), in_size) - xmin, PP_MAX_TAPS);
  float* w = t.alpha + (long long)i * PP_MAX_TAPS;
  float total = 0.f;
  for (int j = 0; j < xsize; j++) {
    const float x = fabsf(__fmul_rn(__fadd_rn(__fsub_rn((float)(j + xmin), center), 0.5f), invscale));
    const float v = x < 1.f ? __fsub_rn(1.f, x) : 0.f;
    w[j] = v;
    total = __fadd_rn(total, v);
  }
  for (int j = 0; j < xsize; j++) w[j] = total != 0.f ? __fdiv_rn(w[j], total) : 0.f;
  t.start[i] = xmin;
  t.count[i] = xsize;
}
// one separable pass along x (axis = 0) or y (axis = 1) over planes [n][Hs][Ws]
__global__ void __launch_bounds__(256) pp_aa_pass_kernel(const float* __restrict__ in, float* __restrict__ out, AreaTab t, int Hs, int Ws,
                                                         int Hd, int Wd, int axis) {
  const int n = blockIdx.z, dy = blockIdx.y, dx = blockIdx.x * blockDim.x + threadIdx.x;
  if (dx >= Wd) return;
  const float* src = in + (long long)n * Hs * Ws;
  const int i = axis == 0 ? dx : dy;
  const int x0 = t.start[i], cnt = t.count[i];
  const float* w = t.alpha + (long long)i * PP_MAX_TAPS;
  float acc = 0.f;
  for (int j = 0; j < cnt; j++) {
    const float v = axis == 0 ? src[(long long)dy * Ws + x0 + j] : src[(long long)(x0 + j) * Ws + dx];
    acc = (j == 0) ? __fmul_rn(v, w[0]) : __fmaf_rn(v, w[j], acc);
  }
  out[((long long)n * Hd + dy) * Wd + dx] = acc;
}

// ---- match post-processing -------------------------------------------------------------------------------------------------
// match_features.py:244-257: valid = matches0 > -1; mkpts0 = kpts0[valid]; mkpts1 = kpts1[matches0[valid]]; mconf;
// *_orig = (k + 0.5) * scale - 0.5 in fp32 (torch float32 tensor times a scalar, then minus 0.5).  Ordered compaction:
// one CTA per pair, ballot prefix sums keep ascending keypoint order, as boolean-mask indexing does.
__global__ void __launch_bounds__(1024) pp_gather_matches_kernel(const float* __restrict__ kpts, const int* __restrict__ matches,
                                                                 const float* __restrict__ mscores, const int* __restrict__ counts,
                                                                 const float* __restrict__ scales, float* __restrict__ mk0,
                                                                 float* __restrict__ mk1, float* __restrict__ mk0o, float* __restrict__ mk1o,
                                                                 float* __restrict__ mconf, int* __restrict__ mcount, int cap) {
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid % 32, wid = tid / 32;
  const int z0 = 2 * p, z1 = 2 * p + 1, n0 = counts[z0], n1 = counts[z1];
  __shared__ int s_w[32];
  __shared__ int s_base;
  if (tid == 0) s_base = 0;
  __syncthreads();
  float sx0 = 1.f, sy0 = 1.f, sx1 = 1.f, sy1 = 1.f;
  if (scales) { sx0 = scales[2 * z0]; sy0 = scales[2 * z0 + 1]; sx1 = scales[2 * z1]; sy1 = scales[2 * z1 + 1]; }
  for (int i0 = 0; i0 < n0; i0 += 1024) {
    const int i = i0 + tid;
    int j = -1;
    if (i < n0) { j = matches[(long long)z0 * cap + i]; if ((unsigned)j >= (unsigned)n1) j = -1; }
    const bool ok = j > -1;
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_w[wid] = __popc(bal);
    __syncthreads();
    int woff = 0, tot = 0;
    for (int w = 0; w < 32; w++) { const int cnt = s_w[w]; if (w < wid) woff += cnt; tot += cnt; }
    if (ok) {
      const long long d = (long long)p * cap + s_base + woff + __popc(bal & ((1u << lane) - 1u));
      const float x0 = kpts[((long long)z0 * cap + i) * 2], y0 = kpts[((long long)z0 * cap + i) * 2 + 1];
      const float x1 = kpts[((long long)z1 * cap + j) * 2], y1 = kpts[((long long)z1 * cap + j) * 2 + 1];
      mk0[2 * d] = x0; mk0[2 * d + 1] = y0; mk1[2 * d] = x1; mk1[2 * d + 1] = y1;
      if (mk0o) {
        mk0o[2 * d] = __fsub_rn(__fmul_rn(__fadd_rn(x0, 0.5f), sx0), 0.5f); mk0o[2 * d + 1] = __fsub_rn(__fmul_rn(__fadd_rn(y0, 0.5f), sy0), 0.5f);
        mk1o[2 * d] = __fsub_rn(__fmul_rn(__fadd_rn(x1, 0.5f), sx1), 0.5f); mk1o[2 * d + 1] = __fsub_rn(__fmul_rn(__fadd_rn(y1, 0.5f), sy1), 0.5f);
      }
      if (mconf) mconf[d] = mscores ? mscores[(long long)z0 * cap + i] : 0.f;
    }
    __syncthreads();
    if (tid == 0) s_base += tot;
    __syncthreads();
  }
  if (tid == 0) mcount[p] = s_base;
}

// keypoints_orig = (k + 0.5) * scale - 0.5 for every keypoint of every image (match_features.py:251-254, match_dense.py:655-656)
__global__ void pp_rescale_kernel(const float* __restrict__ kpts, const int* __restrict__ counts, const float* __restrict__ scales,
                                  float* __restrict__ out, int cap) {
  const int z = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (counts ? counts[z] : cap)) return;
  const long long o = ((long long)z * cap + i) * 2;
  out[o] = __fsub_rn(__fmul_rn(__fadd_rn(kpts[o], 0.5f), scales[2 * z]), 0.5f);
  out[o + 1] = __fsub_rn(__fmul_rn(__fadd_rn(kpts[o + 1], 0.5f), scales[2 * z + 1]), 0.5f);
}

struct Plan { int Co, h1, w1, lin1, h2, w2, lin2, hf, wf, n_resize; };

// size logic of extract_features.py:120-156
bool make_plan(const imw_pre_conf* c, int H0, int W0, int C, Plan& p) {
  p.Co = (C == 3 && c->grayscale) ? 1 : C;
  int h = H0, w = W0;
  p.n_resize = 0; p.h1 = p.w1 = p.h2 = p.w2 = 0; p.lin1 = p.lin2 = 0;
  if (c->resize_max > 0) {
    const double scale = (double)c->resize_max / (double)(w > h ? w : h);
    if (scale < 1.0) {
      const int wn = (int)nearbyint(w * scale), hn = (int)nearbyint(h * scale);   // Python round(): half to even
      p.h1 = hn; p.w1 = wn; p.lin1 = (w < wn || h < hn); p.n_resize = 1; h = hn; w = wn;
    }
  }
  if (c->force_resize) {
    int* hh = p.n_resize ? &p.h2 : &p.h1; int* ww = p.n_resize ? &p.w2 : &p.w1; int* ll = p.n_resize ? &p.lin2 : &p.lin1;
    *hh = c->height; *ww = c->width; *ll = (w < c->width || h < c->height);
    p.n_resize++; h = c->height; w = c->width;
  }
  const int df = c->dfactor > 0 ? c->dfactor : 1;
  p.hf = h / df * df; p.wf = w / df * df;
  return h > 0 && w > 0 && p.hf > 0 && p.wf > 0;
}

struct PPBuffers { float *t0, *t1; AreaTab xt, yt; };
size_t pp_carve(Workspace& ws, PPBuffers& b, int B, int H0, int W0, const Plan& p) {
  size_t m = 0;
  auto upd = [&](int h, int w) { size_t n = (size_t)B * p.Co * h * w; if (n > m) m = n; };
  upd(p.h1, p.w1); upd(p.h2, p.w2); upd(H0, W0);
  b.t0 = ws.take<float>(m); b.t1 = ws.take<float>(m);
  int lim = H0 > W0 ? H0 : W0;
  if (p.h1 > lim) lim = p.h1; if (p.w1 > lim) lim = p.w1; if (p.h2 > lim) lim = p.h2; if (p.w2 > lim) lim = p.w2;
  for (AreaTab* t : {&b.xt, &b.yt}) { t->start = ws.take<int>(lim); t->count = ws.take<int>(lim); t->alpha = ws.take<float>((size_t)lim * PP_MAX_TAPS); }
  return ws.off;
}

}  // namespace

extern "C" int imw_preprocess_plan(const imw_pre_conf* conf, int height, int width, int channels, int* out_channels, int* out_height,
                                   int* out_width, size_t* workspace_bytes) {
  IMW_REQUIRE(conf && (channels == 1 || channels == 3) && height > 0 && width > 0, "imw_preprocess_plan: bad arguments");
  Plan p;
  IMW_REQUIRE(make_plan(conf, height, width, channels, p), "imw_preprocess_plan: empty output for %dx%d", height, width);
  if (out_channels) *out_channels = p.Co;
  if (out_height) *out_height = p.hf;
  if (out_width) *out_width = p.wf;
  if (workspace_bytes) {
    Workspace ws(nullptr, 0);
    PPBuffers b;
    *workspace_bytes = pp_carve(ws, b, 1, height, width, p) + 256;   // per image: multiply the plane part by the batch
  }
  return IMW_OK;
}

extern "C" size_t imw_preprocess_workspace_bytes(const imw_pre_conf* conf, int batch, int height, int width, int channels) {
  Plan p;
  if (!conf || !make_plan(conf, height, width, channels, p)) return 0;
  Workspace ws(nullptr, 0);
  PPBuffers b;
  return pp_carve(ws, b, batch, height, width, p) + 256;
}

extern "C" int imw_preprocess(const imw_pre_conf* conf, int batch, int height, int width, int channels, const unsigned char* images,
                              float* out, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  IMW_REQUIRE(conf && images && out && batch > 0 && (channels == 1 || channels == 3), "imw_preprocess: bad arguments");
  Plan p;
  IMW_REQUIRE(make_plan(conf, height, width, channels, p), "imw_preprocess: empty output for %dx%d", height, width);
  Workspace ws(workspace, workspace_bytes);
  PPBuffers b;
  pp_carve(ws, b, batch, height, width, p);
  if (ws.overflow) { imw_set_error("imw_preprocess: workspace too small (%zu < %zu)", workspace_bytes, ws.off); return IMW_ERR_WORKSPACE; }
  const int Co = p.Co;
  // stage list: resize 1, resize 2 (cv2), then the antialias alignment; /255 rides on the last cv2 stage (or the conversion)
  int hs = height, ws_ = width;
  const bool need_aa = (p.n_resize == 0 ? (p.hf != height || p.wf != width)
                                        : (p.hf != (p.n_resize == 2 ? p.h2 : p.h1) || p.wf != (p.n_resize == 2 ? p.w2 : p.w1)));
  Src src{nullptr, images, height, width, channels, conf->grayscale ? 1 : 0, Co};
  float* cur = nullptr;
  if (p.n_resize == 0) {
    float* dst = need_aa ? b.t0 : out;
    pp_convert_kernel<<<dim3(ceil_div(width, 256), height, batch), 256, 0, st>>>(images, dst, height, width, channels, conf->grayscale ? 1 : 0, Co, 1);
    IMW_CHECK_LAUNCH();
    cur = dst;
  }
  for (int r = 0; r < p.n_resize; r++) {
    const int hd = r == 0 ? p.h1 : p.h2, wd = r == 0 ? p.w1 : p.w2, lin = r == 0 ? p.lin1 : p.lin2;
    const bool last = r == p.n_resize - 1;
    float* dst = (last && !need_aa) ? out : (r == 0 ? b.t0 : b.t1);
    if (r > 0) src = Src{cur, nullptr, hs, ws_, 1, 0, Co};
    const dim3 grid(ceil_div(wd, 256), hd, batch * Co);
    if (lin) {
      pp_linear_kernel<<<grid, 256, 0, st>>>(src, dst, hd, wd, last ? 1 : 0);
      IMW_CHECK_LAUNCH();
    } else {
      const double sxd = 1.0 / ((double)wd / (double)ws_), syd = 1.0 / ((double)hd / (double)hs);
      const int isx = (int)nearbyint(sxd), isy = (int)nearbyint(syd);
      if (fabs(sxd - isx) < 2.220446049250313e-16 && fabs(syd - isy) < 2.220446049250313e-16) {
        pp_area_fast_kernel<<<grid, 256, 0, st>>>(src, dst, hd, wd, isx, isy, last ? 1 : 0);
        IMW_CHECK_LAUNCH();
      } else {
        IMW_REQUIRE(sxd + 2 < PP_MAX_TAPS && syd + 2 < PP_MAX_TAPS, "imw_preprocess: down-scaling factor above %d", PP_MAX_TAPS - 2);
        pp_area_tab_kernel<<<ceil_div(wd, 128), 128, 0, st>>>(b.xt, ws_, wd);
        IMW_CHECK_LAUNCH();
        pp_area_tab_kernel<<<ceil_div(hd, 128), 128, 0, st>>>(b.yt, hs, hd);
        IMW_CHECK_LAUNCH();
        pp_area_kernel<<<grid, 256, 0, st>>>(src, b.xt, b.yt, dst, hd, wd, last ? 1 : 0);
        IMW_CHECK_LAUNCH();
      }
    }
    cur = dst; hs = hd; ws_ = wd;
  }
  if (need_aa) {   // F.resize(image, size_new, antialias=True): width pass, then height pass
    float* a = cur;
    if (p.wf != ws_) {
      float* dst = (p.hf == hs) ? out : (a == b.t0 ? b.t1 : b.t0);
      IMW_REQUIRE((double)ws_ / p.wf * 2 + 2 < PP_MAX_TAPS, "imw_preprocess: antialias support too wide");
      pp_aa_tab_kernel<<<ceil_div(p.wf, 128), 128, 0, st>>>(b.xt, ws_, p.wf);
      IMW_CHECK_LAUNCH();
      pp_aa_pass_kernel<<<dim3(ceil_div(p.wf, 256), hs, batch * Co), 256, 0, st>>>(a, dst, b.xt, hs, ws_, hs, p.wf, 0);
      IMW_CHECK_LAUNCH();
      a = dst; ws_ = p.wf;
    }
    if (p.hf != hs) {
      pp_aa_tab_kernel<<<ceil_div(p.hf, 128), 128, 0, st>>>(b.yt, hs, p.hf);
      IMW_CHECK_LAUNCH();
      pp_aa_pass_kernel<<<dim3(ceil_div(p.wf, 256), p.hf, batch * Co), 256, 0, st>>>(a, out, b.yt, hs, ws_, p.hf, p.wf, 1);
      IMW_CHECK_LAUNCH();
    }
  }
  return IMW_OK;
}

extern "C" int imw_gather_matches(int n_pairs, int cap, const float* keypoints, const int* matches, const float* matching_scores,
                                  const int* counts, const float* scales, float* mkpts0, float* mkpts1, float* mkpts0_orig,
                                  float* mkpts1_orig, float* mconf, int* mcount, cudaStream_t st) {
  IMW_REQUIRE(n_pairs > 0 && cap > 0 && keypoints && matches && counts && mkpts0 && mkpts1 && mcount, "imw_gather_matches: bad arguments");
  IMW_REQUIRE((mkpts0_orig == nullptr) == (mkpts1_orig == nullptr), "imw_gather_matches: both *_orig outputs or none");
  pp_gather_matches_kernel<<<n_pairs, 1024, 0, st>>>(keypoints, matches, matching_scores, counts, scales, mkpts0, mkpts1, mkpts0_orig,
                                                     mkpts1_orig, mconf, mcount, cap);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}

extern "C" int imw_rescale_keypoints(int n_sets, int cap, const float* keypoints, const int* counts, const float* scales, float* out,
                                     cudaStream_t st) {
  IMW_REQUIRE(n_sets > 0 && cap > 0 && keypoints && scales && out, "imw_rescale_keypoints: bad arguments");
  pp_rescale_kernel<<<dim3(ceil_div(cap, 256), n_sets), 256, 0, st>>>(keypoints, counts, scales, out, cap);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}
