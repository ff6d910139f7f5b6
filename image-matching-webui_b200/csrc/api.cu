// C-ABI entry points (include/imw_b200.h): SuperPoint forward and the hloc first-party matchers.
#include <stdarg.h>
#include <string.h>

#include "../../include/imw_b200.h"
#include "common.cuh"
#include "gemm_simt.cuh"
#include "simreduce.cuh"
#include "sp_kernels.h"
#include "tc_simreduce.cuh"

static thread_local char g_err[512] = "";
void imw_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* imw_last_error(void) { return g_err; }
extern "C" int imw_version(void) { return 100; }
unsigned long long g_imw_launches = 0;
extern "C" unsigned long long imw_launch_count(void) { return g_imw_launches; }

int imw_num_sms() {
  static int sms[64] = {0};
  const int d = imw_cur_device() & 63;
  if (!sms[d]) cudaDeviceGetAttribute(&sms[d], cudaDevAttrMultiProcessorCount, d);
  return sms[d] > 0 ? sms[d] : 148;
}

// ---- launch-site profiler ---------------------------------------------------------------------------
// imw_prof_begin(stream): record a start event and switch marking on; every launch of the library then records one
// event behind itself (IMW_COUNT_LAUNCH).  imw_prof_end(): synchronise the events, attribute each interval to the
// launch that ended it, aggregate per launch site (host launcher signature incl. template arguments + line) and
// render "site<TAB>launches<TAB>total_ms" lines into a caller buffer.  Used by bench.py for the live per-kernel
// share of the step and the roofline of the dominant kernel; off (zero cost beyond one branch) otherwise.
#include <map>
#include <string>
#include <vector>
int g_imw_prof_on = 0;
namespace {
struct ProfMark { cudaEvent_t ev; const char* site; int line; };
std::vector<cudaEvent_t> g_prof_pool;
std::vector<ProfMark> g_prof_marks;
cudaEvent_t g_prof_start = nullptr;
cudaEvent_t prof_event(size_t i) {
  while (g_prof_pool.size() <= i) { cudaEvent_t e; cudaEventCreate(&e); g_prof_pool.push_back(e); }
  return g_prof_pool[i];
}
}  // namespace
void imw_prof_mark(const char* site, int line, cudaStream_t st) {
  cudaEvent_t e = prof_event(g_prof_marks.size() + 1);
  cudaEventRecord(e, st);
  g_prof_marks.push_back({e, site, line});
}
extern "C" int imw_prof_begin(cudaStream_t st) {
  g_prof_marks.clear();
  g_prof_start = prof_event(0);
  IMW_CHECK_CUDA(cudaEventRecord(g_prof_start, st));
  g_imw_prof_on = 1;
  return IMW_OK;
}
extern "C" long long imw_prof_end(char* buf, size_t buf_bytes) {
  g_imw_prof_on = 0;
  if (g_prof_marks.empty()) { if (buf && buf_bytes) buf[0] = 0; return 0; }
  if (cudaEventSynchronize(g_prof_marks.back().ev) != cudaSuccess) return IMW_ERR_CUDA;
  std::map<std::string, std::pair<long long, double>> agg;
  cudaEvent_t prev = g_prof_start;
  for (const ProfMark& m : g_prof_marks) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, prev, m.ev);
    prev = m.ev;
    auto& a = agg[std::string(m.site) + ":" + std::to_string(m.line)];
    a.first += 1; a.second += ms;
  }
  std::string out;
  for (auto& kv : agg) out += kv.first + "\t" + std::to_string(kv.second.first) + "\t" + std::to_string(kv.second.second) + "\n";
  if (buf && buf_bytes) { size_t n = out.size() < buf_bytes - 1 ? out.size() : buf_bytes - 1; memcpy(buf, out.data(), n); buf[n] = 0; }
  const long long n = (long long)g_prof_marks.size();
  g_prof_marks.clear();
  return n;
}

// =====================================================================================================
// SuperPoint
// =====================================================================================================
namespace {
int SP_SUB = 32;  // images per pass through the conv stack (bounds the activation workspace); imw_debug_set_sp_sub

struct SPBuffers {
  float *a1, *a2, *a3, *a4, *a5, *a6, *a7, *a8, *pa, *logits, *da, *dense, *nms, *dd;
  unsigned long long* keys;
  size_t key_cap;
};

size_t sp_carve(Workspace& ws, SPBuffers& b, int B, int H, int W) {
  const size_t sb = (size_t)(B < SP_SUB ? B : SP_SUB);
  const size_t h = H / 8, w = W / 8;
  // the same buffers hold either fp32 or two fp16 planes (4 bytes per element either way)
  b.a1 = ws.take<float>(sb * H * W * 64);
  b.a2 = ws.take<float>(sb * (H / 2) * (W / 2) * 64);
  b.a3 = ws.take<float>(sb * (H / 2) * (W / 2) * 64);
  b.a4 = ws.take<float>(sb * (H / 4) * (W / 4) * 64);
  b.a5 = ws.take<float>(sb * (H / 4) * (W / 4) * 128);
  b.a6 = ws.take<float>(sb * h * w * 128);
  b.a7 = ws.take<float>(sb * h * w * 128);
  b.a8 = ws.take<float>(sb * h * w * 128);
  b.pa = ws.take<float>(sb * h * w * 256);
  b.logits = ws.take<float>(sb * h * w * 128);
  b.da = ws.take<float>(sb * h * w * 256);
  b.dense = ws.take<float>((size_t)B * H * W);
  b.nms = ws.take<float>((size_t)B * H * W);
  b.dd = ws.take<float>((size_t)B * h * w * 256);
  b.key_cap = sp_select_key_cap(H, W);
  b.keys = ws.take<unsigned long long>((size_t)B * b.key_cap);
  return ws.off;
}
}  // namespace

extern "C" int imw_debug_set_sp_sub(int n) {   // tuning hook: images per pass of the SuperPoint conv stack
  if (n > 0) SP_SUB = n;
  return SP_SUB;
}

extern "C" size_t imw_superpoint_workspace_bytes(int batch, int height, int width) {
  Workspace ws(nullptr, 0);
  SPBuffers b;
  return sp_carve(ws, b, batch, height, width) + 256;
}

extern "C" int imw_superpoint_forward(const imw_sp_weights* wt, const imw_sp_conf* conf, int B, int H, int W,
                                      const float* image, int cap, float* keypoints, float* scores, float* descriptors,
                                      int* counts, float* dense_out, void* workspace, size_t workspace_bytes,
                                      cudaStream_t st) {
  IMW_REQUIRE(wt && conf && image && B > 0, "imw_superpoint_forward: bad arguments");
  IMW_REQUIRE(H % 8 == 0 && W % 8 == 0 && H >= 16 && W >= 16, "imw_superpoint_forward: H,W must be multiples of 8 (got %dx%d)", H, W);
  // superpoint.py:139-141
  IMW_REQUIRE(!(conf->max_keypoints == 0 || conf->max_keypoints < -1), "\"max_keypoints\" must be positive or \"-1\"");
  IMW_REQUIRE(cap > 0, "imw_superpoint_forward: cap must be positive");
  Workspace ws(workspace, workspace_bytes);
  SPBuffers b;
  sp_carve(ws, b, B, H, W);
  if (ws.overflow) { imw_set_error("imw_superpoint_forward: workspace too small (%zu < %zu)", workspace_bytes, ws.off); return IMW_ERR_WORKSPACE; }
  const int h = H / 8, w = W / 8;
  const bool use_tc = conf->use_tensor_cores != 0;
  if (use_tc) {
    IMW_REQUIRE(W % 16 == 0, "imw_superpoint_forward: the tensor-core path needs W %% 16 == 0 (got %d)", W);
    for (int l : {1, 2, 3, 4, 5, 6, 7, 8, 10}) IMW_REQUIRE(wt->wp[l] != nullptr, "imw_superpoint_forward: fp16-plane weights missing for layer %d", l);
  }
  const bool tc_heads = use_tc && wt->wp[9] && wt->wp[11];
  int rc;
#define RUN(x) do { rc = (x); if (rc) return rc; } while (0)
  for (int b0 = 0; b0 < B; b0 += SP_SUB) {
    const int nb = (B - b0 < SP_SUB) ? (B - b0) : SP_SUB;
    const float* img = image + (size_t)b0 * H * W;
    if (use_tc) {
      // encoder + both 3x3 head convs on tcgen05, activations carried as two fp16 planes (split_planes.cuh)
      if (conf->use_tensor_cores == 1) {   // conv1a evaluated inside the conv1b kernel (no plane traffic for the first layer)
        RUN(tc_conv1ab_fused(img, wt->w[0], wt->b[0], wt->wp[1], wt->b[1], b.a2, nb, H, W, 1, st));
      } else {                             // 2: unfused pair (kept for A/B measurements)
        RUN(sp_conv3x3_c1(img, wt->w[0], wt->b[0], nullptr, b.a1, nb, H, W, st));
        RUN(tc_conv3x3(b.a1, wt->wp[1], wt->b[1], b.a2, nb, H, W, 64, 64, 1, 1, 0, st));
      }
      RUN(tc_conv3x3(b.a2, wt->wp[2], wt->b[2], b.a3, nb, H / 2, W / 2, 64, 64, 1, 0, 0, st));
      RUN(tc_conv3x3(b.a3, wt->wp[3], wt->b[3], b.a4, nb, H / 2, W / 2, 64, 64, 1, 1, 0, st));
      RUN(tc_conv3x3(b.a4, wt->wp[4], wt->b[4], b.a5, nb, H / 4, W / 4, 64, 128, 1, 0, 0, st));
      RUN(tc_conv3x3(b.a5, wt->wp[5], wt->b[5], b.a6, nb, H / 4, W / 4, 128, 128, 1, 1, 0, st));
      RUN(tc_conv3x3(b.a6, wt->wp[6], wt->b[6], b.a7, nb, h, w, 128, 128, 1, 0, 0, st));
      RUN(tc_conv3x3(b.a7, wt->wp[7], wt->b[7], b.a8, nb, h, w, 128, 128, 1, 0, 0, st));
      const int head_planes = tc_heads ? 0 : 1;   // 1x1 heads on tcgen05 too: keep the 3x3 head outputs as planes
      RUN(tc_conv3x3(b.a8, wt->wp[8], wt->b[8], b.pa, nb, h, w, 128, 256, 1, 0, head_planes, st));
      RUN(tc_conv3x3(b.a8, wt->wp[10], wt->b[10], b.da, nb, h, w, 128, 256, 1, 0, head_planes, st));
    } else {
      RUN(sp_conv3x3_c1(img, wt->w[0], wt->b[0], b.a1, nullptr, nb, H, W, st));
      RUN(sp_conv3x3(b.a1, wt->w[1], wt->b[1], b.a2, nb, H, W, 64, 64, 1, 1, st));
      RUN(sp_conv3x3(b.a2, wt->w[2], wt->b[2], b.a3, nb, H / 2, W / 2, 64, 64, 1, 0, st));
      RUN(sp_conv3x3(b.a3, wt->w[3], wt->b[3], b.a4, nb, H / 2, W / 2, 64, 64, 1, 1, st));
      RUN(sp_conv3x3(b.a4, wt->w[4], wt->b[4], b.a5, nb, H / 4, W / 4, 64, 128, 1, 0, st));
      RUN(sp_conv3x3(b.a5, wt->w[5], wt->b[5], b.a6, nb, H / 4, W / 4, 128, 128, 1, 1, st));
      RUN(sp_conv3x3(b.a6, wt->w[6], wt->b[6], b.a7, nb, h, w, 128, 128, 1, 0, st));
      RUN(sp_conv3x3(b.a7, wt->w[7], wt->b[7], b.a8, nb, h, w, 128, 128, 1, 0, st));
      RUN(sp_conv3x3(b.a8, wt->w[8], wt->b[8], b.pa, nb, h, w, 128, 256, 1, 0, st));   // detector head (superpoint.py:165)
      RUN(sp_conv3x3(b.a8, wt->w[10], wt->b[10], b.da, nb, h, w, 128, 256, 1, 0, st)); // descriptor head (superpoint.py:194)
    }
    // detector head 1x1 + softmax + depth-to-space (superpoint.py:166-170)
    if (tc_heads) {  // 65 outputs zero-padded to 128 (weights and bias padded by the host)
      RUN(tc_conv_general(b.pa, wt->wp[9], wt->b[9], nullptr, b.logits, nb, h, w, 256, 128, 1, 1, 0, 1, st));
      RUN(sp_softmax_d2s(b.logits, b.dense + (size_t)b0 * H * W, nb, h, w, st, 128));
    } else {
      GemmArgs g{};
      g.A = b.pa; g.lda = 256; g.W = wt->w[9]; g.ldw = 256; g.M = nb * h * w; g.N = 65; g.K = 256;
      IMW_CHECK_CUDA(launch_gemm(g, 1, EpiBias{b.logits, 0, 65, wt->b[9], 0}, st));
      RUN(sp_softmax_d2s(b.logits, b.dense + (size_t)b0 * H * W, nb, h, w, st));
    }
    // descriptor head 1x1 + channel L2 norm (superpoint.py:195-196)
    {
      float* dd = b.dd + (size_t)b0 * h * w * 256;
      if (tc_heads) {
        RUN(tc_conv_general(b.da, wt->wp[11], wt->b[11], nullptr, dd, nb, h, w, 256, 256, 1, 1, 0, 1, st));
      } else {
        GemmArgs g{};
        g.A = b.da; g.lda = 256; g.W = wt->w[11]; g.ldw = 256; g.M = nb * h * w; g.N = 256; g.K = 256;
        IMW_CHECK_CUDA(launch_gemm(g, 1, EpiBias{dd, 0, 256, wt->b[11], 0}, st));
      }
      RUN(sp_l2norm_rows(dd, (long long)nb * h * w, 256, st));
    }
  }
  if (dense_out) IMW_CHECK_CUDA(cudaMemcpyAsync(dense_out, b.dense, (size_t)B * H * W * sizeof(float), cudaMemcpyDeviceToDevice, st));
  RUN(sp_nms(b.dense, b.nms, B, H, W, conf->nms_radius, st));
  RUN(sp_select(b.nms, b.keys, (int)b.key_cap, keypoints, scores, counts, B, H, W, conf->keypoint_threshold, conf->remove_borders,
                conf->max_keypoints, cap, st));
  RUN(sp_sample_desc(b.dd, keypoints, counts, descriptors, B, h, w, cap, 256, conf->fix_sampling, st));
#undef RUN
  return IMW_OK;
}

// single 3x3 conv layer (NHWC fp32, weights [9][Cin][Cout]); used by bench.py to time the dominant kernel alone
extern "C" int imw_debug_conv3x3(const float* in, const float* w, const float* bias, float* out, int B, int H, int W, int Cin,
                                 int Cout, int relu, int pool, cudaStream_t st) {
  return sp_conv3x3(in, w, bias, out, B, H, W, Cin, Cout, relu, pool, st);
}

// =====================================================================================================
// hloc first-party matchers
// =====================================================================================================
namespace {

struct Top2State { float v1, v2; int j1; };

// top-2 similarities per row + find_nn thresholds (nearest_neighbor.py:6-16)
struct OpTop2 {
  using State = Top2State;
  int* match; float* score; const int* counts; int cap; float ratio2, dist2;  // squared thresholds, <=0: off
  __device__ void init(State& s) const { s.v1 = -INFINITY; s.v2 = -INFINITY; s.j1 = 0x7fffffff; }
  __device__ void accum(State& s, float v, int, int j, int, int) const {
    if (v > s.v1 || (v == s.v1 && j < s.j1)) { s.v2 = s.v1; s.v1 = v; s.j1 = j; }
    else if (v > s.v2) s.v2 = v;
  }
  __device__ State shfl_xor(const State& s, int o) const {
    State t;
    t.v1 = __shfl_xor_sync(0xffffffffu, s.v1, o); t.v2 = __shfl_xor_sync(0xffffffffu, s.v2, o);
    t.j1 = __shfl_xor_sync(0xffffffffu, s.j1, o);
    return t;
  }
  __device__ void merge(State& a, const State& b) const {
    if (b.v1 > a.v1 || (b.v1 == a.v1 && b.j1 < a.j1)) { a.v2 = fmaxf(b.v2, a.v1); a.v1 = b.v1; a.j1 = b.j1; }
    else a.v2 = fmaxf(a.v2, b.v1);
  }
  __device__ void store(const State& s, int own, int i) const {
    const int m = counts[own ^ 1], n = counts[own];
    float d0 = 2.f * (1.f - s.v1), d1 = 2.f * (1.f - s.v2);
    bool ok = true;
    if (ratio2 > 0.f && n > 1 && m > 1) ok = ok && (d0 <= ratio2 * d1);  // :52-53: ratio test off for single descriptors
    if (dist2 > 0.f) ok = ok && (d0 <= dist2);
    ok = ok && (unsigned)s.j1 < (unsigned)m;   // NaN similarities leave j1 at its init value
    match[(long long)own * cap + i] = ok ? s.j1 : -1;
    score[(long long)own * cap + i] = ok ? (s.v1 + 1.f) / 2.f : 0.f;
  }
};

// mutual_check (nearest_neighbor.py:19-24) and packing to [P][cap]
__global__ void nn_finish_kernel(const int* __restrict__ m, const float* __restrict__ sc, const int* __restrict__ counts,
                                 int* __restrict__ matches0, float* __restrict__ scores0, int cap, int mutual) {
  const int p = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  int out = -1;
  float s = 0.f;
  if (i < counts[2 * p] && counts[2 * p + 1] > 0) {
    int j = m[(long long)(2 * p) * cap + i];
    s = sc[(long long)(2 * p) * cap + i];
    out = j;
    if (mutual && j > -1 && m[(long long)(2 * p + 1) * cap + j] != i) out = -1;
  }
  matches0[(long long)p * cap + i] = out;
  scores0[(long long)p * cap + i] = s;
}

// dual_softmax.py:18-20: desc / desc.norm(dim=1)
__global__ void __launch_bounds__(256) normalize_rows_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                             const int* __restrict__ counts, int cap, int dim) {
  const int z = blockIdx.y, row = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (row >= counts[z]) return;
  const float* p = in + ((long long)z * cap + row) * dim;
  float ss = 0.f;
  for (int c = lane; c < dim; c += 32) ss += p[c] * p[c];
  float nrm = sqrtf(warp_sum(ss));
  float* o = out + ((long long)z * cap + row) * dim;
  for (int c = lane; c < dim; c += 32) o[c] = __fdiv_rn(p[c], nrm);
}

// (P == row max) & (P == col max) & (P > thr)  (dual_softmax.py:24-28)
__global__ void dsm_finish_kernel(const float* __restrict__ best_v, const int* __restrict__ best_j, const int* __restrict__ counts,
                                  int* __restrict__ matches0, float* __restrict__ scores0, int cap, float thr) {
  const int p = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  int out = -1;
  float s = 0.f;
  if (i < counts[2 * p] && counts[2 * p + 1] > 0) {
    long long io = (long long)(2 * p) * cap + i;
    int j = best_j[io];
    float v = best_v[io];
    // NaN rows (zero-norm descriptors) leave the arg-max at its init value: "no match", as the reference returns
    if ((unsigned)j < (unsigned)counts[2 * p + 1] && best_j[(long long)(2 * p + 1) * cap + j] == i && v > thr) { out = j; s = v; }
  }
  matches0[(long long)p * cap + i] = out;
  scores0[(long long)p * cap + i] = s;
}

struct MatcherBuffers { float *norm, *f0, *f1, *bv; int *bj, *m; plane_t* planes; };
size_t matcher_carve(Workspace& ws, MatcherBuffers& b, int P, int cap) {
  const size_t T = (size_t)2 * P * cap;
  b.norm = ws.take<float>(T * 256); b.f0 = ws.take<float>(T); b.f1 = ws.take<float>(T); b.bv = ws.take<float>(T);
  b.bj = ws.take<int>(T); b.m = ws.take<int>(T);
  b.planes = ws.take<plane_t>(T * 256 * 2);   // split-fp16 planes of the descriptors for the tcgen05 similarity kernel
  return ws.off;
}
}  // namespace

extern "C" size_t imw_matcher_workspace_bytes(int n_pairs, int cap) {
  Workspace ws(nullptr, 0);
  MatcherBuffers b;
  return matcher_carve(ws, b, n_pairs, cap) + 256;
}

extern "C" int imw_nearest_neighbor(int P, int cap, int dim, const float* desc, const int* counts, float ratio_threshold,
                                    float distance_threshold, int do_mutual_check, int use_tensor_cores, int* matches0,
                                    float* scores0, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  IMW_REQUIRE(P > 0 && cap > 0 && dim > 0 && dim % 4 == 0 && dim <= 256, "imw_nearest_neighbor: dim %% 4 == 0, dim <= 256 (got %d)", dim);
  Workspace ws(workspace, workspace_bytes);
  MatcherBuffers b;
  matcher_carve(ws, b, P, cap);
  if (ws.overflow) { imw_set_error("imw_nearest_neighbor: workspace too small"); return IMW_ERR_WORKSPACE; }
  SimArgs sa{desc, cap, dim, dim, counts, nullptr};
  OpTop2 op{b.m, b.f0, counts, cap, ratio_threshold > 0.f ? ratio_threshold * ratio_threshold : 0.f,
            distance_threshold > 0.f ? distance_threshold * distance_threshold : 0.f};
  if (use_tensor_cores && tc_simreduce_ok(sa)) {
    if (int e = tc_simreduce_split(sa, 2 * P, b.planes, st)) return e;
    if (int e = launch_tc_simreduce(sa, 2 * P, b.planes, op, st)) return e;
  } else IMW_CHECK_CUDA(launch_simreduce(sa, 2 * P, op, st));
  nn_finish_kernel<<<dim3(ceil_div(cap, 256), P), 256, 0, st>>>(b.m, b.f0, counts, matches0, scores0, cap, do_mutual_check);
  IMW_CHECK_LAUNCH_T("nn_finish_kernel");
  return IMW_OK;
}

extern "C" int imw_dual_softmax(int P, int cap, int dim, const float* desc, const int* counts, float match_threshold,
                                float inv_temperature, int use_tensor_cores, int* matches0, float* scores0, void* workspace,
                                size_t workspace_bytes, cudaStream_t st) {
  IMW_REQUIRE(P > 0 && cap > 0 && dim > 0 && dim % 4 == 0 && dim <= 256, "imw_dual_softmax: dim %% 4 == 0, dim <= 256 (got %d)", dim);
  Workspace ws(workspace, workspace_bytes);
  MatcherBuffers b;
  matcher_carve(ws, b, P, cap);
  if (ws.overflow) { imw_set_error("imw_dual_softmax: workspace too small"); return IMW_ERR_WORKSPACE; }
  normalize_rows_kernel<<<dim3(ceil_div(cap, 8), 2 * P), 256, 0, st>>>(desc, b.norm, counts, cap, dim);
  IMW_CHECK_LAUNCH_T("normalize_rows_kernel");
  SimArgs sa{b.norm, cap, dim, dim, counts, nullptr};
  if (use_tensor_cores && tc_simreduce_ok(sa)) {
    if (int e = tc_simreduce_split(sa, 2 * P, b.planes, st)) return e;
    if (int e = launch_tc_simreduce(sa, 2 * P, b.planes, OpSoftmaxStats{b.f0, b.f1, (float*)b.m, cap, inv_temperature}, st)) return e;
    if (int e = launch_tc_simreduce(sa, 2 * P, b.planes, OpDualSoftmaxArgmax{b.f0, b.f1, (const float*)b.m, b.bv, b.bj, cap, inv_temperature}, st)) return e;
  } else {
    IMW_CHECK_CUDA(launch_simreduce(sa, 2 * P, OpSoftmaxStats{b.f0, b.f1, (float*)b.m, cap, inv_temperature}, st));
    IMW_CHECK_CUDA(launch_simreduce(sa, 2 * P, OpDualSoftmaxArgmax{b.f0, b.f1, (const float*)b.m, b.bv, b.bj, cap, inv_temperature}, st));
  }
  dsm_finish_kernel<<<dim3(ceil_div(cap, 256), P), 256, 0, st>>>(b.bv, b.bj, counts, matches0, scores0, cap, match_threshold);
  IMW_CHECK_LAUNCH_T("dsm_finish_kernel");
  return IMW_OK;
}
