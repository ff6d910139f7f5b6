// Shared helpers for the sm_100a kernels of the B200 image-matching hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define IMW_OK 0
#define IMW_ERR_ARG -1
#define IMW_ERR_CUDA -2
#define IMW_ERR_WORKSPACE -3
#define IMW_ERR_UNSUPPORTED -4

void imw_set_error(const char* fmt, ...);

#define IMW_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      imw_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,                 \
                    cudaGetErrorString(_e));                                              \
      return IMW_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

// every kernel launch of the library is counted (bench.py reports it as gpu_launches)
extern unsigned long long g_imw_launches;
// launch-site profiler (imw_prof_begin / imw_prof_end, api.cu): when enabled, a CUDA event is recorded on the launching
// stream after every launch; the interval between consecutive events is that launch's duration inside the real step
extern int g_imw_prof_on;
void imw_prof_mark(const char* site, int line, cudaStream_t st);
#define IMW_COUNT_LAUNCH(stream_)                                             \
  do {                                                                        \
    ++g_imw_launches;                                                         \
    if (g_imw_prof_on) imw_prof_mark(__PRETTY_FUNCTION__, __LINE__, stream_); \
  } while (0)
#define IMW_CHECK_LAUNCH()               \
  do {                                   \
    IMW_COUNT_LAUNCH(st);                \
    IMW_CHECK_CUDA(cudaGetLastError());  \
  } while (0)
// same, with the kernel's name as the profile key (for launches written inline in a long forward function)
#define IMW_CHECK_LAUNCH_T(tag)                                 \
  do {                                                          \
    ++g_imw_launches;                                           \
    if (g_imw_prof_on) imw_prof_mark(tag, __LINE__, st);        \
    IMW_CHECK_CUDA(cudaGetLastError());                         \
  } while (0)

#define IMW_REQUIRE(cond, ...)                                                            \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      imw_set_error(__VA_ARGS__);                                                         \
      return IMW_ERR_ARG;                                                                 \
    }                                                                                     \
  } while (0)

// Function attributes and the SM count are per DEVICE: cache them per device ordinal, not per process
// (one process may drive several GPUs through torch.cuda.device(dev)).
static inline int imw_cur_device() { int d = 0; cudaGetDevice(&d); return d; }
int imw_num_sms();  // SM count of the current device (api.cu, cached per ordinal)
#define IMW_SMEM_ATTR_ONCE(kernel, bytes)                                                                       \
  do {                                                                                                          \
    static unsigned long long done_ = 0;                                                                        \
    const int d_ = imw_cur_device() & 63;                                                                       \
    if (!((done_ >> d_) & 1ull)) {                                                                              \
      IMW_CHECK_CUDA(cudaFuncSetAttribute((kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      done_ |= 1ull << d_;                                                                                      \
    }                                                                                                           \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace (no allocation inside the library).
struct Workspace {
  char* base;
  size_t size;
  size_t off;
  bool overflow;
  Workspace(void* p, size_t n) : base((char*)p), size(n), off(0), overflow(false) {}
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    size_t bytes = count * sizeof(T);
    T* p = (T*)(base ? base + off : nullptr);
    off += bytes;
    if (base && off > size) overflow = true;
    return p;
  }
};

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
