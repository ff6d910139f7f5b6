// Probe (sm_100a): can a tcgen05.mma A-operand descriptor start INSIDE a SWIZZLE_128B atom (pixel-shifted window of a halo
// patch), with a stride-byte-offset that is not a multiple of 1024?  Layout under test: pixels as 128-byte rows written the way
// TMA SWIZZLE_128B writes them (16-byte unit c of row p at c ^ (p & 7), absolute-address pattern, buffer 1024-aligned);
// M index m = (r, c) of a 16 x 8 pixel tile -> halo pixel (r + dy) * PITCH + c + dx.
#include <cstdio>
#include <cuda_fp16.h>
#include "../../image-matching-webui_b200/csrc/tc_common.cuh"

__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) probe(int pitch, int dy, int dx, int mode, int kq, int* mism, float* dump) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // up to 18*16 pixels x 128 B = 36 KB
  uint8_t* sB = smem + 40 * 1024;     // 64 rows x 128 B
  uint64_t* bar = (uint64_t*)(smem + 50 * 1024);
  uint32_t* slot = (uint32_t*)(bar + 1);
  const int t = threadIdx.x, warp = t / 32;
  const int npix = 18 * pitch;
  for (int i = t; i < npix * 64; i += 128) {
    const int p = i / 64, k = i % 64;
    const int c = k / 8, e = k % 8;
    __half v = __float2half((float)((p % 128) * 16 + (k % 16)) + (k >= 16 ? 0.f : 0.f));
    if (k / 16 != kq) v = __float2half(-1.f);   // other k quarters hold -1: a wrong k offset shows
    *(__half*)(sA + p * 128 + ((c ^ (p & 7)) * 16) + e * 2) = v;
  }
  for (int i = t; i < 64 * 64; i += 128) {
    const int n = i / 64, k = i % 64;
    const int c = k / 8, e = k % 8;
    *(__half*)(sB + n * 128 + ((c ^ (n & 7)) * 16) + e * 2) = __float2half((k % 16) == n && (k / 16) == kq ? 1.f : 0.f);
  }
  if (t == 0) { tc::mbar_init(bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc(slot, 64);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *slot;
  if (t == 0) {
    const uint32_t a_addr = tc::smem_u32(sA) + (dy * pitch + dx) * 128 + kq * 32;
    const uint32_t bo = mode == 1 ? ((a_addr >> 7) & 7) : 0;
    const uint64_t ad = desc(a_addr, pitch * 128, bo), bd = desc(tc::smem_u32(sB) + kq * 32, 1024, 0);
    tc::mma_f16(tmem, ad, bd, tc::make_idesc(tc::FMT_F16, 128, 64), 0u);
    tc::mma_commit(bar);
  }
  tc::mbar_wait(bar, 0);
  tc::fence_after_sync();
  float v[16];
  tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
  const int m = t, r = m / 8, c = m % 8;
  const int p = (r + dy) * pitch + c + dx;
  int bad = 0;
  for (int n = 0; n < 16; n++) {
    const float want = (float)((p % 128) * 16 + n);
    if (v[n] != want) bad++;
    if (dump) dump[m * 16 + n] = v[n];
  }
  if (bad) atomicAdd(mism, 1);
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 64);
}

int main() {
  int* mism; float* dump;
  cudaMalloc(&mism, 4); cudaMalloc(&dump, 128 * 16 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int pitches[3] = {8, 10, 16};
  for (int pi = 0; pi < 3; pi++)
    for (int mode = 0; mode < 2; mode++)
      for (int dy = 0; dy < 3; dy++)
        for (int dx = 0; dx < 3; dx++)
          for (int kq = 0; kq < 4; kq += 3) {
            if (pitches[pi] == 8 && dx) continue;
            cudaMemset(mism, 0, 4);
            probe<<<1, 128, 64 * 1024>>>(pitches[pi], dy, dx, mode, kq, mism, dump);
            cudaError_t e = cudaDeviceSynchronize();
            int h = -1; cudaMemcpy(&h, mism, 4, cudaMemcpyDeviceToHost);
            printf("pitch %2d mode %d dy %d dx %d kq %d -> %s bad rows %d\n", pitches[pi], mode, dy, dx, kq, cudaGetErrorString(e), h);
            if (e != cudaSuccess) return 1;
            if (h && pitches[pi] == 10 && dy == 1 && dx == 1 && kq == 0) {
              float hd[128 * 16]; cudaMemcpy(hd, dump, sizeof hd, cudaMemcpyDeviceToHost);
              for (int m = 0; m < 24; m++) printf("  m %3d got pixel %g (k0 %g)\n", m, floorf(hd[m * 16] / 16), hd[m * 16] - 16 * floorf(hd[m * 16] / 16));
            }
          }
  return 0;
}
