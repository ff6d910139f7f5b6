// tcgen05 implicit-GEMM 3x3 convolution with fp32-equivalent precision (two fp16 planes per operand, three products).
//
// Why split precision: SuperPoint's detector branch feeds exact-equality NMS, a hard threshold and a top-k
// cut; single TF32/BF16 operands flip 1.5 % / 8.7 % of the keypoints.  Every fp32
// value is carried as two fp16 planes x = hi + lo * 2^-11 (split_planes.cuh) and a product is assembled from
//     a_hi b_hi + (a_hi b_lo + a_lo b_hi) * 2^-11        (dropped term <= 2^-22 |a||b|)
// accumulated in fp32 in TMEM = 1/3 of the fp16 tensor peak.  The two weight planes of a stage are contiguous in N
// ([b_hi | b_lo] rows), so the three products are issued as TWO kind::f16 MMAs per k-step
//     a_hi x [b_hi | b_lo] -> [main | cross]          a_lo x b_hi -> cross
// (round 1 carried three bf16 planes and six products in four MMAs: twice the tensor work and operand traffic).
// The tensor core adds into its fp32 accumulator with truncation, a bias that grows with the number of
// accumulating MMAs (measured: ~1e-5 relative after 216 MMAs into one accumulator).  The accumulation is therefore
// spread over two [main | cross] accumulator sets that alternate k-step by k-step, summed in fp32 in the epilogue.
//
// Generic kernel (tc_conv3x3_kernel<BN, KS, RES>): GEMM view M = 128 output pixels (8 rows x 16 cols of one image),
// N = Cout tile (64 / 128), K = taps x Cin; 3x3 (pad 1) or 1x1, stride 1 or 2.  The taps are shifted TMA box loads of
// the NHWC activation planes (4-D tensor map, element strides for stride 2, out-of-bounds coordinates zero-filled =
// the conv's zero padding), 64 channels = one 128-byte swizzled row per pixel.
//   warp 0: TMA producer (2 activation planes + 2 weight planes per (tap, 64-channel chunk) stage)
//   warp 1: TMEM alloc + tcgen05.mma issue (8 MMAs per stage)
//   warps 2-9: epilogue (two per TMEM sub-partition): tcgen05.ld, bias, optional residual planes, ReLU / LeakyReLU, optional fused 2x2 max-pool
//              (lane shuffles: the 2x2 window lives in one warp), re-split into fp16 planes (or fp32) and store NHWC.
// Cin = Cout = 64 specialisation (tc_conv3x3_c64_kernel<FUSE>): persistent, halo tile as three dx-shifted copies,
// double-buffered TMEM; FUSE evaluates SuperPoint's conv1a inside the CTA (see below).
#include "split_planes.cuh"

#include "../../include/imw_b200.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int CV_TH = 8, CV_TW = 16;          // pixel tile (rows x cols) = 128 = MMA M
constexpr int CV_CK = 64;                     // channels per stage (128 B of bf16)
constexpr int CV_A_BYTES = 128 * 128;         // one activation plane tile
constexpr int CV_EPI_WARPS = 8;               // two warps per TMEM sub-partition, half of the Cout slice each (the epilogue is exposed time)
constexpr int CV_THREADS = 64 + 32 * CV_EPI_WARPS;

struct ConvArgs {
  int H, W, Cin, Cout, B;     // H, W: INPUT size; output = ceil(H/stride) x ceil(W/stride) (then /2 if pooled)
  int relu, pool, out_fp32;   // relu: 0 none, 1 ReLU, 2 LeakyReLU(0.01)
  const float* bias;
  plane_t* out_planes;        // [NP][B][Ho][Wo][Cout]
  float* out_f32;             // [B][Ho][Wo][Cout]
  int ksize = 3, stride = 1;  // 3x3 (pad 1) or 1x1 (pad 0); stride 1 or 2 (TMA element strides)
  const plane_t* res_planes = nullptr;  // optional residual [NP][B][Ho][Wo][Cout], added before the activation
  // fused first layer (tc_conv3x3_c64_kernel<true>): the activation operand is conv1a(image) computed in the CTA
  const float* img = nullptr;   // [B][H][W] fp32
  const float* w1a = nullptr;   // [9][64]
  const float* b1a = nullptr;   // [64]
};

template <int BN>
constexpr int conv_stage_bytes() { return NP * CV_A_BYTES + NP * BN * 128; }
template <int BN>
constexpr int conv_stages() { return BN == 64 ? 4 : 3; }
template <int BN>
constexpr size_t conv_smem_bytes() { return (size_t)conv_stages<BN>() * conv_stage_bytes<BN>() + 1024 + 256; }

// KS: kernel size (3 or 1) and RES: residual input are compile-time so that the SuperPoint instantiation <BN, 3, false>
// keeps constant tap arithmetic in the single-thread producer / MMA loops and a lean epilogue.
// (Round 2 tried a cluster of two pixel tiles with the weight planes TMA-multicast to both CTAs -- tc::tma_load_2d_mc, multicast
// tcgen05.commit on both `empty` barriers: parity-green, but no faster (LoFTR backbone 177.9 vs 174.9 ms): requests of
// neighbouring SMs for the same L2 lines are already merged below cluster size ~4 and the kernel is bound by the aggregate
// L2 -> SM throughput, see DESIGN.md.  The single-CTA form stays.)
template <int BN, int KS, bool RES>
__global__ void __launch_bounds__(CV_THREADS, 1)
tc_conv3x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, ConvArgs g) {
  constexpr int STAGES = conv_stages<BN>(), STAGE = conv_stage_bytes<BN>(), B_BYTES = BN * 128;
  extern __shared__ uint8_t cv_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)cv_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + STAGES * STAGE);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const int Hc = (g.H + g.stride - 1) / g.stride, Wc = (g.W + g.stride - 1) / g.stride;  // conv output size
  const int tiles_x = (Wc + CV_TW - 1) / CV_TW, tiles_y = (Hc + CV_TH - 1) / CV_TH;
  int t = blockIdx.x;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y; t /= tiles_y;
  const int b = t;
  const int n0 = blockIdx.y * BN;
  const int x0 = tx * CV_TW, y0 = ty * CV_TH;
  constexpr int ntaps = KS * KS, kpad = KS / 2;

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmW);
    for (int s = 0; s < STAGES; s++) { tc::mbar_init(full + s, 1); tc::mbar_init(empty + s, 1); }
    tc::mbar_init(tmem_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 4 * BN);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int chunks = g.Cin / CV_CK, steps = ntaps * chunks;

  if (warp == 0) {
    if (lane == 0) {
      int tap = 0, ck = 0;
      for (int it = 0; it < steps; it++, ck++) {
        if (ck == chunks) { ck = 0; tap++; }
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        const int dy = tap / KS - kpad, dx = tap % KS - kpad;
        tc::mbar_wait(empty + s, ph ^ 1);
        tc::mbar_expect_tx(full + s, STAGE);
        uint8_t* st = smem + s * STAGE;
#pragma unroll
        for (int p = 0; p < NP; p++)  // activation planes: tensor dims (C, W, H, plane*B + b); stride-2 via element strides
          tc::tma_load_4d(st + p * CV_A_BYTES, &tmA, full + s, ck * CV_CK, x0 * g.stride + dx, y0 * g.stride + dy, p * g.B + b);
#pragma unroll
        for (int p = 0; p < NP; p++)  // weight planes: rows (plane*9 + tap)*Cout + n, cols Cin
          tc::tma_load_2d(st + NP * CV_A_BYTES + p * B_BYTES, &tmW, full + s, ck * CV_CK, (p * ntaps + tap) * g.Cout + n0);
      }
    }
  } else if (warp == 1) {
    {
      const bool leader = tc::elect_one();
      constexpr uint32_t idesc = tc::make_idesc(tc::FMT_F16, 128, BN), idesc2 = tc::make_idesc(tc::FMT_F16, 128, 2 * BN);
      int tap = 0, ck = 0;
      for (int it = 0; it < steps; it++, ck++) {
        if (ck == chunks) { ck = 0; tap++; }
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        tc::mbar_wait(full + s, ph);
        tc::fence_after_sync();
        const uint32_t a0 = tc::smem_u32(smem + s * STAGE), b0 = a0 + NP * CV_A_BYTES;
        // The two weight planes are contiguous in N ([b_hi | b_lo] rows): a_hi x [b_hi | b_lo] -> [main | cross] is ONE MMA of
        // N = 2 BN, a_lo x b_hi -> cross the second.  The [main | cross] accumulator set alternates step by step: shorter
        // truncating-add chains.
        const uint32_t d_set = tmem_base + (it & 1) * 2 * BN;
#pragma unroll
        for (int k = 0; k < CV_CK / 16; k++) {
          const uint64_t a_hi = tc::make_smem_desc_sw128(a0 + k * 32), a_lo = tc::make_smem_desc_sw128(a0 + CV_A_BYTES + k * 32);
          const uint64_t b_hi = tc::make_smem_desc_sw128(b0 + k * 32);
          if (leader) {
            tc::mma_f16(d_set, a_hi, b_hi, idesc2, (it >= 2 || k) ? 1u : 0u);
            tc::mma_f16(d_set + BN, a_lo, b_hi, idesc, 1u);
          }
        }
        if (leader) tc::mma_commit(empty + s);
        __syncwarp();
      }
      if (leader) tc::mma_commit(tmem_full);
      __syncwarp();
    }
  } else {
    const int q = warp % 4, chalf = (warp - 2) / 4;   // TMEM sub-partition; columns [chalf * BN / 2, +BN / 2) of the slice
    tc::mbar_wait(tmem_full, 0);
    tc::fence_after_sync();
    const int m = q * 32 + lane;              // pixel index in the tile: row m/16, col m%16
    const int py = y0 + m / CV_TW, px = x0 + m % CV_TW;
    const int Ho = g.pool ? Hc / 2 : Hc, Wo = g.pool ? Wc / 2 : Wc;
    const bool writer = g.pool ? ((lane & 1) == 0 && (lane & 16) == 0) : true;
    const int oy = g.pool ? py / 2 : py, ox = g.pool ? px / 2 : px;
    const bool in_img = (py < Hc) && (px < Wc);
    const size_t plane_stride = (size_t)g.B * Ho * Wo * g.Cout;
    const size_t opix = (((size_t)b * Ho + oy) * Wo + ox) * g.Cout + n0;
#pragma unroll 1
    for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 32) {
      float v[32], t[32];
      const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + c0;
      if (steps >= 2) {   // both [main | cross] sets were written (set 1 = odd k-steps)
        tc::tmem_ld_acc32(lane_base, BN, 2 * BN, PLANE_LO_INV, v);
      } else {
        tc::tmem_ld32(lane_base, v);
        tc::tmem_ld32(lane_base + BN, t);
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = fmaf(t[j], PLANE_LO_INV, v[j]);
      }
      if (RES && in_img) {  // residual branch of a BasicBlock (added before the activation)
        const plane_t* r0 = g.res_planes + opix + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 a = *reinterpret_cast<const uint4*>(r0 + j), bq = *reinterpret_cast<const uint4*>(r0 + plane_stride + j);
          const plane_t *pa = reinterpret_cast<const plane_t*>(&a), *pb = reinterpret_cast<const plane_t*>(&bq);
#pragma unroll
          for (int e = 0; e < 8; e++) t[j + e] = merge2(pa[e], pb[e]);
        }
      } else if (RES) {
#pragma unroll
        for (int j = 0; j < 32; j++) t[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 32; j++) {
        float x = v[j] + (g.bias ? g.bias[n0 + c0 + j] : 0.f);
        if (RES) x += t[j];
        if (g.relu == 1) x = fmaxf(x, 0.f);
        else if (g.relu == 2) x = x > 0.f ? x : 0.01f * x;
        if (g.pool) {  // 2x2 window = lanes {l, l^1, l^16}: all inside this warp (2 image rows x 16 cols)
          x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 1));
          x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 16));
        }
        v[j] = x;
      }
      if (writer && in_img) {
        if (g.out_fp32) {
          float4* o = reinterpret_cast<float4*>(g.out_f32 + opix + c0);
#pragma unroll
          for (int j = 0; j < 8; j++) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
          __align__(16) plane_t p0[32], p1[32];
#pragma unroll
          for (int j = 0; j < 32; j++) split2(v[j], p0[j], p1[j]);
          uint4* o0 = reinterpret_cast<uint4*>(g.out_planes + opix + c0);
          uint4* o1 = reinterpret_cast<uint4*>(g.out_planes + plane_stride + opix + c0);
#pragma unroll
          for (int j = 0; j < 4; j++) {
            o0[j] = reinterpret_cast<const uint4*>(p0)[j];
            o1[j] = reinterpret_cast<const uint4*>(p1)[j];
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 4 * BN);
}

// ---- Cin = Cout = 64 specialisation (conv1b / conv2a / conv2b: 60 % of SuperPoint's FLOPs) ------------------
// The generic kernel re-fetches the activation tile for each of the nine taps and is L2->SM bandwidth bound at
// N = 64 (648 KB per CTA).  Here the pixel tile is 16 rows x 8 cols and, per plane, three copies of the halo
// patch are loaded ONCE, pre-shifted by dx = -1,0,+1 (box 64ch x 8 cols x 18 rows = 18 KB each).  Tap (dy,dx)
// is then the plain K-major SWIZZLE_128B operand starting dy*8 rows = dy*1024 B into copy dx: whole 1024-byte
// swizzle atoms, no descriptor tricks.  Activation traffic drops 432 KB -> 162 KB per CTA; weights stream
// through a 2-stage ring (24 KB per tap).
constexpr int C64_TH = 16, C64_TW = 8;
constexpr int C64_COPY = 18 * 8 * 128;            // one (plane, dx) halo copy: 18 rows x 8 px x 128 B
constexpr int C64_A_BYTES = NP * 3 * C64_COPY;    // NP planes x 3 dx
constexpr int C64_B_STAGE = NP * 64 * 128;        // NP weight planes of one tap
constexpr int C64_B_STAGES = 4;
constexpr size_t C64_SMEM = C64_A_BYTES + C64_B_STAGES * C64_B_STAGE + 1024 + 256 + 2 * 240 * sizeof(float) /*fused: image patches*/;
constexpr int C64_FUSE_PROD = 192;   // conv1a producer threads (6 warps: 512 threads per CTA keep 128 registers per thread)
constexpr int C64_HALO_PX = 18 * 10;                       // conv1a outputs one tile needs: rows y0-1..y0+16, cols x0-1..x0+8
constexpr int C64_STG_BYTES = NP * C64_HALO_PX * 128;      // staging of the split conv1a outputs (fused kernel only)
constexpr size_t C64_FUSE_SMEM = C64_SMEM + C64_STG_BYTES;
// warps 0 (TMA), 1 (MMA), 2..9 (epilogue: two warps per TMEM sub-partition, 32 of the 64 output channels each -- with four warps
// the epilogue of a tile took as long as the tile itself, 5 us, and bounded the kernel), 10..17 (FUSE: conv1a producers)
constexpr int C64_EPI_WARPS = 8;
constexpr int C64_THREADS = 64 + 32 * C64_EPI_WARPS;
constexpr int C64_FUSE_THREADS = C64_THREADS + C64_FUSE_PROD;

// Persistent: one CTA per SM walks the tile list.  The three dx-copy slots, the weight ring and two TMEM
// accumulator sets (2 x 4 x 64 = 512 columns) are all recycled through mbarriers, so the next tile's loads and
// MMAs run while the epilogue warps drain the previous tile.
// FUSE: the input is the 1-channel image; the producer warps (10..15) evaluate conv1a (3x3, Cin = 1, bias, ReLU; superpoint.py:152) on the halo
// patch of every tile, split the result into the three bf16 planes and write the three dx-shifted SWIZZLE_128B copies
// themselves (the layout TMA would have produced) -- conv1a's 118 MB / image of plane traffic never touches HBM.
template <bool FUSE>
__global__ void __launch_bounds__(FUSE ? C64_FUSE_THREADS : C64_THREADS, 1)
tc_conv3x3_c64_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, ConvArgs g, int total_tiles) {
  constexpr int BN = 64;
  extern __shared__ uint8_t cv_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)cv_smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                       // [dx][plane][18*8 rows][128 B]   (plane: hi, lo)
  uint8_t* sB = smem + C64_A_BYTES;         // [stage][plane][64 rows][128 B]
  uint64_t* a_full = (uint64_t*)(sB + C64_B_STAGES * C64_B_STAGE);  // [3] one per dx slot
  uint64_t* a_empty = a_full + 3;
  uint64_t* b_full = a_empty + 3;           // [stages]
  uint64_t* b_empty = b_full + C64_B_STAGES;
  uint64_t* tmem_full = b_empty + C64_B_STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;           // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int tiles_x = g.W / C64_TW, tiles_y = (g.H + C64_TH - 1) / C64_TH;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmW);
    for (int i = 0; i < 3; i++) { tc::mbar_init(a_full + i, FUSE ? C64_FUSE_PROD : 1); tc::mbar_init(a_empty + i, 1); }
    for (int s = 0; s < C64_B_STAGES; s++) { tc::mbar_init(b_full + s, 1); tc::mbar_init(b_empty + s, 1); }
    for (int a = 0; a < 2; a++) { tc::mbar_init(tmem_full + a, 1); tc::mbar_init(tmem_empty + a, 32 * C64_EPI_WARPS); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // taps are visited dx-major so that the MMAs of column dx can start as soon as its three copies landed
  if (warp == 0) {
    if (lane == 0) {
      int i = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, i++) {
        const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
        const int x0 = tx * C64_TW, y0 = ty * C64_TH;
        for (int it = 0; it < 9; it++) {
          const int dxi = it / 3, dyi = it % 3, tap = dyi * 3 + dxi;
          if (!FUSE && dyi == 0) {
            tc::mbar_wait(a_empty + dxi, (i & 1) ^ 1);
            tc::mbar_expect_tx(a_full + dxi, NP * C64_COPY);
#pragma unroll
            for (int p = 0; p < NP; p++)
              tc::tma_load_4d(sA + (dxi * NP + p) * C64_COPY, &tmA, a_full + dxi, 0, x0 + dxi - 1, y0 - 1, p * g.B + b);
          }
          const int c = i * 9 + it, s = c % C64_B_STAGES, ph = (c / C64_B_STAGES) & 1;
          tc::mbar_wait(b_empty + s, ph ^ 1);
          tc::mbar_expect_tx(b_full + s, C64_B_STAGE);
#pragma unroll
          for (int p = 0; p < NP; p++)
            tc::tma_load_2d(sB + s * C64_B_STAGE + p * 64 * 128, &tmW, b_full + s, 0, (p * 9 + tap) * g.Cout);
        }
      }
    }
  } else if (warp == 1) {
    {
      const bool leader = tc::elect_one();
      constexpr uint32_t idesc = tc::make_idesc(tc::FMT_F16, 128, BN), idesc2 = tc::make_idesc(tc::FMT_F16, 128, 2 * BN);
      int i = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, i++) {
        const int acc = i & 1;
        tc::mbar_wait(tmem_empty + acc, ((i >> 1) & 1) ^ 1);  // epilogue drained this accumulator set
        tc::fence_after_sync();
        const uint32_t d_base = tmem_base + acc * 256;
        for (int it = 0; it < 9; it++) {
          const int dxi = it / 3, dyi = it % 3;
          const int c = i * 9 + it, s = c % C64_B_STAGES, ph = (c / C64_B_STAGES) & 1;
          if (dyi == 0) tc::mbar_wait(a_full + dxi, i & 1);
          tc::mbar_wait(b_full + s, ph);
          tc::fence_after_sync();
          const uint32_t a0 = tc::smem_u32(sA + dxi * NP * C64_COPY) + dyi * 1024, b0 = tc::smem_u32(sB + s * C64_B_STAGE);
          const uint32_t d_set = d_base + (it & 1) * 2 * BN;     // [main | cross] sets alternate tap by tap
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const uint64_t a_hi = tc::make_smem_desc_sw128(a0 + k * 32), a_lo = tc::make_smem_desc_sw128(a0 + C64_COPY + k * 32);
            const uint64_t b_hi = tc::make_smem_desc_sw128(b0 + k * 32);
            if (leader) {  // two MMAs for the three partial products ([b_hi | b_lo] contiguous in N), see tc_conv3x3_kernel
              tc::mma_f16(d_set, a_hi, b_hi, idesc2, (it >= 2 || k) ? 1u : 0u);
              tc::mma_f16(d_set + BN, a_lo, b_hi, idesc, 1u);
            }
          }
          if (leader) {
            tc::mma_commit(b_empty + s);
            if (dyi == 2) tc::mma_commit(a_empty + dxi);  // the three taps of this dx slot are done
          }
          __syncwarp();
        }
        if (leader) tc::mma_commit(tmem_full + acc);
        __syncwarp();
      }
    }
  } else if (FUSE && warp >= 2 + C64_EPI_WARPS) {
    const int t = threadIdx.x - C64_THREADS;   // 0..C64_FUSE_PROD-1
    const int chunk = t % 8;                  // output channels [8 chunk, +8) = one 16-byte unit of a pixel's 128-byte row
    float w[9][8], bv[8];
#pragma unroll
    for (int tp = 0; tp < 9; tp++)
#pragma unroll
      for (int k = 0; k < 8; k++) w[tp][k] = g.w1a[tp * 64 + chunk * 8 + k];
#pragma unroll
    for (int k = 0; k < 8; k++) bv[k] = g.b1a[chunk * 8 + k];
    float* s_img = (float*)((uint8_t*)tmem_slot + 64);   // [2][20 rows][12 cols], double-buffered by tile parity
    uint8_t* stg = (uint8_t*)s_img + 2 * 240 * sizeof(float);   // [plane][180 halo pixels][128 B] (16-byte aligned: see C64_SMEM)
    int i = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, i++) {
      const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
      const int x0 = tx * C64_TW, y0 = ty * C64_TH;
      float* im = s_img + (i & 1) * 240;
      const float* src = g.img + (size_t)b * g.H * g.W;
      for (int idx = t; idx < 240; idx += C64_FUSE_PROD) {   // image rows y0-2 .. y0+17, cols x0-2 .. x0+9, zeros outside (conv1a padding)
        const int gy = y0 - 2 + idx / 12, gx = x0 - 2 + idx % 12;
        im[idx] = (gy >= 0 && gy < g.H && gx >= 0 && gx < g.W) ? src[(size_t)gy * g.W + gx] : 0.f;
      }
      asm volatile("bar.sync 2, %0;" ::"n"(C64_FUSE_PROD) : "memory");
      // conv1a ONCE per halo pixel (18 x 10), split, into a linear staging buffer [plane][pixel][8 units of 16 B].  (Round 1
      // evaluated it once per dx-shifted copy: 3 x 144 pixel evaluations per tile, which became the kernel's bottleneck -- issue
      // slots of the producer warps -- once the tensor work was halved.)  Runs ahead of the a_empty waits: overlaps the MMAs of
      // the previous tile.
#pragma unroll 1
      for (int k = 0; k < (C64_HALO_PX + C64_FUSE_PROD / 8 - 1) / (C64_FUSE_PROD / 8); k++) {
        const int pidx = t / 8 + (C64_FUSE_PROD / 8) * k;   // halo pixel: row pidx / 10, column pidx % 10
        if (pidx >= C64_HALO_PX) break;
        const int hy = pidx / 10, hx = pidx % 10;
        const int gy = y0 - 1 + hy, gx = x0 - 1 + hx;       // conv1a output position
        float a[8];
#pragma unroll
        for (int c = 0; c < 8; c++) a[c] = 0.f;
        const bool inside = gy >= 0 && gy < g.H && gx >= 0 && gx < g.W;   // conv1b's zero padding: outside -> 0
        if (inside) {
#pragma unroll
          for (int dy = 0; dy < 3; dy++)
#pragma unroll
            for (int dx = 0; dx < 3; dx++) {
              const float v = im[(hy + dy) * 12 + hx + dx];
#pragma unroll
              for (int c = 0; c < 8; c++) a[c] = fmaf(v, w[dy * 3 + dx][c], a[c]);
            }
#pragma unroll
          for (int c = 0; c < 8; c++) a[c] = fmaxf(a[c] + bv[c], 0.f);
        }
        __align__(16) plane_t p0[8], p1[8];
#pragma unroll
        for (int c = 0; c < 8; c++) split2(a[c], p0[c], p1[c]);
        *reinterpret_cast<uint4*>(stg + pidx * 128 + chunk * 16) = *reinterpret_cast<const uint4*>(p0);
        *reinterpret_cast<uint4*>(stg + C64_HALO_PX * 128 + pidx * 128 + chunk * 16) = *reinterpret_cast<const uint4*>(p1);
      }
      asm volatile("bar.sync 2, %0;" ::"n"(C64_FUSE_PROD) : "memory");
      // three dx-shifted SWIZZLE_128B copies (the layout TMA would have produced): copy dx row (hy, cx) = halo pixel (hy, cx + dx)
      for (int dxi = 0; dxi < 3; dxi++) {
        tc::mbar_wait(a_empty + dxi, (i & 1) ^ 1);   // the MMAs of the previous tile have read this dx slot
        uint8_t* copy = sA + dxi * NP * C64_COPY;
#pragma unroll 1
        for (int k = 0; k < (144 + C64_FUSE_PROD / 8 - 1) / (C64_FUSE_PROD / 8); k++) {
          const int pidx = t / 8 + (C64_FUSE_PROD / 8) * k;   // row of the copy: halo row pidx / 8, column pidx % 8
          if (pidx >= 144) break;
          const int src_px = (pidx / 8) * 10 + (pidx % 8) + dxi;
          const int off = pidx * 128 + ((chunk ^ (pidx & 7)) * 16);   // SWIZZLE_128B: 16-byte unit c of row r sits at c ^ (r & 7)
          *reinterpret_cast<uint4*>(copy + off) = *reinterpret_cast<const uint4*>(stg + src_px * 128 + chunk * 16);
          *reinterpret_cast<uint4*>(copy + C64_COPY + off) = *reinterpret_cast<const uint4*>(stg + C64_HALO_PX * 128 + src_px * 128 + chunk * 16);
        }
        tc::fence_proxy_async();   // generic-proxy writes -> visible to the tensor core
        tc::mbar_arrive(a_full + dxi);
      }
      asm volatile("bar.sync 2, %0;" ::"n"(C64_FUSE_PROD) : "memory");   // staging is rewritten for the next tile
    }
  } else {
    const int q = warp % 4, chalf = (warp - 2) / 4;   // TMEM sub-partition; output channels [32 chalf, +32)
    const int m = q * 32 + lane;              // pixel index in the tile: row m/8, col m%8
    const int Ho = g.pool ? g.H / 2 : g.H, Wo = g.pool ? g.W / 2 : g.W;
    const bool writer = g.pool ? ((lane & 1) == 0 && (lane & 8) == 0) : true;
    const size_t plane_stride = (size_t)g.B * Ho * Wo * g.Cout;
    int i = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, i++) {
      const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
      const int x0 = tx * C64_TW, y0 = ty * C64_TH;
      const int acc = i & 1;
      tc::mbar_wait(tmem_full + acc, (i >> 1) & 1);
      tc::fence_after_sync();
      const int py = y0 + m / C64_TW, px = x0 + m % C64_TW;
      const int oy = g.pool ? py / 2 : py, ox = g.pool ? px / 2 : px;
      const bool in_img = (py < g.H) && (px < g.W);
      const size_t opix = (((size_t)b * Ho + oy) * Wo + ox) * g.Cout;
      {
        const int c0 = chalf * 32;
        float v[32];
        const uint32_t lane_base = tmem_base + acc * 256 + ((uint32_t)(q * 32) << 16) + c0;
        tc::tmem_ld_acc32(lane_base, BN, 2 * BN, PLANE_LO_INV, v);   // (main0 + main1) + (cross0 + cross1) 2^-11
        {  // this warp's only TMEM read of the accumulator set: hand it back to the MMA warp
          tc::fence_before_sync();
          tc::mbar_arrive(tmem_empty + acc);
        }
#pragma unroll
        for (int j = 0; j < 32; j++) {
          float x = v[j] + g.bias[c0 + j];
          if (g.relu) x = fmaxf(x, 0.f);
          if (g.pool) {  // 2x2 window = lanes {l, l^1, l^8}: 4 image rows x 8 cols per warp
            x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 1));
            x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 8));
          }
          v[j] = x;
        }
        if (writer && in_img) {
          if (g.out_fp32) {
            float4* o = reinterpret_cast<float4*>(g.out_f32 + opix + c0);
#pragma unroll
            for (int j = 0; j < 8; j++) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            __align__(16) plane_t p0[32], p1[32];
#pragma unroll
            for (int j = 0; j < 32; j++) split2(v[j], p0[j], p1[j]);
            uint4* o0 = reinterpret_cast<uint4*>(g.out_planes + opix + c0);
            uint4* o1 = reinterpret_cast<uint4*>(g.out_planes + plane_stride + opix + c0);
#pragma unroll
            for (int j = 0; j < 4; j++) {
              o0[j] = reinterpret_cast<const uint4*>(p0)[j];
              o1[j] = reinterpret_cast<const uint4*>(p1)[j];
            }
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

#include "tc_conv_pair.cuh"   // the same conv on a CTA pair (cta_group::2): dense halo patch, resident weights
#include "tc_conv_pair_halo.cuh"   // any channel count on a CTA pair: dense halo patch, weight stages split between the CTAs

// ---- 3x3 stride-1 convs with Cin = 64 k, Cout = 64 / 128 n: the halo-copy scheme of the c64 kernel for any channel count ----
// The generic kernel above fetches the activation tile once per tap: 9 x 32 KB of activations + 9 x 32 KB of weights per
// 64-channel chunk and 128 pixels, and is bound by the chip's aggregate L2 -> SM throughput (ncu round 2: 7.4 TB/s, tensor pipe
// 23-35 %).  Here the 16 x 8 pixel tile keeps, per 64-channel chunk, three dx-shifted halo copies (2 planes x 18 KB each) that
// serve three taps each: 108 KB of activations instead of 288 KB per chunk.  The three dx slots ROLL: a slot is refilled with
// the next chunk (or the next tile) as soon as its three taps have been issued, by a dedicated TMA warp, so activation loads
// run six taps ahead of their use; weights stream through their own ring from a second TMA warp.  Persistent CTAs walk the
// (tile, Cout slice) list; the accumulator (two alternating [main | cross] sets, see above) is single-buffered at BN = 128
// (512 TMEM columns), so only the loads -- not the MMAs -- of the next tile overlap the epilogue.
constexpr int HL_EPI_WARPS = 8;
constexpr int HL_THREADS = (3 + HL_EPI_WARPS) * 32;   // warp 0: weight TMA, 1: MMA, 2: activation TMA, 3..10: epilogue (two warps per
                                                       // TMEM sub-partition, half of the Cout slice each: the accumulator is single-buffered, so the
                                                       // epilogue is exposed time)
template <int BN> constexpr int hl_b_stage() { return NP * BN * 128; }
template <int BN> constexpr int hl_b_stages() { return BN == 128 ? 3 : 6; }
template <int BN> constexpr size_t hl_smem_bytes() { return (size_t)C64_A_BYTES + (size_t)hl_b_stages<BN>() * hl_b_stage<BN>() + 1024 + 256; }

template <int BN, bool RES>
__global__ void __launch_bounds__(HL_THREADS, 1)
tc_conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, ConvArgs g, int total_items) {
  constexpr int STAGES = hl_b_stages<BN>(), B_STAGE = hl_b_stage<BN>();
  extern __shared__ uint8_t cv_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)cv_smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                       // [dx][plane][18*8 rows][128 B]
  uint8_t* sB = smem + C64_A_BYTES;         // [stage][plane][BN rows][128 B]
  uint64_t* a_full = (uint64_t*)(sB + STAGES * B_STAGE);  // [3] one per dx slot
  uint64_t* a_empty = a_full + 3;
  uint64_t* b_full = a_empty + 3;           // [stages]
  uint64_t* b_empty = b_full + STAGES;
  uint64_t* tmem_full = b_empty + STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 1);

  const int tiles_x = g.W / C64_TW, tiles_y = (g.H + C64_TH - 1) / C64_TH;
  const int n_tiles = g.Cout / BN, chunks = g.Cin / CV_CK;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmW);
    for (int i = 0; i < 3; i++) { tc::mbar_init(a_full + i, 1); tc::mbar_init(a_empty + i, 1); }
    for (int s = 0; s < STAGES; s++) { tc::mbar_init(b_full + s, 1); tc::mbar_init(b_empty + s, 1); }
    tc::mbar_init(tmem_full, 1); tc::mbar_init(tmem_empty, 32 * HL_EPI_WARPS);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 4 * BN);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int item, int& b, int& x0, int& y0, int& n0) {
    const int tile = item / n_tiles;             // Cout slices of one tile are neighbours: its activations stay in L2
    n0 = (item % n_tiles) * BN;
    x0 = (tile % tiles_x) * C64_TW; y0 = ((tile / tiles_x) % tiles_y) * C64_TH; b = tile / (tiles_x * tiles_y);
  };

  if (warp == 2) {            // activation TMA: slot dx of (item, chunk) -- use number u = i * chunks + ck of every slot
    if (lane == 0) {
      int i = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, i++) {
        int b, x0, y0, n0;
        decode(item, b, x0, y0, n0);
        for (int ck = 0; ck < chunks; ck++) {
          const int u = i * chunks + ck;
          for (int dxi = 0; dxi < 3; dxi++) {
            tc::mbar_wait(a_empty + dxi, (u & 1) ^ 1);
            tc::mbar_expect_tx(a_full + dxi, NP * C64_COPY);
#pragma unroll
            for (int p = 0; p < NP; p++)
              tc::tma_load_4d(sA + (dxi * NP + p) * C64_COPY, &tmA, a_full + dxi, ck * CV_CK, x0 + dxi - 1, y0 - 1, p * g.B + b);
          }
        }
      }
    }
  } else if (warp == 0) {     // weight TMA: one stage per (chunk, dx, dy) in the MMA's order
    if (lane == 0) {
      int c = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        int b, x0, y0, n0;
        decode(item, b, x0, y0, n0);
        for (int ck = 0; ck < chunks; ck++)
          for (int it = 0; it < 9; it++, c++) {
            const int dxi = it / 3, dyi = it % 3, tap = dyi * 3 + dxi;
            const int s = c % STAGES, ph = (c / STAGES) & 1;
            tc::mbar_wait(b_empty + s, ph ^ 1);
            tc::mbar_expect_tx(b_full + s, B_STAGE);
#pragma unroll
            for (int p = 0; p < NP; p++)
              tc::tma_load_2d(sB + s * B_STAGE + p * BN * 128, &tmW, b_full + s, ck * CV_CK, (p * 9 + tap) * g.Cout + n0);
          }
      }
    }
  } else if (warp == 1) {
    const bool leader = tc::elect_one();
    constexpr uint32_t idesc = tc::make_idesc(tc::FMT_F16, 128, BN), idesc2 = tc::make_idesc(tc::FMT_F16, 128, 2 * BN);
    int i = 0, c = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, i++) {
      tc::mbar_wait(tmem_empty, (i & 1) ^ 1);   // the epilogue has drained the previous item's accumulators
      tc::fence_after_sync();
      int step = 0;
      for (int ck = 0; ck < chunks; ck++) {
        const int u = i * chunks + ck;
        for (int it = 0; it < 9; it++, c++, step++) {
          const int dxi = it / 3, dyi = it % 3;
          const int s = c % STAGES, ph = (c / STAGES) & 1;
          if (dyi == 0) tc::mbar_wait(a_full + dxi, u & 1);
          tc::mbar_wait(b_full + s, ph);
          tc::fence_after_sync();
          const uint32_t a0 = tc::smem_u32(sA + dxi * NP * C64_COPY) + dyi * 1024, b0 = tc::smem_u32(sB + s * B_STAGE);
          const uint32_t d_set = tmem_base + (step & 1) * 2 * BN;     // [main | cross] sets alternate step by step
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const uint64_t a_hi = tc::make_smem_desc_sw128(a0 + k * 32), a_lo = tc::make_smem_desc_sw128(a0 + C64_COPY + k * 32);
            const uint64_t b_hi = tc::make_smem_desc_sw128(b0 + k * 32);
            if (leader) {
              tc::mma_f16(d_set, a_hi, b_hi, idesc2, (step >= 2 || k) ? 1u : 0u);
              tc::mma_f16(d_set + BN, a_lo, b_hi, idesc, 1u);
            }
          }
          if (leader) {
            tc::mma_commit(b_empty + s);
            if (dyi == 2) tc::mma_commit(a_empty + dxi);  // the three taps of this dx slot are done: it may take the next chunk
          }
          __syncwarp();
        }
      }
      if (leader) tc::mma_commit(tmem_full);
      __syncwarp();
    }
  } else {
    const int q = warp % 4, chalf = (warp - 3) / 4;   // TMEM sub-partition; columns [chalf * BN / 2, +BN / 2) of the slice
    const int m = q * 32 + lane;              // pixel index in the tile: row m/8, col m%8
    const int Ho = g.pool ? g.H / 2 : g.H, Wo = g.pool ? g.W / 2 : g.W;
    const bool writer = g.pool ? ((lane & 1) == 0 && (lane & 8) == 0) : true;
    const size_t plane_stride = (size_t)g.B * Ho * Wo * g.Cout;
    int i = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, i++) {
      int b, x0, y0, n0;
      decode(item, b, x0, y0, n0);
      tc::mbar_wait(tmem_full, i & 1);
      tc::fence_after_sync();
      const int py = y0 + m / C64_TW, px = x0 + m % C64_TW;
      const int oy = g.pool ? py / 2 : py, ox = g.pool ? px / 2 : px;
      const bool in_img = (py < g.H) && (px < g.W);
      const size_t opix = (((size_t)b * Ho + oy) * Wo + ox) * g.Cout + n0;
#pragma unroll 1
      for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 32) {
        float v[32], t[32];
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + c0;
        tc::tmem_ld_acc32(lane_base, BN, 2 * BN, PLANE_LO_INV, v);   // (main0 + main1) + (cross0 + cross1) 2^-11
        if (c0 + 32 >= (chalf + 1) * (BN / 2)) {  // this warp's last TMEM read of the item: hand the accumulators back to the MMA warp
          tc::fence_before_sync();
          tc::mbar_arrive(tmem_empty);
        }
        if (RES && in_img) {  // residual branch of a BasicBlock (added before the activation)
          const plane_t* r0 = g.res_planes + opix + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 a = *reinterpret_cast<const uint4*>(r0 + j), bq = *reinterpret_cast<const uint4*>(r0 + plane_stride + j);
            const plane_t *pa = reinterpret_cast<const plane_t*>(&a), *pb = reinterpret_cast<const plane_t*>(&bq);
#pragma unroll
            for (int e = 0; e < 8; e++) t[j + e] = merge2(pa[e], pb[e]);
          }
        } else if (RES) {
#pragma unroll
          for (int j = 0; j < 32; j++) t[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; j++) {
          float x = v[j] + (g.bias ? g.bias[n0 + c0 + j] : 0.f);
          if (RES) x += t[j];
          if (g.relu == 1) x = fmaxf(x, 0.f);
          else if (g.relu == 2) x = x > 0.f ? x : 0.01f * x;
          if (g.pool) {  // 2x2 window = lanes {l, l^1, l^8}: 4 image rows x 8 cols per warp
            x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 1));
            x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 8));
          }
          v[j] = x;
        }
        if (writer && in_img) {
          if (g.out_fp32) {
            float4* o = reinterpret_cast<float4*>(g.out_f32 + opix + c0);
#pragma unroll
            for (int j = 0; j < 8; j++) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            __align__(16) plane_t p0[32], p1[32];
#pragma unroll
            for (int j = 0; j < 32; j++) split2(v[j], p0[j], p1[j]);
            uint4* o0 = reinterpret_cast<uint4*>(g.out_planes + opix + c0);
            uint4* o1 = reinterpret_cast<uint4*>(g.out_planes + plane_stride + opix + c0);
#pragma unroll
            for (int j = 0; j < 4; j++) {
              o0[j] = reinterpret_cast<const uint4*>(p0)[j];
              o1[j] = reinterpret_cast<const uint4*>(p1)[j];
            }
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 4 * BN);
}

// fp32 NHWC -> two fp16 planes (and back): interop with the CUDA-core path and the unit tests
__global__ void split_planes_kernel(const float* __restrict__ in, plane_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  plane_t a, b;
  split2(in[i], a, b);
  out[i] = a; out[n + i] = b;
}
__global__ void merge_planes_kernel(const plane_t* __restrict__ in, float* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = merge2(in[i], in[n + i]);
}

int make_map_act(CUtensorMap* map, const void* base, int BP, int H, int W, int C, int box_w = CV_TW, int box_h = CV_TH, int stride = 1) {
  PFN_encodeTiled fn = tc_get_encode_fn();
  if (!fn) { imw_set_error("cuTensorMapEncodeTiled not available"); return IMW_ERR_CUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)BP};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  // with element strides the box extent is given in INPUT elements: N loaded elements <=> boxDim = N * stride
  cuuint32_t box[4] = {CV_CK, (cuuint32_t)(box_w * stride), (cuuint32_t)(box_h * stride), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { imw_set_error("cuTensorMapEncodeTiled(act) failed: %d", (int)r); return IMW_ERR_CUDA; }
  return IMW_OK;
}
int make_map_wgt(CUtensorMap* map, const void* base, int rows, int Cin, int BN) {
  PFN_encodeTiled fn = tc_get_encode_fn();
  if (!fn) { imw_set_error("cuTensorMapEncodeTiled not available"); return IMW_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
  cuuint32_t box[2] = {CV_CK, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { imw_set_error("cuTensorMapEncodeTiled(wgt) failed: %d", (int)r); return IMW_ERR_CUDA; }
  return IMW_OK;
}

template <int BN, int KS, bool RES>
int launch_conv_t(const CUtensorMap& tmA, const CUtensorMap& tmW, const ConvArgs& g, cudaStream_t st) {
  constexpr size_t smem = conv_smem_bytes<BN>();
  IMW_SMEM_ATTR_ONCE((tc_conv3x3_kernel<BN, KS, RES>), smem);
  const int Hc = ceil_div(g.H, g.stride), Wc = ceil_div(g.W, g.stride);
  dim3 grid((unsigned)(g.B * ceil_div(Hc, CV_TH) * ceil_div(Wc, CV_TW)), g.Cout / BN);
  tc_conv3x3_kernel<BN, KS, RES><<<grid, CV_THREADS, smem, st>>>(tmA, tmW, g);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}

template <int BN, bool RES>
int launch_conv_halo_t(const CUtensorMap& tmA, const CUtensorMap& tmW, const ConvArgs& g, cudaStream_t st) {
  constexpr size_t smem = hl_smem_bytes<BN>();
  IMW_SMEM_ATTR_ONCE((tc_conv3x3_halo_kernel<BN, RES>), smem);
  const int total = g.B * ceil_div(g.H, C64_TH) * (g.W / C64_TW) * (g.Cout / BN);
  const int num_sms = imw_num_sms();
  tc_conv3x3_halo_kernel<BN, RES><<<dim3((unsigned)(total < num_sms ? total : num_sms)), HL_THREADS, smem, st>>>(tmA, tmW, g, total);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}
// 3x3, stride 1, W % 8 == 0: halo-copy kernel (activation map with the 8 x 18 halo box)
int launch_conv_halo(const void* in_planes, const void* w_planes, const ConvArgs& g, cudaStream_t st) {
  CUtensorMap tmA, tmW;
  const int BN = (g.Cout % 128 == 0) ? 128 : 64;
  if (int e = make_map_act(&tmA, in_planes, NP * g.B, g.H, g.W, g.Cin, C64_TW, C64_TH + 2)) return e;
  if (int e = make_map_wgt(&tmW, w_planes, NP * 9 * g.Cout, g.Cin, BN)) return e;
  if (g.res_planes) return BN == 128 ? launch_conv_halo_t<128, true>(tmA, tmW, g, st) : launch_conv_halo_t<64, true>(tmA, tmW, g, st);
  return BN == 128 ? launch_conv_halo_t<128, false>(tmA, tmW, g, st) : launch_conv_halo_t<64, false>(tmA, tmW, g, st);
}
static int g_conv_pair = 1;   // imw_debug_set_conv_pair: 1 = Cin = Cout = 64 convs on CTA pairs (default), 0 = single-CTA c64 kernel, 2 = pair kernel with swapped B halves (probe)
static int g_conv_halo = 1;   // imw_debug_set_conv_halo: 0 = generic kernel everywhere (A/B measurements, unit tests of both)

template <int BN>
int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmW, const ConvArgs& g, cudaStream_t st) {
  if (g.ksize == 1) return launch_conv_t<BN, 1, false>(tmA, tmW, g, st);   // (no 1x1 conv with a residual on the path)
  return g.res_planes ? launch_conv_t<BN, 3, true>(tmA, tmW, g, st) : launch_conv_t<BN, 3, false>(tmA, tmW, g, st);
}

}  // namespace

// CTA-pair launch of the c64 conv (fused first layer when g.img is set)
template <bool FUSE>
static int launch_conv_c64_pair(const CUtensorMap& tmA, const CUtensorMap& tmW, const ConvArgs& g, cudaStream_t st) {
  IMW_SMEM_ATTR_ONCE(tc_conv3x3_c64_pair_kernel<FUSE>, P2_SMEM);
  const int total = g.B * ceil_div(g.H, C64_TH) * (g.W / C64_TW);
  const int pairs = (total + 1) / 2, max_clusters = imw_num_sms() / 2;
  const int clusters = pairs < max_clusters ? pairs : max_clusters;
  tc_conv3x3_c64_pair_kernel<FUSE><<<dim3((unsigned)(2 * clusters)), FUSE ? P2_FUSE_THREADS : P2_PLAIN_THREADS, P2_SMEM, st>>>(tmA, tmW, g, total,
                                                                                                                        (g_conv_pair == 2) | (g_conv_pair & ~3));
  IMW_CHECK_LAUNCH_T(FUSE ? "tc_conv1ab_fused -> tc_conv3x3_c64_pair_kernel<fused conv1a>" : "tc_conv3x3 -> tc_conv3x3_c64_pair_kernel");
  return IMW_OK;
}
extern "C" int imw_debug_set_conv_pair(int mode) {
  if (mode >= 0) g_conv_pair = mode;
  return g_conv_pair;
}

template <int BN, bool RES>
static int launch_conv_halo_pair_t(const CUtensorMap& tmA, const CUtensorMap& tmW, const ConvArgs& g, cudaStream_t st) {
  constexpr size_t smem = ph_smem_bytes<BN>();
  IMW_SMEM_ATTR_ONCE((tc_conv3x3_halo_pair_kernel<BN, RES>), smem);
  const int total = g.B * ceil_div(g.H, C64_TH) * (g.W / C64_TW);
  const int items = ((total + 1) / 2) * (g.Cout / BN), max_clusters = imw_num_sms() / 2;
  const int clusters = items < max_clusters ? items : max_clusters;
  tc_conv3x3_halo_pair_kernel<BN, RES><<<dim3((unsigned)(2 * clusters)), ph_threads<RES>(), smem, st>>>(tmA, tmW, g, total);
  IMW_CHECK_LAUNCH_T(BN == 128 ? "tc_conv3x3 -> tc_conv3x3_halo_pair_kernel<128>" : "tc_conv3x3 -> tc_conv3x3_halo_pair_kernel<64>");
  return IMW_OK;
}
// 3x3, stride 1, W % 8 == 0 on CTA pairs: dense halo patch (10 x 18 box), each CTA fetches its half of the Cout slice
static int launch_conv_halo_pair(const void* in_planes, const void* w_planes, const ConvArgs& g, cudaStream_t st) {
  IMW_REQUIRE(!(g.res_planes && g.pool), "tc_conv (pair halo kernel): a residual input with a fused max-pool is not built");
  CUtensorMap tmA, tmW;
  const int BN = (g.Cout % 128 == 0) ? 128 : 64;
  if (int e = make_map_act(&tmA, in_planes, NP * g.B, g.H, g.W, g.Cin, P2_HALO_W, P2_HALO_H)) return e;
  if (int e = make_map_wgt(&tmW, w_planes, NP * 9 * g.Cout, g.Cin, BN / 2)) return e;
  if (g.res_planes) return BN == 128 ? launch_conv_halo_pair_t<128, true>(tmA, tmW, g, st) : launch_conv_halo_pair_t<64, true>(tmA, tmW, g, st);
  return BN == 128 ? launch_conv_halo_pair_t<128, false>(tmA, tmW, g, st) : launch_conv_halo_pair_t<64, false>(tmA, tmW, g, st);
}

// A/B hook (unit tests, measurements): 1 = 3x3 stride-1 convs with W % 8 == 0 run on the halo-copy kernel (default), 0 = generic kernel
extern "C" int imw_debug_set_conv_halo(int on) {
  if (on >= 0) g_conv_halo = on ? 1 : 0;
  return g_conv_halo;
}

// in_planes [NP][B][H][W][Cin] fp16, w_planes [NP][9][Cout][Cin] fp16 (split_planes.cuh), bias [Cout] fp32.
// out: planes [NP][B][Ho][Wo][Cout] fp16 or fp32 [B][Ho][Wo][Cout].
int tc_conv3x3(const void* in_planes, const void* w_planes, const float* bias, void* out, int B, int H, int W, int Cin,
               int Cout, int relu, int pool, int out_fp32, cudaStream_t st) {
  IMW_REQUIRE(Cin % CV_CK == 0 && Cout % 64 == 0, "tc_conv3x3: Cin %% 64, Cout %% 64 (got %d,%d)", Cin, Cout);
  IMW_REQUIRE(!pool || (W % 2 == 0), "tc_conv3x3: pooled conv needs even W");
  IMW_REQUIRE(!pool || (H % 2 == 0), "tc_conv3x3: pooled conv needs even H");
  CUtensorMap tmA, tmW;
  const int BN = (Cout % 128 == 0) ? 128 : 64;
  if (Cin == 64 && Cout == 64 && W % C64_TW == 0 && g_conv_pair) {  // CTA-pair kernel: dense halo patch (10 x 18 box), resident weights
    if (int e = make_map_act(&tmA, in_planes, NP * B, H, W, Cin, P2_HALO_W, P2_HALO_H)) return e;
    if (int e = make_map_wgt(&tmW, w_planes, NP * 9 * Cout, Cin, 32)) return e;
    ConvArgs g{H, W, Cin, Cout, B, relu, pool, out_fp32, bias, (plane_t*)out, (float*)out};
    return launch_conv_c64_pair<false>(tmA, tmW, g, st);
  }
  if (Cin == 64 && Cout == 64 && W % C64_TW == 0) {  // halo-copy specialisation
    if (int e = make_map_act(&tmA, in_planes, NP * B, H, W, Cin, C64_TW, C64_TH + 2)) return e;
    if (int e = make_map_wgt(&tmW, w_planes, NP * 9 * Cout, Cin, 64)) return e;
    ConvArgs g{H, W, Cin, Cout, B, relu, pool, out_fp32, bias, (plane_t*)out, (float*)out};
    IMW_SMEM_ATTR_ONCE(tc_conv3x3_c64_kernel<false>, C64_SMEM);
    const int total = B * ceil_div(H, C64_TH) * (W / C64_TW);
    const int num_sms = imw_num_sms();
    dim3 grid((unsigned)(total < num_sms ? total : num_sms), 1);
    tc_conv3x3_c64_kernel<false><<<grid, C64_THREADS, C64_SMEM, st>>>(tmA, tmW, g, total);
    IMW_CHECK_LAUNCH();
    return IMW_OK;
  }
  ConvArgs g{H, W, Cin, Cout, B, relu, pool, out_fp32, bias, (plane_t*)out, (float*)out};
  if (g_conv_halo && W % C64_TW == 0) return (g_conv_pair & 1) ? launch_conv_halo_pair(in_planes, w_planes, g, st) : launch_conv_halo(in_planes, w_planes, g, st);
  if (int e = make_map_act(&tmA, in_planes, NP * B, H, W, Cin)) return e;
  if (int e = make_map_wgt(&tmW, w_planes, NP * 9 * Cout, Cin, BN)) return e;
  return BN == 128 ? launch_conv<128>(tmA, tmW, g, st) : launch_conv<64>(tmA, tmW, g, st);
}

// SuperPoint conv1a + conv1b in one kernel: image [B][H][W] fp32 -> conv1b output planes (2x2 max-pooled when pool).
int tc_conv1ab_fused(const float* img, const float* w1a, const float* b1a, const void* w1b_planes, const float* b1b, void* out, int B,
                     int H, int W, int pool, cudaStream_t st) {
  IMW_REQUIRE(W % C64_TW == 0 && (!pool || (H % 2 == 0)), "tc_conv1ab_fused: W %% 8 == 0 (even H when pooled)");
  CUtensorMap tmW;
  if (int e = make_map_wgt(&tmW, w1b_planes, NP * 9 * 64, 64, 64)) return e;
  ConvArgs g{H, W, 64, 64, B, 1, pool, 0, b1b, (plane_t*)out, (float*)out};
  g.img = img; g.w1a = w1a; g.b1a = b1a;
  if (g_conv_pair) {
    CUtensorMap tmW32;
    if (int e = make_map_wgt(&tmW32, w1b_planes, NP * 9 * 64, 64, 32)) return e;
    return launch_conv_c64_pair<true>(tmW32, tmW32, g, st);
  }
  IMW_SMEM_ATTR_ONCE(tc_conv3x3_c64_kernel<true>, C64_FUSE_SMEM);
  const int num_sms = imw_num_sms();
  const int total = B * ceil_div(H, C64_TH) * (W / C64_TW);
  tc_conv3x3_c64_kernel<true><<<dim3((unsigned)(total < num_sms ? total : num_sms)), C64_FUSE_THREADS, C64_FUSE_SMEM, st>>>(tmW, tmW, g, total);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}

// General form used by the LoFTR backbone: 3x3 / 1x1, stride 1 / 2, optional residual, act 0 none / 1 ReLU / 2 LeakyReLU(0.01).
// w_planes [NP][k*k][Cout][Cin] fp16; bias may be NULL (bias-free convs without BatchNorm).
int tc_conv_general(const void* in_planes, const void* w_planes, const float* bias, const void* res_planes, void* out, int B,
                    int H, int W, int Cin, int Cout, int ksize, int stride, int act, int out_fp32, cudaStream_t st) {
  IMW_REQUIRE(Cin % CV_CK == 0 && Cout % 64 == 0, "tc_conv_general: Cin %% 64, Cout %% 64 (got %d,%d)", Cin, Cout);
  IMW_REQUIRE((ksize == 1 || ksize == 3) && (stride == 1 || stride == 2), "tc_conv_general: ksize 1|3, stride 1|2");
  IMW_REQUIRE(!(ksize == 1 && res_planes), "tc_conv_general: residual input is only built for 3x3 convs");
  CUtensorMap tmA, tmW;
  const int BN = (Cout % 128 == 0) ? 128 : 64;
  if (int e = make_map_act(&tmA, in_planes, NP * B, H, W, Cin, CV_TW, CV_TH, stride)) return e;
  if (int e = make_map_wgt(&tmW, w_planes, NP * ksize * ksize * Cout, Cin, BN)) return e;
  ConvArgs g{H, W, Cin, Cout, B, act, 0, out_fp32, bias, (plane_t*)out, (float*)out};
  g.ksize = ksize; g.stride = stride; g.res_planes = (const plane_t*)res_planes;
  if (g_conv_halo && ksize == 3 && stride == 1 && W % C64_TW == 0)
    return (g_conv_pair & 1) ? launch_conv_halo_pair(in_planes, w_planes, g, st) : launch_conv_halo(in_planes, w_planes, g, st);
  return BN == 128 ? launch_conv<128>(tmA, tmW, g, st) : launch_conv<64>(tmA, tmW, g, st);
}

int tc_split_planes(const float* in, void* out_planes, size_t n, cudaStream_t st) {
  split_planes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, (plane_t*)out_planes, n);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}
int tc_merge_planes(const void* in_planes, float* out, size_t n, cudaStream_t st) {
  merge_planes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const plane_t*)in_planes, out, n);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}

__global__ void transpose_taps_kernel(const float* __restrict__ in, float* __restrict__ out, int Cin, int Cout) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int n = 9 * Cin * Cout;
  if (i >= n) return;
  int co = i % Cout, ci = (i / Cout) % Cin, t = i / (Cout * Cin);
  out[((size_t)t * Cout + co) * Cin + ci] = in[i];
}

// unit-test hook: fp32 NHWC in/out, weights fp32 [9][Cin][Cout] (the CUDA-core layout); splits on the fly.
extern "C" int imw_debug_conv3x3_tc(const float* in, const float* w_tap_cin_cout, const float* bias, float* out, int B, int H,
                                    int W, int Cin, int Cout, int relu, int pool, void* scratch, size_t scratch_bytes,
                                    cudaStream_t st) {
  const size_t n_in = (size_t)B * H * W * Cin, n_w = (size_t)9 * Cout * Cin;
  Workspace ws(scratch, scratch_bytes);
  plane_t* in_p = ws.take<plane_t>(NP * n_in);
  float* w_t = ws.take<float>(n_w);
  plane_t* w_p = ws.take<plane_t>(NP * n_w);
  if (ws.overflow) { imw_set_error("imw_debug_conv3x3_tc: scratch too small (%zu needed)", ws.off); return IMW_ERR_WORKSPACE; }
  if (int e = tc_split_planes(in, in_p, n_in, st)) return e;
  transpose_taps_kernel<<<(unsigned)((n_w + 255) / 256), 256, 0, st>>>(w_tap_cin_cout, w_t, Cin, Cout);  // -> [tap][Cout][Cin]
  IMW_CHECK_LAUNCH();
  if (int e = tc_split_planes(w_t, w_p, n_w, st)) return e;
  return tc_conv3x3(in_p, w_p, bias, out, B, H, W, Cin, Cout, relu, pool, 1, st);
}

// bench hook: SuperPoint conv1a + conv1b fused (image in, pooled conv1b planes out)
extern "C" int imw_debug_conv1ab_fused(const float* img, const float* w1a, const float* b1a, const void* w1b_planes, const float* b1b,
                                       void* out_planes, int B, int H, int W, int pool, cudaStream_t st) {
  return tc_conv1ab_fused(img, w1a, b1a, w1b_planes, b1b, out_planes, B, H, W, pool, st);
}

// bench hook: the tcgen05 conv alone on pre-split operands (planes in, planes out)
extern "C" int imw_debug_conv3x3_tc_planes(const void* in_planes, const void* w_planes, const float* bias, void* out_planes, int B,
                                           int H, int W, int Cin, int Cout, int relu, int pool, cudaStream_t st) {
  return tc_conv3x3(in_planes, w_planes, bias, out_planes, B, H, W, Cin, Cout, relu, pool, 0, st);
}
