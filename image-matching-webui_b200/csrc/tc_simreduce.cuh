// Streaming similarity reductions on tcgen05 (3xTF32 split precision = fp32-equivalent products).
//
// Same contract and the same Op functors as simreduce.cuh: for every row i of set X (slot `own`) against all rows j
// of the other slot of the pair, s_ij = <x_i, x_j> is produced tile by tile -- here as 128 x 128 accumulator tiles
// in TMEM -- and folded into a per-row state; the N x M matrix never exists in memory.
//
//   warp 0      TMA producer: own-row tile [128 x 32 f32] and other-slot tile [128 x 32 f32] per k-block (SWIZZLE_128B)
//   warp 1      TMEM allocation + MMA issue: hi*hi into the main accumulator, hi*lo + lo*hi into the cross accumulator
//   warps 2..5  lo splitters (x_lo = x - trunc_tf32(x) into the twin tile; the landed fp32 tile itself is the hi operand)
//   warps 6..13 reduction: one TMEM lane = one row i per thread, two warps per TMEM sub-partition (columns 0..63 /
//               64..127 of the tile); a thread walks its columns in 32-wide chunks (op.accum32 when the functor has a
//               chunked form, else op.accum per element); the state lives in registers across all column tiles of the
//               row tile and the two half-states are merged through shared memory at the end
//
// Persistent CTAs walk the row tiles (m fastest over slots); two accumulator sets alternate so that the reduction
// of column tile t overlaps the MMAs of tile t+1.  Needs cap % 128 == 0, K % 32 == 0, ld % 4 == 0.
#pragma once
#include <type_traits>

#include "simreduce.cuh"
#include "tc_common.cuh"
#include "tc_gemm.cuh"

constexpr int TCS_THREADS = 64 + 128 + 256;

// functors may provide a chunked form  accum32(State&, const float (&s)[32], int i, int j0, int jn, int own, int other)
template <class Op, class = void>
struct tcs_has_accum32 : std::false_type {};
template <class Op>
struct tcs_has_accum32<Op, std::void_t<decltype(&Op::accum32)>> : std::true_type {};

template <class Op>
__global__ void __launch_bounds__(TCS_THREADS, 1) tc_simreduce_kernel(const __grid_constant__ CUtensorMap tmX, SimArgs a, Op op,
                                                                    int m_tiles) {
  constexpr int BN = 128;
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int A_BYTES = TC_BM * 128, B_BYTES = BN * 128, TILE_BYTES = A_BYTES + B_BYTES;
  constexpr int STAGE = 2 * TILE_BYTES;   // [A_hi | A_lo | B_hi | B_lo] (see tc_gemm.cuh)
  constexpr int W_OFF = 2 * A_BYTES;
  constexpr int ACC_COLS = 2 * BN;
  uint64_t* full = (uint64_t*)(smem + TC_STAGES * STAGE);
  uint64_t* empty = full + TC_STAGES;
  uint64_t* ready = empty + TC_STAGES;
  uint64_t* tmem_full = ready + TC_STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
  typename Op::State* half_state = (typename Op::State*)(smem + TC_STAGES * STAGE + 256);   // [128]

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmX);
    for (int s = 0; s < TC_STAGES; s++) { tc::mbar_init(full + s, 1); tc::mbar_init(empty + s, 1); tc::mbar_init(ready + s, 128); }
    for (int i = 0; i < 2; i++) { tc::mbar_init(tmem_full + i, 1); tc::mbar_init(tmem_empty + i, 256); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 2 * ACC_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int KB = a.K / TC_BK, tiles_per_slot = a.cap / TC_BM;

  // every role walks the same (row tile, column tile) sequence
  auto row_tile = [&](int mt, int& own, int& row0, int& n, int& m) -> bool {
    own = mt / tiles_per_slot; row0 = (mt % tiles_per_slot) * TC_BM;
    if (a.skip && a.skip[own >> 1]) return false;
    n = a.counts[own]; m = a.counts[own ^ 1];
    return row0 < n;
  };

  if (warp == 0) {
    if (lane == 0) {
      int c = 0;
      for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
        int own, row0, n, m;
        if (!row_tile(mt, own, row0, n, m)) continue;
        for (int j0 = 0; j0 < m; j0 += BN)
          for (int kb = 0; kb < KB; kb++, c++) {
            const int s = c % TC_STAGES, ph = (c / TC_STAGES) & 1;
            tc::mbar_wait(empty + s, ph ^ 1);
            tc::mbar_expect_tx(full + s, TILE_BYTES);
            tc::tma_load_2d(smem + s * STAGE, &tmX, full + s, kb * TC_BK, own * a.cap + row0);
            tc::tma_load_2d(smem + s * STAGE + W_OFF, &tmX, full + s, kb * TC_BK, (own ^ 1) * a.cap + j0);
          }
      }
    }
  } else if (warp == 1) {
    const bool leader = tc::elect_one();
    constexpr uint32_t idesc = tc::make_idesc(tc::FMT_TF32, TC_BM, BN), idesc2 = tc::make_idesc(tc::FMT_TF32, TC_BM, 2 * BN);
    int c = 0, i = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
      int own, row0, n, m;
      if (!row_tile(mt, own, row0, n, m)) continue;
      for (int j0 = 0; j0 < m; j0 += BN, i++) {
        const int acc = i & 1;
        tc::mbar_wait(tmem_empty + acc, ((i >> 1) & 1) ^ 1);
        tc::fence_after_sync();
        const uint32_t d_main = tmem_base + acc * ACC_COLS, d_cross = d_main + BN;
        for (int kb = 0; kb < KB; kb++, c++) {
          const int s = c % TC_STAGES, ph = (c / TC_STAGES) & 1;
          tc::mbar_wait(ready + s, ph);
          tc::fence_after_sync();
          const uint32_t a_addr = tc::smem_u32(smem + s * STAGE), b_addr = a_addr + W_OFF;
#pragma unroll
          for (int k = 0; k < TC_BK / 8; k++) {
            uint64_t ad = tc::make_smem_desc_sw128(a_addr + k * 32), bd = tc::make_smem_desc_sw128(b_addr + k * 32);
            uint64_t adl = tc::make_smem_desc_sw128(a_addr + A_BYTES + k * 32);
            if (leader) {
              tc::mma_tf32(d_main, ad, bd, idesc2, (kb | k) ? 1u : 0u);   // X_hi x [Y_hi | Y_lo] -> [main | cross]
              tc::mma_tf32(d_cross, adl, bd, idesc, 1u);                  // X_lo x Y_hi -> cross
            }
          }
          if (leader) tc::mma_commit(empty + s);
          __syncwarp();
        }
        if (leader) tc::mma_commit(tmem_full + acc);
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    const int t = threadIdx.x - 64;
    int c = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
      int own, row0, n, m;
      if (!row_tile(mt, own, row0, n, m)) continue;
      for (int j0 = 0; j0 < m; j0 += BN)
        for (int kb = 0; kb < KB; kb++, c++) {
          const int s = c % TC_STAGES, ph = (c / TC_STAGES) & 1;
          tc::mbar_wait(full + s, ph);
          uint4* base = reinterpret_cast<uint4*>(smem + s * STAGE);
#pragma unroll 8
          for (int idx = t; idx < TILE_BYTES / 16; idx += 128) {
            const bool in_a = idx < A_BYTES / 16;
            // the landed tile is the hi operand as it stands (kind::tf32 ignores the 13 low mantissa bits): only x_lo is written
            const uint4* hi = base + (in_a ? idx : idx + (W_OFF - A_BYTES) / 16);
            uint4* lo = const_cast<uint4*>(hi) + (in_a ? A_BYTES : B_BYTES) / 16;
            const uint4 v = *hi;
            uint4 l;
            l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(v.x & 0xFFFFE000u));
            l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(v.y & 0xFFFFE000u));
            l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(v.z & 0xFFFFE000u));
            l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(v.w & 0xFFFFE000u));
            *lo = l;
          }
          tc::fence_proxy_async();
          tc::mbar_arrive(ready + s);
        }
    }
  } else {
    const int q = warp % 4, half = (warp - 6) / 4;   // TMEM sub-partition, column half of the tile
    int i = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
      int own, row0, n, m;
      if (!row_tile(mt, own, row0, n, m)) continue;
      const int row = row0 + q * 32 + lane;
      const bool row_ok = row < n;
      typename Op::State st;
      op.init(st);
      for (int j0 = 0; j0 < m; j0 += BN, i++) {
        const int acc = i & 1;
        tc::mbar_wait(tmem_full + acc, (i >> 1) & 1);
        tc::fence_after_sync();
        const uint32_t lane_addr = tmem_base + acc * ACC_COLS + ((uint32_t)(q * 32) << 16) + half * 64;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 32) {
          float v[32], t[32];
          tc::tmem_ld32(lane_addr + c0, v);
          tc::tmem_ld32(lane_addr + BN + c0, t);
          if (c0 == 32) {
            tc::fence_before_sync();
            tc::mbar_arrive(tmem_empty + acc);
          }
          const int jc = j0 + half * 64 + c0, jn = min(32, m - jc);  // warp-uniform; may be <= 0
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] += t[j];
          if (row_ok && jn > 0) {
            if constexpr (tcs_has_accum32<Op>::value) {
              op.accum32(st, v, row, jc, jn, own, own ^ 1);
            } else {
#pragma unroll
              for (int j = 0; j < 32; j++)
                if (j < jn) op.accum(st, v[j], row, jc + j, own, own ^ 1);
            }
          }
        }
      }
      // merge the two column halves of every row (functor merges are order-insensitive, ties included)
      if (half == 1) half_state[q * 32 + lane] = st;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (half == 0 && row_ok) {
        op.merge(st, half_state[q * 32 + lane]);
        op.store(st, own, row);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 2 * ACC_COLS);
}

// true when the tensor-core version can take this problem (else call launch_simreduce)
static inline bool tc_simreduce_ok(const SimArgs& a) {
  return a.cap % TC_BM == 0 && a.K % TC_BK == 0 && a.ld % 4 == 0 && ((uintptr_t)a.X % 16) == 0;
}

template <class Op>
static inline int launch_tc_simreduce(const SimArgs& a, int slots, Op op, cudaStream_t st) {
  CUtensorMap tmX;
  if (int e = tc_make_map_2d_f32(&tmX, a.X, (uint64_t)slots * a.cap, (uint64_t)a.K, (uint64_t)a.ld, TC_BK, TC_BM)) return e;
  constexpr size_t smem = (size_t)TC_STAGES * 2 * (TC_BM * 128 + 128 * 128) + 1024 + 256 + 128 * sizeof(typename Op::State);
  IMW_SMEM_ATTR_ONCE(tc_simreduce_kernel<Op>, smem);
  const int num_sms = imw_num_sms();
  const int m_tiles = slots * (a.cap / TC_BM);
  if (m_tiles == 0) return IMW_OK;
  tc_simreduce_kernel<Op><<<dim3((unsigned)(m_tiles < num_sms ? m_tiles : num_sms)), TCS_THREADS, smem, st>>>(tmX, a, op, m_tiles);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}
