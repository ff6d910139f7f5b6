// Streaming similarity reductions on tcgen05 (split-fp16 operands = fp32-equivalent products).
//
// Same contract and the same Op functors as simreduce.cuh: for every row i of set X (slot `own`) against all rows j
// of the other slot of the pair, s_ij = <x_i, x_j> is produced tile by tile -- here as 128 x 128 accumulator tiles
// in TMEM -- and folded into a per-row state; the N x M matrix never exists in memory.
//
// Operands: X arrives PRE-SPLIT as two fp16 planes [2][slots * cap][K] (hi = fp16(x), lo = fp16((x - hi) 2^11):
// split_planes.cuh; written once by sim_split_planes / the producing kernel), so that
//   * nothing is split inside the kernel: round 1 / early round 2 ran 3xTF32 with four splitter warps rewriting every landed
//     tile (each own-row tile 32 times per row tile, each column tile once per row tile) and kind::tf32 MMAs at half the
//     fp16 rate; kind::f16 over K = 16 per instruction halves the tensor time and the shared-memory operand bytes;
//   * the OWN row tile (both planes, all of K: 64 KB at K = 128) is loaded ONCE per row tile and stays resident in shared
//     memory while the column tiles stream through a ring: L2 -> SM traffic per tile is the column tile only (64 KB at
//     K = 128) instead of own + column tile.
// s_ij = hi.hi + (hi.lo + lo.hi) 2^-11: the scaled cross terms accumulate in their own TMEM accumulator.
//
//   warp 0       TMA producer: resident own tile (once per row tile), column tiles [hi | lo] per 64-element k-block
//   warp 1       TMEM allocation + MMA issue: A_hi x [B_hi | B_lo] (one N = 256 MMA -> [main | cross]) and A_lo x B_hi -> cross
//   warps 2..17  reduction: one TMEM lane = one row i per thread, FOUR warps per TMEM sub-partition (32 of the tile's 128
//                columns each, op.accum32 when the functor has a chunked form); the state lives in registers across all
//                column tiles of the row tile and the four partial states are merged through shared memory at the end
//
// Persistent CTAs walk the row tiles; two accumulator sets alternate so that the reduction of column tile t overlaps
// the MMAs of tile t+1.  Needs cap % 128 == 0, K % 64 == 0, K <= 256.
#pragma once
#include <type_traits>

#include "simreduce.cuh"
#include "split_planes.cuh"
#include "tc_common.cuh"
#include "tc_gemm.cuh"

constexpr int TCS_RWARPS = 16, TCS_THREADS = 64 + 32 * TCS_RWARPS;
constexpr int TCS_BM = 128, TCS_BN = 128, TCS_BKE = 64;           // k-block = 64 fp16 = one 128-byte swizzled row
constexpr int TCS_PLANE_BYTES = 128 * 128;                         // one plane of one k-block of a 128-row tile: 16 KB
constexpr int TCS_SMEM_BUDGET = 227 * 1024 - 1024 /*alignment slack*/ - 256 /*barriers*/ - 3 * 128 * 16 /*partial states*/;

// functors may provide a chunked form  accum32(State&, const float (&s)[32], int i, int j0, int jn, int own, int other)
template <class Op, class = void>
struct tcs_has_accum32 : std::false_type {};
template <class Op>
struct tcs_has_accum32<Op, std::void_t<decltype(&Op::accum32)>> : std::true_type {};

// X [slots][cap][ld] fp32 (first K columns) -> planes [2][slots * cap][K] fp16.  Rows beyond a slot's count are not
// written: whatever the buffer holds there only reaches output rows / columns that the reduction masks out.
static __global__ void __launch_bounds__(256) sim_split_planes_kernel(const float* __restrict__ X, plane_t* __restrict__ planes,
                                                                      const int* __restrict__ counts, int cap, int ld, int K,
                                                                      long long plane_elems) {
  const int z = blockIdx.y, kq = K / 4;
  const int n = counts ? counts[z] : cap;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)n * kq; idx += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(idx / kq), q = (int)(idx % kq);
    const float4 v = *reinterpret_cast<const float4*>(X + ((long long)z * cap + row) * ld + q * 4);
    plane_t h[4], l[4];
    split2(v.x, h[0], l[0]); split2(v.y, h[1], l[1]); split2(v.z, h[2], l[2]); split2(v.w, h[3], l[3]);
    const long long o = ((long long)z * cap + row) * K + q * 4;
    *reinterpret_cast<uint2*>(planes + o) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(planes + plane_elems + o) = *reinterpret_cast<const uint2*>(l);
  }
}

struct TcSimArgs {
  int cap, K;
  const int* counts;     // [slots]
  const int* skip;       // optional, indexed by pair
  int plane_rows;        // slots * cap: row offset of the lo plane in the tensor map
  int stages;            // column-tile ring depth (what fits beside the resident own tile)
};

template <class Op>
__global__ void __launch_bounds__(TCS_THREADS, 1) tc_simreduce_kernel(const __grid_constant__ CUtensorMap tmX, TcSimArgs a, Op op,
                                                                    int m_tiles) {
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
  const int KB = a.K / TCS_BKE, STAGES = a.stages;
  constexpr int STAGE = 2 * TCS_PLANE_BYTES;                 // [B_hi | B_lo] of one k-block: adjacent -> one N = 256 operand
  constexpr int ACC_COLS = 2 * TCS_BN;
  uint8_t* sA = smem;                                        // [KB][hi | lo][128 rows x 128 B]
  uint8_t* sB = sA + (size_t)KB * STAGE;                     // [STAGES][hi | lo][128 rows x 128 B]
  uint64_t* full = (uint64_t*)(sB + (size_t)STAGES * STAGE);
  uint64_t* empty = full + 8;
  uint64_t* a_full = empty + 8;
  uint64_t* a_empty = a_full + 1;
  uint64_t* tmem_full = a_empty + 1;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;  // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
  typename Op::State* part_state = (typename Op::State*)((uint8_t*)full + 256);   // [3][128]

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmX);
    for (int s = 0; s < STAGES; s++) { tc::mbar_init(full + s, 1); tc::mbar_init(empty + s, 1); }
    tc::mbar_init(a_full, 1); tc::mbar_init(a_empty, 1);
    for (int i = 0; i < 2; i++) { tc::mbar_init(tmem_full + i, 1); tc::mbar_init(tmem_empty + i, 32 * TCS_RWARPS); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 2 * ACC_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_slot = a.cap / TCS_BM;

  // every role walks the same (row tile, column tile) sequence
  auto row_tile = [&](int mt, int& own, int& row0, int& n, int& m) -> bool {
    own = mt / tiles_per_slot; row0 = (mt % tiles_per_slot) * TCS_BM;
    if (a.skip && a.skip[own >> 1]) return false;
    n = a.counts[own]; m = a.counts[own ^ 1];
    return row0 < n;
  };

  if (warp == 0) {
    if (lane == 0) {
      int c = 0, t = 0;
      for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
        int own, row0, n, m;
        if (!row_tile(mt, own, row0, n, m) || m <= 0) continue;
        if (t > 0) tc::mbar_wait(a_empty, (t - 1) & 1);       // the previous row tile's MMAs have read the resident tile
        tc::mbar_expect_tx(a_full, (uint32_t)(KB * STAGE));
        for (int kb = 0; kb < KB; kb++) {
          tc::tma_load_2d(sA + (size_t)kb * STAGE, &tmX, a_full, kb * TCS_BKE, own * a.cap + row0);
          tc::tma_load_2d(sA + (size_t)kb * STAGE + TCS_PLANE_BYTES, &tmX, a_full, kb * TCS_BKE, a.plane_rows + own * a.cap + row0);
        }
        t++;
        for (int j0 = 0; j0 < m; j0 += TCS_BN)
          for (int kb = 0; kb < KB; kb++, c++) {
            const int s = c % STAGES, ph = (c / STAGES) & 1;
            tc::mbar_wait(empty + s, ph ^ 1);
            tc::mbar_expect_tx(full + s, STAGE);
            tc::tma_load_2d(sB + (size_t)s * STAGE, &tmX, full + s, kb * TCS_BKE, (own ^ 1) * a.cap + j0);
            tc::tma_load_2d(sB + (size_t)s * STAGE + TCS_PLANE_BYTES, &tmX, full + s, kb * TCS_BKE, a.plane_rows + (own ^ 1) * a.cap + j0);
          }
      }
    }
  } else if (warp == 1) {
    const bool leader = tc::elect_one();
    constexpr uint32_t idesc = tc::make_idesc(tc::FMT_F16, TCS_BM, TCS_BN), idesc2 = tc::make_idesc(tc::FMT_F16, TCS_BM, 2 * TCS_BN);
    int c = 0, i = 0, t = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
      int own, row0, n, m;
      if (!row_tile(mt, own, row0, n, m) || m <= 0) continue;
      tc::mbar_wait(a_full, t & 1);
      for (int j0 = 0; j0 < m; j0 += TCS_BN, i++) {
        const int acc = i & 1;
        tc::mbar_wait(tmem_empty + acc, ((i >> 1) & 1) ^ 1);
        tc::fence_after_sync();
        const uint32_t d_main = tmem_base + acc * ACC_COLS, d_cross = d_main + TCS_BN;
        for (int kb = 0; kb < KB; kb++, c++) {
          const int s = c % STAGES, ph = (c / STAGES) & 1;
          tc::mbar_wait(full + s, ph);
          tc::fence_after_sync();
          const uint32_t a_addr = tc::smem_u32(sA + (size_t)kb * STAGE), b_addr = tc::smem_u32(sB + (size_t)s * STAGE);
#pragma unroll
          for (int k = 0; k < TCS_BKE / 16; k++) {   // 16 fp16 = 32 bytes along K inside the 128-byte swizzle atom
            uint64_t ad = tc::make_smem_desc_sw128(a_addr + k * 32), bd = tc::make_smem_desc_sw128(b_addr + k * 32);
            uint64_t adl = tc::make_smem_desc_sw128(a_addr + TCS_PLANE_BYTES + k * 32);
            if (leader) {
              tc::mma_f16(d_main, ad, bd, idesc2, (kb | k) ? 1u : 0u);   // X_hi x [Y_hi | Y_lo] -> [main | cross]
              tc::mma_f16(d_cross, adl, bd, idesc, 1u);                  // X_lo x Y_hi -> cross
            }
          }
          if (leader) tc::mma_commit(empty + s);
          __syncwarp();
        }
        if (leader) tc::mma_commit(tmem_full + acc);
        __syncwarp();
      }
      if (leader) tc::mma_commit(a_empty);
      __syncwarp();
      t++;
    }
  } else {
    const int q = warp % 4, qt = (warp - 2) / 4;   // TMEM sub-partition, column quarter of the tile
    const int tr = q * 32 + lane;
    int i = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
      int own, row0, n, m;
      if (!row_tile(mt, own, row0, n, m)) continue;
      const int row = row0 + tr;
      const bool row_ok = row < n;
      typename Op::State st;
      op.init(st);
      for (int j0 = 0; j0 < m; j0 += TCS_BN, i++) {
        const int acc = i & 1;
        tc::mbar_wait(tmem_full + acc, (i >> 1) & 1);
        tc::fence_after_sync();
        const uint32_t lane_addr = tmem_base + acc * ACC_COLS + ((uint32_t)(q * 32) << 16) + qt * 32;
        float v[32], t[32];
        tc::tmem_ld32(lane_addr, v);
        tc::tmem_ld32(lane_addr + TCS_BN, t);
        tc::fence_before_sync();
        tc::mbar_arrive(tmem_empty + acc);
        const int jc = j0 + qt * 32, jn = min(32, m - jc);  // warp-uniform; may be <= 0
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = fmaf(t[j], PLANE_LO_INV, v[j]);
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
      // merge the four column quarters of every row (functor merges are order-insensitive, ties included)
      if (qt > 0) part_state[(qt - 1) * 128 + tr] = st;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * TCS_RWARPS) : "memory");
      if (qt == 0 && row_ok) {
        op.merge(st, part_state[tr]); op.merge(st, part_state[128 + tr]); op.merge(st, part_state[256 + tr]);
        op.store(st, own, row);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * TCS_RWARPS) : "memory");
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 2 * ACC_COLS);
}

// true when the tensor-core version can take this problem (else call launch_simreduce)
static inline bool tc_simreduce_ok(const SimArgs& a) {
  return a.cap % TCS_BM == 0 && a.K % TCS_BKE == 0 && a.K <= 256 && a.ld % 4 == 0 && ((uintptr_t)a.X % 16) == 0;
}
// bytes of the fp16 planes of X: [2][slots * cap][K]
static inline size_t tc_simreduce_plane_bytes(int slots, int cap, int K) { return (size_t)2 * slots * cap * K * sizeof(plane_t); }

// fp32 X -> planes (one launch; callers whose producer writes the planes itself skip it)
static inline int tc_simreduce_split(const SimArgs& a, int slots, plane_t* planes, cudaStream_t st) {
  if (slots == 0 || a.cap == 0) return IMW_OK;
  const long long per_slot = (long long)a.cap * (a.K / 4);
  dim3 grid((unsigned)((per_slot + 255) / 256 < 64 ? (per_slot + 255) / 256 : 64), slots);
  sim_split_planes_kernel<<<grid, 256, 0, st>>>(a.X, planes, a.counts, a.cap, a.ld, a.K, (long long)slots * a.cap * a.K);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}

template <class Op>
static inline int launch_tc_simreduce(const SimArgs& a, int slots, const plane_t* planes, Op op, cudaStream_t st) {
  static_assert(sizeof(typename Op::State) <= 16, "partial-state staging assumes <= 16 bytes per state");
  CUtensorMap tmX;
  if (int e = tc_make_map_2d_f16(&tmX, planes, (uint64_t)2 * slots * a.cap, (uint64_t)a.K, TCS_BKE, TCS_BM)) return e;
  const int KB = a.K / TCS_BKE;
  int stages = (TCS_SMEM_BUDGET - KB * 2 * TCS_PLANE_BYTES) / (2 * TCS_PLANE_BYTES);
  if (stages > 8) stages = 8;
  const size_t smem = (size_t)(KB + stages) * 2 * TCS_PLANE_BYTES + 1024 + 256 + 3 * 128 * 16;
  IMW_SMEM_ATTR_ONCE(tc_simreduce_kernel<Op>, 227 * 1024);
  const int num_sms = imw_num_sms();
  const int m_tiles = slots * (a.cap / TCS_BM);
  if (m_tiles == 0) return IMW_OK;
  TcSimArgs ta{a.cap, a.K, a.counts, a.skip, slots * a.cap, stages};
  tc_simreduce_kernel<Op><<<dim3((unsigned)(m_tiles < num_sms ? m_tiles : num_sms)), TCS_THREADS, smem, st>>>(tmX, ta, op, m_tiles);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}
