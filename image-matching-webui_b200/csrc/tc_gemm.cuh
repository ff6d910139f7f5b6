// tcgen05 TF32 GEMM   C[M,N] = A[M,K] * W[N,K]^T   (fp32 accumulate in TMEM).
//
// SPLIT = 1: operands read as single TF32 (10-bit mantissa): fast mode.
// SPLIT = 3: 3xTF32 split precision = fp32-equivalent products  x = x_hi + x_lo  with x_hi = x truncated to TF32 and
//            x_lo = x - x_hi.  kind::tf32 ignores the 13 low mantissa bits of its fp32 operands, so the tile TMA landed IS
//            the x_hi operand as it stands; four splitter warps only compute x_lo into the slot right behind it (same
//            swizzled offsets).  Constant weights come with their lo plane pre-computed by the host (rows
//            [wlo_rows, 2 wlo_rows) of W, TcGemmArgs::wlo_rows), so the splitters touch the activation tile only: one
//            16-byte shared-memory load and one store per four elements instead of two loads and four stores (the
//            round-1 kernel, bound by exactly this LSU traffic: 63 % of the shared-memory wavefront peak).  The MMA warp
//            issues A_hi x [W_hi | W_lo] (one MMA of N = 2 BN) and A_lo x W_hi (dropped lo*lo term and the TF32
//            rounding of the lo parts are <= 2^-21 relative).  This is the default: LightGlue scores stay within the 1e-3
//            parity tolerance, which single TF32 does not (measured 1.4e-2 on the golden pairs).
//
// Persistent: one CTA per SM walks the (m-tile, n-tile) list; the smem ring runs across tile boundaries and two TMEM
// accumulator sets alternate, so the epilogue of tile i overlaps the main loop of tile i+1.
//   warp 0       TMA producer: A tile [128 x 32 f32] and W tile [BN x 32 f32] per k-block, SWIZZLE_128B,
//                through a STAGES-deep full/empty mbarrier ring
//   warp 1       TMEM allocation + single-thread tcgen05.mma issue (M = 128, K = 8 per instruction)
//   warps 2..5   hi/lo splitters (SPLIT == 3)
//   warps 6..13  epilogue, two warps per TMEM sub-partition (half of the tile's columns each): tcgen05.ld (one TMEM
//                lane = one output row per thread), transpose through a padded smem tile so that global accesses
//                run along the columns, apply the same functor epilogues as the CUDA-core GEMM, store
// Rows are [slots][cap] with cap % 128 == 0: a tile never straddles two slots; finished pairs / rows beyond the
// slot's count are skipped on the device.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "split_planes.cuh"
#include "tc_common.cuh"

struct TcGemmArgs {
  int K;                 // % 32 == 0
  int N;                 // % BN == 0
  int tiles_per_slot;    // cap / 128
  const int* counts;     // [slots] valid rows (nullable: all rows valid)
  const int* skip;       // optional skip[z >> skip_shift]
  int skip_shift;
  const int* wsel_minus1;  // optional per-pair weight slab selection: W rows offset (wsel-1) * wsel_rows
  int wsel_shift;
  int wsel_rows;
  int pair_product;      // 1: "W" is the OTHER slot of the pair in the same row space as A (rows (z^1)*cap + n0) and
                         //    only even slots produce output: C[pair] = X_0 X_1^T (score matrices)
  long long wlo_rows;    // > 0: W holds a second plane W_lo = W - trunc_tf32(W) wlo_rows rows below W (host pre-split);
                         //      0: W_lo is computed in the kernel like A_lo
  const void* w_planes;  // non-null: the weights as split-fp16 planes [2][w_plane_rows][K] (split_planes.cuh, packed by the host):
  long long w_plane_rows;  //         the GEMM runs on kind::f16 (tc_gemm_f16_kernel) -- half the tensor time and 2/3 of the L2 -> SM bytes
  const void* a_planes = nullptr;   // non-null (with w_planes): the ACTIVATIONS as split-fp16 planes [2][a_plane_rows][a_plane_ld] written by
  long long a_plane_rows = 0;       //   their producer (e.g. LayerNorm + GELU): both operand tiles land by TMA, no splitter warps -- the kernel is
  int a_plane_ld = 0;               //   bound by its CUDA-core work (tc_gemm_ablate), and the split is redone for every N tile otherwise
};

constexpr int TC_BM = 128, TC_BK = 32, TC_STAGES = 3, TC_THREADS = 448;

template <int BN, int SPLIT>
constexpr size_t tc_gemm_smem_bytes() {
  return (size_t)TC_STAGES * (SPLIT == 3 ? 2 : 1) * (TC_BM * 128 + BN * 128) + 1024 /*alignment slack*/ + 256 /*barriers*/ +
         8 * 32 * 33 * sizeof(float) /*epilogue transpose tiles*/;
}

// Persistent: one CTA per SM walks the (m-tile, n-tile) list (n fastest, so the CTAs that share an A tile run
// together and hit L2).  Warp roles: 0 TMA producer, 1 MMA issuer, 2-5 hi/lo splitters (SPLIT == 3),
// 6-13 epilogue (two warps per TMEM sub-partition, half of the tile's columns each).  The smem ring runs across tile boundaries and two TMEM accumulator sets alternate, so the
// epilogue of tile i overlaps the main loop of tile i+1.
template <int BN, int SPLIT, bool WLO, class Epi>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                    const __grid_constant__ CUtensorMap tmW, TcGemmArgs g, Epi epi,
                                                                    int m_tiles) {
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
  // stage layout, SPLIT == 3: [A_hi | A_lo | W_hi | W_lo] -- TMA lands the fp32 tiles in the hi slots, the splitters rewrite
  // them in place and fill the lo slots; W_hi and W_lo are adjacent so that A_hi x [W_hi | W_lo] is ONE MMA of N = 2 BN
  // (fewer shared-memory operand reads than three MMAs).  SPLIT == 1: [A | W].
  constexpr int A_BYTES = TC_BM * 128, B_BYTES = BN * 128, TILE_BYTES = A_BYTES + B_BYTES;
  constexpr int STAGE = (SPLIT == 3 ? 2 : 1) * TILE_BYTES;
  constexpr int W_OFF = (SPLIT == 3 ? 2 : 1) * A_BYTES;
  constexpr int ACC_COLS = (SPLIT == 3 ? 2 : 1) * BN;   // SPLIT == 3: hi*hi and cross-term accumulators (see header)
  uint64_t* full = (uint64_t*)(smem + TC_STAGES * STAGE);
  uint64_t* empty = full + TC_STAGES;
  uint64_t* ready = empty + TC_STAGES;   // SPLIT == 3: hi/lo tiles written by the splitter warps
  uint64_t* tmem_full = ready + TC_STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
  float* epi_tiles = (float*)(smem + TC_STAGES * STAGE + 256);  // [4 warps][32][33]

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmW);
    for (int s = 0; s < TC_STAGES; s++) { tc::mbar_init(full + s, 1); tc::mbar_init(empty + s, 1); tc::mbar_init(ready + s, 128); }
    for (int a = 0; a < 2; a++) { tc::mbar_init(tmem_full + a, 1); tc::mbar_init(tmem_empty + a, 256); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 2 * ACC_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int KB = g.K / TC_BK, n_tiles = g.N / BN, total = m_tiles * n_tiles;

  // every role walks the same tile sequence and skips the same tiles (finished pairs, rows beyond the count)
  auto tile_info = [&](int tile, int& m_tile, int& n0, int& z, int& row0, int& nrows) -> bool {
    m_tile = tile / n_tiles; n0 = (tile % n_tiles) * BN;
    z = m_tile / g.tiles_per_slot; row0 = (m_tile % g.tiles_per_slot) * TC_BM;
    if (g.skip && g.skip[z >> g.skip_shift]) return false;
    if (g.pair_product && (z & 1)) return false;
    nrows = g.counts ? g.counts[z] : (g.tiles_per_slot * TC_BM);
    if (g.pair_product && g.counts && n0 >= g.counts[z ^ 1]) return false;
    return row0 < nrows;
  };

  if (warp == 0) {
    if (lane == 0) {
      int c = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        int m_tile, n0, z, row0, nrows;
        if (!tile_info(tile, m_tile, n0, z, row0, nrows)) continue;
        const int w_row0 = g.pair_product ? (z ^ 1) * g.tiles_per_slot * TC_BM + n0
                                          : n0 + (g.wsel_minus1 ? (g.wsel_minus1[z >> g.wsel_shift] - 1) * g.wsel_rows : 0);
        for (int kb = 0; kb < KB; kb++, c++) {
          const int s = c % TC_STAGES, ph = (c / TC_STAGES) & 1;
          tc::mbar_wait(empty + s, ph ^ 1);
          tc::mbar_expect_tx(full + s, TILE_BYTES + (WLO ? B_BYTES : 0));
          tc::tma_load_2d(smem + s * STAGE, &tmA, full + s, kb * TC_BK, m_tile * TC_BM);
          tc::tma_load_2d(smem + s * STAGE + W_OFF, &tmW, full + s, kb * TC_BK, w_row0);
          if (WLO) tc::tma_load_2d(smem + s * STAGE + W_OFF + B_BYTES, &tmW, full + s, kb * TC_BK, (int)(g.wlo_rows + w_row0));
        }
      }
    }
  } else if (warp == 1) {
    const bool leader = tc::elect_one();
    constexpr uint32_t idesc = tc::make_idesc(tc::FMT_TF32, TC_BM, BN), idesc2 = tc::make_idesc(tc::FMT_TF32, TC_BM, 2 * BN);
    int c = 0, i = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      int m_tile, n0, z, row0, nrows;
      if (!tile_info(tile, m_tile, n0, z, row0, nrows)) continue;
      const int acc = i & 1;
      tc::mbar_wait(tmem_empty + acc, ((i >> 1) & 1) ^ 1);  // epilogue drained this accumulator set
      tc::fence_after_sync();
      const uint32_t d_main = tmem_base + acc * ACC_COLS, d_cross = d_main + BN;
      for (int kb = 0; kb < KB; kb++, c++) {
        const int s = c % TC_STAGES, ph = (c / TC_STAGES) & 1;
        tc::mbar_wait(SPLIT == 3 ? ready + s : full + s, ph);
        tc::fence_after_sync();
        const uint32_t a_addr = tc::smem_u32(smem + s * STAGE), b_addr = a_addr + W_OFF;
#pragma unroll
        for (int k = 0; k < TC_BK / 8; k++) {
          // advance 8 tf32 = 32 bytes along K inside the 128-byte swizzle atom
          uint64_t ad = tc::make_smem_desc_sw128(a_addr + k * 32), bd = tc::make_smem_desc_sw128(b_addr + k * 32);
          uint64_t adl = tc::make_smem_desc_sw128(a_addr + A_BYTES + k * 32);
          if (leader) {
            if (SPLIT == 3) {
              tc::mma_tf32(d_main, ad, bd, idesc2, (kb | k) ? 1u : 0u);   // A_hi x [W_hi | W_lo] -> [main | cross]
              tc::mma_tf32(d_cross, adl, bd, idesc, 1u);                  // A_lo x W_hi -> cross
            } else {
              tc::mma_tf32(d_main, ad, bd, idesc, (kb | k) ? 1u : 0u);
            }
          }
        }
        if (leader) tc::mma_commit(empty + s);  // smem stage free once these MMAs have read it
        __syncwarp();
      }
      if (leader) tc::mma_commit(tmem_full + acc);  // accumulator complete
      __syncwarp();
      i++;
    }
  } else if (warp < 6) {
    if (SPLIT == 3) {
      // splitter: x_lo = x - trunc_tf32(x) into the twin tile (the landed tile itself is the hi operand: the tensor core ignores
      // the low 13 mantissa bits); elementwise, so the TMA swizzle is irrelevant.  With a host-provided W_lo plane only the
      // activation tile is touched.
      const int t = threadIdx.x - 64;  // 0..127
      int c = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        int m_tile, n0, z, row0, nrows;
        if (!tile_info(tile, m_tile, n0, z, row0, nrows)) continue;
        for (int kb = 0; kb < KB; kb++, c++) {
          const int s = c % TC_STAGES, ph = (c / TC_STAGES) & 1;
          tc::mbar_wait(full + s, ph);
          uint4* base = reinterpret_cast<uint4*>(smem + s * STAGE);
#pragma unroll 8
          for (int idx = t; idx < (WLO ? A_BYTES : TILE_BYTES) / 16; idx += 128) {
            // A region: hi at idx, lo A_BYTES behind; W region: hi at W_OFF + .., lo B_BYTES behind
            const bool in_a = idx < A_BYTES / 16;
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
          tc::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
          tc::mbar_arrive(ready + s);
        }
      }
    }
  } else {
    const int q = warp % 4;  // TMEM sub-partition this warp may read: lanes [32q, 32q+32)
    const int half = (warp - 6) / 4;  // two warps per sub-partition: columns [0, BN/2) and [BN/2, BN)
    // TMEM gives one output ROW per thread; global memory wants one row per warp instruction.  Transpose each
    // 32x32 block through a padded smem tile so that lanes run along the columns: every load/store of the
    // functor epilogue is one fully coalesced 128-byte line.
    float* T = epi_tiles + (warp - 6) * 32 * 33;
    int i = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      int m_tile, n0, z, row0, nrows;
      if (!tile_info(tile, m_tile, n0, z, row0, nrows)) continue;
      const int acc = i & 1;
      tc::mbar_wait(tmem_full + acc, (i >> 1) & 1);
      tc::fence_after_sync();
      const int row_base = row0 + q * 32;
      const uint32_t lane_addr = tmem_base + acc * ACC_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
        float v[32];
        tc::tmem_ld32(lane_addr + c0, v);
        if (SPLIT == 3) {
          float t[32];
          tc::tmem_ld32(lane_addr + BN + c0, t);
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] += t[j];
        }
        if (c0 + 32 >= (half + 1) * (BN / 2)) {  // this warp's last TMEM read of the accumulator set: hand it back to the MMA warp
          tc::fence_before_sync();
          tc::mbar_arrive(tmem_empty + acc);
        }
        // functors may consume a 32-column chunk in the native thread-per-row layout (V^T for the attention kernel)
        if (epi.rowwise(z, row_base + lane, row_base + lane < nrows, n0 + c0, v)) continue;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; j++) T[lane * 33 + j] = v[j];
        __syncwarp();
        const int rmax = min(32, nrows - row_base);  // warp-uniform
        // all global reads of the epilogue (residual / rotary table) are issued before the first dependent use
        float2 pre[32];
#pragma unroll
        for (int r = 0; r < 32; r++) pre[r] = (r < rmax) ? epi.prefetch(z, row_base + r, n0 + c0 + lane) : make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 32; r++)
          if (r < rmax) epi.elem(z, row_base + r, n0 + c0 + lane, T[r * 33 + lane], pre[r]);
      }
      i++;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 2 * ACC_COLS);
}

// ---------------------------------------------------------------------------------------------------------------------------
// Split-fp16 variant (the default for every linear with host-packed weights):  x = hi + lo 2^-11, three kind::f16 products per
// fp32-equivalent product (split_planes.cuh).  Why: the 3xTF32 kernel above is bound by the L2 -> SM stream of its fp32 operand
// tiles (ncu round 2: 9 TB/s, 1.58 us per 64 elements of K at 96 KB) with the tensor pipe half idle.  Here
//   * weights arrive as fp16 planes [W_hi ; W_lo] (4 B per element instead of 8 B): TMA, adjacent tiles -> one N = 2 BN operand;
//   * the activation tile lands as fp32 (TMA, two 32-column boxes per 64-element k-block) and four splitter warps convert it
//     IN PLACE to the two fp16 operand tiles (K-major, SWIZZLE_128B written by hand: 16-byte chunk c of row r sits at c ^ (r & 7));
//   * MMAs are kind::f16 (K = 16 per instruction): A_hi x [W_hi | W_lo] -> [main | cross], A_lo x W_hi -> cross; the epilogue adds
//     main + cross 2^-11.
// Same persistent tile walk, roles and functor epilogues as the kernel above.
// timing ablations of the split-fp16 GEMM (imw_debug_set_gemm_ablate): bit 1 = splitters skip their work, bit 2 = epilogue functors
// skipped (results are garbage).  Measured on 131072 x 768 x 256: full 0.290 ms, no split 0.250, no epilogue 0.188, neither 0.113 (= 82 %
// of the tensor peak): the kernel is bound by the CUDA-core work of its splitter and epilogue warps, not by the operand stream.
// (A weights-stationary CTA-pair form -- cta_group::2, M = 256, half of the weight slice resident per CTA, activations only
// streamed -- was built and measured in round 2: bit-identical results, 0.407 ms on the same shape: its cross-CTA ready / empty
// round trips cost more than the halved L2 traffic saves.  Removed.)
inline int& tc_gemm_ablate() {
  static int mode = 0;
  return mode;
}
constexpr int TH_BKE = 64, TH_STAGES = 3, TH_SPLIT_WARPS = 4, TH_THREADS = 64 + 32 * TH_SPLIT_WARPS + 256;

// Functors may provide a PAIR form of the epilogue -- two adjacent output columns per lane, sixteen column pairs x two rows per
// warp instruction:  float4 pair_prefetch(z, row, col)  (every global read the pair needs)  and  pair(z, row, col, a0, a1, pre).
// It halves the instruction count of the element-per-lane form (one address, one 8-byte store, no partner shuffle for the
// rotary pairs); the QKV projections, whose epilogues write three re-laid-out tensors, were bound by exactly that count
// (ncu round 2: 48.7 k warp instructions per 128 x 128 tile, tensor pipe 21 %).
template <class Epi, class = void>
struct tc_has_pair : std::false_type {};
template <class Epi>
struct tc_has_pair<Epi, std::void_t<decltype(&Epi::pair)>> : std::true_type {};

// ... and a QUAD form -- four adjacent columns per lane, eight quads x four rows per warp instruction:  bool quad_ok()  (run-time
// alignment check),  float4 quad_col(col)  (per-column constants, once per block),  float4 quad_prefetch(z, row, col)  and
// quad(z, row, col, a, colconst, pre).  The block goes through a swizzled [32 rows][8 x 16 B] smem tile (conflict-free 16-byte
// stores by row and loads by quad).
template <class Epi, class = void>
struct tc_has_quad : std::false_type {};
template <class Epi>
struct tc_has_quad<Epi, std::void_t<decltype(&Epi::quad)>> : std::true_type {};

template <int BN>
constexpr size_t tc_gemm_f16_smem_bytes() {
  return (size_t)TH_STAGES * (2 * TC_BM * 128 /*fp32 landing = A planes, split in place*/ + 2 * BN * 128 /*W planes*/) + 1024 + 256 +
         8 * 32 * 33 * sizeof(float);
}

template <int BN, class Epi, bool APL>
__global__ void __launch_bounds__(TH_THREADS, 1) tc_gemm_f16_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmW, TcGemmArgs g, Epi epi,
                                                                   int m_tiles, int ablate) {
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
  // The fp32 tile lands as two 32-column boxes of 16 KB; the splitters rewrite it IN PLACE as the two fp16 operand tiles
  // (A_hi over box 0, A_lo over box 1: 128 B of fp32 per row and box = 64 B of hi + 64 B of lo, so the bytes match).  One thread
  // owns one whole row (256 B in registers before the first store): no cross-thread hazard, no separate landing buffer --
  // three stages of 64 KB fit where two of 96 KB did (the kernel is bound by the bytes it keeps in flight, not by the tensor pipe).
  constexpr int LAND_BYTES = 2 * TC_BM * 128, AP_BYTES = TC_BM * 128, WP_BYTES = BN * 128;
  constexpr int A_OFF = 0, W_OFF = LAND_BYTES, STAGE = W_OFF + 2 * WP_BYTES;
  constexpr int ACC_COLS = 2 * BN;
  uint64_t* full = (uint64_t*)(smem + TH_STAGES * STAGE);
  uint64_t* empty = full + TH_STAGES;
  uint64_t* ready = empty + TH_STAGES;
  uint64_t* tmem_full = ready + TH_STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
  float* epi_tiles = (float*)(smem + TH_STAGES * STAGE + 256);  // [8 warps][32][33]

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmW);
    for (int s = 0; s < TH_STAGES; s++) { tc::mbar_init(full + s, 1); tc::mbar_init(empty + s, 1); tc::mbar_init(ready + s, 32 * TH_SPLIT_WARPS); }
    for (int a = 0; a < 2; a++) { tc::mbar_init(tmem_full + a, 1); tc::mbar_init(tmem_empty + a, 256); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 2 * ACC_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int KB = g.K / TH_BKE, n_tiles = g.N / BN, total = m_tiles * n_tiles;

  auto tile_info = [&](int tile, int& m_tile, int& n0, int& z, int& row0, int& nrows) -> bool {
    m_tile = tile / n_tiles; n0 = (tile % n_tiles) * BN;
    z = m_tile / g.tiles_per_slot; row0 = (m_tile % g.tiles_per_slot) * TC_BM;
    if (g.skip && g.skip[z >> g.skip_shift]) return false;
    nrows = g.counts ? g.counts[z] : (g.tiles_per_slot * TC_BM);
    return row0 < nrows;
  };

  if (warp == 0) {
    if (lane == 0) {
      int c = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        int m_tile, n0, z, row0, nrows;
        if (!tile_info(tile, m_tile, n0, z, row0, nrows)) continue;
        for (int kb = 0; kb < KB; kb++, c++) {
          const int s = c % TH_STAGES, ph = (c / TH_STAGES) & 1;
          tc::mbar_wait(empty + s, ph ^ 1);
          tc::mbar_expect_tx(full + s, LAND_BYTES + 2 * WP_BYTES);
          uint8_t* st = smem + s * STAGE;
          if (APL) {   // fp16 planes: A_hi and A_lo tiles as they are
            tc::tma_load_2d(st, &tmA, full + s, kb * TH_BKE, m_tile * TC_BM);
            tc::tma_load_2d(st + AP_BYTES, &tmA, full + s, kb * TH_BKE, (int)(g.a_plane_rows + (long long)m_tile * TC_BM));
          } else {
            tc::tma_load_2d(st, &tmA, full + s, kb * TH_BKE, m_tile * TC_BM);
            tc::tma_load_2d(st + LAND_BYTES / 2, &tmA, full + s, kb * TH_BKE + 32, m_tile * TC_BM);
          }
          tc::tma_load_2d(st + W_OFF, &tmW, full + s, kb * TH_BKE, n0);
          tc::tma_load_2d(st + W_OFF + WP_BYTES, &tmW, full + s, kb * TH_BKE, (int)(g.w_plane_rows + n0));
        }
      }
    }
  } else if (warp == 1) {
    const bool leader = tc::elect_one();
    constexpr uint32_t idesc = tc::make_idesc(tc::FMT_F16, TC_BM, BN), idesc2 = tc::make_idesc(tc::FMT_F16, TC_BM, 2 * BN);
    int c = 0, i = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      int m_tile, n0, z, row0, nrows;
      if (!tile_info(tile, m_tile, n0, z, row0, nrows)) continue;
      const int acc = i & 1;
      tc::mbar_wait(tmem_empty + acc, ((i >> 1) & 1) ^ 1);
      tc::fence_after_sync();
      const uint32_t d_main = tmem_base + acc * ACC_COLS, d_cross = d_main + BN;
      for (int kb = 0; kb < KB; kb++, c++) {
        const int s = c % TH_STAGES, ph = (c / TH_STAGES) & 1;
        tc::mbar_wait(APL ? full + s : ready + s, ph);   // splitters done (they waited for the TMA bytes: W planes have landed too)
        tc::fence_after_sync();
        const uint32_t a_addr = tc::smem_u32(smem + s * STAGE + A_OFF), b_addr = tc::smem_u32(smem + s * STAGE + W_OFF);
#pragma unroll
        for (int k = 0; k < TH_BKE / 16; k++) {
          uint64_t ad = tc::make_smem_desc_sw128(a_addr + k * 32), bd = tc::make_smem_desc_sw128(b_addr + k * 32);
          uint64_t adl = tc::make_smem_desc_sw128(a_addr + AP_BYTES + k * 32);
          if (leader) {
            tc::mma_f16(d_main, ad, bd, idesc2, (kb | k) ? 1u : 0u);   // A_hi x [W_hi | W_lo] -> [main | cross]
            tc::mma_f16(d_cross, adl, bd, idesc, 1u);                  // A_lo x W_hi -> cross
          }
        }
        if (leader) tc::mma_commit(empty + s);
        __syncwarp();
      }
      if (leader) tc::mma_commit(tmem_full + acc);
      __syncwarp();
      i++;
    }
  } else if (warp < 2 + TH_SPLIT_WARPS) {
    // splitters: thread = row r of the tile: 64 fp32 (two landed 128-byte rows, SWIZZLE_128B: 16-byte chunk c at c ^ (r & 7))
    // -> 64 hi halves (row r of A_hi, 128 B: element chunk cc at cc ^ (r & 7)) + 64 lo halves, written over the same bytes
    const int r = threadIdx.x - 64;  // 0 .. 127
    const int sw = r & 7;
    int c = 0;
    for (int tile = blockIdx.x; !APL && tile < total; tile += gridDim.x) {
      int m_tile, n0, z, row0, nrows;
      if (!tile_info(tile, m_tile, n0, z, row0, nrows)) continue;
      for (int kb = 0; kb < KB; kb++, c++) {
        const int s = c % TH_STAGES, ph = (c / TH_STAGES) & 1;
        tc::mbar_wait(full + s, ph);
        uint8_t* row_hi = smem + s * STAGE + r * 128;          // box 0 row r -> A_hi row r
        uint8_t* row_lo = row_hi + AP_BYTES;                   // box 1 row r -> A_lo row r
        uint4 v[16];
        if (!(ablate & 2)) {
#pragma unroll
        for (int ch = 0; ch < 8; ch++) {
          v[ch] = *reinterpret_cast<const uint4*>(row_hi + ((ch ^ sw) << 4));        // elements 4 ch .. 4 ch + 3
          v[8 + ch] = *reinterpret_cast<const uint4*>(row_lo + ((ch ^ sw) << 4));    // elements 32 + 4 ch ..
        }
#pragma unroll
        for (int cc = 0; cc < 8; cc++) {   // output chunk cc = elements 8 cc .. 8 cc + 7 = input chunks 2 cc, 2 cc + 1
          const uint4 a = v[2 * cc], b = v[2 * cc + 1];
          const float x[8] = {__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w),
                              __uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z), __uint_as_float(b.w)};
          __align__(16) __half2 h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; e++) split2x2(x[2 * e], x[2 * e + 1], h[e], l[e]);
          *reinterpret_cast<uint4*>(row_hi + ((cc ^ sw) << 4)) = *reinterpret_cast<const uint4*>(h);
          *reinterpret_cast<uint4*>(row_lo + ((cc ^ sw) << 4)) = *reinterpret_cast<const uint4*>(l);
        }
        }
        tc::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
        tc::mbar_arrive(ready + s);
      }
    }
  } else {
    const int ew = warp - 2 - TH_SPLIT_WARPS;   // 0..7
    const int q = warp % 4;          // TMEM sub-partition this warp may read: lanes [32q, 32q+32)
    const int half = ew / 4;         // two warps per sub-partition: columns [0, BN/2) and [BN/2, BN)
    float* T = epi_tiles + ew * 32 * 33;
    int i = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      int m_tile, n0, z, row0, nrows;
      if (!tile_info(tile, m_tile, n0, z, row0, nrows)) continue;
      const int acc = i & 1;
      tc::mbar_wait(tmem_full + acc, (i >> 1) & 1);
      tc::fence_after_sync();
      const int row_base = row0 + q * 32;
      const uint32_t lane_addr = tmem_base + acc * ACC_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
        float v[32];
        tc::tmem_ld32(lane_addr + c0, v);
        {
          float t[32];
          tc::tmem_ld32(lane_addr + BN + c0, t);
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = fmaf(t[j], PLANE_LO_INV, v[j]);
        }
        if (c0 + 32 >= (half + 1) * (BN / 2)) {
          tc::fence_before_sync();
          tc::mbar_arrive(tmem_empty + acc);
        }
        if (ablate & 4) continue;
        if (epi.rowwise(z, row_base + lane, row_base + lane < nrows, n0 + c0, v)) continue;
        const int rmax = min(32, nrows - row_base);  // warp-uniform
        if constexpr (tc_has_quad<Epi>::value) {
          if (epi.quad_ok()) {
            float4* T4 = reinterpret_cast<float4*>(T);   // [32 rows][8 quads], quad q of row r at q ^ (r & 7)
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; j++) T4[lane * 8 + (j ^ (lane & 7))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            __syncwarp();
            const int cq = lane & 7, ro = lane >> 3, col = n0 + c0 + 4 * cq;
            const float4 cc = epi.quad_col(col);
            float4 pre[8];
#pragma unroll
            for (int k = 0; k < 8; k++) pre[k] = (4 * k + ro < rmax) ? epi.quad_prefetch(z, row_base + 4 * k + ro, col) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 8; k++) {
              const int r = 4 * k + ro;
              if (r < rmax) epi.quad(z, row_base + r, col, T4[r * 8 + (cq ^ (r & 7))], cc, pre[k]);
            }
            continue;
          }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; j++) T[lane * 33 + j] = v[j];
        __syncwarp();
        if constexpr (tc_has_pair<Epi>::value) {
          // lane = (row parity, column pair): rows 2k + (lane >> 4), columns 2 (lane & 15), +1.  T reads are conflict-free
          // (row stride 33 words: even rows hit the even banks, odd rows the odd ones)
          const int cp = 2 * (lane & 15), ro = lane >> 4;
          float4 pre[16];
#pragma unroll
          for (int k = 0; k < 16; k++)
            pre[k] = (2 * k + ro < rmax) ? epi.pair_prefetch(z, row_base + 2 * k + ro, n0 + c0 + cp) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int k = 0; k < 16; k++) {
            const int r = 2 * k + ro;
            if (r < rmax) epi.pair(z, row_base + r, n0 + c0 + cp, T[r * 33 + cp], T[r * 33 + cp + 1], pre[k]);
          }
        } else {
          float2 pre[32];
#pragma unroll
          for (int r = 0; r < 32; r++) pre[r] = (r < rmax) ? epi.prefetch(z, row_base + r, n0 + c0 + lane) : make_float2(0.f, 0.f);
#pragma unroll
          for (int r = 0; r < 32; r++)
            if (r < rmax) epi.elem(z, row_base + r, n0 + c0 + lane, T[r * 33 + lane], pre[r]);
        }
      }
      i++;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 2 * ACC_COLS);
}

template <int BN, class Epi>
static inline int launch_tc_gemm_f16(const float* A, long long rows_total, int lda, TcGemmArgs g, Epi epi, cudaStream_t st) {
  CUtensorMap tmA, tmW;
  if (g.a_planes) {
    if (int e = tc_make_map_2d_f16_ld(&tmA, g.a_planes, (uint64_t)(2 * g.a_plane_rows), (uint64_t)g.K, (uint64_t)g.a_plane_ld, TH_BKE, TC_BM)) return e;
  } else {
    if (int e = tc_make_map_2d_f32(&tmA, A, (uint64_t)rows_total, (uint64_t)g.K, (uint64_t)lda, 32, TC_BM)) return e;
  }
  if (int e = tc_make_map_2d_f16(&tmW, g.w_planes, (uint64_t)(2 * g.w_plane_rows), (uint64_t)g.K, TH_BKE, BN)) return e;
  constexpr size_t smem = tc_gemm_f16_smem_bytes<BN>();
  const int num_sms = imw_num_sms();
  const int m_tiles = (int)(rows_total / TC_BM), total = m_tiles * (g.N / BN);
  dim3 grid((unsigned)(total < num_sms ? total : num_sms));
  if (g.a_planes) {
    IMW_SMEM_ATTR_ONCE((tc_gemm_f16_kernel<BN, Epi, true>), smem);
    tc_gemm_f16_kernel<BN, Epi, true><<<grid, TH_THREADS, smem, st>>>(tmA, tmW, g, epi, m_tiles, tc_gemm_ablate());
  } else {
    IMW_SMEM_ATTR_ONCE((tc_gemm_f16_kernel<BN, Epi, false>), smem);
    tc_gemm_f16_kernel<BN, Epi, false><<<grid, TH_THREADS, smem, st>>>(tmA, tmW, g, epi, m_tiles, tc_gemm_ablate());
  }
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}

// A: [slots*cap][lda] fp32 (rows_total x K view), W: [w_rows][K] fp32 (followed by its lo plane when g.wlo_rows > 0).
template <int BN, int SPLIT, bool WLO, class Epi>
static inline int launch_tc_gemm_t(const CUtensorMap& tmA, const CUtensorMap& tmW, long long rows_total, TcGemmArgs g, Epi epi, cudaStream_t st) {
  constexpr size_t smem = tc_gemm_smem_bytes<BN, SPLIT>();
  IMW_SMEM_ATTR_ONCE((tc_gemm_tf32_kernel<BN, SPLIT, WLO, Epi>), smem);
  const int num_sms = imw_num_sms();
  const int m_tiles = (int)(rows_total / TC_BM), total = m_tiles * (g.N / BN);
  dim3 grid((unsigned)(total < num_sms ? total : num_sms));
  tc_gemm_tf32_kernel<BN, SPLIT, WLO, Epi><<<grid, TC_THREADS, smem, st>>>(tmA, tmW, g, epi, m_tiles);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}

template <int BN, int SPLIT, class Epi>
static inline int launch_tc_gemm(const float* A, long long rows_total, int lda, const float* W, long long w_rows, TcGemmArgs g,
                                 Epi epi, cudaStream_t st) {
  if (SPLIT == 3 && g.w_planes && !g.pair_product && !g.wsel_minus1 && g.K % TH_BKE == 0)
    return launch_tc_gemm_f16<BN, Epi>(A, rows_total, lda, g, epi, st);
  CUtensorMap tmA, tmW;
  const bool wlo = SPLIT == 3 && g.wlo_rows > 0 && !g.pair_product;
  if (!wlo) g.wlo_rows = 0;
  if (int e = tc_make_map_2d_f32(&tmA, A, (uint64_t)rows_total, (uint64_t)g.K, (uint64_t)lda, TC_BK, TC_BM)) return e;
  if (int e = tc_make_map_2d_f32(&tmW, W, (uint64_t)(wlo ? g.wlo_rows + w_rows : w_rows), (uint64_t)g.K, (uint64_t)g.K, TC_BK, BN)) return e;
  if (SPLIT == 3 && wlo) return launch_tc_gemm_t<BN, SPLIT, (SPLIT == 3), Epi>(tmA, tmW, rows_total, g, epi, st);
  return launch_tc_gemm_t<BN, SPLIT, false, Epi>(tmA, tmW, rows_total, g, epi, st);
}
