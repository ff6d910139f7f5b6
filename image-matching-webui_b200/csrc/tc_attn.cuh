// tcgen05 flash attention for LightGlue / SuperGlue (head dim 64, fp32-equivalent via split-fp16 operands).
//
//   ctx[z][row][h*64 + d] = softmax_j(scale * <q_row, k_j>) v_j      self: k,v of slot z; cross: of slot z^1
//
// Operands arrive pre-split from the projection epilogues as two fp16 planes (split_planes.cuh: hi = fp16(x),
// lo = fp16((x - hi) 2^11)):  q, k  [plane][slot][head][cap][64],  v transposed  [plane][slot][head][64][cap]  (keys contiguous,
// so that V^T is a plain K-major B operand).  Per CTA: one (slot, head, 128-row q tile); key tiles of 64.
//
// Round-2 structure (the round-1 kernel ran 3xTF32 with 16 softmax warps that each owned a QUARTER of a row: one 512-thread
// barrier and a shared-memory max exchange per key tile, O read back from TMEM and rescaled in registers every tile -- 2.2 us
// per key tile against 0.8 us of MMA time, tensor pipe 40 %):
//   * TWO independent softmax streams: stream A owns the even key tiles, stream B the odd ones; each has its own S buffer, P
//     buffer and O accumulator, so S_{j+1} / P_{j+1} are produced while P_j V_j runs and nothing is exchanged between threads
//     until the end.  Two threads (two warps of the same TMEM sub-partition) share one q row of their stream's tile: each reads,
//     exponentiates, splits and stores its own 32 columns of P; the row maximum is the max of the two half maxima, exchanged
//     through shared memory behind a 64-thread named barrier of the two warps (identical in both threads, so the lazy-rescale
//     decision agrees).
//   * O stays in TMEM and is accumulated by the tensor core across the stream's tiles (use_acc); it is rescaled in place
//     (tcgen05.ld / st) only when a row's running reference maximum has to move by more than 2^8 (lazy rescaling: P <= 256
//     stays far inside fp16 / fp32 range and the final division by the row sum, taken with the same reference, cancels it).
//   * kind::f16 MMAs (K = 16 per instruction): S = Q_hi x [K_hi | K_lo] + Q_lo x K_hi, O = P_hi x [V_hi | V_lo] + P_lo x V_hi,
//     scaled cross terms in their own accumulator columns.  Half the tensor time and half the operand bytes of 3xTF32.
//   * the two streams are merged once: O = (O_a w_a + O_b w_b) / (l_a w_a + l_b w_b), w = 2^(m - max(m_a, m_b)).
//
//   warp 0      TMA producer: Q once, K / V^T tiles through two independent 3-deep rings
//   warp 1      MMA issuer (whichever of S_js / P_jp V_jp has its operands ready is issued next)
//   warps 2-9   softmax stream A, warps 10-17 stream B (two warps per TMEM sub-partition: columns 0..31 / 32..63 of P)
#pragma once
#include "common.cuh"
#include "split_planes.cuh"
#include "tc_common.cuh"

struct TcAttnArgs {
  float* ctx;            // [slots][cap][256]
  const int* counts;     // [slots]
  const int* skip;       // [pairs]
  int cap, slots;
  float scale;
  int cross;
  long long plane_rows_qk;  // slots*4*cap : row offset of the lo plane in the q/k tensor maps
  long long plane_rows_vt;  // slots*4*64
};

constexpr int TA_BQ = 128, TA_BKV = 64, TA_NK = 3, TA_NV = 3, TA_SM_THREADS = 512, TA_THREADS = 64 + TA_SM_THREADS;
constexpr int TA_Q_BYTES = 2 * TA_BQ * 128;         // hi | lo x [128 rows x 64 f16]   = 32 KB
constexpr int TA_K_BYTES = 2 * TA_BKV * 128;        // hi | lo x [64 keys x 64 f16]    = 16 KB per stage
constexpr int TA_V_BYTES = 2 * 64 * 128;            // hi | lo x [64 d x 64 keys f16]  = 16 KB per stage
constexpr int TA_P_BYTES = 2 * TA_BQ * 128;         // hi | lo x [128 rows x 64 keys]  = 32 KB per stream
constexpr size_t TA_SMEM = TA_Q_BYTES + TA_NK * TA_K_BYTES + TA_NV * TA_V_BYTES + 2 * TA_P_BYTES + 1024 + 256 + (6 + 8) * 128 * sizeof(float);
// TMEM columns: S_a [0,128) = main 64 | cross 64, S_b [128,256), O_a [256,384) = main 64 | cross 64, O_b [384,512)
constexpr int TA_TMEM_COLS = 512, TA_S_COL = 0, TA_O_COL = 256;
constexpr float TA_LAZY = 8.f;   // log2 units: the reference maximum of a row moves only when it would grow by more than this

namespace tc {
// 2^x on the SFU (ex2.approx.ftz: relative error 2^-22.5; -inf -> 0)
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 32 lanes x 32 consecutive fp32 columns, registers -> TMEM
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
}  // namespace tc

static __global__ void __launch_bounds__(TA_THREADS, 1)
tc_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, TcAttnArgs g) {
  const int z = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * TA_BQ;
  if (g.skip[z >> 1]) return;
  const int nq = g.counts[z], zk = g.cross ? (z ^ 1) : z, nk = g.counts[zk];
  if (q0 >= nq) return;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (nk == 0) {  // empty key set -> zeros (lightglue.py:113-114)
    for (int i = threadIdx.x; i < TA_BQ * 16; i += TA_THREADS) {
      int r = q0 + i / 16, c = (i % 16) * 4;
      if (r < nq) *reinterpret_cast<float4*>(g.ctx + ((long long)z * g.cap + r) * 256 + head * 64 + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  extern __shared__ uint8_t ta_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)ta_smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                          // [hi | lo][128 rows][128 B]
  uint8_t* sK = sQ + TA_Q_BYTES;               // [stage][hi | lo][64 keys][128 B]     (hi | lo adjacent: one N = 128 operand)
  uint8_t* sV = sK + TA_NK * TA_K_BYTES;       // [stage][hi | lo][64 d rows][128 B]
  uint8_t* sP = sV + TA_NV * TA_V_BYTES;       // [stream][hi | lo][128 rows][128 B]
  uint64_t* bars = (uint64_t*)(sP + 2 * TA_P_BYTES);
  uint64_t *q_full = bars, *k_full = bars + 1 /*[3]*/, *k_empty = bars + 4 /*[3]*/, *v_full = bars + 7 /*[3]*/, *v_empty = bars + 10 /*[3]*/,
           *s_full = bars + 13 /*[2]*/, *s_free = bars + 15 /*[2]*/, *p_full = bars + 17 /*[2]*/, *o_done = bars + 19 /*[2]*/;
  uint32_t* tmem_slot = (uint32_t*)(bars + 21);
  float* xchg = (float*)((uint8_t*)bars + 256);   // reference maxima [2][128] + partial row sums [2][2][128] for the final merge
  float* xhalf = xchg + 768;                      // [stream][tile parity][half][128] half-row maxima

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmQ); tc::tma_prefetch_desc(&tmK); tc::tma_prefetch_desc(&tmV);
    tc::mbar_init(q_full, 1);
    for (int i = 0; i < 3; i++) { tc::mbar_init(k_full + i, 1); tc::mbar_init(k_empty + i, 1); tc::mbar_init(v_full + i, 1); tc::mbar_init(v_empty + i, 1); }
    for (int i = 0; i < 2; i++) { tc::mbar_init(s_full + i, 1); tc::mbar_init(s_free + i, 256); tc::mbar_init(p_full + i, 256); tc::mbar_init(o_done + i, 1); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, TA_TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int T = (nk + TA_BKV - 1) / TA_BKV;
  const int q_row = (z * 4 + head) * g.cap + q0;          // row in the [2][slots*4*cap][64] q/k tensors
  const int k_row0 = (zk * 4 + head) * g.cap;
  const int v_row = (zk * 4 + head) * 64;                 // row in the [2][slots*4*64][cap] v^T tensor

  if (warp == 0) {
    if (lane == 0) {
      tc::mbar_expect_tx(q_full, TA_Q_BYTES);
      tc::tma_load_2d(sQ, &tmQ, q_full, 0, q_row);
      tc::tma_load_2d(sQ + TA_Q_BYTES / 2, &tmQ, q_full, 0, (int)g.plane_rows_qk + q_row);
      // K and V rings are refilled independently, each stage as soon as its own consumer MMAs have completed.  A barrier cannot
      // run ahead of the phase polled here: stage reuse n+1 is only requested after reuse n was observed empty.
      int jk = 0, jv = 0;
      uint32_t spins = 0;
      while (jk < T || jv < T) {
        bool did = false;
        if (jk < T && (jk < TA_NK || tc::mbar_try_wait(k_empty + jk % TA_NK, ((jk / TA_NK) - 1) & 1))) {
          uint8_t* dst = sK + (jk % TA_NK) * TA_K_BYTES;
          tc::mbar_expect_tx(k_full + jk % TA_NK, TA_K_BYTES);
          tc::tma_load_2d(dst, &tmK, k_full + jk % TA_NK, 0, k_row0 + jk * TA_BKV);
          tc::tma_load_2d(dst + TA_K_BYTES / 2, &tmK, k_full + jk % TA_NK, 0, (int)g.plane_rows_qk + k_row0 + jk * TA_BKV);
          jk++; did = true;
        }
        if (jv < T && (jv < TA_NV || tc::mbar_try_wait(v_empty + jv % TA_NV, ((jv / TA_NV) - 1) & 1))) {
          uint8_t* dst = sV + (jv % TA_NV) * TA_V_BYTES;
          tc::mbar_expect_tx(v_full + jv % TA_NV, TA_V_BYTES);
          tc::tma_load_2d(dst, &tmV, v_full + jv % TA_NV, jv * TA_BKV, v_row);
          tc::tma_load_2d(dst + TA_V_BYTES / 2, &tmV, v_full + jv % TA_NV, jv * TA_BKV, (int)g.plane_rows_vt + v_row);
          jv++; did = true;
        }
        if (did) spins = 0;
        else if (++spins > (1u << 24)) { printf("tc_attn producer timeout block (%d,%d,%d) jk %d jv %d\n", blockIdx.x, blockIdx.y, blockIdx.z, jk, jv); __trap(); }
      }
    }
  } else if (warp == 1) {
    const bool leader = tc::elect_one();
    constexpr uint32_t idesc64 = tc::make_idesc(tc::FMT_F16, 128, 64), idesc128 = tc::make_idesc(tc::FMT_F16, 128, 128);
    // operand descriptors = one base descriptor per buffer + (byte offset >> 4) in the 14-bit address field: a 64-bit add per operand
    // instead of a rebuild (the issuing thread's instruction stream sits on the critical path of every key tile)
    const uint64_t dQ = tc::make_smem_desc_sw128(tc::smem_u32(sQ)), dK = tc::make_smem_desc_sw128(tc::smem_u32(sK)),
                   dV = tc::make_smem_desc_sw128(tc::smem_u32(sV)), dP = tc::make_smem_desc_sw128(tc::smem_u32(sP));
    auto issue_S = [&](int j) {  // S_j = Q K_j^T : main = hi*hi, cross = (hi*lo + lo*hi) 2^11
      tc::fence_after_sync();
      const uint32_t d_main = tmem_base + TA_S_COL + (j & 1) * 128, d_cross = d_main + 64;
      const uint64_t kb = dK + (uint64_t)(((j % TA_NK) * TA_K_BYTES) >> 4);
#pragma unroll
      for (int ks = 0; ks < 4; ks++) {
        const uint64_t qh = dQ + (uint64_t)(ks * 2), ql = qh + (uint64_t)((TA_Q_BYTES / 2) >> 4);
        const uint64_t kh = kb + (uint64_t)(ks * 2);   // [K_hi | K_lo]: 128 adjacent rows
        if (leader) {
          tc::mma_f16(d_main, qh, kh, idesc128, ks ? 1u : 0u);   // Q_hi x [K_hi | K_lo] -> [main | cross]
          tc::mma_f16(d_cross, ql, kh, idesc64, 1u);             // Q_lo x K_hi -> cross
        }
      }
      if (leader) {
        tc::mma_commit(k_empty + j % TA_NK);
        tc::mma_commit(s_full + (j & 1));
      }
      __syncwarp();
    };
    auto issue_PV = [&](int j) {  // O_stream += P_j V_j  (K = 64 keys)
      tc::fence_after_sync();
      const int st = j & 1;
      const uint32_t d_main = tmem_base + TA_O_COL + st * 128, d_cross = d_main + 64;
      const uint64_t vb = dV + (uint64_t)(((j % TA_NV) * TA_V_BYTES) >> 4), pb = dP + (uint64_t)((st * TA_P_BYTES) >> 4);
      const uint32_t first = (j < 2) ? 0u : 1u;   // the stream's first tile overwrites its accumulator
#pragma unroll
      for (int ks = 0; ks < 4; ks++) {
        const uint64_t ph = pb + (uint64_t)(ks * 2), pl = ph + (uint64_t)((TA_P_BYTES / 2) >> 4);
        const uint64_t vh = vb + (uint64_t)(ks * 2);   // [V_hi | V_lo]
        if (leader) {
          tc::mma_f16(d_main, ph, vh, idesc128, ks ? 1u : first);
          tc::mma_f16(d_cross, pl, vh, idesc64, 1u);
        }
      }
      if (leader) {
        tc::mma_commit(v_empty + j % TA_NV);
        tc::mma_commit(o_done + st);
      }
      __syncwarp();
    };
    tc::mbar_wait(q_full, 0);
    int js = 0, jp = 0;
    uint32_t spins = 0;
    const int lead_lane = __ffs(__ballot_sync(0xffffffffu, leader)) - 1;
    // the issuing lane polls, the warp follows its verdict.  S_js needs K_js landed and its S buffer drained by the softmax of
    // tile js-2; P_jp V_jp needs P_jp written (which implies the stream's previous P V has completed) and V_jp landed.
    auto ready_S = [&]() -> bool {
      unsigned r = 0;
      if (leader) r = (tc::mbar_try_wait(k_full + js % TA_NK, (js / TA_NK) & 1) &&
                       (js < 2 || tc::mbar_try_wait(s_free + (js & 1), ((js >> 1) - 1) & 1))) ? 1u : 0u;
      return __shfl_sync(0xffffffffu, r, lead_lane) != 0;
    };
    auto ready_PV = [&]() -> bool {
      unsigned r = 0;
      if (leader) r = (tc::mbar_try_wait(p_full + (jp & 1), (jp >> 1) & 1) && tc::mbar_try_wait(v_full + jp % TA_NV, (jp / TA_NV) & 1)) ? 1u : 0u;
      return __shfl_sync(0xffffffffu, r, lead_lane) != 0;
    };
    while (jp < T) {
      bool did = false;
      if (js < T && js <= jp + 2 && ready_S()) { issue_S(js); js++; did = true; }
      if (jp < js && ready_PV()) { issue_PV(jp); jp++; did = true; }
      if (did) spins = 0;
      else if (++spins > (1u << 24)) { printf("tc_attn MMA scheduler timeout block (%d,%d,%d) js %d jp %d\n", blockIdx.x, blockIdx.y, blockIdx.z, js, jp); __trap(); }
    }
  } else {
    const int st = (warp - 2) / 8;                       // stream: key tiles j = st, st + 2, ...
    const int hf = ((warp - 2) / 4) & 1;                 // this thread's half of the tile's 64 columns
    const int q = warp % 4, r = q * 32 + lane;           // TMEM lane = q row inside the tile
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float c2 = g.scale * 1.4426950408889634f;      // softmax in base 2: p = 2^(s c2 - m)
    float m = -INFINITY, l = 0.f;                        // reference maximum (log2 units) and this thread's part of the row sum
    int n = 0;                                           // tiles of this stream processed so far
    uint8_t* pP = sP + st * TA_P_BYTES + r * 128;
    for (int j = st; j < T; j += 2, n++) {
      tc::mbar_wait(s_full + st, n & 1);
      tc::fence_after_sync();
      const int kv0 = j * TA_BKV;
      float s[32];                                       // own half of the row: raw scores
      float mx = -INFINITY;
      {
        const uint32_t a = lane_addr + TA_S_COL + st * 128 + hf * 32;
        uint32_t mn[2][16], cr[2][16];                   // four loads in flight, one wait
        tc::tmem_ld16_async(a, mn[0]);
        tc::tmem_ld16_async(a + 64, cr[0]);
        tc::tmem_ld16_async(a + 16, mn[1]);
        tc::tmem_ld16_async(a + 64 + 16, cr[1]);
        tc::tmem_wait_ld();
        tc::tmem_ld_fence(mn[0]); tc::tmem_ld_fence(cr[0]); tc::tmem_ld_fence(mn[1]); tc::tmem_ld_fence(cr[1]);
        tc::fence_before_sync();
        tc::mbar_arrive(s_free + st);                    // S buffer may take tile j + 2
        if (kv0 + TA_BKV <= nk) {                        // full tile (warp-uniform): no masking
#pragma unroll
          for (int c = 0; c < 32; c++) {
            s[c] = fmaf(__uint_as_float(cr[c >> 4][c & 15]), PLANE_LO_INV, __uint_as_float(mn[c >> 4][c & 15]));
            mx = fmaxf(mx, s[c]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 32; c++) {
            const float v = fmaf(__uint_as_float(cr[c >> 4][c & 15]), PLANE_LO_INV, __uint_as_float(mn[c >> 4][c & 15]));
            s[c] = (kv0 + hf * 32 + c < nk) ? v : -INFINITY;
            mx = fmaxf(mx, s[c]);
          }
        }
      }
      // row maximum = max of the two halves: exchanged with the partner thread (same row, warp +-4) through shared memory and a
      // 64-thread named barrier of the two warps; both threads then hold the identical value (max is commutative)
      {
        float* xb = xhalf + (st * 2 + (n & 1)) * 256;    // double-buffered by tile parity: the partner may run one tile ahead
        xb[hf * 128 + r] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(2 + st * 4 + q) : "memory");
        mx = fmaxf(mx, xb[(hf ^ 1) * 128 + r]) * c2;     // log2 units (scale > 0 commutes with max)
      }
      // lazy reference: move it only when this tile's maximum exceeds it by more than TA_LAZY (always on the first tile)
      const bool move = mx > m + TA_LAZY;                // false for NaN rows (rows beyond the count read unwritten memory)
      const float m_new = move ? mx : m;
      const float alpha = move ? tc::ex2(m - m_new) : 1.f; // 0 on the first tile (m = -inf)
      if (n > 0) {
        tc::mbar_wait(o_done + st, (n - 1) & 1);         // P_{j-2} V_{j-2} complete: O is stable, the P buffer is free
        tc::fence_after_sync();
        if (__any_sync(0xffffffffu, move)) {             // rescale this warp's 32 rows of O in place (rare after the first tiles):
#pragma unroll 1
          for (int cc = 0; cc < 2; cc++) {               // half 0 rescales the main accumulator columns, half 1 the cross terms
            float o[32];
            const uint32_t oa = lane_addr + TA_O_COL + st * 128 + hf * 64 + cc * 32;
            tc::tmem_ld32(oa, o);
#pragma unroll
            for (int c = 0; c < 32; c++) o[c] *= alpha;
            tc::tmem_st32(oa, o);
          }
        }
      }
      l *= alpha;
      m = m_new;
      float ps = 0.f;
#pragma unroll
      for (int c = 0; c < 32; c++) { s[c] = tc::ex2(fmaf(s[c], c2, -m)); ps += s[c]; }   // masked columns: -inf -> 0
      l += ps;
      // P_j -> shared memory as the two fp16 operand planes, K-major rows of 128 B with the 128B swizzle (16-byte chunk ch of
      // row r lives at chunk ch ^ (r & 7)); this thread's 32 keys = chunks 4 hf .. 4 hf + 3
#pragma unroll
      for (int ch = 0; ch < 4; ch++) {
        uint4 hv, lv;
        uint32_t* hp = reinterpret_cast<uint32_t*>(&hv);
        uint32_t* lp = reinterpret_cast<uint32_t*>(&lv);
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const float p0 = s[ch * 8 + 2 * e], p1 = s[ch * 8 + 2 * e + 1];
          const __half2 h2v = __floats2half2_rn(p0, p1);
          const float2 hf2 = __half22float2(h2v);
          const __half2 l2v = __floats2half2_rn((p0 - hf2.x) * PLANE_LO_SCALE, (p1 - hf2.y) * PLANE_LO_SCALE);
          hp[e] = *reinterpret_cast<const uint32_t*>(&h2v);
          lp[e] = *reinterpret_cast<const uint32_t*>(&l2v);
        }
        const int pos = ((4 * hf + ch) ^ (r & 7)) * 16;
        *reinterpret_cast<uint4*>(pP + pos) = hv;
        *reinterpret_cast<uint4*>(pP + TA_P_BYTES / 2 + pos) = lv;
      }
      tc::fence_before_sync();                           // orders the tcgen05.st of the rescale before the MMA that accumulates on top
      tc::fence_proxy_async();
      tc::mbar_arrive(p_full + st);
    }
    // the stream's last P V
    if (n > 0) {
      tc::mbar_wait(o_done + st, (n - 1) & 1);
      tc::fence_after_sync();
    }
    // merge: every thread publishes (m, partial l); thread (stream st, half hf) writes output dims [32 st + 16 hf, +16) of its row
    float* xm = xchg;            // [stream][128]     reference maxima (identical in both halves)
    float* xl = xchg + 256;      // [stream][half][128] partial row sums
    if (hf == 0) xm[st * 128 + r] = m;
    xl[(st * 2 + hf) * 128 + r] = l;
    asm volatile("bar.sync 1, %0;" ::"n"(TA_SM_THREADS) : "memory");
    const float ma = xm[r], mb = xm[128 + r];
    const float la = xl[r] + xl[128 + r], lb = xl[256 + r] + xl[384 + r];
    const float mm = fmaxf(ma, mb);                      // stream A always has a tile (T >= 1): finite for valid rows
    const float wa = tc::ex2(ma - mm), wb = (lb > 0.f) ? tc::ex2(mb - mm) : 0.f;
    const float inv = 1.f / (la * wa + lb * wb);
    const int row = q0 + r;
    {
      float acc[16], t[16];
      const int d0 = st * 32 + hf * 16;
      const uint32_t oa = lane_addr + TA_O_COL + d0;                 // this thread's 16 output dims
      tc::tmem_ld16(oa, acc);
      tc::tmem_ld16(oa + 64, t);
#pragma unroll
      for (int c = 0; c < 16; c++) acc[c] = fmaf(t[c], PLANE_LO_INV, acc[c]) * wa;
      if (T > 1) {                                                   // stream B ran: its accumulator holds data
        float ob[16];
        tc::tmem_ld16(oa + 128, ob);
        tc::tmem_ld16(oa + 128 + 64, t);
#pragma unroll
        for (int c = 0; c < 16; c++) acc[c] = fmaf(fmaf(t[c], PLANE_LO_INV, ob[c]), wb, acc[c]);
      }
      if (row < nq) {
        float4* o = reinterpret_cast<float4*>(g.ctx + ((long long)z * g.cap + row) * 256 + head * 64 + d0);
#pragma unroll
        for (int c = 0; c < 4; c++) o[c] = make_float4(acc[4 * c] * inv, acc[4 * c + 1] * inv, acc[4 * c + 2] * inv, acc[4 * c + 3] * inv);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, TA_TMEM_COLS);
}

// q_planes / k_planes: [2][slots*4*cap][64] fp16, vt_planes: [2][slots*4*64][cap] fp16
static inline int launch_tc_attn(const void* q_planes, const void* k_planes, const void* vt_planes, TcAttnArgs g, cudaStream_t st) {
  CUtensorMap tmQ, tmK, tmV;
  const long long rows_qk = 2 * g.plane_rows_qk, rows_vt = 2 * g.plane_rows_vt;
  if (int e = tc_make_map_2d_f16(&tmQ, q_planes, (uint64_t)rows_qk, 64, 64, TA_BQ)) return e;
  if (int e = tc_make_map_2d_f16(&tmK, k_planes, (uint64_t)rows_qk, 64, 64, TA_BKV)) return e;
  if (int e = tc_make_map_2d_f16(&tmV, vt_planes, (uint64_t)rows_vt, (uint64_t)g.cap, 64, 64)) return e;
  IMW_SMEM_ATTR_ONCE(tc_attn_kernel, TA_SMEM);
  dim3 grid(g.cap / TA_BQ, 4, g.slots);
  tc_attn_kernel<<<grid, TA_THREADS, TA_SMEM, st>>>(tmQ, tmK, tmV, g);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}
