// tcgen05 flash attention for LightGlue (head dim 64, fp32-equivalent via 3xTF32 split operands).
//
//   ctx[z][row][h*64 + d] = softmax_j(scale * <q_row, k_j>) v_j      self: k,v of slot z; cross: of slot z^1
//
// Operands arrive pre-split from the projection epilogues as hi/lo planes (hi = 13 low mantissa bits cleared,
// lo = x - hi): q, k  [plane][slot][head][cap][64],  v transposed  [plane][slot][head][64][cap]  (kv contiguous,
// so that V^T is a plain K-major B operand).  Per CTA: one (slot, head, 128-row q tile).
//   warp 0     TMA producer: Q once, then K_j / V_j tiles of 64 keys (single-buffered; K_{j+1} is fetched while
//              softmax_j and P_j V_j run, V_{j+1} while S_{j+1} and softmax_{j+1} run)
//   warp 1     MMA issuer: S_j = Q K_j^T into one of two TMEM buffers (main hi*hi and cross-term accumulators),
//              O_j = P_j V_j into a fresh TMEM tile (never rescaled in place)
//   warps 2-17 softmax / correction, FOUR warps per TMEM sub-partition: a thread owns one q row and a QUARTER of the columns
//              (16 of the 64 keys of S_j, 16 of the 64 output dims of O): tcgen05.ld S_j, mask, row max (exchanged
//              with the three partner threads through shared memory), P_j = exp(S_j - m) written to shared memory as
//              K-major SWIZZLE_128B tiles (P itself = the hi operand, P - trunc_tf32(P) = the lo operand) for the second
//              MMA, and the running output acc = (acc + O_{j-1}) * exp(m_{j-1} - m_j) kept in registers.
//              (Round 1 ran 8 such warps with half a row each: the per-tile chain S -> softmax -> P -> PV was bound by
//              the LATENCY of those warps -- 2 per scheduler, ~700 dependent instructions per tile, tensor pipe 29 % --
//              not by any throughput; twice the warps with half the work each shortens exactly that chain.)
#pragma once
#include "common.cuh"
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

constexpr int TA_BQ = 128, TA_BKV = 64, TA_SM_THREADS = 512, TA_THREADS = 64 + TA_SM_THREADS;
constexpr int TA_Q_BYTES = 2 * 2 * TA_BQ * 128;     // hi/lo x 2 k-subtiles x [128 x 32 f32]  = 64 KB
constexpr int TA_K_BYTES = 2 * 2 * TA_BKV * 128;    // hi/lo x 2 subtiles x [64 x 32]        = 32 KB
constexpr int TA_V_BYTES = 2 * 2 * 64 * 128;        // hi/lo x 2 kv-subtiles x [64 d x 32 kv] = 32 KB
constexpr int TA_P_BYTES = 2 * 2 * TA_BQ * 128;     // hi/lo x 2 kv-subtiles x [128 x 32]     = 64 KB
constexpr size_t TA_SMEM = TA_Q_BYTES + TA_K_BYTES + TA_V_BYTES + TA_P_BYTES + 1024 + 256 + 2 * 4 * 128 * sizeof(float);

// TMEM columns: S buffers 2 x (main 64 + cross 64) = 256, O main 64 + cross 64 -> 384 (allocate 512)
constexpr int TA_TMEM_COLS = 512, TA_S_COL = 0, TA_O_COL = 256;

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
  uint8_t* sQ = smem;                 // [plane][sub][128 rows][128 B]
  uint8_t* sK = sQ + TA_Q_BYTES;      // [sub][plane][64 rows][128 B]   (hi | lo of one k-subtile adjacent: one N = 128 operand)
  uint8_t* sV = sK + TA_K_BYTES;      // [sub][plane][64 d rows][128 B]
  uint8_t* sP = sV + TA_V_BYTES;      // [plane][sub][128 rows][128 B]
  uint64_t* bars = (uint64_t*)(sP + TA_P_BYTES);
  uint64_t *q_full = bars, *k_full = bars + 1, *k_empty = bars + 2, *v_full = bars + 3, *v_empty = bars + 4,
           *s_full = bars + 5 /*[2]*/, *p_full = bars + 7, *o_full = bars + 8;
  uint32_t* tmem_slot = (uint32_t*)(bars + 9);
  float* xchg = (float*)(sP + TA_P_BYTES + 256);   // [tile parity][quarter][128 rows]: row maxima / final sums of the partner threads

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmQ); tc::tma_prefetch_desc(&tmK); tc::tma_prefetch_desc(&tmV);
    tc::mbar_init(q_full, 1); tc::mbar_init(k_full, 1); tc::mbar_init(k_empty, 1); tc::mbar_init(v_full, 1);
    tc::mbar_init(v_empty, 1); tc::mbar_init(s_full, 1); tc::mbar_init(s_full + 1, 1); tc::mbar_init(p_full, TA_SM_THREADS);
    tc::mbar_init(o_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, TA_TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int T = (nk + TA_BKV - 1) / TA_BKV;
  const int q_row = (z * 4 + head) * g.cap + q0;          // row in the [slots*4*cap][64] q/k tensors
  const int k_row0 = (zk * 4 + head) * g.cap;
  const int v_row = (zk * 4 + head) * 64;                 // row in the [slots*4*64][cap] v^T tensor

  if (warp == 0) {
    if (lane == 0) {
      tc::mbar_expect_tx(q_full, TA_Q_BYTES);
      for (int p = 0; p < 2; p++)
        for (int sub = 0; sub < 2; sub++)
          tc::tma_load_2d(sQ + (p * 2 + sub) * TA_BQ * 128, &tmQ, q_full, sub * 32, (int)(p * g.plane_rows_qk) + q_row);
      // K and V tiles are refilled INDEPENDENTLY, each as soon as its own consumer MMAs have completed (round 1 walked
      // K_j, V_j, K_{j+1} ... in one sequence, so the K_{j+1} load -- and with it S_{j+1} -- waited for P_{j-1} V_{j-1}).
      // A barrier cannot run ahead of the phase polled here: S_{j+1} needs K_{j+1}, which is only requested after
      // k_empty(j) was observed (same for V), so a non-blocking parity test never misses a phase.
      int jk = 0, jv = 0;
      while (jk < T || jv < T) {
        if (jk < T && (jk == 0 || tc::mbar_try_wait(k_empty, (jk - 1) & 1))) {
          tc::mbar_expect_tx(k_full, TA_K_BYTES);
          for (int p = 0; p < 2; p++)
            for (int sub = 0; sub < 2; sub++)
              tc::tma_load_2d(sK + (sub * 2 + p) * TA_BKV * 128, &tmK, k_full, sub * 32, (int)(p * g.plane_rows_qk) + k_row0 + jk * TA_BKV);
          jk++;
        }
        if (jv < T && (jv == 0 || tc::mbar_try_wait(v_empty, (jv - 1) & 1))) {
          tc::mbar_expect_tx(v_full, TA_V_BYTES);
          for (int p = 0; p < 2; p++)
            for (int sub = 0; sub < 2; sub++)
              tc::tma_load_2d(sV + (sub * 2 + p) * 64 * 128, &tmV, v_full, jv * TA_BKV + sub * 32, (int)(p * g.plane_rows_vt) + v_row);
          jv++;
        }
      }
    }
  } else if (warp == 1) {
    {
      const bool leader = tc::elect_one();
      constexpr uint32_t idesc = tc::make_idesc(tc::FMT_TF32, 128, 64), idesc2 = tc::make_idesc(tc::FMT_TF32, 128, 128);
      const uint32_t aQ = tc::smem_u32(sQ), aK = tc::smem_u32(sK), aV = tc::smem_u32(sV), aP = tc::smem_u32(sP);
      auto issue_S = [&](int j) {  // S_j = Q K_j^T : main = hi*hi, cross = hi*lo + lo*hi   (k_full(j) already observed)
        tc::fence_after_sync();
        const uint32_t d_main = tmem_base + TA_S_COL + (j & 1) * 128, d_cross = d_main + 64;
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
          const int sub = ks / 4, ko = (ks % 4) * 32;
          uint64_t qh = tc::make_smem_desc_sw128(aQ + (0 * 2 + sub) * TA_BQ * 128 + ko), ql = tc::make_smem_desc_sw128(aQ + (1 * 2 + sub) * TA_BQ * 128 + ko);
          uint64_t kh = tc::make_smem_desc_sw128(aK + (sub * 2 + 0) * TA_BKV * 128 + ko);   // [K_hi | K_lo]: 128 adjacent rows
          if (leader) {
            tc::mma_tf32(d_main, qh, kh, idesc2, ks ? 1u : 0u);   // Q_hi x [K_hi | K_lo] -> [main | cross]
            tc::mma_tf32(d_cross, ql, kh, idesc, 1u);             // Q_lo x K_hi -> cross
          }
        }
        if (leader) {
          tc::mma_commit(k_empty);
          tc::mma_commit(s_full + (j & 1));
        }
        __syncwarp();
      };
      auto issue_PV = [&](int j) {  // O_j = P_j V_j, K = 64 keys   (p_full(j), v_full(j) already observed)
        tc::fence_after_sync();
        const uint32_t d_main = tmem_base + TA_O_COL, d_cross = d_main + 64;
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
          const int sub = ks / 4, ko = (ks % 4) * 32;
          uint64_t ph = tc::make_smem_desc_sw128(aP + (0 * 2 + sub) * TA_BQ * 128 + ko), pl = tc::make_smem_desc_sw128(aP + (1 * 2 + sub) * TA_BQ * 128 + ko);
          uint64_t vh = tc::make_smem_desc_sw128(aV + (sub * 2 + 0) * 64 * 128 + ko);       // [V_hi | V_lo]
          if (leader) {
            tc::mma_tf32(d_main, ph, vh, idesc2, ks ? 1u : 0u);
            tc::mma_tf32(d_cross, pl, vh, idesc, 1u);
          }
        }
        if (leader) {
          tc::mma_commit(v_empty);
          tc::mma_commit(o_full);
        }
        __syncwarp();
      };
      // Whichever of S_{js} / P_{jp} V_{jp} has its operands ready is issued next (round 1 issued S_{j+1} strictly before
      // P_j V_j and blocked on K_{j+1} while P_j was already waiting).  S_{js} writes TMEM buffer js & 1, free once softmax(js-2)
      // has read it = p_full(js-2) observed = jp >= js - 1.  The polls are warp-uniform (same barrier, same parity in every lane).
      tc::mbar_wait(q_full, 0);
      int js = 0, jp = 0;
      uint32_t spins = 0;
      const int lead_lane = __ffs(__ballot_sync(0xffffffffu, leader)) - 1;
      // the issuing lane polls, the warp follows its verdict (a per-lane poll could split the warp around the __syncwarp()s)
      auto ready = [&](bool cond_s) -> bool {
        unsigned r = 0;
        if (leader) r = cond_s ? (tc::mbar_try_wait(k_full, js & 1) ? 1u : 0u)
                               : ((tc::mbar_try_wait(p_full, jp & 1) && tc::mbar_try_wait(v_full, jp & 1)) ? 1u : 0u);
        return __shfl_sync(0xffffffffu, r, lead_lane) != 0;
      };
      while (jp < T) {
        bool did = false;
        if (js < T && js <= jp + 1 && ready(true)) { issue_S(js); js++; did = true; }
        if (jp < js && ready(false)) { issue_PV(jp); jp++; did = true; }
        if (did) spins = 0;
        else if (++spins > (1u << 24)) { printf("tc_attn MMA scheduler timeout block (%d,%d,%d) js %d jp %d\n", blockIdx.x, blockIdx.y, blockIdx.z, js, jp); __trap(); }
      }
    }
  } else {
    const int q = warp % 4, r = q * 32 + lane;           // TMEM lane = q row inside the tile
    const int qt = (warp - 2) / 4;                       // keys [16 qt, +16) of S_j and dims [16 qt, +16) of O
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    float m = -INFINITY, l = 0.f, acc[16];
#pragma unroll
    for (int c = 0; c < 16; c++) acc[c] = 0.f;
    for (int j = 0; j < T; j++) {
      tc::mbar_wait(s_full + (j & 1), (j >> 1) & 1);
      tc::fence_after_sync();
      float s[16];
      {
        float t[16];
        const uint32_t a = lane_addr + TA_S_COL + (j & 1) * 128 + qt * 16;
        tc::tmem_ld16(a, s);
        tc::tmem_ld16(a + 64, t);
#pragma unroll
        for (int c = 0; c < 16; c++) s[c] += t[c];
      }
      const int kv0 = j * TA_BKV + qt * 16;
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 16; c++) {
        s[c] = (kv0 + c < nk) ? s[c] * g.scale : -INFINITY;
        mx = fmaxf(mx, s[c]);
      }
      // row maximum over the four column quarters (partner threads = same row, other quarters)
      float* xb = xchg + (j & 1) * 512;
      xb[qt * 128 + r] = mx;
      asm volatile("bar.sync 1, %0;" ::"n"(TA_SM_THREADS) : "memory");
      const float m_new = fmaxf(m, fmaxf(fmaxf(xb[r], xb[128 + r]), fmaxf(xb[256 + r], xb[384 + r])));   // finite: the tile holds a valid key
      const float alpha = __expf(m - m_new);     // 0 on the first tile (m = -inf)
      float ps = 0.f;
#pragma unroll
      for (int c = 0; c < 16; c++) { s[c] = __expf(s[c] - m_new); ps += s[c]; }
      l = l * alpha + ps;                        // partial sum over this thread's columns
      m = m_new;
      if (j > 0) {  // fold in O_{j-1} (computed relative to m_{j-1}), then move the reference to m_j
        tc::mbar_wait(o_full, (j - 1) & 1);
        tc::fence_after_sync();
        float t[16];
        tc::tmem_ld16(lane_addr + TA_O_COL + qt * 16, t);        // main
#pragma unroll
        for (int c = 0; c < 16; c++) acc[c] += t[c];
        tc::tmem_ld16(lane_addr + TA_O_COL + 64 + qt * 16, t);   // cross terms
#pragma unroll
        for (int c = 0; c < 16; c++) acc[c] += t[c];
      }
#pragma unroll
      for (int c = 0; c < 16; c++) acc[c] *= alpha;
      // P_j -> shared memory: P (= hi operand: kind::tf32 ignores the 13 low mantissa bits) and P - trunc_tf32(P) (lo), K-major
      // rows of 128 B with the 128B swizzle (16-byte chunk c of row r lives at chunk c ^ (r & 7)); keys [32 sub, +32) form
      // k-subtile `sub`; the previous P V MMA has completed (o_full above)
      tc::fence_before_sync();
      {
        const int sub = qt >> 1, ch0 = (qt & 1) * 4;
        uint8_t* ph = sP + (0 * 2 + sub) * TA_BQ * 128 + r * 128;
        uint8_t* pl = sP + (1 * 2 + sub) * TA_BQ * 128 + r * 128;
#pragma unroll
        for (int ch = 0; ch < 4; ch++) {
          uint4 h4, l4;
          const float* v = &s[ch * 4];
          h4.x = __float_as_uint(v[0]); h4.y = __float_as_uint(v[1]); h4.z = __float_as_uint(v[2]); h4.w = __float_as_uint(v[3]);
          l4.x = __float_as_uint(v[0] - __uint_as_float(h4.x & 0xFFFFE000u)); l4.y = __float_as_uint(v[1] - __uint_as_float(h4.y & 0xFFFFE000u));
          l4.z = __float_as_uint(v[2] - __uint_as_float(h4.z & 0xFFFFE000u)); l4.w = __float_as_uint(v[3] - __uint_as_float(h4.w & 0xFFFFE000u));
          const int pos = ((ch0 + ch) ^ (r & 7)) * 16;
          *reinterpret_cast<uint4*>(ph + pos) = h4;
          *reinterpret_cast<uint4*>(pl + pos) = l4;
        }
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(p_full);
    }
    // last tile's O, normalise, store
    tc::mbar_wait(o_full, (T - 1) & 1);
    tc::fence_after_sync();
    {
      float t[16];
      tc::tmem_ld16(lane_addr + TA_O_COL + qt * 16, t);
#pragma unroll
      for (int c = 0; c < 16; c++) acc[c] += t[c];
      tc::tmem_ld16(lane_addr + TA_O_COL + 64 + qt * 16, t);
#pragma unroll
      for (int c = 0; c < 16; c++) acc[c] += t[c];
    }
    float* xb = xchg + (T & 1) * 512;   // the buffer the last tile did not use
    xb[qt * 128 + r] = l;
    asm volatile("bar.sync 1, %0;" ::"n"(TA_SM_THREADS) : "memory");
    const int row = q0 + r;
    if (row < nq) {
      const float inv = 1.f / ((xb[r] + xb[128 + r]) + (xb[256 + r] + xb[384 + r]));
      float4* o = reinterpret_cast<float4*>(g.ctx + ((long long)z * g.cap + row) * 256 + head * 64 + qt * 16);
#pragma unroll
      for (int c = 0; c < 4; c++) o[c] = make_float4(acc[4 * c] * inv, acc[4 * c + 1] * inv, acc[4 * c + 2] * inv, acc[4 * c + 3] * inv);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, TA_TMEM_COLS);
}

// q_planes / k_planes: [2][slots*4*cap][64] fp32, vt_planes: [2][slots*4*64][cap] fp32
static inline int launch_tc_attn(const float* q_planes, const float* k_planes, const float* vt_planes, TcAttnArgs g, cudaStream_t st) {
  CUtensorMap tmQ, tmK, tmV;
  const long long rows_qk = 2 * g.plane_rows_qk, rows_vt = 2 * g.plane_rows_vt;
  if (int e = tc_make_map_2d_f32(&tmQ, q_planes, (uint64_t)rows_qk, 64, 64, 32, TA_BQ)) return e;
  if (int e = tc_make_map_2d_f32(&tmK, k_planes, (uint64_t)rows_qk, 64, 64, 32, TA_BKV)) return e;
  if (int e = tc_make_map_2d_f32(&tmV, vt_planes, (uint64_t)rows_vt, (uint64_t)g.cap, (uint64_t)g.cap, 32, 64)) return e;
  IMW_SMEM_ATTR_ONCE(tc_attn_kernel, TA_SMEM);
  dim3 grid(g.cap / TA_BQ, 4, g.slots);
  tc_attn_kernel<<<grid, TA_THREADS, TA_SMEM, st>>>(tmQ, tmK, tmV, g);
  IMW_CHECK_LAUNCH();
  return IMW_OK;
}
