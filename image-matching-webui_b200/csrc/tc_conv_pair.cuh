// Cin = Cout = 64 3x3 conv on a CTA PAIR (tcgen05 cta_group::2, M = 256): included by tc_conv.cu inside its anonymous namespace.
//
// What bounded the single-CTA c64 kernel (ncu, profiles/r2c_*): every 16 x 8 pixel tile re-streamed the nine weight taps
// (144 KB) and three dx-shifted halo copies (108 KB) from L2 -- 252 KB per 3456 clk of MMA, and the kernel sustained ~28 B/clk
// per SM -- and, in the fused first layer, the producer warps wrote every conv1a value four times
// (staging + three copies) through the same shared-memory data pipe the tensor core reads its operands from (LSU 40 % +
// tensor 39 % of the wavefront peak).  Here:
//   * ONE dense halo patch per plane (18 rows x 10 px x 128 B, SWIZZLE_128B as TMA writes it) serves all nine taps: the A
//     descriptor of tap (dy, dx) starts (dy * 10 + dx) pixels into the patch with a stride of 1280 B between 8-pixel row groups.
//     The tensor core applies the 128-byte swizzle to absolute shared-memory address bits, so a start inside a 1024-byte atom
//     and a stride that is not a multiple of 1024 address exactly the bytes TMA (or the conv1a producers) put there
//     (tools/probe/swz_probe.cu, run on a B200: 0 bad rows for every (dy, dx, k)).  46 KB per tile instead of 108 KB, written once.
//   * the two CTAs of a cluster work on two pixel tiles as ONE M = 256 MMA; each CTA supplies half of the B operand's N, so the
//     weights of all nine taps stay RESIDENT in shared memory (108 KB per CTA) -- zero weight traffic after the prologue -- and the
//     per-SM operand reads of B halve (N = 64 tiles were bound by them).
//     Per tap and CTA: [0, 8 KB) = this CTA's half of [b_hi | b_lo] (rank 0: b_hi, rank 1: b_lo), [8 KB, 12 KB) = its half of
//     b_hi for the a_lo x b_hi product (rank 0: rows 0..31, rank 1: rows 32..63).
// Roles per CTA: warp 0 TMA (weights once; halo patches unless FUSE), warp 1 TMEM alloc (+ all MMAs, leader CTA only),
// warps 2..9 epilogue (2..17 in the plain conv), warps 10..17 conv1a producers (FUSE).  Cross-CTA signalling: halo patches complete on the LEADER's
// a_full barrier (TMA .cta_group::2 / remote mbarrier.arrive), tcgen05.commit multicasts a_empty and tmem_full to both CTAs, the
// epilogue warps of both CTAs arrive on the leader's tmem_empty.
constexpr int P2_HALO_W = 10, P2_HALO_H = 18, P2_HALO_PX = P2_HALO_W * P2_HALO_H;
constexpr int P2_PLANE = 23552;                 // 180 px x 128 B = 23040, padded to a multiple of 1024
constexpr int P2_A_BUF = NP * P2_PLANE;         // one tile's operand planes
constexpr int P2_A_BYTES = 2 * P2_A_BUF;        // double-buffered
constexpr int P2_W_TAP = 12288;
constexpr int P2_W_BYTES = 9 * P2_W_TAP;
constexpr int P2_BAR_OFF = P2_A_BYTES + P2_W_BYTES;
constexpr size_t P2_SMEM = P2_BAR_OFF + 128 + 2 * 240 * sizeof(float) + 64 * sizeof(float) /*bias*/ + 1024;
constexpr int P2_EPI_WARPS = 8, P2_PROD_WARPS = 8;   // fused first layer: 8 epilogue + 8 conv1a producer warps
constexpr int P2_EPI_WARPS_PLAIN = 16;              // plain conv: the epilogue is the only CUDA-core work, 4 warps per TMEM sub-partition
constexpr int P2_THREADS = 64 + 32 * P2_EPI_WARPS, P2_FUSE_THREADS = P2_THREADS + 32 * P2_PROD_WARPS;
constexpr int P2_PLAIN_THREADS = 64 + 32 * P2_EPI_WARPS_PLAIN;

// (short names for the cta_group::2 primitives of tc_common.cuh)
__device__ __forceinline__ uint64_t p2_desc(uint32_t smem_addr, uint32_t sbo) { return tc::make_smem_desc_sw128_sbo(smem_addr, sbo); }
__device__ __forceinline__ uint32_t p2_mapa(const void* p, uint32_t rank) { return tc::mapa(p, rank); }
__device__ __forceinline__ void p2_arrive_remote(uint32_t cluster_addr) { tc::mbar_arrive_cluster(cluster_addr); }
__device__ __forceinline__ void p2_wait_cluster(uint64_t* bar, uint32_t parity) { tc::mbar_wait_cluster(bar, parity); }
__device__ __forceinline__ void p2_tma_4d_pair(void* smem_dst, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          tc::smem_u32(smem_dst)),
      "l"(m), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void p2_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) { tc::mma_f16_pair(d, a, b, idesc, acc); }
__device__ __forceinline__ void p2_commit(uint64_t* bar) { tc::mma_commit_pair(bar); }
__device__ __forceinline__ void p2_tmem_alloc(uint32_t* smem_result, uint32_t ncols) { tc::tmem_alloc_pair(smem_result, ncols); }
__device__ __forceinline__ void p2_tmem_dealloc(uint32_t taddr, uint32_t ncols) { tc::tmem_dealloc_pair(taddr, ncols); }

// two values -> fp16 plane pairs (split_planes.cuh: packed saturating converts, 4 instructions per value instead of 7)
__device__ __forceinline__ void split2_pos(float x0, float x1, __half2& hi, __half2& lo) { split2x2(x0, x1, hi, lo); }
__device__ __forceinline__ void split2_small(float x0, float x1, __half2& hi, __half2& lo) { split2x2(x0, x1, hi, lo); }

// N (16 / 32) columns of a split-precision accumulator: (set0.main + set1.main) + (set0.cross + set1.cross) 2^-11 (tc::tmem_ld_acc32)
template <int N>
__device__ __forceinline__ void p2_ld_acc(uint32_t lane_base, int cross_off, int set_stride, float (&v)[N]) {
#pragma unroll
  for (int h = 0; h < N / 16; h++) {
    uint32_t m0[16], m1[16], c0[16], c1[16];
    tc::tmem_ld16_async(lane_base + 16 * h, m0);
    tc::tmem_ld16_async(lane_base + set_stride + 16 * h, m1);
    tc::tmem_ld16_async(lane_base + cross_off + 16 * h, c0);
    tc::tmem_ld16_async(lane_base + set_stride + cross_off + 16 * h, c1);
    tc::tmem_wait_ld();
    tc::tmem_ld_fence(m0); tc::tmem_ld_fence(m1); tc::tmem_ld_fence(c0); tc::tmem_ld_fence(c1);
#pragma unroll
    for (int j = 0; j < 16; j++)
      v[16 * h + j] = fmaf(__uint_as_float(c0[j]) + __uint_as_float(c1[j]), PLANE_LO_INV, __uint_as_float(m0[j]) + __uint_as_float(m1[j]));
  }
}

// swap_halves: debugging aid (which CTA's rows are the first half of B's N)
template <bool FUSE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FUSE ? P2_FUSE_THREADS : P2_PLAIN_THREADS, 1)
tc_conv3x3_c64_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, ConvArgs g, int total_tiles,
                           int swap_halves) {
  constexpr int BN = 64;
  constexpr int EPW = FUSE ? P2_EPI_WARPS : P2_EPI_WARPS_PLAIN, CPW = 256 / EPW;   // epilogue warps; output channels per warp (32 / 16)
  extern __shared__ uint8_t cv_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)cv_smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                         // [buf][plane][180 px][128 B]
  uint8_t* sW = smem + P2_A_BYTES;            // [tap][12 KB]
  uint64_t* a_full = (uint64_t*)(smem + P2_BAR_OFF);   // [2]  (used in the leader CTA)
  uint64_t* a_empty = a_full + 2;             // [2]
  uint64_t* tmem_full = a_empty + 2;          // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]  (used in the leader CTA)
  uint64_t* w_full = tmem_empty + 2;          // [1]
  uint32_t* tmem_slot = (uint32_t*)(w_full + 1);
  float* s_img = (float*)(smem + P2_BAR_OFF + 128);    // FUSE: [2][20 rows][12 cols]
  float* s_bias = s_img + 2 * 240;                     // conv bias [64]

  const uint32_t rank = tc::cluster_ctarank();
  const int tiles_x = g.W / C64_TW, tiles_y = (g.H + C64_TH - 1) / C64_TH;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int n_pairs = (total_tiles + 1) / 2, n_clusters = gridDim.x / 2, cluster = blockIdx.x / 2;
  if (warp == 0 && lane == 0) {
    if (!FUSE) tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmW);
    for (int i = 0; i < 2; i++) {
      tc::mbar_init(a_full + i, FUSE ? 2 * P2_PROD_WARPS : 1);
      tc::mbar_init(a_empty + i, 1);
      tc::mbar_init(tmem_full + i, 1);
      tc::mbar_init(tmem_empty + i, 2 * EPW);
    }
    tc::mbar_init(w_full, 1);
    tc::fence_barrier_init();
    // resident weights: this CTA's halves of the nine taps (box = 32 rows x 64 channels = 4 KB)
    tc::mbar_expect_tx(w_full, P2_W_BYTES);
    const int r1 = (int)(rank ^ (uint32_t)(swap_halves & 1));
    for (int tap = 0; tap < 9; tap++) {
      uint8_t* slot = sW + tap * P2_W_TAP;
      const int row_main = (r1 * 9 + tap) * BN;         // rank 0: b_hi, rank 1: b_lo
      tc::tma_load_2d(slot, &tmW, w_full, 0, row_main);
      tc::tma_load_2d(slot + 4096, &tmW, w_full, 0, row_main + 32);
      tc::tma_load_2d(slot + 8192, &tmW, w_full, 0, tap * BN + 32 * r1);   // half of b_hi for a_lo x b_hi
    }
    tc::mbar_wait(w_full, 0);   // landed before the cluster barrier below: the leader's MMAs read BOTH CTAs' weights
  }
  if (threadIdx.x < 64) s_bias[threadIdx.x] = g.bias[threadIdx.x];
  if (warp == 1) p2_tmem_alloc(tmem_slot, 512);
  tc::fence_before_sync();
  tc::cluster_sync();       // barriers of both CTAs initialised before any remote arrive / multicast commit
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_of = [&](int it, int& tile, bool& valid) {
    const int pair = cluster + it * n_clusters;
    tile = 2 * pair + (int)rank;
    valid = tile < total_tiles;
    if (!valid) tile = total_tiles - 1;     // odd tile count: the partner recomputes the last tile and stores nothing
  };
  const int my_iters = (n_pairs - cluster + n_clusters - 1) / n_clusters;

  if (warp == 0) {
    if (!FUSE && lane == 0) {
      for (int i = 0; i < my_iters; i++) {
        int tile; bool valid;
        tile_of(i, tile, valid);
        const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
        const int x0 = tx * C64_TW, y0 = ty * C64_TH, buf = i & 1;
        tc::mbar_wait(a_empty + buf, ((i >> 1) & 1) ^ 1);
        if (rank == 0) tc::mbar_expect_tx(a_full + buf, 2 * NP * P2_HALO_PX * 128);   // both CTAs' patches
        const uint32_t lbar = p2_mapa(a_full + buf, 0);
#pragma unroll
        for (int p = 0; p < NP; p++)
          p2_tma_4d_pair(sA + buf * P2_A_BUF + p * P2_PLANE, &tmA, lbar, 0, x0 - 1, y0 - 1, p * g.B + b);
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      const bool leader = tc::elect_one();
      constexpr uint32_t idesc = tc::make_idesc(tc::FMT_F16, 256, BN), idesc2 = tc::make_idesc(tc::FMT_F16, 256, 2 * BN);
      // Descriptors: the address field is bits [0, 14) in units of 16 B, so every operand of the tile is a compile-time offset added
      // to one base descriptor (the single issuing thread was the bottleneck when it rebuilt four descriptors per k-step:
      // 2.5 us per tile against 1.8 us of MMA).
      const uint64_t a_desc0 = p2_desc(tc::smem_u32(sA), P2_HALO_W * 128), w_desc0 = p2_desc(tc::smem_u32(sW), 1024);
      for (int i = 0; i < my_iters; i++) {
        const int acc = i & 1, buf = i & 1;
        p2_wait_cluster(tmem_empty + acc, ((i >> 1) & 1) ^ 1);   // arrivals come from both CTAs
        p2_wait_cluster(a_full + buf, (i >> 1) & 1);
        tc::fence_after_sync();
        const uint32_t d_base = tmem_base + acc * 256;
        const uint64_t a_d = a_desc0 + (uint64_t)((buf * P2_A_BUF) >> 4);
        if (leader && !(swap_halves & 1024)) {
#pragma unroll
          for (int tap = 0; tap < 9; tap++) {
            const uint32_t d_set = d_base + (tap & 1) * 2 * BN;     // [main | cross] sets alternate tap by tap
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const uint64_t a_hi = a_d + (uint64_t)((((tap / 3) * P2_HALO_W + tap % 3) * 128 + k * 32) >> 4);
              const uint64_t a_lo = a_hi + (uint64_t)(P2_PLANE >> 4);
              const uint64_t b_main = w_desc0 + (uint64_t)((tap * P2_W_TAP + k * 32) >> 4), b_x = b_main + (uint64_t)(8192 >> 4);
              p2_mma(d_set, a_hi, b_main, idesc2, (tap >= 2 || k) ? 1u : 0u);   // a_hi x [b_hi | b_lo] -> [main | cross]
              p2_mma(d_set + BN, a_lo, b_x, idesc, 1u);                          // a_lo x b_hi -> cross
            }
          }
        }
        if (leader) {
          p2_commit(a_empty + buf);
          p2_commit(tmem_full + acc);
        }
        __syncwarp();
      }
    }
  } else if (FUSE && warp >= 2 + EPW) {
    // conv1a producers.  Thread = (segment of 6 halo rows, pair of adjacent halo columns, quad of 4 output channels): its 36 weights
    // live in registers and it walks down its two columns with the 3 x 4 image window in registers (four new values per row):
    // 72 FMAs on 8 independent accumulators per row step, no weight loads, no index arithmetic in the loop.  (8 channels per thread
    // need 72 weight registers, which caps the CTA at 16 warps; weights read from shared memory per tap left the warps waiting on
    // load latency: 4300 clk per tile against 1350 of issue.)  240 of the 256 threads are active; a warp stores 2 pixels x 128 B per
    // instruction (conflict-free through the swizzle).  The next tile's image patch is fetched into a register before the walk and
    // parked in shared memory after it: its global-memory latency hides behind the FMAs.
    constexpr int NT = 32 * P2_PROD_WARPS;
    const int t = threadIdx.x - P2_THREADS;   // 0..NT-1
    const int quad = t % 16;                  // output channels [4 quad, +4) = half of a 16-byte unit of a pixel's 128-byte row
    const int hx = 2 * ((t / 16) % 5), seg = t / 80;   // halo columns hx, hx + 1; halo rows [6 seg, 6 seg + 6)
    const bool active = seg < 3;
    float w[9][4], bv[4];
#pragma unroll
    for (int tp = 0; tp < 9; tp++)
#pragma unroll
      for (int k = 0; k < 4; k++) w[tp][k] = g.w1a[tp * 64 + quad * 4 + k];
#pragma unroll
    for (int k = 0; k < 4; k++) bv[k] = g.b1a[quad * 4 + k];
    auto patch_load = [&](int it) -> float {   // element t of the 20 x 12 image patch of iteration `it` (zeros outside = conv1a padding)
      if (t >= 240 || it >= my_iters) return 0.f;
      int tile; bool valid;
      tile_of(it, tile, valid);
      const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
      const int gy = ty * C64_TH - 2 + t / 12, gx = tx * C64_TW - 2 + t % 12;
      return (gy >= 0 && gy < g.H && gx >= 0 && gx < g.W) ? g.img[((size_t)b * g.H + gy) * g.W + gx] : 0.f;
    };
    if (t < 240) s_img[t] = patch_load(0);
    asm volatile("bar.sync 2, %0;" ::"n"(NT) : "memory");
    for (int i = 0; i < my_iters; i++) {
      int tile; bool valid;
      tile_of(i, tile, valid);
      const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y;
      const int x0 = tx * C64_TW, y0 = ty * C64_TH, buf = i & 1;
      const float* im = s_img + (i & 1) * 240;
      const float nxt = patch_load(i + 1);
      tc::mbar_wait(a_empty + buf, ((i >> 1) & 1) ^ 1);   // the MMAs of tile i-2 have read this buffer
      uint8_t* dst = sA + buf * P2_A_BUF + (quad & 1) * 8;
      if (active && !(swap_halves & 256)) {
        const int gx = x0 - 1 + hx;
        const bool in0 = gx >= 0 && gx < g.W, in1 = gx + 1 >= 0 && gx + 1 < g.W;
        float win[3][4];
#pragma unroll
        for (int r = 0; r < 2; r++) {
          const float2 lo2 = *reinterpret_cast<const float2*>(im + (6 * seg + r) * 12 + hx), hi2 = *reinterpret_cast<const float2*>(im + (6 * seg + r) * 12 + hx + 2);
          win[r + 1][0] = lo2.x; win[r + 1][1] = lo2.y; win[r + 1][2] = hi2.x; win[r + 1][3] = hi2.y;
        }
#pragma unroll
        for (int r = 0; r < 6; r++) {
          const int hy = 6 * seg + r, pidx = hy * P2_HALO_W + hx;
          {
            const float2 lo2 = *reinterpret_cast<const float2*>(im + (hy + 2) * 12 + hx), hi2 = *reinterpret_cast<const float2*>(im + (hy + 2) * 12 + hx + 2);
#pragma unroll
            for (int c = 0; c < 4; c++) { win[0][c] = win[1][c]; win[1][c] = win[2][c]; }
            win[2][0] = lo2.x; win[2][1] = lo2.y; win[2][2] = hi2.x; win[2][3] = hi2.y;
          }
          float a0[4], a1[4];
#pragma unroll
          for (int c = 0; c < 4; c++) a0[c] = a1[c] = 0.f;
#pragma unroll
          for (int dy = 0; dy < 3; dy++)
#pragma unroll
            for (int dx = 0; dx < 3; dx++)
#pragma unroll
              for (int c = 0; c < 4; c++) {
                a0[c] = fmaf(win[dy][dx], w[dy * 3 + dx][c], a0[c]);
                a1[c] = fmaf(win[dy][dx + 1], w[dy * 3 + dx][c], a1[c]);
              }
          const int gy = y0 - 1 + hy;
          const bool row_in = gy >= 0 && gy < g.H;
          const bool i0 = row_in && in0, i1 = row_in && in1;   // conv1b's zero padding: outside the image -> 0
#pragma unroll
          for (int c = 0; c < 4; c++) {   // (same operation order as the stand-alone conv1a kernel: taps, then bias, then ReLU)
            a0[c] = i0 ? fmaxf(a0[c] + bv[c], 0.f) : 0.f;
            a1[c] = i1 ? fmaxf(a1[c] + bv[c], 0.f) : 0.f;
          }
          uint2 h2, l2;
          __half2* hp = reinterpret_cast<__half2*>(&h2);
          __half2* lp = reinterpret_cast<__half2*>(&l2);
          split2_small(a0[0], a0[1], hp[0], lp[0]); split2_small(a0[2], a0[3], hp[1], lp[1]);
          int off = pidx * 128 + (((quad >> 1) ^ (pidx & 7)) * 16);   // SWIZZLE_128B: 16-byte unit c of row r sits at c ^ (r & 7)
          *reinterpret_cast<uint2*>(dst + off) = h2;
          *reinterpret_cast<uint2*>(dst + P2_PLANE + off) = l2;
          split2_small(a1[0], a1[1], hp[0], lp[0]); split2_small(a1[2], a1[3], hp[1], lp[1]);
          off = (pidx + 1) * 128 + (((quad >> 1) ^ ((pidx + 1) & 7)) * 16);
          *reinterpret_cast<uint2*>(dst + off) = h2;
          *reinterpret_cast<uint2*>(dst + P2_PLANE + off) = l2;
        }
      }
      tc::fence_proxy_async();   // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) p2_arrive_remote(p2_mapa(a_full + buf, 0));
      if (t < 240) s_img[((i + 1) & 1) * 240 + t] = nxt;
      asm volatile("bar.sync 2, %0;" ::"n"(NT) : "memory");   // next patch complete; every thread is done reading this one
    }
  } else if (warp >= 2 && warp < 2 + EPW) {
    // epilogue: EPW / 4 warps per TMEM sub-partition, CPW output channels each
    const int q = warp % 4, c0 = ((warp - 2) / 4) * CPW;
    const int m = q * 32 + lane;              // pixel index in the tile: row m/8, col m%8
    const int Ho = g.pool ? g.H / 2 : g.H, Wo = g.pool ? g.W / 2 : g.W;
    const size_t plane_stride = (size_t)g.B * Ho * Wo * g.Cout;
    // pooled outputs: after the 2x2 exchange lane (b0 = lane & 1, b3 = lane & 8) owns channels [cq, cq + CPW / 4) of its window
    const bool b0 = (lane & 1) != 0, b3 = (lane & 8) != 0;
    const int cq = c0 + (b0 ? CPW / 2 : 0) + (b3 ? CPW / 4 : 0);
    float bq[CPW / 4];
#pragma unroll
    for (int e = 0; e < CPW / 4; e++) bq[e] = g.bias[cq + e];
    for (int i = 0; i < my_iters; i++) {
      int tile; bool valid;
      tile_of(i, tile, valid);
      const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
      const int x0 = tx * C64_TW, y0 = ty * C64_TH;
      const int acc = i & 1;
      tc::mbar_wait(tmem_full + acc, (i >> 1) & 1);
      tc::fence_after_sync();
      const int py = y0 + m / C64_TW, px = x0 + m % C64_TW;
      const int oy = g.pool ? py / 2 : py, ox = g.pool ? px / 2 : px;
      const bool in_img = valid && (py < g.H) && (px < g.W);
      const size_t opix = (((size_t)b * Ho + oy) * Wo + ox) * g.Cout;
      float v[CPW];
      const uint32_t lane_base = tmem_base + acc * 256 + ((uint32_t)(q * 32) << 16) + c0;
      p2_ld_acc<CPW>(lane_base, BN, 2 * BN, v);   // (main0 + main1) + (cross0 + cross1) 2^-11
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) p2_arrive_remote(p2_mapa(tmem_empty + acc, 0));   // this warp's only TMEM read of the set: back to the MMA warp
      if (swap_halves & 512) continue;
      if (g.pool) {
        // 2x2 max-pool = lanes {l, l^1, l^8} (4 image rows x 8 cols per warp) as an exchange-and-halve butterfly: 3/4 CPW shuffles per
        // lane instead of 2 CPW, and every lane ends up with CPW / 4 pooled channels to bias / ReLU / split / store (bias and ReLU
        // commute with the max: same values bit for bit)
        float u[CPW / 2], r[CPW / 4];
#pragma unroll
        for (int j = 0; j < CPW / 2; j++) {
          const float keep = b0 ? v[CPW / 2 + j] : v[j], send = b0 ? v[j] : v[CPW / 2 + j];
          u[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
        }
#pragma unroll
        for (int e = 0; e < CPW / 4; e++) {
          const float keep = b3 ? u[CPW / 4 + e] : u[e], send = b3 ? u[e] : u[CPW / 4 + e];
          r[e] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 8));
        }
#pragma unroll
        for (int e = 0; e < CPW / 4; e++) {
          r[e] += bq[e];
          if (g.relu) r[e] = fmaxf(r[e], 0.f);
        }
        if (in_img) {
          if (g.out_fp32) {
#pragma unroll
            for (int e = 0; e < CPW / 4; e += 4) *reinterpret_cast<float4*>(g.out_f32 + opix + cq + e) = make_float4(r[e], r[e + 1], r[e + 2], r[e + 3]);
          } else {
            __align__(16) __half2 hp[CPW / 8], lp[CPW / 8];
            if (g.relu) {
#pragma unroll
              for (int e = 0; e < CPW / 8; e++) split2_pos(r[2 * e], r[2 * e + 1], hp[e], lp[e]);
            } else {
              plane_t* hh = reinterpret_cast<plane_t*>(hp);
              plane_t* ll = reinterpret_cast<plane_t*>(lp);
#pragma unroll
              for (int e = 0; e < CPW / 4; e++) split2(r[e], hh[e], ll[e]);
            }
            if (CPW == 32) {
              *reinterpret_cast<uint4*>(g.out_planes + opix + cq) = *reinterpret_cast<const uint4*>(hp);
              *reinterpret_cast<uint4*>(g.out_planes + plane_stride + opix + cq) = *reinterpret_cast<const uint4*>(lp);
            } else {
              *reinterpret_cast<uint2*>(g.out_planes + opix + cq) = *reinterpret_cast<const uint2*>(hp);
              *reinterpret_cast<uint2*>(g.out_planes + plane_stride + opix + cq) = *reinterpret_cast<const uint2*>(lp);
            }
          }
        }
        continue;
      }
#pragma unroll
      for (int j = 0; j < CPW; j++) {
        float x = v[j] + s_bias[c0 + j];
        if (g.relu) x = fmaxf(x, 0.f);
        v[j] = x;
      }
      if (in_img) {
        if (g.out_fp32) {
          float4* o = reinterpret_cast<float4*>(g.out_f32 + opix + c0);
#pragma unroll
          for (int j = 0; j < CPW / 4; j++) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
          __align__(16) __half2 p0[CPW / 2], p1[CPW / 2];
          if (g.relu) {
#pragma unroll
            for (int j = 0; j < CPW / 2; j++) split2_pos(v[2 * j], v[2 * j + 1], p0[j], p1[j]);
          } else {
            plane_t* hh = reinterpret_cast<plane_t*>(p0);
            plane_t* ll = reinterpret_cast<plane_t*>(p1);
#pragma unroll
            for (int j = 0; j < CPW; j++) split2(v[j], hh[j], ll[j]);
          }
          uint4* o0 = reinterpret_cast<uint4*>(g.out_planes + opix + c0);
          uint4* o1 = reinterpret_cast<uint4*>(g.out_planes + plane_stride + opix + c0);
#pragma unroll
          for (int j = 0; j < CPW / 8; j++) {
            o0[j] = reinterpret_cast<const uint4*>(p0)[j];
            o1[j] = reinterpret_cast<const uint4*>(p1)[j];
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  tc::cluster_sync();       // no CTA leaves while its partner's MMAs / remote arrives may still touch it
  if (warp == 1) p2_tmem_dealloc(tmem_base, 512);
}
