// 3x3 stride-1 convs with Cin = 64 k, Cout = 64 / 128 n on a CTA PAIR (tcgen05 cta_group::2, M = 256): the channel-generic sibling
// of tc_conv_pair.cuh, included by tc_conv.cu inside its anonymous namespace.
//
// The single-CTA halo kernel moved, per 64-channel chunk and 128-pixel tile, 108 KB of dx-shifted activation copies and 9 x 32 KB
// of weight planes (BN = 128) through L2 -> SM: 396 KB per 6912 clk of MMA = 57 B/clk per SM at the tensor peak, twice what it
// sustained.  Here one dense halo patch per plane and chunk (46 KB, see tc_conv_pair.cuh) serves the nine taps, and the two
// CTAs of a pair each fetch HALF of every weight stage ([b_hi | b_lo] rows of their half of the Cout slice): 46 + 144 = 190 KB per
// chunk and CTA = 27.5 B/clk.  Three MMAs per k-step, all with the operand halves at the same offsets in both CTAs:
//     a_hi x b_hi -> main      a_hi x b_lo -> cross      a_lo x b_hi -> cross          (N = BN, each CTA supplies BN / 2 rows)
// (the N-concatenated [b_hi | b_lo] form of the single-CTA kernels would need the two halves of b_hi at different offsets in the two
// CTAs for the third product).  Roles per CTA: warp 0 weight TMA, warp 1 TMEM alloc (+ all MMAs in the leader CTA), warp 2
// activation TMA, warps 3.. epilogue.  All loads complete on the LEADER's full barriers (TMA .cta_group::2); tcgen05.commit
// multicasts the empty barriers and tmem_full to both CTAs; both CTAs' epilogue warps arrive on the leader's tmem_empty.
template <bool RES> constexpr int ph_epi_warps() { return 16; }   // 19 warps = 96 registers per thread: the residual variant works in 16-column chunks
template <bool RES> constexpr int ph_threads() { return (3 + ph_epi_warps<RES>()) * 32; }
template <int BN> constexpr int ph_w_stage() { return BN * 128; }                 // [b_hi half | b_lo half] of one (chunk, tap)
template <int BN> constexpr int ph_w_stages() { return BN == 128 ? 8 : 12; }
template <int BN> constexpr int ph_bar_off() { return P2_A_BYTES + ph_w_stages<BN>() * ph_w_stage<BN>(); }
template <int BN> constexpr size_t ph_smem_bytes() { return (size_t)ph_bar_off<BN>() + 512 + 1024; }

__device__ __forceinline__ void p2_tma_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          tc::smem_u32(smem_dst)),
      "l"(m), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}

template <int BN, bool RES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ph_threads<RES>(), 1)
tc_conv3x3_halo_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, ConvArgs g, int total_tiles) {
  constexpr int STAGES = ph_w_stages<BN>(), W_STAGE = ph_w_stage<BN>(), HALF = BN / 2;
  constexpr int EPW = ph_epi_warps<RES>(), CPW = 4 * BN / EPW, CH = RES ? 16 : (CPW < 32 ? CPW : 32);   // epilogue warps, columns per warp, chunk
  extern __shared__ uint8_t cv_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)cv_smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                         // [buf][plane][180 px][128 B]
  uint8_t* sW = smem + P2_A_BYTES;            // [stage][b_hi half | b_lo half]
  uint64_t* a_full = (uint64_t*)(smem + ph_bar_off<BN>());   // [2]       (used in the leader CTA)
  uint64_t* a_empty = a_full + 2;             // [2]
  uint64_t* b_full = a_empty + 2;             // [STAGES]  (used in the leader CTA)
  uint64_t* b_empty = b_full + STAGES;        // [STAGES]
  uint64_t* tmem_full = b_empty + STAGES;     // [1]
  uint64_t* tmem_empty = tmem_full + 1;       // [1]       (used in the leader CTA)
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 1);

  const uint32_t rank = tc::cluster_ctarank();
  const int tiles_x = g.W / C64_TW, tiles_y = (g.H + C64_TH - 1) / C64_TH;
  const int n_tiles = g.Cout / BN, chunks = g.Cin / CV_CK;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int n_items = ((total_tiles + 1) / 2) * n_tiles, n_clusters = gridDim.x / 2, cluster = blockIdx.x / 2;
  const int my_iters = (n_items - cluster + n_clusters - 1) / n_clusters;
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmW);
    for (int i = 0; i < 2; i++) { tc::mbar_init(a_full + i, 1); tc::mbar_init(a_empty + i, 1); }
    for (int s = 0; s < STAGES; s++) { tc::mbar_init(b_full + s, 1); tc::mbar_init(b_empty + s, 1); }
    tc::mbar_init(tmem_full, 1);
    tc::mbar_init(tmem_empty, 2 * EPW);
    tc::fence_barrier_init();
  }
  if (warp == 1) p2_tmem_alloc(tmem_slot, 4 * BN);
  tc::fence_before_sync();
  tc::cluster_sync();       // barriers of both CTAs initialised before any remote arrive / multicast commit
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // item = (pair of pixel tiles, Cout slice); the slices of one pair are neighbours in the walk: its activations stay in L2
  auto decode = [&](int it, int& b, int& x0, int& y0, int& n0, bool& valid) {
    const int item = cluster + it * n_clusters;
    int tile = 2 * (item / n_tiles) + (int)rank;
    n0 = (item % n_tiles) * BN;
    valid = tile < total_tiles;
    if (!valid) tile = total_tiles - 1;     // odd tile count: the partner recomputes the last tile and stores nothing
    x0 = (tile % tiles_x) * C64_TW; y0 = ((tile / tiles_x) % tiles_y) * C64_TH; b = tile / (tiles_x * tiles_y);
  };

  if (warp == 2) {            // activation TMA: one dense halo patch per plane and (item, chunk)
    if (lane == 0) {
      for (int i = 0; i < my_iters; i++) {
        int b, x0, y0, n0; bool valid;
        decode(i, b, x0, y0, n0, valid);
        for (int ck = 0; ck < chunks; ck++) {
          const int u = i * chunks + ck, buf = u & 1;
          tc::mbar_wait(a_empty + buf, ((u >> 1) & 1) ^ 1);
          if (rank == 0) tc::mbar_expect_tx(a_full + buf, 2 * NP * P2_HALO_PX * 128);   // both CTAs' patches
          const uint32_t lbar = p2_mapa(a_full + buf, 0);
#pragma unroll
          for (int p = 0; p < NP; p++)
            p2_tma_4d_pair(sA + buf * P2_A_BUF + p * P2_PLANE, &tmA, lbar, ck * CV_CK, x0 - 1, y0 - 1, p * g.B + b);
        }
      }
    }
  } else if (warp == 0) {     // weight TMA: this CTA's half of the Cout slice, one stage per (chunk, tap)
    if (lane == 0) {
      int c = 0;
      for (int i = 0; i < my_iters; i++) {
        int b, x0, y0, n0; bool valid;
        decode(i, b, x0, y0, n0, valid);
        const int row0 = n0 + (int)rank * HALF;
        for (int ck = 0; ck < chunks; ck++)
          for (int tap = 0; tap < 9; tap++, c++) {
            const int s = c % STAGES, ph = (c / STAGES) & 1;
            tc::mbar_wait(b_empty + s, ph ^ 1);
            if (rank == 0) tc::mbar_expect_tx(b_full + s, 2 * W_STAGE);   // both CTAs' halves
            const uint32_t lbar = p2_mapa(b_full + s, 0);
#pragma unroll
            for (int p = 0; p < NP; p++)
              p2_tma_2d_pair(sW + s * W_STAGE + p * HALF * 128, &tmW, lbar, ck * CV_CK, (p * 9 + tap) * g.Cout + row0);
          }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      const bool leader = tc::elect_one();
      constexpr uint32_t idesc = tc::make_idesc(tc::FMT_F16, 256, BN);
      // every operand descriptor = one base descriptor + a compile-time offset (+ the ring stage): see tc_conv_pair.cuh
      const uint64_t a_desc0 = p2_desc(tc::smem_u32(sA), P2_HALO_W * 128), w_desc0 = p2_desc(tc::smem_u32(sW), 1024);
      int c = 0;
      for (int i = 0; i < my_iters; i++) {
        p2_wait_cluster(tmem_empty, (i & 1) ^ 1);   // both CTAs' epilogues have drained the previous item's accumulators
        tc::fence_after_sync();
        for (int ck = 0; ck < chunks; ck++) {
          const int u = i * chunks + ck, buf = u & 1;
          p2_wait_cluster(a_full + buf, (u >> 1) & 1);
          const uint64_t a_d = a_desc0 + (uint64_t)((buf * P2_A_BUF) >> 4);
#pragma unroll
          for (int tap = 0; tap < 9; tap++, c++) {
            const int s = c % STAGES, ph = (c / STAGES) & 1;
            p2_wait_cluster(b_full + s, ph);
            tc::fence_after_sync();
            const uint64_t b_d = w_desc0 + (uint64_t)((s * W_STAGE) >> 4);
            const int step = ck * 9 + tap;
            const uint32_t d_set = tmem_base + (step & 1) * 2 * BN;     // [main | cross] sets alternate step by step
            const uint32_t first = step >= 2 ? 1u : 0u;
            if (leader) {
#pragma unroll
              for (int k = 0; k < 4; k++) {
                const uint64_t a_hi = a_d + (uint64_t)((((tap / 3) * P2_HALO_W + tap % 3) * 128 + k * 32) >> 4);
                const uint64_t a_lo = a_hi + (uint64_t)(P2_PLANE >> 4);
                const uint64_t b_hi = b_d + (uint64_t)((k * 32) >> 4), b_lo = b_hi + (uint64_t)((HALF * 128) >> 4);
                p2_mma(d_set, a_hi, b_hi, idesc, k ? 1u : first);        // -> main
                p2_mma(d_set + BN, a_hi, b_lo, idesc, k ? 1u : first);   // -> cross
                p2_mma(d_set + BN, a_lo, b_hi, idesc, 1u);               // -> cross
              }
              p2_commit(b_empty + s);
            }
            __syncwarp();
          }
          if (leader) p2_commit(a_empty + buf);   // the nine taps of this chunk are done: the patch buffer may take the next one
          __syncwarp();
        }
        if (leader) p2_commit(tmem_full);
        __syncwarp();
      }
    }
  } else {
    // epilogue: EPW / 4 warps per TMEM sub-partition, CPW columns of the slice each, in chunks of CH <= 32.  The accumulators are
    // single-buffered (512 TMEM columns at BN = 128), so this is exposed time: 16 warps without a residual input.
    const int q = warp % 4, cw0 = ((warp - 3) / 4) * CPW;
    const int m = q * 32 + lane;              // pixel index in the tile: row m/8, col m%8
    const int Ho = g.pool ? g.H / 2 : g.H, Wo = g.pool ? g.W / 2 : g.W;
    const size_t plane_stride = (size_t)g.B * Ho * Wo * g.Cout;
    const bool b0 = (lane & 1) != 0, b3 = (lane & 8) != 0;   // pooled outputs: lane owns CH / 4 channels of its 2x2 window
    for (int i = 0; i < my_iters; i++) {
      int b, x0, y0, n0; bool valid;
      decode(i, b, x0, y0, n0, valid);
      tc::mbar_wait(tmem_full, i & 1);
      tc::fence_after_sync();
      const int py = y0 + m / C64_TW, px = x0 + m % C64_TW;
      const int oy = g.pool ? py / 2 : py, ox = g.pool ? px / 2 : px;
      const bool in_img = valid && (py < g.H) && (px < g.W);
      const size_t opix = (((size_t)b * Ho + oy) * Wo + ox) * g.Cout + n0;
#pragma unroll 1
      for (int c0 = cw0; c0 < cw0 + CPW; c0 += CH) {
        float v[CH];
        p2_ld_acc<CH>(tmem_base + ((uint32_t)(q * 32) << 16) + c0, BN, 2 * BN, v);   // (main0 + main1) + (cross0 + cross1) 2^-11
        if (c0 + CH >= cw0 + CPW) {  // this warp's last TMEM read of the item: hand the accumulators back to the MMA warp
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) p2_arrive_remote(p2_mapa(tmem_empty, 0));
        }
        if (!RES && g.pool) {
          // 2x2 max-pool = lanes {l, l^1, l^8} as an exchange-and-halve butterfly; bias and the (monotone) activation commute with
          // the max, so they run on the CH / 4 pooled channels this lane ends up with (tc_conv_pair.cuh)
          float u[CH / 2], r[CH / 4];
#pragma unroll
          for (int j = 0; j < CH / 2; j++) {
            const float keep = b0 ? v[CH / 2 + j] : v[j], send = b0 ? v[j] : v[CH / 2 + j];
            u[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
          }
#pragma unroll
          for (int e = 0; e < CH / 4; e++) {
            const float keep = b3 ? u[CH / 4 + e] : u[e], send = b3 ? u[e] : u[CH / 4 + e];
            r[e] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 8));
          }
          const int cq = c0 + (b0 ? CH / 2 : 0) + (b3 ? CH / 4 : 0);
#pragma unroll
          for (int e = 0; e < CH / 4; e++) {
            float x = r[e] + (g.bias ? g.bias[n0 + cq + e] : 0.f);
            if (g.relu == 1) x = fmaxf(x, 0.f);
            else if (g.relu == 2) x = x > 0.f ? x : 0.01f * x;
            r[e] = x;
          }
          if (in_img) {
            if (g.out_fp32) {
#pragma unroll
              for (int e = 0; e < CH / 4; e += 4) *reinterpret_cast<float4*>(g.out_f32 + opix + cq + e) = make_float4(r[e], r[e + 1], r[e + 2], r[e + 3]);
            } else {
              __align__(16) __half2 hp[CH / 8], lp[CH / 8];
#pragma unroll
              for (int e = 0; e < CH / 8; e++) split2x2(r[2 * e], r[2 * e + 1], hp[e], lp[e]);
              if (CH == 32) {
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
        float t[RES ? CH : 1];
        if (RES && in_img) {  // residual branch of a BasicBlock (added before the activation)
          const plane_t* r0 = g.res_planes + opix + c0;
#pragma unroll
          for (int j = 0; j < (RES ? CH : 0); j += 8) {
            uint4 a = *reinterpret_cast<const uint4*>(r0 + j), bq = *reinterpret_cast<const uint4*>(r0 + plane_stride + j);
            const plane_t *pa = reinterpret_cast<const plane_t*>(&a), *pb = reinterpret_cast<const plane_t*>(&bq);
#pragma unroll
            for (int e = 0; e < 8; e++) t[j + e] = merge2(pa[e], pb[e]);
          }
        } else if (RES) {
#pragma unroll
          for (int j = 0; j < (RES ? CH : 0); j++) t[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < CH; j++) {
          float x = v[j] + (g.bias ? g.bias[n0 + c0 + j] : 0.f);
          if (RES) x += t[RES ? j : 0];
          if (g.relu == 1) x = fmaxf(x, 0.f);
          else if (g.relu == 2) x = x > 0.f ? x : 0.01f * x;
          v[j] = x;
        }
        if (in_img) {
          if (g.out_fp32) {
            float4* o = reinterpret_cast<float4*>(g.out_f32 + opix + c0);
#pragma unroll
            for (int j = 0; j < CH / 4; j++) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            __align__(16) __half2 p0[CH / 2], p1[CH / 2];
#pragma unroll
            for (int j = 0; j < CH / 2; j++) split2x2(v[2 * j], v[2 * j + 1], p0[j], p1[j]);
            uint4* o0 = reinterpret_cast<uint4*>(g.out_planes + opix + c0);
            uint4* o1 = reinterpret_cast<uint4*>(g.out_planes + plane_stride + opix + c0);
#pragma unroll
            for (int j = 0; j < CH / 8; j++) {
              o0[j] = reinterpret_cast<const uint4*>(p0)[j];
              o1[j] = reinterpret_cast<const uint4*>(p1)[j];
            }
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  tc::cluster_sync();       // no CTA leaves while its partner's MMAs / remote arrives / TMA completions may still touch it
  if (warp == 1) p2_tmem_dealloc(tmem_base, 4 * BN);
}
