// conv_tc_pair.cuh - the 3x3 stride-1 halo-patch convolution on CTA PAIRS (tcgen05 cta_group::2).
// Included by conv_tc.cu inside its anonymous namespace (shares ConvTcParams, the epilogue and the tile scheduler).
//
// The small-channel 3x3 layers are bound by the rate of tcgen05.mma INSTRUCTIONS, not by the tensor pipe: a
// cta_group::1 MMA (M = 128) costs ~70 cycles for any N <= 64, a cta_group::2 MMA (M = 256 over two SMs) ~50
// (tools/mma_bench.cu, profiles/r01_i_mma_rates.txt).  So two CTAs of a cluster each own one 8 x 16-pixel M tile (their
// halo patch sits at the same shared-memory offset in both), each keeps HALF of the weight rows resident
// ([n_tile/2][K], which also lets Cout = 128 layers issue N = 128), and the leader CTA's single thread issues M = 256
// MMAs for the pair.  Signalling:
//   patches: each CTA's warp 0 issues its own TMA boxes (cta_group::2 signalling): expect_tx + complete_tx land on the
//            LEADER's afull[s] (count 2)
//   leader MMA --commit.multicast--> aempty[s], tfull[b] in BOTH CTAs
//   epilogue --arrive--> the CTA's OWN tempty[b]; each producer waits for it before loading the tile, so the leader's
//            afull implies both accumulators are drained
//   weights: each CTA TMA-loads its half, then its warp 0 arrives on the leader's wready barrier (count 2)
template <int CIN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 2)
conv_tc_halo2_kernel(const __grid_constant__ TmapPack maps, const __grid_constant__ ConvTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bars[2 * MAX_STAGES + 10];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t rank = cluster_ctarank();
  const int half = p.n_tile >> 1;                                     // weight rows held by this CTA
  const uint32_t ones_addr = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bias_addr = ones_addr + ONES_BYTES;
  const uint32_t smem_base = (bias_addr + (uint32_t)half * 32u + 1023u) & ~1023u;
  const int b_block = half * 128;                                     // bytes of one 64-wide K step of this CTA's half
  const int halo_bytes = p.slabs * p.slab_bytes + p.tail_bytes;      // Cin = 80 / 96: full slab(s) + one narrow tail slab
  const uint32_t a_region = smem_base + (uint32_t)(p.ksteps * b_block);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto afull_bar = [&](int s) { return bar0 + 8u * s; };                       // used in the leader only
  auto aempty_bar = [&](int s) { return bar0 + 8u * (MAX_STAGES + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + 4 + b); };  // leader only
  const uint32_t bfull_bar = bar0 + 8u * (2 * MAX_STAGES + 8);                 // local: this CTA's weights landed
  const uint32_t wready_bar = bar0 + 8u * (2 * MAX_STAGES + 9);                // leader: both halves landed

  const int pair = blockIdx.x >> 1, pairs = gridDim.x >> 1;
  const int n_idx = pair % p.n_tiles;
  const int mp_first = pair / p.n_tiles, mp_step = pairs / p.n_tiles;
  const int n0 = n_idx * p.n_tile;
  const int epi_warps_per_tile = p.epi_alt ? 4 : 4 * p.epi_split;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.b);
    prefetch_tmap(&maps.a[0]);
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(afull_bar(s), 2);
      mbar_init(aempty_bar(s), 1);
    }
    for (int b = 0; b < 4; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), epi_warps_per_tile);
    }
    mbar_init(bfull_bar, 1);
    mbar_init(wready_bar, 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc2(smem_u32(&tmem_base_slot), (uint32_t)p.tmem_cols);
  write_bias_tiles(ones_addr, bias_addr, p.bias, n0 + (int)rank * half, half);
  tc_fence_before();
  cluster_sync_all();             // barriers of both CTAs initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_trigger();
  if (warp != 0) pdl_wait();      // warp 0 only streams weights (constants)

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_expect_tx(bfull_bar, (uint32_t)(p.ksteps * b_block));
      for (int ks = 0; ks < p.ksteps; ++ks) tma_load_2d(smem_base + (uint32_t)(ks * b_block), &maps.b, bfull_bar, ks * 64, n0 + (int)rank * half);
      mbar_wait(bfull_bar, 0);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive_rank(wready_bar, 0);
      {
        // Halo patches by TMA (as in conv_tc_halo_kernel): each CTA loads its own tile's patch into its own shared memory;
        // both report (expect_tx + complete_tx) to the LEADER's afull barrier (count 2), which the MMA thread waits on.
        pdl_wait();
        int tcount = 0;
        for (int mp = mp_first; 2 * mp < p.m_tiles; mp += mp_step, ++tcount) {
          const int m = 2 * mp + (int)rank;
          const bool live = m < p.m_tiles;
          const TileCoord t = tile_coord(p, live ? m : p.m_tiles - 1);     // odd tile count: the idle half re-loads the last tile
          const int ab = tcount % p.a_bufs;
          mbar_wait(aempty_bar(ab), (uint32_t)(((tcount / p.a_bufs) & 1) ^ 1));
          // each CTA's producer checks ITS accumulator buffer (local tempty, arrived by its own epilogue warps) before it
          // loads: the leader's afull (both patches landed) then implies both accumulators are drained - no remote
          // epilogue arrivals and no cluster-scope tempty wait on the MMA thread
          mbar_wait(tempty_bar(tcount & (p.n_acc - 1)), (uint32_t)(((tcount >> p.acc_shift) & 1) ^ 1));
          const uint32_t lead_bar = mapa_rank(afull_bar(ab), 0);
          mbar_expect_tx_cluster(lead_bar, (uint32_t)halo_bytes);
          const uint32_t a_dst = a_region + (uint32_t)(ab * halo_bytes);
          for (int sl = 0; sl < p.slabs; ++sl)
            tma_load_4d_pair(a_dst + (uint32_t)(sl * p.slab_bytes), &maps.a[0], lead_bar, sl * 64, t.x0 - 1, t.y0 - 1, t.img);
          if (p.tail_c)
            tma_load_4d_pair(a_dst + (uint32_t)(p.slabs * p.slab_bytes), &maps.a[1], lead_bar, p.slabs * 64, t.x0 - 1, t.y0 - 1, t.img);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the leader CTA's elected thread, M = 256 over the pair =====
    if (rank == 0 && elect_one_sync()) {
      const uint32_t idesc = make_idesc_m(p.n_tile, 256);
      const uint32_t a_layout = p.pitch == 128 ? 2u : (p.pitch == 64 ? 4u : 6u);
      const uint32_t a_hi = desc_hi((uint32_t)(HALO_SPW * p.pitch), a_layout), b_hi = desc_hi(1024u, 2u);
      const uint32_t pitch16 = (uint32_t)p.pitch >> 4, slab16 = (uint32_t)p.slab_bytes >> 4, bblk16 = (uint32_t)b_block >> 4;
      const int groups = (p.Cin < 64 ? p.Cin : 64) / 16;
      const uint32_t b_lo0 = desc_lo(smem_base, 16u);
      const uint64_t ones_desc = smem_desc(ones_addr, 16u, 256u, 6u), bias_desc = smem_desc(bias_addr, 16u, 256u, 6u);
      mbar_wait_cluster(wready_bar, 0);
      int tcount = 0;
      for (int mp = mp_first; 2 * mp < p.m_tiles; mp += mp_step, ++tcount) {
        const int buf = tcount & (p.n_acc - 1);
        const int ab = tcount % p.a_bufs;
        mbar_wait_cluster(afull_bar(ab), (uint32_t)((tcount / p.a_bufs) & 1));
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(buf * p.acc_cols);
        const uint32_t a_lo0 = desc_lo(a_region + (uint32_t)(ab * halo_bytes), 16u);
        umma2_bf16(acc, ones_desc, bias_desc, idesc, 0u);                  // accumulator := bias
        if (CIN == 80 || CIN == 96) {
          // one full slab (64 channels per pixel row, 128B swizzle) + a narrow tail slab (16 / 32 channels, 32B / 64B swizzle):
          // per tap four slices of the first and TG of the second, in the K order of the dense weights (tap * CIN + c)
          constexpr int TAIL = CIN - 64, TG = TAIL / 16, TP16 = TAIL * 2 / 16, SLAB16 = HALO_PH * HALO_SPW * 128 / 16;
          const uint32_t t_hi = desc_hi((uint32_t)(HALO_SPW * TAIL * 2), TAIL == 32 ? 4u : 6u);
#pragma unroll
          for (int j = 0; j < 9 * (4 + TG); ++j) {
            const int tap = j / (4 + TG), g = j % (4 + TG);
            const int pix = (tap / 3) * HALO_SPW + tap % 3;
            const uint64_t ad = g < 4 ? desc64(a_lo0 + (uint32_t)(pix * 8 + 2 * g), a_hi)
                                      : desc64(a_lo0 + (uint32_t)(SLAB16 + pix * TP16 + 2 * (g - 4)), t_hi);
            umma2_acc(acc, ad, desc64(b_lo0 + (uint32_t)(j >> 2) * bblk16 + 2u * (uint32_t)(j & 3), b_hi), idesc);
          }
        } else if (CIN > 0) {
          constexpr int C_ROW = CIN <= 32 ? CIN : 64, PITCH16 = C_ROW * 2 / 16, GROUPS = (CIN < 64 ? CIN : 64) / 16, SLABS = (CIN + 63) / 64;
          constexpr int SLAB16 = HALO_PH * HALO_SPW * C_ROW * 2 / 16;
#pragma unroll
          for (int j = 0; j < 9 * SLABS * GROUPS; ++j) {
            const int tap = j / (SLABS * GROUPS), sl = (j / GROUPS) % SLABS, g = j % GROUPS;
            const int kin_c = j & 3, ks_c = j >> 2;
            umma2_acc(acc, desc64(a_lo0 + (uint32_t)(((tap / 3) * HALO_SPW + tap % 3) * PITCH16 + sl * SLAB16 + 2 * g), a_hi),
                      desc64(b_lo0 + (uint32_t)ks_c * bblk16 + 2u * (uint32_t)kin_c, b_hi), idesc);
          }
        } else {
          int kin = 0, ks = 0;
          for (int ty = 0; ty < 3; ++ty)
            for (int tx = 0; tx < 3; ++tx) {
              const uint32_t tap_lo = a_lo0 + (uint32_t)(ty * HALO_SPW + tx) * pitch16;
              for (int sl = 0; sl < p.slabs; ++sl)
                for (int g = 0; g < groups; ++g) {
                  umma2_acc(acc, desc64(tap_lo + (uint32_t)sl * slab16 + 2u * (uint32_t)g, a_hi),
                            desc64(b_lo0 + (uint32_t)ks * bblk16 + 2u * (uint32_t)kin, b_hi), idesc);
                  if (++kin == 4) { kin = 0; ++ks; }
                }
            }
        }
        umma2_commit_both(aempty_bar(ab));
        umma2_commit_both(tfull_bar(buf));
      }
    }
  } else {
    // ===== epilogue (local accumulator rows; arrival on the leader's tempty) =====
    const EpiCtx ectx = make_epi_ctx(p, warp, lane);
    int tcount = 0;
    for (int mp = mp_first; 2 * mp < p.m_tiles; mp += mp_step, ++tcount) {
      if (p.epi_alt && (tcount & 1) != ectx.group) continue;
      const int m = 2 * mp + (int)rank;
      const int buf = tcount & (p.n_acc - 1);
      const bool live = m < p.m_tiles;
      const TileCoord t = tile_coord(p, live ? m : 0);
      ResPre rp;
      if (live) res_prefetch(p, ectx, t.img, t.x0, t.y0, n0, rp);
      mbar_wait(tfull_bar(buf), (uint32_t)((tcount >> p.acc_shift) & 1));
      tc_fence_after();
      if (live) epilogue_tile(p, ectx, tmem_base + (uint32_t)(buf * p.acc_cols), t.img, t.x0, t.y0, n0, &rp);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));
    }
  }
  tc_fence_before();
  cluster_sync_all();             // the peer's shared memory / barriers stay alive until both CTAs are done
  if (warp == 1) tmem_dealloc2(tmem_base, (uint32_t)p.tmem_cols);
}

// ---- the per-tap TMA kernel on CTA pairs -------------------------------------------------------------------------
// Same division of labour as conv_tc_halo2_kernel for the layers conv_tc_taps_kernel serves (1x1 on the flat view, 3x3
// stride 2, the s2d fold, 3x3 on small maps): each CTA of the pair streams ITS M tile's activation boxes and ITS half of
// the weight rows of every 64-wide K step through its own stage ring (both report expect_tx + complete_tx to the
// leader's full[s], count 2); the leader's thread issues M = 256 MMAs; tcgen05.commit.multicast frees stage s in both
// CTAs.  Pays off where a tile has many MMAs (K >= 512) or wide N: the MMA instruction rate per SM doubles.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 2)
conv_tc_taps2_kernel(const __grid_constant__ TmapPack maps, const __grid_constant__ ConvTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bars[2 * MAX_STAGES + 8];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t rank = cluster_ctarank();
  const int half = p.n_tile >> 1;
  const uint32_t ones_addr = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bias_addr = ones_addr + ONES_BYTES;
  const uint32_t smem_base = (bias_addr + (uint32_t)half * 32u + 1023u) & ~1023u;
  const int b_stage_bytes = half * 128;
  const int stage_bytes = A_STAGE_BYTES + ((b_stage_bytes + 1023) & ~1023);      // the B half keeps its 1024-byte alignment
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };                          // leader only
  auto empty_bar = [&](int s) { return bar0 + 8u * (MAX_STAGES + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + 4 + b); };   // leader only

  const int pair = blockIdx.x >> 1, pairs = gridDim.x >> 1;
  const int n_idx = pair % p.n_tiles;
  const int mp_first = pair / p.n_tiles, mp_step = pairs / p.n_tiles;
  const int n0 = n_idx * p.n_tile;
  const int epi_warps_per_tile = p.epi_alt ? 4 : 4 * p.epi_split;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a[0]);
    prefetch_tmap(&maps.b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 2);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 4; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), epi_warps_per_tile);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc2(smem_u32(&tmem_base_slot), (uint32_t)p.tmem_cols);
  write_bias_tiles(ones_addr, bias_addr, p.bias, n0 + (int)rank * half, half);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    if (elect_one_sync()) {
      const uint32_t tx_bytes = (uint32_t)(p.TW * p.TH * 128 + b_stage_bytes);
      const int sub_bytes = 256 * p.kc;
      int it = 0, tcount = 0;
      for (int mp = mp_first; 2 * mp < p.m_tiles; mp += mp_step, ++tcount) {
        const int m = 2 * mp + (int)rank;
        const TileCoord t = tile_coord(p, m < p.m_tiles ? m : p.m_tiles - 1);    // odd tile count: the idle half re-loads the last tile
        mbar_wait(tempty_bar(tcount & (p.n_acc - 1)), (uint32_t)(((tcount >> p.acc_shift) & 1) ^ 1));   // local accumulator drained
        for (int ks = 0; ks < p.ksteps; ++ks, ++it) {
          const int s = it % p.stages;
          mbar_wait(empty_bar(s), (uint32_t)(((it / p.stages) & 1) ^ 1));
          const uint32_t a_dst = smem_base + (uint32_t)(s * stage_bytes);
          const uint32_t lead_bar = mapa_rank(full_bar(s), 0);
          mbar_expect_tx_cluster(lead_bar, tx_bytes);
          for (int j = 0; j < p.nsub; ++j) {
            int q = ks * p.nsub + j;
            if (q >= p.real_slots) q = p.real_slots - 1;
            const int tap = q / p.chunks_per_tap;
            const int c0 = (q - tap * p.chunks_per_tap) * p.kc;
            tma_load_4d_pair(a_dst + (uint32_t)(j * sub_bytes), &maps.a[p.tap_map[tap]], lead_bar, c0, t.x0 + p.tap_dx[tap],
                             t.y0 + p.tap_dy[tap], t.img);
          }
          tma_load_2d_pair(a_dst + A_STAGE_BYTES, &maps.b, lead_bar, ks * 64, n0 + (int)rank * half);
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one_sync()) {
      const uint32_t idesc = make_idesc_m(p.n_tile, 256);
      const uint32_t a_layout = p.kc == 64 ? 2u : (p.kc == 32 ? 4u : 6u);
      const uint32_t a_hi = desc_hi((uint32_t)(16 * p.kc), a_layout), b_hi = desc_hi(1024u, 2u);
      const int sub_bytes = 256 * p.kc;
      uint32_t a_off[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = k * 16, j = e / p.kc;
        a_off[k] = (uint32_t)((j * sub_bytes + (e - j * p.kc) * 2) >> 4);
      }
      const int last_real = (min(p.real_slots * p.kc, p.Cin * (p.real_slots / p.chunks_per_tap)) - (p.ksteps - 1) * 64 + 15) / 16;
      const uint64_t ones_desc = smem_desc(ones_addr, 16u, 256u, 6u), bias_desc = smem_desc(bias_addr, 16u, 256u, 6u);
      int it = 0, tcount = 0;
      for (int mp = mp_first; 2 * mp < p.m_tiles; mp += mp_step, ++tcount) {
        const int buf = tcount & (p.n_acc - 1);
        const uint32_t acc = tmem_base + (uint32_t)(buf * p.acc_cols);
        for (int ks = 0; ks < p.ksteps; ++ks, ++it) {
          const int s = it % p.stages;
          mbar_wait_cluster(full_bar(s), (uint32_t)((it / p.stages) & 1));
          tc_fence_after();
          if (ks == 0) umma2_bf16(acc, ones_desc, bias_desc, idesc, 0u);
          const uint32_t a_lo = desc_lo(smem_base + (uint32_t)(s * stage_bytes), 16u);
          const uint32_t b_lo = desc_lo(smem_base + (uint32_t)(s * stage_bytes) + A_STAGE_BYTES, 16u);
          const int nk = (ks == p.ksteps - 1) ? last_real : 4;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < nk) umma2_acc(acc, desc64(a_lo + a_off[k], a_hi), desc64(b_lo + 2u * k, b_hi), idesc);
          umma2_commit_both(empty_bar(s));
        }
        umma2_commit_both(tfull_bar(buf));
      }
    }
  } else {
    const EpiCtx ectx = make_epi_ctx(p, warp, lane);
    int tcount = 0;
    for (int mp = mp_first; 2 * mp < p.m_tiles; mp += mp_step, ++tcount) {
      if (p.epi_alt && (tcount & 1) != ectx.group) continue;
      const int m = 2 * mp + (int)rank;
      const bool live = m < p.m_tiles;
      const TileCoord t = tile_coord(p, live ? m : 0);
      const int buf = tcount & (p.n_acc - 1);
      ResPre rp;
      if (live) res_prefetch(p, ectx, t.img, t.x0, t.y0, n0, rp);
      mbar_wait(tfull_bar(buf), (uint32_t)((tcount >> p.acc_shift) & 1));
      tc_fence_after();
      if (live) epilogue_tile(p, ectx, tmem_base + (uint32_t)(buf * p.acc_cols), t.img, t.x0, t.y0, n0, &rp);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2(tmem_base, (uint32_t)p.tmem_cols);
}
