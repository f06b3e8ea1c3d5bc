// conv_tc_s2d.cuh - Conv 3x3 s1 (Cin = 16 -> 32) FUSED with the space_to_depth + 1x1 conv that follows it in the SPD-Conv
// stem of the LPC YAML (layers 1-3: conv.Conv(16, 32, 3, 1) -> space_to_depth -> C2f.cv1, i.e. a 2x2 stride-2 conv over the
// 3x3 conv's output; reference nn/modules/conv.py:36-54, block.py space_to_depth / C2f.forward).
// Included by conv_tc.cu inside its anonymous namespace (shares the PTX helpers, the bias tiles and the epilogue).
//
// Unfused, the 320x320x32 intermediate of a 64-image batch is 419 MB written and 419 MB read back (the two launches are
// 237 us of a 2.0 ms step, both near their own HBM floors).  Here it never leaves the SM:
//   * a SUPERTILE is 16 x 32 intermediate pixels = four 8 x 16 M tiles of the 3x3 conv = one 8 x 16 M tile of the 2x2/s2
//     conv (each of its 128 output pixels consumes a 2x2 block of intermediate pixels: K2 = 4 * 32 = 128);
//   * the 3x3 conv runs exactly as in conv_tc_halo_kernel<16> (TMA halo patch, nine shifted UMMA descriptors, bias MMA,
//     accumulator ring of four 32-column TMEM buffers);
//   * its epilogue (tcgen05.ld -> act -> bf16) does not store to global memory: the thread that owns intermediate pixel
//     (y, x) writes its 32 channels (64 bytes) into row m = (y/2)*8 + x/2, K offset ((y&1)*2 + (x&1))*32 of a K-major,
//     128B-swizzled A2 tile [128 rows x 128 K] in shared memory (two 16 KB K blocks; the layout a TMA box would produce,
//     so the same UMMA descriptors as everywhere else read it; bank-conflict free: the 8 lanes of a store phase hit 8
//     distinct 16-byte chunk columns);
//   * once the four sub-tiles of a supertile have landed (mbarrier a2full, 16 warp arrivals, each after
//     fence.proxy.async), the MMA thread issues the second conv: bias MMA + 8 MMAs (M = 128, N = C2, K = 16) from A2 and
//     the resident W2 into one of two acc2 TMEM buffers;
//   * the two epilogue warp groups take the acc2 tiles alternately: tcgen05.ld -> act2 -> bf16 -> NHWC stores.
// A2 and acc2 are double-buffered; the MMA thread issues conv2(S-1) while it waits for the last patch of supertile S (by then
// every epilogue of S-1 has drained, so the wait is free), which keeps three roles busy without a circular wait:
//   producer: aempty, tempty -> patch            MMA: afull -> conv1 ; a2full, t2empty -> conv2
//   epilogue: a2empty (first sub-tile of a supertile), tfull -> A2 ; t2full -> global
// 2 CTAs per SM: 192 TMEM columns, ~107 KB shared memory (W1 12 KB, W2 8 KB, 2 patches 18 KB, 2 x A2 64 KB).

struct S2dParams {
  int H, W, B;               // the 3x3 conv's input = intermediate size
  int Ho, Wo;                // output of the fused pair (H/2, W/2)
  int st_x, st_y, n_super;   // supertiles per row / column / in total
  float inv_per_img, inv_st_x;
  int n2, acc2_cols, tmem_cols, a_bufs;
  int l2_prefetch;           // the producer pulls the next supertile's patches into L2 ahead of their loads
  int act1, act2;
  const float* bias1;
  const float* bias2;
  bf16* y;
  long long y_ld;
};

constexpr int S2D_N1 = 32, S2D_CIN = 16;
constexpr int S2D_PATCH_BYTES = HALO_PH * HALO_SPW * S2D_CIN * 2;     // 9216
constexpr int S2D_W1_BYTES = 3 * S2D_N1 * 128;                        // K = 144 -> three 64-wide blocks
constexpr int S2D_A2_BYTES = 2 * 128 * 128;                           // two K blocks of [128 rows x 128 B]

template <int ACT>
__device__ __forceinline__ void s2d_act16(const uint32_t* v, int act, uint4& o0, uint4& o1) {
  float f[16];
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const float2 t = make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1]));
    float2 r;
    if (ACT == LPC_ACT_MISH) r = mish2_(t);
    else if (ACT == LPC_ACT_SILU) r = silu2_(t);
    else if (ACT == LPC_ACT_NONE) r = t;
    else r = make_float2(apply_act<false>(t.x, act), apply_act<false>(t.y, act));
    f[i] = r.x;
    f[i + 1] = r.y;
  }
  Vec<bf16> a, b;
  a.pack(f);
  b.pack(f + 8);
  o0 = a.raw;
  o1 = b.raw;
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ACT1 / ACT2 >= 0: compile-time activations; -1: run-time (p.act1 / p.act2)
template <int ACT1, int ACT2>
__global__ void __launch_bounds__(320, 2)
conv_tc_s2d_kernel(const __grid_constant__ TmapPack maps, const __grid_constant__ S2dParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bars[2 * MAX_STAGES + 20];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t ones_addr = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bias1_addr = ones_addr + ONES_BYTES;
  const uint32_t bias2_addr = bias1_addr + S2D_N1 * 32u;
  const uint32_t w1_addr = (bias2_addr + (uint32_t)p.n2 * 32u + 1023u) & ~1023u;
  const uint32_t w2_addr = w1_addr + S2D_W1_BYTES;
  const uint32_t w2_block = (uint32_t)p.n2 * 128u;
  const uint32_t a_region = w2_addr + 2u * w2_block;                       // n2 % 8 == 0 keeps this 1024-aligned
  const uint32_t a2_region = a_region + (uint32_t)p.a_bufs * S2D_PATCH_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto afull_bar = [&](int s) { return bar0 + 8u * s; };
  auto aempty_bar = [&](int s) { return bar0 + 8u * (MAX_STAGES + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + 4 + b); };
  auto a2full_bar = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + 8 + b); };
  auto a2empty_bar = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + 10 + b); };
  auto t2full_bar = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + 12 + b); };
  auto t2empty_bar = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + 14 + b); };
  const uint32_t bfull_bar = bar0 + 8u * (2 * MAX_STAGES + 16);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a[0]);
    prefetch_tmap(&maps.a[1]);
    prefetch_tmap(&maps.b);
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(afull_bar(s), 1);
      mbar_init(aempty_bar(s), 1);
    }
    for (int b = 0; b < 4; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 4);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(a2full_bar(b), 16);          // 4 sub-tiles x 4 epilogue warps
      mbar_init(a2empty_bar(b), 1);
      mbar_init(t2full_bar(b), 1);
      mbar_init(t2empty_bar(b), 4);
    }
    mbar_init(bfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_slot), (uint32_t)p.tmem_cols);
  write_bias_tiles(ones_addr, bias1_addr, p.bias1, 0, S2D_N1);
  write_bias_tiles(ones_addr, bias2_addr, p.bias2, 0, p.n2);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t acc2_base = tmem_base + 4u * S2D_N1;
  pdl_trigger();
  if (warp != 0) pdl_wait();

  const int per_img = p.st_x * p.st_y;
  const int s_first = blockIdx.x, s_step = gridDim.x;

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_expect_tx(bfull_bar, (uint32_t)S2D_W1_BYTES + 2u * w2_block);
      for (int ks = 0; ks < 3; ++ks) tma_load_2d(w1_addr + (uint32_t)(ks * S2D_N1 * 128), &maps.b, bfull_bar, ks * 64, 0);
      for (int ks = 0; ks < 2; ++ks) tma_load_2d(w2_addr + (uint32_t)ks * w2_block, &maps.a[1], bfull_bar, ks * 64, 0);
      pdl_wait();                               // activations of the previous kernel
      // Two patch buffers per CTA is all the shared memory leaves next to the double-buffered A2, i.e. ~2 x 9 KB in flight per
      // CTA against an HBM latency of ~1 us: the patches of the supertile AFTER the current one are pulled into L2 ahead of
      // time (TMA prefetch, no shared memory needed), so the loads below hit L2.
      auto prefetch_super = [&](int s) {
        if (s >= p.n_super || !p.l2_prefetch) return;
        const int img = fast_div(s, per_img, p.inv_per_img);
        const int rem = s - img * per_img;
        const int sy = fast_div(rem, p.st_x, p.inv_st_x);
        const int sx = rem - sy * p.st_x;
        for (int t = 0; t < 4; ++t)
          tma_prefetch_4d(&maps.a[0], 0, sx * (2 * HALO_TW) + (t & 1) * HALO_TW - 1, sy * (2 * HALO_TH) + (t >> 1) * HALO_TH - 1, img);
      };
      prefetch_super(s_first);
      int v = 0;
      for (int s = s_first; s < p.n_super; s += s_step) {
        const int img = fast_div(s, per_img, p.inv_per_img);
        const int rem = s - img * per_img;
        const int sy = fast_div(rem, p.st_x, p.inv_st_x);
        const int sx = rem - sy * p.st_x;
        prefetch_super(s + s_step);
#pragma unroll 1
        for (int t = 0; t < 4; ++t, ++v) {
          const int ab = v % p.a_bufs;
          mbar_wait(aempty_bar(ab), (uint32_t)(((v / p.a_bufs) & 1) ^ 1));
          mbar_wait(tempty_bar(v & 3), (uint32_t)(((v >> 2) & 1) ^ 1));     // "patch full" implies "accumulator drained"
          mbar_expect_tx(afull_bar(ab), (uint32_t)S2D_PATCH_BYTES);
          tma_load_4d(a_region + (uint32_t)(ab * S2D_PATCH_BYTES), &maps.a[0], afull_bar(ab), 0,
                      sx * (2 * HALO_TW) + (t & 1) * HALO_TW - 1, sy * (2 * HALO_TH) + (t >> 1) * HALO_TH - 1, img);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t idesc1 = make_idesc(S2D_N1), idesc2 = make_idesc(p.n2);
      const uint32_t a_hi = desc_hi((uint32_t)(HALO_SPW * S2D_CIN * 2), 6u), b_hi = desc_hi(1024u, 2u);
      const uint32_t w1_lo = desc_lo(w1_addr, 16u), w2_lo = desc_lo(w2_addr, 16u);
      constexpr uint32_t PITCH16 = S2D_CIN * 2 / 16, BBLK16 = S2D_N1 * 128 / 16;
      mbar_wait(bfull_bar, 0);
      auto conv2 = [&](int sc) {
        const int sb = sc & 1;
        const uint32_t ph = (uint32_t)((sc >> 1) & 1);
        mbar_wait(a2full_bar(sb), ph);                 // the four sub-tile epilogues wrote A2[sb]
        mbar_wait(t2empty_bar(sb), ph ^ 1u);           // acc2[sb] drained (supertile sc - 2)
        tc_fence_after();
        const uint32_t acc2 = acc2_base + (uint32_t)(sb * p.acc2_cols);
        umma_bf16(acc2, smem_desc(ones_addr, 16u, 256u, 6u), smem_desc(bias2_addr, 16u, 256u, 6u), idesc2, 0u);
        const uint32_t a2_lo = desc_lo(a2_region + (uint32_t)(sb * S2D_A2_BYTES), 16u);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_acc(acc2, desc64(a2_lo + (uint32_t)((j >> 2) * (128 * 128 / 16) + 2 * (j & 3)), b_hi),
                   desc64(w2_lo + (uint32_t)(j >> 2) * (w2_block >> 4) + 2u * (uint32_t)(j & 3), b_hi), idesc2);
        umma_commit(a2empty_bar(sb));
        umma_commit(t2full_bar(sb));
      };
      int v = 0, sc = 0;
      for (int s = s_first; s < p.n_super; s += s_step, ++sc) {
#pragma unroll 1
        for (int t = 0; t < 4; ++t, ++v) {
          const int ab = v % p.a_bufs;
          mbar_wait(afull_bar(ab), (uint32_t)((v / p.a_bufs) & 1));
          // the producer loaded this patch after the epilogue of (sc-1, 3) drained its accumulator, so every sub-tile of
          // supertile sc-1 is in A2: the waits inside conv2 return at once
          if (t == 3 && sc > 0) conv2(sc - 1);
          tc_fence_after();
          const uint32_t acc = tmem_base + (uint32_t)((v & 3) * S2D_N1);
          const uint32_t a_lo0 = desc_lo(a_region + (uint32_t)(ab * S2D_PATCH_BYTES), 16u);
          umma_bf16(acc, smem_desc(ones_addr, 16u, 256u, 6u), smem_desc(bias1_addr, 16u, 256u, 6u), idesc1, 0u);
#pragma unroll
          for (int j = 0; j < 9; ++j)
            umma_acc(acc, desc64(a_lo0 + (uint32_t)(((j / 3) * HALO_SPW + j % 3) * PITCH16), a_hi),
                     desc64(w1_lo + (uint32_t)(j >> 2) * BBLK16 + 2u * (uint32_t)(j & 3), b_hi), idesc1);
          umma_commit(aempty_bar(ab));
          umma_commit(tfull_bar(v & 3));
        }
      }
      if (sc > 0) conv2(sc - 1);
    }
  } else {
    const int q4 = warp & 3;                 // TMEM lane quarter this warp may read
    const int g = (warp - 2) >> 2;           // epilogue group: sub-tiles t = g, g + 2; acc2 tiles of supertiles sc with (sc & 1) == g
    const int r = q4 * 32 + lane;
    const int ty = r >> 3, tx = r & 7;       // pixel inside an 8 x 16 tile (both convs)
    const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
    // conv2 epilogue of supertile (sc, s): this thread owns output pixel (sy*16 + ty, sx*8 + tx)
    auto epi2 = [&](int sc, int s) {
      const int sb = sc & 1;
      const int img = fast_div(s, per_img, p.inv_per_img);
      const int rem = s - img * per_img;
      const int sy = fast_div(rem, p.st_x, p.inv_st_x);
      const int sx = rem - sy * p.st_x;
      const int oy = sy * HALO_TH + ty, ox = sx * HALO_TW + tx;
      const bool valid = oy < p.Ho && ox < p.Wo;
      bf16* yrow = p.y + (((long long)img * p.Ho + oy) * p.Wo + ox) * p.y_ld;
      mbar_wait(t2full_bar(sb), (uint32_t)((sc >> 1) & 1));
      tc_fence_after();
      const uint32_t trow = acc2_base + (uint32_t)(sb * p.acc2_cols) + lane_off;
      for (int c = 0; c < p.n2; c += 16) {
        uint32_t v0[16];
        tmem_ld16(trow + (uint32_t)c, v0);
        tmem_ld_wait();
        uint4 o0, o1;
        s2d_act16<ACT2>(v0, p.act2, o0, o1);
        if (valid) {
          if ((reinterpret_cast<uintptr_t>(yrow + c) & 31u) == 0) {
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(yrow + c), "r"(o0.x), "r"(o0.y), "r"(o0.z), "r"(o0.w),
                         "r"(o1.x), "r"(o1.y), "r"(o1.z), "r"(o1.w)
                         : "memory");
          } else {
            *reinterpret_cast<uint4*>(yrow + c) = o0;
            *reinterpret_cast<uint4*>(yrow + c + 8) = o1;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t2empty_bar(sb));
    };
    int v = 0, sc = 0, s_prev = -1;
    for (int s = s_first; s < p.n_super; s += s_step, ++sc) {
      const int sb = sc & 1;
      const uint32_t a2 = a2_region + (uint32_t)(sb * S2D_A2_BYTES);
#pragma unroll 1
      for (int t = 0; t < 4; ++t, ++v) {
        if ((t & 1) != g) continue;
        if (t < 2) mbar_wait(a2empty_bar(sb), (uint32_t)(((sc >> 1) & 1) ^ 1));     // conv2(sc - 2) has read A2[sb]
        const int buf = v & 3;
        mbar_wait(tfull_bar(buf), (uint32_t)((v >> 2) & 1));
        tc_fence_after();
        // intermediate pixel (y, x) of the 16-wide x 32-tall supertile -> A2 row m, K quarter q
        const int y = (t >> 1) * HALO_TH + ty, x = (t & 1) * HALO_TW + tx;
        const int m = (y >> 1) * HALO_TW + (x >> 1), q = (y & 1) * 2 + (x & 1);
        const uint32_t row = a2 + (uint32_t)((q >> 1) * (128 * 128) + m * 128);
        const uint32_t trow = tmem_base + (uint32_t)(buf * S2D_N1) + lane_off;
#pragma unroll
        for (int c = 0; c < S2D_N1; c += 16) {
          uint32_t v0[16];
          tmem_ld16(trow + (uint32_t)c, v0);
          tmem_ld_wait();
          uint4 o0, o1;
          s2d_act16<ACT1>(v0, p.act1, o0, o1);
          const int ch = (q & 1) * 4 + (c >> 3);                 // logical 16-byte chunk of the row's 128 bytes
          st_shared_v4(row + (uint32_t)(((ch) ^ (m & 7)) << 4), o0);
          st_shared_v4(row + (uint32_t)(((ch + 1) ^ (m & 7)) << 4), o1);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(tempty_bar(buf));
          mbar_arrive(a2full_bar(sb));
        }
      }
      if (sc > 0 && ((sc - 1) & 1) == g) epi2(sc - 1, s_prev);
      s_prev = s;
    }
    if (sc > 0 && ((sc - 1) & 1) == g) epi2(sc - 1, s_prev);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}
