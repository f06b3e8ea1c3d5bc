// conv_tc.cu - dense convolution as an implicit GEMM on the 5th-gen tensor cores (sm_100a).
//
//   D[M = output pixels, N = Cout] = A[M, K = taps*Cin] * W[N, K]^T,  bf16 operands, fp32 accumulate in TMEM.
//
// Four persistent, warp-specialised kernels share the PTX helpers (tc_ptx.cuh) and the epilogue:
//
//  conv_tc_taps_kernel  (1x1 as a flat GEMM, 3x3 stride 2, the k=2/s=2 space_to_depth fold, 3x3 on small maps)
//     Per 64-wide K step the TMA engine loads, for the current filter tap, a box [kc channels, TW, TH, 1 image] of
//     the NHWC activation into 128B/64B/32B-swizzled shared memory (zero padding and the channel tail of Cin % 64 != 0
//     = TMA out-of-bounds fill; stride-2 convs read through four "parity" tensor maps with doubled W/H strides) plus
//     the [n_tile x 64] weight block.
//  conv_tc_halo_kernel  (3x3 stride 1 - 72 % of LPC-YOLO's FLOPs - for Cin 16 / 32)
//     Per M tile (8 x 16 output pixels) ONE halo patch (18 rows x 16 pixels, all Cin) arrives as one TMA box per
//     64-channel slab, pixel-major with min(Cin,64) channels per row in the 32B/64B/128B-swizzled K-major layout; the
//     nine taps are nine shifted UMMA descriptors into that patch (8-row group = one 8-pixel tile row, SBO = patch row
//     pitch of 16 pixels so that every group starts on a swizzle-atom boundary), so the activation crosses L2->SM 1.6x
//     instead of 9x and no CUDA thread touches an operand byte.  The weights stay resident in shared memory.
//     (An un-swizzled "planar" patch also works but the tensor core reads un-swizzled operands ~8x slower.)
//  conv_tc_halo2_kernel / conv_tc_taps2_kernel  (conv_tc_pair.cuh)
//     The same two on CTA PAIRS (tcgen05 cta_group::2, M = 256 over two SMs): used for every Cin >= 64 halo layer;
//     the per-tap pair kernel is experimental and off by default.
//
// Common structure (two CTAs per SM when TMEM/smem allow): warp 0 = TMA producer, warp 1 = TMEM allocator +
// single-thread tcgen05.mma issuer, warps 2..9 = epilogue.  The accumulator ring in TMEM has 2 or 4 buffers (tmem_full /
// tmem_empty mbarriers), so tile i's epilogue (tcgen05.ld -> packed SiLU/Mish (+gate, +pre-fetched residual) -> bf16 ->
// 16-byte stores into the NHWC output slice) overlaps the main loops of the next tiles; the bias enters through one extra
// K=16 MMA.  All mbarrier waits are bounded: a protocol bug traps instead of hanging the GPU.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int MAX_STAGES = 8;
constexpr int A_STAGE_BYTES = 128 * 64 * 2;  // taps kernel: 128 rows x 64 K-elements of bf16
constexpr int MAX_TAPS = 9;
constexpr int HALO_TW = 8, HALO_TH = 16;
constexpr int HALO_PW = HALO_TW + 2, HALO_PH = HALO_TH + 2;
constexpr int HALO_SPW = 16;             // patch row pitch in pixels: a multiple of 8 so every 8-pixel row group starts on a swizzle-atom boundary

struct TmapPack {
  CUtensorMap a[4];
  CUtensorMap b;
};

struct ConvTcParams {
  int Ho, Wo, B;
  int tiles_x, tiles_y, TW, TH;
  int m_tiles, n_tiles, n_tile, acc_cols, tmem_cols;
  int n_acc, acc_shift;              // TMEM accumulator buffers (2 or 4) and log2 of it
  int Cout, Cin, ksteps, kc, nsub, chunks_per_tap, real_slots, stages;
  int pix_per_img;
  float inv_tiles_per_img, inv_tiles_x, inv_tw;  // reciprocals for the small integer divisions of the tile scheduler
  int a_bufs, b_resident;            // halo kernels: patch buffers in the ring; weights resident (always 1 there)
  int a_tma;                         // halo kernel: 1 = the patch arrives by TMA (one 4-D box per 64-channel slab), no loader warps
  const bf16* x;                     // halo kernel: activation base, pixel pitch, input geometry
  long long x_ld;
  int H, W;
  int pitch, slabs, slab_bytes;      // halo patch: bytes per pixel row (32/64/128), 64-channel slabs
  int tail_c, tail_bytes;            // pair kernel, Cin = 80 / 96: the channels past the last full 64-channel slab sit in a narrow slab of their own (16 / 32 channels per pixel row, its own tensor map: maps.a[1])
  int up_chunks;                     // per-tap kernel, 1x1 over cat[upsample2x(small), skip]: the first up_chunks 64-channel K chunks come from the SMALL map through a 5-D tensor map that repeats every source pixel 2 x 2 (maps.a[1])
  int epi_split;                     // epilogue warps = 4 * epi_split
  int epi_alt;                       // 1: the two epilogue warp groups take alternate tiles (full width each) instead of half the columns of every tile
  int dbg;                           // LPC_TC_DBG bits (profiling only): 1 skip MMAs, 4 skip the epilogue math + stores, 8 trace, 32 skip stores only
  unsigned long long* trace;         // [4 roles][64 tiles][4 stamps] of clock64, CTA 0 only
  signed char tap_map[MAX_TAPS], tap_dx[MAX_TAPS], tap_dy[MAX_TAPS];
  const float* bias;
  const float* chan_scale;
  const bf16* res;
  long long res_ld;
  bf16* y;
  long long y_ld;
  int act;
  unsigned int* rowmax;              // optional: per output pixel, order-preserving key of max_c of the bf16 outputs (fused tail stage 1)
  long long rowmax_img;              // keys between images
  int rowmax_off;                    // first key of this map inside an image's key array
};

// ---- bias through the tensor core ---------------------------------------------------------------------------
// The accumulator is INITIALISED by one extra K=16 MMA: A = "ones" tile (every row = [1, 1, 0...]), B = per output
// channel [bias_hi, bias_lo, 0...] (bf16 hi/lo split: 16 mantissa bits).  Both tiles are un-swizzled K-major core
// matrices written once per CTA with ordinary stores, then published to the async proxy.
constexpr int ONES_BYTES = 128 * 32;
__device__ __forceinline__ void write_bias_tiles(uint32_t ones_addr, uint32_t bias_addr, const float* bias, int n0, int n_tile) {
  // Both tiles are [rows][16 bf16 = 32 B] in the 32B-swizzled K-major layout (16-byte chunk h of row r sits at
  // h ^ ((r >> 2) & 1)); generic-proxy stores through 32-bit shared addresses, published to the async proxy below.
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
    const int row = i >> 3, w = i & 7;
    const int lc = (w >> 2) ^ ((row >> 2) & 1);                            // logical chunk held by this physical word
    const uint32_t v = (lc == 0 && (w & 3) == 0) ? 0x3F803F80u : 0u;       // bf16 (1.0, 1.0) in K = 0, 1
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(ones_addr + (uint32_t)i * 4u), "r"(v) : "memory");
  }
  for (int i = threadIdx.x; i < n_tile * 8; i += blockDim.x) {
    const int row = i >> 3, w = i & 7;
    const int lc = (w >> 2) ^ ((row >> 2) & 1);
    uint32_t v = 0u;
    if (lc == 0 && (w & 3) == 0 && bias) {
      const float b = bias[n0 + row];
      const bf16 hi = __float2bfloat16_rn(b);
      const bf16 lo = __float2bfloat16_rn(b - __bfloat162float(hi));
      v = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
    }
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(bias_addr + (uint32_t)i * 4u), "r"(v) : "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void issue_bias_mma(uint32_t acc, uint32_t ones_addr, uint32_t bias_addr, int n_tile, uint32_t idesc) {
  (void)n_tile;
  umma_bf16(acc, smem_desc(ones_addr, 16u, 256u, 6u), smem_desc(bias_addr, 16u, 256u, 6u), idesc, 0u);
}

// Residual (shortcut) vectors of the first 32 columns of this thread's output row, loaded BEFORE the wait on the
// accumulator: the epilogue is a latency chain (tcgen05.ld -> math -> store) and a global load inside it cost the
// shortcut layers ~30 us each (16->16 160x160 B64: 77 vs 45 us without the shortcut).
struct ResPre {
  uint4 v[4];
};

// ---- epilogue: one accumulator tile (128 rows x n_tile columns) -> NHWC bf16 --------------------------------
template <int ACT>
__device__ __forceinline__ void store16(const uint32_t* v, bf16* yrow, const bf16* rrow, const float* srow, bool no_store = false, bool has_pre = false, uint4 pre0 = uint4(), uint4 pre1 = uint4(),
                                        float* row_max = nullptr) {
  float f[16];
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const float2 t = make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1]));
    float2 r;
    if (ACT == LPC_ACT_MISH) r = mish2_(t);
    else if (ACT == LPC_ACT_SILU) r = silu2_(t);
    else if (ACT == LPC_ACT_NONE) r = t;
    else r = make_float2(apply_act<false>(t.x, ACT), apply_act<false>(t.y, ACT));
    f[i] = r.x;
    f[i + 1] = r.y;
  }
  if (srow) {
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] *= __ldg(srow + i);
  }
  if (rrow) {
    float r8[8];
    Vec<bf16> rv;
    if (has_pre) rv.raw = pre0; else rv = ld_vec<bf16>(rrow);      // pre: residual vectors fetched before the accumulator wait
    rv.unpack(r8);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] += r8[i];
    if (has_pre) rv.raw = pre1; else rv = ld_vec<bf16>(rrow + 8);
    rv.unpack(r8);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[8 + i] += r8[i];
  }
  if (row_max) {
    float m = *row_max;
#pragma unroll
    for (int i = 0; i < 16; ++i) m = fmaxf(m, f[i]);
    *row_max = m;
  }
  Vec<bf16> o, o2;
  o.pack(f);
  o2.pack(f + 8);
  if (no_store) {   // LPC_TC_DBG & 32 (profiling): keep the math alive, drop the global stores
    if ((o.raw.x ^ o2.raw.y) == 0x7fc17fc1u) st_vec<bf16>(yrow, o);
    return;
  }
  // One 32-byte store per lane (STG.E.ENL2.256, sm_100) when the row slice is 32-byte aligned: the lanes of a warp
  // own different pixels, so every store instruction touches 32 separate sectors whatever its width - halving the
  // instruction count halves the L1 wavefronts the epilogue queues (l1tex data-pipe was 60 % busy on the HBM-side layers).
  if ((reinterpret_cast<uintptr_t>(yrow) & 31u) == 0) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(yrow), "r"(o.raw.x), "r"(o.raw.y), "r"(o.raw.z),
                 "r"(o.raw.w), "r"(o2.raw.x), "r"(o2.raw.y), "r"(o2.raw.z), "r"(o2.raw.w)
                 : "memory");
    return;
  }
  st_vec<bf16>(yrow, o);
  st_vec<bf16>(yrow + 8, o2);
}

template <int ACT>
__device__ __forceinline__ void epilogue_cols(uint32_t trow, int c, int c_end, bool valid, bf16* yrow, const bf16* rrow, const float* srow, bool no_store,
                                              const ResPre* pre, float* row_max) {
  const int c_first = c;
  for (; c < c_end; c += 16) {
    uint32_t v0[16];
    tmem_ld16(trow + (uint32_t)c, v0);
    tmem_ld_wait();
    const int ci = (c - c_first) >> 4;
    const bool hp = pre != nullptr && ci < 2;
    uint4 pa = uint4(), pb = uint4();
    if (hp) {                              // static indices + selects: the prefetched vectors stay in registers
      pa = ci == 0 ? pre->v[0] : pre->v[2];
      pb = ci == 0 ? pre->v[1] : pre->v[3];
    }
    if (valid) store16<ACT>(v0, yrow + c, rrow ? rrow + c : nullptr, srow ? srow + c : nullptr, no_store, hp, pa, pb, row_max);
  }
}


// Per-thread epilogue state that does not depend on the tile: computed once, so the per-tile cost is a handful of
// integer ops (the whole SM is instruction-issue bound on the small-channel layers, profiles/r01_e_*).
struct EpiCtx {
  int ty, tx, c0, c_end, group;
  uint32_t lane_off;     // TMEM lane-quarter offset
  bool row_ok;
};
__device__ __forceinline__ EpiCtx make_epi_ctx(const ConvTcParams& p, int warp, int lane) {
  EpiCtx e;
  const int q = warp & 3;              // TMEM lane quarter this warp may read
  const int part = (warp - 2) >> 2;    // column partition when 8 epilogue warps share a tile
  const int r = q * 32 + lane;
  e.ty = fast_div(r, p.TW, p.inv_tw);
  e.tx = r - e.ty * p.TW;
  const int cols = p.epi_alt ? p.n_tile : p.n_tile / p.epi_split;
  e.c0 = p.epi_alt ? 0 : part * cols;
  e.c_end = e.c0 + cols;
  e.group = part;
  e.lane_off = (uint32_t)(q * 32) << 16;
  e.row_ok = e.ty < p.TH && !(p.dbg & 4);
  return e;
}
__device__ __forceinline__ void res_prefetch(const ConvTcParams& p, const EpiCtx& e, int img, int x0, int y0, int n0, ResPre& r) {
  if (!p.res) return;
  const int ox = x0 + e.tx, oy = y0 + e.ty;
  if (!(e.row_ok && ox < p.Wo && oy < p.Ho)) return;
  const long long pix = ((long long)img * p.Ho + oy) * p.Wo + ox;
  const bf16* rrow = p.res + pix * p.res_ld + n0 + e.c0;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (e.c0 + 8 * j < e.c_end) r.v[j] = *reinterpret_cast<const uint4*>(rrow + 8 * j);
}
__device__ __forceinline__ void epilogue_tile(const ConvTcParams& p, const EpiCtx& e, uint32_t tmem_acc, int img, int x0, int y0, int n0, const ResPre* rp = nullptr) {
  const int ox = x0 + e.tx, oy = y0 + e.ty;
  const bool valid = e.row_ok && ox < p.Wo && oy < p.Ho;
  const long long pix = ((long long)img * p.Ho + oy) * p.Wo + ox;
  bf16* yrow = p.y + pix * p.y_ld + n0;
  const bf16* rrow = p.res ? p.res + pix * p.res_ld + n0 : nullptr;
  const float* srow = nullptr;
  if (p.chan_scale) srow = p.chan_scale + (pix / p.pix_per_img) * p.Cout + n0;
  const uint32_t trow = tmem_acc + e.lane_off;
  const bool ns = (p.dbg & 32) != 0;
  const ResPre* pre = (rp && rrow) ? rp : nullptr;
  float rmax = -INFINITY;
  float* rmp = p.rowmax ? &rmax : nullptr;
  switch (p.act) {
    case LPC_ACT_MISH: epilogue_cols<LPC_ACT_MISH>(trow, e.c0, e.c_end, valid, yrow, rrow, srow, ns, pre, rmp); break;
    case LPC_ACT_SILU: epilogue_cols<LPC_ACT_SILU>(trow, e.c0, e.c_end, valid, yrow, rrow, srow, ns, pre, rmp); break;
    case LPC_ACT_NONE: epilogue_cols<LPC_ACT_NONE>(trow, e.c0, e.c_end, valid, yrow, rrow, srow, ns, pre, rmp); break;
    case LPC_ACT_SIGMOID: epilogue_cols<LPC_ACT_SIGMOID>(trow, e.c0, e.c_end, valid, yrow, rrow, srow, ns, pre, rmp); break;
    default: epilogue_cols<LPC_ACT_RELU>(trow, e.c0, e.c_end, valid, yrow, rrow, srow, ns, pre, rmp); break;
  }
  if (p.rowmax && valid) {
    // the thread owns the whole output row (host guarantees one N tile, unsplit columns): max of the ROUNDED outputs =
    // rounding of the max (round-to-nearest is monotonic); key as in tail.cu (order-preserving uint32 of the float)
    const uint32_t u = __float_as_uint(__bfloat162float(__float2bfloat16_rn(rmax)));
    const long long bimg = pix / p.pix_per_img;          // 1x1 convs run on the flat [B*H*W] view: recover (image, pixel)
    p.rowmax[bimg * p.rowmax_img + p.rowmax_off + (pix - bimg * p.pix_per_img)] = u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
  }
}

struct TileCoord {
  int img, x0, y0;
};
__device__ __forceinline__ TileCoord tile_coord(const ConvTcParams& p, int m_idx) {
  const int per_img = p.tiles_x * p.tiles_y;
  TileCoord t;
  t.img = fast_div(m_idx, per_img, p.inv_tiles_per_img);
  const int rem = m_idx - t.img * per_img;
  const int tyi = fast_div(rem, p.tiles_x, p.inv_tiles_x);
  t.x0 = (rem - tyi * p.tiles_x) * p.TW;
  t.y0 = tyi * p.TH;
  return t;
}

// ---- kernel 1: per-tap TMA boxes ----------------------------------------------------------------------------
__global__ void __launch_bounds__(320, 2)
conv_tc_taps_kernel(const __grid_constant__ TmapPack maps, const __grid_constant__ ConvTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bars[2 * MAX_STAGES + 8];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t ones_addr = (smem_u32(smem_raw) + 1023u) & ~1023u;   // [ones | bias | pad] precede the stage ring
  const uint32_t bias_addr = ones_addr + ONES_BYTES;
  const uint32_t smem_base = (bias_addr + (uint32_t)p.n_tile * 32u + 1023u) & ~1023u;
  const int b_stage_bytes = p.n_tile * 128;
  const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (MAX_STAGES + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + 4 + b); };

  const int n_idx = blockIdx.x % p.n_tiles;
  const int m_first = blockIdx.x / p.n_tiles, m_step = gridDim.x / p.n_tiles;
  const int n0 = n_idx * p.n_tile;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a[0]);
    prefetch_tmap(&maps.b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < p.n_acc; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), p.epi_alt ? 4 : 4 * p.epi_split);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_slot), (uint32_t)p.tmem_cols);
  write_bias_tiles(ones_addr, bias_addr, p.bias, n_idx * p.n_tile, p.n_tile);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_trigger();      // the next kernel's prologue may overlap this kernel
  pdl_wait();         // everything above touched only weights / bias / on-chip state

  if (warp == 0) {
    if (elect_one_sync()) {
      const uint32_t tx_bytes = (uint32_t)(p.TW * p.TH * 128 + b_stage_bytes);
      const int sub_bytes = 256 * p.kc;
      int it = 0, tcount = 0;
      for (int m = m_first; m < p.m_tiles; m += m_step, ++tcount) {
        const TileCoord t = tile_coord(p, m);
        // The PRODUCER checks that the tile's accumulator buffer has been drained before it loads the tile's first
        // operands, so "stage full" implies "accumulator free" and the MMA thread - the serial bottleneck, ~185 cycles
        // per already-complete mbarrier wait (tools/mma_bench.cu) - waits on one barrier per K step instead of two.
        mbar_wait(tempty_bar(tcount & (p.n_acc - 1)), (uint32_t)(((tcount >> p.acc_shift) & 1) ^ 1));
        for (int ks = 0; ks < p.ksteps; ++ks, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (uint32_t)((it / p.stages) & 1);
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t a_dst = smem_base + (uint32_t)(s * stage_bytes);
          mbar_expect_tx(full_bar(s), tx_bytes);
          for (int j = 0; j < p.nsub; ++j) {
            int q = ks * p.nsub + j;
            if (q >= p.real_slots) q = p.real_slots - 1;  // padded K: weights are zero there, any finite A will do
            const int tap = q / p.chunks_per_tap;
            const int c0 = (q - tap * p.chunks_per_tap) * p.kc;
            if (p.up_chunks > 0) {
              // 1x1 over cat[upsample2x(small), skip] (kc = 64, one chunk per K step): the box [64 ch, 2, TW/2, 2, TH/2] of the
              // small map lands as the same 128 rows (x + TW * y) an upsampled tile would have
              if (q < p.up_chunks) tma_load_5d(a_dst, &maps.a[1], full_bar(s), c0, 0, t.x0 >> 1, 0, t.img * (p.Ho >> 1) + (t.y0 >> 1));
              else tma_load_4d(a_dst, &maps.a[0], full_bar(s), c0 - p.up_chunks * 64, t.x0, t.y0, t.img);
              continue;
            }
            tma_load_4d(a_dst + (uint32_t)(j * sub_bytes), &maps.a[p.tap_map[tap]], full_bar(s), c0, t.x0 + p.tap_dx[tap],
                        t.y0 + p.tap_dy[tap], t.img);
          }
          tma_load_2d(a_dst + A_STAGE_BYTES, &maps.b, full_bar(s), ks * 64, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t idesc = make_idesc(p.n_tile);
      const uint32_t a_layout = p.kc == 64 ? 2u : (p.kc == 32 ? 4u : 6u);
      const uint32_t a_hi = desc_hi((uint32_t)(16 * p.kc), a_layout), b_hi = desc_hi(1024u, 2u);
      const int sub_bytes = 256 * p.kc;
      uint32_t a_off[4];                       // 16-byte units from the stage base, per K=16 slice
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = k * 16, j = e / p.kc;
        a_off[k] = (uint32_t)((j * sub_bytes + (e - j * p.kc) * 2) >> 4);
      }
      const int last_real = (min(p.real_slots * p.kc, p.Cin * (p.real_slots / p.chunks_per_tap)) - (p.ksteps - 1) * 64 + 15) / 16;   // real K=16 slices of the last step
      int it = 0, tcount = 0;
      for (int m = m_first; m < p.m_tiles; m += m_step, ++tcount) {
        const int buf = tcount & (p.n_acc - 1);
        const uint32_t acc = tmem_base + (uint32_t)(buf * p.acc_cols);
        for (int ks = 0; ks < p.ksteps; ++ks, ++it) {
          const int s = it % p.stages;
          mbar_wait(full_bar(s), (uint32_t)((it / p.stages) & 1));     // implies tempty(buf): see the producer
          tc_fence_after();
          if (ks == 0) issue_bias_mma(acc, ones_addr, bias_addr, p.n_tile, idesc);      // accumulator := bias
          const uint32_t a_lo = desc_lo(smem_base + (uint32_t)(s * stage_bytes), 16u);
          const uint32_t b_lo = desc_lo(smem_base + (uint32_t)(s * stage_bytes) + A_STAGE_BYTES, 16u);
          const int nk = (ks == p.ksteps - 1) ? last_real : 4;           // zero-padded K slices are skipped
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < nk) umma_acc(acc, desc64(a_lo + a_off[k], a_hi), desc64(b_lo + 2u * k, b_hi), idesc);
          umma_commit(empty_bar(s));
        }
        umma_commit(tfull_bar(buf));
      }
    }
  } else {
    const EpiCtx ectx = make_epi_ctx(p, warp, lane);
    int tcount = 0;
    for (int m = m_first; m < p.m_tiles; m += m_step, ++tcount) {
      if (p.epi_alt && (tcount & 1) != ectx.group) continue;   // the other warp group's tile
      const TileCoord t = tile_coord(p, m);
      const int buf = tcount & (p.n_acc - 1);
      ResPre rp;
      res_prefetch(p, ectx, t.img, t.x0, t.y0, n0, rp);
      mbar_wait(tfull_bar(buf), (uint32_t)((tcount >> p.acc_shift) & 1));
      tc_fence_after();
      epilogue_tile(p, ectx, tmem_base + (uint32_t)(buf * p.acc_cols), t.img, t.x0, t.y0, n0, &rp);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

#define TRACE(role, tile, k) do { if ((p.dbg & 8) && blockIdx.x == 0 && (tile) < 64) p.trace[((role) * 64 + (tile)) * 4 + (k)] = clock64(); } while (0)

// ---- kernel 2: 3x3 stride-1 conv from one halo patch per tile ---------------------------------------------
// Warps: 0 = TMA (resident weights once, then one patch box per slab per tile), 1 = MMA issuer, 2..9 = epilogue.
template <int CIN>   // CIN > 0: patch geometry and the MMA issue sequence are compile-time
__global__ void __launch_bounds__(320, 2)
conv_tc_halo_kernel(const __grid_constant__ TmapPack maps, const __grid_constant__ ConvTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bars[4 * MAX_STAGES + 8];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t ones_addr = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bias_addr = ones_addr + ONES_BYTES;
  const uint32_t smem_base = (bias_addr + (uint32_t)p.n_tile * 32u + 1023u) & ~1023u;
  const int b_block = p.n_tile * 128;
  const int b_blocks = p.ksteps;                // the weights stay resident (the host routes everything else to the per-tap kernel)
  const int halo_bytes = p.slabs * p.slab_bytes;
  const uint32_t a_region = smem_base + (uint32_t)(b_blocks * b_block);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto bfull_bar = [&](int s) { return bar0 + 8u * s; };
  auto afull_bar = [&](int s) { return bar0 + 8u * (2 * MAX_STAGES + s); };
  auto aempty_bar = [&](int s) { return bar0 + 8u * (3 * MAX_STAGES + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (4 * MAX_STAGES + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (4 * MAX_STAGES + 4 + b); };

  const int n_idx = blockIdx.x % p.n_tiles;
  const int m_first = blockIdx.x / p.n_tiles, m_step = gridDim.x / p.n_tiles;
  const int n0 = n_idx * p.n_tile;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a[0]);
    prefetch_tmap(&maps.b);
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(bfull_bar(s), 1);
    }
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(afull_bar(s), 1);
      mbar_init(aempty_bar(s), 1);
    }
    for (int b = 0; b < p.n_acc; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), p.epi_alt ? 4 : 4 * p.epi_split);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_slot), (uint32_t)p.tmem_cols);
  write_bias_tiles(ones_addr, bias_addr, p.bias, n_idx * p.n_tile, p.n_tile);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_trigger();      // the next kernel's prologue may overlap this kernel
  if (warp != 0) pdl_wait();   // warp 0 only streams weights (constants); every other role touches activations

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_expect_tx(bfull_bar(0), (uint32_t)(p.ksteps * b_block));
      for (int ks = 0; ks < p.ksteps; ++ks) tma_load_2d(smem_base + (uint32_t)(ks * b_block), &maps.b, bfull_bar(0), ks * 64, n0);
      {
        // The halo patch by TMA: per 64-channel slab ONE 4-D box [min(Cin,64) ch, 16 px, 18 rows, 1 image] at (x0-1, y0-1);
        // the box lands exactly in the pixel-major swizzled slab layout the MMA descriptors view (row pitch 16 pixels) and
        // out-of-image pixels are zero-filled by the TMA unit.  It moves 288 pixels for the 180 the taps read (L2 -> SM
        // traffic only) and replaces ~900 loader warp instructions per tile (profiles/r01_j_ncu_full_conv_16_32.md).
        pdl_wait();                               // activations of the previous kernel
        int tcount = 0;
        for (int m = m_first; m < p.m_tiles; m += m_step, ++tcount) {
          const TileCoord t = tile_coord(p, m);
          const int ab = tcount % p.a_bufs;
          mbar_wait(aempty_bar(ab), (uint32_t)(((tcount / p.a_bufs) & 1) ^ 1));
          // "patch full" must imply "accumulator buffer drained" (the MMA thread waits on afull only)
          mbar_wait(tempty_bar(tcount & (p.n_acc - 1)), (uint32_t)(((tcount >> p.acc_shift) & 1) ^ 1));
          mbar_expect_tx(afull_bar(ab), (uint32_t)halo_bytes);
          const uint32_t a_dst = a_region + (uint32_t)(ab * halo_bytes);
          for (int sl = 0; sl < p.slabs; ++sl)
            tma_load_4d(a_dst + (uint32_t)(sl * p.slab_bytes), &maps.a[0], afull_bar(ab), sl * 64, t.x0 - 1, t.y0 - 1, t.img);
        }
      }
    }
  } else if (warp == 1) {
    // Descriptors are derived arithmetically from kernel parameters and loop counters only (no table look-ups), so the
    // compiler keeps them in the uniform datapath: per MMA a couple of UIADDs + UTCHMMA.  Slice (tap, slab, g) views the
    // patch shifted by (ty, tx) pixels: start = buffer + slab*slab_bytes + (ty*16 + tx)*pitch + g*32 bytes, SBO = one
    // patch row (16 pixels).  base_offset stays 0: the swizzle XOR is a function of the absolute shared-memory address
    // (measured: the PTX-ISA base_offset formula gives wrong results here, base_offset = 0 is bit-exact).
    if (elect_one_sync()) {
      const uint32_t idesc = make_idesc(p.n_tile);
      const uint32_t a_layout = p.pitch == 128 ? 2u : (p.pitch == 64 ? 4u : 6u);
      const uint32_t a_hi = desc_hi((uint32_t)(HALO_SPW * p.pitch), a_layout), b_hi = desc_hi(1024u, 2u);
      const uint32_t pitch16 = (uint32_t)p.pitch >> 4, slab16 = (uint32_t)p.slab_bytes >> 4, bblk16 = (uint32_t)b_block >> 4;
      const int groups = (p.Cin < 64 ? p.Cin : 64) / 16;     // K=16 slices per slab row
      const uint32_t b_lo0 = desc_lo(smem_base, 16u);
      mbar_wait(bfull_bar(0), 0);
      int tcount = 0;
      for (int m = m_first; m < p.m_tiles; m += m_step, ++tcount) {
        const int buf = tcount & (p.n_acc - 1);
        const int ab = tcount % p.a_bufs;
        TRACE(1, tcount, 0);
        TRACE(1, tcount, 1);
        mbar_wait(afull_bar(ab), (uint32_t)((tcount / p.a_bufs) & 1));   // implies tempty(buf): the producer waited for it
        TRACE(1, tcount, 2);
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(buf * p.acc_cols);
        const uint32_t a_lo0 = desc_lo(a_region + (uint32_t)(ab * halo_bytes), 16u);
        issue_bias_mma(acc, ones_addr, bias_addr, p.n_tile, idesc);      // accumulator := bias
        if (CIN > 0) {
          // C_ROW: channels per patch row in shared memory (Cin = 48 rows are zero-filled to 64: 128-byte swizzle);
          // GROUPS: the K=16 slices of a row that carry real channels
          constexpr int C_ROW = CIN <= 32 ? CIN : 64, PITCH16 = C_ROW * 2 / 16, GROUPS = (CIN < 64 ? CIN : 64) / 16, SLABS = (CIN + 63) / 64;
          constexpr int SLAB16 = HALO_PH * HALO_SPW * C_ROW * 2 / 16;
#pragma unroll
          for (int j = 0; j < 9 * SLABS * GROUPS; ++j) {
            const int tap = j / (SLABS * GROUPS), sl = (j / GROUPS) % SLABS, g = j % GROUPS;     // all compile-time
            const int kin_c = j & 3, ks_c = j >> 2;
            if (!(p.dbg & 1))
              umma_acc(acc, desc64(a_lo0 + (uint32_t)(((tap / 3) * HALO_SPW + tap % 3) * PITCH16 + sl * SLAB16 + 2 * g), a_hi),
                       desc64(b_lo0 + (uint32_t)ks_c * bblk16 + 2u * (uint32_t)kin_c, b_hi), idesc);
          }
        } else {
          int kin = 0, ks = 0;
          for (int ty = 0; ty < 3; ++ty)
            for (int tx = 0; tx < 3; ++tx) {
              const uint32_t tap_lo = a_lo0 + (uint32_t)(ty * HALO_SPW + tx) * pitch16;
              for (int sl = 0; sl < p.slabs; ++sl)
                for (int g = 0; g < groups; ++g) {
                  if (!(p.dbg & 1))
                    umma_acc(acc, desc64(tap_lo + (uint32_t)sl * slab16 + 2u * (uint32_t)g, a_hi),
                             desc64(b_lo0 + (uint32_t)ks * bblk16 + 2u * (uint32_t)kin, b_hi), idesc);
                  if (++kin == 4) { kin = 0; ++ks; }
                }
            }
        }
        umma_commit(aempty_bar(ab));
        umma_commit(tfull_bar(buf));
        TRACE(1, tcount, 3);
      }
    }
  } else if (warp < 2 + 4 * p.epi_split) {
    const EpiCtx ectx = make_epi_ctx(p, warp, lane);
    int tcount = 0;
    for (int m = m_first; m < p.m_tiles; m += m_step, ++tcount) {
      if (p.epi_alt && (tcount & 1) != ectx.group) continue;   // the other warp group's tile
      const TileCoord t = tile_coord(p, m);
      const int buf = tcount & (p.n_acc - 1);
      if (threadIdx.x == 64) TRACE(2, tcount, 0);
      ResPre rp;
      res_prefetch(p, ectx, t.img, t.x0, t.y0, n0, rp);
      mbar_wait(tfull_bar(buf), (uint32_t)((tcount >> p.acc_shift) & 1));
      if (threadIdx.x == 64) TRACE(2, tcount, 1);
      tc_fence_after();
      epilogue_tile(p, ectx, tmem_base + (uint32_t)(buf * p.acc_cols), t.img, t.x0, t.y0, n0, &rp);
      if (threadIdx.x == 64) TRACE(2, tcount, 2);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));
      if (threadIdx.x == 64) TRACE(2, tcount, 3);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

#include "conv_tc_pair.cuh"
#include "conv_tc_s2d.cuh"

// ---- host side -------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

int num_sms() { return lpc_num_sms(); }

int pick_kc(int Cin) { return Cin % 64 == 0 ? 64 : (Cin % 32 == 0 ? 32 : 16); }

int pick_ntile(int Cout, int max_tile) {
  for (int nt = (Cout + max_tile - 1) / max_tile; nt <= Cout / 16; ++nt)
    if (Cout % nt == 0 && (Cout / nt) % 16 == 0 && Cout / nt <= max_tile) return Cout / nt;
  return 0;
}

void pick_tile(int Ho, int Wo, int* TW, int* TH) {
  double best = -1;
  int bw = 1, bh = 1;
  for (int tw = 1; tw <= (Wo < 128 ? Wo : 128); ++tw) {
    int th = 128 / tw;
    if (th > Ho) th = Ho;
    if (th < 1) continue;
    const double tiles = (double)((Ho + th - 1) / th) * ((Wo + tw - 1) / tw);
    const double eff = (double)Ho * Wo / (tiles * 128.0);
    // prefer squarer patches among equally efficient ones (less halo re-read from L2): minimise (tw+2)*(th+2)
    const double halo = (double)(tw + 2) * (th + 2);
    const double bhalo = (double)(bw + 2) * (bh + 2);
    if (eff > best + 1e-9 || (eff > best - 1e-9 && halo < bhalo)) { best = eff; bw = tw; bh = th; }
  }
  *TW = bw;
  *TH = bh;
}

int encode_act_map(CUtensorMap* m, const bf16* base, int C, long long Wd, long long Hd, long long Bd, long long sW,
                   long long sH, long long sB, int box_c, int TW, int TH, CUtensorMapSwizzle sw) {
  EncodeTiledFn enc = get_encode();
  if (!enc) LPC_FAIL(LPC_E_CUDA, "conv2d_tc: cuTensorMapEncodeTiled not available");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wd, (cuuint64_t)Hd, (cuuint64_t)Bd};
  cuuint64_t strides[3] = {(cuuint64_t)sW * 2, (cuuint64_t)sH * 2, (cuuint64_t)sB * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)TW, (cuuint32_t)TH, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    LPC_FAIL(LPC_E_CUDA, "conv2d_tc: activation tensor map encode failed (CUresult %d; C=%d W=%lld H=%lld B=%lld box %dx%dx%d)",
             (int)r, C, Wd, Hd, Bd, box_c, TW, TH);
  return LPC_OK;
}

CUtensorMapSwizzle swizzle_of(int kc) {
  return kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

constexpr size_t SMEM_LIMIT = 200 * 1024;   // dynamic smem budget per CTA (of 227 KB)
#define BIAS_REGION(n_tile) ((size_t)ONES_BYTES + (size_t)(n_tile) * 32 + 1024)
int g_force_mode = 0;                       // 0 auto, 1 taps, 2 halo (tests exercise both paths)

}  // namespace

extern "C" int lpc_conv2d_tc_kpad(int Cin, int k) {
  if (Cin <= 0 || Cin % 16 || k < 1 || k > 3) return LPC_E_ARG;
  const int kraw = k * k * Cin;
  return (kraw + 63) / 64 * 64;
}

extern "C" int lpc_conv2d_tc_supported(int Cin, int Cout, int k, int stride, int pad, int x_ld, int y_ld) {
  if (Cin <= 0 || Cin % 16 || Cout <= 0 || Cout % 16) return 0;
  if (!(k == 1 || k == 2 || k == 3)) return 0;
  if (!(stride == 1 || stride == 2)) return 0;
  if (k == 1 && (stride != 1 || pad != 0)) return 0;
  if (k == 2 && (stride != 2 || pad != 0)) return 0;
  if (k == 3 && pad != 1) return 0;
  if (x_ld % 8 || y_ld % 8) return 0;
  if (pick_ntile(Cout, 256) == 0) return 0;
  return 1;
}

extern "C" int lpc_conv2d_tc_set_mode(int mode) {
  const int old = g_force_mode;
  if (mode >= 0 && mode <= 2) g_force_mode = mode;
  return old;
}

extern "C" int lpc_conv2d_tc(const void* x, int x_ld, int B, int H, int W, int Cin, const void* w, const float* bias,
                             int k, int stride, int pad, int Cout, void* y, int y_ld, int act,
                             const float* chan_scale, const void* res, int res_ld, void* stream) {
  return lpc_conv2d_tc_rowmax(x, x_ld, B, H, W, Cin, w, bias, k, stride, pad, Cout, y, y_ld, act, chan_scale, res, res_ld, nullptr, 0, 0, stream);
}

static int conv2d_tc_impl(const void* x, int x_ld, int B, int H, int W, int Cin, const void* w, const float* bias,
                          int k, int stride, int pad, int Cout, void* y, int y_ld, int act,
                          const float* chan_scale, const void* res, int res_ld,
                          unsigned int* rowmax_keys, long long rowmax_img_stride, int rowmax_offset,
                          const void* up_x, int up_ld, int up_c, void* stream);

extern "C" int lpc_conv2d_tc_rowmax(const void* x, int x_ld, int B, int H, int W, int Cin, const void* w, const float* bias,
                                    int k, int stride, int pad, int Cout, void* y, int y_ld, int act,
                                    const float* chan_scale, const void* res, int res_ld,
                                    unsigned int* rowmax_keys, long long rowmax_img_stride, int rowmax_offset, void* stream) {
  return conv2d_tc_impl(x, x_ld, B, H, W, Cin, w, bias, k, stride, pad, Cout, y, y_ld, act, chan_scale, res, res_ld, rowmax_keys,
                        rowmax_img_stride, rowmax_offset, nullptr, 0, 0, stream);
}

// 1x1 conv over cat[upsample2x(x_small), x_skip] without the upsampled tensor: conv_tc.cu taps kernel, `up_chunks`.
extern "C" int lpc_conv1x1_up2cat_tc_supported(int C0, int C1, int Cout, int H, int W, int xs_ld, int xk_ld, int y_ld) {
  static const int on = [] { const char* e = getenv("LPC_TC_UPCAT"); return e ? atoi(e) : 1; }();
  if (!on || C0 <= 0 || C1 <= 0 || C0 % 64 || C1 % 64 || H <= 0 || W <= 0 || (H & 1) || (W & 1) || xs_ld % 8) return 0;
  return lpc_conv2d_tc_supported(C0 + C1, Cout, 1, 1, 0, xk_ld, y_ld);
}

extern "C" int lpc_conv1x1_up2cat_tc(const void* x_small, int xs_ld, int C0, const void* x_skip, int xk_ld, int C1, int B, int H, int W,
                                     const void* w, const float* bias, int Cout, void* y, int y_ld, int act, void* stream) {
  LPC_REQUIRE(x_small && x_skip, "conv1x1_up2cat_tc: null pointer");
  if (!lpc_conv1x1_up2cat_tc_supported(C0, C1, Cout, H, W, xs_ld, xk_ld, y_ld))
    LPC_FAIL(LPC_E_UNSUPPORTED, "conv1x1_up2cat_tc: unsupported shape C0=%d C1=%d Cout=%d H=%d W=%d", C0, C1, Cout, H, W);
  LPC_REQUIRE(aligned16(x_small), "conv1x1_up2cat_tc: pointers must be 16-byte aligned");
  return conv2d_tc_impl(x_skip, xk_ld, B, H, W, C0 + C1, w, bias, 1, 1, 0, Cout, y, y_ld, act, nullptr, nullptr, 0, nullptr, 0, 0, x_small, xs_ld, C0,
                        stream);
}

static int conv2d_tc_impl(const void* x, int x_ld, int B, int H, int W, int Cin, const void* w, const float* bias,
                          int k, int stride, int pad, int Cout, void* y, int y_ld, int act,
                          const float* chan_scale, const void* res, int res_ld,
                          unsigned int* rowmax_keys, long long rowmax_img_stride, int rowmax_offset,
                          const void* up_x, int up_ld, int up_c, void* stream) {
  LPC_REQUIRE(x && w && y, "conv2d_tc: null pointer");
  LPC_REQUIRE(B > 0 && H > 0 && W > 0, "conv2d_tc: bad shape");
  if (!lpc_conv2d_tc_supported(Cin, Cout, k, stride, pad, x_ld, y_ld))
    LPC_FAIL(LPC_E_UNSUPPORTED, "conv2d_tc: unsupported shape Cin=%d Cout=%d k=%d s=%d p=%d x_ld=%d y_ld=%d", Cin, Cout, k, stride, pad, x_ld, y_ld);
  LPC_REQUIRE(aligned16(x) && aligned16(w) && aligned16(y) && aligned16(res), "conv2d_tc: pointers must be 16-byte aligned");
  LPC_REQUIRE(!res || res_ld % 8 == 0, "conv2d_tc: res_ld must be a multiple of 8");
  LPC_REQUIRE(stride == 1 || (H % 2 == 0 && W % 2 == 0), "conv2d_tc: stride-2 needs even H, W");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const bf16* xb = (const bf16*)x;
  EncodeTiledFn enc = get_encode();
  if (!enc) LPC_FAIL(LPC_E_CUDA, "conv2d_tc: cuTensorMapEncodeTiled not available");

  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  TmapPack maps;
  memset(&maps, 0, sizeof(maps));
  const int ntaps = k * k;
  const int kpad = lpc_conv2d_tc_kpad(Cin, k);
  p.Cin = Cin;
  p.Cout = Cout;
  p.ksteps = kpad / 64;
  p.pix_per_img = Ho * Wo;
  p.bias = bias;
  p.chan_scale = chan_scale;
  p.res = (const bf16*)res;
  p.res_ld = res_ld;
  p.y = (bf16*)y;
  p.y_ld = y_ld;
  p.act = act;
  p.rowmax = rowmax_keys;
  p.rowmax_img = rowmax_img_stride;
  p.rowmax_off = rowmax_offset;
  { const char* e = getenv("LPC_TC_DBG"); p.dbg = e ? atoi(e) : 0; }

  // ---- choose the kernel -------------------------------------------------------------------------------
  bool halo = false, pair = false, tpair = false;   // pair: CTA-pair halo kernel, tpair: CTA-pair per-tap kernel
  // Cin = 48 (yolov10m) takes the halo kernel with 64-channel patch rows: the TMA box runs past the tensor's channel
  // extent and is zero-filled, the MMA loop walks only the three real 16-channel groups of every tap (the per-tap kernel
  // needs 27 boxes of 16 channels in the slow 32-byte-swizzled layout per tile: 48->48 @160x160 B256 took 920 us).
  static const int halo48 = [] { const char* e = getenv("LPC_TC_HALO48"); return e ? atoi(e) : 1; }();
  // Cin = 80 / 96 (yolov10x / m): one full 64-channel slab + a narrow 16- / 32-channel slab, CTA-pair kernel only (the weights of
  // these layers do not fit next to two patches in one CTA); LPC_TC_MIXED=0 sends them back to the per-tap kernel
  static const int mixed_env = [] { const char* e = getenv("LPC_TC_MIXED"); return e ? atoi(e) : 1; }();
  const bool mixed = mixed_env && (Cin == 80 || Cin == 96);
  if (k == 3 && stride == 1 && g_force_mode != 1 && (Cin == 16 || Cin == 32 || (Cin == 48 && halo48) || Cin % 64 == 0 || mixed) && Cin <= 256) {
    const long long tiles = (long long)((Wo + HALO_TW - 1) / HALO_TW) * ((Ho + HALO_TH - 1) / HALO_TH);
    const double eff = (double)Ho * Wo / (double)(tiles * 128);
    const int pitch = (Cin > 32 ? 64 : Cin) * 2;
    const size_t halo_bytes = mixed ? (size_t)HALO_PH * HALO_SPW * (128 + (Cin - 64) * 2) : (size_t)((Cin + 63) / 64) * HALO_PH * HALO_SPW * pitch;
    // weights resident if they fit next to two halo buffers (halving the N tile once if that makes them fit: the
    // activation patch is then loaded twice, still far cheaper than re-streaming the weights for every tile),
    // else a 3-block ring with full-width N tiles
    int nt = pick_ntile(Cout, 256);
    size_t need = (size_t)p.ksteps * nt * 128 + 2 * halo_bytes;
    int resident = 1;
    if (need > SMEM_LIMIT && nt % 32 == 0 && (size_t)p.ksteps * (nt / 2) * 128 + 2 * halo_bytes <= SMEM_LIMIT) {
      nt /= 2;
      need = (size_t)p.ksteps * nt * 128 + 2 * halo_bytes;
    }
    if (need > SMEM_LIMIT) resident = 0;          // weights do not fit next to two patches: the per-tap TMA kernel streams them
    // CTA pairs (cta_group::2, conv_tc_pair.cuh): full-width N with HALF the weight rows resident per CTA
    {
      static const int pair_env = [] { const char* e = getenv("LPC_TC_PAIR"); return e ? atoi(e) : 1; }();
      const int ntp = pick_ntile(Cout, 256);
      const size_t bp = (size_t)p.ksteps * (ntp / 2) * 128;
      const long long tot_tiles = tiles * B;
      // Worth it where the tile is MMA-instruction bound (Cin >= 64: >= 36 MMAs per tile); the Cin 16 / 32 layers are
      // bound by per-tile role latency and lose to the cluster-scope hops.  Measured with TMA patches in both kernels
      // (us, single vs pair): 16->32 169 / 265, 32->64 104 / 124, 32->32 17.5 / 21.2, 64->64@80 49.3 / 44.1,
      // 64->64@40 17.9 / 16.2, 64->128 (N = 64 twice) 121 / 66.  LPC_TC_PAIR=2 forces pairs wherever they fit (tests).
      // Cin = 48 on CTA pairs: measured 48->48 @160x160 B256 (us, single / pair): 460 / 575 plain, 700 / 675 with the residual
      // add -> off by default (LPC_TC_PAIR48=1 switches it on; gpurun_out/prof_m256_p48.txt)
      static const int pair48 = [] { const char* e = getenv("LPC_TC_PAIR48"); return e ? atoi(e) : 0; }();
      const bool wanted = (pair_env == 2 && Cin != 48) || Cin >= 64 || (Cin == 48 && pair48);      // (mixed: Cin >= 64)
      if (pair_env && wanted && (eff >= 0.7 || g_force_mode == 2) && bp + 2 * halo_bytes <= SMEM_LIMIT && tot_tiles >= 4) {
        halo = pair = true;
        p.n_tile = ntp;
        p.b_resident = 1;
        const size_t budget = (ntp <= 128 && bp + 3 * halo_bytes <= 100 * 1024) ? 100 * 1024 : SMEM_LIMIT;
        int ab = (int)((budget - bp) / halo_bytes);
        p.a_bufs = ab > MAX_STAGES ? MAX_STAGES : (ab < 2 ? 2 : ab);
      }
    }
    if (!pair && !mixed && resident && (eff >= 0.7 || g_force_mode == 2) && need <= SMEM_LIMIT) {
      halo = true;
      p.n_tile = nt;
      p.b_resident = resident;
      // deep activation prefetch: small-C layers are HBM-bound and need tens of KB in flight per SM
      const size_t b_bytes = need - 2 * halo_bytes;
      const size_t budget = (nt <= 128 && b_bytes + 3 * halo_bytes <= 100 * 1024) ? 100 * 1024 : SMEM_LIMIT;  // 2 CTAs/SM when small
      int ab = (int)((budget - b_bytes) / halo_bytes);
      p.a_bufs = ab > MAX_STAGES ? MAX_STAGES : (ab < 2 ? 2 : ab);
    }
  }
  if (!halo) {
    // N tile of the per-tap kernel.  The widest tile minimises activation re-reads and MMA instructions, but (a) a layer with few M
    // tiles (batch 1; 20x20 maps) then leaves most SMs idle - 128->256 k3 @40x40 at batch 1 is 13 tiles of N = 256 on 148 SMs,
    // each walking 18 K steps alone - and (b) N = 256 needs both TMEM accumulators of an SM, i.e. one resident CTA.  Narrow
    // the tile while the launch has fewer work items than the GPU has CTA slots (never below 64 columns).
    // LPC_TC_NTILE_MAX caps the width outright (measurement).
    static const int nt_cap = [] { const char* e = getenv("LPC_TC_NTILE_MAX"); return e ? atoi(e) : 256; }();
    static const int nt_auto = [] { const char* e = getenv("LPC_TC_NTILE_AUTO"); return e ? atoi(e) : 1; }();
    int nt = pick_ntile(Cout, nt_cap >= 16 ? nt_cap : 256);
    if (nt == 0) nt = pick_ntile(Cout, 256);
    if (nt_auto) {
      const long long m_est = (k == 1) ? ((long long)B * H * W + 127) / 128
                                       : (long long)B * (((H + 2 * pad - k) / stride + 1) * ((W + 2 * pad - k) / stride + 1) + 127) / 128;
      while (nt > 64 && nt % 32 == 0 && m_est * (Cout / nt) < (long long)num_sms() * (nt > 128 ? 1 : 2)) nt /= 2;
    }
    p.n_tile = nt;
  }
  p.n_tiles = Cout / p.n_tile;
  p.acc_cols = 32;
  while (p.acc_cols < p.n_tile) p.acc_cols <<= 1;
  p.n_acc = 2;
  p.acc_shift = 1;
  p.tmem_cols = 2 * p.acc_cols;
  p.epi_split = (p.n_tile % 32 == 0) ? 2 : 1;

  size_t smem = 0;
  if (halo) {
    p.B = B; p.Ho = Ho; p.Wo = Wo; p.TW = HALO_TW; p.TH = HALO_TH;
    p.tiles_x = (Wo + HALO_TW - 1) / HALO_TW;
    p.tiles_y = (Ho + HALO_TH - 1) / HALO_TH;
    p.x = xb; p.x_ld = x_ld; p.H = H; p.W = W;
    p.pitch = (Cin > 32 ? 64 : Cin) * 2;
    p.slabs = (Cin + 63) / 64;
    p.slab_bytes = HALO_PH * HALO_SPW * p.pitch;
    const bool mixed_slabs = Cin > 64 && Cin % 64 != 0;          // only reached through the pair kernel (selection above)
    if (mixed_slabs) {
      p.slabs = Cin / 64;                          // full slabs; the tail slab follows them
      p.tail_c = Cin % 64;
      p.tail_bytes = HALO_PH * HALO_SPW * p.tail_c * 2;
    }
    smem = (size_t)p.ksteps * (pair ? p.n_tile / 2 : p.n_tile) * 128 + (size_t)p.a_bufs * ((size_t)p.slabs * p.slab_bytes + p.tail_bytes) + 1024 + BIAS_REGION(p.n_tile);
    {
      p.a_tma = 1;                                // both halo kernels receive their patches by TMA
      const int cb = Cin > 32 ? 64 : Cin;
      if (int e = encode_act_map(&maps.a[0], xb, Cin, W, H, B, x_ld, (long long)W * x_ld, (long long)H * W * x_ld, cb, HALO_SPW, HALO_PH, swizzle_of(cb))) return e;
      if (mixed_slabs)
        if (int e = encode_act_map(&maps.a[1], xb, Cin, W, H, B, x_ld, (long long)W * x_ld, (long long)H * W * x_ld, p.tail_c, HALO_SPW, HALO_PH, swizzle_of(p.tail_c))) return e;
    }
  } else {
    // 1x1: always 64-channel boxes in the 128B-swizzled layout; when Cin is not a multiple of 64 the last box runs past the
    // tensor's channel extent and TMA zero-fills the rest (the packed weights are zero there too), e.g. Cin = 80:
    // 2 requests per tile instead of 5 boxes of 16 channels in the slow 32B-swizzled layout.
    p.kc = (k == 1) ? 64 : pick_kc(Cin);
    p.nsub = 64 / p.kc;
    p.chunks_per_tap = (Cin + p.kc - 1) / p.kc;
    p.real_slots = ntaps * p.chunks_per_tap;
    const CUtensorMapSwizzle sw = swizzle_of(p.kc);
    if (k == 1 && up_x) {
      // spatial 16 x 8 tiles instead of the flat view: the skip half through a 4-D map over its C1 channels, the upsampled half
      // through a 5-D map over the SMALL tensor [C0, 2 (stride 0), W/2, 2 (stride 0), B*H/2]
      p.B = B; p.Ho = Ho; p.Wo = Wo; p.TW = 16; p.TH = 8;
      p.tiles_x = (Wo + 15) / 16; p.tiles_y = (Ho + 7) / 8;
      p.up_chunks = up_c / 64;
      if (int e = encode_act_map(&maps.a[0], xb, Cin - up_c, W, H, B, x_ld, (long long)W * x_ld, (long long)H * W * x_ld, 64, 16, 8, sw)) return e;
      {
        const int Ws = W / 2, Hs = H / 2;
        cuuint64_t dims[5] = {(cuuint64_t)up_c, 2, (cuuint64_t)Ws, 2, (cuuint64_t)B * Hs};
        cuuint64_t strides[4] = {0, (cuuint64_t)up_ld * 2, 0, (cuuint64_t)Ws * up_ld * 2};
        cuuint32_t box[5] = {64, 2, 8, 2, 4};
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&maps.a[1], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(up_x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) LPC_FAIL(LPC_E_UNSUPPORTED, "conv1x1_up2cat_tc: the repeating (stride-0) tensor map was refused (CUresult %d)", (int)r);
      }
    } else if (k == 1) {
      const long long M = (long long)B * H * W;
      LPC_REQUIRE(M < (1ll << 31), "conv2d_tc: too many pixels");
      p.B = 1; p.Ho = 1; p.Wo = (int)M; p.TW = 128; p.TH = 1;
      p.tiles_x = (int)((M + 127) / 128); p.tiles_y = 1;
      if (int e = encode_act_map(&maps.a[0], xb, Cin, M, 1, 1, x_ld, (long long)M * x_ld, (long long)M * x_ld, p.kc, 128, 1, sw)) return e;
    } else {
      p.B = B; p.Ho = Ho; p.Wo = Wo;
      pick_tile(Ho, Wo, &p.TW, &p.TH);
      p.tiles_x = (Wo + p.TW - 1) / p.TW;
      p.tiles_y = (Ho + p.TH - 1) / p.TH;
      if (stride == 1) {
        for (int t = 0; t < ntaps; ++t) { p.tap_map[t] = 0; p.tap_dx[t] = (signed char)(t % k - pad); p.tap_dy[t] = (signed char)(t / k - pad); }
        if (int e = encode_act_map(&maps.a[0], xb, Cin, W, H, B, x_ld, (long long)W * x_ld, (long long)H * W * x_ld, p.kc, p.TW, p.TH, sw)) return e;
      } else {
        for (int py = 0; py < 2; ++py)
          for (int px = 0; px < 2; ++px)
            if (int e = encode_act_map(&maps.a[py * 2 + px], xb + ((long long)py * W + px) * x_ld, Cin, (W - px + 1) / 2, (H - py + 1) / 2, B,
                                       2ll * x_ld, 2ll * W * x_ld, (long long)H * W * x_ld, p.kc, p.TW, p.TH, sw))
              return e;
        for (int t = 0; t < ntaps; ++t) {
          const int oy = t / k - pad, ox = t % k - pad;
          const int py = ((oy % 2) + 2) % 2, px = ((ox % 2) + 2) % 2;
          p.tap_map[t] = (signed char)(py * 2 + px);
          p.tap_dy[t] = (signed char)((oy - py) / 2);
          p.tap_dx[t] = (signed char)((ox - px) / 2);
        }
      }
    }
    {
      // The CTA-pair variant of this kernel (conv_tc_taps2_kernel) is correct (the parity tests pass with LPC_TC_TPAIR=2)
      // but OFF by default: its full/empty handshake crosses the cluster once per 64-wide K step (4 MMAs), and that
      // round trip costs more than the doubled MMA rate returns - measured B=64 (us, single / pair): 192->64 k3 40x40
      // 46 / 102, 128->256 k3 40x40 63 / 124, 512->256 1x1 20x20 15 / 23, 256->64 1x1 80x80 44 / 67.  The halo pair
      // kernel hands over once per TILE (>= 36 MMAs) and wins.  Next step: several K steps per barrier round.
      static const int tp_env = [] { const char* e = getenv("LPC_TC_TPAIR"); return e ? atoi(e) : 0; }();
      const long long m_est = (long long)p.tiles_x * p.tiles_y * p.B;
      const bool wanted = tp_env == 2 || p.ksteps >= 8 || (p.n_tile >= 128 && p.ksteps >= 2);
      tpair = tp_env && wanted && p.n_tile % 16 == 0 && m_est >= 4 && !up_x;
    }
    const int stage_bytes = A_STAGE_BYTES + (tpair ? ((p.n_tile / 2 * 128 + 1023) & ~1023) : p.n_tile * 128);
    // two CTAs per SM when the double-buffered accumulators leave TMEM room, else one CTA with a deeper ring
    // ... unless the launch has no more work items than SMs (batch 1, small maps): every CTA then walks its K loop alone, the
    // second resident CTA would stay empty, and the loop's speed is (stages in flight) / (TMA latency) - take the deep ring.
    static const int deep_env = [] { const char* e = getenv("LPC_TC_DEEP_RING"); return e ? atoi(e) : 1; }();
    const long long items_est = (long long)p.tiles_x * p.tiles_y * p.B * (Cout / p.n_tile);
    const bool few_items = deep_env && items_est <= num_sms();
    const size_t budget = (p.tmem_cols <= 256 && !few_items) ? 100 * 1024 : SMEM_LIMIT;
    int stages = (int)(budget / stage_bytes);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) stages = 2;
    p.stages = stages;
    smem = (size_t)stages * stage_bytes + 1024 + BIAS_REGION(p.n_tile);
  }
  p.m_tiles = p.tiles_x * p.tiles_y * p.B;
  LPC_REQUIRE(p.m_tiles < (1 << 24), "conv2d_tc: too many tiles");
  p.inv_tiles_per_img = 1.0f / (float)(p.tiles_x * p.tiles_y);
  p.inv_tiles_x = 1.0f / (float)p.tiles_x;
  p.inv_tw = 1.0f / (float)p.TW;
  {
    cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)kpad * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)((pair || tpair) ? p.n_tile / 2 : p.n_tile)};   // a CTA of a pair loads its half of the rows
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&maps.b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) LPC_FAIL(LPC_E_CUDA, "conv2d_tc: weight tensor map encode failed (CUresult %d)", (int)r);
  }
  static unsigned long long attr_set = 0;     // per device
  if (lpc_first_on_device(&attr_set)) {
    cudaError_t e1 = cudaFuncSetAttribute(conv_tc_taps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT + 16 * 1024);
    if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(conv_tc_taps2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT + 16 * 1024);
    cudaError_t e2 = cudaSuccess;
    const int lim = (int)SMEM_LIMIT + 16 * 1024;
#define HALO_ATTR(C_) if (cudaFuncSetAttribute(conv_tc_halo_kernel<C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim) != cudaSuccess) e2 = cudaErrorUnknown;
#define PAIR_ATTR(C_) if (cudaFuncSetAttribute(conv_tc_halo2_kernel<C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim) != cudaSuccess) e2 = cudaErrorUnknown;
    PAIR_ATTR(0) PAIR_ATTR(16) PAIR_ATTR(32) PAIR_ATTR(48) PAIR_ATTR(64) PAIR_ATTR(128) PAIR_ATTR(80) PAIR_ATTR(96)
#undef PAIR_ATTR
    HALO_ATTR(0) HALO_ATTR(16) HALO_ATTR(32) HALO_ATTR(48) HALO_ATTR(64) HALO_ATTR(128)
#undef HALO_ATTR
    if (e1 != cudaSuccess || e2 != cudaSuccess) { attr_set = 0; LPC_FAIL(LPC_E_CUDA, "conv2d_tc: smem attribute: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2)); }
  }
  // Four accumulator buffers when TMEM allows without costing a resident CTA: with two, the per-buffer chain
  // MMA(t) -> epilogue(t) -> MMA(t+2) makes the tile period (M + E) / 2 instead of max(M, E / groups).
  {
    static const int force = [] { const char* e = getenv("LPC_TC_NACC"); return e ? atoi(e) : 0; }();
    const bool one_cta = smem > 110 * 1024;
    if (4 * p.acc_cols <= 256 || (one_cta && 4 * p.acc_cols <= 512)) p.n_acc = 4;
    if (force == 2 || (force == 4 && 4 * p.acc_cols <= 512)) p.n_acc = force;
    p.acc_shift = p.n_acc == 4 ? 2 : 1;
    p.tmem_cols = p.n_acc * p.acc_cols;
  }
  // LPC_CTA_CAP=1 (measurement): one CTA per SM per launch, so that kernels of two independent chains (detect(streams=2)) can be
  // co-resident on every SM
  static const int cta_cap = [] { const char* e = getenv("LPC_CTA_CAP"); return e ? atoi(e) : 2; }();
  const int ctas_per_sm = (p.tmem_cols <= 256 && smem <= 110 * 1024 && cta_cap >= 2) ? 2 : 1;
  int per_n = (num_sms() * ctas_per_sm) / p.n_tiles;
  if (per_n < 1) per_n = 1;
  if (per_n > p.m_tiles) per_n = p.m_tiles;
  if (pair || tpair) {                          // per_n counts CTA PAIRS here; each pair walks two M tiles at a time
    per_n = (num_sms() * ctas_per_sm / 2) / p.n_tiles;
    if (per_n < 1) per_n = 1;
    if (per_n > (p.m_tiles + 1) / 2) per_n = (p.m_tiles + 1) / 2;
  }
  const unsigned grid = (unsigned)(per_n * p.n_tiles) * ((pair || tpair) ? 2u : 1u);
  // Each role is latency-bound per tile (~1.5-2k cycles: barrier round trips, tcgen05.ld, MUFU chains), so with enough
  // tiles per CTA the two epilogue warp groups take alternate tiles (two epilogues in flight) instead of splitting the
  // columns of one; LPC_TC_EPI_ALT=0/1 overrides (profiling).
  {
    static const int force = [] { const char* e = getenv("LPC_TC_EPI_ALT"); return e ? atoi(e) : -1; }();
    const bool many_tiles = p.m_tiles / (per_n * ((pair || tpair) ? 2 : 1)) >= 4;
    p.epi_alt = (p.epi_split == 2 && many_tiles) ? 1 : 0;
    if (force >= 0 && p.epi_split == 2) p.epi_alt = force;
    // N tiles that cannot be split between two warp groups (Cout = 16, 48, 80 ...: half a tile is not a multiple of 16
    // columns) ran with ONE epilogue group - 6 warps per CTA, one epilogue latency chain per CTA (16->16 @160x160 B64:
    // 18 % warps active, 53 / 75 us against a 16 / 24 us floor).  Alternating whole tiles needs no column split.
    if (p.epi_split == 1 && many_tiles && force != 0) { p.epi_split = 2; p.epi_alt = 1; }
  }
  const unsigned threads = 64 + 128 * p.epi_split;
  static unsigned long long* trace_buf = nullptr;
  if (p.dbg & 8) {
    if (!trace_buf) cudaMalloc(&trace_buf, 4 * 64 * 4 * 8);
    cudaMemset(trace_buf, 0, 4 * 64 * 4 * 8);
    p.trace = trace_buf;
  }
  if (p.rowmax && !(p.n_tiles == 1 && (p.epi_split == 1 || p.epi_alt)))
    LPC_FAIL(LPC_E_UNSUPPORTED, "conv2d_tc_rowmax: needs one N tile whose columns are not split between epilogue warps (Cout=%d)", Cout);
  if (halo) {
    LPC_REQUIRE((long long)H * W * x_ld < (1ll << 31), "conv2d_tc: image too large for 32-bit offsets");
    const unsigned th = threads;
    cudaStream_t st = (cudaStream_t)stream;
#define HALO_LAUNCH(K_) switch (Cin) {                                               \
      case 16: lpc_launch_pdl(K_<16>, grid, th, smem, st, maps, p); break;              \
      case 32: lpc_launch_pdl(K_<32>, grid, th, smem, st, maps, p); break;              \
      case 64: lpc_launch_pdl(K_<64>, grid, th, smem, st, maps, p); break;              \
      case 128: lpc_launch_pdl(K_<128>, grid, th, smem, st, maps, p); break;            \
      default: lpc_launch_pdl(K_<0>, grid, th, smem, st, maps, p); break;               \
    }
    if (!pair && Cin == 48) lpc_launch_pdl(conv_tc_halo_kernel<48>, grid, th, smem, st, maps, p);
    else if (pair && Cin == 48) lpc_launch_pdl(conv_tc_halo2_kernel<48>, grid, th, smem, st, maps, p);
    else if (pair && Cin == 80) lpc_launch_pdl(conv_tc_halo2_kernel<80>, grid, th, smem, st, maps, p);
    else if (pair && Cin == 96) lpc_launch_pdl(conv_tc_halo2_kernel<96>, grid, th, smem, st, maps, p);
    else if (pair) { HALO_LAUNCH(conv_tc_halo2_kernel) } else { HALO_LAUNCH(conv_tc_halo_kernel) }
#undef HALO_LAUNCH
  }
  else if (tpair)
    lpc_launch_pdl(conv_tc_taps2_kernel, grid, threads, smem, (cudaStream_t)stream, maps, p);
  else
    lpc_launch_pdl(conv_tc_taps_kernel, grid, threads, smem, (cudaStream_t)stream, maps, p);
  LPC_CHECK_LAUNCH("conv2d_tc");
  if ((p.dbg & 8) && halo) {
    static int dumps = 0;
    cudaDeviceSynchronize();
    if (dumps++ == 5) {
      unsigned long long h[4 * 64 * 4];
      cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
      const unsigned long long t0 = h[0];
      for (int t = 0; t < 40; ++t) {
        printf("tile %2d MMAx: fence %6lld bias %6lld loop %6lld c1 %6lld\n", t, (long long)(h[(3 * 64 + t) * 4 + 0] - t0), (long long)(h[(3 * 64 + t) * 4 + 1] - t0), (long long)(h[(3 * 64 + t) * 4 + 2] - t0), (long long)(h[(3 * 64 + t) * 4 + 3] - t0));
        printf("tile %2d  LD: w0 %6lld w1 %6lld iss %6lld sig %6lld | MMA: s %6lld te %6lld af %6lld done %6lld | EPI: s %6lld tf %6lld ep %6lld ar %6lld\n", t,
               (long long)(h[(0 * 64 + t) * 4 + 0] - t0), (long long)(h[(0 * 64 + t) * 4 + 1] - t0), (long long)(h[(0 * 64 + t) * 4 + 2] - t0), (long long)(h[(0 * 64 + t) * 4 + 3] - t0),
               (long long)(h[(1 * 64 + t) * 4 + 0] - t0), (long long)(h[(1 * 64 + t) * 4 + 1] - t0), (long long)(h[(1 * 64 + t) * 4 + 2] - t0), (long long)(h[(1 * 64 + t) * 4 + 3] - t0),
               (long long)(h[(2 * 64 + t) * 4 + 0] - t0), (long long)(h[(2 * 64 + t) * 4 + 1] - t0), (long long)(h[(2 * 64 + t) * 4 + 2] - t0), (long long)(h[(2 * 64 + t) * 4 + 3] - t0));
      }
      fflush(stdout);
    }
  }
  return LPC_OK;
}

// ---- Conv 3x3 s1 fused with the following space_to_depth + 1x1 conv (= 2x2 stride-2 conv), conv_tc_s2d.cuh -------------
extern "C" int lpc_conv3x3_s2d_tc_supported(int Cin, int C1, int C2, int H, int W, int x_ld, int y_ld) {
  static const int on = [] { const char* e = getenv("LPC_TC_S2D"); return e ? atoi(e) : 1; }();
  if (!on) return 0;
  if (Cin != S2D_CIN || C1 != S2D_N1) return 0;
  if (C2 <= 0 || C2 % 16 || C2 > 64) return 0;
  if (H <= 0 || W <= 0 || H % 2 || W % 2) return 0;
  if (x_ld % 8 || y_ld % 8) return 0;
  return 1;
}

extern "C" int lpc_conv3x3_s2d_tc(const void* x, int x_ld, int B, int H, int W, int Cin, const void* w1, const float* bias1, int C1,
                                  int act1, const void* w2, const float* bias2, int C2, int act2, void* y, int y_ld, void* stream) {
  LPC_REQUIRE(x && w1 && w2 && y, "conv3x3_s2d_tc: null pointer");
  LPC_REQUIRE(B > 0, "conv3x3_s2d_tc: bad shape");
  if (!lpc_conv3x3_s2d_tc_supported(Cin, C1, C2, H, W, x_ld, y_ld))
    LPC_FAIL(LPC_E_UNSUPPORTED, "conv3x3_s2d_tc: unsupported shape Cin=%d C1=%d C2=%d H=%d W=%d x_ld=%d y_ld=%d", Cin, C1, C2, H, W, x_ld, y_ld);
  LPC_REQUIRE(aligned16(x) && aligned16(w1) && aligned16(w2) && aligned16(y), "conv3x3_s2d_tc: pointers must be 16-byte aligned");
  LPC_REQUIRE((long long)H * W * x_ld < (1ll << 31), "conv3x3_s2d_tc: image too large for 32-bit offsets");
  EncodeTiledFn enc = get_encode();
  if (!enc) LPC_FAIL(LPC_E_CUDA, "conv3x3_s2d_tc: cuTensorMapEncodeTiled not available");
  S2dParams p;
  memset(&p, 0, sizeof(p));
  TmapPack maps;
  memset(&maps, 0, sizeof(maps));
  p.H = H; p.W = W; p.B = B;
  p.Ho = H / 2; p.Wo = W / 2;
  p.st_x = (W + 2 * HALO_TW - 1) / (2 * HALO_TW);
  p.st_y = (H + 2 * HALO_TH - 1) / (2 * HALO_TH);
  const long long n_super = (long long)p.st_x * p.st_y * B;
  LPC_REQUIRE(n_super < (1 << 22), "conv3x3_s2d_tc: too many tiles");
  p.n_super = (int)n_super;
  p.inv_per_img = 1.0f / (float)(p.st_x * p.st_y);
  p.inv_st_x = 1.0f / (float)p.st_x;
  p.n2 = C2;
  p.acc2_cols = C2 <= 32 ? 32 : 64;
  p.tmem_cols = 256;
  p.act1 = act1; p.act2 = act2;
  { static const int pf = [] { const char* e = getenv("LPC_TC_S2D_PREFETCH"); return e ? atoi(e) : 0; }(); p.l2_prefetch = pf; }   // measured: 1.957 vs 1.930 ms per step with / without - off
  p.bias1 = bias1; p.bias2 = bias2;
  p.y = (bf16*)y; p.y_ld = y_ld;
  const size_t fixed = (size_t)ONES_BYTES + S2D_N1 * 32 + (size_t)C2 * 32 + 1024 + S2D_W1_BYTES + 2 * (size_t)C2 * 128 + 2 * S2D_A2_BYTES + 1024;
  {
    static const int ab_env = [] { const char* e = getenv("LPC_TC_S2D_ABUFS"); return e ? atoi(e) : 0; }();
    int ab = (int)(((size_t)110 * 1024 - fixed) / S2D_PATCH_BYTES);          // two CTAs per SM
    if (ab_env > 0) ab = ab_env;
    p.a_bufs = ab > MAX_STAGES ? MAX_STAGES : (ab < 2 ? 2 : ab);
  }
  const size_t smem = fixed + (size_t)p.a_bufs * S2D_PATCH_BYTES;
  if (int e = encode_act_map(&maps.a[0], (const bf16*)x, Cin, W, H, B, x_ld, (long long)W * x_ld, (long long)H * W * x_ld, Cin, HALO_SPW, HALO_PH,
                             swizzle_of(Cin)))
    return e;
  auto weight_map = [&](CUtensorMap* m, const void* w, int kpad, int rows) -> int {
    cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kpad * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) LPC_FAIL(LPC_E_CUDA, "conv3x3_s2d_tc: weight tensor map encode failed (CUresult %d)", (int)r);
    return LPC_OK;
  };
  if (int e = weight_map(&maps.b, w1, lpc_conv2d_tc_kpad(Cin, 3), C1)) return e;          // [32][192]
  if (int e = weight_map(&maps.a[1], w2, lpc_conv2d_tc_kpad(C1, 2), C2)) return e;        // [C2][128], K = (ky, kx, c)
  static unsigned long long attr_set = 0;     // per device
  if (lpc_first_on_device(&attr_set)) {
    const int lim = (int)SMEM_LIMIT + 16 * 1024;
    cudaError_t e1 = cudaFuncSetAttribute(conv_tc_s2d_kernel<LPC_ACT_SILU, LPC_ACT_MISH>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(conv_tc_s2d_kernel<-1, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    if (e1 != cudaSuccess) { attr_set = 0; LPC_FAIL(LPC_E_CUDA, "conv3x3_s2d_tc: smem attribute: %s", cudaGetErrorString(e1)); }
  }
  const int ctas_per_sm = smem <= 110 * 1024 ? 2 : 1;
  long long grid = (long long)num_sms() * ctas_per_sm;
  if (grid > n_super) grid = n_super;
  if (act1 == LPC_ACT_SILU && act2 == LPC_ACT_MISH)
    lpc_launch_pdl(conv_tc_s2d_kernel<LPC_ACT_SILU, LPC_ACT_MISH>, (unsigned)grid, 320u, smem, (cudaStream_t)stream, maps, p);
  else
    lpc_launch_pdl(conv_tc_s2d_kernel<-1, -1>, (unsigned)grid, 320u, smem, (cudaStream_t)stream, maps, p);
  LPC_CHECK_LAUNCH("conv3x3_s2d_tc");
  return LPC_OK;
}
