// dwpw_tc.cu - depthwise 3x3 (stride 1, pad 1) -> pointwise 1x1 [-> pointwise 1x1] in ONE kernel (sm_100a).
//
// The v10Detect class branch (head.py:504-505: [dw3x3(x) + 1x1(x -> c3)] -> [dw3x3(c3) + 1x1] -> 1x1(c3 -> nc)) and the
// CIB blocks (block.py:735-756) are chains of a bandwidth-bound depthwise conv feeding a 1x1 conv on the same pixels.
// Unfused, every link writes its map to HBM and the next one reads it back: five launches and ~10 passes over an
// 80x80 map per image at the P3 level of the LPC head (139 us of the 2.04 ms step at B = 64, against a 38 us floor for
// reading the input once and writing the output once).  Here the depthwise result never leaves the SM:
//
//   warp 8   TMA producer: per tile (8 x 16 output pixels) and per 64-channel block ONE halo box [64 ch, 10 px, 18 rows]
//            of the NHWC input (out-of-image pixels / channels zero-filled = the conv padding), ring of patch slots
//   warps 0-3 depthwise workers: dw 3x3 on the patch (bf16 x bf16 -> fp32 FMAs on packed words, runs of four adjacent
//            pixels per thread: 18 LDS.128 for 4 x 8 outputs), bias + activation, bf16, written as the K-major,
//            128B-swizzled A tile [128 pixels x 64 channels] of the pointwise GEMM (ring of A slots)
//   warp 9   one elected thread issues tcgen05.mma M128 x N=C1 x K16 per 16 channels against the resident 1x1 weights;
//            two accumulators in TMEM, GEMM 1 of tile i is issued before GEMM 2 of tile i - 1
//   warps 4-7 epilogue (a thread owns one pixel): tcgen05.ld -> + bias -> activation -> either bf16 NHWC stores, or
//            (second pointwise stage) bf16 into the A operand of GEMM 2 (same accumulator, re-used) -> epilogue 2 ->
//            stores, optionally with the per-pixel max-logit key of the fused v10 tail (lpc_conv2d_tc_rowmax's contract).
//
// The first version ran all phases of a tile in order on eight do-everything warps: 10.7 k warp instructions per tile
// (activation switch per element pair, slot = counter % n arithmetic) at 39 % issue utilisation - slower than the
// unfused chain (profiles/r02_d_ncu_dwpw_first.md).  Dedicated roles keep the depthwise FMAs and the epilogue math of
// different tiles in flight together; activations are template parameters.  All mbarrier waits are bounded.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int DP_TW = 8, DP_TH = 16;                   // output tile: 8 x 16 pixels = 128 GEMM rows, row r = ty * 8 + tx
constexpr int DP_PW = DP_TW + 2, DP_PH = DP_TH + 2;    // halo patch
constexpr int DP_PATCH_BYTES = DP_PW * DP_PH * 128;    // 64 channels (128 B) per pixel
constexpr int DP_A_BYTES = 128 * 128;                  // one 64-channel K block of the A operand
constexpr int DP_MAX_NA = 4, DP_MAX_NP = 8;

struct DwPwParams {
  int B, H, W, Cin, C1, C2;
  int kb1, kb2;                      // 64-channel K blocks of GEMM 1 (Cin) and GEMM 2 (C1)
  int cin_pad;                       // kb1 * 64
  int tiles_x, tiles_y, ntiles;
  float inv_tiles_x, inv_tiles_y;
  int dw_act, act1, act2;
  int tmem_cols, acc_cols;
  int na, np;                        // A ring slots, patch ring slots
  const float* dw_w;                 // [9][Cin] fp32
  const float* dw_b;                 // [Cin] or null
  const float* b1;
  const float* b2;
  bf16* y;
  long long y_ld;
  unsigned int* rowmax;
  long long rowmax_img;
  int rowmax_off;
  unsigned off_w1, off_w2, off_a, off_a2, off_patch, off_dww, off_dwb, off_b1, off_b2;   // from the 1024-aligned base
};

__device__ __forceinline__ void dp_fhfma2(float& a0, float& a1, uint32_t x, uint32_t w) {
  asm("{\n\t.reg .b16 xl, xh, wl, wh;\n\t"
      "mov.b32 {xl, xh}, %2;\n\t"
      "mov.b32 {wl, wh}, %3;\n\t"
      "fma.rn.f32.bf16 %0, xl, wl, %0;\n\t"
      "fma.rn.f32.bf16 %1, xh, wh, %1;\n\t}"
      : "+f"(a0), "+f"(a1)
      : "r"(x), "r"(w));
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
template <int ACT> __device__ __forceinline__ float2 actT(float2 v) {
  if (ACT == LPC_ACT_SILU) return silu2_(v);
  if (ACT == LPC_ACT_MISH) return mish2_(v);
  return v;
}
__device__ __forceinline__ uint32_t pack2(float2 v) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// One accumulator chunk (16 columns of this thread's row): + bias, activation -> eight packed bf16 pairs; m = running max
template <int ACT>
__device__ __forceinline__ void epi_chunk(const uint32_t* v, uint32_t bias_addr, uint32_t* w, float& m) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 bb = lds_f4(bias_addr + 16u * i);
    const float2 o0 = actT<ACT>(make_float2(__uint_as_float(v[4 * i]) + bb.x, __uint_as_float(v[4 * i + 1]) + bb.y));
    const float2 o1 = actT<ACT>(make_float2(__uint_as_float(v[4 * i + 2]) + bb.z, __uint_as_float(v[4 * i + 3]) + bb.w));
    m = fmaxf(m, fmaxf(fmaxf(o0.x, o0.y), fmaxf(o1.x, o1.y)));
    w[2 * i] = pack2(o0);
    w[2 * i + 1] = pack2(o1);
  }
}

// Warp roles: 0-3 depthwise workers, 4-7 epilogue, 8 TMA producer, 9 MMA issuer.  ACT: activation of the depthwise conv
// and of the first pointwise conv (the head and CIB use one activation for both); the second pointwise conv (B2B) is
// a plain conv (+ bias).
template <int ACT, bool B2B>
__global__ void __launch_bounds__(320, 2)
dwpw_tc_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap w1map, const __grid_constant__ CUtensorMap w2map,
               const __grid_constant__ DwPwParams p) {
  extern __shared__ __align__(1024) unsigned char dp_smem[];
  __shared__ __align__(8) unsigned long long bars[2 * DP_MAX_NA + 2 * DP_MAX_NP + 9];
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(dp_smem) + 1023u) & ~1023u;
  const uint32_t w1s = base + p.off_w1, w2s = base + p.off_w2, a_ring = base + p.off_a, a2s = base + p.off_a2, patch = base + p.off_patch;
  const uint32_t dww = base + p.off_dww, dwb = base + p.off_dwb, b1s = base + p.off_b1, b2s = base + p.off_b2;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_free = [&](int s) { return bar0 + 8u * (DP_MAX_NA + s); };
  auto p_full = [&](int s) { return bar0 + 8u * (2 * DP_MAX_NA + s); };
  auto p_empty = [&](int s) { return bar0 + 8u * (2 * DP_MAX_NA + DP_MAX_NP + s); };
  const uint32_t bx = bar0 + 8u * (2 * DP_MAX_NA + 2 * DP_MAX_NP);
  const uint32_t w_full = bx;
  auto acc1_full = [&](int b) { return bx + 8u * (1 + b); };
  auto acc_free = [&](int b) { return bx + 8u * (3 + b); };
  auto acc2_full = [&](int b) { return bx + 8u * (5 + b); };
  auto a2_full = [&](int b) { return bx + 8u * (7 + b); };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    prefetch_tmap(&xmap);
    prefetch_tmap(&w1map);
    if (B2B) prefetch_tmap(&w2map);
    for (int s = 0; s < DP_MAX_NA; ++s) { mbar_init(a_full(s), 4); mbar_init(a_free(s), 1); }
    for (int s = 0; s < DP_MAX_NP; ++s) { mbar_init(p_full(s), 1); mbar_init(p_empty(s), 4); }
    mbar_init(w_full, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(acc1_full(b), 1); mbar_init(acc_free(b), 4); mbar_init(acc2_full(b), 1); mbar_init(a2_full(b), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)p.tmem_cols);
  // constants into shared memory: depthwise taps (fp32 -> bf16, zero beyond Cin), biases
  for (int i = tid; i < 9 * (p.cin_pad / 2); i += blockDim.x) {
    const int tap = i / (p.cin_pad / 2), c = (i - tap * (p.cin_pad / 2)) * 2;
    float2 wf = make_float2(0.f, 0.f);
    if (c < p.Cin) wf = __ldg(reinterpret_cast<const float2*>(p.dw_w + (long long)tap * p.Cin + c));
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(dww + (uint32_t)(tap * p.cin_pad + c) * 2u), "r"(pack2(wf)) : "memory");
  }
  for (int i = tid; i < p.cin_pad; i += blockDim.x)
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(dwb + 4u * i), "f"((p.dw_b && i < p.Cin) ? __ldg(p.dw_b + i) : 0.f) : "memory");
  for (int i = tid; i < p.C1; i += blockDim.x)
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(b1s + 4u * i), "f"(p.b1 ? __ldg(p.b1 + i) : 0.f) : "memory");
  for (int i = tid; i < p.C2; i += blockDim.x)
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(b2s + 4u * i), "f"(p.b2 ? __ldg(p.b2 + i) : 0.f) : "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int kb1 = p.kb1, kb2 = p.kb2, na = p.na, np = p.np;
  pdl_trigger();

  if (warp == 8) {
    // ===== TMA producer: resident pointwise weights once, then the (tile, channel block) patch stream =====
    if (elect_one_sync()) {
      mbar_expect_tx(w_full, (uint32_t)(kb1 * p.C1 * 128 + kb2 * p.C2 * 128));
      for (int kb = 0; kb < kb1; ++kb) tma_load_2d(w1s + (uint32_t)(kb * p.C1 * 128), &w1map, w_full, kb * 64, 0);
      if (B2B)
        for (int kb = 0; kb < kb2; ++kb) tma_load_2d(w2s + (uint32_t)(kb * p.C2 * 128), &w2map, w_full, kb * 64, 0);
      pdl_wait();                                 // activations of the previous kernel
      int ps = 0;
      uint32_t pph = 1;                           // parity to wait for on p_empty: "previous phase" on a fresh barrier
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        const int r1 = fast_div(t, p.tiles_x, p.inv_tiles_x), tx = t - r1 * p.tiles_x;
        const int n = fast_div(r1, p.tiles_y, p.inv_tiles_y), ty = r1 - n * p.tiles_y;
        for (int kb = 0; kb < kb1; ++kb) {
          mbar_wait(p_empty(ps), pph);
          mbar_expect_tx(p_full(ps), (uint32_t)DP_PATCH_BYTES);
          tma_load_4d(patch + (uint32_t)(ps * DP_PATCH_BYTES), &xmap, p_full(ps), kb * 64, tx * DP_TW - 1, ty * DP_TH - 1, n);
          if (++ps == np) { ps = 0; pph ^= 1u; }
        }
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer: GEMM 1 of tile i, then (B2B) GEMM 2 of tile i - 1, so the epilogue of a tile overlaps the next GEMM 1 =====
    if (elect_one_sync()) {
      const uint32_t idesc1 = make_idesc(p.C1), idesc2 = make_idesc(B2B ? p.C2 : 16);
      const uint32_t hi = desc_hi(1024u, 2u);                  // K-major, 128B swizzle, 8-row groups 1024 B apart (A and W alike)
      const uint32_t a2_lo = desc_lo(a2s, 16u);
      mbar_wait(w_full, 0);
      int as = 0;
      uint32_t aph = 0;
      int i = 0;
      auto gemm2 = [&](int it) {
        const int b = it & 1;
        mbar_wait(a2_full(b), (uint32_t)((it >> 1) & 1));
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(b * p.acc_cols);
        for (int kb = 0; kb < kb2; ++kb) {
          const uint32_t a_lo = a2_lo + (uint32_t)((kb * DP_A_BYTES) >> 4), b_lo = desc_lo(w2s + (uint32_t)(kb * p.C2 * 128), 16u);
          const int nk = min(4, (p.C1 - kb * 64 + 15) >> 4);
          for (int k = 0; k < nk; ++k) umma_bf16(acc, desc64(a_lo + 2u * k, hi), desc64(b_lo + 2u * k, hi), idesc2, (uint32_t)((kb | k) != 0));
        }
        umma_commit(acc2_full(b));
      };
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++i) {
        const int b = i & 1;
        mbar_wait(acc_free(b), (uint32_t)(((i >> 1) & 1) ^ 1));       // the epilogue has drained this accumulator (tile i - 2)
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(b * p.acc_cols);
        for (int kb = 0; kb < kb1; ++kb) {
          mbar_wait(a_full(as), aph);
          tc_fence_after();
          const uint32_t a_lo = desc_lo(a_ring + (uint32_t)(as * DP_A_BYTES), 16u), b_lo = desc_lo(w1s + (uint32_t)(kb * p.C1 * 128), 16u);
          const int nk = min(4, (p.Cin - kb * 64 + 15) >> 4);
          for (int k = 0; k < nk; ++k) umma_bf16(acc, desc64(a_lo + 2u * k, hi), desc64(b_lo + 2u * k, hi), idesc1, (uint32_t)((kb | k) != 0));
          umma_commit(a_free(as));
          if (++as == na) { as = 0; aph ^= 1u; }
        }
        umma_commit(acc1_full(b));
        if (B2B && i > 0) gemm2(i - 1);
      }
      if (B2B && i > 0) gemm2(i - 1);
    }
  } else if (warp < 4) {
    // ===== depthwise workers: patch -> dw 3x3 + bias + act -> bf16 A tile =====
    pdl_wait();
    const int j = tid & 7;                        // 16-byte chunk (8 channels) of the 64-channel block
    const int dty = tid >> 3;                     // 0..15: tile row; the thread computes its 8 pixels in two runs of four
    const uint32_t patch_off = (uint32_t)(dty * DP_PW * 128 + j * 16);
    int ps = 0, as = 0;
    uint32_t pph = 0, aph = 1;
    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
      for (int kb = 0; kb < kb1; ++kb) {
        const bool live = kb * 64 + j * 8 < p.Cin;              // chunks beyond Cin are never read by the MMAs
        uint4 wreg[9];
        const uint32_t wsrc = dww + (uint32_t)(kb * 64 + j * 8) * 2u;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) wreg[tap] = lds128(wsrc + (uint32_t)(tap * p.cin_pad) * 2u);
        const float4 bv0 = lds_f4(dwb + 4u * (uint32_t)(kb * 64 + j * 8)), bv1 = lds_f4(dwb + 4u * (uint32_t)(kb * 64 + j * 8 + 4));
        mbar_wait(p_full(ps), pph);
        mbar_wait(a_free(as), aph);                                   // the MMAs that read this A slot have completed (long ago)
        const uint32_t src = patch + (uint32_t)(ps * DP_PATCH_BYTES) + patch_off;
        const uint32_t dst = a_ring + (uint32_t)(as * DP_A_BYTES) + (uint32_t)(dty * 8 * 128);
        if (live) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float acc[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              acc[i][0] = bv0.x; acc[i][1] = bv0.y; acc[i][2] = bv0.z; acc[i][3] = bv0.w;
              acc[i][4] = bv1.x; acc[i][5] = bv1.y; acc[i][6] = bv1.z; acc[i][7] = bv1.w;
            }
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              uint4 in[6];
#pragma unroll
              for (int i = 0; i < 6; ++i) in[i] = lds128(src + (uint32_t)((ky * DP_PW + h * 4 + i) * 128));
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const uint4 wq = wreg[ky * 3 + kx];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const uint4 xv = in[i + kx];
                  dp_fhfma2(acc[i][0], acc[i][1], xv.x, wq.x);
                  dp_fhfma2(acc[i][2], acc[i][3], xv.y, wq.y);
                  dp_fhfma2(acc[i][4], acc[i][5], xv.z, wq.z);
                  dp_fhfma2(acc[i][6], acc[i][7], xv.w, wq.w);
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 o;
              o.x = pack2(actT<ACT>(make_float2(acc[i][0], acc[i][1])));
              o.y = pack2(actT<ACT>(make_float2(acc[i][2], acc[i][3])));
              o.z = pack2(actT<ACT>(make_float2(acc[i][4], acc[i][5])));
              o.w = pack2(actT<ACT>(make_float2(acc[i][6], acc[i][7])));
              const int idx = h * 4 + i;                              // row r = dty * 8 + idx: r & 7 = idx
              sts128(dst + (uint32_t)(idx * 128) + (uint32_t)(((j ^ idx) & 7) << 4), o);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(p_empty(ps));                      // this warp has read the patch slot
        if (++ps == np) { ps = 0; pph ^= 1u; }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full(as));
        if (++as == na) { as = 0; aph ^= 1u; }
      }
    }
  } else {
    // ===== epilogue warps 4..7: one thread = one accumulator row (pixel) =====
    pdl_wait();                                   // (outputs / keys may alias buffers the previous kernel still reads)
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;            // accumulator row (TMEM lane) = pixel ety * 8 + etx of the tile
    const int ety = r >> 3, etx = r & 7;
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    const uint32_t a_row = (uint32_t)(r * 128), sw = (uint32_t)(r & 7);
    const int chunks1 = p.C1 >> 4, chunks2 = p.C2 >> 4;
    int i = 0;
    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++i) {
      const int b = i & 1;
      const uint32_t ph = (uint32_t)((i >> 1) & 1);
      const int r1 = fast_div(t, p.tiles_x, p.inv_tiles_x), tx = t - r1 * p.tiles_x;
      const int n = fast_div(r1, p.tiles_y, p.inv_tiles_y), ty = r1 - n * p.tiles_y;
      const int oy = ty * DP_TH + ety, ox = tx * DP_TW + etx;
      const bool valid = oy < p.H && ox < p.W;
      const long long pix = ((long long)n * p.H + oy) * p.W + ox;
      const uint32_t trow = tmem_base + (uint32_t)(b * p.acc_cols) + lane_sel;
      bf16* yrow = p.y + pix * p.y_ld;
      float rmax = -INFINITY;
      mbar_wait(acc1_full(b), ph);
      tc_fence_after();
      if (B2B) {
        // epilogue 1: act(acc + b1) -> bf16 -> A operand of GEMM 2 (the previous tile's GEMM 2 has completed: this warp waited for it)
        float dummy = 0.f;
        for (int c = 0; c < chunks1; ++c) {
          uint32_t v[16], w[8];
          tmem_ld16(trow + (uint32_t)(c * 16), v);
          tmem_ld_wait();
          epi_chunk<ACT>(v, b1s + 64u * (uint32_t)c, w, dummy);
          const uint32_t dst = a2s + (uint32_t)((c >> 2) * DP_A_BYTES) + a_row;
          const uint32_t jj = (uint32_t)(c & 3) * 2u;
          sts128(dst + ((jj ^ sw) << 4), make_uint4(w[0], w[1], w[2], w[3]));
          sts128(dst + (((jj + 1u) ^ sw) << 4), make_uint4(w[4], w[5], w[6], w[7]));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a2_full(b));
        mbar_wait(acc2_full(b), ph);
        tc_fence_after();
      }
      {
        const int chunks = B2B ? chunks2 : chunks1;
        const uint32_t bs = B2B ? b2s : b1s;
        for (int c = 0; c < chunks; ++c) {
          uint32_t v[16], w[8];
          tmem_ld16(trow + (uint32_t)(c * 16), v);
          tmem_ld_wait();
          if (B2B) epi_chunk<LPC_ACT_NONE>(v, bs + 64u * (uint32_t)c, w, rmax);
          else epi_chunk<ACT>(v, bs + 64u * (uint32_t)c, w, rmax);
          if (valid) {
            bf16* dstp = yrow + c * 16;
            if ((reinterpret_cast<uintptr_t>(dstp) & 31u) == 0) {
              asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dstp), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]),
                           "r"(w[5]), "r"(w[6]), "r"(w[7])
                           : "memory");
            } else {
              *reinterpret_cast<uint4*>(dstp) = make_uint4(w[0], w[1], w[2], w[3]);
              *reinterpret_cast<uint4*>(dstp + 8) = make_uint4(w[4], w[5], w[6], w[7]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free(b));
      if (p.rowmax && valid) {
        // max of the ROUNDED outputs = rounding of the max (round-to-nearest is monotonic); key as in tail.cu
        const uint32_t u = __float_as_uint(__bfloat162float(__float2bfloat16_rn(rmax)));
        p.rowmax[(long long)n * p.rowmax_img + p.rowmax_off + (long long)oy * p.W + ox] = u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

typedef CUresult (*DpEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
DpEncodeFn dp_get_encode() {
  static DpEncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<DpEncodeFn>(f);
  });
  return fn;
}

constexpr size_t DP_SMEM_LIMIT = 220 * 1024;

// shared-memory plan; returns total dynamic bytes (0 = does not fit)
size_t dp_plan(DwPwParams& p) {
  p.kb1 = (p.Cin + 63) / 64;
  p.kb2 = p.C2 ? (p.C1 + 63) / 64 : 0;
  p.cin_pad = p.kb1 * 64;
  const int cmax = p.C1 > p.C2 ? p.C1 : p.C2;
  p.acc_cols = 32;
  while (p.acc_cols < cmax) p.acc_cols <<= 1;
  p.tmem_cols = 2 * p.acc_cols;                    // two accumulators: the epilogue of tile i overlaps GEMM 1 of tile i + 1
  if (p.tmem_cols > 512) return 0;
  size_t off = 0;
  auto take = [&](size_t bytes, size_t align) { off = (off + align - 1) / align * align; const size_t o = off; off += bytes; return (unsigned)o; };
  p.off_w1 = take((size_t)p.kb1 * p.C1 * 128, 1024);
  p.off_w2 = take((size_t)p.kb2 * p.C2 * 128, 1024);
  p.off_a2 = take((size_t)p.kb2 * DP_A_BYTES, 1024);
  const size_t fixed_tail = (size_t)9 * p.cin_pad * 2 + (size_t)p.cin_pad * 4 + (size_t)(p.C1 + p.C2) * 4 + 6 * 128;
  // Two CTAs per SM (<= 110 KB each) when the rings fit, else one CTA with deeper rings.  Measured (dw64->80 @80x80 B64, us):
  // 2 CTAs x 2 patch slots 44, 1 CTA x 7 patch slots 65, 1 CTA x 4 slots 52 - the kernel is bound by warp instruction issue
  // (53 % issue-active at 20 warps per SM, 8.3 k warp instructions per tile), not by the patch latency; LPC_DWPW_NP caps the
  // patch ring for such sweeps.
  static const int np_cap = [] { const char* e = getenv("LPC_DWPW_NP"); return e ? atoi(e) : DP_MAX_NP; }();
  for (int pass = 0; pass < 2; ++pass) {
    const size_t budget = pass == 0 ? 110 * 1024 : DP_SMEM_LIMIT;
    int na = 2, np = 2;
    size_t need = off + 1024 + (size_t)na * DP_A_BYTES + (size_t)np * DP_PATCH_BYTES + fixed_tail + 1024;
    if (need > budget) continue;
    while (np < DP_MAX_NP && np < np_cap && np < p.kb1 + 2 && need + DP_PATCH_BYTES <= budget) { ++np; need += DP_PATCH_BYTES; }
    while (na < DP_MAX_NA && na < 2 * p.kb1 && need + DP_A_BYTES <= budget) { ++na; need += DP_A_BYTES; }
    p.na = na;
    p.np = np;
    p.off_a = take((size_t)na * DP_A_BYTES, 1024);
    p.off_patch = take((size_t)np * DP_PATCH_BYTES, 128);
    p.off_dww = take((size_t)9 * p.cin_pad * 2, 16);
    p.off_dwb = take((size_t)p.cin_pad * 4, 16);
    p.off_b1 = take((size_t)p.C1 * 4, 16);
    p.off_b2 = take((size_t)(p.C2 ? p.C2 : 4) * 4, 16);
    return off + 1024;     // + slack for the 1024-byte alignment of the base
  }
  return 0;
}

bool dp_shape_ok(int Cin, int C1, int C2) {
  if (Cin <= 0 || Cin % 16 || Cin > 1024) return false;     // whole K = 16 slices: every A chunk an MMA reads is written
  if (C1 <= 0 || C1 % 16 || C1 > 256) return false;
  if (C2 < 0 || C2 % 16 || C2 > 256) return false;
  return true;
}

}  // namespace

extern "C" int lpc_dwpw_tc_supported(int Cin, int C1, int C2, int x_ld, int y_ld, int dw_act, int act1, int act2) {
  if (!dp_shape_ok(Cin, C1, C2) || x_ld % 8 || y_ld % 8 || x_ld < Cin) return 0;
  // one activation for the depthwise conv and the first pointwise conv (SiLU in the head, Mish in CIB, or none); the second
  // pointwise stage is a plain conv
  if (dw_act != act1 || !(act1 == LPC_ACT_SILU || act1 == LPC_ACT_MISH || act1 == LPC_ACT_NONE)) return 0;
  if (C2 && act2 != LPC_ACT_NONE) return 0;
  DwPwParams p;
  memset(&p, 0, sizeof(p));
  p.Cin = Cin; p.C1 = C1; p.C2 = C2;
  return dp_plan(p) ? 1 : 0;
}

extern "C" int lpc_dwpw_tc(const void* x, int x_ld, int B, int H, int W, int Cin, const float* dw_w, const float* dw_bias, int dw_act,
                           const void* w1, const float* b1, int C1, int act1, const void* w2, const float* b2, int C2, int act2,
                           void* y, int y_ld, unsigned int* rowmax_keys, long long rowmax_img_stride, int rowmax_offset, void* stream) {
  LPC_REQUIRE(x && dw_w && w1 && y, "dwpw_tc: null pointer");
  LPC_REQUIRE(B > 0 && H > 0 && W > 0, "dwpw_tc: bad shape");
  LPC_REQUIRE((C2 == 0) == (w2 == nullptr), "dwpw_tc: w2 / C2 mismatch");
  if (!lpc_dwpw_tc_supported(Cin, C1, C2, x_ld, y_ld, dw_act, act1, act2))
    LPC_FAIL(LPC_E_UNSUPPORTED, "dwpw_tc: unsupported shape / activations Cin=%d C1=%d C2=%d x_ld=%d y_ld=%d act %d/%d/%d", Cin, C1, C2, x_ld, y_ld, dw_act, act1, act2);
  LPC_REQUIRE(aligned16(x) && aligned16(w1) && aligned16(w2) && aligned16(y), "dwpw_tc: pointers must be 16-byte aligned");
  LPC_REQUIRE(y_ld >= (C2 ? C2 : C1), "dwpw_tc: output pitch smaller than the channel count");
  DpEncodeFn enc = dp_get_encode();
  if (!enc) LPC_FAIL(LPC_E_CUDA, "dwpw_tc: cuTensorMapEncodeTiled not available");
  DwPwParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.C1 = C1; p.C2 = C2;
  const size_t smem = dp_plan(p);
  p.tiles_x = (W + DP_TW - 1) / DP_TW;
  p.tiles_y = (H + DP_TH - 1) / DP_TH;
  const long long nt = (long long)p.tiles_x * p.tiles_y * B;
  LPC_REQUIRE(nt < (1 << 24), "dwpw_tc: too many tiles");
  p.ntiles = (int)nt;
  p.inv_tiles_x = 1.0f / (float)p.tiles_x;
  p.inv_tiles_y = 1.0f / (float)p.tiles_y;
  p.dw_act = dw_act; p.act1 = act1; p.act2 = act2;
  p.dw_w = dw_w; p.dw_b = dw_bias; p.b1 = b1; p.b2 = b2;
  p.y = (bf16*)y; p.y_ld = y_ld;
  p.rowmax = rowmax_keys; p.rowmax_img = rowmax_img_stride; p.rowmax_off = rowmax_offset;

  CUtensorMap xmap, w1map, w2map;
  memset(&w2map, 0, sizeof(w2map));
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)x_ld * 2, (cuuint64_t)W * x_ld * 2, (cuuint64_t)H * W * x_ld * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)DP_PW, (cuuint32_t)DP_PH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&xmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) LPC_FAIL(LPC_E_CUDA, "dwpw_tc: activation tensor map encode failed (CUresult %d)", (int)r);
  }
  auto wmap = [&](CUtensorMap* m, const void* w, int rows, int k) -> int {
    const int kpad = (k + 63) / 64 * 64;
    cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kpad * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
  };
  if (int r = wmap(&w1map, w1, C1, Cin)) LPC_FAIL(LPC_E_CUDA, "dwpw_tc: weight tensor map encode failed (CUresult %d)", r);
  if (C2)
    if (int r = wmap(&w2map, w2, C2, C1)) LPC_FAIL(LPC_E_CUDA, "dwpw_tc: second weight tensor map encode failed (CUresult %d)", r);

  typedef void (*Kern)(CUtensorMap, CUtensorMap, CUtensorMap, DwPwParams);
  Kern kern = nullptr;
  if (C2) kern = act1 == LPC_ACT_SILU ? dwpw_tc_kernel<LPC_ACT_SILU, true> : act1 == LPC_ACT_MISH ? dwpw_tc_kernel<LPC_ACT_MISH, true> : dwpw_tc_kernel<LPC_ACT_NONE, true>;
  else kern = act1 == LPC_ACT_SILU ? dwpw_tc_kernel<LPC_ACT_SILU, false> : act1 == LPC_ACT_MISH ? dwpw_tc_kernel<LPC_ACT_MISH, false> : dwpw_tc_kernel<LPC_ACT_NONE, false>;
  static unsigned long long attr_done = 0;
  if (lpc_first_on_device(&attr_done)) {
    const int lim = (int)DP_SMEM_LIMIT + 2048;
    cudaError_t e = cudaSuccess;
#define DP_ATTR(K_) if (cudaFuncSetAttribute(K_, cudaFuncAttributeMaxDynamicSharedMemorySize, lim) != cudaSuccess) e = cudaErrorUnknown;
    DP_ATTR((dwpw_tc_kernel<LPC_ACT_SILU, true>)) DP_ATTR((dwpw_tc_kernel<LPC_ACT_MISH, true>)) DP_ATTR((dwpw_tc_kernel<LPC_ACT_NONE, true>))
    DP_ATTR((dwpw_tc_kernel<LPC_ACT_SILU, false>)) DP_ATTR((dwpw_tc_kernel<LPC_ACT_MISH, false>)) DP_ATTR((dwpw_tc_kernel<LPC_ACT_NONE, false>))
#undef DP_ATTR
    if (e != cudaSuccess) { attr_done = 0; LPC_FAIL(LPC_E_CUDA, "dwpw_tc: smem attribute"); }
  }
  const int per_sm = (smem <= 112 * 1024 && p.tmem_cols <= 256) ? 2 : 1;
  long long grid = (long long)lpc_num_sms() * per_sm;
  if (grid > nt) grid = nt;
  lpc_launch_pdl(kern, dim3((unsigned)grid), dim3(320), smem, (cudaStream_t)stream, xmap, w1map, w2map, p);
  LPC_CHECK_LAUNCH("dwpw_tc");
  return LPC_OK;
}
