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
//   warps 0-7 (workers): depthwise 3x3 on the patch (bf16 x bf16 -> fp32 FMAs on packed words, four adjacent pixels per
//            thread: 18 LDS.128 for 4 x 8 outputs), bias + activation, bf16, written as the K-major, 128B-swizzled A
//            tile [128 pixels x 64 channels] of the pointwise GEMM
//   warp 9   one elected thread issues tcgen05.mma M128 x N=C1 x K16 per 16 channels against the resident 1x1 weights,
//            accumulator in TMEM
//   workers  epilogue: tcgen05.ld -> + bias -> activation -> either bf16 NHWC stores, or (second pointwise stage) bf16
//            back into the A ring as the operand of GEMM 2 (same accumulator columns, re-used) -> epilogue 2 -> stores,
//            optionally with the per-pixel max-logit key of the fused v10 tail (lpc_conv2d_tc_rowmax's contract).
//
// One CTA runs its tiles' phases in order (two CTAs per SM overlap each other); all mbarrier waits are bounded.
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
constexpr int DP_MAX_NA = 4, DP_MAX_NP = 4;
constexpr int DP_WORKERS = 256;

struct DwPwParams {
  int B, H, W, Cin, C1, C2;
  int kb1, kb2;                      // 64-channel K blocks of GEMM 1 (Cin) and GEMM 2 (C1)
  int cin_pad;                       // kb1 * 64
  int tiles_x, tiles_y, ntiles;
  float inv_tiles_x, inv_tiles_y;
  int dw_act, act1, act2;
  int tmem_cols;
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
  unsigned off_w1, off_w2, off_a, off_patch, off_dww, off_dwb, off_b1, off_b2, off_rm;   // from the 1024-aligned base
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
__device__ __forceinline__ float2 act2_(float2 v, int act) {
  switch (act) {
    case LPC_ACT_NONE: return v;
    case LPC_ACT_SILU: return silu2_(v);
    case LPC_ACT_MISH: return mish2_(v);
    default: return make_float2(apply_act<false>(v.x, act), apply_act<false>(v.y, act));
  }
}
__device__ __forceinline__ uint32_t pack2(float2 v) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(320, 2)
dwpw_tc_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap w1map, const __grid_constant__ CUtensorMap w2map,
               const __grid_constant__ DwPwParams p) {
  extern __shared__ __align__(1024) unsigned char dp_smem[];
  __shared__ __align__(8) unsigned long long bars[2 * DP_MAX_NA + 2 * DP_MAX_NP + 2];
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(dp_smem) + 1023u) & ~1023u;
  const uint32_t w1s = base + p.off_w1, w2s = base + p.off_w2, a_ring = base + p.off_a, patch = base + p.off_patch;
  const uint32_t dww = base + p.off_dww, dwb = base + p.off_dwb, b1s = base + p.off_b1, b2s = base + p.off_b2, rms = base + p.off_rm;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_free = [&](int s) { return bar0 + 8u * (DP_MAX_NA + s); };
  auto p_full = [&](int s) { return bar0 + 8u * (2 * DP_MAX_NA + s); };
  auto p_empty = [&](int s) { return bar0 + 8u * (2 * DP_MAX_NA + DP_MAX_NP + s); };
  const uint32_t w_full = bar0 + 8u * (2 * DP_MAX_NA + 2 * DP_MAX_NP);
  const uint32_t acc_full = w_full + 8u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    prefetch_tmap(&xmap);
    prefetch_tmap(&w1map);
    if (p.C2) prefetch_tmap(&w2map);
    for (int s = 0; s < DP_MAX_NA; ++s) { mbar_init(a_full(s), 8); mbar_init(a_free(s), 1); }
    for (int s = 0; s < DP_MAX_NP; ++s) { mbar_init(p_full(s), 1); mbar_init(p_empty(s), 8); }
    mbar_init(w_full, 1);
    mbar_init(acc_full, 1);
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
  const uint32_t tmem_acc = tmem_slot;
  pdl_trigger();

  if (warp == 8) {
    // ===== TMA producer: resident pointwise weights once, then the (tile, channel block) patch stream =====
    if (elect_one_sync()) {
      mbar_expect_tx(w_full, (uint32_t)(p.kb1 * p.C1 * 128 + p.kb2 * p.C2 * 128));
      for (int kb = 0; kb < p.kb1; ++kb) tma_load_2d(w1s + (uint32_t)(kb * p.C1 * 128), &w1map, w_full, kb * 64, 0);
      for (int kb = 0; kb < p.kb2; ++kb) tma_load_2d(w2s + (uint32_t)(kb * p.C2 * 128), &w2map, w_full, kb * 64, 0);
      pdl_wait();                                 // activations of the previous kernel
      int pu = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        const int r1 = fast_div(t, p.tiles_x, p.inv_tiles_x), tx = t - r1 * p.tiles_x;
        const int n = fast_div(r1, p.tiles_y, p.inv_tiles_y), ty = r1 - n * p.tiles_y;
        for (int kb = 0; kb < p.kb1; ++kb, ++pu) {
          const int s = pu % p.np;
          mbar_wait(p_empty(s), (uint32_t)(((pu / p.np) & 1) ^ 1));
          mbar_expect_tx(p_full(s), (uint32_t)DP_PATCH_BYTES);
          tma_load_4d(patch + (uint32_t)(s * DP_PATCH_BYTES), &xmap, p_full(s), kb * 64, tx * DP_TW - 1, ty * DP_TH - 1, n);
        }
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer =====
    if (elect_one_sync()) {
      const uint32_t idesc1 = make_idesc(p.C1), idesc2 = make_idesc(p.C2 ? p.C2 : 16);
      const uint32_t hi = desc_hi(1024u, 2u);                  // K-major, 128B swizzle, 8-row groups 1024 B apart (A and W alike)
      mbar_wait(w_full, 0);
      int au = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        for (int kb = 0; kb < p.kb1; ++kb, ++au) {
          const int s = au % p.na;
          mbar_wait(a_full(s), (uint32_t)((au / p.na) & 1));
          tc_fence_after();
          const uint32_t a_lo = desc_lo(a_ring + (uint32_t)(s * DP_A_BYTES), 16u), b_lo = desc_lo(w1s + (uint32_t)(kb * p.C1 * 128), 16u);
          const int nk = min(4, (p.Cin - kb * 64 + 15) >> 4);
          for (int k = 0; k < nk; ++k) umma_bf16(tmem_acc, desc64(a_lo + 2u * k, hi), desc64(b_lo + 2u * k, hi), idesc1, (uint32_t)((kb | k) != 0));
          umma_commit(a_free(s));
        }
        umma_commit(acc_full);
        if (p.C2) {
          for (int kb = 0; kb < p.kb2; ++kb, ++au) {
            const int s = au % p.na;
            mbar_wait(a_full(s), (uint32_t)((au / p.na) & 1));
            tc_fence_after();
            const uint32_t a_lo = desc_lo(a_ring + (uint32_t)(s * DP_A_BYTES), 16u), b_lo = desc_lo(w2s + (uint32_t)(kb * p.C2 * 128), 16u);
            const int nk = min(4, (p.C1 - kb * 64 + 15) >> 4);
            for (int k = 0; k < nk; ++k) umma_bf16(tmem_acc, desc64(a_lo + 2u * k, hi), desc64(b_lo + 2u * k, hi), idesc2, (uint32_t)((kb | k) != 0));
            umma_commit(a_free(s));
          }
          umma_commit(acc_full);
        }
      }
    }
  } else {
    // ===== workers: depthwise conv -> A tile, then the epilogue(s) =====
    pdl_wait();                                   // (rowmax keys / outputs may alias buffers the previous kernel still reads)
    const int j = tid & 7;                        // 16-byte chunk (8 channels) of the 64-channel block
    const int q = tid >> 3;                       // 0..31: tile row ty = q >> 1, four pixels from x = (q & 1) * 4
    const int dty = q >> 1, dx0 = (q & 1) * 4;
    const uint32_t patch_off = (uint32_t)((dty * DP_PW + dx0) * 128 + j * 16);
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane;            // accumulator row (TMEM lane) = pixel ety * 8 + etx of the tile
    const int ety = r >> 3, etx = r & 7;
    const uint32_t trow = tmem_acc + ((uint32_t)(quarter * 32) << 16);
    const uint32_t a_row = (uint32_t)(r * 128), sw = (uint32_t)(r & 7);
    int pu = 0, au = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
      const int r1 = fast_div(t, p.tiles_x, p.inv_tiles_x), tx = t - r1 * p.tiles_x;
      const int n = fast_div(r1, p.tiles_y, p.inv_tiles_y), ty = r1 - n * p.tiles_y;
      // ---- depthwise 3x3 per 64-channel block -> A ring ----
      for (int kb = 0; kb < p.kb1; ++kb, ++pu, ++au) {
        const int ps = pu % p.np, as = au % p.na;
        uint4 wreg[9];
        const uint32_t wsrc = dww + (uint32_t)(kb * 64 + j * 8) * 2u;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) wreg[tap] = lds128(wsrc + (uint32_t)(tap * p.cin_pad) * 2u);
        float bv[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) bv[v] = lds_f32(dwb + 4u * (uint32_t)(kb * 64 + j * 8 + v));
        mbar_wait(p_full(ps), (uint32_t)((pu / p.np) & 1));
        const uint32_t src = patch + (uint32_t)(ps * DP_PATCH_BYTES) + patch_off;
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int v = 0; v < 8; ++v) acc[i][v] = bv[v];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          uint4 in[6];
#pragma unroll
          for (int i = 0; i < 6; ++i) in[i] = lds128(src + (uint32_t)((ky * DP_PW + i) * 128));
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
        __syncwarp();
        if (lane == 0) mbar_arrive(p_empty(ps));                      // this warp has read the patch slot
        mbar_wait(a_free(as), (uint32_t)(((au / p.na) & 1) ^ 1));       // the MMAs that read this A slot have completed
        const uint32_t dst = a_ring + (uint32_t)(as * DP_A_BYTES);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o;
          o.x = pack2(act2_(make_float2(acc[i][0], acc[i][1]), p.dw_act));
          o.y = pack2(act2_(make_float2(acc[i][2], acc[i][3]), p.dw_act));
          o.z = pack2(act2_(make_float2(acc[i][4], acc[i][5]), p.dw_act));
          o.w = pack2(act2_(make_float2(acc[i][6], acc[i][7]), p.dw_act));
          const int rr = dty * 8 + dx0 + i;
          sts128(dst + (uint32_t)(rr * 128) + (uint32_t)(((j ^ (rr & 7)) & 7) << 4), o);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full(as));
      }
      // ---- epilogue 1 ----
      const int oy = ty * DP_TH + ety, ox = tx * DP_TW + etx;
      const bool valid = oy < p.H && ox < p.W;
      const long long pix = ((long long)n * p.H + oy) * p.W + ox;
      mbar_wait(acc_full, acc_phase);
      acc_phase ^= 1u;
      tc_fence_after();
      if (p.C2) {
        const int chunks = p.C1 >> 4, c_lo = half ? (chunks + 1) / 2 : 0, c_hi = half ? chunks : (chunks + 1) / 2;
        const int au2 = au;                                           // first A slot use of GEMM 2
        // slots of GEMM 2 were read by GEMM 1's MMAs, which have all completed (acc_full): keep the a_free phases in step
        for (int kb = 0; kb < p.kb2; ++kb) mbar_wait(a_free((au2 + kb) % p.na), (uint32_t)((((au2 + kb) / p.na) & 1) ^ 1));
        for (int c = c_lo; c < c_hi; ++c) {
          uint32_t v[16];
          tmem_ld16(trow + (uint32_t)(c * 16), v);
          tmem_ld_wait();
          uint4 o0, o1;
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 bb = make_float2(lds_f32(b1s + 4u * (uint32_t)(c * 16 + 2 * i)), lds_f32(b1s + 4u * (uint32_t)(c * 16 + 2 * i + 1)));
            w[i] = pack2(act2_(make_float2(__uint_as_float(v[2 * i]) + bb.x, __uint_as_float(v[2 * i + 1]) + bb.y), p.act1));
          }
          o0 = make_uint4(w[0], w[1], w[2], w[3]);
          o1 = make_uint4(w[4], w[5], w[6], w[7]);
          const int kb = c >> 2, jj = (c & 3) * 2;
          const uint32_t dst = a_ring + (uint32_t)(((au2 + kb) % p.na) * DP_A_BYTES) + a_row;
          sts128(dst + ((((uint32_t)jj) ^ sw) << 4), o0);
          sts128(dst + ((((uint32_t)jj + 1u) ^ sw) << 4), o1);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0)
          for (int kb = 0; kb < p.kb2; ++kb) mbar_arrive(a_full((au2 + kb) % p.na));
        au += p.kb2;
        mbar_wait(acc_full, acc_phase);
        acc_phase ^= 1u;
        tc_fence_after();
      }
      // ---- final epilogue: + bias, activation, (row max), bf16 NHWC stores ----
      {
        const int Cn = p.C2 ? p.C2 : p.C1;
        const uint32_t bs = p.C2 ? b2s : b1s;
        const int act = p.C2 ? p.act2 : p.act1;
        const int chunks = Cn >> 4, c_lo = half ? (chunks + 1) / 2 : 0, c_hi = half ? chunks : (chunks + 1) / 2;
        bf16* yrow = p.y + pix * p.y_ld;
        float rmax = -INFINITY;
        for (int c = c_lo; c < c_hi; ++c) {
          uint32_t v[16];
          tmem_ld16(trow + (uint32_t)(c * 16), v);
          tmem_ld_wait();
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 bb = make_float2(lds_f32(bs + 4u * (uint32_t)(c * 16 + 2 * i)), lds_f32(bs + 4u * (uint32_t)(c * 16 + 2 * i + 1)));
            const float2 o = act2_(make_float2(__uint_as_float(v[2 * i]) + bb.x, __uint_as_float(v[2 * i + 1]) + bb.y), act);
            rmax = fmaxf(rmax, fmaxf(o.x, o.y));
            w[i] = pack2(o);
          }
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
        tc_fence_before();
        if (p.rowmax) {
          // the two warps of a lane quarter own different column ranges: combine their maxima through shared memory
          if (half) asm volatile("st.shared.f32 [%0], %1;" ::"r"(rms + 4u * (uint32_t)r), "f"(rmax) : "memory");
          worker_bar();
          if (!half && valid) {
            const float m = fmaxf(rmax, lds_f32(rms + 4u * (uint32_t)r));
            // max of the ROUNDED outputs = rounding of the max (round-to-nearest is monotonic); key as in tail.cu
            const uint32_t u = __float_as_uint(__bfloat162float(__float2bfloat16_rn(m)));
            p.rowmax[(long long)n * p.rowmax_img + p.rowmax_off + (long long)oy * p.W + ox] = u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
          }
          worker_bar();                            // rms is free for the next tile
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_acc, (uint32_t)p.tmem_cols);
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
  size_t off = 0;
  auto take = [&](size_t bytes, size_t align) { off = (off + align - 1) / align * align; const size_t o = off; off += bytes; return (unsigned)o; };
  p.off_w1 = take((size_t)p.kb1 * p.C1 * 128, 1024);
  p.off_w2 = take((size_t)p.kb2 * p.C2 * 128, 1024);
  const size_t fixed_tail = (size_t)9 * p.cin_pad * 2 + (size_t)p.cin_pad * 4 + (size_t)(p.C1 + p.C2) * 4 + 512 + 4 * 128;
  // as many ring slots as fit: two CTAs per SM (<= 110 KB each) when possible
  for (int pass = 0; pass < 2; ++pass) {
    const size_t budget = pass == 0 ? 110 * 1024 : DP_SMEM_LIMIT;
    int na = p.kb2 > 2 ? p.kb2 : 2, np = 2;
    if (na > DP_MAX_NA) return 0;
    size_t need = off + 1024 + (size_t)na * DP_A_BYTES + (size_t)np * DP_PATCH_BYTES + fixed_tail + 1024;
    if (need > budget) continue;
    while (np < DP_MAX_NP && np < p.kb1 + 1 && need + DP_PATCH_BYTES <= budget) { ++np; need += DP_PATCH_BYTES; }
    while (na < DP_MAX_NA && na < p.kb1 && need + DP_A_BYTES <= budget) { ++na; need += DP_A_BYTES; }
    p.na = na;
    p.np = np;
    p.off_a = take((size_t)na * DP_A_BYTES, 1024);
    p.off_patch = take((size_t)np * DP_PATCH_BYTES, 128);
    p.off_dww = take((size_t)9 * p.cin_pad * 2, 16);
    p.off_dwb = take((size_t)p.cin_pad * 4, 16);
    p.off_b1 = take((size_t)p.C1 * 4, 16);
    p.off_b2 = take((size_t)(p.C2 ? p.C2 : 4) * 4, 16);
    p.off_rm = take(512, 16);
    return off + 1024;     // + slack for the 1024-byte alignment of the base
  }
  return 0;
}

bool dp_shape_ok(int Cin, int C1, int C2) {
  if (Cin <= 0 || Cin % 8 || Cin > 1024) return false;
  if (C1 <= 0 || C1 % 16 || C1 > 256) return false;
  if (C2 < 0 || C2 % 16 || C2 > 256) return false;
  return true;
}

}  // namespace

extern "C" int lpc_dwpw_tc_supported(int Cin, int C1, int C2, int x_ld, int y_ld) {
  if (!dp_shape_ok(Cin, C1, C2) || x_ld % 8 || y_ld % 8 || x_ld < Cin) return 0;
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
  if (!lpc_dwpw_tc_supported(Cin, C1, C2, x_ld, y_ld))
    LPC_FAIL(LPC_E_UNSUPPORTED, "dwpw_tc: unsupported shape Cin=%d C1=%d C2=%d x_ld=%d y_ld=%d", Cin, C1, C2, x_ld, y_ld);
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
  const int cmax = C1 > C2 ? C1 : C2;
  p.tmem_cols = 32;
  while (p.tmem_cols < cmax) p.tmem_cols <<= 1;

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

  static unsigned long long attr_done = 0;
  if (lpc_first_on_device(&attr_done))
    if (cudaFuncSetAttribute(dwpw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DP_SMEM_LIMIT + 2048) != cudaSuccess) {
      attr_done = 0;
      LPC_FAIL(LPC_E_CUDA, "dwpw_tc: smem attribute");
    }
  const int per_sm = (smem <= 112 * 1024 && p.tmem_cols <= 256) ? 2 : 1;
  long long grid = (long long)lpc_num_sms() * per_sm;
  if (grid > nt) grid = nt;
  lpc_launch_pdl(dwpw_tc_kernel, dim3((unsigned)grid), dim3(320), smem, (cudaStream_t)stream, xmap, w1map, w2map, p);
  LPC_CHECK_LAUNCH("dwpw_tc");
  return LPC_OK;
}
