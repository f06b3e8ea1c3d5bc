// stem_tc.cu - the Cin = 3 stem convolution (YAML layer 0: Conv(3, c, 3, 2), conv.py:36-54) on the tensor cores.
//
// K = 27 is too ragged for TMA-fed implicit GEMM, so the im2col row is BUILT in shared memory by CUDA threads:
// the input is NHWC with C padded to 4 (one 8-byte pixel), an output pixel's row is its 9 taps x 4 channels = 36
// K-elements, followed by two 1.0 entries that multiply the bias (bf16 hi / lo split) stored in the matching K rows of the
// weight tile, zero-padded to K = 48 = three tcgen05.mma of K = 16.  M tile = 8 x 16 output pixels (128 rows),
// N = Cout (16 ... 128), accumulator double-buffered in TMEM.
//
// Warps 0-3: builders (thread = A row: its 3 x 3-pixel window arrives through a thread-private cp.async ring, five
// tiles ahead; 5 swizzled 16-byte shared stores per row); warp 4: MMA issuer; warps 5-8: epilogue (tcgen05.ld ->
// SiLU / Mish -> bf16 -> 16-byte stores).  Persistent; every role is latency-bound per tile (~1.3k cycles, measured with
// LPC_STEM_DBG), so throughput comes from 4 independent CTAs per SM (2 when Cout > 32 limits TMEM).
//
// Per image at 640x640: 1.64 MB in + 3.28 MB out (Cout = 16): HBM-bound (the CUDA-core version was FMA-bound at
// 5x the HBM floor, profiles/r01_f_*).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

#define STRACE(role, tile, k) do { if (p.trace && blockIdx.x == 0 && (tile) < 64) p.trace[((role) * 64 + (tile)) * 4 + (k)] = clock64(); } while (0)

constexpr int SB_TW = 16, SB_TH = 8;          // output tile
constexpr int SB_THREADS = 288;
constexpr int SB_A_BYTES = 128 * 128;         // 128 rows x 128 B (K = 64 slots, 48 used), 128B swizzle
constexpr int SB_DEPTH = 3, SB_SLOT = 80;     // input ring: tiles in flight per CTA, bytes per thread slot
// TMA input box of one 8 x 16 output tile: input rows 2*oy0-1 .. 2*oy0+15, pixels 2*ox0-2 .. 2*ox0+33 (36 pixels of 8 bytes;
// 33 are read; the box starts on an EVEN pixel because every box row must start on a 16-byte boundary in global memory -
// a box starting at pixel 2*ox0-1 raised 'illegal instruction'), zero fill outside the image = the conv padding.  The image is described as [B][H][W*4 bf16] so that the
// box's inner extent (144 elements = 288 bytes) is a multiple of 16 bytes.
constexpr int SB_BOX_PX = 36, SB_BOX_ROWS = 2 * SB_TH + 1, SB_BOX_ROW_B = SB_BOX_PX * 8;
constexpr int SB_BOX_BYTES = SB_BOX_ROWS * SB_BOX_ROW_B, SB_BOX_SLOT = (SB_BOX_BYTES + 127) & ~127;
static_assert(SB_DEPTH * SB_BOX_SLOT <= SB_DEPTH * 128 * SB_SLOT, "the TMA ring reuses the cp.async ring's shared memory");
// uint8 HWC input (lpc_stem_conv_u8): the image is [B][H][W*3 bytes]; the box of a tile starts 16 bytes before byte
// 3 * 2*ox0 (a multiple of 96, so every box row starts on a 16-byte boundary) and is 112 bytes wide: pixel 2*ox0-1 is at
// byte 13, pixel 2*ox0+32 ends at byte 111.  Zero fill outside the image = the conv padding (0 / 255 = 0).
constexpr int SB_U8_ROW_B = 112, SB_U8_LEAD = 13, SB_U8_BOX_BYTES = SB_BOX_ROWS * SB_U8_ROW_B;
static_assert(SB_U8_BOX_BYTES <= SB_BOX_SLOT, "the uint8 box fits a ring slot");


struct StemTcParams {
  const uint2* x;      // [B, H, W] pixels of 4 bf16
  const float* w;      // [27][Cout] fp32 (tap-major, then input channel)
  const float* bias;   // [Cout] or null
  bf16* y;
  long long y_ld;
  int B, H, W, Ho, Wo, Cout, act;
  int tiles_x, tiles_y, m_tiles;
  int acc_cols, tmem_cols;
  float inv_per_img, inv_tiles_x;
  int tma_in;          // the 3x3 stride-2 input windows of a tile arrive as ONE TMA box (else: per-thread cp.async ring)
  int swap_rb;         // U8 kernels: the image bytes are BGR: byte 2 of a pixel goes to channel slot 0
  unsigned long long* trace;   // LPC_STEM_DBG=1: [3 roles][64 tiles][4 stamps] clock64 of CTA 0
};

template <int ACT>
__device__ __forceinline__ void stem_epilogue_row(uint32_t trow, int Cout, bf16* yrow, bool valid) {
  for (int c = 0; c < Cout; c += 16) {
    uint32_t v[16];
    tmem_ld16(trow + (uint32_t)c, v);
    tmem_ld_wait();
    if (valid) {
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float t = __uint_as_float(v[i]);
        f[i] = ACT == LPC_ACT_SILU ? silu_<false>(t) : ACT == LPC_ACT_MISH ? mish_<false>(t) : ACT == LPC_ACT_NONE ? t : apply_act<false>(t, ACT);
      }
      Vec<bf16> o, o2;
      o.pack(f);
      o2.pack(f + 8);
      if ((reinterpret_cast<uintptr_t>(yrow + c) & 31u) == 0) {      // one 32-byte store per lane (see conv_tc.cu store16)
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(yrow + c), "r"(o.raw.x), "r"(o.raw.y),
                     "r"(o.raw.z), "r"(o.raw.w), "r"(o2.raw.x), "r"(o2.raw.y), "r"(o2.raw.z), "r"(o2.raw.w)
                     : "memory");
      } else {
        st_vec<bf16>(yrow + c, o);
        st_vec<bf16>(yrow + c + 8, o2);
      }
    }
  }
}

// U8: the input is the uint8 HWC image itself (always by TMA): the builders normalise (v / 255, bit-identical to lpc_pack_u8's
// bf16 output: fma(2^23 + v, 1/255f, -2^23/255f) == v * (1/255f) exactly, and bf16_rn of that equals bf16_rn(v / 255f) for
// all 256 values - checked exhaustively) and pad to 4 channels while they build the im2col row.
template <int ACT, bool U8>
__global__ void __launch_bounds__(SB_THREADS, 4)
stem_tc_kernel(const __grid_constant__ StemTcParams p, const __grid_constant__ CUtensorMap in_map) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bars[8 + SB_DEPTH];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t a_base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // two A stages
  const uint32_t b_base = a_base + 2 * SB_A_BYTES;                      // weight tile [Cout][128 B]
  unsigned char* gen_base = smem_raw + (a_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto afull = [&](int s) { return bar0 + 8u * s; };
  auto aempty = [&](int s) { return bar0 + 8u * (2 + s); };
  auto tfull = [&](int s) { return bar0 + 8u * (4 + s); };
  auto tempty = [&](int s) { return bar0 + 8u * (6 + s); };
  auto ifull = [&](int s) { return bar0 + 8u * (8 + s); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(afull(s), 128);
      mbar_init(aempty(s), 1);
      mbar_init(tfull(s), 1);
      mbar_init(tempty(s), 4);
    }
    for (int s = 0; s < SB_DEPTH; ++s) mbar_init(ifull(s), 1);
    if (U8 || p.tma_in) prefetch_tmap(&in_map);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(smem_u32(&tmem_base_slot), (uint32_t)p.tmem_cols);
  // zero both A stages (the K padding is never rewritten) and build the weight tile: B[co][k], k = tap*4 + ci,
  // k = 36 / 37 hold the bias hi / lo; element (co, k) lives at co*128 + ((k/8) ^ (co%8))*16 + (k%8)*2.
  for (int i = threadIdx.x; i < 2 * SB_A_BYTES / 16; i += SB_THREADS) reinterpret_cast<uint4*>(gen_base)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < p.Cout * 64; i += SB_THREADS) {
    const int co = i >> 6, k = i & 63;
    float v = 0.f;
    if (k < 36) {
      const int tap = k >> 2, ci = k & 3;
      if (ci < 3) v = p.w[(tap * 3 + ci) * p.Cout + co];
    } else if (k < 38 && p.bias) {
      const float b = p.bias[co];
      const float hi = __bfloat162float(__float2bfloat16_rn(b));
      v = (k == 36) ? hi : b - hi;
    }
    const int off = co * 128 + (((k >> 3) ^ (co & 7)) << 4) + (k & 7) * 2;
    *reinterpret_cast<bf16*>(gen_base + 2 * SB_A_BYTES + off) = __float2bfloat16_rn(v);
  }
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int per_img = p.tiles_x * p.tiles_y;
  pdl_trigger();
  pdl_wait();          // the prologue above read only weights / bias

  if (warp < 4) {
    // ===== builders =====
    const int r = threadIdx.x;                      // A row = output pixel (ty, tx) of the tile
    const int ty = r >> 4, tx = r & 15;
    const uint32_t row_addr = (uint32_t)(r * 128);
    const uint32_t sw = (uint32_t)(r & 7);
    if (U8 || p.tma_in) {
      // One elected builder thread fetches the whole input window of tile it + SB_DEPTH - 1 with a single TMA box; every
      // builder then reads its own 3 x (16 + 8) bytes from shared memory.  The per-thread cp.async ring below needed
      // six LSU instructions per output pixel on the global side alone and made the L1 data pipe the stem's limiter.
      const uint32_t ring0 = b_base + (uint32_t)(p.Cout * 128);
      auto issue_box = [&](int m, int slot) {
        if (m < p.m_tiles) {
          const int img = fast_div(m, per_img, p.inv_per_img), rem = m - img * per_img;
          const int tyi = fast_div(rem, p.tiles_x, p.inv_tiles_x), txi = rem - tyi * p.tiles_x;
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_expect_tx(ifull(slot), (uint32_t)(U8 ? SB_U8_BOX_BYTES : SB_BOX_BYTES));
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
              ::"r"(ring0 + (uint32_t)(slot * SB_BOX_SLOT)), "l"(&in_map), "r"(ifull(slot)),
                "r"(U8 ? 2 * txi * SB_TW * 3 - 16 : (2 * txi * SB_TW - 2) * 4), "r"(2 * tyi * SB_TH - 1), "r"(img)
              : "memory");
        }
      };
      int m = blockIdx.x;
      if (r == 0)
        for (int d = 0; d < SB_DEPTH - 1; ++d) issue_box(m + d * (int)gridDim.x, d);
      const uint32_t my_off = U8 ? (uint32_t)(2 * ty * SB_U8_ROW_B + SB_U8_LEAD + tx * 6) : (uint32_t)(2 * ty * SB_BOX_ROW_B + tx * 16);
      int it = 0;
      for (; m < p.m_tiles; m += gridDim.x, ++it) {
        if (r == 0) STRACE(0, it, 0);
        // the slot refilled now was read in iteration it - 1: all 128 builders have passed their loads of it
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (r == 0) issue_box(m + (SB_DEPTH - 1) * (int)gridDim.x, (it + SB_DEPTH - 1) % SB_DEPTH);
        const int slot = it % SB_DEPTH;
        mbar_wait(ifull(slot), (uint32_t)((it / SB_DEPTH) & 1));
        const uint32_t src = ring0 + (uint32_t)(slot * SB_BOX_SLOT) + my_off;
        uint4 e12[3];     // pixels 2ox, 2ox+1 of input row k (box pixels 2tx+2, 2tx+3)
        uint2 e0[3];      // pixel 2ox-1 (box pixel 2tx+1)
        if (U8) {
          // 9 bytes (3 pixels x BGR / RGB) per input row, at an arbitrary byte offset: three aligned words, two funnel shifts
          const uint32_t al = src & ~3u, sh = (src & 3u) * 8u;
          const float inv = 1.0f / 255.0f, off = -8388608.0f * (1.0f / 255.0f);
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            uint32_t w0, w1, w2;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w0) : "r"(al + (uint32_t)(k * SB_U8_ROW_B)));
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w1) : "r"(al + (uint32_t)(k * SB_U8_ROW_B + 4)));
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w2) : "r"(al + (uint32_t)(k * SB_U8_ROW_B + 8)));
            const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh), b8 = w2 >> sh;
            float f[9];
#pragma unroll
            for (int j = 0; j < 9; ++j) {
              const uint32_t word = j < 4 ? lo : (j < 8 ? hi : b8);
              // bytes -> float without I2F: 0x4B0000vv is 2^23 + v; the FMA removes the 2^23 and scales in one rounding
              f[j] = fmaf(__uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540u | (uint32_t)(j & 3))), inv, off);
            }
            if (p.swap_rb) {      // BGR bytes -> RGB channel slots (the K order, hence the accumulation order, stays the packed path's)
#pragma unroll
              for (int q = 0; q < 9; q += 3) { const float t = f[q]; f[q] = f[q + 2]; f[q + 2] = t; }
            }
            auto pk = [](float a, float b) {
              const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
              return *reinterpret_cast<const uint32_t*>(&h);
            };
            e0[k] = make_uint2(pk(f[0], f[1]), pk(f[2], 0.f));
            e12[k] = make_uint4(pk(f[3], f[4]), pk(f[5], 0.f), pk(f[6], f[7]), pk(f[8], 0.f));
          }
        } else {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(e0[k].x), "=r"(e0[k].y) : "r"(src + (uint32_t)(k * SB_BOX_ROW_B + 8)));
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(e12[k].x), "=r"(e12[k].y), "=r"(e12[k].z), "=r"(e12[k].w) : "r"(src + (uint32_t)(k * SB_BOX_ROW_B + 16)));
          }
        }
        const int s = it & 1;
        if (r == 0) STRACE(0, it, 1);
        mbar_wait(aempty(s), (uint32_t)(((it >> 1) & 1) ^ 1));
        if (r == 0) STRACE(0, it, 2);
        unsigned char* rowp = gen_base + s * SB_A_BYTES + row_addr;
        // K chunks of 8 elements = two taps, exactly as in the cp.async path below
        *reinterpret_cast<uint4*>(rowp + ((0u ^ sw) << 4)) = make_uint4(e0[0].x, e0[0].y, e12[0].x, e12[0].y);
        *reinterpret_cast<uint4*>(rowp + ((1u ^ sw) << 4)) = make_uint4(e12[0].z, e12[0].w, e0[1].x, e0[1].y);
        *reinterpret_cast<uint4*>(rowp + ((2u ^ sw) << 4)) = e12[1];
        *reinterpret_cast<uint4*>(rowp + ((3u ^ sw) << 4)) = make_uint4(e0[2].x, e0[2].y, e12[2].x, e12[2].y);
        *reinterpret_cast<uint4*>(rowp + ((4u ^ sw) << 4)) = make_uint4(e12[2].z, e12[2].w, 0x3F803F80u, 0u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(afull(s));
        if (r == 0) STRACE(0, it, 3);
      }
    } else {
    // Input ring: every thread stages the 72 bytes of its own 3 x 3-pixel window (3 x [16 B: pixels 2ox, 2ox+1] then
    // 3 x [8 B: pixel 2ox-1]) for tile it + SB_DEPTH - 1 with cp.async (zero-fill outside the image) into a
    // thread-private 80-byte slot, so SB_DEPTH - 1 tiles of loads are in flight per CTA without holding registers.
    unsigned char* ring = gen_base + 2 * SB_A_BYTES + p.Cout * 128 + r * SB_SLOT;
    const uint32_t ring_u32 = b_base + (uint32_t)(p.Cout * 128 + r * SB_SLOT);
    auto issue_tile = [&](int m, int slot) {
      if (m < p.m_tiles) {
        const int img = fast_div(m, per_img, p.inv_per_img), rem = m - img * per_img;
        const int tyi = fast_div(rem, p.tiles_x, p.inv_tiles_x), txi = rem - tyi * p.tiles_x;
        const int oy = tyi * SB_TH + ty, ox = txi * SB_TW + tx;
        const uint2* base = p.x + (long long)img * p.H * p.W;
        const uint32_t dst = ring_u32 + (uint32_t)(slot * 128 * SB_SLOT);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int iy = 2 * oy - 1 + ky;
          const bool rok = (unsigned)iy < (unsigned)p.H && ox < p.Wo;
          const uint2* rowp = rok ? base + (long long)iy * p.W + 2 * ox : p.x;
          const uint32_t n12 = rok ? (2 * ox + 1 < p.W ? 16u : 8u) : 0u;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst + 16u * ky), "l"(rowp), "r"(n12) : "memory");
          const bool lok = rok && ox > 0;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst + 48u + 8u * ky), "l"(lok ? rowp - 1 : p.x), "r"(lok ? 8u : 0u) : "memory");
        }
      }
      cp_async_commit();
    };
    int m = blockIdx.x;
#pragma unroll
    for (int d = 0; d < SB_DEPTH - 1; ++d) issue_tile(m + d * (int)gridDim.x, d);
    int it = 0;
    for (; m < p.m_tiles; m += gridDim.x, ++it) {
      if (r == 0) STRACE(0, it, 0);
      issue_tile(m + (SB_DEPTH - 1) * (int)gridDim.x, (it + SB_DEPTH - 1) % SB_DEPTH);
      asm volatile("cp.async.wait_group %0;" ::"n"(SB_DEPTH - 1) : "memory");
      const unsigned char* slot = ring + (it % SB_DEPTH) * 128 * SB_SLOT;
      uint4 e12[3];
      uint2 e0[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        e12[k] = *reinterpret_cast<const uint4*>(slot + 16 * k);
        e0[k] = *reinterpret_cast<const uint2*>(slot + 48 + 8 * k);
      }
      const int s = it & 1;
      if (r == 0) STRACE(0, it, 1);
      mbar_wait(aempty(s), (uint32_t)(((it >> 1) & 1) ^ 1));
      if (r == 0) STRACE(0, it, 2);
      unsigned char* rowp = gen_base + s * SB_A_BYTES + row_addr;
      // K chunks of 8 elements = two taps: (t0,t1) (t2,t3) (t4,t5) (t6,t7) (t8, 1, 1, 0...)
      const uint4 c0 = make_uint4(e0[0].x, e0[0].y, e12[0].x, e12[0].y);
      const uint4 c1 = make_uint4(e12[0].z, e12[0].w, e0[1].x, e0[1].y);
      const uint4 c2 = e12[1];
      const uint4 c3 = make_uint4(e0[2].x, e0[2].y, e12[2].x, e12[2].y);
      const uint4 c4 = make_uint4(e12[2].z, e12[2].w, 0x3F803F80u, 0u);
      *reinterpret_cast<uint4*>(rowp + ((0u ^ sw) << 4)) = c0;
      *reinterpret_cast<uint4*>(rowp + ((1u ^ sw) << 4)) = c1;
      *reinterpret_cast<uint4*>(rowp + ((2u ^ sw) << 4)) = c2;
      *reinterpret_cast<uint4*>(rowp + ((3u ^ sw) << 4)) = c3;
      *reinterpret_cast<uint4*>(rowp + ((4u ^ sw) << 4)) = c4;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(afull(s));
      if (r == 0) STRACE(0, it, 3);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
  } else if (warp == 4) {
    // ===== MMA issuer =====
    if (elect_one_sync()) {
      const uint32_t idesc = make_idesc(p.Cout);
      const uint32_t hi = desc_hi(1024u, 2u);
      const uint32_t b_lo = desc_lo(b_base, 16u);
      int it = 0;
      for (int m = blockIdx.x; m < p.m_tiles; m += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t ph = (uint32_t)((it >> 1) & 1);
        STRACE(1, it, 0);
        mbar_wait(tempty(s), ph ^ 1u);
        STRACE(1, it, 1);
        mbar_wait(afull(s), ph);
        STRACE(1, it, 2);
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(s * p.acc_cols);
        const uint32_t a_lo = desc_lo(a_base + (uint32_t)(s * SB_A_BYTES), 16u);
        umma_bf16(acc, desc64(a_lo, hi), desc64(b_lo, hi), idesc, 0u);
        umma_acc(acc, desc64(a_lo + 2u, hi), desc64(b_lo + 2u, hi), idesc);
        umma_acc(acc, desc64(a_lo + 4u, hi), desc64(b_lo + 4u, hi), idesc);
        umma_commit(aempty(s));
        umma_commit(tfull(s));
        STRACE(1, it, 3);
      }
    }
  } else {
    // ===== epilogue =====
    const int q = warp & 3;                         // TMEM lane quarter
    const int r = q * 32 + lane;
    const int ty = r >> 4, tx = r & 15;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    int it = 0;
    for (int m = blockIdx.x; m < p.m_tiles; m += gridDim.x, ++it) {
      const int img = fast_div(m, per_img, p.inv_per_img), rem = m - img * per_img;
      const int tyi = fast_div(rem, p.tiles_x, p.inv_tiles_x), txi = rem - tyi * p.tiles_x;
      const int oy = tyi * SB_TH + ty, ox = txi * SB_TW + tx;
      const bool valid = oy < p.Ho && ox < p.Wo;
      bf16* yrow = p.y + (((long long)img * p.Ho + oy) * p.Wo + ox) * p.y_ld;
      const int s = it & 1;
      if (threadIdx.x == 160) STRACE(2, it, 0);
      mbar_wait(tfull(s), (uint32_t)((it >> 1) & 1));
      if (threadIdx.x == 160) STRACE(2, it, 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + (uint32_t)(s * p.acc_cols) + lane_off;
      stem_epilogue_row<ACT>(trow, p.Cout, yrow, valid);
      if (threadIdx.x == 160) STRACE(2, it, 2);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(s));
      if (threadIdx.x == 160) STRACE(2, it, 3);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace

namespace {

typedef CUresult (*StemEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
StemEncodeFn stem_encode() {
  static StemEncodeFn enc = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return reinterpret_cast<StemEncodeFn>(f);
  }();
  return enc;
}

template <bool U8>
void (*stem_kernel_of(int act))(StemTcParams, CUtensorMap) {
  return act == LPC_ACT_SILU ? stem_tc_kernel<LPC_ACT_SILU, U8> : act == LPC_ACT_MISH ? stem_tc_kernel<LPC_ACT_MISH, U8>
       : act == LPC_ACT_NONE ? stem_tc_kernel<LPC_ACT_NONE, U8> : stem_tc_kernel<LPC_ACT_RELU, U8>;
}

// Common part of the two entry points: tile geometry, TMEM / CTA plan, tensor map, launch.
int stem_launch(StemTcParams& p, bool u8, const void* x, int B, int H, int W, int Cout, int act, cudaStream_t stream) {
  if (Cout % 16 || Cout > 128 || (H & 1) || (W & 1)) return LPC_E_UNSUPPORTED;
  if (!(act == LPC_ACT_SILU || act == LPC_ACT_MISH || act == LPC_ACT_NONE || act == LPC_ACT_RELU)) return LPC_E_UNSUPPORTED;
  p.B = B; p.H = H; p.W = W; p.Ho = H / 2; p.Wo = W / 2; p.Cout = Cout; p.act = act;
  p.tiles_x = (p.Wo + SB_TW - 1) / SB_TW;
  p.tiles_y = (p.Ho + SB_TH - 1) / SB_TH;
  const long long tiles = (long long)p.tiles_x * p.tiles_y * B;
  if (tiles >= (1ll << 24)) return LPC_E_UNSUPPORTED;   // fast_div range
  p.m_tiles = (int)tiles;
  p.inv_per_img = 1.0f / (float)(p.tiles_x * p.tiles_y);
  p.inv_tiles_x = 1.0f / (float)p.tiles_x;
  p.acc_cols = 32;
  while (p.acc_cols < Cout) p.acc_cols <<= 1;
  p.tmem_cols = 2 * p.acc_cols;
  const size_t smem = 2 * SB_A_BYTES + (size_t)Cout * 128 + (size_t)SB_DEPTH * 128 * SB_SLOT + 1024;
  static unsigned long long attr = 0;     // per device
  if (lpc_first_on_device(&attr)) {
    const int lim = 2 * SB_A_BYTES + 128 * 128 + SB_DEPTH * 128 * SB_SLOT + 1024;
    bool ok = true;
    for (int a : {LPC_ACT_SILU, LPC_ACT_MISH, LPC_ACT_NONE, LPC_ACT_RELU}) {
      ok = ok && cudaFuncSetAttribute(stem_kernel_of<false>(a), cudaFuncAttributeMaxDynamicSharedMemorySize, lim) == cudaSuccess;
      ok = ok && cudaFuncSetAttribute(stem_kernel_of<true>(a), cudaFuncAttributeMaxDynamicSharedMemorySize, lim) == cudaSuccess;
    }
    if (!ok) { attr = 0; LPC_FAIL(LPC_E_CUDA, "stem_conv: smem attribute"); }
  }
  const int sms = lpc_num_sms();
  int per_sm = 512 / p.tmem_cols;
  if (per_sm > 4) per_sm = 4;
  while (per_sm > 1 && (size_t)per_sm * (smem + 2048) > 227 * 1024) --per_sm;
  { static const int cta_cap = [] { const char* e = getenv("LPC_CTA_CAP"); return e ? atoi(e) : 4; }(); if (per_sm > cta_cap) per_sm = cta_cap < 1 ? 1 : cta_cap; }
  long long grid = (long long)sms * per_sm;
  if (grid > tiles) grid = tiles;
  static unsigned long long* trace_buf = nullptr;
  static const bool dbg = getenv("LPC_STEM_DBG") != nullptr;
  p.trace = nullptr;
  if (dbg) {
    if (!trace_buf) cudaMalloc(&trace_buf, 3 * 64 * 4 * 8);
    cudaMemset(trace_buf, 0, 3 * 64 * 4 * 8);
    p.trace = trace_buf;
  }
  CUtensorMap in_map;
  memset(&in_map, 0, sizeof(in_map));
  p.tma_in = 0;
  StemEncodeFn enc = stem_encode();
  if (u8) {
    // input tensor map: [B][H][W*3 bytes], box = 112 bytes x 17 rows (see SB_U8_*)
    if (!enc) LPC_FAIL(LPC_E_CUDA, "stem_conv_u8: cuTensorMapEncodeTiled not available");
    cuuint64_t dims[3] = {(cuuint64_t)W * 3, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)W * 3, (cuuint64_t)H * W * 3};
    cuuint32_t box[3] = {(cuuint32_t)SB_U8_ROW_B, (cuuint32_t)SB_BOX_ROWS, 1};
    cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = enc(&in_map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) LPC_FAIL(LPC_E_CUDA, "stem_conv_u8: tensor map encode failed (CUresult %d)", (int)r);
    p.tma_in = 1;
  } else {
    // input tensor map: [B][H][W*4 bf16], box = 144 elements x 17 rows (see SB_BOX_*)
    static const int tma_env = [] { const char* e = getenv("LPC_STEM_TMA"); return e ? atoi(e) : 1; }();
    if (tma_env && enc && W % 2 == 0 && W * 4 >= SB_BOX_PX * 4 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
      cuuint64_t dims[3] = {(cuuint64_t)W * 4, (cuuint64_t)H, (cuuint64_t)B};
      cuuint64_t strides[2] = {(cuuint64_t)W * 8, (cuuint64_t)H * W * 8};
      cuuint32_t box[3] = {(cuuint32_t)SB_BOX_PX * 4, (cuuint32_t)SB_BOX_ROWS, 1};
      cuuint32_t es[3] = {1, 1, 1};
      if (enc(&in_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(x), dims, strides, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
        p.tma_in = 1;
    }
  }
  lpc_launch_pdl(u8 ? stem_kernel_of<true>(act) : stem_kernel_of<false>(act), (unsigned)grid, SB_THREADS, smem, stream, p, in_map);
  LPC_CHECK_LAUNCH("stem_conv_tc");
  if (dbg) {
    static int calls = 0;
    cudaDeviceSynchronize();
    if (calls++ == 3) {
      unsigned long long h[3 * 64 * 4];
      cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
      const unsigned long long t0 = h[0];
      for (int t = 0; t < 40; ++t) {
        printf("tile %2d  BLD:", t);
        for (int k = 0; k < 4; ++k) printf(" %6lld", (long long)(h[(0 * 64 + t) * 4 + k] - t0));
        printf(" | MMA:");
        for (int k = 0; k < 4; ++k) printf(" %6lld", (long long)(h[(1 * 64 + t) * 4 + k] - t0));
        printf(" | EPI:");
        for (int k = 0; k < 4; ++k) printf(" %6lld", (long long)(h[(2 * 64 + t) * 4 + k] - t0));
        printf("\n");
      }
      fflush(stdout);
    }
  }
  return LPC_OK;
}

}  // namespace

// Returns LPC_E_UNSUPPORTED (without setting an error) when the shape is not for this kernel; lpc_stem_conv then
// uses the CUDA-core kernel (fp32 validation mode, stride 1, odd sizes).
int lpc_stem_conv_tc(const void* x, int B, int H, int W, const float* w, const float* bias, int Cout, void* y, int y_ld,
                     int act, cudaStream_t stream) {
  StemTcParams p;
  memset(&p, 0, sizeof(p));
  p.x = (const uint2*)x; p.w = w; p.bias = bias; p.y = (bf16*)y; p.y_ld = y_ld;
  return stem_launch(p, false, x, B, H, W, Cout, act, stream);
}

// ---- the stem straight from the uint8 image ----------------------------------------------------------------------
extern "C" int lpc_stem_conv_u8_supported(int H, int W, int stride, int Cout, int y_ld, int act) {
  static const int on = [] { const char* e = getenv("LPC_STEM_U8"); return e ? atoi(e) : 1; }();
  if (!on || stride != 2 || H <= 0 || W <= 0 || (H & 1) || W % 16) return 0;       // W * 3 bytes per row: a multiple of 16
  if (Cout % 16 || Cout > 128 || y_ld % 8 || y_ld < Cout) return 0;
  if (!(act == LPC_ACT_SILU || act == LPC_ACT_MISH || act == LPC_ACT_NONE || act == LPC_ACT_RELU)) return 0;
  return 1;
}

extern "C" int lpc_stem_conv_u8(const void* x, int B, int H, int W, int swap_rb, const float* w, const float* bias, int stride, int Cout,
                                void* y, int y_ld, int act, void* stream) {
  LPC_REQUIRE(x && w && y, "stem_conv_u8: null pointer");
  LPC_REQUIRE(B > 0, "stem_conv_u8: bad shape");
  if (!lpc_stem_conv_u8_supported(H, W, stride, Cout, y_ld, act))
    LPC_FAIL(LPC_E_UNSUPPORTED, "stem_conv_u8: unsupported shape H=%d W=%d stride=%d Cout=%d y_ld=%d act=%d", H, W, stride, Cout, y_ld, act);
  LPC_REQUIRE(aligned16(x) && aligned16(y) && aligned16(w), "stem_conv_u8: pointers must be 16-byte aligned");
  StemTcParams p;
  memset(&p, 0, sizeof(p));
  p.x = nullptr; p.w = w; p.bias = bias; p.y = (bf16*)y; p.y_ld = y_ld;
  p.swap_rb = swap_rb ? 1 : 0;
  const int r = stem_launch(p, true, x, B, H, W, Cout, act, (cudaStream_t)stream);
  if (r == LPC_E_UNSUPPORTED) LPC_FAIL(LPC_E_UNSUPPORTED, "stem_conv_u8: shape not taken (H=%d W=%d Cout=%d)", H, W, Cout);
  return r;
}
