// conv_tc.cu - dense convolution as an implicit GEMM on the 5th-gen tensor cores (sm_100a).
//
//   D[M = output pixels, N = Cout] = A[M, K = taps*Cin] * W[N, K]^T,  bf16 operands, fp32 accumulate.
//
// * A is never materialised (no im2col buffer): for every filter tap the TMA engine loads a box
//   [kc channels, TW, TH, 1 image] of the NHWC activation straight into 128B/64B/32B-swizzled shared
//   memory; zero padding is the TMA out-of-bounds fill.  Stride-2 convs (and the space_to_depth + 1x1
//   fold, k=2) read through four "parity" tensor maps whose W/H strides are doubled, so tap (ky,kx)
//   is again a plain shifted box.  1x1 convs see the activation as a flat [M, Cin] matrix.
// * W ([Cout][Kpad] bf16, K-major) is loaded by a 2-D TMA map, 64 K-elements x n_tile rows per stage.
// * tcgen05.mma (cta_group::1, M=128, N=n_tile<=256, K=16) is issued by one elected thread; the
//   accumulator lives in TMEM; tcgen05.commit releases smem stages and signals the epilogue.
// * Epilogue warps read TMEM with tcgen05.ld (32 lanes x 16 columns), apply bias + activation
//   (+ SPCA channel gate, + residual), convert to bf16 and store 32-byte runs into the NHWC output
//   view (which may be a channel slice of a concat buffer).
// * Warp roles: w0 = TMA producer, w1 = TMEM allocator + MMA issuer, w2..w5 = epilogue.  Up to two
//   CTAs are resident per SM, so one tile's epilogue overlaps another tile's main loop.
#include <cuda.h>

#include <mutex>
#include <unordered_map>
#include <string>
#include <cstring>

#include "common.cuh"

namespace {

constexpr int TC_NT = 192;
constexpr int MAX_STAGES = 8;
constexpr int A_STAGE_BYTES = 128 * 64 * 2;  // 128 rows x 64 K-elements of bf16
constexpr int MAX_TAPS = 9;

struct TmapPack {
  CUtensorMap a[4];
  CUtensorMap b;
};

struct ConvTcParams {
  int Ho, Wo, B;
  int tiles_x, tiles_y, TW, TH;
  int Cout, n_tile, tmem_cols;
  int ksteps, kc, nsub, chunks_per_tap, real_slots, stages;
  int pix_per_img;
  signed char tap_map[MAX_TAPS], tap_dx[MAX_TAPS], tap_dy[MAX_TAPS];
  const float* bias;
  const float* chan_scale;
  const bf16* res;
  long long res_ld;
  bf16* y;
  long long y_ld;
  int act;
};

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      printf("lpc conv_tc: mbarrier timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor (sm_100 format: version 1 at bit 46).
//   layout: 2 = 128B swizzle, 4 = 64B, 6 = 32B; sbo = bytes between 8-row groups; lbo unused (=1).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// ---- the kernel ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_NT)
conv_tc_kernel(const __grid_constant__ TmapPack maps, const __grid_constant__ ConvTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bars[2 * MAX_STAGES + 1];
  __shared__ uint32_t tmem_base_slot;

  // dynamic smem may not be 1024-aligned by default: align manually (host adds 1 KB of slack)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b_stage_bytes = p.n_tile * 128;
  const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (MAX_STAGES + s); };
  const uint32_t accum_bar = bar0 + 8u * (2 * MAX_STAGES);

  // tile coordinates
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int img = blockIdx.x / tiles_per_img;
  const int trem = blockIdx.x - img * tiles_per_img;
  const int tyi = trem / p.tiles_x, txi = trem - tyi * p.tiles_x;
  const int x0 = txi * p.TW, y0 = tyi * p.TH;
  const int n0 = blockIdx.y * p.n_tile;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a[0]);
    prefetch_tmap(&maps.b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_slot), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const int rows = p.TW * p.TH;
      const uint32_t tx_bytes = (uint32_t)(rows * 128 + b_stage_bytes);
      const int sub_bytes = 256 * p.kc;
      for (int ks = 0; ks < p.ksteps; ++ks) {
        const int s = ks % p.stages;
        const uint32_t ph = (uint32_t)((ks / p.stages) & 1);
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t a_dst = smem_base + (uint32_t)(s * stage_bytes);
        const uint32_t b_dst = a_dst + A_STAGE_BYTES;
        mbar_expect_tx(full_bar(s), tx_bytes);
        for (int j = 0; j < p.nsub; ++j) {
          int q = ks * p.nsub + j;
          if (q >= p.real_slots) q = p.real_slots - 1;  // padded K: weights are zero there, any finite A will do
          const int tap = q / p.chunks_per_tap;
          const int c0 = (q - tap * p.chunks_per_tap) * p.kc;
          tma_load_4d(a_dst + (uint32_t)(j * sub_bytes), &maps.a[p.tap_map[tap]], full_bar(s), c0,
                      x0 + p.tap_dx[tap], y0 + p.tap_dy[tap], img);
        }
        tma_load_2d(b_dst, &maps.b, full_bar(s), ks * 64, n0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_tile >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t a_layout = p.kc == 64 ? 2u : (p.kc == 32 ? 4u : 6u);
      const uint32_t a_sbo = (uint32_t)(16 * p.kc);  // 8 rows * (kc*2) bytes
      const int sub_bytes = 256 * p.kc;
      for (int ks = 0; ks < p.ksteps; ++ks) {
        const int s = ks % p.stages;
        const uint32_t ph = (uint32_t)((ks / p.stages) & 1);
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t a_base = smem_base + (uint32_t)(s * stage_bytes);
        const uint32_t b_base = a_base + A_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int e = k * 16;
          const int j = e / p.kc;
          const uint32_t a_addr = a_base + (uint32_t)(j * sub_bytes + (e - j * p.kc) * 2);
          const uint64_t adesc = smem_desc(a_addr, a_sbo, a_layout);
          const uint64_t bdesc = smem_desc(b_base + (uint32_t)(k * 32), 1024u, 2u);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (ks > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(s));  // frees this smem stage once the MMAs above have read it
      }
      umma_commit(accum_bar);  // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
    const int q = warp & 3;
    const int r = q * 32 + lane;  // accumulator row == pixel slot inside the tile
    const int ty = r / p.TW, tx = r - ty * p.TW;
    const int ox = x0 + tx, oy = y0 + ty;
    const bool valid = (r < p.TW * p.TH) && ox < p.Wo && oy < p.Ho;
    const long long pix = ((long long)img * p.Ho + oy) * p.Wo + ox;
    bf16* yrow = p.y + pix * p.y_ld + n0;
    const bf16* rrow = p.res ? p.res + pix * p.res_ld + n0 : nullptr;
    const float* srow = p.chan_scale ? p.chan_scale + (pix / p.pix_per_img) * p.Cout + n0 : nullptr;
    const float* brow = p.bias ? p.bias + n0 : nullptr;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int c = 0; c < p.n_tile; c += 16) {
      uint32_t v[16];
      tmem_ld16(trow + (uint32_t)c, v);
      tmem_ld_wait();
      if (valid) {
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float t = __uint_as_float(v[i]) + (brow ? __ldg(brow + c + i) : 0.f);
          f[i] = apply_act<false>(t, p.act);
        }
        if (srow) {
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] *= __ldg(srow + c + i);
        }
        if (rrow) {
          float r8[8];
          ld_vec<bf16>(rrow + c).unpack(r8);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] += r8[i];
          ld_vec<bf16>(rrow + c + 8).unpack(r8);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[8 + i] += r8[i];
        }
        Vec<bf16> o;
        o.pack(f);
        st_vec<bf16>(yrow + c, o);
        o.pack(f + 8);
        st_vec<bf16>(yrow + c + 8, o);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

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

int pick_kc(int Cin) { return Cin % 64 == 0 ? 64 : (Cin % 32 == 0 ? 32 : 16); }

int pick_ntile(int Cout) {
  for (int nt = (Cout + 255) / 256; nt <= Cout / 16; ++nt)
    if (Cout % nt == 0 && (Cout / nt) % 16 == 0 && Cout / nt <= 256) return Cout / nt;
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
    if (eff > best + 1e-9 || (eff > best - 1e-9 && tw > bw)) { best = eff; bw = tw; bh = th; }
  }
  *TW = bw;
  *TH = bh;
}

int encode_act_map(CUtensorMap* m, const bf16* base, long long ld, int C, long long Wd, long long Hd, long long Bd,
                   long long sW, long long sH, long long sB, int kc, int TW, int TH) {
  EncodeTiledFn enc = get_encode();
  if (!enc) LPC_FAIL(LPC_E_CUDA, "conv2d_tc: cuTensorMapEncodeTiled not available");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wd, (cuuint64_t)Hd, (cuuint64_t)Bd};
  cuuint64_t strides[3] = {(cuuint64_t)sW * 2, (cuuint64_t)sH * 2, (cuuint64_t)sB * 2};
  cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)TW, (cuuint32_t)TH, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  (void)ld;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    LPC_FAIL(LPC_E_CUDA, "conv2d_tc: activation tensor map encode failed (CUresult %d; C=%d W=%lld H=%lld B=%lld kc=%d box %dx%d)",
             (int)r, C, Wd, Hd, Bd, kc, TW, TH);
  return LPC_OK;
}

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
  if (pick_ntile(Cout) == 0) return 0;
  return 1;
}

extern "C" int lpc_conv2d_tc(const void* x, int x_ld, int B, int H, int W, int Cin, const void* w, const float* bias,
                             int k, int stride, int pad, int Cout, void* y, int y_ld, int act,
                             const float* chan_scale, const void* res, int res_ld, void* stream) {
  LPC_REQUIRE(x && w && y, "conv2d_tc: null pointer");
  LPC_REQUIRE(B > 0 && H > 0 && W > 0, "conv2d_tc: bad shape");
  if (!lpc_conv2d_tc_supported(Cin, Cout, k, stride, pad, x_ld, y_ld))
    LPC_FAIL(LPC_E_UNSUPPORTED, "conv2d_tc: unsupported shape Cin=%d Cout=%d k=%d s=%d p=%d x_ld=%d y_ld=%d", Cin, Cout, k, stride, pad, x_ld, y_ld);
  LPC_REQUIRE(aligned16(x) && aligned16(w) && aligned16(y) && aligned16(res), "conv2d_tc: pointers must be 16-byte aligned");
  LPC_REQUIRE(!res || res_ld % 8 == 0, "conv2d_tc: res_ld must be a multiple of 8");
  LPC_REQUIRE(stride == 1 || (H % 2 == 0 && W % 2 == 0), "conv2d_tc: stride-2 needs even H, W");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const bf16* xb = (const bf16*)x;

  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  TmapPack maps;
  memset(&maps, 0, sizeof(maps));
  p.kc = pick_kc(Cin);
  p.nsub = 64 / p.kc;
  p.chunks_per_tap = Cin / p.kc;
  const int ntaps = k * k;
  p.real_slots = ntaps * p.chunks_per_tap;
  const int kpad = lpc_conv2d_tc_kpad(Cin, k);
  p.ksteps = kpad / 64;
  p.Cout = Cout;
  p.n_tile = pick_ntile(Cout);
  p.tmem_cols = 32;
  while (p.tmem_cols < p.n_tile) p.tmem_cols <<= 1;
  p.pix_per_img = Ho * Wo;
  p.bias = bias;
  p.chan_scale = chan_scale;
  p.res = (const bf16*)res;
  p.res_ld = res_ld;
  p.y = (bf16*)y;
  p.y_ld = y_ld;
  p.act = act;

  if (k == 1) {
    // flat GEMM view: one "image" of M x 1 pixels
    const long long M = (long long)B * H * W;
    LPC_REQUIRE(M < (1ll << 31), "conv2d_tc: too many pixels");
    p.B = 1; p.Ho = 1; p.Wo = (int)M; p.TW = 128; p.TH = 1;
    p.tiles_x = (int)((M + 127) / 128); p.tiles_y = 1;
    p.tap_map[0] = 0; p.tap_dx[0] = 0; p.tap_dy[0] = 0;
    if (int e = encode_act_map(&maps.a[0], xb, x_ld, Cin, M, 1, 1, x_ld, (long long)M * x_ld, (long long)M * x_ld, p.kc, 128, 1)) return e;
  } else {
    p.B = B; p.Ho = Ho; p.Wo = Wo;
    pick_tile(Ho, Wo, &p.TW, &p.TH);
    p.tiles_x = (Wo + p.TW - 1) / p.TW;
    p.tiles_y = (Ho + p.TH - 1) / p.TH;
    if (stride == 1) {
      for (int t = 0; t < ntaps; ++t) { p.tap_map[t] = 0; p.tap_dx[t] = (signed char)(t % k - pad); p.tap_dy[t] = (signed char)(t / k - pad); }
      if (int e = encode_act_map(&maps.a[0], xb, x_ld, Cin, W, H, B, x_ld, (long long)W * x_ld, (long long)H * W * x_ld, p.kc, p.TW, p.TH)) return e;
    } else {
      // parity maps: map (py,px) views pixels (2*hy+py, 2*hx+px)
      for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px)
          if (int e = encode_act_map(&maps.a[py * 2 + px], xb + ((long long)py * W + px) * x_ld, x_ld, Cin, (W - px + 1) / 2,
                                     (H - py + 1) / 2, B, 2ll * x_ld, 2ll * W * x_ld, (long long)H * W * x_ld, p.kc, p.TW, p.TH))
            return e;
      for (int t = 0; t < ntaps; ++t) {
        const int oy = t / k - pad, ox = t % k - pad;          // input offset relative to 2*o
        const int py = ((oy % 2) + 2) % 2, px = ((ox % 2) + 2) % 2;
        p.tap_map[t] = (signed char)(py * 2 + px);
        p.tap_dy[t] = (signed char)((oy - py) / 2);
        p.tap_dx[t] = (signed char)((ox - px) / 2);
      }
    }
  }
  {
    EncodeTiledFn enc = get_encode();
    if (!enc) LPC_FAIL(LPC_E_CUDA, "conv2d_tc: cuTensorMapEncodeTiled not available");
    cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)kpad * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)p.n_tile};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&maps.b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) LPC_FAIL(LPC_E_CUDA, "conv2d_tc: weight tensor map encode failed (CUresult %d)", (int)r);
  }
  const int stage_bytes = A_STAGE_BYTES + p.n_tile * 128;
  int stages = (100 * 1024) / stage_bytes;
  if (stages > 4) stages = 4;
  if (stages > p.ksteps) stages = p.ksteps;
  if (stages < 2) stages = p.ksteps < 2 ? 1 : 2;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) LPC_FAIL(LPC_E_CUDA, "conv2d_tc: smem attribute: %s", cudaGetErrorString(e));
    smem_set = 200 * 1024;
  }
  dim3 grid((unsigned)(p.tiles_x * p.tiles_y * p.B), (unsigned)(Cout / p.n_tile));
  conv_tc_kernel<<<grid, TC_NT, smem, (cudaStream_t)stream>>>(maps, p);
  LPC_CHECK_LAUNCH("conv2d_tc");
  return LPC_OK;
}
