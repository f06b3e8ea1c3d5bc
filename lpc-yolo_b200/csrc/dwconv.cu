// dwconv.cu - depthwise convolution and SPPF pooling, NHWC, bandwidth-bound CUDA-core kernels.
//
// Channel vectors (16 bytes: 8 bf16 / 4 fp32) are the fastest-varying index across threads -> coalesced 128-bit accesses.
#include "common.cuh"
#include "tc_ptx.cuh"

#include <stdlib.h>
#include <string.h>

#include <mutex>

namespace {

__device__ __forceinline__ float relu_(float v) { return fmaxf(v, 0.0f); }

// Row-sliding register window: a thread owns one 16-byte channel vector of PX horizontally adjacent outputs.  Per
// filter row it loads the NV = (PX-1)*S + (K-1)*D + 1 input vectors that row needs ONCE, unpacks them to fp32
// registers once, and every (kx, pixel) pair is then 8 (bf16) / 4 (fp32) FMAs on registers; the K weight vectors of
// the row are 16-byte loads.  3x3 s1: 18 loads + 72 FMA per output vector instead of 36 loads + 72 scalar weight loads.
template <typename T, int K, int S, int D, int PX>
__global__ void __launch_bounds__(128)
dwconv_kernel(const T* __restrict__ x, int x_ld, int B, int H, int W, int C,
              const float* __restrict__ w, const float* __restrict__ bias, int pad,
              int Ho, int Wo, T* __restrict__ y, int y_ld, int act,
              const T* __restrict__ res, int res_ld) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = Vec<T>::N;
  constexpr bool PR = Precise<T>::value;
  constexpr int NV = (PX - 1) * S + (K - 1) * D + 1;
  // thread index = (image, 8-row block, x group, row in block, channel vector): a CTA covers an 8-row strip, so the
  // vertical halo (K-1 extra input rows per 8 output rows) is re-read from L1, not from L2 by another SM
  // grid = (x-group/channel chunks, 8-row blocks, images); thread in block -> (x group, row in block, channel vector)
  const int cvecs = C / V;
  const int wgroups = (Wo + PX - 1) / PX;
  const unsigned local = blockIdx.x * blockDim.x + threadIdx.x;
  if (local >= (unsigned)(wgroups * 8 * cvecs)) return;
  const unsigned q = local / (unsigned)cvecs;
  const int cv = (int)(local - q * (unsigned)cvecs);
  const int ry = (int)(q & 7u);
  const int xg = (int)(q >> 3);
  const int oy = blockIdx.y * 8 + ry;
  const int n = blockIdx.z;
  if (oy >= Ho) return;
  const int c0 = cv * V;
  const int ox0 = xg * PX;
  const int ix0 = ox0 * S - pad;

  float2 acc2[PX][V / 2];
  {
    float bv[V];
#pragma unroll
    for (int v = 0; v < V; v += 4) {
      const float4 b4 = bias ? __ldg(reinterpret_cast<const float4*>(bias + c0 + v)) : make_float4(0.f, 0.f, 0.f, 0.f);
      bv[v] = b4.x; bv[v + 1] = b4.y; bv[v + 2] = b4.z; bv[v + 3] = b4.w;
    }
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
      for (int v = 0; v < V; v += 2) acc2[p][v / 2] = make_float2(bv[v], bv[v + 1]);
  }
  const T* img = x + (long long)n * H * W * x_ld + c0;
  // RB filter rows per batch: all their loads are issued before any is consumed (memory-level parallelism per thread;
  // the kernel is latency-bound otherwise: ~28 resident warps per SM)
  constexpr int RB = 1;
#pragma unroll
  for (int ky0 = 0; ky0 < K; ky0 += RB) {
    uint4 raw[RB][NV];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int ky = ky0 + r;
      const int iy = oy * S - pad + ky * D;
      const bool rok = ky < K && iy >= 0 && iy < H;
      const T* row = img + (long long)iy * W * x_ld;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int ix = ix0 + j;
        raw[r][j] = make_uint4(0, 0, 0, 0);
        if (rok && ix >= 0 && ix < W) raw[r][j] = __ldg(reinterpret_cast<const uint4*>(row + (long long)ix * x_ld));
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int ky = ky0 + r;
      if (ky >= K) break;
      float2 in[NV][V / 2];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        Vec<T> t;
        t.raw = raw[r][j];
        float f[V];
        t.unpack(f);
#pragma unroll
        for (int v = 0; v < V; v += 2) in[j][v / 2] = make_float2(f[v], f[v + 1]);
      }
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        float2 wv[V / 2];
#pragma unroll
        for (int v = 0; v < V; v += 4) {
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + (ky * K + kx) * C + c0 + v));
          wv[v / 2] = make_float2(w4.x, w4.y);
          wv[v / 2 + 1] = make_float2(w4.z, w4.w);
        }
        // packed fp32 FMA (FFMA2, sm_100): two channels per instruction - the kernel is issue-bound
#pragma unroll
        for (int p = 0; p < PX; ++p)
#pragma unroll
          for (int v = 0; v < V / 2; ++v) acc2[p][v] = __ffma2_rn(in[p * S + kx * D][v], wv[v], acc2[p][v]);
      }
    }
  }
  float o[PX][V];
#pragma unroll
  for (int p = 0; p < PX; ++p)
#pragma unroll
    for (int v = 0; v < V; v += 2) { o[p][v] = acc2[p][v / 2].x; o[p][v + 1] = acc2[p][v / 2].y; }
  // one uniform branch on the activation for the whole tile (a per-element switch costs a branch per element)
#define DW_ACT(fn)                                   \
  _Pragma("unroll") for (int p = 0; p < PX; ++p)     \
  _Pragma("unroll") for (int v = 0; v < V; ++v) o[p][v] = fn(o[p][v]);
  switch (act) {
    case LPC_ACT_NONE: break;
    case LPC_ACT_SILU: DW_ACT(silu_<PR>) break;
    case LPC_ACT_MISH: DW_ACT(mish_<PR>) break;
    case LPC_ACT_SIGMOID: DW_ACT(sigmoid_<PR>) break;
    default: DW_ACT(relu_) break;
  }
#undef DW_ACT
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    const int ox = ox0 + p;
    if (ox >= Wo) break;
    const long long opix = (long long)(n * Ho + oy) * Wo + ox;
    if (res) {
      float r[V];
      ld_vec<T>(res + opix * res_ld + c0).unpack(r);
#pragma unroll
      for (int v = 0; v < V; ++v) o[p][v] += r[v];
    }
    Vec<T> ov;
    ov.pack(o[p]);
    st_vec<T>(y + opix * y_ld + c0, ov);
  }
}

// SPPF: three chained MaxPool2d(5,1,2) (-inf padding).  One CTA owns one image x VP 16-byte channel vectors per pixel
// (2 x 8 bf16 channels), keeps the whole map in shared memory IN ITS NATIVE TYPE - max needs no arithmetic, so bf16
// pairs are compared packed (HMNMX2.BF16) and nothing is converted - and runs each 5x5 pool as a separable row pass +
// column pass; the three pool outputs go to channel slices [0,C), [C,2C), [2C,3C) of y with 16-byte stores.  The first
// version staged fp32 float4 pairs per pixel: twice the shared-memory traffic for the same channels and 2-way bank
// conflicts in the column pass (33 us for 6.5 MB, profiles/r01_k_per_launch_lpc_b64.csv).
template <typename T> __device__ __forceinline__ uint4 vec_max(uint4 a, uint4 b);
template <> __device__ __forceinline__ uint4 vec_max<bf16>(uint4 a, uint4 b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}
template <> __device__ __forceinline__ uint4 vec_max<float>(uint4 a, uint4 b) {
  uint4 r;
  r.x = __float_as_uint(fmaxf(__uint_as_float(a.x), __uint_as_float(b.x)));
  r.y = __float_as_uint(fmaxf(__uint_as_float(a.y), __uint_as_float(b.y)));
  r.z = __float_as_uint(fmaxf(__uint_as_float(a.z), __uint_as_float(b.z)));
  r.w = __float_as_uint(fmaxf(__uint_as_float(a.w), __uint_as_float(b.w)));
  return r;
}

template <typename T>
__global__ void __launch_bounds__(256)
sppf_pool_kernel(const T* __restrict__ x, int x_ld, int H, int W, int C, T* __restrict__ y, int y_ld, int VP) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = Vec<T>::N;
  extern __shared__ uint4 sp4[];
  const int HW = H * W, items = HW * VP;
  uint4* a = sp4;                 // [HW][VP]
  uint4* b = sp4 + items;
  const int n = blockIdx.y, c0 = blockIdx.x * VP * V;
  const T* xin = x + (long long)n * HW * x_ld + c0;
  T* yo = y + (long long)n * HW * y_ld + c0;
  for (int e = threadIdx.x; e < items; e += blockDim.x) {
    const int p = e / VP, j = e - p * VP;
    a[e] = __ldg(reinterpret_cast<const uint4*>(xin + (long long)p * x_ld + j * V));
  }
  __syncthreads();
  for (int pool = 0; pool < 3; ++pool) {
    for (int e = threadIdx.x; e < items; e += blockDim.x) {     // row pass a -> b: consecutive lanes, consecutive 16-byte words
      const int p = e / VP, px = p % W;
      uint4 m = a[e];
#pragma unroll
      for (int d = -2; d <= 2; ++d)
        if (d != 0 && px + d >= 0 && px + d < W) m = vec_max<T>(m, a[e + d * VP]);
      b[e] = m;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < items; e += blockDim.x) {     // column pass b -> a (+ 16-byte stores)
      const int p = e / VP, j = e - p * VP, py = p / W;
      uint4 m = b[e];
#pragma unroll
      for (int d = -2; d <= 2; ++d)
        if (d != 0 && py + d >= 0 && py + d < H) m = vec_max<T>(m, b[e + d * W * VP]);
      a[e] = m;
      *reinterpret_cast<uint4*>(yo + (long long)p * y_ld + pool * C + j * V) = m;
    }
    __syncthreads();
  }
}

// The same three pools in TWO passes and one barrier: chained stride-1 max pools with -inf padding are single max pools over
// 5x5 / 9x9 / 13x13 windows clipped to the map, and a square window is separable, so one row pass produces the 5 / 9 / 13-wide
// row maxima of every pixel from the input tile (13 shared-memory reads) and one column pass reduces each over 5 / 9 / 13 rows
// (27 reads) straight into the three output slices.  The chained version above runs six passes with a barrier between
// each and was bound by that latency chain (20 x 20 map, 128 channels, B = 64: 32 us in the step for 26 MB of traffic).
template <typename T>
__global__ void __launch_bounds__(256)
sppf_pool_direct_kernel(const T* __restrict__ x, int x_ld, int H, int W, int C, T* __restrict__ y, int y_ld, int VP) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = Vec<T>::N;
  extern __shared__ uint4 sp4[];
  const int HW = H * W, items = HW * VP;
  uint4* a = sp4;                 // input [HW][VP]
  uint4* r5 = sp4 + items;        // row maxima of width 5 / 9 / 13
  uint4* r9 = r5 + items;
  uint4* r13 = r9 + items;
  const int n = blockIdx.y, c0 = blockIdx.x * VP * V;
  const T* xin = x + (long long)n * HW * x_ld + c0;
  T* yo = y + (long long)n * HW * y_ld + c0;
  for (int e = threadIdx.x; e < items; e += blockDim.x) {
    const int p = e / VP, j = e - p * VP;
    a[e] = __ldg(reinterpret_cast<const uint4*>(xin + (long long)p * x_ld + j * V));
  }
  __syncthreads();
  for (int e = threadIdx.x; e < items; e += blockDim.x) {
    const int p = e / VP, px = p % W;
    uint4 m = a[e];
#pragma unroll
    for (int d = 1; d <= 2; ++d) {
      if (px - d >= 0) m = vec_max<T>(m, a[e - d * VP]);
      if (px + d < W) m = vec_max<T>(m, a[e + d * VP]);
    }
    r5[e] = m;
#pragma unroll
    for (int d = 3; d <= 4; ++d) {
      if (px - d >= 0) m = vec_max<T>(m, a[e - d * VP]);
      if (px + d < W) m = vec_max<T>(m, a[e + d * VP]);
    }
    r9[e] = m;
#pragma unroll
    for (int d = 5; d <= 6; ++d) {
      if (px - d >= 0) m = vec_max<T>(m, a[e - d * VP]);
      if (px + d < W) m = vec_max<T>(m, a[e + d * VP]);
    }
    r13[e] = m;
  }
  __syncthreads();
  const int rs = W * VP;
  for (int e = threadIdx.x; e < items; e += blockDim.x) {
    const int p = e / VP, j = e - p * VP, py = p / W;
    uint4 m5 = r5[e], m9 = r9[e], m13 = r13[e];
#pragma unroll
    for (int d = 1; d <= 6; ++d) {
      const bool up = py - d >= 0, dn = py + d < H;
      if (d <= 2) {
        if (up) m5 = vec_max<T>(m5, r5[e - d * rs]);
        if (dn) m5 = vec_max<T>(m5, r5[e + d * rs]);
      }
      if (d <= 4) {
        if (up) m9 = vec_max<T>(m9, r9[e - d * rs]);
        if (dn) m9 = vec_max<T>(m9, r9[e + d * rs]);
      }
      if (up) m13 = vec_max<T>(m13, r13[e - d * rs]);
      if (dn) m13 = vec_max<T>(m13, r13[e + d * rs]);
    }
    T* dst = yo + (long long)p * y_ld + j * V;
    *reinterpret_cast<uint4*>(dst) = m5;
    *reinterpret_cast<uint4*>(dst + C) = m9;
    *reinterpret_cast<uint4*>(dst + 2 * C) = m13;
  }
}

// ---- bf16 depthwise conv, TMA-staged (production path) ---------------------------------------------------------------
// The register-window kernel above spends ~45 % of its ~1060 instructions per thread on addressing and bounds
// predicates of its 18 global loads and hides their latency only through occupancy (ncu: issue-active 52 %, warps
// active 22 %, profiles/r01_h_ncu_dwconv.md).  Here a persistent CTA receives the whole input tile - TH x TW outputs
// plus the filter halo, CB channels - as ONE 4-D TMA box per tile (zero fill outside the map = the conv padding, so no
// predicates), double buffered on two mbarriers; every thread owns one 16-byte channel vector of PX adjacent outputs
// and reads its (K rows) x (NV columns) window with LDS.128 at tile-constant offsets.  Outputs go straight from
// registers to global memory (16-byte stores, a 64/128-byte run per pixel).

// acc0 += lo(x) * lo(w), acc1 += hi(x) * hi(w): bf16 operands taken straight from the packed words, fp32 accumulate
__device__ __forceinline__ void fhfma2(float& a0, float& a1, uint32_t x, uint32_t w) {
  asm("{\n\t.reg .b16 xl, xh, wl, wh;\n\t"
      "mov.b32 {xl, xh}, %2;\n\t"
      "mov.b32 {wl, wh}, %3;\n\t"
      "fma.rn.f32.bf16 %0, xl, wl, %0;\n\t"
      "fma.rn.f32.bf16 %1, xh, wh, %1;\n\t}"
      : "+f"(a0), "+f"(a1)
      : "r"(x), "r"(w));
}

struct DwTmaParams {
  int C, Ho, Wo, B;
  int CB, CV, XG, TH;          // tile = TH rows x (XG * PX) columns x CB channels; CV = CB / 8 channel vectors
  int IW, IH;                  // input tile extent (with halo)
  int tiles_x, tiles_y, cblks, ntiles;
  int pad, act;
  float inv_tiles_x, inv_tiles_y;   // fast_div reciprocals (ntiles < 2^24 checked on the host)
  unsigned stage_bytes, tx_bytes;   // ring slot pitch (128-byte multiple) / bytes one box delivers
  const float* w;
  const float* bias;
  bf16* y;
  int y_ld;
  const bf16* res;
  int res_ld;
};

template <int K, int S, int D, int PX>
__global__ void __launch_bounds__(256, 2)
dwconv_tma_kernel(const __grid_constant__ CUtensorMap map, const __grid_constant__ DwTmaParams p) {
  extern __shared__ __align__(128) unsigned char dw_smem[];
  __shared__ __align__(8) unsigned long long dw_bars[2];
  constexpr int NV = (PX - 1) * S + (K - 1) * D + 1;
  constexpr bool WREG = (K == 3);            // 9 taps x 4 packed words = 36 registers; larger filters read shared memory
  const uint32_t base = (smem_u32(dw_smem) + 127u) & ~127u;
  const uint32_t wsm = base + 2u * p.stage_bytes;          // [K*K][CB] bf16: this CTA's filter taps
  const uint32_t bar0 = smem_u32(&dw_bars[0]);
  const int tid = threadIdx.x;
  const int cb = blockIdx.y;                 // the channel block is fixed per CTA: its taps are staged ONCE
  if (tid == 0) {
    prefetch_tmap(&map);
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // fp32 taps -> bf16 in shared memory (weights are not produced by the previous kernel: this runs before pdl_wait)
  for (int i = tid; i < K * K * (p.CB / 2); i += 256) {
    const int tap = i / (p.CB / 2), c = (i - tap * (p.CB / 2)) * 2;
    const float2 wf = __ldg(reinterpret_cast<const float2*>(p.w + (long long)tap * p.C + cb * p.CB + c));
    const __nv_bfloat162 h = __floats2bfloat162_rn(wf.x, wf.y);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(wsm + (uint32_t)(tap * p.CB + c) * 2u), "r"(*reinterpret_cast<const uint32_t*>(&h)));
  }
  __syncthreads();
  pdl_trigger();
  pdl_wait();

  auto issue = [&](int t, int stage) {       // one elected thread: arm the barrier, fetch spatial tile t of this channel block
    const int r1 = fast_div(t, p.tiles_x, p.inv_tiles_x), tx = t - r1 * p.tiles_x;
    const int n = fast_div(r1, p.tiles_y, p.inv_tiles_y), ty = r1 - n * p.tiles_y;
    mbar_expect_tx(bar0 + 8u * stage, p.tx_bytes);
    tma_load_4d(base + stage * p.stage_bytes, &map, bar0 + 8u * stage, cb * p.CB, tx * (p.XG * PX) * S - p.pad,
                ty * p.TH * S - p.pad, n);
  };

  // this thread's work item inside every tile
  const int items = p.TH * p.XG * p.CV;
  const bool worker = tid < items;
  const int cv = tid % p.CV;
  const int xg = (tid / p.CV) % p.XG;
  const int ry = tid / (p.CV * p.XG);
  const uint32_t pix_b = (uint32_t)p.CB * 2u, row_b = (uint32_t)p.IW * pix_b;
  const uint32_t my_off = (uint32_t)(ry * S) * row_b + (uint32_t)(xg * PX * S) * pix_b + (uint32_t)cv * 16u;
  const int c0 = cb * p.CB + cv * 8;
  const uint32_t wsrc = wsm + (uint32_t)cv * 16u;

  uint4 wreg[WREG ? K * K : 1];
  if (WREG) {
#pragma unroll
    for (int tap = 0; tap < K * K; ++tap)
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(wreg[tap].x), "=r"(wreg[tap].y), "=r"(wreg[tap].z), "=r"(wreg[tap].w)
                   : "r"(wsrc + (uint32_t)tap * pix_b));
  }
  float bv[8];
  {
    const float4 b0 = (p.bias && worker) ? __ldg(reinterpret_cast<const float4*>(p.bias + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 b1 = (p.bias && worker) ? __ldg(reinterpret_cast<const float4*>(p.bias + c0 + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
  }

  int t = blockIdx.x;
  if (tid == 0 && t < p.ntiles) issue(t, 0);
  int it = 0;
  for (; t < p.ntiles; t += gridDim.x, ++it) {
    const int stage = it & 1;
    const int tn = t + gridDim.x;
    if (tid == 0 && tn < p.ntiles) issue(tn, stage ^ 1);    // the other buffer was released by the barrier below
    const int r1 = fast_div(t, p.tiles_x, p.inv_tiles_x), tx = t - r1 * p.tiles_x;
    const int n = fast_div(r1, p.tiles_y, p.inv_tiles_y), ty = r1 - n * p.tiles_y;
    const int oy = ty * p.TH + ry, ox0 = (tx * p.XG + xg) * PX;

    mbar_wait(bar0 + 8u * stage, (uint32_t)((it >> 1) & 1));
    if (worker) {
      const uint32_t src = base + stage * p.stage_bytes + my_off;
      // bf16 x bf16 -> fp32 mixed-precision FMA (fma.rn.f32.bf16 = one FHFMA.BF16 with .H0/.H1 operand selectors): the
      // packed input words are consumed as they are, so there is no bf16 -> fp32 unpack at all (it was ~45 % of the
      // ALU-pipe work of the fp32 FFMA2 formulation); the BN-folded filter taps are rounded to bf16 like the weights of
      // every dense conv of this mode, products are exact and the accumulation stays fp32.
      float acc[PX][8];
#pragma unroll
      for (int q = 0; q < PX; ++q)
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[q][v] = bv[v];
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        uint4 in[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j)
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(in[j].x), "=r"(in[j].y), "=r"(in[j].z), "=r"(in[j].w)
                       : "r"(src + (uint32_t)(ky * D) * row_b + (uint32_t)j * pix_b));
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          uint4 wq;
          if (WREG) {
            wq = wreg[ky * K + kx];
          } else {
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(wq.x), "=r"(wq.y), "=r"(wq.z), "=r"(wq.w)
                         : "r"(wsrc + (uint32_t)(ky * K + kx) * pix_b));
          }
#pragma unroll
          for (int q = 0; q < PX; ++q) {
            const uint4 xv = in[q * S + kx * D];
            fhfma2(acc[q][0], acc[q][1], xv.x, wq.x);
            fhfma2(acc[q][2], acc[q][3], xv.y, wq.y);
            fhfma2(acc[q][4], acc[q][5], xv.z, wq.z);
            fhfma2(acc[q][6], acc[q][7], xv.w, wq.w);
          }
        }
      }
      float2 acc2[PX][4];
#pragma unroll
      for (int q = 0; q < PX; ++q)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc2[q][v] = make_float2(acc[q][2 * v], acc[q][2 * v + 1]);
      switch (p.act) {
        case LPC_ACT_NONE: break;
        case LPC_ACT_SILU:
#pragma unroll
          for (int q = 0; q < PX; ++q)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc2[q][v] = silu2_(acc2[q][v]);
          break;
        case LPC_ACT_MISH:
#pragma unroll
          for (int q = 0; q < PX; ++q)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc2[q][v] = mish2_(acc2[q][v]);
          break;
        default:
#pragma unroll
          for (int q = 0; q < PX; ++q)
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              acc2[q][v].x = apply_act<false>(acc2[q][v].x, p.act);
              acc2[q][v].y = apply_act<false>(acc2[q][v].y, p.act);
            }
          break;
      }
      if (oy < p.Ho) {
#pragma unroll
        for (int q = 0; q < PX; ++q) {
          const int ox = ox0 + q;
          if (ox >= p.Wo) break;
          const long long opix = (long long)(n * p.Ho + oy) * p.Wo + ox;
          if (p.res) {
            float rr[8];
            ld_vec<bf16>(p.res + opix * p.res_ld + c0).unpack(rr);
#pragma unroll
            for (int v = 0; v < 4; ++v) { acc2[q][v].x += rr[2 * v]; acc2[q][v].y += rr[2 * v + 1]; }
          }
          uint4 o;
          __nv_bfloat162 h;
          h = __floats2bfloat162_rn(acc2[q][0].x, acc2[q][0].y); o.x = *reinterpret_cast<uint32_t*>(&h);
          h = __floats2bfloat162_rn(acc2[q][1].x, acc2[q][1].y); o.y = *reinterpret_cast<uint32_t*>(&h);
          h = __floats2bfloat162_rn(acc2[q][2].x, acc2[q][2].y); o.z = *reinterpret_cast<uint32_t*>(&h);
          h = __floats2bfloat162_rn(acc2[q][3].x, acc2[q][3].y); o.w = *reinterpret_cast<uint32_t*>(&h);
          *reinterpret_cast<uint4*>(p.y + opix * p.y_ld + c0) = o;
        }
      }
    }
    __syncthreads();          // every read of this stage is done: the next iteration may refill it
  }
}

typedef CUresult (*DwEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
DwEncodeFn dw_get_encode() {
  static DwEncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<DwEncodeFn>(f);
  });
  return fn;
}
int dw_num_sms() { return lpc_num_sms(); }

// returns LPC_OK when launched, 1 when the shape is left to the register-window kernel
template <int K, int S, int D>
int launch_dw_tma(const void* x, int x_ld, int B, int H, int W, int C, const float* w, const float* bias, int pad,
                  int Ho, int Wo, void* y, int y_ld, int act, const void* res, int res_ld, cudaStream_t s) {
  constexpr int PX = (K == 7) ? 2 : 4;
  static const int enabled = [] { const char* e = getenv("LPC_DW_TMA"); return e ? atoi(e) : 1; }();
  DwEncodeFn enc = dw_get_encode();
  if (!enabled || !enc || C % 8 || H > 65535 || W > 65535) return 1;
  // tile search: CB channels (a divisor of C, multiple of 8, <= 128), XG x-groups of PX outputs, TH rows, at most 256
  // work items; minimise the number of tile passes, then the halo overhead
  DwTmaParams p;
  memset(&p, 0, sizeof(p));
  double best = 1e30;
  for (int cbv = 1; cbv <= 16; ++cbv) {
    if ((C / 8) % cbv) continue;
    if (cbv % 4 && cbv != C / 8) continue;      // 64-byte runs per pixel (or the whole pixel): full 32-byte sectors
    // Two outputs per thread (7x7) and 64-byte pixels put consecutive x-groups 128 bytes apart: every quarter-warp of a
    // 16-byte shared-memory load hits the same banks twice.  Measured 512ch 20x20 B256: cbv 4 -> 522 us, cbv 8 / 16 ->
    // 207 / 201 us (gpurun_out/dw7_tiles.txt); 128-byte pixels are conflict-free.
    if (PX == 2 && cbv < 8 && (C / 8) % 8 == 0) continue;
    for (int th = 4; th <= 16; th *= 2)
      for (int xg = 1; xg * cbv * th <= 256; ++xg) {
        const int tw = xg * PX;
        const int iw = (tw - 1) * S + (K - 1) * D + 1, ih = (th - 1) * S + (K - 1) * D + 1;
        if (iw > 256 || ih > 256) continue;
        const size_t bytes = (size_t)iw * ih * cbv * 16;
        if (bytes > 44 * 1024) continue;
        const double tiles = (double)((Wo + tw - 1) / tw) * ((Ho + th - 1) / th) * (C / 8 / cbv);
        // cost ~ per-tile pass (fixed 256-thread sweep + barrier) plus the bytes the tile moves, weighted against short
        // per-pixel runs: a TMA box row of 64 bytes costs about as much as one of 192 (measured 192ch 80x80 B256, us:
        // cbv 4 -> 434, cbv 8 -> 399, cbv 12 -> 276; 64ch 80x80 B64: cbv 4 -> 27.7, cbv 8 -> 26.8)
        const double cost = tiles * (1.0 + (double)bytes / (24.0 * 1024)) * (1.0 + 4.0 / cbv);
        if (cost < best) {
          best = cost;
          p.CV = cbv; p.CB = cbv * 8; p.XG = xg; p.TH = th; p.IW = iw; p.IH = ih;
          p.stage_bytes = (unsigned)((bytes + 127) & ~(size_t)127);
          p.tx_bytes = (unsigned)bytes;
        }
      }
  }
  if (const char* e = getenv("LPC_DW_TILE")) {           // "cbv,xg,th": force a tile (tuning runs only)
    int cbv = 0, xg = 0, th = 0;
    if (sscanf(e, "%d,%d,%d", &cbv, &xg, &th) == 3 && cbv > 0 && (C / 8) % cbv == 0 && xg * cbv * th <= 256) {
      const int tw = xg * PX, iw = (tw - 1) * S + (K - 1) * D + 1, ih = (th - 1) * S + (K - 1) * D + 1;
      const size_t bytes = (size_t)iw * ih * cbv * 16;
      if (iw <= 256 && ih <= 256 && bytes <= 44 * 1024) {
        best = 0;
        p.CV = cbv; p.CB = cbv * 8; p.XG = xg; p.TH = th; p.IW = iw; p.IH = ih;
        p.stage_bytes = (unsigned)((bytes + 127) & ~(size_t)127);
        p.tx_bytes = (unsigned)bytes;
      }
    }
  }
  if (getenv("LPC_DW_TILE_PRINT")) fprintf(stderr, "dw tile K=%d S=%d C=%d %dx%d: cbv=%d xg=%d th=%d iw=%d ih=%d\n", K, S, C, Ho, Wo, p.CV, p.XG, p.TH, p.IW, p.IH);
  if (best >= 1e30) return 1;
  p.C = C; p.Ho = Ho; p.Wo = Wo; p.B = B;
  p.tiles_x = (Wo + p.XG * PX - 1) / (p.XG * PX);
  p.tiles_y = (Ho + p.TH - 1) / p.TH;
  p.cblks = C / p.CB;
  const long long nt = (long long)p.tiles_x * p.tiles_y * B;      // spatial tiles per channel block
  if (nt >= (1ll << 24)) return 1;
  p.ntiles = (int)nt;
  p.inv_tiles_x = 1.0f / (float)p.tiles_x;
  p.inv_tiles_y = 1.0f / (float)p.tiles_y;
  p.pad = pad; p.act = act; p.w = w; p.bias = bias;
  p.y = (bf16*)y; p.y_ld = y_ld; p.res = (const bf16*)res; p.res_ld = res_ld;
  CUtensorMap map;
  {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)x_ld * 2, (cuuint64_t)W * x_ld * 2, (cuuint64_t)H * W * x_ld * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.CB, (cuuint32_t)p.IW, (cuuint32_t)p.IH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return 1;
  }
  const size_t smem = 2 * (size_t)p.stage_bytes + (size_t)K * K * p.CB * 2 + 128;
  static unsigned long long attr_done = 0;     // per device
  if (lpc_first_on_device(&attr_done)) cudaFuncSetAttribute(dwconv_tma_kernel<K, S, D, PX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int per_sm = (int)((200 * 1024) / (smem + 1024));
  (void)per_sm;
  static const int cta_cap = [] { const char* e = getenv("LPC_CTA_CAP"); return e ? atoi(e) : 2; }();
  long long gx = ((long long)(cta_cap >= 2 ? 2 : 1) * dw_num_sms() + p.cblks - 1) / p.cblks;      // two resident CTAs per SM in total, split over the channel blocks
  if (gx > nt) gx = nt;
  if (gx < 1) gx = 1;
  if (p.cblks > 65535) return 1;
  lpc_launch_pdl(dwconv_tma_kernel<K, S, D, PX>, dim3((unsigned)gx, (unsigned)p.cblks), 256, smem, s, map, p);
  LPC_CHECK_LAUNCH("dwconv2d(tma)");
  return LPC_OK;
}

template <typename T, int K, int S, int D>
int launch_dw(const void* x, int x_ld, int B, int H, int W, int C, const float* w, const float* bias, int pad,
              int Ho, int Wo, void* y, int y_ld, int act, const void* res, int res_ld, cudaStream_t s) {
  constexpr int PX = (K == 7 || (sizeof(T) == 4 && K >= 5)) ? 2 : 4;   // keeps the fp32 row window within the register file
  constexpr int V = Vec<T>::N;
  const long long per_strip = (long long)8 * ((Wo + PX - 1) / PX) * (C / V);
  LPC_REQUIRE(per_strip < (1ll << 31) && B <= 65535 && (Ho + 7) / 8 <= 65535, "dwconv2d: shape too large");
  dim3 grid(cdiv(per_strip, 128), (Ho + 7) / 8, B);
  lpc_launch_pdl(dwconv_kernel<T, K, S, D, PX>, grid, 128, 0, s, (const T*)x, x_ld, B, H, W, C, w, bias, pad, Ho, Wo,
                                                                (T*)y, y_ld, act, (const T*)res, res_ld);
  LPC_CHECK_LAUNCH("dwconv2d");
  return LPC_OK;
}

template <typename T>
int dispatch_dw(const void* x, int x_ld, int B, int H, int W, int C, const float* w, const float* bias, int k,
                int stride, int pad, int dil, int Ho, int Wo, void* y, int y_ld, int act, const void* res,
                int res_ld, cudaStream_t s) {
#define DW(K_, S_, D_)                                                                                                      \
  if (k == K_ && stride == S_ && dil == D_) {                                                                               \
    if (sizeof(T) == 2) {                                                                                                   \
      const int r = launch_dw_tma<K_, S_, D_>(x, x_ld, B, H, W, C, w, bias, pad, Ho, Wo, y, y_ld, act, res, res_ld, s);      \
      if (r != 1) return r;                                                                                                 \
    }                                                                                                                       \
    return launch_dw<T, K_, S_, D_>(x, x_ld, B, H, W, C, w, bias, pad, Ho, Wo, y, y_ld, act, res, res_ld, s);               \
  }
  DW(3, 1, 1) DW(3, 2, 1) DW(3, 1, 2) DW(3, 1, 3) DW(5, 1, 1) DW(5, 2, 1) DW(7, 1, 1) DW(7, 2, 1)
#undef DW
  LPC_FAIL(LPC_E_UNSUPPORTED, "dwconv2d: k=%d stride=%d dilation=%d not supported", k, stride, dil);
}

}  // namespace

// fp32 validation mode: one thread per (output pixel, channel), products and sums in fp64, bias / activation / residual
// in fp64, ONE rounding to fp32 per output (same contract as conv_direct_kernel<float>; speed is irrelevant here).
__global__ void __launch_bounds__(256)
dwconv_f32_validate_kernel(const float* __restrict__ x, int x_ld, int B, int H, int W, int C, const float* __restrict__ w,
                           const float* __restrict__ bias, int K, int S, int pad, int D, int Ho, int Wo,
                           float* __restrict__ y, int y_ld, int act, const float* __restrict__ res, int res_ld) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)B * Ho * Wo * C;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const long long pix = idx / C;
  const int ox = (int)(pix % Wo);
  const int oy = (int)((pix / Wo) % Ho);
  const int n = (int)(pix / ((long long)Wo * Ho));
  double acc = bias ? (double)bias[c] : 0.0;
  for (int ky = 0; ky < K; ++ky) {
    const int iy = oy * S - pad + ky * D;
    if (iy < 0 || iy >= H) continue;
    for (int kx = 0; kx < K; ++kx) {
      const int ix = ox * S - pad + kx * D;
      if (ix < 0 || ix >= W) continue;
      acc = fma((double)x[((long long)(n * H + iy) * W + ix) * x_ld + c], (double)w[(ky * K + kx) * C + c], acc);
    }
  }
  acc = apply_act_f64(acc, act);
  if (res) acc += (double)res[pix * res_ld + c];
  y[pix * y_ld + c] = (float)acc;
}

extern "C" int lpc_dwconv2d(int dtype, const void* x, int x_ld, int B, int H, int W, int C, const float* w,
                            const float* bias, int k, int stride, int pad, int dil, void* y, int y_ld, int act,
                            const void* res, int res_ld, void* stream) {
  LPC_REQUIRE(x && w && y, "dwconv2d: null pointer");
  LPC_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && dil >= 1 && pad >= 0, "dwconv2d: bad shape");
  const int V = dtype == LPC_F32 ? 4 : 8;
  LPC_REQUIRE(C % V == 0 && x_ld % V == 0 && y_ld % V == 0 && (!res || res_ld % V == 0),
              "dwconv2d: C / pitches must be multiples of %d", V);
  LPC_REQUIRE(aligned16(x) && aligned16(y) && aligned16(res), "dwconv2d: pointers must be 16-byte aligned");
  const int ke = dil * (k - 1) + 1;
  const int Ho = (H + 2 * pad - ke) / stride + 1, Wo = (W + 2 * pad - ke) / stride + 1;
  LPC_REQUIRE(Ho > 0 && Wo > 0, "dwconv2d: empty output");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == LPC_F32) {
    static const int fast32 = [] { const char* e = getenv("LPC_DW_F32_FAST"); return e ? atoi(e) : 0; }();
    if (fast32) return dispatch_dw<float>(x, x_ld, B, H, W, C, w, bias, k, stride, pad, dil, Ho, Wo, y, y_ld, act, res, res_ld, s);
    const long long total = (long long)B * Ho * Wo * C;
    lpc_launch_pdl(dwconv_f32_validate_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, (const float*)x, x_ld, B, H, W, C, w, bias,
                   k, stride, pad, dil, Ho, Wo, (float*)y, y_ld, act, (const float*)res, res_ld);
    LPC_CHECK_LAUNCH("dwconv2d (fp32 validation)");
    return LPC_OK;
  }
  if (dtype == LPC_BF16) return dispatch_dw<bf16>(x, x_ld, B, H, W, C, w, bias, k, stride, pad, dil, Ho, Wo, y, y_ld, act, res, res_ld, s);
  LPC_FAIL(LPC_E_ARG, "dwconv2d: unknown dtype %d", dtype);
}

extern "C" int lpc_sppf_pool(int dtype, const void* x, int x_ld, int B, int H, int W, int C, void* y, int y_ld,
                             void* stream) {
  LPC_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && C > 0, "sppf_pool: bad argument");
  const int V = dtype == LPC_F32 ? 4 : 8;
  LPC_REQUIRE(C % V == 0 && x_ld % V == 0 && y_ld % V == 0 && y_ld >= 3 * C, "sppf_pool: C / pitch constraints");
  LPC_REQUIRE(aligned16(x) && aligned16(y), "sppf_pool: pointers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  LPC_REQUIRE(B <= 65535, "sppf_pool: batch too large");
  static const int chained = [] { const char* e = getenv("LPC_SPPF_CHAINED"); return e ? atoi(e) : 0; }();
  if (!chained && (size_t)H * W * 16 * 4 <= 100 * 1024) {
    // direct two-pass kernel: four planes (input + three row-maxima planes); two 16-byte channel vectors per pixel (whole 32-byte
    // sectors on both the loads and the stores) when the planes still leave room for four CTAs per SM
    const int VP = (C % (2 * V) == 0 && (size_t)H * W * 2 * 16 * 4 <= 56 * 1024) ? 2 : 1;
    const size_t smem = (size_t)H * W * VP * 16 * 4;
    dim3 grid(C / (VP * V), B);
    if (dtype == LPC_F32) {
      cudaFuncSetAttribute(sppf_pool_direct_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      lpc_launch_pdl(sppf_pool_direct_kernel<float>, grid, 256, smem, s, (const float*)x, x_ld, H, W, C, (float*)y, y_ld, VP);
    } else if (dtype == LPC_BF16) {
      cudaFuncSetAttribute(sppf_pool_direct_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      lpc_launch_pdl(sppf_pool_direct_kernel<bf16>, grid, 256, smem, s, (const bf16*)x, x_ld, H, W, C, (bf16*)y, y_ld, VP);
    } else
      LPC_FAIL(LPC_E_ARG, "sppf_pool: unknown dtype %d", dtype);
    LPC_CHECK_LAUNCH("sppf_pool");
    return LPC_OK;
  }
  // two 16-byte channel vectors per pixel per CTA when the channel count and the shared-memory budget allow
  int VP = (C % (2 * V) == 0 && (size_t)H * W * 2 * 16 * 2 <= 100 * 1024) ? 2 : 1;
  const size_t smem = (size_t)H * W * VP * 16 * 2;
  LPC_REQUIRE(smem <= 200 * 1024, "sppf_pool: map too large for the shared-memory pooling kernel (%d x %d)", H, W);
  dim3 grid(C / (VP * V), B);
  if (dtype == LPC_F32) {
    cudaFuncSetAttribute(sppf_pool_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    lpc_launch_pdl(sppf_pool_kernel<float>, grid, 256, smem, s, (const float*)x, x_ld, H, W, C, (float*)y, y_ld, VP);
  } else if (dtype == LPC_BF16) {
    cudaFuncSetAttribute(sppf_pool_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    lpc_launch_pdl(sppf_pool_kernel<bf16>, grid, 256, smem, s, (const bf16*)x, x_ld, H, W, C, (bf16*)y, y_ld, VP);
  } else
    LPC_FAIL(LPC_E_ARG, "sppf_pool: unknown dtype %d", dtype);
  LPC_CHECK_LAUNCH("sppf_pool");
  return LPC_OK;
}
