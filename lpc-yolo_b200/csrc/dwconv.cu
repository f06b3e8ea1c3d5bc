// dwconv.cu - depthwise convolution and SPPF pooling, NHWC, bandwidth-bound CUDA-core kernels.
//
// Channel vectors (16 bytes: 8 bf16 / 4 fp32) are the fastest-varying index across threads -> coalesced 128-bit accesses.
#include "common.cuh"

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

// SPPF: three chained MaxPool2d(5,1,2) (-inf padding).  One CTA owns one image x 8 channels, keeps the whole map in
// shared memory and runs each 5x5 pool as a separable row pass + column pass (10 compares instead of 25); pool i's
// result is stored to its concat slice and is the input of pool i+1.
template <typename T>
__global__ void __launch_bounds__(256)
sppf_pool_kernel(const T* __restrict__ x, int x_ld, int H, int W, int C, T* __restrict__ y, int y_ld) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = Vec<T>::N;
  extern __shared__ float4 sp4[];
  const int HW = H * W;
  float4* a = sp4;               // [HW][2] float4 = 8 channels per pixel
  float4* b = sp4 + HW * 2;
  const int n = blockIdx.y, c0 = blockIdx.x * 8;
  const T* xin = x + (long long)n * HW * x_ld + c0;
  T* yo = y + (long long)n * HW * y_ld + c0;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {          // thread = pixel: 16-byte loads
    float f[8];
#pragma unroll
    for (int v = 0; v < 8; v += V) ldg_vec<T>(xin + (long long)p * x_ld + v).unpack(f + v);
    a[p * 2] = make_float4(f[0], f[1], f[2], f[3]);
    a[p * 2 + 1] = make_float4(f[4], f[5], f[6], f[7]);
  }
  __syncthreads();
  auto max4 = [](float4 u, float4 v) { return make_float4(fmaxf(u.x, v.x), fmaxf(u.y, v.y), fmaxf(u.z, v.z), fmaxf(u.w, v.w)); };
  for (int pool = 0; pool < 3; ++pool) {
    for (int e = threadIdx.x; e < HW * 2; e += blockDim.x) {     // row pass a -> b (thread = pixel x 4 channels)
      const int p = e >> 1, py = p / W, px = p - py * W;
      float4 m = a[e];
#pragma unroll
      for (int d = -2; d <= 2; ++d) {
        const int xx = px + d;
        if (d != 0 && xx >= 0 && xx < W) m = max4(m, a[e + 2 * d]);
      }
      b[e] = m;
    }
    __syncthreads();
    for (int p = threadIdx.x; p < HW; p += blockDim.x) {         // column pass b -> a (+ 16-byte stores)
      const int py = p / W;
      float4 m0 = b[p * 2], m1 = b[p * 2 + 1];
#pragma unroll
      for (int d = -2; d <= 2; ++d) {
        const int yy = py + d;
        if (d != 0 && yy >= 0 && yy < H) {
          m0 = max4(m0, b[(p + d * W) * 2]);
          m1 = max4(m1, b[(p + d * W) * 2 + 1]);
        }
      }
      a[p * 2] = m0;
      a[p * 2 + 1] = m1;
      const float f[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
      for (int v = 0; v < 8; v += V) {
        Vec<T> o;
        o.pack(f + v);
        st_vec<T>(yo + (long long)p * y_ld + pool * C + v, o);
      }
    }
    __syncthreads();
  }
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
#define DW(K_, S_, D_) \
  if (k == K_ && stride == S_ && dil == D_) return launch_dw<T, K_, S_, D_>(x, x_ld, B, H, W, C, w, bias, pad, Ho, Wo, y, y_ld, act, res, res_ld, s);
  DW(3, 1, 1) DW(3, 2, 1) DW(3, 1, 2) DW(3, 1, 3) DW(5, 1, 1) DW(5, 2, 1) DW(7, 1, 1) DW(7, 2, 1)
#undef DW
  LPC_FAIL(LPC_E_UNSUPPORTED, "dwconv2d: k=%d stride=%d dilation=%d not supported", k, stride, dil);
}

}  // namespace

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
  if (dtype == LPC_F32) return dispatch_dw<float>(x, x_ld, B, H, W, C, w, bias, k, stride, pad, dil, Ho, Wo, y, y_ld, act, res, res_ld, s);
  if (dtype == LPC_BF16) return dispatch_dw<bf16>(x, x_ld, B, H, W, C, w, bias, k, stride, pad, dil, Ho, Wo, y, y_ld, act, res, res_ld, s);
  LPC_FAIL(LPC_E_ARG, "dwconv2d: unknown dtype %d", dtype);
}

extern "C" int lpc_sppf_pool(int dtype, const void* x, int x_ld, int B, int H, int W, int C, void* y, int y_ld,
                             void* stream) {
  LPC_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && C > 0, "sppf_pool: bad argument");
  const int V = dtype == LPC_F32 ? 4 : 8;
  LPC_REQUIRE(C % V == 0 && x_ld % V == 0 && y_ld % V == 0 && y_ld >= 3 * C, "sppf_pool: C / pitch constraints");
  LPC_REQUIRE(aligned16(x) && aligned16(y), "sppf_pool: pointers must be 16-byte aligned");
  LPC_REQUIRE(C % 8 == 0, "sppf_pool: C must be a multiple of 8");
  const size_t smem = (size_t)H * W * 8 * 2 * sizeof(float);
  LPC_REQUIRE(smem <= 200 * 1024, "sppf_pool: map too large for the shared-memory pooling kernel (%d x %d)", H, W);
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid(C / 8, B);
  if (dtype == LPC_F32) {
    cudaFuncSetAttribute(sppf_pool_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    lpc_launch_pdl(sppf_pool_kernel<float>, grid, 256, smem, s, (const float*)x, x_ld, H, W, C, (float*)y, y_ld);
  } else if (dtype == LPC_BF16) {
    cudaFuncSetAttribute(sppf_pool_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    lpc_launch_pdl(sppf_pool_kernel<bf16>, grid, 256, smem, s, (const bf16*)x, x_ld, H, W, C, (bf16*)y, y_ld);
  } else
    LPC_FAIL(LPC_E_ARG, "sppf_pool: unknown dtype %d", dtype);
  LPC_CHECK_LAUNCH("sppf_pool");
  return LPC_OK;
}
