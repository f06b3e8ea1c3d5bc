// dwconv.cu - depthwise convolution and SPPF pooling, NHWC, bandwidth-bound CUDA-core kernels.
//
// One thread owns one 16-byte channel vector (8 bf16 / 4 fp32) of PX horizontally adjacent output
// pixels, so a 3x3 stride-1 window loads 3*(PX+2) input vectors for PX outputs instead of 9*PX.
// Channel vectors are the fastest-varying index across threads -> fully coalesced 128-bit accesses.
#include "common.cuh"

namespace {

template <typename T, int K, int S, int PX>
__global__ void __launch_bounds__(256)
dwconv_kernel(const T* __restrict__ x, int x_ld, int B, int H, int W, int C,
              const float* __restrict__ w, const float* __restrict__ bias, int pad, int dil,
              int Ho, int Wo, T* __restrict__ y, int y_ld, int act,
              const T* __restrict__ res, int res_ld) {
  constexpr int V = Vec<T>::N;
  constexpr bool PR = Precise<T>::value;
  const int cvecs = C / V;
  const int wgroups = (Wo + PX - 1) / PX;
  const long long total = (long long)B * Ho * wgroups * cvecs;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = (int)(idx % cvecs);
  long long t = idx / cvecs;
  const int xg = (int)(t % wgroups);
  t /= wgroups;
  const int oy = (int)(t % Ho);
  const int n = (int)(t / Ho);
  const int c0 = cv * V;
  const int ox0 = xg * PX;

  float acc[PX][V];
#pragma unroll
  for (int p = 0; p < PX; ++p)
#pragma unroll
    for (int v = 0; v < V; ++v) acc[p][v] = bias ? bias[c0 + v] : 0.f;

  // input columns touched by the PX outputs: ix = (ox0+p)*S - pad + kx*dil
  constexpr int SPAN = (PX - 1) * S + 1;  // per-kx span of distinct input columns
#pragma unroll
  for (int ky = 0; ky < K; ++ky) {
    const int iy = oy * S - pad + ky * dil;
    if (iy < 0 || iy >= H) continue;
    const T* row = x + ((long long)(n * H + iy) * W) * x_ld + c0;
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
      float wv[V];
#pragma unroll
      for (int v = 0; v < V; ++v) wv[v] = __ldg(w + (ky * K + kx) * C + c0 + v);
#pragma unroll
      for (int p = 0; p < PX; ++p) {
        const int ix = (ox0 + p) * S - pad + kx * dil;
        if (ix < 0 || ix >= W || ox0 + p >= Wo) continue;
        float f[V];
        ldg_vec<T>(row + (long long)ix * x_ld).unpack(f);
#pragma unroll
        for (int v = 0; v < V; ++v) acc[p][v] = fmaf(f[v], wv[v], acc[p][v]);
      }
    }
  }
  (void)SPAN;
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    const int ox = ox0 + p;
    if (ox >= Wo) break;
    const long long opix = (long long)(n * Ho + oy) * Wo + ox;
    float o[V];
#pragma unroll
    for (int v = 0; v < V; ++v) o[v] = apply_act<PR>(acc[p][v], act);
    if (res) {
      float r[V];
      ld_vec<T>(res + opix * res_ld + c0).unpack(r);
#pragma unroll
      for (int v = 0; v < V; ++v) o[v] += r[v];
    }
    Vec<T> ov;
    ov.pack(o);
    st_vec<T>(y + opix * y_ld + c0, ov);
  }
}

// SPPF: three chained MaxPool2d(5,1,2) (-inf padding).  One CTA owns one image x 8 channels, keeps the whole map in
// shared memory and runs each 5x5 pool as a separable row pass + column pass (10 compares instead of 25); pool i's
// result is stored to its concat slice and is the input of pool i+1.
template <typename T>
__global__ void __launch_bounds__(256)
sppf_pool_kernel(const T* __restrict__ x, int x_ld, int H, int W, int C, T* __restrict__ y, int y_ld) {
  extern __shared__ float sp[];
  const int HW = H * W;
  float* a = sp;             // [HW][8]
  float* b = sp + HW * 8;    // [HW][8]
  const int n = blockIdx.y, c0 = blockIdx.x * 8;
  const T* xin = x + (long long)n * HW * x_ld + c0;
  T* yo = y + (long long)n * HW * y_ld + c0;
  for (int e = threadIdx.x; e < HW * 8; e += blockDim.x) a[e] = to_f(xin[(long long)(e >> 3) * x_ld + (e & 7)]);
  __syncthreads();
  for (int pool = 0; pool < 3; ++pool) {
    for (int e = threadIdx.x; e < HW * 8; e += blockDim.x) {       // row pass a -> b
      const int p = e >> 3, ch = e & 7, py = p / W, px = p - py * W;
      float m = -INFINITY;
#pragma unroll
      for (int d = -2; d <= 2; ++d) {
        const int xx = px + d;
        if (xx >= 0 && xx < W) m = fmaxf(m, a[(py * W + xx) * 8 + ch]);
      }
      b[e] = m;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < HW * 8; e += blockDim.x) {       // column pass b -> a (+ store)
      const int p = e >> 3, ch = e & 7, py = p / W, px = p - py * W;
      float m = -INFINITY;
#pragma unroll
      for (int d = -2; d <= 2; ++d) {
        const int yy = py + d;
        if (yy >= 0 && yy < H) m = fmaxf(m, b[(yy * W + px) * 8 + ch]);
      }
      a[e] = m;
      yo[(long long)p * y_ld + pool * C + ch] = from_f<T>(m);
    }
    __syncthreads();
  }
}

template <typename T, int K, int S>
int launch_dw(const void* x, int x_ld, int B, int H, int W, int C, const float* w, const float* bias, int pad,
              int dil, int Ho, int Wo, void* y, int y_ld, int act, const void* res, int res_ld, cudaStream_t s) {
  constexpr int PX = (S == 1) ? 4 : 2;
  constexpr int V = Vec<T>::N;
  const long long total = (long long)B * Ho * ((Wo + PX - 1) / PX) * (C / V);
  dwconv_kernel<T, K, S, PX><<<cdiv(total, 256), 256, 0, s>>>((const T*)x, x_ld, B, H, W, C, w, bias, pad, dil, Ho, Wo,
                                                             (T*)y, y_ld, act, (const T*)res, res_ld);
  LPC_CHECK_LAUNCH("dwconv2d");
  return LPC_OK;
}

template <typename T>
int dispatch_dw(const void* x, int x_ld, int B, int H, int W, int C, const float* w, const float* bias, int k,
                int stride, int pad, int dil, int Ho, int Wo, void* y, int y_ld, int act, const void* res,
                int res_ld, cudaStream_t s) {
#define DW(K_, S_) \
  if (k == K_ && stride == S_) return launch_dw<T, K_, S_>(x, x_ld, B, H, W, C, w, bias, pad, dil, Ho, Wo, y, y_ld, act, res, res_ld, s);
  DW(3, 1) DW(3, 2) DW(5, 1) DW(5, 2) DW(7, 1) DW(7, 2)
#undef DW
  LPC_FAIL(LPC_E_UNSUPPORTED, "dwconv2d: k=%d stride=%d not supported", k, stride);
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
    sppf_pool_kernel<float><<<grid, 256, smem, s>>>((const float*)x, x_ld, H, W, C, (float*)y, y_ld);
  } else if (dtype == LPC_BF16) {
    cudaFuncSetAttribute(sppf_pool_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sppf_pool_kernel<bf16><<<grid, 256, smem, s>>>((const bf16*)x, x_ld, H, W, C, (bf16*)y, y_ld);
  } else
    LPC_FAIL(LPC_E_ARG, "sppf_pool: unknown dtype %d", dtype);
  LPC_CHECK_LAUNCH("sppf_pool");
  return LPC_OK;
}
