// conv_direct.cu - CUDA-core implicit-GEMM convolution (NHWC), any k / stride / pad.
//
// Role: (1) the fp32 VALIDATION mode (tensor cores have no fp32 path; SURVEY.md section 7 hard part 4),
// (2) the bf16 path for shapes the tcgen05 kernel does not take (the Cin=3 stem).  The production bf16
// dense convs run in conv_tc.cu.
//
// Tiling: CTA = 64 output pixels x 64 output channels, 256 threads, 4x4 register tile per thread,
// K loop over (tap, 16-channel slab) staged through shared memory.
//
// fp32 validation mode: products and sums are carried in fp64 (an fp32 x fp32 product is exact in fp64), bias,
// activation, channel gate and residual are applied in fp64 too and the result is rounded to fp32 ONCE per layer, so the
// only error a layer adds is the storage rounding of its output.  The reference's own fp32 path (fp32 accumulation in
// ATen's order + separate BN / activation / add roundings) sits 3e-6 ... 5e-5 from an fp64 run; this mode must land inside
// that (tests/test_gpu_e2e.py::test_fp32_mode_raw_and_y).
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

template <typename T> struct Acc { typedef float type; };
template <> struct Acc<float> { typedef double type; };

template <typename T>
__global__ void __launch_bounds__(NT)
conv_direct_kernel(const T* __restrict__ x, int x_ld, int B, int H, int W, int Cin,
                   const T* __restrict__ w, const float* __restrict__ bias, int k, int stride, int pad,
                   int Cout, int Ho, int Wo, T* __restrict__ y, int y_ld, int act,
                   const float* __restrict__ chan_scale, const T* __restrict__ res, int res_ld) {
  pdl_trigger();
  pdl_wait();
  constexpr bool PR = Precise<T>::value;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN];

  const int tid = threadIdx.x;
  const long long M = (long long)B * Ho * Wo;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // A-load role: thread -> (pixel lp, 4 consecutive channels lc)
  const int lp = tid >> 2, lc = (tid & 3) * 4;
  long long mA = m0 + lp;
  const bool mvalid = mA < M;
  int an = 0, aoy = 0, aox = 0;
  if (mvalid) {
    an = (int)(mA / ((long long)Ho * Wo));
    int r = (int)(mA - (long long)an * Ho * Wo);
    aoy = r / Wo;
    aox = r - aoy * Wo;
  }
  // B-load role: thread -> (k row bk, 4 consecutive couts bn)
  const int bk = tid >> 4, bn = (tid & 15) * 4;

  const int ty = tid >> 4, tx = tid & 15;
  typedef typename Acc<T>::type A;
  A acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = (A)0;

  const int taps = k * k;
  for (int tap = 0; tap < taps; ++tap) {
    const int ky = tap / k, kx = tap - ky * k;
    const int iy = aoy * stride - pad + ky, ix = aox * stride - pad + kx;
    const bool pvalid = mvalid && iy >= 0 && iy < H && ix >= 0 && ix < W;
    const T* xp = x + ((long long)(an * H + iy) * W + ix) * x_ld;
    const T* wp = w + (long long)tap * Cin * Cout;
    for (int c0 = 0; c0 < Cin; c0 += BK) {
      // stage A (transposed: As[k][m])
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int c = c0 + lc + j;
        As[lc + j][lp] = (pvalid && c < Cin) ? to_f(xp[c]) : 0.f;
      }
      // stage B
      {
        int c = c0 + bk;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int n = n0 + bn + j;
          Bs[bk][bn + j] = (c < Cin && n < Cout) ? to_f(wp[(long long)c * Cout + n]) : 0.f;
        }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma((A)av[i], (A)bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    int n_img = (int)(m / ((long long)Ho * Wo));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= Cout) continue;
      if (PR) {
        double v = (double)acc[i][j] + (bias ? (double)bias[n] : 0.0);
        v = apply_act_f64(v, act);
        if (chan_scale) v *= (double)chan_scale[(long long)n_img * Cout + n];
        if (res) v += (double)to_f(res[m * res_ld + n]);
        y[m * y_ld + n] = from_f<T>((float)v);
      } else {
        float v = (float)acc[i][j] + (bias ? bias[n] : 0.f);
        v = apply_act<PR>(v, act);
        if (chan_scale) v *= chan_scale[(long long)n_img * Cout + n];
        if (res) v += to_f(res[m * res_ld + n]);
        y[m * y_ld + n] = from_f<T>(v);
      }
    }
  }
}

}  // namespace

extern "C" int lpc_conv2d_direct(int dtype, const void* x, int x_ld, int B, int H, int W, int Cin,
                                 const void* w, const float* bias, int k, int stride, int pad, int Cout,
                                 void* y, int y_ld, int act, const float* chan_scale,
                                 const void* res, int res_ld, void* stream) {
  LPC_REQUIRE(x && w && y, "conv2d_direct: null pointer");
  LPC_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv2d_direct: bad shape");
  LPC_REQUIRE(k >= 1 && k <= 7 && stride >= 1 && pad >= 0, "conv2d_direct: bad k/stride/pad");
  LPC_REQUIRE(x_ld >= Cin && y_ld >= Cout, "conv2d_direct: pitch smaller than channel count");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  LPC_REQUIRE(Ho > 0 && Wo > 0, "conv2d_direct: empty output");
  const long long M = (long long)B * Ho * Wo;
  dim3 grid(cdiv(M, BM), cdiv(Cout, BN));
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == LPC_F32) {
    lpc_launch_pdl(conv_direct_kernel<float>, grid, NT, 0, s, (const float*)x, x_ld, B, H, W, Cin, (const float*)w, bias, k,
                                                  stride, pad, Cout, Ho, Wo, (float*)y, y_ld, act, chan_scale,
                                                  (const float*)res, res_ld);
  } else if (dtype == LPC_BF16) {
    lpc_launch_pdl(conv_direct_kernel<bf16>, grid, NT, 0, s, (const bf16*)x, x_ld, B, H, W, Cin, (const bf16*)w, bias, k,
                                                 stride, pad, Cout, Ho, Wo, (bf16*)y, y_ld, act, chan_scale,
                                                 (const bf16*)res, res_ld);
  } else {
    LPC_FAIL(LPC_E_ARG, "conv2d_direct: unknown dtype %d", dtype);
  }
  LPC_CHECK_LAUNCH("conv2d_direct");
  return LPC_OK;
}
