// stem.cu - the Cin=3 stem convolution (YAML layer 0: Conv(3, c, 3, 2), conv.py:36-54), K4 in SURVEY.md 2.2.
//
// K = 27 is far too small for the tensor-core path and the layer is bandwidth / FMA-issue bound, so this is a
// register-tiled CUDA-core kernel: a CTA stages the input patch of a 8 x 64 output tile in shared memory
// (NHWC, 4 channels per pixel = one 8-byte bf16x4 / 16-byte fp32x4 vector), every thread owns TWO horizontally
// adjacent output pixels (their 3 x 5 x 3 input window lives in registers) and loops over groups of 8 output
// channels whose 27 x 8 weights are broadcast 16-byte shared-memory reads.  432 FMA per 54 LDS.128.
#include "common.cuh"

namespace {

constexpr int ST_TH = 8, ST_TW = 64, ST_NT = 256;  // 8 x 64 outputs per CTA, 2 per thread
constexpr int ST_MAXC = 96;

template <typename T> struct In4;
template <> struct In4<bf16> {
  typedef uint2 V;
  static __device__ __forceinline__ void load(const V& v, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
    float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    f[0] = a.x; f[1] = a.y; f[2] = b.x;
  }
};
template <> struct In4<float> {
  typedef float4 V;
  static __device__ __forceinline__ void load(const V& v, float* f) { f[0] = v.x; f[1] = v.y; f[2] = v.z; }
};

template <typename T, int S>
__global__ void __launch_bounds__(ST_NT)
stem_conv_kernel(const T* __restrict__ x, int B, int H, int W, const float* __restrict__ w, const float* __restrict__ bias,
                 int Cout, int Ho, int Wo, T* __restrict__ y, int y_ld, int act) {
  pdl_trigger();
  pdl_wait();
  typedef typename In4<T>::V V;
  constexpr bool PR = Precise<T>::value;
  constexpr int IN_H = (ST_TH - 1) * S + 3, IN_W = (ST_TW - 1) * S + 3;
  __shared__ V tile[IN_H][IN_W + 1];
  __shared__ __align__(16) float ws[27 * ST_MAXC];
  __shared__ float bs[ST_MAXC];

  const int tid = threadIdx.x;
  const int tiles_x = (Wo + ST_TW - 1) / ST_TW;
  const int n = blockIdx.z;
  const int oy0 = blockIdx.y * ST_TH, ox0 = (blockIdx.x % tiles_x) * ST_TW;
  const int iy0 = oy0 * S - 1, ix0 = ox0 * S - 1;

  for (int i = tid; i < 27 * Cout; i += ST_NT) ws[i] = w[i];
  for (int i = tid; i < Cout; i += ST_NT) bs[i] = bias ? bias[i] : 0.f;
  const V* xin = reinterpret_cast<const V*>(x) + (long long)n * H * W;
  V zero;
  memset(&zero, 0, sizeof(V));
  for (int i = tid; i < IN_H * IN_W; i += ST_NT) {
    const int r = i / IN_W, c = i - r * IN_W;
    const int iy = iy0 + r, ix = ix0 + c;
    tile[r][c] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(xin + (long long)iy * W + ix) : zero;
  }
  __syncthreads();

  // thread -> output row ty, output columns 2*tp, 2*tp+1
  const int ty = tid >> 5, tp = tid & 31;
  constexpr int WIN_W = S + 3;  // input columns spanned by two adjacent outputs
  float in[3][WIN_W][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < WIN_W; ++c) In4<T>::load(tile[ty * S + r][tp * 2 * S + c], in[r][c]);

  const int oy = oy0 + ty, ox = ox0 + tp * 2;
  if (oy >= Ho || ox >= Wo) return;
  const bool second = ox + 1 < Wo;
  T* yo = y + ((long long)(n * Ho + oy) * Wo + ox) * y_ld;
  for (int g = 0; g < Cout; g += 8) {
    float a0[8], a1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a0[j] = a1[j] = bs[g + j];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          const float* wp = &ws[((ky * 3 + kx) * 3 + ci) * Cout + g];
          const float4 w0 = *reinterpret_cast<const float4*>(wp), w1 = *reinterpret_cast<const float4*>(wp + 4);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          const float v0 = in[ky][kx][ci], v1 = in[ky][kx + S][ci];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            a0[j] = fmaf(v0, wv[j], a0[j]);
            a1[j] = fmaf(v1, wv[j], a1[j]);
          }
        }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a0[j] = apply_act<PR>(a0[j], act);
      a1[j] = apply_act<PR>(a1[j], act);
    }
    if (sizeof(T) == 2) {
      Vec<bf16> o;
      o.pack(a0);
      st_vec<bf16>(reinterpret_cast<bf16*>(yo) + g, o);
      if (second) {
        o.pack(a1);
        st_vec<bf16>(reinterpret_cast<bf16*>(yo) + y_ld + g, o);
      }
    } else {
      Vec<float> o;
      float* yf = reinterpret_cast<float*>(yo);
      o.pack(a0); st_vec<float>(yf + g, o);
      o.pack(a0 + 4); st_vec<float>(yf + g + 4, o);
      if (second) {
        o.pack(a1); st_vec<float>(yf + y_ld + g, o);
        o.pack(a1 + 4); st_vec<float>(yf + y_ld + g + 4, o);
      }
    }
  }
}

}  // namespace

int lpc_stem_conv_tc(const void* x, int B, int H, int W, const float* w, const float* bias, int Cout, void* y, int y_ld,
                     int act, cudaStream_t stream);   // stem_tc.cu

extern "C" int lpc_stem_conv(int dtype, const void* x, int B, int H, int W, const float* w, const float* bias, int stride,
                             int Cout, void* y, int y_ld, int act, void* stream) {
  LPC_REQUIRE(x && w && y, "stem_conv: null pointer");
  LPC_REQUIRE(B > 0 && H > 0 && W > 0, "stem_conv: bad shape");
  LPC_REQUIRE(stride == 1 || stride == 2, "stem_conv: stride must be 1 or 2");
  LPC_REQUIRE(Cout % 8 == 0 && Cout <= ST_MAXC && y_ld % 8 == 0 && y_ld >= Cout, "stem_conv: Cout must be a multiple of 8, <= %d", ST_MAXC);
  LPC_REQUIRE(aligned16(x) && aligned16(y) && aligned16(w), "stem_conv: pointers must be 16-byte aligned");
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  dim3 grid(cdiv(Wo, ST_TW), cdiv(Ho, ST_TH), B);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == LPC_BF16 && stride == 2) {   // tensor-core path; falls through for shapes it does not take
    const int r = lpc_stem_conv_tc(x, B, H, W, w, bias, Cout, y, y_ld, act, s);
    if (r != LPC_E_UNSUPPORTED) return r;
  }
#define STEM(T_, S_) lpc_launch_pdl(stem_conv_kernel<T_, S_>, grid, ST_NT, 0, s, (const T_*)x, B, H, W, w, bias, Cout, Ho, Wo, (T_*)y, y_ld, act)
  if (dtype == LPC_BF16) { if (stride == 2) STEM(bf16, 2); else STEM(bf16, 1); }
  else if (dtype == LPC_F32) { if (stride == 2) STEM(float, 2); else STEM(float, 1); }
  else LPC_FAIL(LPC_E_ARG, "stem_conv: unknown dtype %d", dtype);
#undef STEM
  LPC_CHECK_LAUNCH("stem_conv");
  return LPC_OK;
}
