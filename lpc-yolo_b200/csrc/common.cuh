// common.cuh - shared device/host helpers for liblpcyolo (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <functional>
#include <tuple>
#include <utility>

#include "../../include/lpcyolo.h"

typedef __nv_bfloat16 bf16;

// ---- error plumbing -------------------------------------------------------------------------------
void lpc_set_error(const char* fmt, ...);
#define LPC_FAIL(code, ...)      \
  do {                           \
    lpc_set_error(__VA_ARGS__);  \
    return (code);               \
  } while (0)
#define LPC_REQUIRE(cond, ...)                  \
  do {                                          \
    if (!(cond)) LPC_FAIL(LPC_E_ARG, __VA_ARGS__); \
  } while (0)
void lpc_count_launch();
#define LPC_CHECK_LAUNCH(name)                                                            \
  do {                                                                                    \
    lpc_count_launch();                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess) LPC_FAIL(LPC_E_CUDA, "%s: %s", name, cudaGetErrorString(e__)); \
  } while (0)

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// Kernels launched through lpc_launch_pdl may start while the previous kernel of the stream is still draining: their
// prologue (barrier init, TMEM allocation, weight / bias staging - nothing the previous kernel produced) overlaps its
// tail; they call pdl_wait() before touching activations and pdl_trigger() once their own dependents may be scheduled.
// LPC_PDL=0 in the environment turns the attribute off (plain stream order).
bool lpc_pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t lpc_launch_raw(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = lpc_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
// ---- plan recording (api.cu: lpc_plan_begin / lpc_plan_end / lpc_plan_run) -----------------------------------------
// Every kernel of the library is launched through lpc_launch_pdl.  While the calling thread records a plan, each launch is
// ALSO stored - kernel pointer, grid, block, shared memory and a by-value copy of the kernel arguments (parameter structs and
// tensor maps included) - so that the whole launch sequence can be re-issued later from C, on any stream, without the host
// code that produced it.
bool lpc_plan_recording();
void lpc_plan_push(cudaStream_t recorded_on, std::function<cudaError_t(cudaStream_t)> op);
template <typename... KArgs, typename... Args>
static inline cudaError_t lpc_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  if (lpc_plan_recording()) {
    std::tuple<std::decay_t<KArgs>...> saved(args...);
    lpc_plan_push(stream, [kern, grid, block, smem, saved](cudaStream_t s) {
      return std::apply([&](const auto&... a) { return lpc_launch_raw(kern, grid, block, smem, s, a...); }, saved);
    });
  }
  return lpc_launch_raw(kern, grid, block, smem, stream, std::forward<Args>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// One-time per-DEVICE work (cudaFuncSetAttribute is a per-context setting: a process that predicts on cuda:0 and later on
// cuda:1 must raise the shared-memory limit of every kernel on both).  ``mask`` is a function-local static bit set.
static inline bool lpc_first_on_device(unsigned long long* mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (*mask & bit) return false;
  *mask |= bit;
  return true;
}
static inline int lpc_num_sms() {
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int& v = n[dev & 63];
  if (!v) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
  }
  return v;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- scalar conversions ---------------------------------------------------------------------------
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// ---- 16-byte vectors: 8 bf16 or 4 float --------------------------------------------------------------
template <typename T> struct Vec;
template <> struct Vec<bf16> {
  static constexpr int N = 8;
  uint4 raw;
  __device__ __forceinline__ void unpack(float* f) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  __device__ __forceinline__ void pack(const float* f) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  }
};
template <> struct Vec<float> {
  static constexpr int N = 4;
  uint4 raw;
  __device__ __forceinline__ void unpack(float* f) const {
    const float* p = reinterpret_cast<const float*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) f[i] = p[i];
  }
  __device__ __forceinline__ void pack(const float* f) {
    float* p = reinterpret_cast<float*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = f[i];
  }
};
template <typename T> __device__ __forceinline__ Vec<T> ldg_vec(const T* p) {
  Vec<T> v;
  v.raw = __ldg(reinterpret_cast<const uint4*>(p));
  return v;
}
template <typename T> __device__ __forceinline__ Vec<T> ld_vec(const T* p) {
  Vec<T> v;
  v.raw = *reinterpret_cast<const uint4*>(p);
  return v;
}
template <typename T> __device__ __forceinline__ void st_vec(T* p, const Vec<T>& v) {
  *reinterpret_cast<uint4*>(p) = v.raw;
}

// ---- activations -------------------------------------------------------------------------------------
// PRECISE=true (fp32 validation mode): expf + IEEE division.  PRECISE=false (bf16 mode): ex2.approx /
// rcp.approx, ~2 ulp, far below bf16 rounding.
template <bool PRECISE> __device__ __forceinline__ float exp_(float x) { return PRECISE ? expf(x) : __expf(x); }
template <bool PRECISE> __device__ __forceinline__ float div_(float a, float b) { return PRECISE ? a / b : __fdividef(a, b); }

template <bool PRECISE> __device__ __forceinline__ float sigmoid_(float x) {
  return div_<PRECISE>(1.0f, 1.0f + exp_<PRECISE>(-x));
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool PRECISE> __device__ __forceinline__ float silu_(float x) {
  if (PRECISE) return x / (1.0f + expf(-x));
  // x*sigmoid(x) = 0.5x(1 + tanh(x/2)): ONE MUFU op (tanh.approx, ~2^-11 rel. error, below bf16 rounding)
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx(h), h);
}
// mish(x) = x*tanh(softplus(x)) = x*n/(n+2) = x - 2x/(n+2) with n = e^x (e^x + 2)   (one exp, one reciprocal)
template <bool PRECISE> __device__ __forceinline__ float mish_(float x) {
  if (PRECISE) {
    if (x > 20.0f) return x;
    const float e = expf(x);
    const float n = e * (e + 2.0f);
    return x * (n / (n + 2.0f));
  }
  // 7 instructions, 2 of them MUFU (the epilogue of the small-channel convs is MUFU / issue bound, profiles/r01_g_*).
  // No clamp needed: e = +inf gives d = +inf, rcp = 0 and the result is x; x -> -inf gives e = 0, d = 2, result -0.
  const float e = ex2_approx(x * 1.4426950408889634f);
  const float d = fmaf(e, e + 2.0f, 2.0f);            // n + 2
  return x * fmaf(-2.0f, rcp_approx(d), 1.0f);        // x (1 - 2/(n+2)) = x n/(n+2)
}
// Packed (two elements per instruction: FFMA2 / FMUL2 / FADD2, sm_100) bf16-mode activations for the conv epilogues,
// which are issue-bound on the small-channel layers (profiles/r01_j_ncu_full_conv_16_32.md).
__device__ __forceinline__ float2 mish2_(float2 x) {
#ifdef LPC_MISH_ONE_MUFU
  const float2 xl = __fmul2_rn(x, make_float2(1.4426950408889634f, 1.4426950408889634f));
  float2 e;
  e.x = ex2_approx(fminf(xl.x, 30.0f));      // clamp: d = e^2 + 2e + 2 stays finite; mish(x) = x beyond
  e.y = ex2_approx(fminf(xl.y, 30.0f));
  const float2 two = make_float2(2.0f, 2.0f);
  const float2 t = __fadd2_rn(e, two);
  const float2 n = __fmul2_rn(e, t);
  const float2 d = __ffma2_rn(e, t, two);
  float2 r;
  r.x = __uint_as_float(0x7EF311C7u - __float_as_uint(d.x));
  r.y = __uint_as_float(0x7EF311C7u - __float_as_uint(d.y));
  const float2 nd = make_float2(-d.x, -d.y);
  r = __fmul2_rn(r, __ffma2_rn(nd, r, two));
  r = __fmul2_rn(r, __ffma2_rn(nd, r, two));
  return __fmul2_rn(__fmul2_rn(x, n), r);
#else
  // 9 instructions per PAIR (2 MUFU per element).  The conv epilogues are bound by instruction issue, not by the MUFU
  // pipe (ncu: issue 68 %, xu 31 %), so this beats the 19-instruction one-MUFU variant above (integer-seeded Newton
  // reciprocal on the FMA pipe, 6.6e-6 relative error) that is kept under LPC_MISH_ONE_MUFU.
  // x (1 - 2/(n+2)) with n + 2 = e (e + 2) + 2; no clamp: e = +inf gives rcp = 0 and the result x.
  const float2 xl = __fmul2_rn(x, make_float2(1.4426950408889634f, 1.4426950408889634f));
  float2 e;
  e.x = ex2_approx(xl.x);
  e.y = ex2_approx(xl.y);
  const float2 two = make_float2(2.0f, 2.0f);
  const float2 d = __ffma2_rn(e, __fadd2_rn(e, two), two);
  float2 r;
  r.x = rcp_approx(d.x);
  r.y = rcp_approx(d.y);
  return __fmul2_rn(x, __ffma2_rn(r, make_float2(-2.0f, -2.0f), make_float2(1.0f, 1.0f)));
#endif
}
__device__ __forceinline__ float2 silu2_(float2 x) {
  const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  float2 th;
  th.x = tanh_approx(h.x);
  th.y = tanh_approx(h.y);
  return __ffma2_rn(h, th, h);
}
template <bool PRECISE> __device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case LPC_ACT_SILU: return silu_<PRECISE>(v);
    case LPC_ACT_MISH: return mish_<PRECISE>(v);
    case LPC_ACT_SIGMOID: return sigmoid_<PRECISE>(v);
    case LPC_ACT_RELU: return fmaxf(v, 0.0f);
    default: return v;
  }
}
// fp64 activations of the fp32 validation mode (one rounding per layer output, see conv_direct.cu)
__device__ __forceinline__ double apply_act_f64(double v, int act) {
  switch (act) {
    case LPC_ACT_SILU: return v / (1.0 + exp(-v));
    case LPC_ACT_MISH: {
      if (v > 40.0) return v;
      const double e = exp(v);
      const double n = e * (e + 2.0);
      return v * (n / (n + 2.0));
    }
    case LPC_ACT_SIGMOID: return 1.0 / (1.0 + exp(-v));
    case LPC_ACT_RELU: return v > 0.0 ? v : 0.0;
    default: return v;
  }
}
template <typename T> struct Precise { static constexpr bool value = false; };
template <> struct Precise<float> { static constexpr bool value = true; };
