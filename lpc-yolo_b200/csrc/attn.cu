// attn.cu - PSA attention core: fused QK^T -> softmax -> PV with an online softmax (never materialises
// the [N,N] matrix the reference builds, block.py:789-793).  fp32 math on CUDA cores; this block is
// <= 0.4 % of the network's FLOPs and N <= 1600 tokens.
//
// CTA = 64 queries of one (image, head); 256 threads as a 16x16 grid of 4x4 score tiles.
// Loop over 64-key tiles: S = scale * Q K^T, running row max / sum, P -> smem, O += P V.
#include "common.cuh"

namespace {

constexpr int BQ = 64, BKEY = 64, ATT_NT = 256;
constexpr int MAX_KD = 40, MAX_HD = 80;  // yolov10m: kd 36, hd 72
constexpr int HD_IT = MAX_HD / 16;       // output columns per thread: tx, tx+16, ...

template <typename T>
__global__ void __launch_bounds__(ATT_NT)
psa_attention_kernel(const T* __restrict__ qkv, int ld, int N, int heads, int kd, int hd, T* __restrict__ out, int out_ld) {
  constexpr bool PR = Precise<T>::value;
  extern __shared__ float smem[];
  float* Qs = smem;                       // [kd][BQ+4]   (transposed: Qs[c][i])
  float* Ks = Qs + MAX_KD * (BQ + 4);     // [kd][BKEY+4] (Ks[c][j])
  float* Vs = Ks + MAX_KD * (BKEY + 4);   // [BKEY][MAX_HD]
  float* Ps = Vs + BKEY * MAX_HD;         // [BQ][BKEY+1]

  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const float scale = rsqrtf((float)kd);
  const int q_off = h * kd, k_off = heads * kd + h * kd, v_off = 2 * heads * kd + h * hd;
  const T* base = qkv + (long long)b * N * ld;

  for (int e = tid; e < BQ * kd; e += ATT_NT) {
    int i = e / kd, c = e - i * kd;
    Qs[c * (BQ + 4) + i] = (q0 + i < N) ? to_f(base[(long long)(q0 + i) * ld + q_off + c]) * scale : 0.f;
  }

  float m_run[4], l_run[4], o[4][HD_IT];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -INFINITY;
    l_run[i] = 0.f;
#pragma unroll
    for (int j = 0; j < HD_IT; ++j) o[i][j] = 0.f;
  }

  for (int j0 = 0; j0 < N; j0 += BKEY) {
    __syncthreads();  // previous tile's Ps / Vs fully consumed (and Qs visible on the first pass)
    for (int e = tid; e < BKEY * kd; e += ATT_NT) {
      int j = e / kd, c = e - j * kd;
      Ks[c * (BKEY + 4) + j] = (j0 + j < N) ? to_f(base[(long long)(j0 + j) * ld + k_off + c]) : 0.f;
    }
    for (int e = tid; e < BKEY * hd; e += ATT_NT) {
      int j = e / hd, d = e - j * hd;
      Vs[j * MAX_HD + d] = (j0 + j < N) ? to_f(base[(long long)(j0 + j) * ld + v_off + d]) : 0.f;
    }
    __syncthreads();

    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
    for (int c = 0; c < kd; ++c) {
      float4 a = *reinterpret_cast<const float4*>(&Qs[c * (BQ + 4) + ty * 4]);
      float4 k4 = *reinterpret_cast<const float4*>(&Ks[c * (BKEY + 4) + tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, kv[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(av[i], kv[j], s[i][j]);
    }
    // mask keys beyond N, online softmax per row (16 threads tx=0..15 share a row group)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j0 + tx * 4 + j >= N) s[i][j] = -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int off = 8; off; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m_run[i], mx);
      const float corr = (m_run[i] == -INFINITY) ? 0.f : exp_<PR>(m_run[i] - m_new);
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float p = (s[i][j] == -INFINITY) ? 0.f : exp_<PR>(s[i][j] - m_new);
        Ps[(ty * 4 + i) * (BKEY + 1) + tx * 4 + j] = p;
        rs += p;
      }
#pragma unroll
      for (int off = 8; off; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      l_run[i] = l_run[i] * corr + rs;
      m_run[i] = m_new;
#pragma unroll
      for (int j = 0; j < HD_IT; ++j) o[i][j] *= corr;
    }
    __syncthreads();
    for (int j = 0; j < BKEY; ++j) {
      float pv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pv[i] = Ps[(ty * 4 + i) * (BKEY + 1) + j];
#pragma unroll
      for (int d = 0; d < HD_IT; ++d) {
        const int col = tx + 16 * d;
        if (col < hd) {
          const float v = Vs[j * MAX_HD + col];
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i][d] = fmaf(pv[i], v, o[i][d]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty * 4 + i;
    if (q >= N) continue;
    const float inv = 1.0f / l_run[i];
#pragma unroll
    for (int d = 0; d < HD_IT; ++d) {
      const int col = tx + 16 * d;
      if (col < hd) out[((long long)b * N + q) * out_ld + h * hd + col] = from_f<T>(o[i][d] * inv);
    }
  }
}

constexpr size_t ATT_SMEM = sizeof(float) * (MAX_KD * (BQ + 4) + MAX_KD * (BKEY + 4) + BKEY * MAX_HD + BQ * (BKEY + 1));

}  // namespace

extern "C" int lpc_psa_attention(int dtype, const void* qkv, int qkv_ld, int B, int N, int heads, int kd, int hd,
                                 void* out, int out_ld, void* stream) {
  LPC_REQUIRE(qkv && out && B > 0 && N > 0 && heads > 0, "psa_attention: bad argument");
  LPC_REQUIRE(kd > 0 && kd <= MAX_KD && hd > 0 && hd <= MAX_HD, "psa_attention: kd <= %d, hd <= %d", MAX_KD, MAX_HD);
  LPC_REQUIRE(qkv_ld >= heads * (2 * kd + hd) && out_ld >= heads * hd, "psa_attention: pitch too small");
  dim3 grid(cdiv(N, BQ), heads, B);
  cudaStream_t s = (cudaStream_t)stream;
  static bool attr_done[2] = {false, false};
  if (dtype == LPC_F32) {
    if (!attr_done[0]) { cudaFuncSetAttribute(psa_attention_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM); attr_done[0] = true; }
    psa_attention_kernel<float><<<grid, ATT_NT, ATT_SMEM, s>>>((const float*)qkv, qkv_ld, N, heads, kd, hd, (float*)out, out_ld);
  } else if (dtype == LPC_BF16) {
    if (!attr_done[1]) { cudaFuncSetAttribute(psa_attention_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM); attr_done[1] = true; }
    psa_attention_kernel<bf16><<<grid, ATT_NT, ATT_SMEM, s>>>((const bf16*)qkv, qkv_ld, N, heads, kd, hd, (bf16*)out, out_ld);
  } else {
    LPC_FAIL(LPC_E_ARG, "psa_attention: unknown dtype %d", dtype);
  }
  LPC_CHECK_LAUNCH("psa_attention");
  return LPC_OK;
}
