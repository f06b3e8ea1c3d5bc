// attn.cu - PSA attention core: fused QK^T -> softmax -> PV with an online softmax (never materialises
// the [N,N] matrix the reference builds, block.py:789-793).  fp32 math on CUDA cores; this block is
// <= 0.4 % of the network's FLOPs and N <= 1600 tokens.
//
// CTA = 64 queries of one (image, head); 256 threads as a 16x16 grid of 4x4 score tiles.
// Loop over 64-key tiles: S = scale * Q K^T, running row max / sum, P -> smem, O += P V.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BQ = 64, BKEY = 64, ATT_NT = 256;
constexpr int MAX_KD = 40, MAX_HD = 80;  // yolov10m: kd 36, hd 72
constexpr int HD_IT = MAX_HD / 16;       // output columns per thread: tx, tx+16, ...

template <typename T>
__global__ void __launch_bounds__(ATT_NT)
psa_attention_kernel(const T* __restrict__ qkv, int ld, int N, int heads, int kd, int hd, T* __restrict__ out, int out_ld) {
  pdl_trigger();
  pdl_wait();
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

// ---- fp32 validation mode --------------------------------------------------------------------------------
// One warp per (image, head, query); scores, softmax and the PV sum are carried in fp64 and the output is rounded to
// fp32 once (same contract as conv_direct_kernel<float>: a layer adds only the storage rounding of its output).
constexpr int VAL_WARPS = 4;
__global__ void __launch_bounds__(VAL_WARPS * 32)
psa_attention_f32_validate_kernel(const float* __restrict__ qkv, int ld, int N, int heads, int kd, int hd, float* __restrict__ out, int out_ld) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ double val_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * VAL_WARPS + warp, h = blockIdx.y, b = blockIdx.z;
  if (q >= N) return;
  double* sc = val_smem + (size_t)warp * N;
  const float* base = qkv + (long long)b * N * ld;
  const int q_off = h * kd, k_off = heads * kd + h * kd, v_off = 2 * heads * kd + h * hd;
  const double scale = 1.0 / sqrt((double)kd);
  const float* qrow = base + (long long)q * ld + q_off;
  double mx = -INFINITY;
  for (int j = lane; j < N; j += 32) {
    const float* krow = base + (long long)j * ld + k_off;
    double a = 0.0;
    for (int c = 0; c < kd; ++c) a = fma((double)qrow[c], (double)krow[c], a);
    a *= scale;
    sc[j] = a;
    mx = fmax(mx, a);
  }
  for (int off = 16; off; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  double sum = 0.0;
  for (int j = lane; j < N; j += 32) {
    const double p = exp(sc[j] - mx);
    sc[j] = p;
    sum += p;
  }
  for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  __syncwarp();
  for (int d = lane; d < hd; d += 32) {
    double o = 0.0;
    for (int j = 0; j < N; ++j) o = fma(sc[j], (double)base[(long long)j * ld + v_off + d], o);
    out[((long long)b * N + q) * out_ld + h * hd + d] = (float)(o / sum);
  }
}

// ---- bf16 tensor-core variant ------------------------------------------------------------------------------
// Flash-attention style: CTA = 64 queries of one (image, head), 4 warps x 16 query rows; keys/values stream through a
// double-buffered cp.async ring in 64-key tiles; S = Q K^T and O += P V are mma.sync m16n8k16 (bf16 in, fp32
// accumulate); the softmax runs on the accumulator fragments (quad shuffles), P is re-packed to bf16 in registers as
// the A operand of the second MMA (never touches shared memory).  This block is <= 0.4 % of the network's FLOPs, but
// at fp32-FMA speed it was 7-9 % of the step (profiles/r01_f_*): the point of the tensor path is latency, not peak.
__host__ __device__ constexpr int pitch_of(int dim) { return ((dim / 8) & 1) ? dim : dim + 8; }   // odd number of 16-byte units per row -> conflict-free ldmatrix

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <int BYTES> __device__ __forceinline__ void cp_async_zfill(uint32_t dst, const void* src, bool valid) {
  const uint32_t n = valid ? BYTES : 0;
  if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}

constexpr int MQ = 64, MK = 64, MMA_NT = 128;

template <int KD, int HD>
__global__ void __launch_bounds__(MMA_NT)
psa_attention_mma_kernel(const bf16* __restrict__ qkv, int ld, int N, int heads, bf16* __restrict__ out, int out_ld) {
  pdl_trigger();
  pdl_wait();
  constexpr int KDP = (KD + 15) / 16 * 16, KP = pitch_of(KDP), VP = pitch_of(HD);
  constexpr int CH = (KD % 8 == 0) ? 16 : 8;            // global->shared copy granularity in bytes
  constexpr int KCH = KD * 2 / CH, VCH = HD * 2 / 16;   // chunks per row
  constexpr int KSTEPS = KDP / 16, NT_S = MK / 8, NT_O = HD / 8;
  __shared__ __align__(16) bf16 Qs[MQ * KP];
  __shared__ __align__(16) bf16 Ks[2][MK * KP];
  __shared__ __align__(16) bf16 Vs[2][MK * VP];

  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * MQ;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q_off = h * KD, k_off = heads * KD + h * KD, v_off = 2 * heads * KD + h * HD;
  const bf16* base = qkv + (long long)b * N * ld;

  // zero the K-padding columns (KD..KDP) of Q and both K stages once; the copies below never touch them
  if (KDP > KD) {
    for (int e = tid; e < MQ * (KDP - KD); e += MMA_NT) {
      const int r = e / (KDP - KD), c = KD + e % (KDP - KD);
      Qs[r * KP + c] = __float2bfloat16(0.f);
      Ks[0][r * KP + c] = __float2bfloat16(0.f);
      Ks[1][r * KP + c] = __float2bfloat16(0.f);
    }
  }
  auto load_kv = [&](int stage, int j0) {
    for (int e = tid; e < MK * KCH; e += MMA_NT) {
      const int r = e / KCH, c = e - r * KCH;
      const bool ok = j0 + r < N;
      cp_async_zfill<CH>((uint32_t)__cvta_generic_to_shared(&Ks[stage][r * KP]) + c * CH,
                         reinterpret_cast<const char*>(base + (long long)(ok ? j0 + r : 0) * ld + k_off) + c * CH, ok);
    }
    for (int e = tid; e < MK * VCH; e += MMA_NT) {
      const int r = e / VCH, c = e - r * VCH;
      const bool ok = j0 + r < N;
      cp_async_zfill<16>((uint32_t)__cvta_generic_to_shared(&Vs[stage][r * VP]) + c * 16,
                         reinterpret_cast<const char*>(base + (long long)(ok ? j0 + r : 0) * ld + v_off) + c * 16, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int e = tid; e < MQ * KCH; e += MMA_NT) {
    const int r = e / KCH, c = e - r * KCH;
    const bool ok = q0 + r < N;
    cp_async_zfill<CH>((uint32_t)__cvta_generic_to_shared(&Qs[r * KP]) + c * CH,
                       reinterpret_cast<const char*>(base + (long long)(ok ? q0 + r : 0) * ld + q_off) + c * CH, ok);
  }
  load_kv(0, 0);

  const float sl2 = rsqrtf((float)KD) * 1.4426950408889634f;   // softmax scale folded into the exp2 argument
  float o[NT_O][4];
#pragma unroll
  for (int i = 0; i < NT_O; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  uint32_t qf[KSTEPS][4];

  const int ntiles = (N + MK - 1) / MK;
  for (int t = 0; t < ntiles; ++t) {
    const int st = t & 1;
    if (t + 1 < ntiles) {
      load_kv(st ^ 1, (t + 1) * MK);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (t == 0) {
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, col = ks * 16 + (lane >> 4) * 8;
        ldsm_x4((uint32_t)__cvta_generic_to_shared(&Qs[row * KP + col]), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    // ---- S = Q K^T : 16 x 64 per warp ----
    float s[NT_S][4];
#pragma unroll
    for (int nt = 0; nt < NT_S; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      const uint32_t krow = (uint32_t)__cvta_generic_to_shared(&Ks[st][(nt * 8 + (lane & 7)) * KP]);
#pragma unroll
      for (int ks = 0; ks + 1 < KSTEPS; ks += 2) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(krow + (uint32_t)(ks * 16 + (lane >> 3) * 8) * 2u, b0, b1, b2, b3);
        mma_bf16(s[nt], qf[ks], b0, b1);
        mma_bf16(s[nt], qf[ks + 1], b2, b3);
      }
      if (KSTEPS & 1) {
        uint32_t b0, b1;
        ldsm_x2(krow + (uint32_t)((KSTEPS - 1) * 16 + ((lane >> 3) & 1) * 8) * 2u, b0, b1);
        mma_bf16(s[nt], qf[KSTEPS - 1], b0, b1);
      }
    }
    // ---- online softmax on the fragments (rows g = lane/4 and g + 8; a quad shares a row) ----
    const int jbase = t * MK + 2 * (lane & 3);
    if (t == ntiles - 1) {
#pragma unroll
      for (int nt = 0; nt < NT_S; ++nt) {
        if (jbase + nt * 8 >= N) s[nt][0] = s[nt][2] = -INFINITY;
        if (jbase + nt * 8 + 1 >= N) s[nt][1] = s[nt][3] = -INFINITY;
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < NT_S; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
    float corr[2], msc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);                  // finite: every tile holds at least one valid key
      corr[r] = ex2_approx((m_run[r] - m_new) * sl2);              // first tile: exp2(-inf) = 0
      m_run[r] = m_new;
      msc[r] = m_new * sl2;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[NT_S / 2][4];
#pragma unroll
    for (int nt = 0; nt < NT_S; ++nt) {
      const float p0 = ex2_approx(fmaf(s[nt][0], sl2, -msc[0])), p1 = ex2_approx(fmaf(s[nt][1], sl2, -msc[0]));
      const float p2 = ex2_approx(fmaf(s[nt][2], sl2, -msc[1])), p3 = ex2_approx(fmaf(s[nt][3], sl2, -msc[1]));
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];   // per-thread partial sums; reduced over the quad at the end
#pragma unroll
    for (int i = 0; i < NT_O; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0];
      o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
    // ---- O += P V ----
#pragma unroll
    for (int dt = 0; dt < NT_O; ++dt) {
#pragma unroll
      for (int kk = 0; kk < MK / 32; ++kk) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t((uint32_t)__cvta_generic_to_shared(&Vs[st][(kk * 32 + lane) * VP + dt * 8]), b0, b1, b2, b3);
        mma_bf16(o[dt], pf[2 * kk], b0, b1);
        mma_bf16(o[dt], pf[2 * kk + 1], b2, b3);
      }
    }
    __syncthreads();   // this stage is refilled by the next iteration's prefetch
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const int g = lane >> 2;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int q = q0 + warp * 16 + g + r * 8;
    if (q >= N) continue;
    const float inv = 1.0f / l_run[r];
    bf16* orow = out + ((long long)b * N + q) * out_ld + h * HD + 2 * (lane & 3);
#pragma unroll
    for (int dt = 0; dt < NT_O; ++dt)
      *reinterpret_cast<uint32_t*>(orow + dt * 8) = pack_bf16x2(o[dt][r * 2] * inv, o[dt][r * 2 + 1] * inv);
  }
}

// ---- the same attention with K and V of one (image, head) RESIDENT in shared memory ---------------------------------------
// psa_attention_mma_kernel re-loads K / V once per 64-query CTA (seven CTAs per head at N = 400) and pays one cp.async round
// trip + two __syncthreads per 64-key tile: 28-30 us at LPC B = 64 for 4 us of HBM time and well under 1 us of math.  Here a
// CTA owns a (image, head) pair - or a contiguous share of its query blocks when there are few pairs -, loads all N keys and
// values once (N <= 512: <= 115 KB, two CTAs per SM at N = 400) and its eight warps walk 16-query blocks without any
// block-level synchronisation after the load: Q fragments come straight from global memory in the mma A layout, S = QK^T,
// online softmax and O += PV stay in registers exactly as in the kernel above.
constexpr int RES_WARPS = 8, RES_NT = RES_WARPS * 32;
template <int KD, int HD>
__host__ __device__ constexpr size_t psa_res_smem(int n_pad) {
  return (size_t)n_pad * (pitch_of((KD + 15) / 16 * 16) + pitch_of(HD)) * sizeof(bf16);
}

template <int KD, int HD>
__global__ void __launch_bounds__(RES_NT)
psa_attention_res_kernel(const bf16* __restrict__ qkv, int ld, int N, int heads, int n_pad, bf16* __restrict__ out, int out_ld) {
  pdl_trigger();
  pdl_wait();
  constexpr int KDP = (KD + 15) / 16 * 16, KP = pitch_of(KDP), VP = pitch_of(HD);
  constexpr int CH = (KD % 8 == 0) ? 16 : 8;
  constexpr int KCH = KD * 2 / CH, VCH = HD * 2 / 16;
  constexpr int KSTEPS = KDP / 16, NT_S = MK / 8, NT_O = HD / 8;
  extern __shared__ __align__(16) unsigned char res_smem[];
  bf16* Ks = reinterpret_cast<bf16*>(res_smem);
  bf16* Vs = Ks + (size_t)n_pad * KP;

  const int b = blockIdx.z, h = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q_off = h * KD, k_off = heads * KD + h * KD, v_off = 2 * heads * KD + h * HD;
  const bf16* base = qkv + (long long)b * N * ld;

  if (KDP > KD) {       // K padding columns, never touched by the copies
    for (int e = tid; e < n_pad * (KDP - KD); e += RES_NT) {
      const int r = e / (KDP - KD), c = KD + e % (KDP - KD);
      Ks[r * KP + c] = __float2bfloat16(0.f);
    }
  }
  for (int e = tid; e < n_pad * KCH; e += RES_NT) {
    const int r = e / KCH, c = e - r * KCH;
    const bool ok = r < N;
    cp_async_zfill<CH>((uint32_t)__cvta_generic_to_shared(&Ks[r * KP]) + c * CH,
                       reinterpret_cast<const char*>(base + (long long)(ok ? r : 0) * ld + k_off) + c * CH, ok);
  }
  for (int e = tid; e < n_pad * VCH; e += RES_NT) {
    const int r = e / VCH, c = e - r * VCH;
    const bool ok = r < N;
    cp_async_zfill<16>((uint32_t)__cvta_generic_to_shared(&Vs[r * VP]) + c * 16,
                       reinterpret_cast<const char*>(base + (long long)(ok ? r : 0) * ld + v_off) + c * 16, ok);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");

  // this CTA's share of the 16-query blocks of the head, block-cyclic over the warps
  const int nqb = (N + 15) / 16;
  const int per_cta = (nqb + (int)gridDim.x - 1) / (int)gridDim.x;
  const int qb_lo = blockIdx.x * per_cta, qb_hi = min(nqb, qb_lo + per_cta);
  const float sl2 = rsqrtf((float)KD) * 1.4426950408889634f;
  const int ntiles = n_pad / MK;
  const int g = lane >> 2, c2 = 2 * (lane & 3);

  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  for (int qb = qb_lo + warp; qb < qb_hi; qb += RES_WARPS) {
    // Q fragments (mma.m16n8k16 A layout: rows g / g + 8, columns c2 + {0, 1} and + 8) straight from global memory
    uint32_t qf[KSTEPS][4];
    const int r0 = qb * 16 + g, r1 = r0 + 8;
    const bf16* q0p = base + (long long)min(r0, N - 1) * ld + q_off;
    const bf16* q1p = base + (long long)min(r1, N - 1) * ld + q_off;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
      const int ca = ks * 16 + c2, cb = ca + 8;
      qf[ks][0] = ca < KD ? *reinterpret_cast<const uint32_t*>(q0p + ca) : 0u;
      qf[ks][1] = ca < KD ? *reinterpret_cast<const uint32_t*>(q1p + ca) : 0u;
      qf[ks][2] = cb < KD ? *reinterpret_cast<const uint32_t*>(q0p + cb) : 0u;
      qf[ks][3] = cb < KD ? *reinterpret_cast<const uint32_t*>(q1p + cb) : 0u;
    }
    float o[NT_O][4];
#pragma unroll
    for (int i = 0; i < NT_O; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

    for (int t = 0; t < ntiles; ++t) {
      const bf16* Kt = Ks + (size_t)t * MK * KP;
      const bf16* Vt = Vs + (size_t)t * MK * VP;
      float s[NT_S][4];
#pragma unroll
      for (int nt = 0; nt < NT_S; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        const uint32_t krow = (uint32_t)__cvta_generic_to_shared(&Kt[(nt * 8 + (lane & 7)) * KP]);
#pragma unroll
        for (int ks = 0; ks + 1 < KSTEPS; ks += 2) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4(krow + (uint32_t)(ks * 16 + (lane >> 3) * 8) * 2u, b0, b1, b2, b3);
          mma_bf16(s[nt], qf[ks], b0, b1);
          mma_bf16(s[nt], qf[ks + 1], b2, b3);
        }
        if (KSTEPS & 1) {
          uint32_t b0, b1;
          ldsm_x2(krow + (uint32_t)((KSTEPS - 1) * 16 + ((lane >> 3) & 1) * 8) * 2u, b0, b1);
          mma_bf16(s[nt], qf[KSTEPS - 1], b0, b1);
        }
      }
      const int jbase = t * MK + c2;
      if ((t + 1) * MK > N) {
#pragma unroll
        for (int nt = 0; nt < NT_S; ++nt) {
          if (jbase + nt * 8 >= N) s[nt][0] = s[nt][2] = -INFINITY;
          if (jbase + nt * 8 + 1 >= N) s[nt][1] = s[nt][3] = -INFINITY;
        }
      }
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < NT_S; ++nt) {
        mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
        mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
      }
      float corr[2], msc[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        const float m_new = fmaxf(m_run[r], mx[r]);
        corr[r] = ex2_approx((m_run[r] - m_new) * sl2);
        m_run[r] = m_new;
        msc[r] = m_new * sl2;
      }
      float rs[2] = {0.f, 0.f};
      uint32_t pf[NT_S / 2][4];
#pragma unroll
      for (int nt = 0; nt < NT_S; ++nt) {
        const float p0 = ex2_approx(fmaf(s[nt][0], sl2, -msc[0])), p1 = ex2_approx(fmaf(s[nt][1], sl2, -msc[0]));
        const float p2 = ex2_approx(fmaf(s[nt][2], sl2, -msc[1])), p3 = ex2_approx(fmaf(s[nt][3], sl2, -msc[1]));
        rs[0] += p0 + p1;
        rs[1] += p2 + p3;
        pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
        pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
      for (int i = 0; i < NT_O; ++i) {
        o[i][0] *= corr[0]; o[i][1] *= corr[0];
        o[i][2] *= corr[1]; o[i][3] *= corr[1];
      }
#pragma unroll
      for (int dt = 0; dt < NT_O; ++dt) {
#pragma unroll
        for (int kk = 0; kk < MK / 32; ++kk) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t((uint32_t)__cvta_generic_to_shared(&Vt[(kk * 32 + lane) * VP + dt * 8]), b0, b1, b2, b3);
          mma_bf16(o[dt], pf[2 * kk], b0, b1);
          mma_bf16(o[dt], pf[2 * kk + 1], b2, b3);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
      l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int q = qb * 16 + g + r * 8;
      if (q >= N) continue;
      const float inv = 1.0f / l_run[r];
      bf16* orow = out + ((long long)b * N + q) * out_ld + h * HD + c2;
#pragma unroll
      for (int dt = 0; dt < NT_O; ++dt)
        *reinterpret_cast<uint32_t*>(orow + dt * 8) = pack_bf16x2(o[dt][r * 2] * inv, o[dt][r * 2 + 1] * inv);
    }
  }
}

constexpr size_t ATT_SMEM = sizeof(float) * (MAX_KD * (BQ + 4) + MAX_KD * (BKEY + 4) + BKEY * MAX_HD + BQ * (BKEY + 1));

}  // namespace

int lpc_psa_attention_tc(const void* qkv, int ld, int B, int N, int heads, void* out, int out_ld, cudaStream_t stream);   // attn_tc.cu

// Which bf16 kernel serves kd 32 / hd 64.  Measured on B200 (us, tcgen05 / mma.sync, tools/run_attn.py, profiles/r02_g_attention.md):
// N 1600 x 5 heads x B 32 (yolov10x @1280) 301 / 341; N 400 x 2 x B 64 (LPC) 31.1 / 31.1; N 400 x 4 x B 256 192 / 182; batch 1:
// N 900 22.8 / 18.7, N 400 14.6 / 10.5.  The tensor-memory kernel wins where the key loop is long (its per-block MMA +
// softmax hand-offs amortise), the register-fragment kernel where a CTA lives for three or four blocks: tcgen05 from
// N > 512.  LPC_ATT_TC=1 / 0 forces one or the other (the parity tests run both on every shape).
static bool att_use_tc(int N) {
  const char* e = getenv("LPC_ATT_TC");
  if (e && e[0] == '1') return true;
  if (e && e[0] == '0') return false;
  return N > 512;
}

extern "C" int lpc_psa_attention(int dtype, const void* qkv, int qkv_ld, int B, int N, int heads, int kd, int hd,
                                 void* out, int out_ld, void* stream) {
  LPC_REQUIRE(qkv && out && B > 0 && N > 0 && heads > 0, "psa_attention: bad argument");
  LPC_REQUIRE(kd > 0 && kd <= MAX_KD && hd > 0 && hd <= MAX_HD, "psa_attention: kd <= %d, hd <= %d", MAX_KD, MAX_HD);
  LPC_REQUIRE(qkv_ld >= heads * (2 * kd + hd) && out_ld >= heads * hd, "psa_attention: pitch too small");
  dim3 grid(cdiv(N, BQ), heads, B);
  cudaStream_t s = (cudaStream_t)stream;
  static unsigned long long attr_done[3] = {0, 0, 0};  // per device: generic fp32 / generic bf16 / K-V-resident kernels
  if (dtype == LPC_F32 && !getenv("LPC_ATT_F32_FAST")) {
    const size_t sm = (size_t)VAL_WARPS * N * sizeof(double);
    LPC_REQUIRE(sm <= 200 * 1024, "psa_attention (fp32 validation): N = %d too large", N);
    cudaFuncSetAttribute(psa_attention_f32_validate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    lpc_launch_pdl(psa_attention_f32_validate_kernel, dim3(cdiv(N, VAL_WARPS), heads, B), dim3(VAL_WARPS * 32), sm, s, (const float*)qkv, qkv_ld, N, heads, kd,
                   hd, (float*)out, out_ld);
  } else if (dtype == LPC_F32) {
    if (lpc_first_on_device(&attr_done[0])) cudaFuncSetAttribute(psa_attention_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM);
    lpc_launch_pdl(psa_attention_kernel<float>, grid, ATT_NT, ATT_SMEM, s, (const float*)qkv, qkv_ld, N, heads, kd, hd, (float*)out, out_ld);
  } else if (dtype == LPC_BF16 && kd == 32 && hd == 64 && att_use_tc(N) && qkv_ld % 8 == 0 && out_ld % 8 == 0 && aligned16(qkv) && aligned16(out)) {
    // tcgen05 path (attn_tc.cu): S and P in tensor memory, K / V by TMA, V consumed MN-major as it lies in memory
    return lpc_psa_attention_tc(qkv, qkv_ld, B, N, heads, out, out_ld, s);
  } else if (dtype == LPC_BF16 && ((kd == 32 && hd == 64) || (kd == 36 && hd == 72)) && qkv_ld % 8 == 0 && out_ld % 2 == 0 &&
             aligned16(qkv) && (reinterpret_cast<uintptr_t>(out) & 3) == 0) {
    // tensor-core path (the two head geometries of the YOLOv10 / LPC family)
    // K / V resident per (image, head) when there are enough pairs to fill the GPU and they fit two CTAs per SM or one
    // (LPC_ATT_RES=0 / 1 forces the streaming / the resident kernel - the parity tests run both)
    {
      const char* re_ = getenv("LPC_ATT_RES");            // read at every call, like LPC_ATT_TC (the parity tests switch it)
      const int res_env = re_ ? atoi(re_) : -1;
      const int n_pad = cdiv(N, MK) * MK;
      const size_t sm = kd == 32 ? psa_res_smem<32, 64>(n_pad) : psa_res_smem<36, 72>(n_pad);
      const long long pairs = (long long)B * heads;
      const bool fits = sm <= 200 * 1024 && N >= 16;
      bool use = res_env == 1 ? fits : (res_env == 0 ? false : (fits && pairs >= lpc_num_sms() / 2 && N <= 640));
      if (use) {
        if (lpc_first_on_device(&attr_done[2])) {
          cudaFuncSetAttribute(psa_attention_res_kernel<32, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
          cudaFuncSetAttribute(psa_attention_res_kernel<36, 72>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        }
        // few pairs: several CTAs per pair share its query blocks (each loads K / V itself)
        const int per_sm = sm <= 110 * 1024 ? 2 : 1;
        int split = (int)((long long)lpc_num_sms() * per_sm / pairs);
        const int nqb = cdiv(N, 16);
        if (split < 1) split = 1;
        if (split > cdiv(nqb, RES_WARPS)) split = cdiv(nqb, RES_WARPS);
        dim3 g3(split, heads, B);
        if (kd == 32) lpc_launch_pdl(psa_attention_res_kernel<32, 64>, g3, RES_NT, sm, s, (const bf16*)qkv, qkv_ld, N, heads, n_pad, (bf16*)out, out_ld);
        else lpc_launch_pdl(psa_attention_res_kernel<36, 72>, g3, RES_NT, sm, s, (const bf16*)qkv, qkv_ld, N, heads, n_pad, (bf16*)out, out_ld);
        LPC_CHECK_LAUNCH("psa_attention");
        return LPC_OK;
      }
    }
    dim3 g2(cdiv(N, MQ), heads, B);
    if (kd == 32) lpc_launch_pdl(psa_attention_mma_kernel<32, 64>, g2, MMA_NT, 0, s, (const bf16*)qkv, qkv_ld, N, heads, (bf16*)out, out_ld);
    else lpc_launch_pdl(psa_attention_mma_kernel<36, 72>, g2, MMA_NT, 0, s, (const bf16*)qkv, qkv_ld, N, heads, (bf16*)out, out_ld);
  } else if (dtype == LPC_BF16) {
    if (lpc_first_on_device(&attr_done[1])) cudaFuncSetAttribute(psa_attention_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM);
    lpc_launch_pdl(psa_attention_kernel<bf16>, grid, ATT_NT, ATT_SMEM, s, (const bf16*)qkv, qkv_ld, N, heads, kd, hd, (bf16*)out, out_ld);
  } else {
    LPC_FAIL(LPC_E_ARG, "psa_attention: unknown dtype %d", dtype);
  }
  LPC_CHECK_LAUNCH("psa_attention");
  return LPC_OK;
}
