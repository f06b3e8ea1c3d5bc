// attn_tc.cu - PSA attention core (block.py:783-795) on the 5th-generation tensor cores, bf16, kd = 32 / hd = 64
// (every YOLOv10 / LPC model except yolov10m's kd 36 / hd 72, which stays on the mma.sync kernel of attn.cu).
//
//   out[i, :] = sum_j softmax_j(scale * q_i . k_j) v_j          per (image, head), N = H * W tokens (<= 1600 here)
//
// One CTA = 128 queries of one (image, head).  Keys / values stream through a TMA ring in blocks of 128 tokens:
//   S block [128 q x 128 keys] = Q K_j^T      tcgen05.mma, A = Q and B = K_j K-major in 64B-swizzled shared memory, fp32 in TMEM
//   softmax                                   four warps, one thread per query row, tcgen05.ld of its S row
//   O [128 q x 64]            += P_j V_j      tcgen05.mma with A = P_j read from TENSOR MEMORY (bf16 pairs written back with
//                                             tcgen05.st) and B = V_j exactly as it lies in memory: tokens x 64 channels,
//                                             i.e. MN-major (instruction descriptor bit 16), 128B-swizzled - no transpose
// Two passes over the keys instead of an online softmax: pass 1 only reduces the row maxima (QK^T is a K = 32
// contraction, cheap to repeat), pass 2 forms P = exp2((S - max) * scale * log2 e) against the FINAL maximum, so the
// O accumulator in TMEM never needs rescaling.  The N x N matrix the reference materialises never exists.
// Warps: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = softmax / epilogue.  All mbarrier waits are bounded.
#include <cuda.h>

#include <cstring>
#include <mutex>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int AT_KD = 32, AT_HD = 64, AT_BQ = 128, AT_BK = 128, AT_STAGES = 3;
constexpr int AT_Q_BYTES = AT_BQ * AT_KD * 2;            // 8 KB, 64-byte rows
constexpr int AT_K_BYTES = AT_BK * AT_KD * 2;            // 8 KB
constexpr int AT_V_BYTES = AT_BK * AT_HD * 2;            // 16 KB, 128-byte rows
constexpr int AT_STAGE_BYTES = AT_K_BYTES + AT_V_BYTES;
constexpr int AT_SMEM = AT_Q_BYTES + AT_STAGES * AT_STAGE_BYTES + 1024;
constexpr uint32_t AT_TS = 0, AT_TP = 128, AT_TO = 192, AT_TMEM_COLS = 256;   // S fp32 | P bf16 pairs | O fp32

struct PsaTcParams {
  int N, heads, nblk;
  float scale_log2e;
  bf16* out;
  long long out_ld;
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                 "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(192, 2)
psa_attention_tc_kernel(const __grid_constant__ CUtensorMap qk_map, const __grid_constant__ CUtensorMap v_map, const __grid_constant__ PsaTcParams p) {
  extern __shared__ __align__(1024) unsigned char at_smem[];
  __shared__ __align__(8) unsigned long long bars[2 * AT_STAGES + 4];
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(at_smem) + 1023u) & ~1023u;
  const uint32_t q_s = base, ring = base + AT_Q_BYTES;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto kv_full = [&](int s) { return bar0 + 8u * s; };
  auto kv_empty = [&](int s) { return bar0 + 8u * (AT_STAGES + s); };
  const uint32_t q_full = bar0 + 8u * (2 * AT_STAGES), s_full = q_full + 8u, sm_done = q_full + 16u, o_full = q_full + 24u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * AT_BQ, h = blockIdx.y, b = blockIdx.z;
  const int nblk = p.nblk, total = 2 * nblk;

  if (tid == 0) {
    prefetch_tmap(&qk_map);
    prefetch_tmap(&v_map);
    for (int s = 0; s < AT_STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    mbar_init(q_full, 1);
    mbar_init(s_full, 1);
    mbar_init(sm_done, 4);
    mbar_init(o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), AT_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_trigger();

  if (warp == 0) {
    if (elect_one_sync()) {
      pdl_wait();
      mbar_expect_tx(q_full, (uint32_t)AT_Q_BYTES);
      tma_load_3d(q_s, &qk_map, q_full, h * AT_KD, q0, b);
      const int kc = p.heads * AT_KD + h * AT_KD, vc = 2 * p.heads * AT_KD + h * AT_HD;
      int s = 0;
      uint32_t ph = 1;
      for (int it = 0; it < total; ++it) {
        const int j = it < nblk ? it : it - nblk;
        const bool pass2 = it >= nblk;
        mbar_wait(kv_empty(s), ph);
        mbar_expect_tx(kv_full(s), (uint32_t)(pass2 ? AT_STAGE_BYTES : AT_K_BYTES));
        const uint32_t dst = ring + (uint32_t)(s * AT_STAGE_BYTES);
        tma_load_3d(dst, &qk_map, kv_full(s), kc, j * AT_BK, b);
        if (pass2) tma_load_3d(dst + AT_K_BYTES, &v_map, kv_full(s), vc, j * AT_BK, b);
        if (++s == AT_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t idesc_s = make_idesc(AT_BK);                          // M 128 x N 128, both operands K-major
      const uint32_t idesc_pv = make_idesc(AT_HD) | (1u << 16);            // M 128 x N 64, B (= V) MN-major
      const uint32_t hi_qk = desc_hi(512u, 4u);                            // 64-byte rows, 64B swizzle: 8-row groups 512 B apart
      const uint32_t hi_v = desc_hi(1024u, 2u);                            // 128-byte rows (one token), 128B swizzle: 8-token groups 1024 B apart
      const uint32_t q_lo = desc_lo(q_s, 16u);
      mbar_wait(q_full, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < total; ++it) {
        const bool pass2 = it >= nblk;
        const int j = pass2 ? it - nblk : it;
        mbar_wait(kv_full(s), ph);
        if (it > 0 && it <= nblk) mbar_wait(sm_done, (uint32_t)((it - 1) & 1));   // pass 1 (and the first pass-2 block): S has been read
        tc_fence_after();
        const uint32_t k_lo = desc_lo(ring + (uint32_t)(s * AT_STAGE_BYTES), 16u);
#pragma unroll
        for (int k = 0; k < AT_KD / 16; ++k) umma_bf16(tmem + AT_TS, desc64(q_lo + 2u * k, hi_qk), desc64(k_lo + 2u * k, hi_qk), idesc_s, (uint32_t)(k != 0));
        umma_commit(s_full);
        if (pass2) {
          mbar_wait(sm_done, (uint32_t)(it & 1));                          // this block's P is in tensor memory (and S has been read)
          tc_fence_after();
          const uint32_t v_lo = desc_lo(ring + (uint32_t)(s * AT_STAGE_BYTES + AT_K_BYTES), 16u);
#pragma unroll
          for (int kk = 0; kk < AT_BK / 16; ++kk)
            umma_ts(tmem + AT_TO, tmem + AT_TP + 8u * kk, desc64(v_lo + (uint32_t)(kk * 128), hi_v), idesc_pv, (uint32_t)((j | kk) != 0));
        }
        umma_commit(kv_empty(s));
        if (++s == AT_STAGES) { s = 0; ph ^= 1u; }
      }
      umma_commit(o_full);
    }
  } else {
    // ===== softmax / epilogue: thread = one query row =====
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    const uint32_t ts = tmem + AT_TS + lane_sel, tp = tmem + AT_TP + lane_sel, to = tmem + AT_TO + lane_sel;
    float m = -INFINITY;
    for (int it = 0; it < nblk; ++it) {                       // pass 1: row maxima
      mbar_wait(s_full, (uint32_t)(it & 1));
      tc_fence_after();
      const int col0 = it * AT_BK;
#pragma unroll 2
      for (int c = 0; c < AT_BK / 16; ++c) {
        uint32_t v[16];
        tmem_ld16(ts + (uint32_t)(c * 16), v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (col0 + c * 16 + i < p.N) m = fmaxf(m, __uint_as_float(v[i]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sm_done);
    }
    const float sl2 = p.scale_log2e, msc = m * sl2;
    float sum = 0.f;
    for (int it = nblk; it < total; ++it) {                   // pass 2: P = exp2((S - m) * scale * log2 e) -> bf16 -> tensor memory
      mbar_wait(s_full, (uint32_t)(it & 1));
      tc_fence_after();
      const int col0 = (it - nblk) * AT_BK;
#pragma unroll 1
      for (int c = 0; c < AT_BK / 32; ++c) {
        uint32_t v0[16], v1[16], w[16];
        tmem_ld16(ts + (uint32_t)(c * 32), v0);
        tmem_ld16(ts + (uint32_t)(c * 32 + 16), v1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k0 = col0 + c * 32 + 2 * i, k1 = k0 + 16;
          const float a0 = k0 < p.N ? ex2_approx(fmaf(__uint_as_float(v0[2 * i]), sl2, -msc)) : 0.f;
          const float a1 = k0 + 1 < p.N ? ex2_approx(fmaf(__uint_as_float(v0[2 * i + 1]), sl2, -msc)) : 0.f;
          const float b0 = k1 < p.N ? ex2_approx(fmaf(__uint_as_float(v1[2 * i]), sl2, -msc)) : 0.f;
          const float b1 = k1 + 1 < p.N ? ex2_approx(fmaf(__uint_as_float(v1[2 * i + 1]), sl2, -msc)) : 0.f;
          sum += (a0 + a1) + (b0 + b1);
          __nv_bfloat162 ha = __floats2bfloat162_rn(a0, a1), hb = __floats2bfloat162_rn(b0, b1);
          w[i] = *reinterpret_cast<uint32_t*>(&ha);           // keys 32c + 2i, +1   -> P column 16c + i
          w[8 + i] = *reinterpret_cast<uint32_t*>(&hb);       // keys 32c + 16 + 2i  -> P column 16c + 8 + i
        }
        tmem_st16(tp + (uint32_t)(c * 16), w);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sm_done);
    }
    mbar_wait(o_full, 0);
    tc_fence_after();
    const int q = q0 + r;
    const float inv = 1.0f / sum;
    bf16* orow = p.out + ((long long)b * p.N + q) * p.out_ld + h * AT_HD;
#pragma unroll
    for (int c = 0; c < AT_HD / 16; ++c) {
      uint32_t v[16], w[8];
      tmem_ld16(to + (uint32_t)(c * 16), v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        __nv_bfloat162 hv = __floats2bfloat162_rn(__uint_as_float(v[2 * i]) * inv, __uint_as_float(v[2 * i + 1]) * inv);
        w[i] = *reinterpret_cast<uint32_t*>(&hv);
      }
      if (q < p.N) {
        *reinterpret_cast<uint4*>(orow + c * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(orow + c * 16 + 8) = make_uint4(w[4], w[5], w[6], w[7]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, AT_TMEM_COLS);
}

typedef CUresult (*AtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
AtEncodeFn at_get_encode() {
  static AtEncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<AtEncodeFn>(f);
  });
  return fn;
}

}  // namespace

// Internal (called by lpc_psa_attention in attn.cu): returns LPC_OK when launched, LPC_E_UNSUPPORTED when this path declines.
int lpc_psa_attention_tc(const void* qkv, int ld, int B, int N, int heads, void* out, int out_ld, cudaStream_t stream) {
  AtEncodeFn enc = at_get_encode();
  if (!enc) return LPC_E_UNSUPPORTED;
  if (ld % 8 || out_ld % 8 || !aligned16(qkv) || !aligned16(out) || N < 1 || B > 65535 || heads > 65535) return LPC_E_UNSUPPORTED;
  CUtensorMap qk_map, v_map;
  cuuint64_t dims[3] = {(cuuint64_t)(heads * (2 * AT_KD + AT_HD)), (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)N * ld * 2};
  cuuint32_t es[3] = {1, 1, 1};
  cuuint32_t box_qk[3] = {(cuuint32_t)AT_KD, (cuuint32_t)AT_BK, 1};
  cuuint32_t box_v[3] = {(cuuint32_t)AT_HD, (cuuint32_t)AT_BK, 1};
  if (enc(&qk_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), dims, strides, box_qk, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    LPC_FAIL(LPC_E_CUDA, "psa_attention_tc: q/k tensor map encode failed");
  if (enc(&v_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), dims, strides, box_v, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    LPC_FAIL(LPC_E_CUDA, "psa_attention_tc: v tensor map encode failed");
  PsaTcParams p;
  memset(&p, 0, sizeof(p));
  p.N = N;
  p.heads = heads;
  p.nblk = (N + AT_BK - 1) / AT_BK;
  p.scale_log2e = 1.4426950408889634f / sqrtf((float)AT_KD);
  p.out = (bf16*)out;
  p.out_ld = out_ld;
  static unsigned long long attr_done = 0;
  if (lpc_first_on_device(&attr_done))
    if (cudaFuncSetAttribute(psa_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM) != cudaSuccess) {
      attr_done = 0;
      LPC_FAIL(LPC_E_CUDA, "psa_attention_tc: smem attribute");
    }
  lpc_launch_pdl(psa_attention_tc_kernel, dim3((unsigned)((N + AT_BQ - 1) / AT_BQ), (unsigned)heads, (unsigned)B), dim3(192), (size_t)AT_SMEM, stream, qk_map, v_map, p);
  LPC_CHECK_LAUNCH("psa_attention_tc");
  return LPC_OK;
}
