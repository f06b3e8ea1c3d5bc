// mma_bench.cu - microbenchmark of tcgen05.mma (kind::f16, cta_group::1, M = 128) issue / execution rate on sm_100a as a
// function of N, accumulator dependence, operand swizzle and A-operand placement.  No global traffic: operands are
// whatever sits in shared memory.  One CTA per SM; one elected thread issues `iters` x `chain` MMAs and commits.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I lpc-yolo_b200/csrc tools/mma_bench.cu -o tools/mma_bench
//   tools/mma_bench            -> table of cycles per MMA
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "tc_ptx.cuh"

struct Cfg {
  int n;          // MMA N
  int accs;       // accumulators rotated over (1 = every MMA depends on the previous one)
  int a_layout;   // 2 = 128B swizzle, 4 = 64B, 6 = 32B
  int a_shift;    // A start offset in bytes from the 1024-aligned buffer (multiple of the row pitch): tap shift
  int a_sbo;      // bytes between 8-row groups of A
  int chain;      // MMAs per commit
  int iters;
  int m;          // MMA M (128 or 64)
  int ats;        // 1: A operand from TMEM (tcgen05.mma [d], [a_tmem], b_desc)
  int per_sm;     // CTAs per SM (1 or 2): 2 halves the TMEM / shared memory per CTA
  int ldtm;       // 1: four other warps hammer tcgen05.ld on a second accumulator region meanwhile
};

__global__ void __launch_bounds__(192, 2) mma_bench_kernel(Cfg c, unsigned long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < (c.per_sm == 2 ? 20 : 48) * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), c.per_sm == 2 ? 256u : 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  volatile __shared__ int stop;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t idesc = (make_idesc(c.n) & ~(0x1Fu << 24)) | ((uint32_t)(c.m >> 4) << 24);
      const uint32_t a_hi = desc_hi((uint32_t)c.a_sbo, (uint32_t)c.a_layout), b_hi = desc_hi(1024u, 2u);
      const uint32_t a_lo = desc_lo(base + (uint32_t)c.a_shift, 16u), b_lo = desc_lo(base + (c.per_sm == 2 ? 40 : 96) * 1024, 16u);
      const int acc_cols = c.n < 32 ? 32 : c.n;
      uint32_t ph = 0;
      // warm-up
      for (int k = 0; k < 8; ++k) umma_bf16(tmem, desc64(a_lo, a_hi), desc64(b_lo, b_hi), idesc, 0u);
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), ph); ph ^= 1;
      const long long t0 = clock64();
      long long t_issue = 0;
      for (int it = 0; it < c.iters; ++it) {
        for (int k = 0; k < c.chain; ++k) {
          const uint32_t acc = tmem + (uint32_t)((k % c.accs) * acc_cols);
          if (c.ats) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(acc), "r"(tmem + (c.per_sm == 2 ? 192u : 384u) + 8u * (uint32_t)(k & 3)), "l"(desc64(b_lo + 2u * (uint32_t)(k & 3), b_hi)), "r"(idesc) : "memory");
          } else {
            umma_acc(acc, desc64(a_lo + 2u * (uint32_t)(k & 3), a_hi), desc64(b_lo + 2u * (uint32_t)(k & 3), b_hi), idesc);
          }
        }
        umma_commit(smem_u32(&bar));
        if (it == c.iters - 1) t_issue = clock64();
        mbar_wait(smem_u32(&bar), ph); ph ^= 1;
      }
      const long long t1 = clock64();
      if (blockIdx.x == 0) { out[0] = (unsigned long long)(t1 - t0); out[1] = (unsigned long long)(t_issue - t0); }
      stop = 1;
    }
  } else if (warp >= 2 && c.ldtm) {
    uint32_t v[16];
    uint32_t sink = 0;
    const uint32_t trow = tmem + (c.per_sm == 2 ? 128u : 256u) + ((uint32_t)((warp & 3) * 32) << 16);
    while (!stop) {
      for (int cidx = 0; cidx < 64; cidx += 16) {
        tmem_ld16(trow + (uint32_t)cidx, v);
        tmem_ld_wait();
        sink += v[0] + v[15];
      }
    }
    if (sink == 0x12345678u) out[3] = sink;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, c.per_sm == 2 ? 256u : 512u);
  (void)lane;
}

// ---- cost of the per-tile bookkeeping ops of the issuing thread ---------------------------------------------
__global__ void __launch_bounds__(64, 1) op_cost_kernel(unsigned long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[4];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);            // completes on every commit
    mbar_init(smem_u32(&bars[1]), 1u << 20);     // never completes: absorbs arrivals
    mbar_init(smem_u32(&bars[2]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1 && elect_one_sync()) {
    const uint32_t idesc = make_idesc(16);
    const uint64_t ad = desc64(desc_lo(base, 16u), desc_hi(1024u, 2u)), bd = desc64(desc_lo(base + 32768, 16u), desc_hi(1024u, 2u));
    const uint32_t b0 = smem_u32(&bars[0]), b1 = smem_u32(&bars[1]), b2 = smem_u32(&bars[2]);
    const int R = 200;
    long long t[8];
    // complete phase 0 of bars[2] so that waiting on parity 0 is an "already complete" wait
    mbar_arrive(b2);
    t[0] = clock64();
    for (int i = 0; i < R; ++i) mbar_wait(b2, 0);                       // ready wait
    t[1] = clock64();
    for (int i = 0; i < R; ++i) tc_fence_after();
    t[2] = clock64();
    for (int i = 0; i < R; ++i) umma_commit(b1);                         // commit with nothing pending
    t[3] = clock64();
    for (int i = 0; i < R; ++i) { umma_bf16(tmem, ad, bd, idesc, 0u); umma_commit(b1); }   // 1 MMA + commit
    t[4] = clock64();
    for (int i = 0; i < R; ++i) { umma_bf16(tmem, ad, bd, idesc, 0u); umma_commit(b1); umma_commit(b1); }
    t[5] = clock64();
    uint32_t ph = 0;
    for (int i = 0; i < R; ++i) { umma_bf16(tmem, ad, bd, idesc, 0u); umma_commit(b0); mbar_wait(b0, ph); ph ^= 1; }   // full round trip
    t[6] = clock64();
    for (int i = 0; i < R; ++i) { mbar_wait(b2, 0); mbar_wait(b2, 0); tc_fence_after(); umma_bf16(tmem, ad, bd, idesc, 0u); umma_acc(tmem, ad, bd, idesc); umma_acc(tmem, ad, bd, idesc); umma_commit(b1); umma_commit(b1); }
    t[7] = clock64();
    if (blockIdx.x == 0) for (int i = 0; i < 7; ++i) out[i] = (unsigned long long)(t[i + 1] - t[i]) / R;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 64);
}

// ---- cta_group::2: one instruction drives the tensor cores of a CTA pair (M = 256) ---------------------------
// Does the ~70-cycle per-instruction floor of small-N MMAs halve per SM when one issue serves two SMs?
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64, 1) pair_bench_kernel(int n, int chain, int iters, unsigned long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 20 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (rank == 0 && warp == 1 && elect_one_sync()) {
    const uint32_t idesc = (make_idesc(n) & ~(0x1Fu << 24)) | ((256u >> 4) << 24);     // M = 256 across the pair
    const uint32_t a_hi = desc_hi(1024u, 2u), b_hi = desc_hi(1024u, 2u);
    const uint32_t a_lo = desc_lo(base, 16u), b_lo = desc_lo(base + 40 * 1024, 16u);
    uint32_t ph = 0;
    auto mma2 = [&](uint32_t acc, uint32_t k, uint32_t accumulate) {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(acc), "l"(desc64(a_lo + 2u * (k & 3), a_hi)), "l"(desc64(b_lo + 2u * (k & 3), b_hi)), "r"(idesc), "r"(accumulate) : "memory");
    };
    auto commit2 = [&]() {
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((unsigned short)1) : "memory");
    };
    for (int k = 0; k < 8; ++k) mma2(tmem, k, 0u);
    commit2();
    mbar_wait(smem_u32(&bar), ph); ph ^= 1;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int k = 0; k < chain; ++k) mma2(tmem, (uint32_t)k, 1u);
      commit2();
      mbar_wait(smem_u32(&bar), ph); ph ^= 1;
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
  }
  tc_fence_before();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

int main() {
  unsigned long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("%4s %4s %6s %7s %6s %5s %5s | %10s %10s %10s\n", "N", "accs", "layout", "shift", "sbo", "chain", "ldtm", "cyc/MMA", "issue/MMA", "floor");
  auto run = [&](Cfg c, int grid) {
    cudaMemset(d, 0, 64);
    mma_bench_kernel<<<grid * c.per_sm, 192, (c.per_sm == 2 ? 90 : 200) * 1024>>>(c, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
    unsigned long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    const double n = (double)c.iters * c.chain;
    printf("M%3d ts%d %4d %4d %6d %7d %6d %5d %5d | %10.1f %10.1f %10.1f  (grid %d)\n", c.m, c.ats, c.n, c.accs, c.a_layout, c.a_shift, c.a_sbo, c.chain, c.ldtm, h[0] / n, h[1] / n,
           128.0 * c.n / 256.0, grid);
    fflush(stdout);
  };
  const int ns[] = {16, 32, 64, 128, 256};
  for (int n : ns) {
    run({n, 1, 2, 0, 1024, 36, 50, 128, 0, 1, 0}, sms);
    if (n <= 128) run({n, 2, 2, 0, 1024, 36, 50, 128, 0, 1, 0}, sms);
    if (n <= 64) run({n, 4, 2, 0, 1024, 36, 50, 128, 0, 1, 0}, sms);
  }
  // halo-style A: 128B rows, 8-row groups 2048 B apart, shifted start
  for (int n : {32, 64, 128}) {
    run({n, 1, 2, 0, 2048, 36, 50, 128, 0, 1, 0}, sms);
    run({n, 1, 2, 128, 2048, 36, 50, 128, 0, 1, 0}, sms);
    run({n, 1, 2, 17 * 128, 2048, 36, 50, 128, 0, 1, 0}, sms);
  }
  // 64B / 32B swizzle (Cin 32 / 16 patches)
  run({32, 1, 4, 0, 1024, 36, 50, 128, 0, 1, 0}, sms);
  run({32, 1, 4, 64 * 17, 1024, 36, 50, 128, 0, 1, 0}, sms);
  run({32, 1, 6, 0, 512, 36, 50, 128, 0, 1, 0}, sms);
  run({32, 1, 6, 32 * 17, 512, 36, 50, 128, 0, 1, 0}, sms);
  run({64, 1, 6, 32 * 17, 512, 36, 50, 128, 0, 1, 0}, sms);
  // short chains (9 MMAs per tile as in Cin = 16)
  run({32, 1, 6, 0, 512, 9, 200, 128, 0, 1, 0}, sms);
  run({32, 2, 6, 0, 512, 9, 200, 128, 0, 1, 0}, sms);
  run({64, 1, 2, 0, 1024, 9, 200, 128, 0, 1, 0}, sms);
  // concurrent TMEM reads by the epilogue warps
  run({64, 1, 2, 0, 1024, 36, 50, 128, 0, 1, 1}, sms);
  run({128, 1, 2, 0, 1024, 36, 50, 128, 0, 1, 1}, sms);
  // M = 64 (is the ~69-cycle floor the A read?) and A from TMEM
  for (int n : {32, 64, 128, 256}) run({n, 1, 2, 0, 1024, 36, 50, 64, 0, 1, 0}, sms);
  for (int n : {32, 64, 128, 256}) run({n, 1, 2, 0, 1024, 36, 50, 128, 1, 1, 0}, sms);
  for (int n : {32, 64, 128, 256}) run({n, 1, 2, 0, 1024, 36, 50, 64, 1, 1, 0}, sms);
  // two CTAs per SM, each with its own issuing thread
  for (int n : {16, 32, 64, 128}) run({n, 1, 2, 0, 1024, 36, 50, 128, 0, 2, 0}, sms);
  for (int n : {32, 64}) run({n, 1, 2, 0, 1024, 36, 50, 128, 0, 2, 1}, sms);
  // a single CTA on the chip (no neighbours)
  run({64, 1, 2, 0, 1024, 36, 50, 128, 0, 1, 0}, 1);
  run({256, 1, 2, 0, 1024, 36, 50, 128, 0, 1, 0}, 1);
  {
    cudaFuncSetAttribute(pair_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int n : {32, 64, 128, 256}) {
      cudaMemset(d, 0, 64);
      pair_bench_kernel<<<sms / 2 * 2, 64, 100 * 1024>>>(n, 36, 50, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("pair bench error: %s\n", cudaGetErrorString(e)); return 1; }
      unsigned long long h[1];
      cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
      printf("cta_group::2  M256 N%3d chain 36 | %10.1f cycles per MMA (covers 2 SMs; single-CTA floor 128*N/256 = %.0f)\n", n, h[0] / (36.0 * 50), 128.0 * n / 256.0);
      fflush(stdout);
    }
  }
  {
    cudaFuncSetAttribute(op_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaMemset(d, 0, 64);
    op_cost_kernel<<<sms, 64, 100 * 1024>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    unsigned long long h[8];
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("cycles per op (issuing thread): ready mbar_wait %llu | fence::after %llu | commit (idle) %llu | mma+commit %llu | mma+2 commits %llu | mma+commit+wait round trip %llu | stem-like tile loop (2 waits, fence, 3 mma, 2 commits) %llu\n",
           h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
  }
  return 0;
}
