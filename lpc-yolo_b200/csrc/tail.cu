// tail.cu - the v10Detect tail: DFL expectation, anchor decode, sigmoid scores, NMS-free top-k.
//
// Reference semantics (utils/ops.py:851-864): stage 1 keeps the K anchors with the largest max-class
// score, stage 2 keeps the K largest of the K*nc (anchor, class) scores among them, sorted descending.
// We reproduce the two stages on order-preserving integer keys of the LOGITS (sigmoid is monotonic, so
// ranking logits == ranking scores; scores are produced only for the K winners), with a deterministic
// tie rule: equal keys are ordered by ascending flat index anchor*nc + class.
//
//   kernel 1  amax_keys      one pass over the class logits (HBM-bound): key of max_c logit per anchor; on the engine
//                            path the class-branch conv epilogue writes these keys and this kernel is not launched
//   kernel 2  select_decode  one CTA per image: radix-select K anchors, keep the pairs of those anchors that reach the
//                            stage-1 threshold, rank them (all-pairs comparison), box decode of the K winners, write [K,6]
//
// Box decode follows Detect.inference (head.py:45-71): DFL softmax expectation over 16 bins per side
// (block.py:57-60), anchors at cell centre (tal.py:294-306), dist2bbox xywh (tal.py:309-319), * stride,
// then xywh2xyxy (ops.py:402-421) and clip_boxes (ops.py:305-324) as predict.py:20,35 do.
#include "common.cuh"

namespace {

constexpr int REG_MAX = 16;
constexpr int SEL_NT = 1024;
constexpr int KMAX = 1024;

struct TailSrc {
  const void* ptr[3];       // per level base pointer (image 0, cell 0, channel 0)
  long long img_stride[3];  // elements between images
  int a_start[4];           // first anchor of each level; a_start[3] = A
  int lvl_w[3];             // cells per row of each level
  float stride[3];
  long long sa, sc;         // element strides between anchors / between classes
  long long c_off;          // element offset of class 0 inside an anchor row
  int nc, A;
};

__device__ __forceinline__ uint32_t fkey(float f) {
  uint32_t u = __float_as_uint(f);
  return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
  return __uint_as_float(u);
}
// Level lookups as select chains on constant indices: a dynamically indexed kernel parameter would be copied to a
// local-memory stack frame and every row address would start with local loads.
__device__ __forceinline__ int level_of(const TailSrc& s, int a) { return (a >= s.a_start[1]) + (a >= s.a_start[2]); }
template <typename T>
__device__ __forceinline__ const T* anchor_row(const TailSrc& s, int b, int a) {
  const bool l1 = a >= s.a_start[1], l2 = a >= s.a_start[2];
  const void* p = l2 ? s.ptr[2] : (l1 ? s.ptr[1] : s.ptr[0]);
  const long long is = l2 ? s.img_stride[2] : (l1 ? s.img_stride[1] : s.img_stride[0]);
  const int a0 = l2 ? s.a_start[2] : (l1 ? s.a_start[1] : s.a_start[0]);
  return reinterpret_cast<const T*>(p) + (long long)b * is + (long long)(a - a0) * s.sa;
}

// ---------------------------------------------------------------------------------------------------
// kernel 1: per-anchor max key.  VECTOR: 4 lanes share one anchor, lane j loads every 4th 16-byte class vector (an
// anchor's classes are one contiguous run: a warp-wide load covers 8 anchors x 64 B instead of 32 distinct
// lines when every thread walks its own row), then a 2-step shuffle max (16 lanes per anchor was slower: 4x the thread instructions).  Scalar path: one thread
// per anchor (strided class layouts of lpc_v10_postprocess).
template <typename T, bool VECTOR>
__global__ void __launch_bounds__(256)
amax_keys_kernel(TailSrc s, int B, uint32_t* __restrict__ amax) {
  pdl_trigger();
  pdl_wait();
  if (VECTOR) {
    constexpr int V = Vec<T>::N, LPA = 4;
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gid = gt / LPA;                 // anchor (the grid is sized so that whole lane groups are in or out)
    const int sub = (int)(gt & (LPA - 1));
    const bool live = gid < (long long)B * s.A;
    float m = -INFINITY;
    if (live) {
      const int b = (int)(gid / s.A), a = (int)(gid - (long long)b * s.A);
      const T* row = anchor_row<T>(s, b, a) + s.c_off;
      for (int c = sub * V; c < s.nc; c += LPA * V) {
        float f[V];
        ldg_vec<T>(row + c).unpack(f);
#pragma unroll
        for (int v = 0; v < V; ++v) m = fmaxf(m, f[v]);
      }
    }
#pragma unroll
    for (int o = LPA / 2; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (live && sub == 0) amax[gid] = fkey(m);
  } else {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * s.A) return;
    const int b = (int)(gid / s.A), a = (int)(gid - (long long)b * s.A);
    const T* row = anchor_row<T>(s, b, a) + s.c_off;
    float m = -INFINITY;
    for (int c = 0; c < s.nc; ++c) m = fmaxf(m, to_f(row[(long long)c * s.sc]));
    amax[gid] = fkey(m);
  }
}

// ---------------------------------------------------------------------------------------------------
// block-wide helpers (SEL_NT threads)
#ifdef LPC_TAIL_CLOCKS   // tools/run_tail.py: clock64 of CTA 0 at the phase boundaries
__device__ long long g_tail_clk[16];
#define TAIL_CLK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_tail_clk[i] = clock64(); } while (0)
#else
#define TAIL_CLK(i) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------------
// block-wide helpers (SEL_NT threads)
constexpr int CAND_CAP = 4096;   // stage-2 candidates kept in shared memory
constexpr int RANK_CAP = 1024;   // entries ranked by all-pairs comparison (>= KMAX)
constexpr int DIRECT_MAX = 512;  // candidate count up to which they are ranked directly (no second radix select)
struct SelShared {
  uint32_t hist[256];
  uint32_t warp_gt[SEL_NT / 32], warp_eq[SEL_NT / 32];
  uint32_t prefix, need, gt_count, eq_total, digit, ncand;
};

__device__ __forceinline__ void hist_add(SelShared& sh, uint32_t k, int shift) { atomicAdd(&sh.hist[(k >> shift) & 255u], 1u); }

// Radix select over n keys provided by key(i), 8-bit digits from the top.  On return: elements with (key >> kshift) >
// sh.prefix are all winners; the first sh.need elements (in index order) with (key >> kshift) == sh.prefix complete K
// (sh.eq_total elements carry that digit string).
// sh.hist must be ZERO on entry (hist0_ready: it already holds the histogram of the leading digit, accumulated by the
// caller while it loaded the keys) and is zero again on return: the suffix scan and the digit search run inside warp
// 0 (8 bins per lane + shuffles), which clears the bins it has read - two block barriers per pass.
template <int PASSES, typename KeyFn>
__device__ void radix_select(SelShared& sh, int n, int K, KeyFn key, bool hist0_ready) {
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) { sh.prefix = 0; sh.need = (uint32_t)K; }
  for (int pass = 0; pass < PASSES; ++pass) {
    const int shift = 32 - 8 * (pass + 1);
    if (pass > 0 || !hist0_ready) {
      const uint32_t prefix = pass ? sh.prefix : 0u;
#pragma unroll 2
      for (int i = tid; i < n; i += SEL_NT) {
        const uint32_t k = key(i);
        if (pass == 0 || (k >> (shift + 8)) == prefix) hist_add(sh, k, shift);
      }
    }
    __syncthreads();
    if (tid < 32) {
      uint32_t h[8], tot = 0;                      // h[j] = sum of this lane's bins j..7
#pragma unroll
      for (int j = 7; j >= 0; --j) { tot += sh.hist[8 * lane + j]; h[j] = tot; sh.hist[8 * lane + j] = 0; }
      uint32_t incl = tot;                         // -> sum of the lane totals of lanes >= lane
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const uint32_t v = __shfl_down_sync(0xffffffffu, incl, off);
        if (lane + off < 32) incl += v;
      }
      const uint32_t above_lane = incl - tot, need = sh.need;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t cum = h[j] + above_lane, above = (j == 7 ? 0u : h[j + 1]) + above_lane;
        if (cum >= need && above < need) { sh.digit = (uint32_t)(8 * lane + j); sh.gt_count = above; sh.eq_total = cum - above; }
      }
      __syncwarp();
      if (lane == 0) { sh.prefix = (sh.prefix << 8) | sh.digit; sh.need -= sh.gt_count; }
    }
    __syncthreads();
  }
}

// Collect the winners of a finished radix_select IN ASCENDING INDEX ORDER: emit(i, key, slot) is called exactly K times,
// slot = rank of i among the winners; ties with the threshold digit string are admitted in ascending i.  Every thread
// owns a contiguous index range (odd length: conflict-free shared-memory reads); one block-wide exclusive scan of the
// per-thread (greater, tie) counts gives the first slot of each range - one barrier, no atomics, and the output needs
// no sort.  admit_all: every tie is emitted (K - need + eq_total calls; the caller ranks them).  Must be entered right
// after radix_select (reads sh.prefix / sh.need behind its last barrier).
template <int PASSES, typename KeyFn, typename EmitFn>
__device__ void collect(SelShared& sh, int n, KeyFn key, EmitFn emit, bool admit_all = false) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int kshift = 32 - 8 * PASSES;
  const uint32_t prefix = sh.prefix, need = admit_all ? 0xFFFFFFFFu : sh.need;
  const int per = ((n + SEL_NT - 1) / SEL_NT) | 1;
  const int lo = min(n, tid * per), hi = min(n, lo + per);
  uint32_t gt = 0, eq = 0;
  for (int i = lo; i < hi; ++i) {
    const uint32_t top = key(i) >> kshift;
    gt += top > prefix;
    eq += top == prefix;
  }
  uint32_t igt = gt, ieq = eq;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, igt, off), w = __shfl_up_sync(0xffffffffu, ieq, off);
    if (lane >= off) { igt += v; ieq += w; }
  }
  if (lane == 31) { sh.warp_gt[wid] = igt; sh.warp_eq[wid] = ieq; }
  __syncthreads();
  // every warp sums the totals of the warps before it (no second barrier)
  const uint32_t wg = __reduce_add_sync(0xffffffffu, lane < wid ? sh.warp_gt[lane] : 0u);
  const uint32_t we = __reduce_add_sync(0xffffffffu, lane < wid ? sh.warp_eq[lane] : 0u);
  uint32_t r = we + ieq - eq;                                // ties before this thread's range
  int slot = (int)(wg + igt - gt + min(r, need));
  if (gt || (eq && r < need))
    for (int i = lo; i < hi; ++i) {
      const uint32_t k = key(i), top = k >> kshift;
      if (top > prefix) emit(i, k, slot++);
      else if (top == prefix) { if (r < need) emit(i, k, slot++); ++r; }
    }
}

// out[rank] = in[i] for every entry whose rank (number of LARGER entries of in[0..n)) is below K.  Entries are distinct
// and non-zero; in[n..n+3] must be zero (padding); n <= 4 * SEL_NT.  One image is ranked by ONE SM, so the instruction
// count is what matters: a thread holds four entries in registers and walks a slice of in[] with 16-byte loads (8
// comparisons per shared-memory instruction); the P slices of an entry group sit in adjacent lanes (slice length / 2
// odd: no bank conflicts between them) and their partial ranks meet in a shuffle butterfly - no atomics, no barrier.
__device__ void rank_emit(const unsigned long long* __restrict__ in, int n, unsigned long long* __restrict__ out, int K) {
  const int npad = (n + 3) & ~3, G = npad >> 2;
  int P = 32;
  while (P > 1 && G * P > SEL_NT) P >>= 1;
  const int chunk = 2 * (((npad + 2 * P - 1) / (2 * P)) | 1);
  __syncthreads();
  for (int t0 = 0; t0 < G * P; t0 += SEL_NT) {
    const int t = t0 + threadIdx.x, part = t & (P - 1), gq = t / P, g = min(gq, G - 1);
    if ((t & ~31) / P >= G) continue;              // the whole warp is past the last entry group (warp-uniform)
    const int j0 = min(npad, part * chunk), j1 = min(npad, j0 + chunk);
    const ulonglong2 ea = *reinterpret_cast<const ulonglong2*>(in + 4 * g), eb = *reinterpret_cast<const ulonglong2*>(in + 4 * g + 2);
    uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;
#pragma unroll 2
    for (int j = j0; j < j1; j += 2) {
      const ulonglong2 o = *reinterpret_cast<const ulonglong2*>(in + j);
      r0 += (o.x > ea.x) + (o.y > ea.x);
      r1 += (o.x > ea.y) + (o.y > ea.y);
      r2 += (o.x > eb.x) + (o.y > eb.x);
      r3 += (o.x > eb.y) + (o.y > eb.y);
    }
    for (int off = P >> 1; off; off >>= 1) {
      r0 += __shfl_xor_sync(0xffffffffu, r0, off);
      r1 += __shfl_xor_sync(0xffffffffu, r1, off);
      r2 += __shfl_xor_sync(0xffffffffu, r2, off);
      r3 += __shfl_xor_sync(0xffffffffu, r3, off);
    }
    if (part == 0 && gq < G) {
      if (4 * g + 0 < n && r0 < (uint32_t)K) out[r0] = ea.x;
      if (4 * g + 1 < n && r1 < (uint32_t)K) out[r1] = ea.y;
      if (4 * g + 2 < n && r2 < (uint32_t)K) out[r2] = eb.x;
      if (4 * g + 3 < n && r3 < (uint32_t)K) out[r3] = eb.y;
    }
  }
  __syncthreads();
}

// Append the `cnt` (<= 8) flagged candidates of every lane to sh-resident pairs[]: one shared-memory atomic per warp.
__device__ __forceinline__ int warp_reserve(SelShared& sh, int cnt) {
  const int lane = threadIdx.x & 31;
  int incl = cnt;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += v;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  int base = 0;
  if (total) {
    if (lane == 31) base = (int)atomicAdd(&sh.ncand, (uint32_t)total);
    base = __shfl_sync(0xffffffffu, base, 31);
  }
  return base + incl - cnt;
}

// DFL expectation of one box side (block.py:57-60): softmax over the 16 bins, dot with arange(16).  The softmax weights
// are v * (1 / sum) (one IEEE division per side instead of sixteen; <= 1 ulp per weight from v / sum).
template <typename T>
__device__ __forceinline__ float dfl_side(const T* __restrict__ p, bool vec_ok) {
  float v[REG_MAX], mx = -INFINITY;
  if (vec_ok) {
    constexpr int V = Vec<T>::N;
#pragma unroll
    for (int i = 0; i < REG_MAX; i += V) ldg_vec<T>(p + i).unpack(v + i);
  } else {
#pragma unroll
    for (int i = 0; i < REG_MAX; ++i) v[i] = to_f(p[i]);
  }
#pragma unroll
  for (int i = 0; i < REG_MAX; ++i) mx = fmaxf(mx, v[i]);
  float se = 0.f;
#pragma unroll
  for (int i = 0; i < REG_MAX; ++i) { v[i] = expf(v[i] - mx); se += v[i]; }
  const float inv = 1.0f / se;
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < REG_MAX; ++i) acc += (v[i] * inv) * (float)i;
  return acc;
}

// Which elements of a 16-byte vector can still be among the K best pairs: (key >> KSHIFT) >= thr.  bf16 (KSHIFT 16: thr is
// the key of a bf16 value) compares two elements per instruction in the bf16 domain; the float comparison admits a
// superset of the key comparison (-0 >= +0), which only adds a candidate.
template <typename T, int KSHIFT> struct HitTest {
  uint32_t thr;
  __device__ explicit HitTest(uint32_t t) : thr(t) {}
  __device__ __forceinline__ uint32_t mask(const Vec<T>& raw) const {
    float f[Vec<T>::N];
    raw.unpack(f);
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < Vec<T>::N; ++j) m |= (uint32_t)((fkey(f[j]) >> KSHIFT) >= thr) << j;
    return m;
  }
};
template <> struct HitTest<bf16, 16> {
  __nv_bfloat162 t2;
  __device__ explicit HitTest(uint32_t t) {
    const unsigned short h = (unsigned short)((t & 0x8000u) ? (t ^ 0x8000u) : ~t);     // inverse of fkey on the top 16 bits
    t2 = __halves2bfloat162(__ushort_as_bfloat16(h), __ushort_as_bfloat16(h));
  }
  __device__ __forceinline__ uint32_t mask(const Vec<bf16>& raw) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw.raw);
    uint32_t m = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const uint32_t g = __hge2_mask(h[w], t2) & 0x00010001u;        // bit 0: low element, bit 16: high element
      m |= ((g | (g >> 15)) & 3u) << (2 * w);
    }
    return m;
  }
};

// MODE 0: raw head maps -> decoded xyxy dets[B][K][6] (+anchor_idx);  MODE 1: preds passthrough.
//
// One CTA per image.  Stage 1: radix-select the K anchors with the largest max-class key (leading-digit histogram
// accumulated while the keys stream into shared memory), collected in ascending anchor order.  Stage 2: a pair
// (anchor, class) can only be among the K best pairs if its key reaches the threshold digits of stage 1 (each of the K
// selected anchors owns one such pair), so while the K*nc logits of the selected anchors stream in, only those pairs
// are kept: a few hundred to a thousand candidates instead of K*nc keys.  Up to DIRECT_MAX candidates are ranked
// directly (key desc, flat index asc) by one all-pairs comparison and the K best land sorted; more are first cut by a
// radix select over the candidates that keeps every tie of the K-th key, then ranked.  Mass ties (more than CAND_CAP
// candidates or RANK_CAP survivors) fall back to the radix select over all K*nc keys in flat-index order.  Boxes are
// decoded for the winners only, four lanes per winner (one per side).
template <typename T, int PASSES, int MODE>
__global__ void __launch_bounds__(SEL_NT)
select_decode_kernel(const __grid_constant__ TailSrc s, int K, int sortn, const uint32_t* __restrict__ amax, int cache_cap_keys, int vec_ok,
                     int img_h, int img_w, const float* __restrict__ scale_back, float* __restrict__ dets, int* __restrict__ anchor_idx,
                     float* __restrict__ boxes_out, float* __restrict__ scores_out, long long* __restrict__ labels_out) {
  pdl_trigger();
  extern __shared__ __align__(16) unsigned char dsm[];
  __shared__ SelShared sh;
  unsigned long long* pairs = reinterpret_cast<unsigned long long*>(dsm);        // [CAND_CAP + 4] stage-2 candidates
  unsigned long long* win = pairs + CAND_CAP + 4;                                // [RANK_CAP + 4] survivors to rank
  unsigned long long* sorted = win + RANK_CAP + 4;                               // [sortn] the K best pairs, sorted
  uint32_t* sel = reinterpret_cast<uint32_t*>(sorted + sortn);                   // [sortn] selected anchors (ascending)
  uint32_t* cache = sel + sortn;                                                 // [cache_cap_keys] (16-byte aligned)
  const int b = blockIdx.x, tid = threadIdx.x;
  const int A = s.A, nc = s.nc;
  constexpr int kshift = 32 - 8 * PASSES;
  if (tid < 256) sh.hist[tid] = 0;
  if (tid == 0) sh.ncand = 0;
  pdl_wait();
  __syncthreads();
  TAIL_CLK(0);

  // ---- stage 1: K anchors with the largest max-class key ------------------------------------------
  const uint32_t* ak = amax + (long long)b * A;
  const bool cache1 = A <= cache_cap_keys;
  if (cache1) {
    if ((A & 3) == 0 && (reinterpret_cast<uintptr_t>(ak) & 15) == 0) {
      const uint4* ak4 = reinterpret_cast<const uint4*>(ak);
      uint4* c4 = reinterpret_cast<uint4*>(cache);
      for (int i = tid; i < (A >> 2); i += SEL_NT) {
        const uint4 k = __ldg(ak4 + i);
        c4[i] = k;
        hist_add(sh, k.x, 24); hist_add(sh, k.y, 24); hist_add(sh, k.z, 24); hist_add(sh, k.w, 24);
      }
    } else {
      for (int i = tid; i < A; i += SEL_NT) {
        const uint32_t k = ak[i];
        cache[i] = k;
        hist_add(sh, k, 24);
      }
    }
  }
  const uint32_t* kp = cache1 ? cache : ak;        // generic loads: one loop body for both homes of the keys
  auto key1 = [&](int i) -> uint32_t { return kp[i]; };
  TAIL_CLK(1);
  radix_select<PASSES>(sh, A, K, key1, cache1);
  const uint32_t thr = sh.prefix;          // a winning pair has (key >> kshift) >= thr
  TAIL_CLK(2);
  collect<PASSES>(sh, A, key1, [&](int i, uint32_t, int slot) { sel[slot] = (uint32_t)i; });
  __syncthreads();
  TAIL_CLK(3);

  // ---- stage 2: the K largest of the K*nc pair keys ---------------------------------------------------
  const int n2 = K * nc;
  auto load2 = [&](int i) -> uint32_t {
    const int slot = i / nc, c = i - slot * nc;
    const T* row = anchor_row<T>(s, b, (int)sel[slot]) + s.c_off;
    return fkey(to_f(row[(long long)c * s.sc]));
  };
  constexpr int V = Vec<T>::N;
  if (vec_ok) {                                    // 16-byte loads along the classes of each selected anchor,
    const int cvn = nc / V, nv = K * cvn;          // four per thread in flight, one slot reservation per warp
    const float inv_cvn = 1.0f / (float)cvn;
    const HitTest<T, kshift> test(thr);
    constexpr int U = 4;
    for (int v0 = 0; v0 < nv; v0 += U * SEL_NT) {
      Vec<T> raw[U];
      uint32_t flat0[U], hit[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int v = v0 + u * SEL_NT + tid;
        hit[u] = 0;
        if (v < nv) {
          int slot = (int)((float)v * inv_cvn), cv = v - slot * cvn;       // v < 2^23: off by one at most
          if (cv < 0) { --slot; cv += cvn; } else if (cv >= cvn) { ++slot; cv -= cvn; }
          const uint32_t a = sel[slot];
          raw[u] = ldg_vec<T>(anchor_row<T>(s, b, (int)a) + s.c_off + cv * V);
          flat0[u] = a * (uint32_t)nc + (uint32_t)(cv * V);
        }
      }
      int cnt = 0;
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (v0 + u * SEL_NT + tid < nv) { hit[u] = test.mask(raw[u]); cnt += __popc(hit[u]); }
      int pos = warp_reserve(sh, cnt);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (hit[u]) {
          float f[V];
          raw[u].unpack(f);
#pragma unroll
          for (int j = 0; j < V; ++j)
            if ((hit[u] >> j) & 1u) {
              if (pos < CAND_CAP) pairs[pos] = ((unsigned long long)fkey(f[j]) << 32) | (unsigned long long)(0xFFFFFFFFu - (flat0[u] + j));
              ++pos;
            }
        }
    }
  } else {
    for (int i0 = 0; i0 < n2; i0 += SEL_NT) {
      const int i = i0 + tid;
      uint32_t k = 0;
      bool hit = false;
      if (i < n2) { k = load2(i); hit = (k >> kshift) >= thr; }
      const int pos = warp_reserve(sh, hit ? 1 : 0);
      if (hit && pos < CAND_CAP) {
        const int slot = i / nc, c = i - slot * nc;
        pairs[pos] = ((unsigned long long)k << 32) | (unsigned long long)(0xFFFFFFFFu - (sel[slot] * (uint32_t)nc + (uint32_t)c));
      }
    }
  }
  __syncthreads();
  TAIL_CLK(4);
  const int n_cand = (int)sh.ncand;
#ifdef LPC_TAIL_CLOCKS
  if (blockIdx.x == 0 && tid == 0) g_tail_clk[15] = n_cand;
#endif
  const unsigned long long* ranked = pairs;        // what rank_emit sees
  int n_ranked = n_cand;
  bool fallback = n_cand > CAND_CAP || n_cand < K;     // fewer than K: NaN logits (never equal to themselves)
  if (!fallback && n_cand > DIRECT_MAX && n_cand > K) {
    // cut the candidates to the K best keys plus every tie of the K-th key; their order is settled by the ranking
    const uint32_t* pk = reinterpret_cast<const uint32_t*>(pairs);
    auto keyc = [&](int i) -> uint32_t { return pk[2 * i + 1]; };
    radix_select<PASSES>(sh, n_cand, K, keyc, false);
    const int m = K - (int)sh.need + (int)sh.eq_total;
    if (m <= RANK_CAP) {
      collect<PASSES>(sh, n_cand, keyc, [&](int i, uint32_t, int slot) { win[slot] = pairs[i]; }, true);
      ranked = win;
      n_ranked = m;
    } else {
      fallback = true;
      __syncthreads();                             // sh.need / sh.eq_total are rewritten below
    }
  }
  if (fallback) {
    // mass ties: radix select over all K*nc keys in flat-index order (cached in shared memory when they fit)
    const bool cache2 = n2 <= cache_cap_keys;
    if (cache2) {
      for (int i = tid; i < n2; i += SEL_NT) cache[i] = load2(i);
      __syncthreads();
    }
    auto key2 = [&](int i) -> uint32_t { return cache2 ? cache[i] : load2(i); };
    radix_select<PASSES>(sh, n2, K, key2, false);
    collect<PASSES>(sh, n2, key2, [&](int i, uint32_t k, int slot) {
      const int sl = i / nc, c = i - sl * nc;
      win[slot] = ((unsigned long long)k << 32) | (unsigned long long)(0xFFFFFFFFu - (sel[sl] * (uint32_t)nc + (uint32_t)c));
    });
    ranked = win;
    n_ranked = K;
  }
  if (tid < 4) const_cast<unsigned long long*>(ranked)[n_ranked + tid] = 0;
  TAIL_CLK(5);
  rank_emit(ranked, n_ranked, sorted, K);       // key desc, then flat index asc (entries are distinct)
  TAIL_CLK(6);

  // ---- winners: scores, labels, boxes ------------------------------------------------------------------
  if (MODE == 0) {
    const int side = tid & 3, lane = tid & 31;
#pragma unroll 1
    for (int r0 = 0; r0 < K; r0 += SEL_NT / 4) {
      const int r = r0 + (tid >> 2);
      if (r0 + ((tid & ~31) >> 2) >= K) continue;  // none of this warp's eight winners exists (warp-uniform)
      const unsigned long long e = sorted[r < K ? r : 0];
      const uint32_t flat = 0xFFFFFFFFu - (uint32_t)(e & 0xFFFFFFFFull);
      const int a = (int)(flat / (uint32_t)nc), c = (int)(flat - (uint32_t)a * nc);
      const float ds = dfl_side<T>(anchor_row<T>(s, b, a) + side * REG_MAX, vec_ok != 0);
      const float d0 = __shfl_sync(0xffffffffu, ds, lane & ~3), d1 = __shfl_sync(0xffffffffu, ds, (lane & ~3) + 1);
      const float d2 = __shfl_sync(0xffffffffu, ds, (lane & ~3) + 2), d3 = __shfl_sync(0xffffffffu, ds, (lane & ~3) + 3);
      if (r < K && side == 0) {
        const bool l1 = a >= s.a_start[1], l2 = a >= s.a_start[2];
        const int cell = a - (l2 ? s.a_start[2] : (l1 ? s.a_start[1] : s.a_start[0]));
        const int lw = l2 ? s.lvl_w[2] : (l1 ? s.lvl_w[1] : s.lvl_w[0]);
        const float st = l2 ? s.stride[2] : (l1 ? s.stride[1] : s.stride[0]);
        const int cy = cell / lw, cx = cell - cy * lw;
        const float ax = (float)cx + 0.5f, ay = (float)cy + 0.5f;
        const float x1 = ax - d0, y1 = ay - d1, x2 = ax + d2, y2 = ay + d3;
        const float bx = (x1 + x2) / 2 * st, by = (y1 + y2) / 2 * st, bw = (x2 - x1) * st, bh = (y2 - y1) * st;
        float X1 = bx - bw / 2, Y1 = by - bh / 2, X2 = bx + bw / 2, Y2 = by + bh / 2;
        if (img_h > 0) {
          X1 = fminf(fmaxf(X1, 0.f), (float)img_w); X2 = fminf(fmaxf(X2, 0.f), (float)img_w);
          Y1 = fminf(fmaxf(Y1, 0.f), (float)img_h); Y2 = fminf(fmaxf(Y2, 0.f), (float)img_h);
        }
        if (scale_back) {
          // ops.scale_boxes (utils/ops.py:89-124) + clip_boxes (:305-324) of this image: subtract the LetterBox padding,
          // divide by the resize gain (IEEE division, as torch's fp32 ``boxes /= gain``), clamp to the original image
          const float* sp = scale_back + (long long)b * 5;
          const float px = sp[0], py = sp[1], gain = sp[2], ow = sp[3], oh = sp[4];
          X1 = fminf(fmaxf((X1 - px) / gain, 0.f), ow); X2 = fminf(fmaxf((X2 - px) / gain, 0.f), ow);
          Y1 = fminf(fmaxf((Y1 - py) / gain, 0.f), oh); Y2 = fminf(fmaxf((Y2 - py) / gain, 0.f), oh);
        }
        float2* o = reinterpret_cast<float2*>(dets + ((long long)b * K + r) * 6);   // 24-byte rows: 8-byte aligned
        o[0] = make_float2(X1, Y1);
        o[1] = make_float2(X2, Y2);
        o[2] = make_float2(1.0f / (1.0f + expf(-key2f((uint32_t)(e >> 32)))), (float)c);
        if (anchor_idx) anchor_idx[(long long)b * K + r] = a;
      }
    }
  } else {
    for (int r = tid; r < K; r += SEL_NT) {
      const unsigned long long e = sorted[r];
      const uint32_t flat = 0xFFFFFFFFu - (uint32_t)(e & 0xFFFFFFFFull);
      const int a = (int)(flat / (uint32_t)nc), c = (int)(flat - (uint32_t)a * nc);
      const T* row = anchor_row<T>(s, b, a);
      float* bo = boxes_out + ((long long)b * K + r) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) bo[i] = to_f(row[(long long)i * s.sc]);
      scores_out[(long long)b * K + r] = key2f((uint32_t)(e >> 32));  // keys were taken on the scores themselves
      labels_out[(long long)b * K + r] = c;
    }
  }
  TAIL_CLK(7);
}

// ---------------------------------------------------------------------------------------------------
// Detect.inference as a standalone op: y[B][4+nc][A] fp32 (module-level drop-in for v10Detect.forward)
template <typename T>
__global__ void __launch_bounds__(128)
decode_kernel(TailSrc s, int B, float* __restrict__ y) {
  pdl_trigger();
  pdl_wait();
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)B * s.A) return;
  const int b = (int)(gid / s.A), a = (int)(gid - (long long)b * s.A);
  const int l = level_of(s, a);
  const int cell = a - s.a_start[l];
  const int cy = cell / s.lvl_w[l], cx = cell - cy * s.lvl_w[l];
  const T* row = anchor_row<T>(s, b, a);
  float* yo = y + (long long)b * (4 + s.nc) * s.A + a;
  if (Precise<T>::value) {
    // fp32 validation mode: softmax, expectation, box arithmetic and sigmoid in fp64, one rounding per output, so y adds
    // nothing to the distance the raw maps already have from an fp64 run (tests/test_gpu_e2e.py::test_fp32_mode_raw_and_y)
    double d[4];
#pragma unroll 1
    for (int side = 0; side < 4; ++side) {
      double v[REG_MAX], mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < REG_MAX; ++i) { v[i] = (double)to_f(row[side * REG_MAX + i]); mx = fmax(mx, v[i]); }
      double se = 0.0, acc = 0.0;
#pragma unroll
      for (int i = 0; i < REG_MAX; ++i) { v[i] = exp(v[i] - mx); se += v[i]; }
#pragma unroll
      for (int i = 0; i < REG_MAX; ++i) acc += (v[i] / se) * (double)i;
      d[side] = acc;
    }
    const double ax = (double)cx + 0.5, ay = (double)cy + 0.5, st = (double)s.stride[l];
    const double x1 = ax - d[0], y1 = ay - d[1], x2 = ax + d[2], y2 = ay + d[3];
    yo[0] = (float)((x1 + x2) / 2 * st);
    yo[(long long)s.A] = (float)((y1 + y2) / 2 * st);
    yo[2ll * s.A] = (float)((x2 - x1) * st);
    yo[3ll * s.A] = (float)((y2 - y1) * st);
    for (int c = 0; c < s.nc; ++c) yo[(long long)(4 + c) * s.A] = (float)(1.0 / (1.0 + exp(-(double)to_f(row[4 * REG_MAX + c]))));
    return;
  }
  float d[4];
#pragma unroll
  for (int side = 0; side < 4; ++side) {
    float v[REG_MAX], mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < REG_MAX; ++i) { v[i] = to_f(row[side * REG_MAX + i]); mx = fmaxf(mx, v[i]); }
    float se = 0.f;
#pragma unroll
    for (int i = 0; i < REG_MAX; ++i) { v[i] = expf(v[i] - mx); se += v[i]; }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < REG_MAX; ++i) acc += (v[i] / se) * (float)i;
    d[side] = acc;
  }
  const float ax = (float)cx + 0.5f, ay = (float)cy + 0.5f, st = s.stride[l];
  const float x1 = ax - d[0], y1 = ay - d[1], x2 = ax + d[2], y2 = ay + d[3];
  yo[0] = (x1 + x2) / 2 * st;
  yo[(long long)s.A] = (y1 + y2) / 2 * st;
  yo[2ll * s.A] = (x2 - x1) * st;
  yo[3ll * s.A] = (y2 - y1) * st;
  for (int c = 0; c < s.nc; ++c) {
    const float v = to_f(row[4 * REG_MAX + c]);
    yo[(long long)(4 + c) * s.A] = 1.0f / (1.0f + expf(-v));
  }
}

int make_raw_src(TailSrc& s, const char* name, int dtype, const void* raw0, const void* raw1, const void* raw2, int ld,
                 int B, int H0, int W0, int nc, const float* strides) {
  LPC_REQUIRE(raw0 && raw1 && raw2 && strides, "%s: null pointer", name);
  LPC_REQUIRE(B > 0 && H0 > 0 && W0 > 0 && nc > 0, "%s: bad shape", name);
  LPC_REQUIRE(H0 % 4 == 0 && W0 % 4 == 0, "%s: level-0 map must be divisible by 4 (three levels)", name);
  LPC_REQUIRE(ld >= 4 * REG_MAX + nc, "%s: pitch smaller than 64+nc", name);
  const void* p[3] = {raw0, raw1, raw2};
  int a = 0;
  for (int l = 0; l < 3; ++l) {
    const int h = H0 >> l, w = W0 >> l;
    s.ptr[l] = p[l];
    s.img_stride[l] = (long long)h * w * ld;
    s.a_start[l] = a;
    s.lvl_w[l] = w;
    s.stride[l] = strides[l];
    a += h * w;
  }
  s.a_start[3] = a;
  s.A = a;
  s.nc = nc;
  s.sa = ld;
  s.sc = 1;
  s.c_off = 4 * REG_MAX;
  (void)dtype;
  return LPC_OK;
}

int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

constexpr int SEL_SMEM_BUDGET = 200 * 1024;

template <typename T, int PASSES, int MODE>
int launch_select(const TailSrc& s, int B, int K, const uint32_t* amax, int img_h, int img_w, const float* scale_back, float* dets, int* aidx,
                  float* boxes, float* scores, long long* labels, cudaStream_t st) {
  const int sortn = next_pow2(K) < 4 ? 4 : next_pow2(K);   // multiple of 4: keeps the key cache 16-byte aligned
  const size_t fixed = (size_t)(CAND_CAP + 4 + RANK_CAP + 4) * 8 + (size_t)sortn * (8 + 4);
  int cap = (int)((SEL_SMEM_BUDGET - fixed) / 4);
  const int want = s.A > K * s.nc ? s.A : K * s.nc;
  if (cap > want) cap = want;
  const size_t smem = fixed + (size_t)cap * 4;
  auto kern = select_decode_kernel<T, PASSES, MODE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) LPC_FAIL(LPC_E_CUDA, "select_decode: smem attribute: %s", cudaGetErrorString(e));
  // 16-byte loads along an anchor row: contiguous classes, rows and class block 16-byte aligned
  constexpr int V = Vec<T>::N;
  const int vec_ok = MODE == 0 && s.sc == 1 && s.nc % V == 0 && s.sa % V == 0 && s.c_off % V == 0 && aligned16(s.ptr[0]) && aligned16(s.ptr[1]) &&
                     aligned16(s.ptr[2]) && s.img_stride[0] % V == 0 && s.img_stride[1] % V == 0 && s.img_stride[2] % V == 0;
  lpc_launch_pdl(kern, B, SEL_NT, smem, st, s, K, sortn, amax, cap, vec_ok, img_h, img_w, scale_back, dets, aidx, boxes, scores, labels);
  LPC_CHECK_LAUNCH("select_decode");
  return LPC_OK;
}

}  // namespace

#ifdef LPC_TAIL_CLOCKS
extern "C" int lpc_debug_tail_clocks(long long* out16) {   // tools only (not declared in include/lpcyolo.h)
  return cudaMemcpyFromSymbol(out16, g_tail_clk, sizeof(long long) * 16) == cudaSuccess ? 0 : -3;
}
#endif

extern "C" size_t lpc_v10_topk_workspace_bytes(int B, int A, int K) {
  (void)K;
  return (size_t)B * (size_t)A * sizeof(uint32_t) + 256;
}

extern "C" int lpc_v10_decode(int dtype, const void* raw0, const void* raw1, const void* raw2, int ld, int B, int H0,
                              int W0, int nc, const float* strides, float* y, void* stream) {
  TailSrc s;
  if (int e = make_raw_src(s, "v10_decode", dtype, raw0, raw1, raw2, ld, B, H0, W0, nc, strides)) return e;
  LPC_REQUIRE(y, "v10_decode: null output");
  cudaStream_t st = (cudaStream_t)stream;
  const int g = cdiv((long long)B * s.A, 128);
  if (dtype == LPC_F32) lpc_launch_pdl(decode_kernel<float>, g, 128, 0, st, s, B, y);
  else if (dtype == LPC_BF16) lpc_launch_pdl(decode_kernel<bf16>, g, 128, 0, st, s, B, y);
  else LPC_FAIL(LPC_E_ARG, "v10_decode: unknown dtype %d", dtype);
  LPC_CHECK_LAUNCH("v10_decode");
  return LPC_OK;
}

extern "C" int lpc_v10_decode_topk(int dtype, const void* raw0, const void* raw1, const void* raw2, int ld, int B,
                                   int H0, int W0, int nc, const float* strides, int K, int img_h, int img_w,
                                   void* workspace, size_t ws_bytes, float* dets, int* anchor_idx, void* stream) {
  return lpc_v10_decode_topk_keys(dtype, raw0, raw1, raw2, ld, B, H0, W0, nc, strides, K, img_h, img_w, workspace, ws_bytes, 0, dets,
                                  anchor_idx, stream);
}

extern "C" int lpc_v10_decode_topk_keys(int dtype, const void* raw0, const void* raw1, const void* raw2, int ld, int B,
                                        int H0, int W0, int nc, const float* strides, int K, int img_h, int img_w,
                                        void* workspace, size_t ws_bytes, int keys_ready, float* dets, int* anchor_idx, void* stream) {
  return lpc_v10_decode_topk_scaled(dtype, raw0, raw1, raw2, ld, B, H0, W0, nc, strides, K, img_h, img_w, workspace, ws_bytes, keys_ready,
                                    nullptr, dets, anchor_idx, stream);
}

extern "C" int lpc_v10_decode_topk_scaled(int dtype, const void* raw0, const void* raw1, const void* raw2, int ld, int B,
                                          int H0, int W0, int nc, const float* strides, int K, int img_h, int img_w,
                                          void* workspace, size_t ws_bytes, int keys_ready, const float* scale_back,
                                          float* dets, int* anchor_idx, void* stream) {
  TailSrc s;
  if (int e = make_raw_src(s, "v10_decode_topk", dtype, raw0, raw1, raw2, ld, B, H0, W0, nc, strides)) return e;
  LPC_REQUIRE(dets && workspace, "v10_decode_topk: null pointer");
  LPC_REQUIRE(K > 0 && K <= KMAX && s.A >= K, "v10_decode_topk: need 0 < K <= %d and A >= K (A=%d, K=%d)", KMAX, s.A, K);
  if (ws_bytes < lpc_v10_topk_workspace_bytes(B, s.A, K)) LPC_FAIL(LPC_E_WORKSPACE, "v10_decode_topk: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* amax = reinterpret_cast<uint32_t*>(workspace);
  const int g = cdiv((long long)B * s.A, 256);
  if (dtype == LPC_BF16) {
    const bool vec = (nc % 8 == 0) && (ld % 8 == 0) && aligned16(raw0) && aligned16(raw1) && aligned16(raw2);
    if (!keys_ready) {
      if (vec) lpc_launch_pdl(amax_keys_kernel<bf16, true>, cdiv((long long)B * s.A * 4, 256), 256, 0, st, s, B, amax);
      else lpc_launch_pdl(amax_keys_kernel<bf16, false>, g, 256, 0, st, s, B, amax);
      LPC_CHECK_LAUNCH("amax_keys");
    }
    return launch_select<bf16, 2, 0>(s, B, K, amax, img_h, img_w, scale_back, dets, anchor_idx, nullptr, nullptr, nullptr, st);
  } else if (dtype == LPC_F32) {
    const bool vec = (nc % 4 == 0) && (ld % 4 == 0) && aligned16(raw0) && aligned16(raw1) && aligned16(raw2);
    if (!keys_ready) {
      if (vec) lpc_launch_pdl(amax_keys_kernel<float, true>, cdiv((long long)B * s.A * 4, 256), 256, 0, st, s, B, amax);
      else lpc_launch_pdl(amax_keys_kernel<float, false>, g, 256, 0, st, s, B, amax);
      LPC_CHECK_LAUNCH("amax_keys");
    }
    return launch_select<float, 4, 0>(s, B, K, amax, img_h, img_w, scale_back, dets, anchor_idx, nullptr, nullptr, nullptr, st);
  }
  LPC_FAIL(LPC_E_ARG, "v10_decode_topk: unknown dtype %d", dtype);
}

extern "C" int lpc_v10_postprocess(const float* preds, long long stride_b, long long stride_a, long long stride_c,
                                   int B, int A, int nc, int K, void* workspace, size_t ws_bytes, float* boxes,
                                   float* scores, long long* labels, void* stream) {
  LPC_REQUIRE(preds && boxes && scores && labels && workspace, "v10_postprocess: null pointer");
  LPC_REQUIRE(B > 0 && nc > 0 && K > 0 && K <= KMAX && A >= K, "v10_postprocess: need 0 < K <= %d and A >= K (A=%d, K=%d)", KMAX, A, K);
  if (ws_bytes < lpc_v10_topk_workspace_bytes(B, A, K)) LPC_FAIL(LPC_E_WORKSPACE, "v10_postprocess: workspace too small");
  TailSrc s;
  for (int l = 0; l < 3; ++l) { s.ptr[l] = preds; s.img_stride[l] = stride_b; s.lvl_w[l] = 1; s.stride[l] = 1.f; }
  s.a_start[0] = 0; s.a_start[1] = s.a_start[2] = s.a_start[3] = A;  // every anchor resolves to level 0
  s.a_start[1] = A; s.a_start[2] = A;
  s.A = A; s.nc = nc; s.sa = stride_a; s.sc = stride_c; s.c_off = 4 * stride_c;
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* amax = reinterpret_cast<uint32_t*>(workspace);
  const int g = cdiv((long long)B * A, 256);
  lpc_launch_pdl(amax_keys_kernel<float, false>, g, 256, 0, st, s, B, amax);
  LPC_CHECK_LAUNCH("amax_keys");
  return launch_select<float, 4, 1>(s, B, K, amax, 0, 0, nullptr, nullptr, nullptr, boxes, scores, labels, st);
}
