// tail.cu - the v10Detect tail: DFL expectation, anchor decode, sigmoid scores, NMS-free top-k.
//
// Reference semantics (utils/ops.py:851-864): stage 1 keeps the K anchors with the largest max-class
// score, stage 2 keeps the K largest of the K*nc (anchor, class) scores among them, sorted descending.
// We reproduce the two stages on order-preserving integer keys of the LOGITS (sigmoid is monotonic, so
// ranking logits == ranking scores; scores are produced only for the K winners), with a deterministic
// tie rule: equal keys are ordered by ascending flat index anchor*nc + class.
//
//   kernel 1  amax_keys      one pass over the class logits (HBM-bound): key of max_c logit per anchor
//   kernel 2  select_decode  one CTA per image: radix-select K anchors, radix-select K pairs, bitonic sort,
//                            box decode of the winners only, write [K,6]
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
__device__ __forceinline__ int level_of(const TailSrc& s, int a) { return (a >= s.a_start[1]) + (a >= s.a_start[2]); }
template <typename T>
__device__ __forceinline__ const T* anchor_row(const TailSrc& s, int b, int a) {
  const int l = level_of(s, a);
  return reinterpret_cast<const T*>(s.ptr[l]) + (long long)b * s.img_stride[l] + (long long)(a - s.a_start[l]) * s.sa;
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
struct SelShared {
  uint32_t hist[256];
  uint32_t cum[256];
  uint32_t warp_cnt[SEL_NT / 32];
  uint32_t prefix, need, gt_count, eq_base, digit;
};

// Radix select over n keys provided by key(i).  On return: elements with (key >> kshift) > sh.prefix are
// all winners; the first sh.need elements (in index order) with (key >> kshift) == sh.prefix complete K.
// Three block barriers per pass: the 256-bin suffix scan and the digit search run inside warp 0 (8 bins per lane +
// shuffles) instead of a 16-barrier Hillis-Steele scan over shared memory.
template <int PASSES, typename KeyFn>
__device__ void radix_select(SelShared& sh, int n, int K, KeyFn key) {
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) { sh.prefix = 0; sh.need = (uint32_t)K; }
  for (int pass = 0; pass < PASSES; ++pass) {
    if (tid < 256) sh.hist[tid] = 0;
    __syncthreads();
    const int shift = 32 - 8 * (pass + 1);
    const uint32_t prefix = sh.prefix;
    for (int i = tid; i < n; i += SEL_NT) {
      const uint32_t k = key(i);
      if (pass == 0 || (k >> (shift + 8)) == prefix) atomicAdd(&sh.hist[(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid < 32) {
      uint32_t h[8], tot = 0;                      // h[j] = sum of this lane's bins j..7
#pragma unroll
      for (int j = 7; j >= 0; --j) { tot += sh.hist[8 * lane + j]; h[j] = tot; }
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
        if (cum >= need && above < need) { sh.digit = (uint32_t)(8 * lane + j); sh.gt_count = above; }
      }
      __syncwarp();
      if (lane == 0) { sh.prefix = (sh.prefix << 8) | sh.digit; sh.need -= sh.gt_count; }
    }
    __syncthreads();
  }
}

// Collect winners of a finished radix_select: emit(i, key, slot) is called exactly K times in total; ties are
// admitted in ascending i.  Output slots: [0, n_gt) for strictly-greater (arbitrary order), then ties in index order.
// Every thread owns a CONTIGUOUS index range (odd length: conflict-free shared-memory reads), so the tie ranks come
// from one block-wide exclusive scan of per-thread tie counts (4 barriers in total, not 3 per 1024 elements).
template <int PASSES, typename KeyFn, typename EmitFn>
__device__ void collect(SelShared& sh, int n, int K, KeyFn key, EmitFn emit) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int kshift = 32 - 8 * PASSES;
  const uint32_t prefix = sh.prefix, need = sh.need;
  const uint32_t n_gt = (uint32_t)K - need;
  __syncthreads();
  if (tid == 0) sh.gt_count = 0;
  __syncthreads();
  const int per = ((n + SEL_NT - 1) / SEL_NT) | 1;
  const int lo = min(n, tid * per), hi = min(n, lo + per);
  uint32_t cnt = 0;
  for (int i = lo; i < hi; ++i) {
    const uint32_t k = key(i), top = k >> kshift;
    if (top > prefix) emit(i, k, (int)atomicAdd(&sh.gt_count, 1u));
    else if (top == prefix) ++cnt;
  }
  uint32_t incl = cnt;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += v;
  }
  if (lane == 31) sh.warp_cnt[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = sh.warp_cnt[lane], wi = w;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, wi, off);
      if (lane >= off) wi += v;
    }
    sh.warp_cnt[lane] = wi - w;                    // exclusive prefix of the warp totals
  }
  __syncthreads();
  uint32_t r = sh.warp_cnt[wid] + incl - cnt;      // ties before this thread's range
  if (cnt && r < need)
    for (int i = lo; i < hi && r < need; ++i) {
      const uint32_t k = key(i);
      if ((k >> kshift) == prefix) { emit(i, k, (int)(n_gt + r)); ++r; }
    }
}

// out[rank] = in[i] where rank = number of entries of in[0..n) that precede in[i] (DESC: larger first).  Entries are
// distinct, so the ranks are a permutation; every comparison operand is a shared-memory broadcast.  One barrier on
// each side replaces the 45 barrier steps of a 512-element bitonic network.
template <bool DESC, typename U>
__device__ void rank_sort(const U* in, U* out, int n) {
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += SEL_NT) {
    const U e = in[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const U o = in[j];
      rank += DESC ? (o > e) : (o < e);
    }
    out[rank] = e;
  }
  __syncthreads();
}

// MODE 0: raw head maps -> decoded xyxy dets[B][K][6] (+anchor_idx);  MODE 1: preds passthrough.
template <typename T, int PASSES, int MODE>
__global__ void __launch_bounds__(SEL_NT)
select_decode_kernel(TailSrc s, int K, int sortn, const uint32_t* __restrict__ amax, int cache_cap_keys, int vec_ok,
                     int img_h, int img_w, float* __restrict__ dets, int* __restrict__ anchor_idx,
                     float* __restrict__ boxes_out, float* __restrict__ scores_out, long long* __restrict__ labels_out) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) unsigned char dsm[];
  __shared__ SelShared sh;
  unsigned long long* sortbuf = reinterpret_cast<unsigned long long*>(dsm);      // [sortn] stage-2 winners (unsorted)
  unsigned long long* sorted = sortbuf + sortn;                                  // [sortn] ... sorted
  uint32_t* sel = reinterpret_cast<uint32_t*>(sorted + sortn);                   // [sortn] selected anchors (ascending)
  uint32_t* sel_raw = sel + sortn;                                               // [sortn] ... in collection order
  uint32_t* cache = sel_raw + sortn;                                             // [cache_cap_keys]
  const int b = blockIdx.x, tid = threadIdx.x;
  const int A = s.A, nc = s.nc;
  const uint32_t* ak = amax + (long long)b * A;

  // ---- stage 1: K anchors with the largest max-class key ------------------------------------------
  const bool cache1 = A <= cache_cap_keys;
  if (cache1) {
    for (int i = tid; i < A; i += SEL_NT) cache[i] = ak[i];
    __syncthreads();
  }
  auto key1 = [&](int i) -> uint32_t { return cache1 ? cache[i] : ak[i]; };
  radix_select<PASSES>(sh, A, K, key1);
  collect<PASSES>(sh, A, K, key1, [&](int i, uint32_t, int slot) { sel_raw[slot] = (uint32_t)i; });
  // ascending anchor order (so that candidate position order == flat (anchor, class) order)
  rank_sort<false>(sel_raw, sel, K);

  // ---- stage 2: K largest of the K*nc pair keys ------------------------------------------------------
  const int n2 = K * nc;
  const bool cache2 = n2 <= cache_cap_keys;
  auto load2 = [&](int i) -> uint32_t {
    const int slot = i / nc, c = i - slot * nc;
    const T* row = anchor_row<T>(s, b, (int)sel[slot]) + s.c_off;
    return fkey(to_f(row[(long long)c * s.sc]));
  };
  if (cache2) {
    constexpr int V = Vec<T>::N;
    if (vec_ok) {                                  // 16-byte loads along the classes of each selected anchor
      const int cvn = nc / V;
      for (int v = tid; v < K * cvn; v += SEL_NT) {
        const int slot = v / cvn, cv = v - slot * cvn;
        const T* row = anchor_row<T>(s, b, (int)sel[slot]) + s.c_off + cv * V;
        float f[V];
        ldg_vec<T>(row).unpack(f);
#pragma unroll
        for (int j = 0; j < V; ++j) cache[slot * nc + cv * V + j] = fkey(f[j]);
      }
    } else {
      for (int i = tid; i < n2; i += SEL_NT) cache[i] = load2(i);
    }
    __syncthreads();
  }
  auto key2 = [&](int i) -> uint32_t { return cache2 ? cache[i] : load2(i); };
  radix_select<PASSES>(sh, n2, K, key2);
  collect<PASSES>(sh, n2, K, key2, [&](int i, uint32_t k, int slot) {
    const int sl = i / nc, c = i - sl * nc;
    const uint32_t flat = sel[sl] * (uint32_t)nc + (uint32_t)c;
    sortbuf[slot] = ((unsigned long long)k << 32) | (unsigned long long)(0xFFFFFFFFu - flat);
  });
  rank_sort<true>(sortbuf, sorted, K);   // key desc, then flat index asc (entries are distinct)
  sortbuf = sorted;

  // ---- winners: scores, labels, boxes ------------------------------------------------------------------
  for (int r = tid; r < K; r += SEL_NT) {
    const unsigned long long e = sortbuf[r];
    const uint32_t k = (uint32_t)(e >> 32);
    const uint32_t flat = 0xFFFFFFFFu - (uint32_t)(e & 0xFFFFFFFFull);
    const int a = (int)(flat / (uint32_t)nc), c = (int)(flat - (uint32_t)a * nc);
    const float logit = key2f(k);
    if (MODE == 0) {
      const int l = level_of(s, a);
      const int cell = a - s.a_start[l];
      const int cy = cell / s.lvl_w[l], cx = cell - cy * s.lvl_w[l];
      const T* row = anchor_row<T>(s, b, a);
      float d[4];
#pragma unroll
      for (int side = 0; side < 4; ++side) {
        float v[REG_MAX], mx = -INFINITY;
        if (vec_ok) {
          constexpr int V = Vec<T>::N;
#pragma unroll
          for (int i = 0; i < REG_MAX; i += V) ldg_vec<T>(row + side * REG_MAX + i).unpack(v + i);
        } else {
#pragma unroll
          for (int i = 0; i < REG_MAX; ++i) v[i] = to_f(row[side * REG_MAX + i]);
        }
#pragma unroll
        for (int i = 0; i < REG_MAX; ++i) mx = fmaxf(mx, v[i]);
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
      const float bx = (x1 + x2) / 2 * st, by = (y1 + y2) / 2 * st, bw = (x2 - x1) * st, bh = (y2 - y1) * st;
      float X1 = bx - bw / 2, Y1 = by - bh / 2, X2 = bx + bw / 2, Y2 = by + bh / 2;
      if (img_h > 0) {
        X1 = fminf(fmaxf(X1, 0.f), (float)img_w); X2 = fminf(fmaxf(X2, 0.f), (float)img_w);
        Y1 = fminf(fmaxf(Y1, 0.f), (float)img_h); Y2 = fminf(fmaxf(Y2, 0.f), (float)img_h);
      }
      float* o = dets + ((long long)b * K + r) * 6;
      o[0] = X1; o[1] = Y1; o[2] = X2; o[3] = Y2;
      o[4] = 1.0f / (1.0f + expf(-logit));
      o[5] = (float)c;
      if (anchor_idx) anchor_idx[(long long)b * K + r] = a;
    } else {
      const T* row = anchor_row<T>(s, b, a);
      float* bo = boxes_out + ((long long)b * K + r) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) bo[i] = to_f(row[(long long)i * s.sc]);
      scores_out[(long long)b * K + r] = logit;  // keys were taken on the scores themselves
      labels_out[(long long)b * K + r] = c;
    }
  }
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
  float* yo = y + (long long)b * (4 + s.nc) * s.A + a;
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
int launch_select(const TailSrc& s, int B, int K, const uint32_t* amax, int img_h, int img_w, float* dets, int* aidx,
                  float* boxes, float* scores, long long* labels, cudaStream_t st) {
  const int sortn = next_pow2(K);
  const size_t fixed = (size_t)sortn * (8 + 8 + 4 + 4);
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
  lpc_launch_pdl(kern, B, SEL_NT, smem, st, s, K, sortn, amax, cap, vec_ok, img_h, img_w, dets, aidx, boxes, scores, labels);
  LPC_CHECK_LAUNCH("select_decode");
  return LPC_OK;
}

}  // namespace

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
    return launch_select<bf16, 2, 0>(s, B, K, amax, img_h, img_w, dets, anchor_idx, nullptr, nullptr, nullptr, st);
  } else if (dtype == LPC_F32) {
    const bool vec = (nc % 4 == 0) && (ld % 4 == 0) && aligned16(raw0) && aligned16(raw1) && aligned16(raw2);
    if (!keys_ready) {
      if (vec) lpc_launch_pdl(amax_keys_kernel<float, true>, cdiv((long long)B * s.A * 4, 256), 256, 0, st, s, B, amax);
      else lpc_launch_pdl(amax_keys_kernel<float, false>, g, 256, 0, st, s, B, amax);
      LPC_CHECK_LAUNCH("amax_keys");
    }
    return launch_select<float, 4, 0>(s, B, K, amax, img_h, img_w, dets, anchor_idx, nullptr, nullptr, nullptr, st);
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
  return launch_select<float, 4, 1>(s, B, K, amax, 0, 0, nullptr, nullptr, boxes, scores, labels, st);
}
