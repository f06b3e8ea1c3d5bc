// glue.cu - bandwidth-bound neck glue and pooled-attention gates (NHWC, 128-bit vector accesses).
//   upsample2x / copy_channels / space_to_depth / channel_deinterleave / pack_input      (K7)
//   global_avgpool / channel_mlp / cbam_stats / cbam_apply                              (K8)
#include "common.cuh"

namespace {

// y[n, y, x, :] = x[n, y/2, x/2, :]
template <typename T>
__global__ void upsample2x_kernel(const T* __restrict__ x, int x_ld, int B, int H, int W, int C, T* __restrict__ y, int y_ld) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = Vec<T>::N;
  const int cvecs = C / V;
  const unsigned total = (unsigned)B * H * W * cvecs;  // one thread per INPUT vector, writes 4 outputs (host checks < 2^32)
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const unsigned p = idx / (unsigned)cvecs;
  const int cv = (int)(idx - p * (unsigned)cvecs);
  const unsigned t = p / (unsigned)W;
  const int ix = (int)(p - t * (unsigned)W);
  const int n = (int)(t / (unsigned)H);
  const int iy = (int)(t - (unsigned)n * (unsigned)H);
  Vec<T> v = ldg_vec<T>(x + (long long)p * x_ld + cv * V);
  const int Wo = 2 * W;
  T* o = y + (((long long)n * 2 * H + 2 * iy) * Wo + 2 * ix) * y_ld + cv * V;
  st_vec<T>(o, v);
  st_vec<T>(o + y_ld, v);
  st_vec<T>(o + (long long)Wo * y_ld, v);
  st_vec<T>(o + (long long)Wo * y_ld + y_ld, v);
}

template <typename T>
__global__ void copy_channels_kernel(const T* __restrict__ x, int x_ld, long long npix, int C, T* __restrict__ y, int y_ld) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = Vec<T>::N;
  const int cvecs = C / V;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npix * cvecs) return;
  const int cv = (int)(idx % cvecs);
  const long long p = idx / cvecs;
  st_vec<T>(y + p * y_ld + cv * V, ldg_vec<T>(x + p * x_ld + cv * V));
}

// out[n, oy, ox, q*C + c] = x[n, 2*oy + (q&1), 2*ox + (q>>1), c]   q = 0..3
// (block.py:4070: [::2,::2], [1::2,::2], [::2,1::2], [1::2,1::2] -> row parity varies fastest)
template <typename T>
__global__ void space_to_depth_kernel(const T* __restrict__ x, int x_ld, int B, int H, int W, int C, T* __restrict__ y, int y_ld) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = Vec<T>::N;
  const int cvecs = C / V;
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)B * Ho * Wo * 4 * cvecs;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = (int)(idx % cvecs);
  long long t = idx / cvecs;
  const int q = (int)(t & 3);
  t >>= 2;
  const int ox = (int)(t % Wo);
  t /= Wo;
  const int oy = (int)(t % Ho);
  const int n = (int)(t / Ho);
  const int iy = 2 * oy + (q & 1), ix = 2 * ox + (q >> 1);
  Vec<T> v = ldg_vec<T>(x + ((long long)(n * H + iy) * W + ix) * x_ld + cv * V);
  st_vec<T>(y + ((long long)(n * Ho + oy) * Wo + ox) * y_ld + q * C + cv * V, v);
}

// y[p, j] = x[p, 2j], y[p, C/2 + j] = x[p, 2j+1]  (LPC's channel de-interleave, block.py:5824)
template <typename T>
__global__ void deinterleave_kernel(const T* __restrict__ x, int x_ld, long long npix, int C, T* __restrict__ y, int y_ld) {
  pdl_trigger();
  pdl_wait();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npix * C) return;
  const int c = (int)(idx % C);
  const long long p = idx / C;
  const int src = c < C / 2 ? 2 * c : 2 * (c - C / 2) + 1;
  y[p * y_ld + c] = x[p * x_ld + src];
}
// Vector version: a thread takes 32 input bytes (16 bf16 / 8 fp32 channels) and writes one 16-byte vector of the even
// and one of the odd channels (byte permutes, no arithmetic).  Needs C % (2 * V) == 0 and 16-byte aligned rows.
template <typename T>
__global__ void deinterleave_vec_kernel(const T* __restrict__ x, int x_ld, unsigned npix, int C, T* __restrict__ y, int y_ld) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = Vec<T>::N;
  const int groups = C / (2 * V);
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npix * (unsigned)groups) return;
  const unsigned p = idx / (unsigned)groups;
  const int g = (int)(idx - p * (unsigned)groups);
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(x + (long long)p * x_ld + g * 2 * V));
  const uint4 b = __ldg(reinterpret_cast<const uint4*>(x + (long long)p * x_ld + g * 2 * V + V));
  uint4 ev, od;
  if (sizeof(T) == 2) {
    ev = make_uint4(__byte_perm(a.x, a.y, 0x5410), __byte_perm(a.z, a.w, 0x5410), __byte_perm(b.x, b.y, 0x5410), __byte_perm(b.z, b.w, 0x5410));
    od = make_uint4(__byte_perm(a.x, a.y, 0x7632), __byte_perm(a.z, a.w, 0x7632), __byte_perm(b.x, b.y, 0x7632), __byte_perm(b.z, b.w, 0x7632));
  } else {
    ev = make_uint4(a.x, a.z, b.x, b.z);
    od = make_uint4(a.y, a.w, b.y, b.w);
  }
  *reinterpret_cast<uint4*>(y + (long long)p * y_ld + g * V) = ev;
  *reinterpret_cast<uint4*>(y + (long long)p * y_ld + C / 2 + g * V) = od;
}

template <typename T>
__global__ void pack_input_kernel(const float* __restrict__ x, int B, int C, int H, int W, T* __restrict__ y, int y_ld, int Cpad) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)B * H * W;
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const long long hw = (long long)H * W;
  const int n = (int)(p / hw);
  const long long r = p - n * hw;
  for (int c = 0; c < Cpad; ++c) {
    float v = c < C ? __ldg(x + ((long long)n * C + c) * hw + r) : 0.f;
    y[p * y_ld + c] = from_f<T>(v);
  }
}


// uint8 HWC (cv2 / BGR) image batch -> NHWC network input with C padded to 4: the reference's preprocess for array sources
// (engine/predictor.py:115-133: stack -> BGR->RGB -> BHWC->BCHW -> float -> /255) plus LetterBox's constant border
// (data/augment.py:725-731, value 114) for the no-resize case, in ONE pass: the image is already NHWC, so the
// reference's transpose disappears.  Thread = output pixel: 3 byte loads (a warp reads 96 contiguous bytes), one
// 8-byte (bf16) / 16-byte (fp32) store.  v / 255 is an IEEE division: bit-identical to torch's `im /= 255` in fp32.
// bf16 fast path of pack_u8 (the production array source: W, Ws, left multiples of 4, 4-byte aligned source): a thread
// converts FOUR pixels - three aligned 32-bit loads (12 bytes), a 256-entry lookup of bf16(v / 255) built once per CTA with
// the same IEEE division (bit-identical to the per-pixel kernel, which spent most of its instructions on three
// divisions and 64-bit index arithmetic per pixel), one 32-byte store.  grid = (W/4 groups, H, B): no index divisions.
__global__ void __launch_bounds__(128)
pack_u8x4_kernel(const unsigned char* __restrict__ src, int Hs, int Ws, int top, int left, int H, int W, float pad, int swap_rb,
                 bf16* __restrict__ y) {
  __shared__ unsigned short lut[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    const bf16 v = __float2bfloat16_rn(__fdiv_rn((float)i, 255.0f));
    lut[i] = *reinterpret_cast<const unsigned short*>(&v);
  }
  __syncthreads();
  pdl_trigger();
  pdl_wait();
  const int g = blockIdx.x * blockDim.x + threadIdx.x;      // group of 4 pixels in the row
  if (g * 4 >= W) return;
  const int h = blockIdx.y, n = blockIdx.z;
  const int sy = h - top, sx = g * 4 - left;
  uint32_t o[8];
  if ((unsigned)sy < (unsigned)Hs && sx >= 0 && sx < Ws) {   // the whole group is inside (Ws, left multiples of 4)
    const uint32_t* q = reinterpret_cast<const uint32_t*>(src + (((long long)n * Hs + sy) * Ws + sx) * 3);
    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
    const uint32_t by[12] = {w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u, w0 >> 24, w1 & 255u, (w1 >> 8) & 255u,
                             (w1 >> 16) & 255u, w1 >> 24, w2 & 255u, (w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t a = lut[by[3 * i]], b = lut[by[3 * i + 1]], c = lut[by[3 * i + 2]];
      o[2 * i] = (swap_rb ? c : a) | (b << 16);
      o[2 * i + 1] = swap_rb ? a : c;
    }
  } else {
    const bf16 pv = __float2bfloat16_rn(__fdiv_rn(pad, 255.0f));
    const uint32_t pb = *reinterpret_cast<const unsigned short*>(&pv);
#pragma unroll
    for (int i = 0; i < 4; ++i) { o[2 * i] = pb | (pb << 16); o[2 * i + 1] = pb; }
  }
  bf16* dst = y + (((long long)n * H + h) * W + g * 4) * 4;
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
               "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
               : "memory");
}

template <typename T>
__global__ void pack_u8_kernel(const unsigned char* __restrict__ src, int B, int Hs, int Ws, int top, int left, int H, int W,
                               float pad, int swap_rb, T* __restrict__ y) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)B * H * W;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const int w = (int)(p % W);
  const long long t = p / W;
  const int h = (int)(t % H);
  const int n = (int)(t / H);
  const int sy = h - top, sx = w - left;
  float c0 = pad, c1 = pad, c2 = pad;
  if ((unsigned)sy < (unsigned)Hs && (unsigned)sx < (unsigned)Ws) {
    const unsigned char* q = src + (((long long)n * Hs + sy) * Ws + sx) * 3;
    const float a = (float)q[0], b = (float)q[1], c = (float)q[2];
    c0 = swap_rb ? c : a;
    c1 = b;
    c2 = swap_rb ? a : c;
  }
  c0 = __fdiv_rn(c0, 255.0f);
  c1 = __fdiv_rn(c1, 255.0f);
  c2 = __fdiv_rn(c2, 255.0f);
  if (sizeof(T) == 2) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(c0, c1), hi = __floats2bfloat162_rn(c2, 0.f);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&lo);
    o.y = *reinterpret_cast<const uint32_t*>(&hi);
    reinterpret_cast<uint2*>(y)[p] = o;
  } else {
    reinterpret_cast<float4*>(y)[p] = make_float4(c0, c1, c2, 0.f);
  }
}

// LetterBox with resize (data/augment.py:726-727: cv2.resize(img, new_unpad, INTER_LINEAR)) fused with the border, BGR->RGB,
// /255 and the NHWC-4 pack.  Bit-exact to OpenCV's 8-bit linear resize (imgproc/resize.cpp, HResizeLinear / VResizeLinear with
// 11-bit fixed-point coefficients): per output column xtab = {x0, a0, a1} (x1 = min(x0 + 1, ws - 1)), per output row
// ytab = {y0, y1, b0, b1}; the tables are built on the host exactly as OpenCV builds them (engine.resize_tables) and
// verified against cv2 in tests/test_preprocess.py.
//   h_r = S[y_r][x0] * a0 + S[y_r][x1] * a1;   out = ((b0 * (h_0 >> 4) >> 16) + (b1 * (h_1 >> 4) >> 16) + 2) >> 2
template <typename T>
__global__ void letterbox_u8_kernel(const unsigned char* __restrict__ src, int B, int hs, int ws, int nh, int nw,
                                    const int* __restrict__ xtab, const int* __restrict__ ytab, int top, int left, int H, int W,
                                    float pad, int swap_rb, T* __restrict__ y) {
  pdl_trigger();
  pdl_wait();
  const unsigned total = (unsigned)B * H * W;
  const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const unsigned t = p / (unsigned)W;
  const int w = (int)(p - t * (unsigned)W);
  const int n = (int)(t / (unsigned)H);
  const int h = (int)(t - (unsigned)n * (unsigned)H);
  const int dy = h - top, dx = w - left;
  float c0 = pad, c1 = pad, c2 = pad;
  if ((unsigned)dy < (unsigned)nh && (unsigned)dx < (unsigned)nw) {
    const int x0 = xtab[dx * 3], a0 = xtab[dx * 3 + 1], a1 = xtab[dx * 3 + 2];
    const int x1 = min(x0 + 1, ws - 1);
    const int y0 = ytab[dy * 4], y1 = ytab[dy * 4 + 1], b0 = ytab[dy * 4 + 2], b1 = ytab[dy * 4 + 3];
    const unsigned char* r0 = src + ((long long)n * hs + y0) * ws * 3;
    const unsigned char* r1 = src + ((long long)n * hs + y1) * ws * 3;
    int o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int h0 = (int)r0[x0 * 3 + c] * a0 + (int)r0[x1 * 3 + c] * a1;
      const int h1 = (int)r1[x0 * 3 + c] * a0 + (int)r1[x1 * 3 + c] * a1;
      const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
      o[c] = min(max(v, 0), 255);
    }
    c0 = (float)(swap_rb ? o[2] : o[0]);
    c1 = (float)o[1];
    c2 = (float)(swap_rb ? o[0] : o[2]);
  }
  c0 = __fdiv_rn(c0, 255.0f);
  c1 = __fdiv_rn(c1, 255.0f);
  c2 = __fdiv_rn(c2, 255.0f);
  if (sizeof(T) == 2) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(c0, c1), hi = __floats2bfloat162_rn(c2, 0.f);
    uint2 ov;
    ov.x = *reinterpret_cast<const uint32_t*>(&lo);
    ov.y = *reinterpret_cast<const uint32_t*>(&hi);
    reinterpret_cast<uint2*>(y)[p] = ov;
  } else {
    reinterpret_cast<float4*>(y)[p] = make_float4(c0, c1, c2, 0.f);
  }
}

// partial[b, chunk, c] = sum over the chunk's pixels (fixed order: deterministic, no atomics).  grid (chunks, B);
// 256 threads = 8 pixel lanes x 32 channel lanes looping over channel blocks.
template <typename T>
__global__ void global_avgpool_kernel(const T* __restrict__ x, int x_ld, int HW, int C, int chunk_pix, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  // 256 threads = PL pixel lanes x CV channel-vector lanes (16-byte loads); per channel block the pixel lanes are reduced
  // through shared memory in a fixed order (deterministic).
  constexpr int V = Vec<T>::N;
  __shared__ float part[256 * V];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int p0 = chunk * chunk_pix, p1 = min(HW, p0 + chunk_pix);
  const int cvecs = C / V;
  const int CV = cvecs < 32 ? cvecs : 32, PL = 256 / CV;
  const int cvl = threadIdx.x % CV, pl = threadIdx.x / CV;
  const T* base = x + (long long)b * HW * x_ld;
  for (int cb = 0; cb < cvecs; cb += CV) {
    const int cv = cb + cvl;
    float acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.f;
    if (cv < cvecs && pl < PL) {
      int p = p0 + pl;
      for (; p + 3 * PL < p1; p += 4 * PL) {          // four independent 16-byte loads in flight, summed in pixel order
        Vec<T> t0 = ldg_vec<T>(base + (long long)p * x_ld + cv * V);
        Vec<T> t1 = ldg_vec<T>(base + (long long)(p + PL) * x_ld + cv * V);
        Vec<T> t2 = ldg_vec<T>(base + (long long)(p + 2 * PL) * x_ld + cv * V);
        Vec<T> t3 = ldg_vec<T>(base + (long long)(p + 3 * PL) * x_ld + cv * V);
        float f0[V], f1[V], f2[V], f3[V];
        t0.unpack(f0); t1.unpack(f1); t2.unpack(f2); t3.unpack(f3);
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = (((acc[v] + f0[v]) + f1[v]) + f2[v]) + f3[v];
      }
      for (; p < p1; p += PL) {
        float f[V];
        ldg_vec<T>(base + (long long)p * x_ld + cv * V).unpack(f);
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] += f[v];
      }
    }
#pragma unroll
    for (int v = 0; v < V; ++v) part[threadIdx.x * V + v] = acc[v];
    __syncthreads();
    for (int i = threadIdx.x; i < CV * V; i += 256) {
      const int c = cb * V + i;
      if (c < C) {
        float t = 0.f;
        for (int q = 0; q < PL; ++q) t += part[(q * CV) * V + i];
        out[((long long)b * gridDim.x + chunk) * C + c] = t;
      }
    }
    __syncthreads();
  }
}

// tiny per-image MLP on pooled vectors; one block per image.  A warp owns an output: its lanes stride over the weight row
// (coalesced) and reduce with shuffles - the per-thread serial dot product it replaces walked the rows with a stride of C0
// floats between threads and was pure load latency (7-13 us for a few KFLOP, profiles/r01_k_per_launch_lpc_b64.csv).
__global__ void channel_mlp_kernel(const float* __restrict__ in, int parts, float in_scale, int C0, const float* __restrict__ W1,
                                   const float* __restrict__ b1, int C1, int act1, const float* __restrict__ W2,
                                   const float* __restrict__ b2, int C2, int act2, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  float* v0 = sm;
  float* v1 = sm + C0;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < C0; i += blockDim.x) {
    float t = 0.f;
    for (int q = 0; q < parts; ++q) t += in[((long long)b * parts + q) * C0 + i];   // fixed order: deterministic
    v0[i] = t * in_scale;
  }
  __syncthreads();
  for (int o = warp; o < C1; o += nwarps) {
    const float* wr = W1 + (long long)o * C0;
    float s = 0.f;
    for (int i = lane; i < C0; i += 32) s = fmaf(__ldg(wr + i), v0[i], s);
#pragma unroll
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) {
      s = apply_act<true>(s + (b1 ? b1[o] : 0.f), act1);
      if (W2) v1[o] = s; else out[(long long)b * C1 + o] = s;
    }
  }
  if (!W2) return;
  __syncthreads();
  if (C1 <= 32) {        // short hidden vector (C/16): a thread per output, the row is a handful of floats
    for (int o = threadIdx.x; o < C2; o += blockDim.x) {
      float s = b2 ? b2[o] : 0.f;
      const float* wr = W2 + (long long)o * C1;
      for (int i = 0; i < C1; ++i) s = fmaf(__ldg(wr + i), v1[i], s);
      out[(long long)b * C2 + o] = apply_act<true>(s, act2);
    }
  } else {
    for (int o = warp; o < C2; o += nwarps) {
      const float* wr = W2 + (long long)o * C1;
      float s = 0.f;
      for (int i = lane; i < C1; i += 32) s = fmaf(__ldg(wr + i), v1[i], s);
#pragma unroll
      for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
      if (lane == 0) out[(long long)b * C2 + o] = apply_act<true>(s + (b2 ? b2[o] : 0.f), act2);
    }
  }
}

// LPP lanes (a power of two <= 32) share one pixel, each owning 16-byte channel vectors: mean_c and max_c of x*ca.
// grid (chunks, B): a CTA stays inside one image, so a lane's channel-attention weights ca[b][c] are loaded ONCE into
// registers (the first version re-loaded 2 x 16 bytes of ca for every 16 bytes of x: the LSU, not HBM, set its speed);
// 4 pixels are in flight per lane group.
template <typename T>
__global__ void __launch_bounds__(256)
cbam_stats_kernel(const T* __restrict__ x, int x_ld, int B, int HW, int C, int lpp, const float* __restrict__ ca,
                  float* __restrict__ stats) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = Vec<T>::N;
  const int b = blockIdx.y;
  const int sub = threadIdx.x & (lpp - 1);
  const int grp = threadIdx.x / lpp, ngrp = 256 / lpp;
  const int chunk_pix = (HW + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * chunk_pix, p1 = min(HW, p0 + chunk_pix);
  const T* img = x + (long long)b * HW * x_ld;
  const float* cab = ca + (long long)b * C;
  float2* so = reinterpret_cast<float2*>(stats) + (long long)b * HW;
  const bool single = C <= lpp * V;          // one vector per lane (C <= 256 in bf16): ca lives in registers
  float g0[V];
  if (single) {
#pragma unroll
    for (int v = 0; v < V; v += 4) {
      const float4 c4 = sub * V < C ? __ldg(reinterpret_cast<const float4*>(cab + sub * V + v)) : make_float4(0.f, 0.f, 0.f, 0.f);
      g0[v] = c4.x; g0[v + 1] = c4.y; g0[v + 2] = c4.z; g0[v + 3] = c4.w;
    }
  }
  constexpr int U = 4;
  for (int pb = p0 + grp; pb < p1; pb += ngrp * U) {
    float s[U], m[U];
    if (single) {
      Vec<T> xv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = pb + u * ngrp;
        xv[u].raw = make_uint4(0, 0, 0, 0);
        if (p < p1 && sub * V < C) xv[u] = ldg_vec<T>(img + (long long)p * x_ld + sub * V);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float f[V];
        xv[u].unpack(f);
        s[u] = 0.f;
        m[u] = -INFINITY;
        if (sub * V < C) {
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const float t = f[v] * g0[v];
            s[u] += t;
            m[u] = fmaxf(m[u], t);
          }
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = pb + u * ngrp;
        s[u] = 0.f;
        m[u] = -INFINITY;
        if (p < p1)
          for (int c = sub * V; c < C; c += lpp * V) {
            float f[V];
            ldg_vec<T>(img + (long long)p * x_ld + c).unpack(f);
#pragma unroll
            for (int v = 0; v < V; ++v) {
              const float t = f[v] * __ldg(cab + c + v);
              s[u] += t;
              m[u] = fmaxf(m[u], t);
            }
          }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      for (int o = lpp >> 1; o; o >>= 1) {
        s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
        m[u] = fmaxf(m[u], __shfl_xor_sync(0xffffffffu, m[u], o));
      }
      const int p = pb + u * ngrp;
      if (p < p1 && sub == 0) so[p] = make_float2(s[u] / (float)C, m[u]);
    }
  }
}

// gate = sigmoid(conv_kxk([mean_c, max_c])) ; y = x * ca * gate (conv.py:300-320).  CTA = one 16 x 16 pixel tile of one
// image: the stats halo ((16+k-1)^2 float2), the 2*k*k filter taps and ca[b][:] are staged in shared memory, each
// thread computes the gate of one pixel (2*k*k FMAs on shared memory), then the 256 threads stream the tile's channel
// vectors (16-byte loads / stores).  The earlier version split the taps over the lanes of a pixel with scattered 8-byte
// global loads and was 5x above its HBM floor (profiles/r01_j_per_launch_lpc_b64.csv).
constexpr int CB_T = 16;
template <typename T>
__global__ void __launch_bounds__(CB_T * CB_T)
cbam_apply_kernel(const T* __restrict__ x, int x_ld, int H, int W, int C, const float* __restrict__ ca,
                  const float* __restrict__ stats, const float* __restrict__ w, int k, T* __restrict__ y, int y_ld) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = Vec<T>::N;
  extern __shared__ float cb_sm[];
  const int pad = k / 2, hw = CB_T + k - 1;
  float2* st = reinterpret_cast<float2*>(cb_sm);          // [hw][hw]
  float* wsm = cb_sm + 2 * hw * hw;                        // [2][k*k]
  float* casm = wsm + 2 * k * k;                           // [C]
  float* gate = casm + C;                                  // [256]
  const int b = blockIdx.z, y0 = blockIdx.y * CB_T, x0 = blockIdx.x * CB_T;
  const int tid = threadIdx.x;
  const float2* sb = reinterpret_cast<const float2*>(stats) + (long long)b * H * W;
  for (int i = tid; i < hw * hw; i += CB_T * CB_T) {
    const int r = i / hw, c = i - r * hw;
    const int iy = y0 - pad + r, ix = x0 - pad + c;
    st[i] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(sb + (long long)iy * W + ix) : make_float2(0.f, 0.f);
  }
  for (int i = tid; i < 2 * k * k; i += CB_T * CB_T) wsm[i] = w[i];
  for (int i = tid; i < C; i += CB_T * CB_T) casm[i] = ca[(long long)b * C + i];
  __syncthreads();
  {
    const int ty = tid / CB_T, tx = tid - ty * CB_T;
    float s = 0.f;
    for (int ky = 0; ky < k; ++ky)
      for (int kx = 0; kx < k; ++kx) {
        const float2 v = st[(ty + ky) * hw + tx + kx];
        s = fmaf(v.x, wsm[ky * k + kx], s);
        s = fmaf(v.y, wsm[k * k + ky * k + kx], s);
      }
    gate[tid] = sigmoid_<true>(s);
  }
  __syncthreads();
  const int cvecs = C / V;
  if ((CB_T * CB_T) % cvecs == 0) {
    // the lane's channel vector is the same in every round: its ca weights stay in registers, and the pixel index
    // advances by a constant (no division in the loop)
    const int cv = tid % cvecs, ppr = CB_T * CB_T / cvecs;
    float cg[V];
#pragma unroll
    for (int v = 0; v < V; ++v) cg[v] = casm[cv * V + v];
    for (int pix = tid / cvecs; pix < CB_T * CB_T; pix += ppr) {
      const int ty = pix / CB_T, tx = pix % CB_T;
      const int oy = y0 + ty, ox = x0 + tx;
      if (oy >= H || ox >= W) continue;
      const long long gp = ((long long)b * H + oy) * W + ox;
      float f[V];
      ldg_vec<T>(x + gp * x_ld + cv * V).unpack(f);
      const float g = gate[pix];
#pragma unroll
      for (int v = 0; v < V; ++v) f[v] *= cg[v] * g;
      Vec<T> o;
      o.pack(f);
      st_vec<T>(y + gp * y_ld + cv * V, o);
    }
    return;
  }
  for (int i = tid; i < CB_T * CB_T * cvecs; i += CB_T * CB_T) {
    const int pix = i / cvecs, cv = i - pix * cvecs;
    const int ty = pix / CB_T, tx = pix - ty * CB_T;
    const int oy = y0 + ty, ox = x0 + tx;
    if (oy >= H || ox >= W) continue;
    const long long gp = ((long long)b * H + oy) * W + ox;
    float f[V];
    ldg_vec<T>(x + gp * x_ld + cv * V).unpack(f);
    const float g = gate[pix];
#pragma unroll
    for (int v = 0; v < V; ++v) f[v] *= casm[cv * V + v] * g;
    Vec<T> o;
    o.pack(f);
    st_vec<T>(y + gp * y_ld + cv * V, o);
  }
}

template <typename T> int vec_of() { return Vec<T>::N; }

}  // namespace

#define DISPATCH_T(dtype, CALL_F32, CALL_BF16, name)                   \
  if ((dtype) == LPC_F32) { CALL_F32; }                                \
  else if ((dtype) == LPC_BF16) { CALL_BF16; }                         \
  else LPC_FAIL(LPC_E_ARG, name ": unknown dtype %d", (dtype));        \
  LPC_CHECK_LAUNCH(name);                                              \
  return LPC_OK;

static int check_vec(const char* name, int dtype, int C, int x_ld, int y_ld, const void* x, const void* y) {
  const int V = dtype == LPC_F32 ? 4 : 8;
  if (C % V || x_ld % V || y_ld % V) LPC_FAIL(LPC_E_ARG, "%s: C and pitches must be multiples of %d", name, V);
  if (!aligned16(x) || !aligned16(y)) LPC_FAIL(LPC_E_ARG, "%s: pointers must be 16-byte aligned", name);
  if (!x || !y) LPC_FAIL(LPC_E_ARG, "%s: null pointer", name);
  return LPC_OK;
}

extern "C" int lpc_upsample2x(int dtype, const void* x, int x_ld, int B, int H, int W, int C, void* y, int y_ld, void* stream) {
  if (int e = check_vec("upsample2x", dtype, C, x_ld, y_ld, x, y)) return e;
  cudaStream_t s = (cudaStream_t)stream;
  const int V = dtype == LPC_F32 ? 4 : 8;
  LPC_REQUIRE((long long)B * H * W * (C / V) < (1ll << 32), "upsample2x: tensor too large for 32-bit indexing");
  const int g = cdiv((long long)B * H * W * (C / V), 256);
  DISPATCH_T(dtype, (lpc_launch_pdl(upsample2x_kernel<float>, g, 256, 0, s, (const float*)x, x_ld, B, H, W, C, (float*)y, y_ld)),
             (lpc_launch_pdl(upsample2x_kernel<bf16>, g, 256, 0, s, (const bf16*)x, x_ld, B, H, W, C, (bf16*)y, y_ld)), "upsample2x")
}

extern "C" int lpc_copy_channels(int dtype, const void* x, int x_ld, long long npix, int C, void* y, int y_ld, void* stream) {
  if (int e = check_vec("copy_channels", dtype, C, x_ld, y_ld, x, y)) return e;
  cudaStream_t s = (cudaStream_t)stream;
  const int V = dtype == LPC_F32 ? 4 : 8;
  const int g = cdiv(npix * (C / V), 256);
  DISPATCH_T(dtype, (lpc_launch_pdl(copy_channels_kernel<float>, g, 256, 0, s, (const float*)x, x_ld, npix, C, (float*)y, y_ld)),
             (lpc_launch_pdl(copy_channels_kernel<bf16>, g, 256, 0, s, (const bf16*)x, x_ld, npix, C, (bf16*)y, y_ld)), "copy_channels")
}

extern "C" int lpc_space_to_depth(int dtype, const void* x, int x_ld, int B, int H, int W, int C, void* y, int y_ld, void* stream) {
  if (int e = check_vec("space_to_depth", dtype, C, x_ld, y_ld, x, y)) return e;
  LPC_REQUIRE(H % 2 == 0 && W % 2 == 0 && y_ld >= 4 * C, "space_to_depth: H, W must be even and y_ld >= 4C");
  cudaStream_t s = (cudaStream_t)stream;
  const int V = dtype == LPC_F32 ? 4 : 8;
  const int g = cdiv((long long)B * H * W * (C / V), 256);
  DISPATCH_T(dtype, (lpc_launch_pdl(space_to_depth_kernel<float>, g, 256, 0, s, (const float*)x, x_ld, B, H, W, C, (float*)y, y_ld)),
             (lpc_launch_pdl(space_to_depth_kernel<bf16>, g, 256, 0, s, (const bf16*)x, x_ld, B, H, W, C, (bf16*)y, y_ld)), "space_to_depth")
}

extern "C" int lpc_channel_deinterleave(int dtype, const void* x, int x_ld, long long npix, int C, void* y, int y_ld, void* stream) {
  LPC_REQUIRE(x && y && C % 2 == 0 && x_ld >= C && y_ld >= C, "channel_deinterleave: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  {
    const int V = dtype == LPC_F32 ? 4 : 8;
    if (C % (2 * V) == 0 && x_ld % V == 0 && y_ld % V == 0 && aligned16(x) && aligned16(y) && npix * (C / (2 * V)) < (1ll << 32) &&
        (dtype == LPC_F32 || dtype == LPC_BF16)) {
      const int gv = cdiv(npix * (C / (2 * V)), 256);
      DISPATCH_T(dtype, (lpc_launch_pdl(deinterleave_vec_kernel<float>, gv, 256, 0, s, (const float*)x, x_ld, (unsigned)npix, C, (float*)y, y_ld)),
                 (lpc_launch_pdl(deinterleave_vec_kernel<bf16>, gv, 256, 0, s, (const bf16*)x, x_ld, (unsigned)npix, C, (bf16*)y, y_ld)), "channel_deinterleave")
    }
  }
  const int g = cdiv(npix * C, 256);
  DISPATCH_T(dtype, (lpc_launch_pdl(deinterleave_kernel<float>, g, 256, 0, s, (const float*)x, x_ld, npix, C, (float*)y, y_ld)),
             (lpc_launch_pdl(deinterleave_kernel<bf16>, g, 256, 0, s, (const bf16*)x, x_ld, npix, C, (bf16*)y, y_ld)), "channel_deinterleave")
}

extern "C" int lpc_pack_input(int dtype, const float* x, int B, int C, int H, int W, void* y, int y_ld, int Cpad, void* stream) {
  LPC_REQUIRE(x && y && C > 0 && Cpad >= C && y_ld >= Cpad, "pack_input: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  const int g = cdiv((long long)B * H * W, 256);
  DISPATCH_T(dtype, (lpc_launch_pdl(pack_input_kernel<float>, g, 256, 0, s, x, B, C, H, W, (float*)y, y_ld, Cpad)),
             (lpc_launch_pdl(pack_input_kernel<bf16>, g, 256, 0, s, x, B, C, H, W, (bf16*)y, y_ld, Cpad)), "pack_input")
}

extern "C" int lpc_pack_u8(int dtype, const void* src, int B, int Hs, int Ws, int top, int left, int H, int W, int pad_value,
                           int swap_rb, void* y, void* stream) {
  LPC_REQUIRE(src && y && B > 0 && Hs > 0 && Ws > 0 && H >= Hs && W >= Ws, "pack_u8: bad argument");
  LPC_REQUIRE(top >= 0 && left >= 0 && top + Hs <= H && left + Ws <= W, "pack_u8: the image does not fit at (%d, %d)", top, left);
  LPC_REQUIRE(pad_value >= 0 && pad_value <= 255, "pack_u8: pad value must be a byte");
  LPC_REQUIRE(aligned16(y), "pack_u8: output must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == LPC_BF16 && W % 4 == 0 && Ws % 4 == 0 && left % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0 &&
      (reinterpret_cast<uintptr_t>(y) & 31) == 0 && H <= 65535 && B <= 65535) {
    dim3 grid(cdiv(W / 4, 128), H, B);
    lpc_launch_pdl(pack_u8x4_kernel, grid, 128, 0, s, (const unsigned char*)src, Hs, Ws, top, left, H, W, (float)pad_value, swap_rb, (bf16*)y);
    LPC_CHECK_LAUNCH("pack_u8");
    return LPC_OK;
  }
  const int g = cdiv((long long)B * H * W, 256);
  DISPATCH_T(dtype, (lpc_launch_pdl(pack_u8_kernel<float>, g, 256, 0, s, (const unsigned char*)src, B, Hs, Ws, top, left, H, W, (float)pad_value, swap_rb, (float*)y)),
             (lpc_launch_pdl(pack_u8_kernel<bf16>, g, 256, 0, s, (const unsigned char*)src, B, Hs, Ws, top, left, H, W, (float)pad_value, swap_rb, (bf16*)y)), "pack_u8")
}

extern "C" int lpc_letterbox_u8(int dtype, const void* src, int B, int hs, int ws, int nh, int nw, const int* xtab, const int* ytab,
                                int top, int left, int H, int W, int pad_value, int swap_rb, void* y, void* stream) {
  LPC_REQUIRE(src && y && xtab && ytab && B > 0 && hs > 0 && ws > 0 && nh > 0 && nw > 0, "letterbox_u8: bad argument");
  LPC_REQUIRE(top >= 0 && left >= 0 && top + nh <= H && left + nw <= W, "letterbox_u8: the resized image does not fit at (%d, %d)", top, left);
  LPC_REQUIRE(pad_value >= 0 && pad_value <= 255, "letterbox_u8: pad value must be a byte");
  LPC_REQUIRE(aligned16(y) && (long long)B * H * W < (1ll << 32), "letterbox_u8: output alignment / size");
  cudaStream_t s = (cudaStream_t)stream;
  const int g = cdiv((long long)B * H * W, 256);
  DISPATCH_T(dtype, (lpc_launch_pdl(letterbox_u8_kernel<float>, g, 256, 0, s, (const unsigned char*)src, B, hs, ws, nh, nw, xtab, ytab, top, left, H, W, (float)pad_value, swap_rb, (float*)y)),
             (lpc_launch_pdl(letterbox_u8_kernel<bf16>, g, 256, 0, s, (const unsigned char*)src, B, hs, ws, nh, nw, xtab, ytab, top, left, H, W, (float)pad_value, swap_rb, (bf16*)y)), "letterbox_u8")
}

extern "C" int lpc_global_avgpool_chunks(int B, int HW) {
  int chunks = (592 + B - 1) / B;             // ~4 CTAs per SM in total
  const int max_chunks = (HW + 63) / 64;      // at least 64 pixels per chunk
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks > 64) chunks = 64;
  return chunks < 1 ? 1 : chunks;
}

extern "C" int lpc_global_avgpool(int dtype, const void* x, int x_ld, int B, int HW, int C, float* partial, void* stream) {
  LPC_REQUIRE(x && partial && B > 0 && HW > 0 && C > 0 && x_ld >= C, "global_avgpool: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  const int chunks = lpc_global_avgpool_chunks(B, HW);
  const int chunk_pix = (HW + chunks - 1) / chunks;
  dim3 g(chunks, B);
  DISPATCH_T(dtype, (lpc_launch_pdl(global_avgpool_kernel<float>, g, 256, 0, s, (const float*)x, x_ld, HW, C, chunk_pix, partial)),
             (lpc_launch_pdl(global_avgpool_kernel<bf16>, g, 256, 0, s, (const bf16*)x, x_ld, HW, C, chunk_pix, partial)), "global_avgpool")
}

extern "C" int lpc_channel_mlp(const float* in, int B, int parts, float in_scale, int C0, const float* W1, const float* b1, int C1,
                               int act1, const float* W2, const float* b2, int C2, int act2, float* out, void* stream) {
  LPC_REQUIRE(in && W1 && out && B > 0 && C0 > 0 && C1 > 0 && parts > 0, "channel_mlp: bad argument");
  LPC_REQUIRE((size_t)(C0 + C1) * 4 <= 48 * 1024, "channel_mlp: vectors too large");
  lpc_launch_pdl(channel_mlp_kernel, B, 256, (size_t)(C0 + C1) * 4, (cudaStream_t)stream, in, parts, in_scale, C0, W1, b1, C1, act1, W2, b2, C2, act2, out);
  LPC_CHECK_LAUNCH("channel_mlp");
  return LPC_OK;
}

static int lanes_per_pixel(int dtype, int C) {
  const int V = dtype == LPC_F32 ? 4 : 8;
  int l = 1;
  while (l * 2 <= 32 && l * 2 * V <= C) l *= 2;
  return l;
}

extern "C" int lpc_cbam_stats(int dtype, const void* x, int x_ld, int B, int HW, int C, const float* ca, float* stats, void* stream) {
  LPC_REQUIRE(x && ca && stats && x_ld >= C, "cbam_stats: bad argument");
  if (int e = check_vec("cbam_stats", dtype, C, x_ld, x_ld, x, x)) return e;
  cudaStream_t s = (cudaStream_t)stream;
  const int lpp = lanes_per_pixel(dtype, C);
  LPC_REQUIRE(B <= 65535, "cbam_stats: batch too large");
  // ~8 CTAs per SM in total, at least 128 pixels (4 rounds of the 4-pixel unroll for 8 lane groups) per CTA
  int chunks = (8 * 148 + B - 1) / B;
  const int max_chunks = (HW + 127) / 128;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  dim3 g(chunks, B);
  DISPATCH_T(dtype, (lpc_launch_pdl(cbam_stats_kernel<float>, g, 256, 0, s, (const float*)x, x_ld, B, HW, C, lpp, ca, stats)),
             (lpc_launch_pdl(cbam_stats_kernel<bf16>, g, 256, 0, s, (const bf16*)x, x_ld, B, HW, C, lpp, ca, stats)), "cbam_stats")
}

extern "C" int lpc_cbam_apply(int dtype, const void* x, int x_ld, int B, int H, int W, int C, const float* ca,
                              const float* stats, const float* w, int k, void* y, int y_ld, void* stream) {
  LPC_REQUIRE(x && ca && stats && w && y && (k == 3 || k == 7), "cbam_apply: bad argument");
  if (int e = check_vec("cbam_apply", dtype, C, x_ld, y_ld, x, y)) return e;
  cudaStream_t s = (cudaStream_t)stream;
  dim3 g(cdiv(W, CB_T), cdiv(H, CB_T), B);
  LPC_REQUIRE(B <= 65535 && cdiv(H, CB_T) <= 65535, "cbam_apply: shape too large");
  const int hw = CB_T + k - 1;
  const size_t smem = sizeof(float) * (size_t)(2 * hw * hw + 2 * k * k + C + CB_T * CB_T);
  LPC_REQUIRE(smem <= 48 * 1024, "cbam_apply: C too large for the shared-memory gate kernel");
  DISPATCH_T(dtype, (lpc_launch_pdl(cbam_apply_kernel<float>, g, CB_T * CB_T, smem, s, (const float*)x, x_ld, H, W, C, ca, stats, w, k, (float*)y, y_ld)),
             (lpc_launch_pdl(cbam_apply_kernel<bf16>, g, CB_T * CB_T, smem, s, (const bf16*)x, x_ld, H, W, C, ca, stats, w, k, (bf16*)y, y_ld)), "cbam_apply")
}

// ---- small box utilities behind the reference's utils/tal.py / utils/ops.py function API -----------------------------
// (the engine never calls these: the fused tail computes anchors, dist2bbox, xywh2xyxy, scale_boxes and clip_boxes itself)
namespace {

// anchors[a] = (x + offset, y + offset), stride_out[a] = stride of a's level; a = level offset + y * W_l + x
__global__ void make_anchors_kernel(int n_levels, int4 starts, int4 widths, float4 strides, float offset, int A,
                                    float* __restrict__ anchors, float* __restrict__ stride_out) {
  pdl_trigger();
  pdl_wait();
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= A) return;
  const int st[4] = {starts.x, starts.y, starts.z, starts.w};
  const int wd[4] = {widths.x, widths.y, widths.z, widths.w};
  const float sv[4] = {strides.x, strides.y, strides.z, strides.w};
  int l = 0;
  for (int i = 1; i < n_levels; ++i)
    if (a >= st[i]) l = i;
  const int cell = a - st[l];
  const int y = cell / wd[l], x = cell - y * wd[l];
  anchors[2 * a] = (float)x + offset;
  anchors[2 * a + 1] = (float)y + offset;
  stride_out[a] = sv[l];
}

// rows of (l, t, r, b) distances + (ax, ay) anchor points -> (cx, cy, w, h) or (x1, y1, x2, y2)
__global__ void dist2bbox_kernel(const float4* __restrict__ dist, const float2* __restrict__ anchors, long long n, long long anchor_mod, int xywh,
                                 float4* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 d = dist[i];
  const float2 a = anchors[i % anchor_mod];
  const float x1 = a.x - d.x, y1 = a.y - d.y, x2 = a.x + d.z, y2 = a.y + d.w;
  out[i] = xywh ? make_float4((x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1) : make_float4(x1, y1, x2, y2);
}

// in place on rows whose first four floats are a box: optional xywh -> xyxy, minus padding, / gain, optional clip
__global__ void scale_boxes_kernel(float* __restrict__ boxes, long long n, int row_stride, int xywh_in, float pad_x, float pad_y, float gain,
                                   float clip_w, float clip_h) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* b = boxes + i * row_stride;
  float x1 = b[0], y1 = b[1], x2 = b[2], y2 = b[3];
  if (xywh_in) {
    const float dw = x2 / 2, dh = y2 / 2, cx = x1, cy = y1;
    x1 = cx - dw; y1 = cy - dh; x2 = cx + dw; y2 = cy + dh;
  }
  x1 -= pad_x; x2 -= pad_x; y1 -= pad_y; y2 -= pad_y;
  if (gain != 1.0f) { x1 /= gain; y1 /= gain; x2 /= gain; y2 /= gain; }
  if (clip_w > 0.f) {
    x1 = fminf(fmaxf(x1, 0.f), clip_w); x2 = fminf(fmaxf(x2, 0.f), clip_w);
    y1 = fminf(fmaxf(y1, 0.f), clip_h); y2 = fminf(fmaxf(y2, 0.f), clip_h);
  }
  b[0] = x1; b[1] = y1; b[2] = x2; b[3] = y2;
}

}  // namespace

extern "C" int lpc_make_anchors(int n_levels, const int* hw_host, const float* strides_host, float offset, float* anchors,
                                float* stride_out, void* stream) {
  LPC_REQUIRE(n_levels >= 1 && n_levels <= 4 && hw_host && strides_host && anchors && stride_out, "make_anchors: bad argument (1..4 levels)");
  int st[4] = {0, 0, 0, 0}, wd[4] = {1, 1, 1, 1};
  float sv[4] = {0, 0, 0, 0};
  int A = 0;
  for (int l = 0; l < n_levels; ++l) {
    LPC_REQUIRE(hw_host[2 * l] > 0 && hw_host[2 * l + 1] > 0, "make_anchors: empty level");
    st[l] = A; wd[l] = hw_host[2 * l + 1]; sv[l] = strides_host[l];
    A += hw_host[2 * l] * hw_host[2 * l + 1];
  }
  lpc_launch_pdl(make_anchors_kernel, dim3(cdiv(A, 256)), dim3(256), 0, (cudaStream_t)stream, n_levels, make_int4(st[0], st[1], st[2], st[3]),
                 make_int4(wd[0], wd[1], wd[2], wd[3]), make_float4(sv[0], sv[1], sv[2], sv[3]), offset, A, anchors, stride_out);
  LPC_CHECK_LAUNCH("make_anchors");
  return LPC_OK;
}

extern "C" int lpc_dist2bbox(const float* dist, const float* anchors, long long n, long long n_anchors, int xywh, float* out, void* stream) {
  LPC_REQUIRE(dist && anchors && out && n > 0 && n_anchors > 0 && n % n_anchors == 0, "dist2bbox: bad argument");
  LPC_REQUIRE(aligned16(dist) && aligned16(out) && (reinterpret_cast<uintptr_t>(anchors) & 7) == 0, "dist2bbox: rows must be 16-byte aligned");
  lpc_launch_pdl(dist2bbox_kernel, dim3(cdiv(n, 256)), dim3(256), 0, (cudaStream_t)stream, (const float4*)dist, (const float2*)anchors, n, n_anchors, xywh,
                 (float4*)out);
  LPC_CHECK_LAUNCH("dist2bbox");
  return LPC_OK;
}

extern "C" int lpc_scale_boxes(float* boxes, long long n, int row_stride, int xywh_in, float pad_x, float pad_y, float gain, float clip_w,
                               float clip_h, void* stream) {
  LPC_REQUIRE(boxes && n > 0 && row_stride >= 4 && gain > 0.f, "scale_boxes: bad argument");
  lpc_launch_pdl(scale_boxes_kernel, dim3(cdiv(n, 256)), dim3(256), 0, (cudaStream_t)stream, boxes, n, row_stride, xywh_in, pad_x, pad_y, gain, clip_w, clip_h);
  LPC_CHECK_LAUNCH("scale_boxes");
  return LPC_OK;
}
