/*
 * lpcyolo.h - C ABI of liblpcyolo.so: the B200 (sm_100a) kernels behind the LPC-YOLO / YOLOv10
 * inference hot path.
 *
 * The reference (a fork of ultralytics 8.1.34) has no FFI / plugin boundary of its own: the path is
 * pure Python on torch ATen.  This header therefore *defines* the boundary:
 * each entry point below replaces one group of ATen calls made by a reference module and cites it
 * (paths relative to the reference's ultralytics/ directory).  The Python host code in
 * lpc-yolo_b200/ binds these with ctypes (lpc-yolo_b200/_lib.py); INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer owned by the caller.  The library never allocates device
 *     memory, never synchronises and keeps no global state besides a per-process TMA descriptor
 *     cache and the last-error string (thread local).
 *   - Activations are NHWC "views": element (n,y,x,c) of a view with pixel pitch `ld` lives at
 *     ptr[((n*H + y)*W + x)*ld + c].  ld >= C lets a layer read or write a channel slice of a wider
 *     buffer, which is how chunk/cat (block.py:229-231), Concat (conv.py:331) and the detect head's
 *     cat (head.py:76) are made free.
 *   - dtype: LPC_BF16 (storage bf16, fp32 accumulate; the production mode) or LPC_F32 (fp32 storage
 *     and arithmetic on CUDA cores; the validation mode, 1e-5 of the reference).
 *   - `stream` is a cudaStream_t passed as void*.
 *   - Return value: 0 on success, negative LPC_E_* otherwise; lpc_last_error() gives the text.
 */
#ifndef LPCYOLO_H_
#define LPCYOLO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LPC_ABI_VERSION 1

enum { LPC_BF16 = 0, LPC_F32 = 1 };
/* per-layer activation, read from the reference module's actual `.act` (SURVEY.md finding 1):
 * conv.Conv -> SiLU (conv.py:39), block.Conv -> Mish (block.py:4920), act=False -> identity,
 * ChannelAttention / SPCA gates -> sigmoid (conv.py:286, block.py:5740), SPCA squeeze -> ReLU. */
enum { LPC_ACT_NONE = 0, LPC_ACT_SILU = 1, LPC_ACT_MISH = 2, LPC_ACT_SIGMOID = 3, LPC_ACT_RELU = 4 };
enum {
  LPC_OK = 0,
  LPC_E_ARG = -1,         /* bad shape / alignment / null pointer */
  LPC_E_UNSUPPORTED = -2, /* valid request the kernels do not cover */
  LPC_E_CUDA = -3,        /* a CUDA runtime / driver call failed */
  LPC_E_WORKSPACE = -4    /* workspace too small */
};

int lpc_abi_version(void);
const char* lpc_last_error(void);
/* compute capability of the current device as major*10+minor (100 on B200); <0 on error. */
int lpc_device_arch(void);
/* number of kernels this library has launched in this process (bench.py reports it as gpu_launches). */
unsigned long long lpc_launch_count(void);

/* ---- dense convolution -------------------------------------------------------------------------
 * y = act(conv2d(x, w) + bias) [* chan_scale[n,c]] [+ res]
 * replaces nn.Conv2d + BatchNorm2d (folded on the host: utils/torch_utils.py:171-198) + activation in
 * conv.Conv.forward (conv.py:48-54) and block.Conv.forward (block.py:4922-4926); `res` is the
 * Bottleneck / PSA / SPCA shortcut add (block.py:340, :813-814, :5749); `chan_scale` ([B,Cout] fp32) is
 * SPCA's `spatial * attn` gate (block.py:5748).
 *
 * lpc_conv2d_direct: CUDA-core implicit GEMM, any k/stride/pad, both dtypes.
 *   w: [k*k][Cin][Cout] in the activation dtype; bias: [Cout] fp32 or NULL.
 * lpc_conv2d_tc: tcgen05 / TMEM implicit GEMM fed by TMA, bf16 only, k in {1,2,3}, stride in {1,2}
 *   (k=2 requires stride 2 pad 0: the space_to_depth + 1x1 fold, block.py:4069 + :222).
 *   w: [Cout][Kpad] bf16, K-major, k index = tap*Cin + cin, zero padded to Kpad (lpc_conv2d_tc_kpad).
 *   Requires Cin % 16 == 0, Cout % 16 == 0, x_ld % 8 == 0, 16-byte aligned pointers.
 */
int lpc_conv2d_direct(int dtype, const void* x, int x_ld, int B, int H, int W, int Cin,
                      const void* w, const float* bias, int k, int stride, int pad, int Cout,
                      void* y, int y_ld, int act, const float* chan_scale,
                      const void* res, int res_ld, void* stream);
int lpc_conv2d_tc(const void* x, int x_ld, int B, int H, int W, int Cin,
                  const void* w, const float* bias, int k, int stride, int pad, int Cout,
                  void* y, int y_ld, int act, const float* chan_scale,
                  const void* res, int res_ld, void* stream);
/* K padding rule of the tcgen05 weight layout; returns Kpad (multiple of 64) or <0. */
int lpc_conv2d_tc_kpad(int Cin, int k);
/* Kernel selection override for tests / profiling: 0 = auto, 1 = always the per-tap TMA kernel, 2 = the halo-patch
 * kernel for every 3x3 stride-1 conv whose buffers fit.  Returns the previous mode. */
int lpc_conv2d_tc_set_mode(int mode);
/* lpc_conv2d_tc that ALSO writes, per output pixel, the order-preserving uint32 key of max over the output channels of
 * the (bf16-rounded) outputs: keys[b * img_stride + offset + oy * Wo + ox].  Used on the last conv of the v10Detect class
 * branch (head.py:504-505) so that stage 1 of ops.v10postprocess (utils/ops.py:853, `scores.amax(-1)`) costs no extra
 * pass over the class logits.  Returns LPC_E_UNSUPPORTED (nothing launched) when Cout is split over several N tiles or
 * epilogue warps; rowmax_keys == NULL behaves exactly like lpc_conv2d_tc. */
int lpc_conv2d_tc_rowmax(const void* x, int x_ld, int B, int H, int W, int Cin, const void* w, const float* bias,
                         int k, int stride, int pad, int Cout, void* y, int y_ld, int act,
                         const float* chan_scale, const void* res, int res_ld,
                         unsigned int* rowmax_keys, long long rowmax_img_stride, int rowmax_offset, void* stream);
/* 1 if lpc_conv2d_tc accepts this shape, else 0. */
int lpc_conv2d_tc_supported(int Cin, int Cout, int k, int stride, int pad, int x_ld, int y_ld);

/* 1x1 conv over cat[upsample2x(x_small), x_skip] (the neck's nn.Upsample -> Concat -> C2f.cv1, LPC YAML layers 15-17 / 18-20;
 * reference nn/modules/conv.py:323-333 Concat, torch nn.Upsample(scale_factor=2, mode="nearest"), block.py:226-230) WITHOUT the
 * upsampled tensor: the K chunks of the first C0 input channels are fetched from the SMALL map [B,H/2,W/2,C0] through a 5-D tensor
 * map whose two repeat dimensions have stride 0, so every source pixel lands in the four rows of the A tile its 2x2 output pixels
 * own; the other C1 channels come from x_skip [B,H,W,C1] as usual.  w: [Cout][C0 + C1] (lpc_conv2d_tc's 1x1 packing, upsampled
 * channels first).  Bit-identical to lpc_upsample2x + lpc_conv2d_tc.  C0, C1 multiples of 64, even H and W (output size). */
int lpc_conv1x1_up2cat_tc_supported(int C0, int C1, int Cout, int H, int W, int xs_ld, int xk_ld, int y_ld);
int lpc_conv1x1_up2cat_tc(const void* x_small, int xs_ld, int C0, const void* x_skip, int xk_ld, int C1, int B, int H, int W,
                          const void* w, const float* bias, int Cout, void* y, int y_ld, int act, void* stream);

/* Conv 3x3 s1 (Cin 16 -> 32) fused with the space_to_depth + 1x1 conv that follows it in the SPD-Conv stem of the LPC YAML
 * (layers 1-3: conv.Conv(16,32,3,1) -> space_to_depth -> C2f.cv1; reference nn/modules/conv.py:36-54, nn/modules/block.py:4063-4070
 * space_to_depth, block.py:226-230 C2f.forward): y[B,H/2,W/2,C2] = act2(W2 * s2d(act1(W1 * x + b1)) + b2), where w2 is the 1x1
 * conv re-packed as a 2x2 stride-2 conv ([C2][kpad(32,2)=128], K = (ky, kx, c)) and w1 is lpc_conv2d_tc's [32][kpad(16,3)=192].
 * The 32-channel intermediate (419 MB at 320x320 batch 64) stays in shared memory.  bf16 only; results are bit-identical to
 * lpc_conv2d_tc(k=3) followed by lpc_conv2d_tc(k=2, s=2).  _supported: Cin == 16, C1 == 32, C2 % 16 == 0, C2 <= 64, even H and W. */
int lpc_conv3x3_s2d_tc_supported(int Cin, int C1, int C2, int H, int W, int x_ld, int y_ld);
int lpc_conv3x3_s2d_tc(const void* x, int x_ld, int B, int H, int W, int Cin, const void* w1, const float* bias1, int C1, int act1,
                       const void* w2, const float* bias2, int C2, int act2, void* y, int y_ld, void* stream);

/* Depthwise 3x3 (stride 1, pad 1, bias, dw_act) -> pointwise 1x1 (C1, bias, act1) [-> pointwise 1x1 (C2, bias, act2)] as
 * ONE tcgen05 kernel: the depthwise result is written straight into the shared-memory A operand of the 1x1 GEMM, and
 * the first GEMM's activated output into the A operand of the second; nothing between x and y touches HBM.  Replaces
 * the class-branch chain of v10Detect (nn/modules/head.py:504-505, three launches -> one, five -> two per level) and the
 * dw -> 1x1 links of CIB (block.py:744-750).  bf16 NHWC; dw_w [9][Cin] fp32; w1 [C1][pad64(Cin)], w2 [C2][pad64(C1)]
 * bf16 (lpc_conv2d_tc's 1x1 packing); C1, C2 multiples of 16 and <= 256, C2 = 0 / w2 = NULL for a single pointwise
 * stage.  rowmax_keys: as lpc_conv2d_tc_rowmax (per-pixel key of max_c of the final outputs), or NULL.
 * Cin a multiple of 16; dw_act == act1 in {none, SiLU, Mish}; act2 none.
 * lpc_dwpw_tc_supported: 1 if the shape (incl. its shared-memory plan) and the activations are taken. */
int lpc_dwpw_tc_supported(int Cin, int C1, int C2, int x_ld, int y_ld, int dw_act, int act1, int act2);
int lpc_dwpw_tc(const void* x, int x_ld, int B, int H, int W, int Cin, const float* dw_w, const float* dw_bias, int dw_act,
                const void* w1, const float* b1, int C1, int act1, const void* w2, const float* b2, int C2, int act2,
                void* y, int y_ld, unsigned int* rowmax_keys, long long rowmax_img_stride, int rowmax_offset, void* stream);

/* Stem: Conv(3, Cout, 3, stride, pad 1) on an NHWC input whose pixel pitch is exactly 4 (lpc_pack_input's Cpad=4).
 * bf16, stride 2, even W: tcgen05 (stem_tc.cu - the 3x3 windows of an 8x16 output tile arrive as one 3-D TMA box, builder
 * warps turn them into swizzled K-major im2col rows, K = 36 + 2 bias slots -> 48, three UMMAs per 128 pixels); other
 * shapes: a register-tiled CUDA-core kernel (stem.cu).  The fp32 validation mode does not call this entry (the host
 * routes it to lpc_conv2d_direct, which accumulates in fp64).  w: [27][Cout] fp32 with row index (ky*3+kx)*3+cin;
 * Cout % 8 == 0, Cout <= 96.  Replaces layer 0 of every v10 / LPC YAML (conv.py:36-54). */
int lpc_stem_conv(int dtype, const void* x, int B, int H, int W, const float* w, const float* bias, int stride,
                  int Cout, void* y, int y_ld, int act, void* stream);

/* The same stem straight from the uint8 HWC image (the LoadPilAndNumpy source, [B][H][W][3] bytes, BGR when swap_rb != 0):
 * y = act(conv3x3_s2(bf16(x / 255)) + b), bit-identical to lpc_pack_u8 (bf16) followed by lpc_stem_conv - the /255, the
 * BGR->RGB swap and the NHWC-4 padding of predictor.preprocess (engine/predictor.py:115-133) happen while the builder
 * warps assemble the im2col rows (the image arrives as one 112-byte x 17-row TMA box per tile), so the 4-channel bf16 copy of the batch (210 MB at B=64) is never written or read.
 * bf16 output only; stride 2, even H, W % 16 == 0 (16-byte row pitch), Cout % 16 == 0, Cout <= 128; no LetterBox border
 * (callers with a border or a resize use lpc_pack_u8 / lpc_letterbox_u8 + lpc_stem_conv). */
int lpc_stem_conv_u8_supported(int H, int W, int stride, int Cout, int y_ld, int act);
int lpc_stem_conv_u8(const void* x, int B, int H, int W, int swap_rb, const float* w, const float* bias, int stride, int Cout,
                     void* y, int y_ld, int act, void* stream);

/* ---- depthwise convolution ---------------------------------------------------------------------
 * groups == C convs: CIB / RepVGGDW (block.py:700-756; the 3x3 branch is merged into the 7x7 on the
 * host as RepVGGDW.fuse does, :714-733), SCDown.cv2 (:822), Attention.pe (:780), LPC.cv2 5x5 (:5806),
 * SPCA dilated 3x3 d=1,2,3 (:5728-5731), v10Detect cls branch (head.py:504-505).
 * w: [k*k][C] fp32; bias [C] fp32 or NULL; k in {3,5,7}; stride in {1,2}; C % 8 == 0. */
int lpc_dwconv2d(int dtype, const void* x, int x_ld, int B, int H, int W, int C,
                 const float* w, const float* bias, int k, int stride, int pad, int dil,
                 void* y, int y_ld, int act, const void* res, int res_ld, void* stream);

/* ---- SPPF pooling (block.py:171-176): y[:, 0:C]=m(x), [C:2C]=m(m(x)), [2C:3C]=m(m(m(x))) with
 * m = MaxPool2d(5,1,2), computed in one pass as 5x5 / 9x9 / 13x13 windows (-inf padding). */
int lpc_sppf_pool(int dtype, const void* x, int x_ld, int B, int H, int W, int C,
                  void* y, int y_ld, void* stream);

/* ---- PSA attention core (block.py:783-793): out[n, i, h*hd+d] = sum_j softmax_j(scale q_i.k_j) v[j,d].
 * qkv holds, per pixel, [q: heads*kd | k: heads*kd | v: heads*hd] (the host permutes the qkv conv's
 * output channels from the reference's per-head interleave, block.py:787). N = H*W tokens. */
int lpc_psa_attention(int dtype, const void* qkv, int qkv_ld, int B, int N, int heads, int kd, int hd,
                      void* out, int out_ld, void* stream);

/* ---- neck glue (nn.Upsample(None,2,'nearest'), Concat conv.py:331, space_to_depth block.py:4069,
 * LPC channel de-interleave block.py:5819-5825) */
int lpc_upsample2x(int dtype, const void* x, int x_ld, int B, int H, int W, int C, void* y, int y_ld, void* stream);
int lpc_copy_channels(int dtype, const void* x, int x_ld, long long npix, int C, void* y, int y_ld, void* stream);
int lpc_space_to_depth(int dtype, const void* x, int x_ld, int B, int H, int W, int C, void* y, int y_ld, void* stream);
int lpc_channel_deinterleave(int dtype, const void* x, int x_ld, long long npix, int C, void* y, int y_ld, void* stream);
/* fp32 NCHW image batch -> NHWC view (channels c >= C zero filled up to Cpad). predictor.py:115-133. */
int lpc_pack_input(int dtype, const float* x_nchw, int B, int C, int H, int W, void* y, int y_ld, int Cpad, void* stream);
/* uint8 HWC image batch [B,Hs,Ws,3] (device) -> NHWC network input [B,H,W,4] (pixel pitch 4, channel 3 = 0), values / 255:
 * the array-source preprocess (engine/predictor.py:115-133: BGR->RGB when swap_rb, BHWC->BCHW, float, /255) fused with
 * LetterBox's constant border (data/augment.py:725-731): the image lands at (top, left), the rest is pad_value / 255.
 * No resize (LetterBox ratio 1) in this round. */
int lpc_pack_u8(int dtype, const void* src_u8, int B, int Hs, int Ws, int top, int left, int H, int W, int pad_value,
                int swap_rb, void* y, void* stream);

/* The same with LetterBox's resize (data/augment.py:726-727, cv2.resize INTER_LINEAR) in front: the [hs,ws] images are
 * resized to [nh,nw] with OpenCV's 8-bit fixed-point linear interpolation (bit-exact), placed at (top,left) of [H,W].
 * xtab [nw][3] = {x0, a0, a1}, ytab [nh][4] = {y0, y1, b0, b1}: int32 device arrays (11-bit coefficients, built as
 * imgproc/resize.cpp builds them; see engine.resize_tables). */
int lpc_letterbox_u8(int dtype, const void* src_u8, int B, int hs, int ws, int nh, int nw, const int* xtab, const int* ytab,
                     int top, int left, int H, int W, int pad_value, int swap_rb, void* y, void* stream);

/* ---- CBAM / SPCA pooled gates (conv.py:278-320, block.py:5735-5747) */
/* partial[b][chunk][c] = sum of x over the chunk's pixels, chunk count = lpc_global_avgpool_chunks(B, HW)
 * (a fixed-order two-stage reduction: deterministic, no atomics). */
int lpc_global_avgpool_chunks(int B, int HW);
int lpc_global_avgpool(int dtype, const void* x, int x_ld, int B, int HW, int C, float* partial, void* stream);
/* v[b] = in_scale * sum_parts in[b][part][:];  out[b] = act2(W2 . act1(W1 . v[b] + b1) + b2); W2 may be NULL
 * (single layer). W row-major [Cout][Cin]. */
int lpc_channel_mlp(const float* in, int B, int parts, float in_scale, int C0, const float* W1, const float* b1, int C1,
                    int act1, const float* W2, const float* b2, int C2, int act2, float* out, void* stream);
/* stats[b,p,0] = mean_c(x*ca), stats[b,p,1] = max_c(x*ca) */
int lpc_cbam_stats(int dtype, const void* x, int x_ld, int B, int HW, int C, const float* ca, float* stats, void* stream);
/* y = x * ca * sigmoid(conv_kxk(stats, w)) ; w: [2][k][k] fp32 (cin-major as nn.Conv2d(2,1,k)) */
int lpc_cbam_apply(int dtype, const void* x, int x_ld, int B, int H, int W, int C, const float* ca,
                   const float* stats, const float* w, int k, void* y, int y_ld, void* stream);

/* ---- v10Detect tail ------------------------------------------------------------------------------
 * raw[l]: level-l head map, NHWC view with 4*16 box-distribution channels then nc class logits
 * (head.py:76 cat order); level l has (H0>>l) x (W0>>l) cells and stride strides[l].
 *
 * lpc_v10_decode: Detect.inference (head.py:45-71) = DFL (block.py:57-60) + make_anchors
 *   (utils/tal.py:294-306) + dist2bbox xywh (tal.py:309-319) + sigmoid -> y[B][4+nc][A] fp32.
 * lpc_v10_decode_topk: the same fused with ops.v10postprocess (utils/ops.py:851-864), xywh2xyxy
 *   (ops.py:402-421) and clip_boxes (ops.py:305-324; skipped when img_h<=0):
 *   dets[B][K][6] = x1,y1,x2,y2,score,label sorted by score desc; anchor_idx[B][K] (may be NULL).
 *   Ties in score are ordered by ascending (anchor*nc + class).  Requires A*nc >= K, K <= 1024.
 * lpc_v10_postprocess: ops.v10postprocess on an already decoded preds[B][A][4+nc] fp32 tensor with
 *   arbitrary strides (elements): boxes[B][K][4] (gathered, unchanged), scores[B][K], labels[B][K] i64.
 */
size_t lpc_v10_topk_workspace_bytes(int B, int A, int K);
int lpc_v10_decode(int dtype, const void* raw0, const void* raw1, const void* raw2, int ld,
                   int B, int H0, int W0, int nc, const float* strides3_host, float* y, void* stream);
int lpc_v10_decode_topk(int dtype, const void* raw0, const void* raw1, const void* raw2, int ld,
                        int B, int H0, int W0, int nc, const float* strides3_host, int K,
                        int img_h, int img_w, void* workspace, size_t ws_bytes,
                        float* dets, int* anchor_idx, void* stream);
/* lpc_v10_decode_topk with the per-anchor keys optionally ALREADY in the workspace (keys_ready != 0): written by
 * lpc_conv2d_tc_rowmax on the three class-branch convs (keys[b * A + level_offset + cell]); stage 1 then needs no pass
 * over the class logits. */
int lpc_v10_decode_topk_keys(int dtype, const void* raw0, const void* raw1, const void* raw2, int ld,
                             int B, int H0, int W0, int nc, const float* strides3_host, int K,
                             int img_h, int img_w, void* workspace, size_t ws_bytes, int keys_ready,
                             float* dets, int* anchor_idx, void* stream);
/* lpc_v10_decode_topk_keys followed, inside the same kernel, by the predictor's rescale to the ORIGINAL image
 * (ops.scale_boxes utils/ops.py:89-124 + clip_boxes :305-324, called from models/yolov10/predict.py:35):
 * scale_back = device array [B][5] of (pad_x, pad_y, gain, orig_w, orig_h) per image, or NULL for none. */
int lpc_v10_decode_topk_scaled(int dtype, const void* raw0, const void* raw1, const void* raw2, int ld,
                               int B, int H0, int W0, int nc, const float* strides3_host, int K,
                               int img_h, int img_w, void* workspace, size_t ws_bytes, int keys_ready,
                               const float* scale_back, float* dets, int* anchor_idx, void* stream);
/* The reference's small box helpers as kernels (function-level API parity; the fused tail above does not call them):
 * lpc_make_anchors  utils/tal.py:294-306: anchors[A][2] = (x+offset, y+offset), stride_out[A]; levels in the given
 *                   order, row-major cells; hw_host = {h0,w0,h1,w1,...} (1..4 levels).
 * lpc_dist2bbox     utils/tal.py:309-319 on rows: dist[n][4] = (l,t,r,b), anchors[n_anchors][2] re-used every
 *                   n_anchors rows (n % n_anchors == 0) -> out[n][4] = (cx,cy,w,h) if xywh else (x1,y1,x2,y2).
 * lpc_scale_boxes   in place on n rows of row_stride floats whose first four are a box: xywh2xyxy (utils/ops.py:402-421)
 *                   when xywh_in, then scale_boxes (ops.py:89-124: minus padding, / gain) and clip_boxes (:305-324)
 *                   when clip_w > 0. */
int lpc_make_anchors(int n_levels, const int* hw_host, const float* strides_host, float offset, float* anchors,
                     float* stride_out, void* stream);
int lpc_dist2bbox(const float* dist, const float* anchors, long long n, long long n_anchors, int xywh, float* out, void* stream);
int lpc_scale_boxes(float* boxes, long long n, int row_stride, int xywh_in, float pad_x, float pad_y, float gain,
                    float clip_w, float clip_h, void* stream);
int lpc_v10_postprocess(const float* preds, long long stride_b, long long stride_a, long long stride_c,
                        int B, int A, int nc, int K, void* workspace, size_t ws_bytes,
                        float* boxes, float* scores, long long* labels, void* stream);

/* ---- launch plans: the plan-level entry of SURVEY.md section 8(b) ----------------------------------------------------
 * The reference's layer loop is Python (BaseModel._predict_once, nn/tasks.py:83-111).  A plan is the launch sequence of one
 * step, recorded once from whatever host code issues it and replayed from C: between lpc_plan_begin() and lpc_plan_end()
 * every kernel this thread launches through the library is stored with its grid, shared memory and a copy of its
 * arguments (the launches also run).  lpc_plan_run re-issues them with the recorded stream structure; lpc_plan_run_graph
 * captures that into a CUDA graph on first use and launches the graph afterwards.  The device buffers of the recorded
 * step must stay alive at the same addresses (record inside a private memory pool); inputs are refreshed by writing into
 * the recorded input buffer.  Returns 0 / a negative status; lpc_plan_size = number of recorded launches. */
typedef struct lpc_plan lpc_plan;
int lpc_plan_begin(void);
int lpc_plan_end(lpc_plan** plan);
int lpc_plan_size(const lpc_plan* plan);
/* while recording: stream `waiter` waits for everything issued so far on stream `signaler` (the host code's side-stream
 * forks / joins); a no-op outside a recording */
int lpc_plan_wait(void* waiter_stream, void* signaler_stream);
int lpc_plan_run(lpc_plan* plan, void* stream);
int lpc_plan_run_graph(lpc_plan* plan, void* stream);
void lpc_plan_destroy(lpc_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* LPCYOLO_H_ */
