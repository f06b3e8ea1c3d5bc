"""Tensor-level wrappers over the C ABI.

Activations are torch CUDA tensors of LOGICAL shape [B,C,H,W] and PHYSICAL layout NHWC, possibly a
channel slice of a wider buffer (``buf[:, c0:c1]``): the kernels take (pointer, pixel pitch).  Keeping the
logical shape NCHW keeps every reference module signature unchanged.
"""
import ctypes as C

import os

import torch

from . import _lib
from ._lib import ACT_MISH, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SILU, BF16, F32, LpcError, check  # noqa: F401

_DT = {torch.bfloat16: BF16, torch.float32: F32}


# Optional per-launch profiling (bench.py): when PROFILE is a list, every op appends
# (kind, start_event, end_event, algorithmic_flops, algorithmic_bytes, shape tag).
PROFILE = None
# Optional launch recording (bench.py): when REPLAY is a list, the tcgen05 conv and fused-tail ops append
# (kind, closure, algorithmic_flops, algorithmic_bytes, tag); calling a closure re-issues exactly that launch (same
# pointers), so bench.py can capture e.g. all dense-conv launches of one step into a CUDA graph and time them back to back.
REPLAY = None


class _prof:
    def __init__(self, kind, flops=0.0, nbytes=0.0, tag=""):
        self.rec = None
        if PROFILE is not None:
            self.rec = [kind, torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), flops, nbytes, tag]

    def __enter__(self):
        if self.rec is not None:
            self.rec[1].record()
        return self

    def __exit__(self, *a):
        if self.rec is not None:
            self.rec[2].record()
            PROFILE.append(tuple(self.rec))


def _profiled(fn):
    import functools

    @functools.wraps(fn)
    def wrap(*a, **k):
        if PROFILE is None:
            return fn(*a, **k)
        t = a[0] if a and torch.is_tensor(a[0]) else None
        tag = "x".join(map(str, t.shape)) if t is not None else ""
        nb = 2.0 * t.numel() * t.element_size() if t is not None else 0.0
        with _prof(fn.__name__, 0.0, nb, tag):
            return fn(*a, **k)
    return wrap


def dt_code(dtype):
    try:
        return _DT[dtype]
    except KeyError:
        raise LpcError(f"unsupported activation dtype {dtype}; use torch.bfloat16 or torch.float32") from None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---- independent op chains on side streams -------------------------------------------------------------------------
_SIDE_STREAMS = {}
BRANCH_STREAMS = os.environ.get("LPC_BRANCH_STREAMS", os.environ.get("LPC_HEAD_STREAMS", "1")) != "0"


def fork_join(jobs, device):
    """Run independent callables concurrently: jobs[0] on the current stream, the others on cached side streams that
    first wait for everything issued so far; the current stream then waits for all of them.  Under CUDA-graph capture
    this records parallel branches.  Every tensor a job READS must have been produced on the current stream before the
    call, every tensor it WRITES must outlive the call (the callers pre-allocate outputs on the current stream);
    temporaries a job allocates belong to its side stream in the caching allocator, so they are never handed to another
    stream while in use.  LPC_BRANCH_STREAMS=0 serialises."""
    if not BRANCH_STREAMS or len(jobs) < 2 or device.type != "cuda":
        for j in jobs:
            j()
        return
    main = torch.cuda.current_stream(device)
    # one pool per (device, forking stream): a nested fork (the head's branches inside a batch-half running on a side stream)
    # must not be handed the stream it is running on
    key = (device.index if device.index is not None else torch.cuda.current_device(), main.cuda_stream)
    pool = _SIDE_STREAMS.setdefault(key, [])
    while len(pool) < len(jobs) - 1:
        pool.append(torch.cuda.Stream(device=device))
    fork = torch.cuda.Event()
    fork.record(main)
    L = _lib.lib()
    for j, st in zip(jobs[1:], pool):
        st.wait_event(fork)
        L.lpc_plan_wait(C.c_void_p(st.cuda_stream), C.c_void_p(main.cuda_stream))      # no-op unless a launch plan is being recorded
        with torch.cuda.stream(st):
            j()
    jobs[0]()
    for st in pool[:len(jobs) - 1]:
        main.wait_stream(st)
        L.lpc_plan_wait(C.c_void_p(main.cuda_stream), C.c_void_p(st.cuda_stream))


def new_act(B, Cc, H, W, dtype, device):
    """Fresh NHWC activation with logical NCHW shape."""
    return torch.empty((B, H, W, Cc), dtype=dtype, device=device).permute(0, 3, 1, 2)


def view_of(t):
    """-> (data_ptr, pitch) after checking that ``t`` is an NHWC view (size-1 dims are unconstrained)."""
    if not t.is_cuda:
        raise LpcError("lpc-yolo_b200 ops run on CUDA tensors only (there is no CPU fallback)")
    B, Cc, H, W = t.shape
    sB, sC, sH, sW = t.stride()
    if W > 1:
        ld = sW
    elif H > 1:
        ld = sH
    elif B > 1:
        ld = sB
    else:
        ld = Cc
    ok = (Cc == 1 or sC == 1) and (W == 1 or sW == ld) and (H == 1 or sH == W * ld) and (B == 1 or sB == H * W * ld) and ld >= Cc
    if not ok:
        raise LpcError(f"tensor of shape {tuple(t.shape)} / strides {t.stride()} is not an NHWC view")
    return t.data_ptr(), ld


def is_nhwc_view(t):
    try:
        view_of(t)
        return True
    except LpcError:
        return False


def as_act(x, dtype):
    """Boundary conversion for tensors that arrive NCHW-contiguous (tests, user code)."""
    if x.dtype == dtype and is_nhwc_view(x):
        return x
    if x.dim() == 4 and x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] <= 4:
        return pack_input(x, dtype)
    y = new_act(*x.shape[:1], x.shape[1], x.shape[2], x.shape[3], dtype, x.device)
    y.copy_(x)
    return y


def _fp(t):
    return None if t is None else C.c_void_p(t.data_ptr())


# ---------------------------------------------------------------------------------------------------------
@_profiled
def pack_input(x, dtype, cpad=4):
    """fp32 NCHW image batch -> NHWC activation with C padded to ``cpad`` (logical C stays x.shape[1])."""
    B, Cc, H, W = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    buf = torch.empty((B, H, W, cpad), dtype=dtype, device=x.device)
    check(_lib.lib().lpc_pack_input(dt_code(dtype), _fp(x), B, Cc, H, W, _fp(buf), cpad, cpad, _stream()), "pack_input")
    return buf.permute(0, 3, 1, 2)[:, :Cc]


@_profiled
def pack_u8(src, dtype, top=0, left=0, H=None, W=None, pad_value=114, swap_rb=True, out=None):
    """uint8 HWC images [B,Hs,Ws,3] on the device -> NHWC network input (logical [B,3,H,W]), /255, BGR->RGB,
    LetterBox border (no resize).  ``out``: an optional [B,H,W,4] buffer of ``dtype`` (CUDA-graph static input)."""
    if not src.is_cuda:
        raise LpcError("lpc-yolo_b200 ops run on CUDA tensors only (there is no CPU fallback)")
    assert src.dtype == torch.uint8 and src.dim() == 4 and src.shape[3] == 3 and src.is_contiguous()
    B, Hs, Ws, _ = src.shape
    H, W = H or Hs, W or Ws
    buf = out if out is not None else torch.empty((B, H, W, 4), dtype=dtype, device=src.device)
    assert buf.shape == (B, H, W, 4) and buf.dtype == dtype and buf.is_contiguous()
    check(_lib.lib().lpc_pack_u8(dt_code(dtype), _fp(src), B, Hs, Ws, top, left, H, W, pad_value, int(swap_rb), _fp(buf), _stream()), "pack_u8")
    return buf.permute(0, 3, 1, 2)[:, :3]


def resize_tables(hs, ws, nh, nw, device):
    """Coefficient tables of OpenCV's 8-bit INTER_LINEAR resize (imgproc/resize.cpp: fx = (float)((dx+0.5)*scale-0.5),
    cvFloor, border handling, saturate_cast<short>(w * 2048)), as int32 device tensors: xtab [nw,3], ytab [nh,4]."""
    import numpy as np

    def sat(v):
        return int(max(-32768, min(32767, np.rint(v))))

    xt = np.empty((nw, 3), np.int32)
    sx = ws / nw
    for d in range(nw):
        f = np.float32((d + 0.5) * sx - 0.5)
        s0 = int(np.floor(f))
        f = np.float32(f - np.float32(s0))
        if s0 < 0:
            f, s0 = np.float32(0), 0
        if s0 >= ws - 1:
            f, s0 = np.float32(0), ws - 1
        xt[d] = (s0, sat(np.float32(np.float32(1) - f) * np.float32(2048)), sat(f * np.float32(2048)))
    yt = np.empty((nh, 4), np.int32)
    sy = hs / nh
    for d in range(nh):
        f = np.float32((d + 0.5) * sy - 0.5)
        s0 = int(np.floor(f))
        f = np.float32(f - np.float32(s0))        # rows are clamped, the weights are not (unlike x)
        yt[d] = (min(max(s0, 0), hs - 1), min(max(s0 + 1, 0), hs - 1), sat(np.float32(np.float32(1) - f) * np.float32(2048)), sat(f * np.float32(2048)))
    return torch.from_numpy(xt).to(device), torch.from_numpy(yt).to(device)


@_profiled
def letterbox_u8(src, dtype, nh, nw, tables, top, left, H, W, pad_value=114, swap_rb=True, out=None):
    """uint8 HWC images [B,hs,ws,3] -> resized to [nh,nw] (cv2 INTER_LINEAR, bit-exact), bordered to [H,W], /255, BGR->RGB."""
    if not src.is_cuda:
        raise LpcError("lpc-yolo_b200 ops run on CUDA tensors only (there is no CPU fallback)")
    assert src.dtype == torch.uint8 and src.dim() == 4 and src.shape[3] == 3 and src.is_contiguous()
    B, hs, ws, _ = src.shape
    xt, yt = tables
    assert tuple(xt.shape) == (nw, 3) and tuple(yt.shape) == (nh, 4) and xt.dtype == torch.int32 and yt.dtype == torch.int32
    buf = out if out is not None else torch.empty((B, H, W, 4), dtype=dtype, device=src.device)
    assert buf.shape == (B, H, W, 4) and buf.dtype == dtype and buf.is_contiguous()
    check(_lib.lib().lpc_letterbox_u8(dt_code(dtype), _fp(src), B, hs, ws, nh, nw, _fp(xt), _fp(yt), top, left, H, W, pad_value,
                                      int(swap_rb), _fp(buf), _stream()), "letterbox_u8")
    return buf.permute(0, 3, 1, 2)[:, :3]


def conv2d(x, pc, out=None, res=None, chan_scale=None, rowmax=None):
    """Dense conv through a PackedConv ``pc`` (see pack.py). Chooses tcgen05 when the shape allows."""
    B, Cin, H, W = x.shape
    assert Cin == pc.cin, (Cin, pc.cin)
    Ho = (H + 2 * pc.p - pc.k) // pc.s + 1
    Wo = (W + 2 * pc.p - pc.k) // pc.s + 1
    if out is None:
        out = new_act(B, pc.cout, Ho, Wo, x.dtype, x.device)
    assert tuple(out.shape) == (B, pc.cout, Ho, Wo), (tuple(out.shape), (B, pc.cout, Ho, Wo))
    xp, xld = view_of(x)
    yp, yld = view_of(out)
    rp, rld = (0, 0) if res is None else view_of(res)
    L = _lib.lib()
    use_tc = (pc.w_tc is not None and x.dtype == torch.bfloat16 and xp % 16 == 0 and yp % 16 == 0 and rp % 16 == 0
              and rld % 8 == 0 and L.lpc_conv2d_tc_supported(Cin, pc.cout, pc.k, pc.s, pc.p, xld, yld))
    flops = 2.0 * B * Ho * Wo * pc.cout * Cin * pc.k * pc.k
    # (fp32 validation mode: the stem goes through conv_direct's fp64-accumulating kernel like every other dense conv)
    if pc.w_stem is not None and x.dtype == torch.bfloat16 and xld == 4 and chan_scale is None and res is None and yld % 8 == 0 and xp % 16 == 0 and yp % 16 == 0:
        nb = x.element_size() * (B * H * W * 4 + B * Ho * Wo * pc.cout)
        with _prof("stem_conv", flops, nb, f"3->{pc.cout} k3s{pc.s} {H}x{W} B{B}"):
            check(L.lpc_stem_conv(dt_code(x.dtype), xp, B, H, W, _fp(pc.w_stem), _fp(pc.bias), pc.s, pc.cout, yp, yld, pc.act,
                                  _stream()), "stem_conv")
        return out
    nbytes = x.element_size() * (B * H * W * Cin + B * Ho * Wo * pc.cout * (2 if res is not None else 1) + pc.cout * Cin * pc.k * pc.k)
    tag = f"{Cin}->{pc.cout} k{pc.k}s{pc.s} {H}x{W} B{B}"
    if use_tc:
        # rowmax = (uint8 workspace tensor, keys per image, first key of this map): the conv also emits the per-pixel key
        # of max_c(output) for the fused tail (lpc_conv2d_tc_rowmax); rowmax["ok"] reports whether the kernel took it
        rm = None
        if rowmax is not None:
            rm = (C.c_void_p(rowmax["ws"].data_ptr()), int(rowmax["A"]), int(rowmax["off"]))

        def launch(keep=(x, out, res, chan_scale, pc, rowmax)):
            if rm is not None:
                st = L.lpc_conv2d_tc_rowmax(xp, xld, B, H, W, Cin, _fp(pc.w_tc), _fp(pc.bias), pc.k, pc.s, pc.p, pc.cout, yp, yld,
                                            pc.act, _fp(chan_scale), rp or None, rld, rm[0], rm[1], rm[2], _stream())
                if st == 0:
                    rowmax["ok"] = rowmax.get("ok", True)
                    return
                if st != _lib.E_UNSUPPORTED:
                    check(st, "conv2d_tc_rowmax")
                rowmax["ok"] = False          # shape not taken: plain conv, the tail runs its own key pass
            check(L.lpc_conv2d_tc(xp, xld, B, H, W, Cin, _fp(pc.w_tc), _fp(pc.bias), pc.k, pc.s, pc.p, pc.cout, yp, yld,
                                  pc.act, _fp(chan_scale), rp or None, rld, _stream()), "conv2d_tc")
        if REPLAY is not None:
            plain = None
            if rm is not None:        # the same launch without the fused stage-1 keys (bench.py measures what they cost)
                def plain(keep=(x, out, res, chan_scale, pc)):
                    check(L.lpc_conv2d_tc(xp, xld, B, H, W, Cin, _fp(pc.w_tc), _fp(pc.bias), pc.k, pc.s, pc.p, pc.cout, yp, yld,
                                          pc.act, _fp(chan_scale), rp or None, rld, _stream()), "conv2d_tc")
            REPLAY.append(("conv2d_tc", launch, flops, nbytes, tag, plain))
        with _prof("conv2d_tc", flops, nbytes, tag):
            launch()
    else:
        if rowmax is not None:
            rowmax["ok"] = False
        with _prof("conv2d_direct", flops, nbytes, tag):
            check(L.lpc_conv2d_direct(dt_code(x.dtype), xp, xld, B, H, W, Cin, _fp(pc.w_direct), _fp(pc.bias), pc.k, pc.s,
                                      pc.p, pc.cout, yp, yld, pc.act, _fp(chan_scale), rp or None, rld, _stream()),
                  "conv2d_direct")
    return out


def stem_u8_supported(src, pc, out_ld=None):
    """Can the stem conv ``pc`` read the uint8 HWC images ``src`` [B,H,W,3] directly (lpc_stem_conv_u8)?"""
    if pc.w_stem is None or not (torch.is_tensor(src) and src.is_cuda and src.dtype == torch.uint8 and src.dim() == 4 and src.shape[3] == 3):
        return False
    B, H, W, _ = src.shape
    return (src.is_contiguous() and src.data_ptr() % 16 == 0 and (pc.k, pc.p) == (3, 1)
            and bool(_lib.lib().lpc_stem_conv_u8_supported(H, W, pc.s, pc.cout, out_ld if out_ld is not None else pc.cout, pc.act)))


def stem_conv_u8(src, pc, swap_rb=True, out=None):
    """Stem conv on uint8 HWC images (BGR when ``swap_rb``): /255, channel swap and the conv in one kernel -> bf16 NHWC."""
    B, H, W, _ = src.shape
    Ho, Wo = (H + 2 - 3) // pc.s + 1, (W + 2 - 3) // pc.s + 1
    if out is None:
        out = new_act(B, pc.cout, Ho, Wo, torch.bfloat16, src.device)
    assert tuple(out.shape) == (B, pc.cout, Ho, Wo) and out.dtype == torch.bfloat16
    yp, yld = view_of(out)
    flops = 2.0 * B * Ho * Wo * pc.cout * 27
    nb = B * H * W * 3 + 2 * B * Ho * Wo * pc.cout
    with _prof("stem_conv_u8", flops, nb, f"u8 3->{pc.cout} k3s{pc.s} {H}x{W} B{B}"):
        check(_lib.lib().lpc_stem_conv_u8(_fp(src), B, H, W, int(swap_rb), _fp(pc.w_stem), _fp(pc.bias), pc.s, pc.cout, yp, yld, pc.act,
                                         _stream()), "stem_conv_u8")
    return out


_UPCAT_REFUSED = False      # set when cuTensorMapEncodeTiled refused the stride-0 map once (then the fold is off for the process)


def conv1x1_upcat_supported(x_small, x_skip, pc, out_ld=None):
    """Can ``pc`` (a 1x1 conv over cat[upsample2x(x_small), x_skip]) run without the upsampled tensor (lpc_conv1x1_up2cat_tc)?"""
    if _UPCAT_REFUSED or x_small.dtype != torch.bfloat16 or x_skip.dtype != torch.bfloat16 or pc.w_tc is None or (pc.k, pc.s, pc.p) != (1, 1, 0):
        return False
    B, C0, Hs, Ws = x_small.shape
    B1, C1, H, W = x_skip.shape
    if B1 != B or (H, W) != (2 * Hs, 2 * Ws) or pc.cin != C0 + C1:
        return False
    sp, sld = view_of(x_small)
    kp, kld = view_of(x_skip)
    return sp % 16 == 0 and kp % 16 == 0 and bool(_lib.lib().lpc_conv1x1_up2cat_tc_supported(C0, C1, pc.cout, H, W, sld, kld,
                                                                                            out_ld if out_ld is not None else pc.cout))


def conv1x1_upcat(x_small, x_skip, pc, out=None):
    """act(W * cat[upsample2x(x_small), x_skip] + b); callers check ``conv1x1_upcat_supported``.  -> None when the driver refuses
    the repeating tensor map (nothing launched)."""
    B, C0, Hs, Ws = x_small.shape
    _, C1, H, W = x_skip.shape
    if out is None:
        out = new_act(B, pc.cout, H, W, x_skip.dtype, x_skip.device)
    assert tuple(out.shape) == (B, pc.cout, H, W)
    sp, sld = view_of(x_small)
    kp, kld = view_of(x_skip)
    yp, yld = view_of(out)
    L = _lib.lib()
    flops = 2.0 * B * H * W * pc.cout * (C0 + C1)
    nbytes = x_skip.element_size() * (B * Hs * Ws * C0 + B * H * W * (C1 + pc.cout) + pc.cout * (C0 + C1))
    tag = f"up2({C0})+{C1}->{pc.cout} k1s1 {H}x{W} B{B}"

    def launch(keep=(x_small, x_skip, out, pc)):
        st = L.lpc_conv1x1_up2cat_tc(sp, sld, C0, kp, kld, C1, B, H, W, _fp(pc.w_tc), _fp(pc.bias), pc.cout, yp, yld, pc.act, _stream())
        if st == _lib.E_UNSUPPORTED:
            return False
        check(st, "conv1x1_up2cat_tc")
        return True
    with _prof("conv1x1_up2cat_tc", flops, nbytes, tag):
        ok = launch()
    if not ok:
        # the driver refused the repeating (stride-0) tensor map: nothing was launched; the caller materialises the upsampled half
        global _UPCAT_REFUSED
        _UPCAT_REFUSED = True
        return None
    if REPLAY is not None:
        REPLAY.append(("conv2d_tc", launch, flops, nbytes, tag, None))
    return out


def conv3x3_s2d_supported(x, pc1, pc2, out_ld=None):
    """Can conv3x3(pc1) -> space_to_depth -> conv1x1 (pc2 = its 2x2 stride-2 re-packing) run as ONE kernel (lpc_conv3x3_s2d_tc)?"""
    if x.dtype != torch.bfloat16 or pc1.w_tc is None or pc2.w_tc is None:
        return False
    if (pc1.k, pc1.s, pc1.p) != (3, 1, 1) or (pc2.k, pc2.s, pc2.p) != (2, 2, 0) or pc2.cin != pc1.cout:
        return False
    B, Cin, H, W = x.shape
    xp, xld = view_of(x)
    return xp % 16 == 0 and bool(_lib.lib().lpc_conv3x3_s2d_tc_supported(Cin, pc1.cout, pc2.cout, H, W, xld, out_ld if out_ld is not None else pc2.cout))


def conv3x3_s2d(x, pc1, pc2, out=None):
    """act2(conv2x2s2(act1(conv3x3(x)))) without writing the 3x3 conv's output; callers check ``conv3x3_s2d_supported``."""
    B, Cin, H, W = x.shape
    if out is None:
        out = new_act(B, pc2.cout, H // 2, W // 2, x.dtype, x.device)
    assert tuple(out.shape) == (B, pc2.cout, H // 2, W // 2)
    xp, xld = view_of(x)
    yp, yld = view_of(out)
    L = _lib.lib()
    flops = 2.0 * B * H * W * pc1.cout * Cin * 9 + 2.0 * B * (H // 2) * (W // 2) * pc2.cout * pc2.cin * 4
    nbytes = x.element_size() * (B * H * W * Cin + B * (H // 2) * (W // 2) * pc2.cout)
    tag = f"{Cin}->{pc1.cout} k3s1 + s2d {4 * pc1.cout}->{pc2.cout} {H}x{W} B{B}"

    def launch(keep=(x, out, pc1, pc2)):
        check(L.lpc_conv3x3_s2d_tc(xp, xld, B, H, W, Cin, _fp(pc1.w_tc), _fp(pc1.bias), pc1.cout, pc1.act, _fp(pc2.w_tc), _fp(pc2.bias),
                                   pc2.cout, pc2.act, yp, yld, _stream()), "conv3x3_s2d_tc")
    if REPLAY is not None:
        REPLAY.append(("conv2d_tc", launch, flops, nbytes, tag, None))
    with _prof("conv3x3_s2d_tc", flops, nbytes, tag):
        launch()
    return out


@_profiled
def dwconv2d(x, pd, out=None, res=None):
    B, Cc, H, W = x.shape
    ke = pd.d * (pd.k - 1) + 1
    Ho = (H + 2 * pd.p - ke) // pd.s + 1
    Wo = (W + 2 * pd.p - ke) // pd.s + 1
    if out is None:
        out = new_act(B, Cc, Ho, Wo, x.dtype, x.device)
    assert tuple(out.shape) == (B, Cc, Ho, Wo)
    xp, xld = view_of(x)
    yp, yld = view_of(out)
    rp, rld = (0, 0) if res is None else view_of(res)
    check(_lib.lib().lpc_dwconv2d(dt_code(x.dtype), xp, xld, B, H, W, Cc, _fp(pd.w), _fp(pd.bias), pd.k, pd.s, pd.p, pd.d,
                                  yp, yld, pd.act, rp or None, rld, _stream()), "dwconv2d")
    return out


# "1": use lpc_dwpw_tc wherever the kernel takes the shape; "auto" (default): only where it measured faster than the unfused
# chain on B200 (single pointwise stage, one 64-channel block, large maps: dw64->80 @80x80 B64 44 vs 53 us; every other
# head shape loses 10-70 %, profiles/r02_e_dwpw.md - the kernel is issue-bound at 20 warps per SM); "0": never.
FUSE_DWPW = os.environ.get("LPC_FUSE_DWPW", "auto")


def dwpw_supported(x, pd, pc1, pc2=None, out_ld=None):
    """Can depthwise ``pd`` -> pointwise ``pc1`` [-> pointwise ``pc2``] on x run as one lpc_dwpw_tc launch?"""
    if FUSE_DWPW in ("0", False) or x.dtype != torch.bfloat16 or not x.is_cuda:
        return False
    if FUSE_DWPW == "auto" and not (pc2 is None and pd.c <= 64 and x.shape[0] * x.shape[2] * x.shape[3] >= 128 * 1024):
        return False
    if not (pd.k == 3 and pd.s == 1 and pd.p == 1 and pd.d == 1 and pd.c == x.shape[1]):
        return False
    for pc in (pc1, pc2):
        if pc is not None and not (pc.k == 1 and pc.s == 1 and pc.p == 0 and pc.w_tc is not None):
            return False
    if pc1.cin != pd.c or (pc2 is not None and pc2.cin != pc1.cout):
        return False
    xp, xld = view_of(x)
    if xp % 16:
        return False
    cl = pc2.cout if pc2 is not None else pc1.cout
    return bool(_lib.lib().lpc_dwpw_tc_supported(pd.c, pc1.cout, pc2.cout if pc2 is not None else 0, xld, out_ld if out_ld is not None else cl,
                                                 pd.act, pc1.act, pc2.act if pc2 is not None else ACT_NONE))


def dwpw(x, pd, pc1, pc2=None, out=None, rowmax=None):
    """act2(W2 act1(W1 dw_act(dw3x3(x)))) in one kernel (lpc_dwpw_tc); callers check ``dwpw_supported`` first."""
    B, Cin, H, W = x.shape
    cl = pc2.cout if pc2 is not None else pc1.cout
    if out is None:
        out = new_act(B, cl, H, W, x.dtype, x.device)
    assert tuple(out.shape) == (B, cl, H, W)
    xp, xld = view_of(x)
    yp, yld = view_of(out)
    L = _lib.lib()
    rm = (None, 0, 0)
    if rowmax is not None:
        rm = (C.c_void_p(rowmax["ws"].data_ptr()), int(rowmax["A"]), int(rowmax["off"]))
    flops = 2.0 * B * H * W * (9 * Cin + Cin * pc1.cout + (pc1.cout * pc2.cout if pc2 is not None else 0))
    nbytes = x.element_size() * B * H * W * (Cin + cl)
    tag = f"dw{Cin}->{pc1.cout}" + (f"->{pc2.cout}" if pc2 is not None else "") + f" {H}x{W} B{B}"

    def launch(keep=(x, out, pd, pc1, pc2, rowmax)):
        check(L.lpc_dwpw_tc(xp, xld, B, H, W, Cin, _fp(pd.w), _fp(pd.bias), pd.act, _fp(pc1.w_tc), _fp(pc1.bias), pc1.cout, pc1.act,
                            _fp(pc2.w_tc) if pc2 is not None else None, _fp(pc2.bias) if pc2 is not None else None,
                            pc2.cout if pc2 is not None else 0, pc2.act if pc2 is not None else ACT_NONE, yp, yld, rm[0], rm[1], rm[2], _stream()),
              "dwpw_tc")
    if REPLAY is not None:
        REPLAY.append(("dwpw_tc", launch, flops, nbytes, tag, None))
    with _prof("dwpw_tc", flops, nbytes, tag):
        launch()
    if rowmax is not None:
        rowmax["ok"] = rowmax.get("ok", True)
    return out


@_profiled
def sppf_pool(x, out):
    """out (3C channels) <- [pool5(x), pool9(x), pool13(x)]."""
    B, Cc, H, W = x.shape
    assert out.shape[1] == 3 * Cc
    xp, xld = view_of(x)
    yp, yld = view_of(out)
    check(_lib.lib().lpc_sppf_pool(dt_code(x.dtype), xp, xld, B, H, W, Cc, yp, yld, _stream()), "sppf_pool")
    return out


@_profiled
def psa_attention(qkv, heads, kd, hd, out=None):
    B, Ct, H, W = qkv.shape
    assert Ct == heads * (2 * kd + hd)
    if out is None:
        out = new_act(B, heads * hd, H, W, qkv.dtype, qkv.device)
    xp, xld = view_of(qkv)
    yp, yld = view_of(out)
    check(_lib.lib().lpc_psa_attention(dt_code(qkv.dtype), xp, xld, B, H * W, heads, kd, hd, yp, yld, _stream()), "psa_attention")
    return out


@_profiled
def upsample2x(x, out=None):
    B, Cc, H, W = x.shape
    if out is None:
        out = new_act(B, Cc, 2 * H, 2 * W, x.dtype, x.device)
    assert tuple(out.shape) == (B, Cc, 2 * H, 2 * W)
    xp, xld = view_of(x)
    yp, yld = view_of(out)
    check(_lib.lib().lpc_upsample2x(dt_code(x.dtype), xp, xld, B, H, W, Cc, yp, yld, _stream()), "upsample2x")
    return out


@_profiled
def copy_channels(x, out):
    B, Cc, H, W = x.shape
    assert tuple(out.shape) == tuple(x.shape)
    xp, xld = view_of(x)
    yp, yld = view_of(out)
    check(_lib.lib().lpc_copy_channels(dt_code(x.dtype), xp, xld, B * H * W, Cc, yp, yld, _stream()), "copy_channels")
    return out


@_profiled
def space_to_depth(x, out=None):
    B, Cc, H, W = x.shape
    if out is None:
        out = new_act(B, 4 * Cc, H // 2, W // 2, x.dtype, x.device)
    xp, xld = view_of(x)
    yp, yld = view_of(out)
    check(_lib.lib().lpc_space_to_depth(dt_code(x.dtype), xp, xld, B, H, W, Cc, yp, yld, _stream()), "space_to_depth")
    return out


@_profiled
def channel_deinterleave(x, out=None):
    B, Cc, H, W = x.shape
    if out is None:
        out = new_act(B, Cc, H, W, x.dtype, x.device)
    xp, xld = view_of(x)
    yp, yld = view_of(out)
    check(_lib.lib().lpc_channel_deinterleave(dt_code(x.dtype), xp, xld, B * H * W, Cc, yp, yld, _stream()), "channel_deinterleave")
    return out


@_profiled
def global_avgpool(x):
    """-> (partial sums [B, chunks, C] fp32, scale = 1/HW); channel_mlp finishes the reduction."""
    B, Cc, H, W = x.shape
    L = _lib.lib()
    chunks = L.lpc_global_avgpool_chunks(B, H * W)
    out = torch.empty((B, chunks, Cc), dtype=torch.float32, device=x.device)
    xp, xld = view_of(x)
    check(L.lpc_global_avgpool(dt_code(x.dtype), xp, xld, B, H * W, Cc, _fp(out), _stream()), "global_avgpool")
    return out, 1.0 / (H * W)


@_profiled
def channel_mlp(v, W1, b1, act1, W2=None, b2=None, act2=ACT_NONE, scale=1.0):
    """v: [B, C0] or partial sums [B, parts, C0] (summed and scaled by ``scale`` on load)."""
    if v.dim() == 2:
        v = v.unsqueeze(1)
    B, parts, C0 = v.shape
    C1 = W1.shape[0]
    C2 = W2.shape[0] if W2 is not None else 0
    out = torch.empty((B, C2 if W2 is not None else C1), dtype=torch.float32, device=v.device)
    check(_lib.lib().lpc_channel_mlp(_fp(v), B, parts, float(scale), C0, _fp(W1), _fp(b1), C1, act1, _fp(W2), _fp(b2), C2, act2,
                                     _fp(out), _stream()), "channel_mlp")
    return out


def pooled_gate(x, W1, b1, act1, W2=None, b2=None, act2=ACT_NONE):
    """act2(W2 act1(W1 avgpool(x) + b1) + b2): CBAM's channel attention / SPCA's gate."""
    part, scale = global_avgpool(x)
    return channel_mlp(part, W1, b1, act1, W2, b2, act2, scale)


@_profiled
def cbam_spatial(x, ca, w7, k, out=None):
    """y = x*ca*sigmoid(conv_kxk([mean_c(x*ca), max_c(x*ca)]))."""
    B, Cc, H, W = x.shape
    if out is None:
        out = new_act(B, Cc, H, W, x.dtype, x.device)
    stats = torch.empty((B, H * W, 2), dtype=torch.float32, device=x.device)
    xp, xld = view_of(x)
    yp, yld = view_of(out)
    L = _lib.lib()
    check(L.lpc_cbam_stats(dt_code(x.dtype), xp, xld, B, H * W, Cc, _fp(ca), _fp(stats), _stream()), "cbam_stats")
    check(L.lpc_cbam_apply(dt_code(x.dtype), xp, xld, B, H, W, Cc, _fp(ca), _fp(stats), _fp(w7), k, yp, yld, _stream()), "cbam_apply")
    return out


# ---- v10Detect tail ---------------------------------------------------------------------------------------
def _raw_args(raw, strides):
    r0, r1, r2 = raw
    B, Ct, H0, W0 = r0.shape
    ptrs, lds = zip(*(view_of(r) for r in raw))
    if len(set(lds)) != 1:
        raise LpcError("the three head maps must share one pixel pitch")
    assert tuple(r1.shape[2:]) == (H0 // 2, W0 // 2) and tuple(r2.shape[2:]) == (H0 // 4, W0 // 4)
    st = (C.c_float * 3)(*[float(s) for s in strides])
    return ptrs, lds[0], B, Ct, H0, W0, st


@_profiled
def v10_decode(raw, strides, nc):
    """Detect.inference: three NHWC head maps -> y [B, 4+nc, A] fp32."""
    ptrs, ld, B, Ct, H0, W0, st = _raw_args(raw, strides)
    assert Ct == 64 + nc
    A = H0 * W0 + (H0 // 2) * (W0 // 2) + (H0 // 4) * (W0 // 4)
    y = torch.empty((B, 4 + nc, A), dtype=torch.float32, device=raw[0].device)
    check(_lib.lib().lpc_v10_decode(dt_code(raw[0].dtype), ptrs[0], ptrs[1], ptrs[2], ld, B, H0, W0, nc, st, _fp(y), _stream()), "v10_decode")
    return y


def topk_workspace(B, A, max_det, device):
    """Workspace of the fused tail (per-anchor keys first): allocate it BEFORE the head convs to let them fill the keys."""
    n = _lib.lib().lpc_v10_topk_workspace_bytes(B, A, max_det)
    return torch.empty((n,), dtype=torch.uint8, device=device)


def v10_decode_topk(raw, strides, nc, max_det=300, img_hw=None, return_index=False, ws=None, keys_ready=False, scale_back=None, out=None):
    """Fused decode + v10postprocess + xywh2xyxy (+clip): -> dets [B,K,6] fp32 (x1,y1,x2,y2,score,label).
    ``scale_back``: optional fp32 device tensor [B,5] = (pad_x, pad_y, gain, orig_w, orig_h) per image: the predictor's
    ops.scale_boxes + clip_boxes (utils/ops.py:89-124, 305-324) applied in the same kernel."""
    ptrs, ld, B, Ct, H0, W0, st = _raw_args(raw, strides)
    assert Ct == 64 + nc
    A = H0 * W0 + (H0 // 2) * (W0 // 2) + (H0 // 4) * (W0 // 4)
    dev = raw[0].device
    L = _lib.lib()
    ws_bytes = L.lpc_v10_topk_workspace_bytes(B, A, max_det)
    if ws is None:
        ws, keys_ready = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev), False
    assert ws.numel() >= ws_bytes
    dets = out if out is not None else torch.empty((B, max_det, 6), dtype=torch.float32, device=dev)
    assert tuple(dets.shape) == (B, max_det, 6) and dets.dtype == torch.float32 and dets.is_contiguous()
    aidx = torch.empty((B, max_det), dtype=torch.int32, device=dev) if return_index else None
    ih, iw = (img_hw if img_hw is not None else (0, 0))
    # algorithmic bytes: raw maps read once + detections written
    nbytes = B * A * (64 + nc) * raw[0].element_size() + B * max_det * 6 * 4
    if scale_back is not None:
        assert scale_back.is_cuda and scale_back.dtype == torch.float32 and tuple(scale_back.shape) == (B, 5) and scale_back.is_contiguous()

    def launch(keep=(raw, ws, dets, aidx, scale_back)):
        check(L.lpc_v10_decode_topk_scaled(dt_code(raw[0].dtype), ptrs[0], ptrs[1], ptrs[2], ld, B, H0, W0, nc, st, max_det, ih, iw,
                                           _fp(ws), ws_bytes, int(bool(keys_ready)), _fp(scale_back), _fp(dets), _fp(aidx), _stream()),
              "v10_decode_topk")
    if REPLAY is not None:
        REPLAY.append(("v10_decode_topk", launch, 0.0, nbytes, ""))
    with _prof("v10_decode_topk", 0.0, nbytes):
        launch()
    return (dets, aidx) if return_index else dets


@_profiled
def v10_postprocess(preds, max_det, nc):
    """ops.v10postprocess on preds [B,A,4+nc] fp32 (any strides) -> boxes [B,K,4], scores [B,K], labels [B,K] i64."""
    if not preds.is_cuda:
        raise LpcError("v10postprocess runs on CUDA tensors only")
    if preds.dtype != torch.float32:
        preds = preds.float()
    B, A, Ct = preds.shape
    assert Ct == 4 + nc
    dev = preds.device
    L = _lib.lib()
    ws_bytes = L.lpc_v10_topk_workspace_bytes(B, A, max_det)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    boxes = torch.empty((B, max_det, 4), dtype=torch.float32, device=dev)
    scores = torch.empty((B, max_det), dtype=torch.float32, device=dev)
    labels = torch.empty((B, max_det), dtype=torch.int64, device=dev)
    sb, sa, sc = preds.stride()
    check(L.lpc_v10_postprocess(_fp(preds), sb, sa, sc, B, A, nc, max_det, _fp(ws), ws_bytes, _fp(boxes), _fp(scores),
                                _fp(labels), _stream()), "v10_postprocess")
    return boxes, scores, labels
