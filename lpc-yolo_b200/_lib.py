"""ctypes binding of liblpcyolo.so (the C ABI declared in include/lpcyolo.h).

There is no CPU fallback: if the library cannot be built/loaded, or an op is asked to run on a
non-CUDA tensor, the call raises.  The library is built in-tree (csrc/build.py) so the binary that the
tests and the bench load is the one shipped with the repository snapshot.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "liblpcyolo.so")
if os.environ.get("LPC_LIB"):            # A/B measurements of two builds in one process tree (tools only)
    LIB_PATH = os.environ["LPC_LIB"]

BF16, F32 = 0, 1
E_ARG, E_UNSUPPORTED, E_CUDA, E_WORKSPACE = -1, -2, -3, -4
ACT_NONE, ACT_SILU, ACT_MISH, ACT_SIGMOID, ACT_RELU = 0, 1, 2, 3, 4

_p, _i, _ll, _f32p, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_size_t

# name -> (restype, argtypes); mirrors include/lpcyolo.h one to one (tests/test_host_cpu.py::test_abi_exports_every_declared_symbol checks this)
SIGNATURES = {
    "lpc_abi_version": (_i, []),
    "lpc_last_error": (C.c_char_p, []),
    "lpc_device_arch": (_i, []),
    "lpc_launch_count": (C.c_ulonglong, []),
    "lpc_plan_begin": (_i, []),
    "lpc_plan_end": (_i, [C.POINTER(C.c_void_p)]),
    "lpc_plan_size": (_i, [_p]),
    "lpc_plan_wait": (_i, [_p, _p]),
    "lpc_plan_run": (_i, [_p, _p]),
    "lpc_plan_run_graph": (_i, [_p, _p]),
    "lpc_plan_destroy": (None, [_p]),
    "lpc_conv2d_direct": (_i, [_i, _p, _i, _i, _i, _i, _i, _p, _f32p, _i, _i, _i, _i, _p, _i, _i, _f32p, _p, _i, _p]),
    "lpc_conv2d_tc": (_i, [_p, _i, _i, _i, _i, _i, _p, _f32p, _i, _i, _i, _i, _p, _i, _i, _f32p, _p, _i, _p]),
    "lpc_conv2d_tc_rowmax": (_i, [_p, _i, _i, _i, _i, _i, _p, _f32p, _i, _i, _i, _i, _p, _i, _i, _f32p, _p, _i, _p, _ll, _i, _p]),
    "lpc_conv2d_tc_kpad": (_i, [_i, _i]),
    "lpc_conv2d_tc_set_mode": (_i, [_i]),
    "lpc_conv2d_tc_supported": (_i, [_i, _i, _i, _i, _i, _i, _i]),
    "lpc_conv1x1_up2cat_tc_supported": (_i, [_i, _i, _i, _i, _i, _i, _i, _i]),
    "lpc_conv1x1_up2cat_tc": (_i, [_p, _i, _i, _p, _i, _i, _i, _i, _i, _p, _f32p, _i, _p, _i, _i, _p]),
    "lpc_conv3x3_s2d_tc_supported": (_i, [_i, _i, _i, _i, _i, _i, _i]),
    "lpc_conv3x3_s2d_tc": (_i, [_p, _i, _i, _i, _i, _i, _p, _f32p, _i, _i, _p, _f32p, _i, _i, _p, _i, _p]),
    "lpc_dwpw_tc_supported": (_i, [_i, _i, _i, _i, _i, _i, _i, _i]),
    "lpc_dwpw_tc": (_i, [_p, _i, _i, _i, _i, _i, _f32p, _f32p, _i, _p, _f32p, _i, _i, _p, _f32p, _i, _i, _p, _i, _p, _ll, _i, _p]),
    "lpc_stem_conv": (_i, [_i, _p, _i, _i, _i, _f32p, _f32p, _i, _i, _p, _i, _i, _p]),
    "lpc_stem_conv_u8_supported": (_i, [_i, _i, _i, _i, _i, _i]),
    "lpc_stem_conv_u8": (_i, [_p, _i, _i, _i, _i, _f32p, _f32p, _i, _i, _p, _i, _i, _p]),
    "lpc_dwconv2d": (_i, [_i, _p, _i, _i, _i, _i, _i, _f32p, _f32p, _i, _i, _i, _i, _p, _i, _i, _p, _i, _p]),
    "lpc_sppf_pool": (_i, [_i, _p, _i, _i, _i, _i, _i, _p, _i, _p]),
    "lpc_psa_attention": (_i, [_i, _p, _i, _i, _i, _i, _i, _i, _p, _i, _p]),
    "lpc_upsample2x": (_i, [_i, _p, _i, _i, _i, _i, _i, _p, _i, _p]),
    "lpc_copy_channels": (_i, [_i, _p, _i, _ll, _i, _p, _i, _p]),
    "lpc_space_to_depth": (_i, [_i, _p, _i, _i, _i, _i, _i, _p, _i, _p]),
    "lpc_channel_deinterleave": (_i, [_i, _p, _i, _ll, _i, _p, _i, _p]),
    "lpc_pack_input": (_i, [_i, _f32p, _i, _i, _i, _i, _p, _i, _i, _p]),
    "lpc_pack_u8": (_i, [_i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "lpc_letterbox_u8": (_i, [_i, _p, _i, _i, _i, _i, _i, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p]),
    "lpc_global_avgpool": (_i, [_i, _p, _i, _i, _i, _i, _f32p, _p]),
    "lpc_global_avgpool_chunks": (_i, [_i, _i]),
    "lpc_channel_mlp": (_i, [_f32p, _i, _i, C.c_float, _i, _f32p, _f32p, _i, _i, _f32p, _f32p, _i, _i, _f32p, _p]),
    "lpc_cbam_stats": (_i, [_i, _p, _i, _i, _i, _i, _f32p, _f32p, _p]),
    "lpc_cbam_apply": (_i, [_i, _p, _i, _i, _i, _i, _i, _f32p, _f32p, _f32p, _i, _p, _i, _p]),
    "lpc_v10_topk_workspace_bytes": (_sz, [_i, _i, _i]),
    "lpc_v10_decode": (_i, [_i, _p, _p, _p, _i, _i, _i, _i, _i, C.POINTER(C.c_float), _f32p, _p]),
    "lpc_v10_decode_topk": (_i, [_i, _p, _p, _p, _i, _i, _i, _i, _i, C.POINTER(C.c_float), _i, _i, _i, _p, _sz,
                                 _f32p, _p, _p]),
    "lpc_v10_decode_topk_keys": (_i, [_i, _p, _p, _p, _i, _i, _i, _i, _i, C.POINTER(C.c_float), _i, _i, _i, _p, _sz, _i,
                                      _f32p, _p, _p]),
    "lpc_v10_decode_topk_scaled": (_i, [_i, _p, _p, _p, _i, _i, _i, _i, _i, C.POINTER(C.c_float), _i, _i, _i, _p, _sz, _i,
                                        _f32p, _f32p, _p, _p]),
    "lpc_make_anchors": (_i, [_i, C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_float, _f32p, _f32p, _p]),
    "lpc_dist2bbox": (_i, [_f32p, _f32p, _ll, _ll, _i, _f32p, _p]),
    "lpc_scale_boxes": (_i, [_f32p, _ll, _i, _i, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _p]),
    "lpc_v10_postprocess": (_i, [_f32p, _ll, _ll, _ll, _i, _i, _i, _i, _p, _sz, _f32p, _f32p, _p, _p]),
}

_lib = None


class LpcError(RuntimeError):
    pass


def _build_mod():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_lpc_build", os.path.join(_HERE, "csrc", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build(force=False):
    return _build_mod().build(force=force)


def lib():
    """Load the library and return the ctypes handle.  The binary must correspond to the sources in the tree: with nvcc
    present ``build()`` runs first (a sha1 stamp over the sources makes it a no-op when nothing changed); without nvcc the
    shipped binary's stamp is compared with the sources and a mismatch raises (LPC_ALLOW_STALE=1 downgrades it to a warning)."""
    global _lib
    if _lib is None:
        if not os.environ.get("LPC_LIB"):
            bm = _build_mod()
            import shutil
            if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
                bm.build(force=bool(os.environ.get("LPC_REBUILD")))
            elif not os.path.exists(LIB_PATH):
                raise LpcError(f"{LIB_PATH} is missing and there is no nvcc to build it (no fallback path exists)")
            else:
                stamp = open(bm.STAMP).read().strip() if os.path.exists(bm.STAMP) else None
                if stamp != bm._digest():
                    msg = f"{LIB_PATH} was not built from the sources in this tree (stamp {stamp} != {bm._digest()})"
                    if not os.environ.get("LPC_ALLOW_STALE"):
                        raise LpcError(msg + "; rebuild it or set LPC_ALLOW_STALE=1")
                    import warnings
                    warnings.warn(msg)
        try:
            handle = C.CDLL(LIB_PATH)
        except OSError as e:  # loud failure: there is no fallback path
            raise LpcError(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.lpc_abi_version() != 1:
            raise LpcError("liblpcyolo.so ABI version mismatch")
        _lib = handle
    return _lib


def check(status, what=""):
    if status != 0:
        raise LpcError(f"{what or 'lpc call'} failed ({status}): {lib().lpc_last_error().decode()}")
