"""Leave-one-out cost of every launch of one step: the library handle is replaced by a proxy that turns exactly ONE C call (the
k-th) into a no-op, the step is captured into a CUDA graph and timed, for k = 1..n.  full - without_k = what launch k
contributes to the CRITICAL PATH of the graph (launches on forked side streams that hide under another chain contribute
~0 although their prefix-difference in tools/prefix_times.py is several microseconds).  The skipped launch leaves its output
uninitialised; the kernels' timing does not depend on the values (the tail's does slightly).

    python tools/leave_one_out.py [model] [batch] [size] > gpurun_out/loo_lpc_b64.csv
"""
import ctypes
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

pkg = importlib.import_module("lpc-yolo_b200")
Fn = importlib.import_module("lpc-yolo_b200.functional")
libmod = importlib.import_module("lpc-yolo_b200._lib")
synth = importlib.import_module("lpc-yolo_b200.utils.synth")

LAUNCHING = {"lpc_conv2d_direct", "lpc_conv2d_tc", "lpc_conv3x3_s2d_tc", "lpc_conv1x1_up2cat_tc", "lpc_stem_conv_u8", "lpc_dwpw_tc", "lpc_conv2d_tc_rowmax", "lpc_stem_conv", "lpc_dwconv2d", "lpc_sppf_pool", "lpc_psa_attention",
             "lpc_upsample2x", "lpc_copy_channels", "lpc_space_to_depth", "lpc_channel_deinterleave", "lpc_pack_input", "lpc_pack_u8",
             "lpc_letterbox_u8", "lpc_global_avgpool", "lpc_channel_mlp", "lpc_cbam_stats", "lpc_cbam_apply", "lpc_v10_decode",
             "lpc_v10_decode_topk", "lpc_v10_decode_topk_keys", "lpc_v10_decode_topk_scaled", "lpc_v10_postprocess"}


class Proxy:
    def __init__(self, real):
        self.real, self.limit, self.count, self.log = real, 1 << 30, 0, []

    def __getattr__(self, name):
        fn = getattr(self.real, name)
        if name not in LAUNCHING:
            return fn

        def call(*a):
            self.count += 1
            if self.limit == 1 << 30:                  # the first, complete pass names every launch
                self.log.append((name, [int(v) for v in a if isinstance(v, int) and not isinstance(v, bool) and abs(v) < 100000]))
            if self.count == self.limit:
                return 0
            return fn(*a)
        return call


def main():
    name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "lpc"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 640
    every = int(sys.argv[sys.argv.index("--every") + 1]) if "--every" in sys.argv else 1
    yolo = pkg.YOLO(bench.FILES[name])
    synth.init_synthetic(yolo.model)
    m = yolo.model.cuda().eval()
    m.compute_dtype = torch.bfloat16
    g = torch.Generator().manual_seed(1)
    x = Fn.pack_u8(torch.randint(0, 256, (B, S, S, 3), generator=g, dtype=torch.uint8).cuda(), torch.bfloat16)
    real = libmod.lib()
    px = Proxy(real)
    libmod._lib = px
    with torch.no_grad():
        m.detect(x, 300)
        torch.cuda.synchronize()
        n = px.count
        log = list(px.log)

        def timed(limit, reps=20):
            px.limit = limit
            px.count = 0
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                m.detect(x, 300)
            torch.cuda.current_stream().wait_stream(s)
            px.count = 0
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                m.detect(x, 300)
            for _ in range(3):
                gr.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                gr.replay()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps * 1e3
        full = min(timed(1 << 29), timed(1 << 29))
        print(f"# full step {full:.1f} us, {n} launches")
        print("idx,call,int_args,without_us,critical_us")
        for k in range(1, n + 1):
            t = timed(k)
            nm, ints = log[k - 1] if k - 1 < len(log) else ("?", [])
            print(f"{k},{nm},{' '.join(map(str, ints))},{t:.1f},{full - t:.1f}", flush=True)


if __name__ == "__main__":
    main()
