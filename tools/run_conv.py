"""Run one conv shape a few times (for ncu): python tools/run_conv.py c1 c2 k s H B [mode] [act]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("lpc-yolo_b200")
M = importlib.import_module("lpc-yolo_b200.nn.modules")
B_ = importlib.import_module("lpc-yolo_b200.nn.modules.block")
Fn = importlib.import_module("lpc-yolo_b200.functional")
c1, c2, k, s, H, B = map(int, sys.argv[1:7])
mode = int(sys.argv[7]) if len(sys.argv) > 7 else 0
mish = (sys.argv[8] == "mish") if len(sys.argv) > 8 else True
pkg.lib().lpc_conv2d_tc_set_mode(mode)
mod = (B_.Conv if mish else M.Conv)(c1, c2, k, s).cuda().eval()
x = Fn.new_act(B, c1, H, H, torch.bfloat16, "cuda"); x.normal_()
with torch.no_grad():
    for _ in range(3):
        y = mod(x)
    torch.cuda.synchronize()
    # 10 launches captured in one CUDA graph: no host launch cost (tensor-map encoding, ctypes) inside the timed region
    sd = torch.cuda.Stream(); sd.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(sd):
        y = mod(x)
    torch.cuda.current_stream().wait_stream(sd)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            y = mod(x)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record(); torch.cuda.synchronize()
us = a.elapsed_time(b) * 20
Ho = y.shape[2]
by = 2 * (x.numel() + y.numel())
print(f"conv {c1}->{c2} k{k}s{s} {H}x{H} B{B} mode {mode}: {us:.1f} us  {by/us/1e3:.0f} GB/s  {2*B*Ho*Ho*c1*c2*k*k/us/1e6:.1f} TFLOP/s")
