"""Run the stem conv a few times: python tools/run_stem.py cout B [S]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("lpc-yolo_b200")
M = importlib.import_module("lpc-yolo_b200.nn.modules")
Fn = importlib.import_module("lpc-yolo_b200.functional")
c2, B = int(sys.argv[1]), int(sys.argv[2])
S = int(sys.argv[3]) if len(sys.argv) > 3 else 640
mod = M.Conv(3, c2, 3, 2).cuda().eval()
x = Fn.pack_input(torch.rand(B, 3, S, S).cuda(), torch.bfloat16)
with torch.no_grad():
    for _ in range(5):
        y = mod(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        y = mod(x)
    b.record(); torch.cuda.synchronize()
us = a.elapsed_time(b) * 100
by = 2 * (B * S * S * 4 + y.numel())
print(f"stem 3->{c2} {S}x{S} B{B}: {us:.1f} us  {by/us/1e3:.0f} GB/s")
