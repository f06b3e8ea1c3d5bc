"""Time one depthwise conv shape: python tools/run_dw.py C k s H B [dil]"""
import importlib, os, sys, ctypes as C_
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("lpc-yolo_b200")
Fn = importlib.import_module("lpc-yolo_b200.functional")
c, k, s, H, B = map(int, sys.argv[1:6])
d = int(sys.argv[6]) if len(sys.argv) > 6 else 1
x = Fn.new_act(B, c, H, H, torch.bfloat16, "cuda"); x.normal_()
pad = d * (k - 1) // 2
Ho = (H + 2 * pad - d * (k - 1) - 1) // s + 1
y = Fn.new_act(B, c, Ho, Ho, torch.bfloat16, "cuda")
w = torch.randn(k * k, c, device="cuda"); b = torch.randn(c, device="cuda")
L = pkg.lib()
xp, xld = Fn.view_of(x); yp, yld = Fn.view_of(y)
def run():
    r = L.lpc_dwconv2d(0, C_.c_void_p(xp), xld, B, H, H, c, C_.cast(w.data_ptr(), C_.POINTER(C_.c_float)), C_.cast(b.data_ptr(), C_.POINTER(C_.c_float)), k, s, pad, d,
                       C_.c_void_p(yp), yld, 1, None, 0, C_.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert r == 0, pkg.last_error() if hasattr(pkg, "last_error") else r
for _ in range(3): run()
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): run()
e.record(); torch.cuda.synchronize()
us = a.elapsed_time(e) * 50
by = 2 * (x.numel() + y.numel())
print(f"dw C{c} k{k}s{s}d{d} {H}x{H} B{B}: {us:.1f} us  {by/us/1e3:.0f} GB/s")
