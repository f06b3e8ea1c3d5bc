"""Time lpc_psa_attention (bf16) on the PSA shapes of the BASELINE configs; LPC_ATT_MMA_SYNC=1 selects the mma.sync kernel.
    python tools/run_attn.py"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

Fn = importlib.import_module("lpc-yolo_b200.functional")


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


for tag, B, heads, N in (("LPC B64 @640", 64, 2, 400), ("yolov10s B256 @640", 256, 4, 400), ("yolov10x B32 @1280", 32, 5, 1600), ("yolov10b B1 @960", 1, 4, 900),
                         ("yolov10b B1 @640", 1, 4, 400), ("yolov10b B1 @320", 1, 4, 100)):
    kd, hd = 32, 64
    Ct = heads * (2 * kd + hd)
    x = (torch.randn(B, 1, N, Ct, device="cuda") * 0.8).to(torch.bfloat16).permute(0, 3, 1, 2)
    out = Fn.new_act(B, heads * hd, 1, N, torch.bfloat16, "cuda")
    with torch.no_grad():
        t = timed(lambda: Fn.psa_attention(x, heads, kd, hd, out=out))
    fl = 2.0 * B * heads * N * N * (kd + hd)
    print(f"{tag}: heads {heads} N {N}: {t:8.1f} us  ({fl / t / 1e6:7.1f} TFLOP/s; exp floor {B * heads * N * N / (148 * 16 * 1.965e3):6.1f} us)", flush=True)
