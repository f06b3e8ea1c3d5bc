"""Time lpc_dwpw_tc against the unfused chain (dwconv -> 1x1 [-> 1x1]) on the LPC head shapes (B = 64).
    python tools/run_dwpw.py [B]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

pkg = importlib.import_module("lpc-yolo_b200")
Fn = importlib.import_module("lpc-yolo_b200.functional")
M = importlib.import_module("lpc-yolo_b200.nn.modules")
head = importlib.import_module("lpc-yolo_b200.nn.modules.head")


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


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    cases = [(64, 80, 0, 80), (80, 80, 80, 80), (192, 80, 0, 40), (80, 80, 80, 40), (384, 80, 0, 20), (80, 80, 80, 20)]
    if len(sys.argv) > 2:
        cases = [cases[int(v)] for v in sys.argv[2].split(",")]
    for Cin, C1, C2, S in cases:
        dw = M.Conv(Cin, Cin, 3, g=Cin).cuda()
        pw = M.Conv(Cin, C1, 1).cuda()
        pl = head._Plain1x1(C1, C2).cuda() if C2 else None
        x = Fn.new_act(B, Cin, S, S, torch.bfloat16, "cuda")
        x.normal_()
        with torch.no_grad():
            pd, p1 = dw._packed(x, dw._build), pw._packed(x, pw._build)
            p2 = pl._packed(x, pl._build) if pl is not None else None
            out = Fn.new_act(B, C2 or C1, S, S, torch.bfloat16, "cuda")
            mid = Fn.new_act(B, Cin, S, S, torch.bfloat16, "cuda")
            mid2 = Fn.new_act(B, C1, S, S, torch.bfloat16, "cuda")
            tf = timed(lambda: Fn.dwpw(x, pd, p1, p2, out=out))

            def unf():
                Fn.dwconv2d(x, pd, out=mid)
                if p2 is None:
                    Fn.conv2d(mid, p1, out=out)
                else:
                    Fn.conv2d(mid, p1, out=mid2)
                    Fn.conv2d(mid2, p2, out=out)
            tu = timed(unf)
        nb = 2.0 * B * S * S * (Cin + (C2 or C1))
        print(f"dw{Cin}->{C1}" + (f"->{C2}" if C2 else "") + f" {S}x{S} B{B}: fused {tf:7.1f} us  unfused {tu:7.1f} us  floor {nb / 6551e3:6.1f} us", flush=True)


if __name__ == "__main__":
    main()
