"""Time the fused v10 tail alone (select_decode with the keys ready, and amax_keys + select_decode) on synthetic raw
head maps, CUDA-graph replays bracketed by CUDA events; prints the per-phase clock64 deltas of CTA 0 when the library
was built with -DLPC_TAIL_CLOCKS (tools only).  A/B two builds with LPC_LIB=<path>.

  python tools/run_tail.py [B] [S] [--real]     # --real: raw maps from the LPC model on a seeded batch
"""
import ctypes as C
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
pkg = importlib.import_module("lpc-yolo_b200")
Fn = importlib.import_module("lpc-yolo_b200.functional")


def graph_time(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    B = int(args[0]) if args else 64
    S = int(args[1]) if len(args) > 1 else 640
    nc, K = 80, 300
    torch.manual_seed(0)
    if "--real" in sys.argv:
        import lpc_oracle as O
        om = O.build("lpc")
        model = pkg.YOLOv10DetectionModel(O.MODEL_FILES["lpc"])
        model.load_state_dict(om.sd, strict=True)
        model = model.cuda().eval()
        model.compute_dtype = torch.bfloat16
        with torch.no_grad():
            raw = model(torch.rand(B, 3, S, S).cuda())["one2one"][1]
    else:
        raw = []
        for l in range(3):
            h = S // (8 << l)
            t = Fn.new_act(B, 64 + nc, h, h, torch.bfloat16, "cuda")
            t.copy_(torch.randn(B, 64 + nc, h, h, device="cuda") * 1.5 - 3.0)
            raw.append(t)
    A = sum((S // (8 << l)) ** 2 for l in range(3))
    L = pkg.lib()
    ws = torch.empty((L.lpc_v10_topk_workspace_bytes(B, A, K),), dtype=torch.uint8, device="cuda")
    ref = Fn.v10_decode_topk(raw, [8.0, 16.0, 32.0], nc, K, (S, S), ws=ws, keys_ready=False)   # fills the keys
    t_sel = graph_time(lambda: Fn.v10_decode_topk(raw, [8.0, 16.0, 32.0], nc, K, (S, S), ws=ws, keys_ready=True))
    t_all = graph_time(lambda: Fn.v10_decode_topk(raw, [8.0, 16.0, 32.0], nc, K, (S, S), ws=ws, keys_ready=False))
    got = Fn.v10_decode_topk(raw, [8.0, 16.0, 32.0], nc, K, (S, S), ws=ws, keys_ready=True)
    torch.cuda.synchronize()
    assert torch.equal(ref, got)
    nbytes = B * A * (64 + nc) * 2 + B * K * 6 * 4
    print(f"lib {os.environ.get('LPC_LIB', 'in-tree')}  B={B} S={S}: select_decode {t_sel:.1f} us "
          f"({nbytes / t_sel / 1e3:.0f} GB/s algorithmic), amax_keys+select_decode {t_all:.1f} us; "
          f"checksum {ref.double().sum().item():.6f}")
    try:
        fn = C.CDLL(os.environ.get("LPC_LIB") or os.path.join(ROOT, "lpc-yolo_b200", "csrc", "liblpcyolo.so")).lpc_debug_tail_clocks
    except AttributeError:
        return
    buf = (C.c_longlong * 16)()
    assert fn(buf) == 0
    names = ["load keys + digit-0 histogram", "select 1", "collect 1", "gather candidates", "cut candidates", "rank", "decode"]
    c = list(buf)
    print("CTA 0 phase cycles: " + ", ".join(f"{n} {c[i + 1] - c[i]}" for i, n in enumerate(names)) + f"; total {c[len(names)] - c[0]}; stage-2 candidates {c[15]}")


if __name__ == "__main__":
    main()
