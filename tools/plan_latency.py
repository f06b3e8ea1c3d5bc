"""Batch-1 (or any batch) step latency: torch CUDA graph of the Python-driven step vs the library's launch plan (lpc_plan_run_graph).
    python tools/plan_latency.py [model] [batch] [size]"""
import importlib, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
pkg = importlib.import_module("lpc-yolo_b200")
Fn = importlib.import_module("lpc-yolo_b200.functional")
synth = importlib.import_module("lpc-yolo_b200.utils.synth")
name = sys.argv[1] if len(sys.argv) > 1 else "yolov10b"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
S = int(sys.argv[3]) if len(sys.argv) > 3 else 640
yolo = pkg.YOLO(bench.FILES[name]); synth.init_synthetic(yolo.model)
m = yolo.model.cuda().eval(); m.compute_dtype = torch.bfloat16
x = Fn.pack_u8(torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8).cuda(), torch.bfloat16)


def p50(fn, n=100):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


with torch.no_grad():
    g, out = bench.graph_of(m, x)
    t_torch = p50(g.replay)
    plan = pkg.Plan(m, x, 300)
    t_plan = p50(lambda: plan.run(graph=True))
    t_eager = p50(lambda: plan.run(graph=False), 30)
    same = bool(torch.equal(plan.out, out))
print(f"{name} B={B} @{S}: torch CUDA graph (multi-stream) {t_torch:.4f} ms | library plan graph ({plan.launches} launches, recorded stream structure) {t_plan:.4f} ms | "
      f"plan re-issued launch by launch from C {t_eager:.4f} ms | identical results {same}")
