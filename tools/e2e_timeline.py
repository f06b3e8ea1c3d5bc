"""Device timeline of one end-to-end predict() step (uint8 host source): where the chunks' H2D copies and graph replays
start and end, from CUDA events (no nsys in the image).  python tools/e2e_timeline.py [batch]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
pkg = importlib.import_module("lpc-yolo_b200")
eng = importlib.import_module("lpc-yolo_b200.engine")
synth = importlib.import_module("lpc-yolo_b200.utils.synth")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
yolo = pkg.YOLO(bench.FILES["lpc"]); synth.init_synthetic(yolo.model)
x = torch.randint(0, 256, (B, 640, 640, 3), dtype=torch.uint8).pin_memory()
xn = x.numpy()
for _ in range(3):
    yolo.predict(xn, conf=0.25, half=True, imgsz=640)
torch.cuda.synchronize()
P = yolo.predictor
plan = eng._chunk_plan(B)
gds = [v for k, v in P._graphed.items() if k[0] == "u8"]
bysize = {g.inp[0].shape[0]: g for g in gds}
cur = torch.cuda.current_stream(); cs = P._copy_stream
E_ = lambda: torch.cuda.Event(enable_timing=True)
for rep in range(3):
    torch.cuda.synchronize()
    t_host0 = time.perf_counter()
    start = E_(); start.record(cur); cs.wait_stream(cur)
    marks = []; lo = 0
    for cb in plan:
        gd = bysize[cb]
        c0, c1, g0, g1 = E_(), E_(), E_(), E_()
        with torch.cuda.stream(cs):
            c0.record(cs); gd.inp[0].copy_(x[lo:lo + cb], non_blocking=True); c1.record(cs)
        cur.wait_event(c1)
        g0.record(cur); gd.graphs[0].replay(); g1.record(cur)
        marks.append((cb, c0, c1, g0, g1)); lo += cb
    t_host1 = time.perf_counter()
    torch.cuda.synchronize()
    t_host2 = time.perf_counter()
    if rep == 2:
        print(f"plan {plan}: host enqueue {1e3*(t_host1-t_host0):.3f} ms, host until GPU idle {1e3*(t_host2-t_host0):.3f} ms")
        for cb, c0, c1, g0, g1 in marks:
            print(f"  chunk {cb:3d}: copy {start.elapsed_time(c0):.3f} -> {start.elapsed_time(c1):.3f} ms   graph {start.elapsed_time(g0):.3f} -> {start.elapsed_time(g1):.3f} ms ({g0.elapsed_time(g1):.3f})")
# the same graphs replayed with nothing else running
for cb in sorted(bysize):
    gd = bysize[cb]; a, b = E_(), E_()
    torch.cuda.synchronize(); a.record()
    for _ in range(10): gd.graphs[0].replay()
    b.record(); torch.cuda.synchronize()
    print(f"graph B={cb}: {a.elapsed_time(b)/10:.3f} ms alone")
t0 = time.perf_counter()
for _ in range(10):
    r = yolo.predict(xn, conf=0.25, half=True, imgsz=640)
torch.cuda.synchronize()
print(f"predict(): {(time.perf_counter()-t0)*100:.3f} ms/step; speed {r[0].speed}")
