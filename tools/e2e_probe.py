"""Where the end-to-end predict() time goes: python tools/e2e_probe.py [batch]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
pkg = importlib.import_module("lpc-yolo_b200")
synth = importlib.import_module("lpc-yolo_b200.utils.synth")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
yolo = pkg.YOLO(bench.FILES["lpc"]); synth.init_synthetic(yolo.model)
x = torch.randint(0, 256, (B, 640, 640, 3), dtype=torch.uint8).pin_memory()
xn = x.numpy()
for _ in range(3):
    yolo.predict(xn, conf=0.25, half=True, imgsz=640)
torch.cuda.synchronize()
t0 = time.perf_counter(); n = 10
for _ in range(n):
    r = yolo.predict(xn, conf=0.25, half=True, imgsz=640); h = yolo.predictor.last_preds.cpu()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
print(f"chunks={os.environ.get('LPC_E2E_CHUNKS','auto')} B={B}: {dt*1e3:.3f} ms/step  {B/dt:.0f} img/s  speed={r[0].speed}")
# raw H2D bandwidth
d = torch.empty_like(x, device="cuda"); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 10
print(f"H2D {x.numel()/1e6:.1f} MB in {t*1e3:.3f} ms = {x.numel()/t/1e9:.1f} GB/s")
