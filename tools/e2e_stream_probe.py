"""Streamed end-to-end throughput (predict(stream=True, batch=B)) with the host time per step beside it.

    python tools/e2e_stream_probe.py [model] [B] [S] [NB] [calls]
Env: LPC_STREAM_DEPTH=1|2 (steps queued ahead), LPC_E2E_SKIP_H2D=1 (drop the copies: what is left is compute + fixed costs)."""
import importlib
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("lpc-yolo_b200")
synth = importlib.import_module("lpc-yolo_b200.utils.synth")
from bench import FILES  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "lpc"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
S = int(sys.argv[3]) if len(sys.argv) > 3 else 640
NB = int(sys.argv[4]) if len(sys.argv) > 4 else 16
calls = int(sys.argv[5]) if len(sys.argv) > 5 else 4
yolo = pkg.YOLO(FILES[name])
synth.init_synthetic(yolo.model)
g = torch.Generator().manual_seed(0)
x = torch.randint(0, 256, (NB * B, S, S, 3), generator=g, dtype=torch.uint8).pin_memory().numpy()
kw = dict(conf=0.25, half=True, imgsz=S, batch=B)
for _ in yolo.predict(x[:3 * B], stream=True, **kw):
    pass
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 0
host = 0.0
for _ in range(calls):
    it = iter(yolo.predict(x, stream=True, **kw))
    while True:
        h0 = time.perf_counter()
        try:
            next(it)
        except StopIteration:
            break
        host += time.perf_counter() - h0
        n += 1
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"{name} B={B} S={S} NB={NB} depth={os.environ.get('LPC_STREAM_DEPTH', '1')} skip_h2d={os.environ.get('LPC_E2E_SKIP_H2D', '0')}: "
      f"{n / dt:.0f} img/s, {1e3 * dt * B / n:.3f} ms/step (time inside next(): {1e3 * host * B / n:.3f} ms/step)")
