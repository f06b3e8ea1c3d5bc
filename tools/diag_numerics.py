"""Diagnostic (not a test): numerical yardsticks on the GPU box.  Prints, per model,
  ours-fp32 vs oracle-fp64, oracle-fp32 vs oracle-fp64 (the reference's own fp32 noise),
  ours-bf16 vs oracle-fp32, oracle-bf16(CPU) vs oracle-fp32 (the reference's own bf16 drift)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch
import lpc_oracle as O
pkg = importlib.import_module("lpc-yolo_b200")

def err(a, b):
    return ((a - b).abs().max() / b.abs().max()).item(), ((a - b).norm() / b.norm()).item()

names = sys.argv[1:] or ["yolov10n", "lpc", "yolov10s", "yolov10m", "yolov10b", "yolov10l", "yolov10x"]
S = 160
for name in names:
    t0 = time.time()
    om = O.build(name)
    pm = pkg.YOLOv10DetectionModel(O.MODEL_FILES[name]); pm.load_state_dict(om.sd); pm = pm.cuda().eval()
    x = O.synth_input(2, S)
    y32, raw32 = om.forward(x)
    om64 = O.OracleModel(om.name, om.layers, om.save, om.meta, om.sd).to(torch.float64)
    y64, raw64 = om64.forward(x.double())
    om16 = O.OracleModel(om.name, om.layers, om.save, om.meta, om.sd).to(torch.bfloat16)
    try:
        raw16 = om16.features(x.bfloat16())
    except Exception as e:
        raw16 = None
    with torch.no_grad():
        pm.compute_dtype = torch.float32
        o = pm(x.cuda())["one2one"]; g32 = [r.float().cpu() for r in o[1]]; gy32 = o[0].cpu()
        d32 = pm.detect(x.cuda(), 300).cpu()
        pm.compute_dtype = torch.bfloat16
        o = pm(x.cuda())["one2one"]; g16 = [r.float().cpu() for r in o[1]]
    cat = lambda rs: torch.cat([r.reshape(r.shape[0], r.shape[1], -1) for r in rs], 2)
    R64, R32, G32, G16 = cat(raw64), cat(raw32), cat(g32), cat(g16)
    line = f"{name:9s} ours32-vs-64 {err(G32.double(), R64)[0]:.2e}  oracle32-vs-64 {err(R32.double(), R64)[0]:.2e}  "
    line += f"y: ours {err(gy32.double(), y64)[0]:.2e} oracle {err(y32.double(), y64)[0]:.2e} | ours16-vs-32 max {err(G16, R32)[0]:.3f} l2 {err(G16, R32)[1]:.4f}"
    if raw16 is not None:
        R16 = cat([r.float() for r in raw16])
        line += f"  oracle16-vs-32 max {err(R16, R32)[0]:.3f} l2 {err(R16, R32)[1]:.4f}"
    # box / score deltas of fp32 detections
    od, _, _, _ = om.predict(x)
    same = (d32[..., 5] == od[..., 5])
    db = (d32[..., :4] - od[..., :4]).abs().max(-1).values
    ds = (d32[..., 4] - od[..., 4]).abs() / od[..., 4]
    line += f" | dets: same-rank class eq {same.float().mean():.3f} box dmax(px) median {db[same].median():.1e} p99 {db[same].quantile(0.99):.1e} score rel p99 {ds[same].quantile(0.99):.1e}"
    print(line, f"[{time.time()-t0:.0f}s]", flush=True)
