"""Per-launch profile of one step of the bench workload (CUDA events around every op, eager mode).
usage: python tools/profile_step.py [model] [batch] [size]   -> table sorted by time, written to stdout."""
import importlib, os, sys
os.environ.setdefault("LPC_BRANCH_STREAMS", "0")    # per-op attribution needs one stream (side streams overlap ops inside the event brackets)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("lpc-yolo_b200")
Fn = importlib.import_module("lpc-yolo_b200.functional")
synth = importlib.import_module("lpc-yolo_b200.utils.synth")
sys.path.insert(0, ROOT)
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "lpc"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
S = int(sys.argv[3]) if len(sys.argv) > 3 else 640
yolo = pkg.YOLO(bench.FILES[name]); synth.init_synthetic(yolo.model)
m = yolo.model.cuda().eval(); m.compute_dtype = torch.bfloat16
x = Fn.pack_input(torch.rand(B, 3, S, S).cuda(), torch.bfloat16)
with torch.no_grad():
    for _ in range(3):
        m.detect(x, 300)
    torch.cuda.synchronize()
    Fn.PROFILE = []
    torch.cuda.cudart().cudaProfilerStart()      # ncu --profile-from-start off captures exactly this step
    m.detect(x, 300)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
prof, Fn.PROFILE = Fn.PROFILE, None
rows = [(k, a.elapsed_time(b) * 1e3, fl, by, tag) for (k, a, b, fl, by, tag) in prof]
tot = sum(r[1] for r in rows)
print(f"# {name} B={B} S={S}: {len(rows)} ops, {tot/1e3:.3f} ms total (eager, event-bracketed)")
agg = {}
for k, us, fl, by, tag in rows:
    a = agg.setdefault(k, [0, 0.0, 0.0, 0.0]); a[0] += 1; a[1] += us; a[2] += fl; a[3] += by
print("kind,count,us,share,TFLOP/s,GB/s")
for k, (n, us, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{n},{us:.1f},{us/tot:.3f},{fl/us/1e6:.1f},{by/us/1e3:.0f}")
print("\nkind,tag,us,TFLOP/s,GB/s(algorithmic)")
for k, us, fl, by, tag in sorted(rows, key=lambda r: -r[1])[:45]:
    print(f"{k},{tag},{us:.1f},{fl/us/1e6:.1f},{by/us/1e3:.0f}")

# every op in launch order with its roofline floor (max of FLOPs / tensor peak and bytes / HBM peak); joined by
# tools/join_ncu.py with the ncu launch list of the same command
import json
pk = bench.peaks()
print("\n#ORDER idx,kind,tag,us_eager,gflop,mbytes,floor_us")
for i, (k, us, fl, by, tag) in enumerate(rows):
    floor = max(fl / (pk["tf_burst"] * 1e12), by / (pk["hbm"] * 1e9)) * 1e6
    print(f"#ORDER {i},{k},{tag},{us:.1f},{fl/1e9:.3f},{by/1e6:.2f},{floor:.1f}")
