"""sha1 of every layer's output (and of the head's intermediate maps) for one seeded batch: run on two devices and diff (T7 debugging)."""
import hashlib, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
pkg = importlib.import_module("lpc-yolo_b200")
Fn = importlib.import_module("lpc-yolo_b200.functional")
synth = importlib.import_module("lpc-yolo_b200.utils.synth")
name = sys.argv[1] if len(sys.argv) > 1 else "lpc"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
yolo = pkg.YOLO(bench.FILES[name]); synth.init_synthetic(yolo.model, seed=0)
m = yolo.model.cuda().eval(); m.compute_dtype = torch.bfloat16
S = 640
g = torch.Generator().manual_seed(2)
x = Fn.pack_u8(torch.randint(0, 256, (B, S, S, 3), generator=g, dtype=torch.uint8).cuda(), torch.bfloat16)


def h(t):
    return hashlib.sha1(t.float().contiguous().cpu().numpy().tobytes()).hexdigest()[:12]


def hook(mod, inp, out):
    if torch.is_tensor(out):
        print(f"layer {mod.i:2d} {type(mod).__name__:16s} {h(out)}", flush=True)


for layer in m.model:
    layer.register_forward_hook(hook)
# finer: every LpcModule inside layers 21-26 and the head
mods = importlib.import_module("lpc-yolo_b200.nn.modules.base")
for nm, sub in m.named_modules():
    if isinstance(sub, mods.LpcModule) and nm.count(".") >= 2 and any(nm.startswith(f"model.{i}.") for i in (21, 22, 24, 25, 28)):
        sub.register_forward_hook(lambda mod, inp, out, nm=nm: print(f"   {nm:44s} {h(out)}", flush=True) if torch.is_tensor(out) else None)
with torch.no_grad():
    raw = m(x)["one2one"][1]
for l, r in enumerate(raw):
    print("raw", l, h(r))
