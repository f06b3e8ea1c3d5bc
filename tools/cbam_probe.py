"""T7 debugging: which piece of CBAM differs between two devices for the same input."""
import hashlib, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
pkg = importlib.import_module("lpc-yolo_b200")
Fn = importlib.import_module("lpc-yolo_b200.functional")
synth = importlib.import_module("lpc-yolo_b200.utils.synth")
yolo = pkg.YOLO(bench.FILES["lpc"]); synth.init_synthetic(yolo.model, seed=0)
m = yolo.model.cuda().eval(); m.compute_dtype = torch.bfloat16
B, S = 64, 640
g = torch.Generator().manual_seed(2)
x = Fn.pack_u8(torch.randint(0, 256, (B, S, S, 3), generator=g, dtype=torch.uint8).cuda(), torch.bfloat16)
h = lambda t: hashlib.sha1(t.float().contiguous().cpu().numpy().tobytes()).hexdigest()[:12]
grab = {}
m.model[20].register_forward_hook(lambda mod, i, o: grab.__setitem__("x20", o))
with torch.no_grad():
    m(x)
    x20 = grab["x20"]
    print("x20", h(x20), "nan", int(torch.isnan(x20.float()).sum()), "inf", int(torch.isinf(x20.float()).sum()), "absmax", x20.float().abs().max().item())
    cb = m.model[21]
    ca_mod = cb.channel_attention
    w, b = ca_mod._packed(x20, ca_mod._build)
    part, scale = Fn.global_avgpool(x20)
    print("partial sums", h(part), tuple(part.shape), "nan", int(torch.isnan(part).sum()), "absmax", part.abs().max().item())
    ca = Fn.channel_mlp(part, w, b, 3, None, None, 0, scale)
    print("gate", h(ca), "nan", int(torch.isnan(ca).sum()))
    sa = cb.spatial_attention
    w7 = sa._packed(x20, sa._build)
    stats = torch.empty((B, 80 * 80, 2), dtype=torch.float32, device="cuda")
    L = pkg.lib()
    import ctypes as C
    xp, xld = Fn.view_of(x20)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.lpc_cbam_stats(0, xp, xld, B, 6400, 64, C.c_void_p(ca.data_ptr()), C.c_void_p(stats.data_ptr()), st)
    print("stats", h(stats), "nan", int(torch.isnan(stats).sum()), "absmax", stats.abs().max().item())
    y = cb(x20)
    print("cbam out", h(y), "nan", int(torch.isnan(y.float()).sum()))
    ref = x20.float().permute(0, 2, 3, 1).reshape(B, 6400, 64)
    print("torch mean hash (fp32)", h(ref.mean(1)), "gate from torch", h(torch.sigmoid(ref.mean(1) @ w.t() + b)))
