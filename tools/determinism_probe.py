"""Is model.detect bit-reproducible run to run (eager, fresh allocations) with the bench's synthetic weights?  (T7 debugging)"""
import importlib, os, sys
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
outs = []
with torch.no_grad():
    for i in range(4):
        junk = torch.full((int(3e8),), float(i), device="cuda")      # perturb the allocator / leave different garbage behind
        del junk
        raw = [r.clone() for r in m(x)["one2one"][1]]
        d = m.detect(x, 300).clone()
        outs.append((raw, d))
    for i in range(1, 4):
        for l in range(3):
            print(f"run {i} raw level {l} identical to run 0:", torch.equal(outs[i][0][l], outs[0][0][l]))
        dd = (outs[i][1] - outs[0][1]).abs()
        print(f"run {i} detections identical: {bool((dd == 0).all())}; differing values {int((dd > 0).sum())}, images {int((dd.flatten(1).amax(1) > 0).sum())}")
    sc = outs[0][1][..., 4]
    print("distinct scores per image (first 4):", [int(torch.unique(sc[b]).numel()) for b in range(4)], "score range", sc.min().item(), sc.max().item())
    # same raw maps through the standalone tail twice
    t0 = Fn.v10_decode_topk(outs[0][0], [8.0, 16.0, 32.0], 80, 300, (640, 640)).clone()
    t1 = Fn.v10_decode_topk(outs[0][0], [8.0, 16.0, 32.0], 80, 300, (640, 640)).clone()
    print("standalone tail on identical raw maps identical:", torch.equal(t0, t1))
    # graph replay vs eager, and checksums to compare across devices
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        m.detect(x, 300)
    torch.cuda.current_stream().wait_stream(s)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        og = m.detect(x, 300)
    gr.replay(); torch.cuda.synchronize()
    print("graph replay identical to eager:", torch.equal(og, outs[0][1]))
    import hashlib
    for l in range(3):
        print(f"raw level {l} sha1", hashlib.sha1(outs[0][0][l].float().cpu().numpy().tobytes()).hexdigest()[:16])
    print("dets sha1", hashlib.sha1(outs[0][1].cpu().numpy().tobytes()).hexdigest()[:16], "device", torch.cuda.get_device_name(), os.environ.get("CUDA_VISIBLE_DEVICES"))
