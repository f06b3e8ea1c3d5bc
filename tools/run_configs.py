"""Run the BASELINE.json configs that are not the default bench line (device-timed, CUDA graph) and print one JSON per config:
   python tools/run_configs.py            # s/m B=256 @640, x B=32 @1280, b B=1 latency @320/640/960"""
import importlib, json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
pkg = importlib.import_module("lpc-yolo_b200")
Fn = importlib.import_module("lpc-yolo_b200.functional")
synth = importlib.import_module("lpc-yolo_b200.utils.synth")
PK = bench.peaks()


def run(name, B, S, steps=10, latency=False):
    yolo = pkg.YOLO(bench.FILES[name]); synth.init_synthetic(yolo.model)
    m = yolo.model.cuda().eval(); m.compute_dtype = torch.bfloat16
    x = Fn.pack_input(torch.rand(B, 3, S, S).cuda(), torch.bfloat16)
    with torch.no_grad():
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                m.detect(x, 300)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = m.detect(x, 300)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        if latency:
            ts = []
            for _ in range(100):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); g.replay(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ms = statistics.median(ts)
        else:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                g.replay()
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / steps
    gf = bench.GFLOP_IMG_640[name] * (S / 640) ** 2
    rec = {"model": name, "batch": B, "size": S, "ms_per_step": round(ms, 4), "img_per_s": round(B / ms * 1e3, 1),
           "conv_tflops_whole_step": round(B * gf / ms, 1), "frac_of_bf16_peak_whole_step": round(B * gf / ms / PK["tf_burst"], 4)}
    if latency:
        rec["p50_latency_ms"] = round(ms, 4)
    print(json.dumps(rec), flush=True)
    del g, m, yolo
    torch.cuda.empty_cache()


if __name__ == "__main__":
    which = sys.argv[1:] or ["s", "m", "x", "b"]
    if "s" in which: run("yolov10s", 256, 640)
    if "m" in which: run("yolov10m", 256, 640)
    if "x" in which: run("yolov10x", 32, 1280)
    if "b" in which:
        for S in (320, 640, 960):
            run("yolov10b", 1, S, latency=True)
    if "n" in which: run("yolov10n", 64, 640)
