"""Device-timed step (CUDA-graph replay) of one model / batch for several batch-stream counts (nn/tasks.py detect(streams=n)).
    python tools/stream_sweep.py [model] [batch] [size] [n1,n2,...]"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

pkg = importlib.import_module("lpc-yolo_b200")
Fn = importlib.import_module("lpc-yolo_b200.functional")
synth = importlib.import_module("lpc-yolo_b200.utils.synth")


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "lpc"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 640
    ns = [int(v) for v in (sys.argv[4] if len(sys.argv) > 4 else "1,2,3,4").split(",")]
    yolo = pkg.YOLO(bench.FILES[name])
    synth.init_synthetic(yolo.model)
    m = yolo.model.cuda().eval()
    m.compute_dtype = torch.bfloat16
    g = torch.Generator().manual_seed(1)
    x = Fn.pack_u8(torch.randint(0, 256, (B, S, S, 3), generator=g, dtype=torch.uint8).cuda(), torch.bfloat16)
    ref = None
    with torch.no_grad():
        for n in ns:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    m.detect(x, 300, streams=n)
            torch.cuda.current_stream().wait_stream(s)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                out = m.detect(x, 300, streams=n)
            for _ in range(3):
                gr.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(30):
                gr.replay()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 30
            same = None
            if ref is None:
                ref = out.clone()
            else:
                same = bool(torch.equal(out, ref))
            print(json.dumps({"model": name, "batch": B, "size": S, "streams": n, "ms_per_step": round(ms, 4), "img_per_s": round(B / ms * 1e3, 1),
                              "identical_to_first": same}), flush=True)
            del gr


if __name__ == "__main__":
    main()
