"""Generate tests/golden/preprocess.npz from the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY; build container only).

Runs the reference's own ``LetterBox`` (data/augment.py:684-742) and ``BasePredictor.preprocess`` arithmetic
(engine/predictor.py:115-133) on small seeded uint8 BGR images: same-size (no pad), narrower (auto: minimum-rectangle pad
modulo stride) and a non-auto case (full pad to the square).  Stored: the images and the reference's float32 result.
    python oracle/gen_golden_preprocess.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
from ultralytics.data.augment import LetterBox  # noqa: E402


def ref_preprocess(ims, imgsz, stride, pt=True):
    same_shapes = len({x.shape for x in ims}) == 1                       # predictor.py:153-155
    lb = LetterBox(imgsz, auto=same_shapes and pt, stride=stride)
    im = np.stack([lb(image=x) for x in ims])                            # :123
    im = im[..., ::-1].transpose((0, 3, 1, 2))                           # :124
    im = np.ascontiguousarray(im)
    im = torch.from_numpy(im).float()
    im /= 255                                                            # :131
    return im.numpy()


rng = np.random.default_rng(7)
cases = {}
for name, (h, w, imgsz) in {"same": (64, 64, 64), "narrow": (64, 40, 64), "short": (44, 64, 64), "odd": (64, 50, 64)}.items():
    ims = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(2)]
    cases[f"{name}_img"] = np.stack(ims)
    cases[f"{name}_imgsz"] = np.int64(imgsz)
    cases[f"{name}_out"] = ref_preprocess(ims, imgsz, 32)
# non-auto (mixed shapes in the batch): each image padded to the full square; stored per image
ims = [rng.integers(0, 256, (64, 40, 3), dtype=np.uint8), rng.integers(0, 256, (48, 64, 3), dtype=np.uint8)]
cases["mixed_a"], cases["mixed_b"] = ims
cases["mixed_out"] = ref_preprocess(ims, 64, 32)
out = os.path.join(os.path.dirname(HERE), "tests", "golden", "preprocess.npz")
np.savez_compressed(out, **cases)
print("wrote", out, {k: v.shape for k, v in cases.items() if hasattr(v, "shape")})
