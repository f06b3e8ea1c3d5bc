"""GPU parity tests, tier T3: the fused tail (DFL + anchor decode + sigmoid + NMS-free top-k) fed IDENTICAL raw
head maps as the oracle / the reference.  Integer results (anchor index, class id) must be bit-exact wherever
the oracle's selected scores are strictly separated; tie groups are compared as sets; scores within 2 ulp;
boxes within 1e-3 px."""
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
STRIDES = [8.0, 16.0, 32.0]


@pytest.fixture(scope="module")
def Fn(pkg):
    return importlib.import_module("lpc-yolo_b200.functional")


def _to_dev(raw, dtype, Fn):
    out = []
    for r in raw:
        t = Fn.new_act(*r.shape, dtype, "cuda")
        t.copy_(r.cuda())
        out.append(t)
    return out


def _ulp_close(a, b, n=2):
    return (a.view(torch.int32) - b.view(torch.int32)).abs().max().item() <= n


def _check_against_oracle(oracle, dets, aidx, raw_cpu, K, nc, img_hw):
    """dets/aidx from the GPU vs the oracle run on the same (already rounded) raw maps."""
    y = oracle.decode(raw_cpu, STRIDES, nc)
    odets, oaidx = oracle.postprocess(y, K, nc, img_hw=img_hw)
    B = y.shape[0]
    dets, aidx = dets.cpu(), aidx.cpu().long()
    for b in range(B):
        gs, os_ = dets[b, :, 4], odets[b, :, 4]
        assert _ulp_close(gs.contiguous(), os_.contiguous()), "score multiset differs"
        assert (gs[:-1] >= gs[1:]).all(), "scores must be sorted descending"
        # strictly separated ranks: exact anchor index and class id
        # "strictly separated": torch's CPU sigmoid is not a pure function of its input (the vectorised and
        # the scalar-tail code paths differ by 1 ulp for the SAME logit), so ties are judged with a 4-ulp margin
        sep = torch.ones(K, dtype=torch.bool)
        eq = (os_[:-1] - os_[1:]).abs() <= 5e-7 * os_[:-1].abs()
        sep[:-1] &= ~eq
        sep[1:] &= ~eq
        sep[-1] = False
        assert torch.equal(aidx[b, sep], oaidx[b, sep])
        assert torch.equal(dets[b, sep, 5], odets[b, sep, 5])
        assert not sep.any() or (dets[b, sep, :4] - odets[b, sep, :4]).abs().max() < 1e-3
        # every selected pair must carry the score it claims, and our tie order is ascending flat index
        flat = aidx[b] * nc + dets[b, :, 5].long()
        got_scores = y[b, 4:, :].t().reshape(-1)[flat]
        assert _ulp_close(got_scores.contiguous(), gs.contiguous())
        lg = torch.cat([r.reshape(r.shape[0], r.shape[1], -1) for r in raw_cpu], 2)[b, 64:, :].t().reshape(-1)[flat]
        tie = lg[:-1] == lg[1:]                      # equal LOGITS are ordered by ascending flat index
        assert (flat[:-1][tie] < flat[1:][tie]).all()
        assert (lg[:-1] >= lg[1:]).all()
        assert flat.unique().numel() == K


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name", ["yolov10n", "lpc"])
def test_tail_on_reference_raw_maps(oracle, Fn, dtype, name):
    """Golden raw maps produced by the unmodified reference (tests/golden, 160x160)."""
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    raw = [torch.from_numpy(g[f"raw_small_{l}"])[None] for l in range(3)]
    if dtype == torch.bfloat16:
        raw = [r.bfloat16().float() for r in raw]
    dev = _to_dev(raw, dtype, Fn)
    dets, aidx = Fn.v10_decode_topk(dev, STRIDES, 80, 300, (160, 160), return_index=True)
    _check_against_oracle(oracle, dets, aidx, raw, 300, 80, (160, 160))
    if dtype == torch.float32:
        # against the reference's own outputs: y and the final [300,6]
        y = Fn.v10_decode(dev, STRIDES, 80).cpu()
        ref_y = torch.from_numpy(g["y_small"])
        assert (y[0, 4:] - ref_y[4:]).abs().max() < 1e-6
        assert (y[0, :4] - ref_y[:4]).abs().max() < 1e-3
        rd = torch.from_numpy(g["dets_small"])
        assert _ulp_close(dets[0, :, 4].cpu().contiguous(), rd[:, 4].contiguous())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,S,K", [(3, 640, 300), (2, 320, 300), (1, 160, 300), (2, 160, 100), (1, 1280, 300)])
def test_tail_random_maps(oracle, Fn, dtype, B, S, K):
    """Seeded synthetic head maps at the BASELINE sizes (A = 8400 @640, 33600 @1280), incl. heavy ties in bf16."""
    g = torch.Generator().manual_seed(S + B)
    raw = []
    for l in range(3):
        h = S // 8 >> l
        r = torch.randn(B, 144, h, h, generator=g) * 2.0
        r[:, 64:] -= 6.0
        raw.append(r.to(dtype).float())
    dev = _to_dev(raw, dtype, Fn)
    dets, aidx = Fn.v10_decode_topk(dev, STRIDES, 80, K, (S, S), return_index=True)
    _check_against_oracle(oracle, dets, aidx, raw, K, 80, (S, S))


def _stage2_counts(raw, K):
    """(candidates, survivors) per image as select_decode_kernel counts them: pairs of the K selected anchors whose
    logit reaches the K-th anchor maximum, and pairs that reach the K-th pair logit (all its ties included)."""
    lg = torch.cat([r.reshape(r.shape[0], r.shape[1], -1) for r in raw], 2)[:, 64:, :]
    out = []
    for b in range(lg.shape[0]):
        m = lg[b].amax(0)
        t = m.topk(K).values[-1]
        gt, eq = (m > t).nonzero().flatten(), (m == t).nonzero().flatten()
        pairs = lg[b][:, torch.cat([gt, eq[: K - len(gt)]])].t().reshape(-1)
        out.append((int((pairs >= t).sum()), int((pairs >= pairs.topk(K).values[-1]).sum())))
    return out


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("path,csd,q", [("direct", 3.0, 0), ("cut", 1.5, 8), ("cut", 2.0, 0.5), ("survivors", 1.5, 0.5),
                                        ("overflow", 1.0, 1)])
def test_tail_stage2_paths(oracle, Fn, dtype, path, csd, q):
    """Every stage-2 route of select_decode_kernel: candidates ranked directly (<= 512), cut by a radix select over the
    candidates first (<= 4096, ties of the K-th key kept and ordered by the ranking), and the two fall-backs to the
    flat-order radix select (more than 1024 survivors / more than 4096 candidates).  Class logits = per-anchor level +
    class noise, optionally quantised so that the K-th key sits inside a large tie group."""
    K, S, B = 300, 640, 2
    g = torch.Generator().manual_seed(0)
    raw = []
    for l in range(3):
        h = S // 8 >> l
        r = torch.randn(B, 144, h, h, generator=g) * 2.0
        v = torch.randn(B, 1, h, h, generator=g) * 2.0 - 6.0 + csd * torch.randn(B, 80, h, h, generator=g)
        r[:, 64:] = torch.round(v * q) / q if q else v
        raw.append(r.to(dtype).float())
    cnt = _stage2_counts(raw, K)
    route = {"direct": lambda c, m: c <= 512, "cut": lambda c, m: 512 < c <= 4096 and m <= 1024,
             "survivors": lambda c, m: 512 < c <= 4096 and m > 1024, "overflow": lambda c, m: c > 4096}[path]
    assert any(route(c, m) for c, m in cnt), f"construction does not reach the {path} route: {cnt}"
    dev = _to_dev(raw, dtype, Fn)
    dets, aidx = Fn.v10_decode_topk(dev, STRIDES, 80, K, (S, S), return_index=True)
    _check_against_oracle(oracle, dets, aidx, raw, K, 80, (S, S))


def test_tail_all_equal_scores(oracle, Fn):
    """Degenerate input (SURVEY.md finding 5): every class logit identical -> pure tie-break by flat index."""
    raw = [torch.zeros(1, 144, 20 >> l, 20 >> l) for l in range(3)]
    dev = _to_dev(raw, torch.bfloat16, Fn)
    dets, aidx = Fn.v10_decode_topk(dev, STRIDES, 80, 300, None, return_index=True)
    assert (dets[0, :, 4] == 0.5).all()
    # two-stage semantics: stage 1 keeps anchors 0..299 (ties by index), stage 2 keeps the first 300 pairs
    flat = aidx[0].cpu().long() * 80 + dets[0, :, 5].cpu().long()
    assert torch.equal(flat, torch.arange(300))


def test_v10postprocess_api(oracle, Fn, pkg):
    """utils.ops.v10postprocess drop-in on a decoded preds tensor (both contiguous and transposed views)."""
    ops = importlib.import_module("lpc-yolo_b200.utils.ops")
    g = torch.Generator().manual_seed(7)
    y = torch.rand(2, 84, 2100, generator=g)
    y[:, 4:] = torch.sigmoid(torch.randn(2, 80, 2100, generator=g) * 2 - 5)
    ob, os_, ol, _ = oracle.v10postprocess(y.transpose(-1, -2), 300, 80)
    for preds in (y.cuda().transpose(-1, -2), y.cuda().transpose(-1, -2).contiguous()):
        b, s, l = ops.v10postprocess(preds, 300, 80)
        assert torch.equal(s.cpu(), os_)
        sep = torch.ones(2, 300, dtype=torch.bool)
        eq = os_[:, :-1] == os_[:, 1:]
        sep[:, :-1] &= ~eq
        sep[:, 1:] &= ~eq
        sep[:, -1] = False
        assert torch.equal(l.cpu()[sep], ol[sep]) and torch.equal(b.cpu()[sep], ob[sep])
        assert l.dtype == torch.int64


def test_conv_rowmax_keys_bit_exact(pkg, Fn):
    """lpc_conv2d_tc_rowmax (stage 1 of v10postprocess fused into the class-branch conv): keys == key(max_c(output))."""
    head = importlib.import_module("lpc-yolo_b200.nn.modules.head")
    torch.manual_seed(5)
    B, C1, NC, H, W = 3, 80, 80, 20, 12
    conv = head._Plain1x1(C1, NC).cuda().eval()
    with torch.no_grad():
        conv.weight.normal_(0, 0.2)
        conv.bias.normal_(0, 1.0)
    x = Fn.new_act(B, C1, H, W, torch.bfloat16, "cuda")
    x.normal_()
    A, off = 2 * H * W + 7, H * W + 7                     # keys land at an offset inside a larger per-image key array
    ws = torch.zeros((B * A * 4 + 256,), dtype=torch.uint8, device="cuda")
    out = Fn.new_act(B, 64 + NC, H, W, torch.bfloat16, "cuda")
    rm = {"ws": ws, "A": A, "off": off}
    with torch.no_grad():
        conv(x, out=out[:, 64:], rowmax=rm)
    assert rm.get("ok") is True
    torch.cuda.synchronize()
    keys = ws[: B * A * 4].view(torch.int32).view(B, A)[:, off: off + H * W].cpu()
    m = out[:, 64:].float().amax(1).reshape(B, H * W).cpu()                      # max over classes of the stored bf16 logits
    u = m.view(torch.int32)
    want = torch.where(u < 0, ~u, u ^ torch.tensor(-2 ** 31, dtype=torch.int32))    # fkey(): order-preserving uint32 of the float
    assert torch.equal(keys, want)
    assert int(ws[: B * A * 4].view(torch.int32).view(B, A)[:, :off].abs().sum()) == 0   # nothing written outside the slot


def test_tal_and_ops_function_api(pkg, oracle):
    """utils/tal.py make_anchors / dist2bbox and utils/ops.py xywh2xyxy / clip_boxes / scale_boxes (kernel-backed function
    API) against the oracle's restatements of tal.py:294-319 and ops.py:89-124, 305-324, 402-421."""
    tal = importlib.import_module("lpc-yolo_b200.utils.tal")
    ops = importlib.import_module("lpc-yolo_b200.utils.ops")
    feats = [torch.empty(2, 8, h, w, device="cuda") for h, w in ((12, 20), (6, 10), (3, 5))]
    pts, st = tal.make_anchors(feats, [8.0, 16.0, 32.0], 0.5)
    opts, ost = oracle.make_anchors([(12, 20), (6, 10), (3, 5)], [8.0, 16.0, 32.0])
    assert torch.equal(pts.cpu(), opts) and torch.equal(st.cpu(), ost)
    g = torch.Generator().manual_seed(4)
    A = pts.shape[0]
    dist = torch.rand(2, 4, A, generator=g) * 6
    for xywh in (True, False):
        got = tal.dist2bbox(dist.cuda(), pts.t().unsqueeze(0), xywh=xywh, dim=1).cpu()           # the head's call (head.py:97-101)
        lt, rb = dist.split([2, 2], 1)
        x1y1, x2y2 = opts.t().unsqueeze(0) - lt, opts.t().unsqueeze(0) + rb
        want = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) if xywh else torch.cat((x1y1, x2y2), 1)
        assert torch.equal(got, want)
    b = torch.rand(3, 50, 4, generator=g) * 300
    assert torch.equal(ops.xywh2xyxy(b.cuda()).cpu(), oracle.xywh2xyxy(b))
    want = oracle.scale_boxes((192, 256), b.clone(), (150, 200))
    assert torch.allclose(ops.scale_boxes((192, 256), b.clone().cuda(), (150, 200)).cpu(), want, atol=1e-4)
    rows = torch.cat((b, torch.rand(3, 50, 2, generator=g)), -1)                                   # [.., 6] rows: only the box moves
    got = rows.clone().cuda()
    ops.clip_boxes(got, (100, 120))
    assert torch.equal(got[..., 4:].cpu(), rows[..., 4:])
    assert torch.equal(got[..., 0].cpu(), rows[..., 0].clamp(0, 120)) and torch.equal(got[..., 3].cpu(), rows[..., 3].clamp(0, 100))
    with pytest.raises(pkg.LpcError):
        tal.dist2bbox(dist, pts.t().unsqueeze(0).cpu(), dim=1)                                     # CPU tensors: no fallback
