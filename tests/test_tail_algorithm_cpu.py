"""CPU check of the selection ALGORITHM of select_decode_kernel (lpc-yolo_b200/csrc/tail.cu), independent of the GPU:
a numpy model of the kernel's routes (stage-1 threshold digits, ties admitted in ascending anchor order, stage-2 candidates
= pairs of the selected anchors whose key reaches that threshold, optional cut that keeps every tie of the K-th key, final
ranking by (key desc, flat index asc)) must pick exactly what the two-stage definition of ops.v10postprocess
(utils/ops.py:851-864) picks under the same tie rule.  The GPU parity tests (tests/test_gpu_tail.py) compare the kernel
itself with the oracle; this file pins the argument the kernel's short cut rests on:

    every pair among the K best has key >= the K-th largest anchor maximum,

for bf16 keys (two 8-bit digits = the top 16 bits) and fp32 keys (four digits)."""
import numpy as np
import pytest
import torch

DIRECT_MAX, CAND_CAP, RANK_CAP = 512, 4096, 1024      # tail.cu


def fkey(x):
    """Order-preserving uint32 key of a float32 (tail.cu fkey())."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return np.where(u >> 31 != 0, ~u, u ^ np.uint32(0x80000000)).astype(np.uint32)


def two_stage_definition(logits, K):
    """v10postprocess with the deterministic tie rule: stage 1 = K anchors by (max key desc, anchor asc); stage 2 = K pairs
    of those anchors by (key desc, flat index asc)."""
    A, nc = logits.shape
    keys = fkey(logits)
    amax = keys.max(1)
    order1 = np.lexsort((np.arange(A), -amax.astype(np.int64)))[:K]
    sel = np.sort(order1)
    pk = keys[sel].reshape(-1)
    flat = (sel[:, None] * nc + np.arange(nc)[None, :]).reshape(-1)
    order2 = np.lexsort((flat, -pk.astype(np.int64)))[:K]
    return flat[order2], pk[order2]


def kernel_model(logits, K, passes):
    """The routes of select_decode_kernel; returns (flat indices, keys, route)."""
    A, nc = logits.shape
    shift = 32 - 8 * passes
    keys = fkey(logits)
    top1 = keys.max(1) >> shift
    thr = np.sort(top1)[-K]                               # radix select: digit string of the K-th largest anchor key
    gt, eq = np.nonzero(top1 > thr)[0], np.nonzero(top1 == thr)[0]
    need = K - len(gt)
    assert 0 < need <= len(eq)
    sel = np.sort(np.concatenate([gt, eq[:need]]))        # ordered collect: ties admitted in ascending anchor order
    pk = keys[sel].reshape(-1)
    flat = (sel[:, None] * nc + np.arange(nc)[None, :]).reshape(-1)
    cand = (pk >> shift) >= thr
    n_cand = int(cand.sum())
    assert n_cand >= K, "each selected anchor owns a pair that reaches the threshold"
    route = "direct"
    if n_cand > CAND_CAP:
        route = "overflow"                                # fall back: radix select over all K*nc keys in flat order
        ck, cf = pk, flat
    else:
        ck, cf = pk[cand], flat[cand]
    if route == "overflow" or n_cand > max(DIRECT_MAX, K):
        top2 = ck >> shift
        t2 = np.sort(top2)[-K]
        m = int((top2 >= t2).sum())
        if route != "overflow" and m <= RANK_CAP:
            route = "cut"
            keep = top2 >= t2                             # every tie of the K-th key survives; the ranking orders them
            ck, cf = ck[keep], cf[keep]
        else:
            if route != "overflow":
                route = "survivors"
                ck, cf = pk, flat
                top2 = ck >> shift
                t2 = np.sort(top2)[-K]
            g2, e2 = np.nonzero(top2 > t2)[0], np.nonzero(top2 == t2)[0]
            keep = np.sort(np.concatenate([g2, e2[: K - len(g2)]]))   # ties in ascending flat order (the arrays are flat-ordered)
            ck, cf = ck[keep], cf[keep]
    order = np.lexsort((cf, -ck.astype(np.int64)))[:K]
    return cf[order], ck[order], route


def _logits(kind, dtype, A=8400, nc=80, seed=0):
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(A, 1, generator=g) * 2.0 - 6.0
    if kind == "random":
        v = torch.randn(A, nc, generator=g) * 2.0 - 6.0
    elif kind == "equal":
        v = torch.zeros(A, nc)
    else:
        csd, q = kind
        v = base + csd * torch.randn(A, nc, generator=g)
        if q:
            v = torch.round(v * q) / q
    return v.to(dtype).float().numpy()


@pytest.mark.parametrize("dtype,passes", [(torch.bfloat16, 2), (torch.float32, 4)])
@pytest.mark.parametrize("kind,expect", [("random", "direct"), ((3.0, 0), "direct"), ((1.5, 8), "cut"), ((2.0, 0.5), "cut"),
                                         ((1.5, 0.5), "survivors"), ((0.5, 1), "overflow"), ((0.6, 0), "overflow"), ("equal", "overflow")])
def test_kernel_routes_equal_the_two_stage_definition(dtype, passes, kind, expect):
    K = 300
    x = _logits(kind, dtype)
    want_flat, want_keys = two_stage_definition(x, K)
    got_flat, got_keys, route = kernel_model(x, K, passes)
    assert route == expect
    assert np.array_equal(got_keys, want_keys)
    assert np.array_equal(got_flat, want_flat)


@pytest.mark.parametrize("A,K", [(300, 300), (2100, 100), (525, 300)])
def test_small_maps(A, K):
    x = _logits("random", torch.bfloat16, A=A, seed=A)
    want_flat, want_keys = two_stage_definition(x, K)
    got_flat, got_keys, _ = kernel_model(x, K, 2)
    assert np.array_equal(got_flat, want_flat) and np.array_equal(got_keys, want_keys)


def test_definition_matches_the_oracle_on_separated_scores():
    """The deterministic two-stage definition used above is the oracle's v10postprocess wherever scores are distinct."""
    import lpc_oracle as O
    g = torch.Generator().manual_seed(3)
    A, nc, K = 2100, 80, 300
    logits = torch.randn(A, nc, generator=g) * 2.0 - 5.0
    y = torch.cat([torch.rand(A, 4, generator=g), torch.sigmoid(logits)], 1)[None]        # [1, A, 4+nc]
    _, scores, labels, idx = O.v10postprocess(y, K, nc)
    flat, keys = two_stage_definition(torch.sigmoid(logits).numpy(), K)
    s = scores[0].numpy()
    sep = np.ones(K, dtype=bool)
    eq = s[:-1] == s[1:]
    sep[:-1] &= ~eq
    sep[1:] &= ~eq
    assert np.array_equal((flat % nc)[sep], labels[0].numpy()[sep])
    assert np.array_equal((flat // nc)[sep], idx[0].numpy()[sep])
