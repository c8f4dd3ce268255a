"""GPU parity: tcgen05 attention forward vs PyTorch fp32 softmax(QK^T)V on the same bf16 inputs.

The kernel keeps no row maximum: two threads share a query row and use a LAGGED exponent reference (the row's first
score, raised from the probability sums of the tile two back), with an exact per-row fallback for rows whose scores
outgrow it by 2^100. The cases below exercise all three regimes: the plain one, reference moves (scores growing along
kv), and the exact fallback (a first score ~100 nats under the row maximum)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from audio_llama_b200 import ops

LOG2E = 1.4426950408889634


def ref_attention(qkv, H, q_log2=False):
    B, T, d3 = qkv.shape
    d = d3 // 3
    q, k, v = qkv.float().split(d, dim=-1)
    sh = lambda t: t.view(B, T, H, 64).transpose(1, 2)
    s = sh(q) @ sh(k).transpose(2, 3)
    if q_log2:
        s = s * math.log(2.0)          # the scores are in log2 units
    a = torch.softmax(s, dim=-1) @ sh(v)
    return a.transpose(1, 2).reshape(B, T, d)


def check(y, ref, tol_rel=1e-2):
    err = (y - ref).abs().max().item()
    # P and the output are rounded to bf16: 2^-8 relative each, on values of the scale of V (|v| <~ 4)
    assert err <= 3e-2 * max(1.0, ref.abs().max().item()), err
    rel = ((y - ref).norm() / ref.norm()).item()
    assert rel <= tol_rel, rel


@pytest.mark.parametrize("q_log2", [False, True])
@pytest.mark.parametrize("B,T,H,qs", [(1, 128, 1, 1.0), (1, 256, 2, 1.0), (2, 300, 3, 2.0), (1, 1500, 6, 1.0),
                                      (2, 1500, 20, 3.0), (1, 92, 1, 1.0), (1, 1000, 2, 0.2),
                                      (6, 700, 20, 1.0), (3, 100, 120, 1.0),    # many CTAs per SM slot, many heads
                                      (1, 1, 1, 1.0), (1, 65, 2, 1.0), (1, 129, 1, 4.0)])
def test_attention(B, T, H, qs, q_log2):
    g = torch.Generator().manual_seed(B * 1000 + T + H)
    qkv = torch.randn(B, T, 3 * H * 64, generator=g)
    qkv[..., : H * 64] *= qs * 0.125 * 3 * (LOG2E if q_log2 else 1.0)   # q part: pre-scaled, spread so softmax is not flat
    qkv = qkv.bfloat16()
    ref = ref_attention(qkv, H, q_log2)
    y = ops.attention(qkv.cuda(), H, q_log2).float().cpu()
    check(y, ref)


def _structured(T, H, row_scores, q_log2, seed=0):
    """qkv whose head-0 scores are row_scores[kv] for every query (q = e_0 * 8, k[kv] = e_0 * row_scores[kv] / 8),
    random v; the other heads are plain random."""
    g = torch.Generator().manual_seed(seed)
    d = H * 64
    qkv = torch.randn(1, T, 3 * d, generator=g) * 0.3
    qkv[0, :, :64] = 0
    qkv[0, :, 0] = 8.0
    qkv[0, :, d:d + 64] = 0
    qkv[0, :, d] = torch.as_tensor(row_scores, dtype=torch.float32) / 8.0
    return qkv.bfloat16()


@pytest.mark.parametrize("q_log2", [False, True])
def test_attention_reference_moves(q_log2):
    """Scores that climb by ~45 (log2 units) per few tiles: the lagged reference has to follow (rescale of O and l),
    in both halves of every row consistently."""
    T, H = 1500, 2
    ramp = torch.linspace(-60.0, 60.0, T)                  # +10 per 128-wide tile
    qkv = _structured(T, H, ramp, q_log2)
    ref = ref_attention(qkv, H, q_log2)
    y = ops.attention(qkv.cuda(), H, q_log2).float().cpu()
    check(y, ref)


@pytest.mark.parametrize("q_log2", [False, True])
@pytest.mark.parametrize("where", [3, 700, 1499])
def test_attention_exact_fallback(q_log2, where):
    """One kv position scores ~110 above everything else, including the row's first score: the fast path's sums
    overflow its 2^100 guard and the CTA recomputes its rows exactly. The result is (almost) one-hot on `where`."""
    T, H = 1500, 2
    sc = torch.full((T,), -55.0)
    sc[where] = 55.0
    qkv = _structured(T, H, sc, q_log2, seed=where)
    ref = ref_attention(qkv, H, q_log2)
    y = ops.attention(qkv.cuda(), H, q_log2).float().cpu()
    assert torch.isfinite(y).all()
    check(y, ref)


def test_attention_large_uniform_offset():
    """All scores ~ +300 (a large common offset, small spread): only differences matter; the first-score reference
    absorbs the offset (general form; the raw form of q_log2 declines rows like this by itself)."""
    T, H = 700, 2
    g = torch.Generator().manual_seed(5)
    sc = 300.0 + torch.randn(T, generator=g) * 2
    for q_log2 in (False, True):
        qkv = _structured(T, H, sc, q_log2)
        ref = ref_attention(qkv, H, q_log2)
        y = ops.attention(qkv.cuda(), H, q_log2).float().cpu()
        assert torch.isfinite(y).all()
        check(y, ref, tol_rel=2e-2)    # scores of 300 carry bf16 steps of 2: the reference sees the same rounded inputs
