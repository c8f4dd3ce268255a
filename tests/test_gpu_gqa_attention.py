"""GPU parity: causal grouped-query attention (head_dim 128) of the LLaMA layers, forward and backward, against the
float32 restatement of what HF LlamaAttention computes (oracle/llama.py::gqa_attention and torch autograd of it)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from audio_llama_b200 import llama_native as LN
from oracle import llama as OL


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


def make(B, S, Hq, Hkv, seed, spread=1.0):
    g = torch.Generator().manual_seed(seed)
    q = (torch.randn(B, S, Hq, 128, generator=g) * spread).bfloat16()
    k = (torch.randn(B, S, Hkv, 128, generator=g) * spread).bfloat16()
    v = torch.randn(B, S, Hkv, 128, generator=g).bfloat16()
    return q, k, v


@pytest.mark.parametrize("B,S,Hq,Hkv,kv_len", [(1, 128, 1, 1, None), (2, 300, 4, 2, None), (1, 2014, 6, 2, None),
                                               (3, 515, 3, 1, [515, 130, 1]), (2, 2014, 24, 8, [2014, 1700]),
                                               (1, 77, 2, 2, [50])])
def test_gqa_forward(B, S, Hq, Hkv, kv_len):
    q, k, v = make(B, S, Hq, Hkv, seed=S + Hq, spread=1.5)
    scale = 128 ** -0.5
    ref = OL.gqa_attention(q, k, v, scale, kv_len)
    kl = torch.tensor(kv_len, dtype=torch.int32).cuda() if kv_len is not None else None
    out, lse = LN.gqa_attention_forward(q.cuda(), k.cuda(), v.cuda(), kl, scale)
    assert out.shape == ref.shape and lse.shape == (B, Hq, S)
    assert torch.isfinite(out).all() and torch.isfinite(lse).all()
    assert rel(out.cpu(), ref) <= 1e-2
    # log-sum-exp (base 2) of the scaled, masked scores
    g = Hq // Hkv
    s = (q.float().permute(0, 2, 1, 3) @ k.float().permute(0, 2, 1, 3).repeat_interleave(g, 1).transpose(2, 3)) * scale
    m = torch.ones(S, S, dtype=torch.bool).tril()[None, None].expand(B, 1, S, S)
    if kv_len is not None:
        m = m & (torch.arange(S)[None, :] < torch.tensor(kv_len)[:, None])[:, None, None, :]
    ref_lse = torch.logsumexp(s.masked_fill(~m, float("-inf")), dim=-1) * 1.4426950408889634
    assert (lse.cpu() - ref_lse).abs().max() <= 2e-2


@pytest.mark.parametrize("B,S,Hq,Hkv,kv_len", [(1, 128, 1, 1, None), (2, 300, 4, 2, None), (1, 1000, 6, 2, [700]),
                                               (2, 515, 3, 1, [515, 130]), (1, 2014, 6, 2, None)])
def test_gqa_backward(B, S, Hq, Hkv, kv_len):
    q, k, v = make(B, S, Hq, Hkv, seed=7 * S + Hq, spread=1.2)
    scale = 128 ** -0.5
    g = torch.Generator().manual_seed(1)
    d_out = torch.randn(B, S, Hq, 128, generator=g).bfloat16()
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    ref = OL.gqa_attention(qr, kr, vr, scale, kv_len)
    ref.backward(d_out.float())
    kl = torch.tensor(kv_len, dtype=torch.int32).cuda() if kv_len is not None else None
    qc, kc, vc = (t.cuda().requires_grad_(True) for t in (q, k, v))
    out = LN.gqa_attention(qc, kc, vc, kl, scale)
    out.backward(d_out.cuda())
    assert rel(out.detach().cpu(), ref.detach()) <= 1e-2
    for name, got, want in (("dq", qc.grad, qr.grad), ("dk", kc.grad, kr.grad), ("dv", vc.grad, vr.grad)):
        assert torch.isfinite(got).all(), name
        assert rel(got.cpu(), want) <= 3e-2, (name, rel(got.cpu(), want))
    if kv_len is not None:                                  # padded keys receive exactly zero gradient
        for b, n in enumerate(kv_len):
            assert float(kc.grad[b, n:].abs().max() if n < S else 0.0) == 0.0
            assert float(vc.grad[b, n:].abs().max() if n < S else 0.0) == 0.0
