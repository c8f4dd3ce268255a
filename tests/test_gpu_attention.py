"""GPU parity: tcgen05 attention forward vs PyTorch fp32 softmax(QK^T)V on the same bf16 inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from audio_llama_b200 import ops


def ref_attention(qkv, H):
    B, T, d3 = qkv.shape
    d = d3 // 3
    q, k, v = qkv.float().split(d, dim=-1)
    sh = lambda t: t.view(B, T, H, 64).transpose(1, 2)
    a = torch.softmax(sh(q) @ sh(k).transpose(2, 3), dim=-1) @ sh(v)
    return a.transpose(1, 2).reshape(B, T, d)


@pytest.mark.parametrize("B,T,H,qs", [(1, 128, 1, 1.0), (1, 256, 2, 1.0), (2, 300, 3, 2.0), (1, 1500, 6, 1.0),
                                      (2, 1500, 20, 3.0), (1, 92, 1, 1.0), (1, 1000, 2, 0.2),
                                      (6, 700, 20, 1.0), (3, 100, 120, 1.0)])   # many CTAs per SM slot, many heads
def test_attention(B, T, H, qs):
    g = torch.Generator().manual_seed(B * 1000 + T + H)
    qkv = torch.randn(B, T, 3 * H * 64, generator=g)
    qkv[..., : H * 64] *= qs * 0.125 * 3        # q part: pre-scaled, with spread so softmax is not flat
    qkv = qkv.bfloat16()
    ref = ref_attention(qkv, H)
    y = ops.attention(qkv.cuda(), H).float().cpu()
    err = (y - ref).abs().max().item()
    # P and the output are rounded to bf16: 2^-8 relative each, on values of the scale of V (|v| <~ 4)
    assert err <= 3e-2 * max(1.0, ref.abs().max().item()), err
    rel = ((y - ref).norm() / ref.norm()).item()
    assert rel <= 1e-2, rel
