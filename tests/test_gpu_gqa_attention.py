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
                                               (2, 515, 3, 1, [515, 130]), (1, 2014, 6, 2, None),
                                               # more work items than SMs in all three persistent kernels (288 / 288 / 192),
                                               # with fully padded kv tiles in the middle of a CTA's item list
                                               (4, 700, 16, 8, [700, 130, 5, 384])])
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


def test_gqa_persistent_kernels_are_deterministic_across_item_schedules():
    """The persistent kernels walk their item lists in a serpentine order over the SMs; a tile's result must not depend
    on which CTA took it or on what that CTA did before: the same rows computed alone (few items) and inside a large
    batch (many items per CTA) are bit-identical, forward and backward."""
    B, S, Hq, Hkv = 6, 900, 12, 4
    q, k, v = make(B, S, Hq, Hkv, seed=5)
    d_out = torch.randn(B, S, Hq, 128, generator=torch.Generator().manual_seed(2)).bfloat16()
    scale = 128 ** -0.5

    def run(sl):
        qc, kc, vc = (t[sl].contiguous().cuda().requires_grad_(True) for t in (q, k, v))
        out = LN.gqa_attention(qc, kc, vc, None, scale)
        out.backward(d_out[sl].contiguous().cuda())
        return out.detach().cpu(), qc.grad.cpu(), kc.grad.cpu(), vc.grad.cpu()

    big = run(slice(0, B))
    one = run(slice(3, 4))
    for a, b_ in zip(big, one):
        assert torch.equal(a[3:4], b_)


def test_native_attention_inside_audio_llm_matches_hf_with_padding():
    """AudioLLM on a bf16 LLaMA with head_dim 128 and GQA: the native attention (causal + kv_len from a right-padded
    mask, labels carrying pad ids at the padded positions as the reference's dataset produces them) gives HF's loss and
    LoRA gradients; a left-padded mask keeps the stock path."""
    import os
    from unittest.mock import patch
    from transformers import LlamaConfig, LlamaForCausalLM
    from audio_llama_b200.models import base as Bm
    from audio_llama_b200.models.allm import AudioLLM
    from audio_llama_b200.config import EncoderConfig
    from audio_llama_b200.encoder import WhisperEncoderModule
    from audio_llama_b200 import synth

    def fake(lp, wp):
        torch.manual_seed(0)
        lc = LlamaConfig(vocab_size=320, hidden_size=512, intermediate_size=512, num_hidden_layers=2,
                         num_attention_heads=4, num_key_value_heads=2)          # head_dim 128, 2 query heads per kv head
        ec = EncoderConfig(d_model=128, n_layers=1, n_heads=2, ffn_dim=256, n_mels=80)
        return (Bm.FrozenModelWrapper(LlamaForCausalLM(lc).to(torch.bfloat16)),
                Bm.FrozenModelWrapper(WhisperEncoderModule(ec, synth.init_encoder_weights(ec), max_batch=2)))

    ids, mask, labels = (t.cuda() for t in synth.synth_text(2, 200, 320))
    mask[0, 150:] = 0                                        # right padding; its labels stay token ids
    mask[1, :] = 1

    def run(attention, m=mask, fused_layers=True):
        with patch.object(Bm, "load_base_models", fake), patch.dict(os.environ, {"AUDIOLLM_B200_NATIVE": "0"}):
            mdl = AudioLLM("x", "y", lora_rank=8).to("cuda")
        g = torch.Generator().manual_seed(3)
        for l in mdl.lora_layers.values():
            with torch.no_grad():
                l.lora_A.copy_(torch.randn(l.lora_A.shape, generator=g) * 0.05)
                l.lora_B.copy_(torch.randn(l.lora_B.shape, generator=g) * 0.05)
        mdl.enable_fused_lora()
        mdl.enable_native_llama_ops(attention=attention, fused_layers=fused_layers)
        assert mdl.native_attention == attention
        # the fused decoder-layer forward (adds in the GEMM / rmsnorm-backward epilogues) is armed with the native attention
        assert (getattr(mdl.llama.model.model.layers[0], "_al_lora", None) is not None) == (attention and fused_layers)
        out = mdl(input_ids=ids, attention_mask=m, labels=labels)
        out.loss.backward()
        l0 = mdl.lora_layers["model.layers.0.self_attn.q_proj"]
        l1 = mdl.lora_layers["model.layers.1.self_attn.v_proj"]
        return float(out.loss), l0.lora_B.grad.float().cpu(), l1.lora_A.grad.float().cpu()

    try:
        a, b = run(False), run(True)
        assert abs(a[0] - b[0]) <= 2e-2 * abs(a[0]), (a[0], b[0])
        assert rel(b[1], a[1]) <= 8e-2 and rel(b[2], a[2]) <= 8e-2, (rel(b[1], a[1]), rel(b[2], a[2]))
        # same kernels with autograd's own adds instead of the epilogue adds: only the rounding points differ
        u = run(True, fused_layers=False)
        assert abs(u[0] - b[0]) <= 5e-3 * abs(u[0]), (u[0], b[0])
        assert rel(b[1], u[1]) <= 3e-2 and rel(b[2], u[2]) <= 3e-2, (rel(b[1], u[1]), rel(b[2], u[2]))
        left = mask.clone()
        left[0] = 0
        left[0, 50:] = 1                                     # left padding: not a kv_len mask -> stock path, still finite
        c = run(True, left)
        assert c[0] == c[0]
    finally:
        LN.disable_rope_patch()
