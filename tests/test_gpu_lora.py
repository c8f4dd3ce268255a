"""GPU parity: fused frozen-linear + LoRA GEMM (L1) vs the oracle (= the reference's hook, pinned by
tests/golden/reference_modules.npz) on LLaMA-3.2-3B's four LoRA-targeted shapes."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _manual(model, device="cuda"):
    """Move to the GPU WITHOUT the default switch to the fused / native paths (AUDIOLLM_B200_NATIVE=0): these tests
    turn each path on themselves and compare it with the reference-style hook path."""
    import os
    from unittest.mock import patch
    with patch.dict(os.environ, {"AUDIOLLM_B200_NATIVE": "0"}):
        return model.to(device)

from audio_llama_b200 import ops
from oracle import encoder as O


@pytest.mark.parametrize("in_dim,out_dim,rank,rows", [(3072, 3072, 64, 300), (3072, 1024, 64, 2014), (3072, 8192, 64, 257),
                                                      (8192, 3072, 64, 130), (256, 256, 8, 77)])
def test_lora_linear(in_dim, out_dim, rank, rows):
    g = torch.Generator().manual_seed(in_dim + out_dim)
    x = torch.randn(rows, in_dim, generator=g).bfloat16()
    W = (torch.randn(out_dim, in_dim, generator=g) * 0.02).bfloat16()
    bias = torch.randn(out_dim, generator=g) * 0.1
    A = torch.randn(rank, in_dim, generator=g) * 0.05          # reference init is zeros (update == 0): randomise
    B = torch.randn(out_dim, rank, generator=g) * 0.05
    scaling = 16 / rank
    ref = O.lora_linear(x.float(), W.float(), bias, A, B, scaling)
    base = O.lora_linear(x.float(), W.float(), bias, torch.zeros_like(A), B, scaling)
    y = ops.lora_linear(x.cuda(), W.cuda(), bias.cuda(), A.cuda(), B.cuda(), scaling, out_dtype=torch.float32).cpu()
    assert O.rel_l2(y, ref) <= 1e-2
    # the low-rank part itself (what the hook adds) is reproduced, not just the frozen product
    assert O.rel_l2(y - base, ref - base) <= 3e-2
    # zero A (the reference's init) -> exactly the frozen linear
    y0 = ops.lora_linear(x.cuda(), W.cuda(), bias.cuda(), torch.zeros_like(A).cuda(), B.cuda(), scaling, out_dtype=torch.float32).cpu()
    assert O.rel_l2(y0, base) <= 2e-3


def test_lora_linear_matches_reference_hook_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "reference_modules.npz"))
    t = lambda k: torch.from_numpy(g[k])
    y = ops.lora_linear(t("lora_x").cuda().bfloat16(), t("lora_W").cuda().bfloat16(), t("lora_b").cuda(), t("lora_A").cuda(),
                        t("lora_B").cuda(), float(g["lora_scaling"][0]), out_dtype=torch.float32).cpu()
    assert O.rel_l2(y, t("lora_y")) <= 2e-2


def test_fused_lora_in_llama_matches_hooks():
    """AudioLLM.enable_fused_lora(): logits and LoRA gradients of a bf16 LLaMA equal the hook path's."""
    from unittest.mock import patch
    from transformers import LlamaConfig, LlamaForCausalLM
    from audio_llama_b200.models import base as B
    from audio_llama_b200.models.allm import AudioLLM
    from audio_llama_b200.config import EncoderConfig
    from audio_llama_b200.encoder import WhisperEncoderModule
    from audio_llama_b200 import synth

    def fake(lp, wp):
        torch.manual_seed(0)
        lc = LlamaConfig(vocab_size=320, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                         num_attention_heads=4, num_key_value_heads=4)
        ec = EncoderConfig(d_model=128, n_layers=1, n_heads=2, ffn_dim=256, n_mels=80)
        return (B.FrozenModelWrapper(LlamaForCausalLM(lc).to(torch.bfloat16)),
                B.FrozenModelWrapper(WhisperEncoderModule(ec, synth.init_encoder_weights(ec), max_batch=2)))

    def run(fused):
        with patch.object(B, "load_base_models", fake):
            m = _manual(AudioLLM("x", "y", lora_rank=8))
        g = torch.Generator().manual_seed(3)
        for l in m.lora_layers.values():
            with torch.no_grad():
                l.lora_A.copy_(torch.randn(l.lora_A.shape, generator=g) * 0.05)
                l.lora_B.copy_(torch.randn(l.lora_B.shape, generator=g) * 0.05)
        if fused:
            m.enable_fused_lora()
        ids, mask, labels = (t.cuda() for t in synth.synth_text(2, 16, 320))
        out = m(input_ids=ids, attention_mask=mask, labels=labels)            # text-only: isolates the LLaMA linears
        out.loss.backward()
        l0 = m.lora_layers["model.layers.0.self_attn.q_proj"]
        return out.logits.float().cpu(), l0.lora_A.grad.float().cpu(), l0.lora_B.grad.float().cpu()

    a, b = run(False), run(True)
    assert O.rel_l2(b[0], a[0]) <= 3e-2                       # bf16 LLaMA both ways; different rounding points
    assert O.rel_l2(b[1], a[1]) <= 1e-1 and O.rel_l2(b[2], a[2]) <= 1e-1


@pytest.mark.parametrize("in_dim,out_dim,rank,rows", [(3072, 1024, 64, 2014), (1024, 3072, 64, 517), (256, 512, 8, 77)])
def test_lora_linear_backward(in_dim, out_dim, rank, rows):
    """al_lora_linear_backward (dx, dA, dB) vs torch autograd of the oracle's hook formula in fp32."""
    g = torch.Generator().manual_seed(in_dim * 7 + out_dim)
    x = torch.randn(rows, in_dim, generator=g).bfloat16()
    W = (torch.randn(out_dim, in_dim, generator=g) * 0.02).bfloat16()
    A = (torch.randn(rank, in_dim, generator=g) * 0.05)
    B = (torch.randn(out_dim, rank, generator=g) * 0.05)
    dy = torch.randn(rows, out_dim, generator=g).bfloat16()
    scaling = 16 / rank
    xr = x.float().requires_grad_(True)
    Ar = A.bfloat16().float().requires_grad_(True)      # the kernel's operands are the bf16 roundings
    Br = B.clone().requires_grad_(True)
    y = O.lora_linear(xr, W.float(), None, Ar, Br, scaling)
    y.backward(dy.float())
    xc, Wc = x.cuda(), W.cuda()
    _, (a_pad, b_pad, t) = ops.lora_linear(xc, Wc, None, A.cuda(), B.cuda(), scaling, return_saved=True)
    dx, dA, dB_raw = ops.lora_linear_backward(xc, dy.cuda(), Wc.t().contiguous(), a_pad, b_pad, t, rank)
    assert O.rel_l2(dx.float().cpu(), xr.grad) <= 1e-2
    assert O.rel_l2(dA.cpu(), Ar.grad) <= 2e-2
    assert O.rel_l2((dB_raw * scaling).cpu(), Br.grad) <= 2e-2
    # dx skipped when the input needs no gradient
    dx2, dA2, _ = ops.lora_linear_backward(xc, dy.cuda(), None, a_pad, b_pad, t, rank, need_dx=False)
    assert dx2 is None and O.rel_l2(dA2.cpu(), dA.cpu()) <= 1e-5      # (split-K reduce-add: not bit-reproducible)


@pytest.mark.parametrize("in_dim,out_dim,rank,rows", [(3072, 1024, 64, 2014), (256, 512, 8, 77)])
def test_lora_linear_epilogue_adds(in_dim, out_dim, rank, rows):
    """The _ex forms: forward with an addend == forward + addend; backward accumulating into an existing dx (in place)
    == the sum of the separate input gradients. One rounding to bf16 instead of two: compared in fp32 at bf16 tolerance."""
    g = torch.Generator().manual_seed(in_dim + out_dim)
    x = torch.randn(rows, in_dim, generator=g).bfloat16().cuda()
    W = (torch.randn(out_dim, in_dim, generator=g) * 0.02).bfloat16().cuda()
    A = (torch.randn(rank, in_dim, generator=g) * 0.05).cuda()
    B = (torch.randn(out_dim, rank, generator=g) * 0.05).cuda()
    r = torch.randn(rows, out_dim, generator=g).bfloat16().cuda()
    dy = torch.randn(rows, out_dim, generator=g).bfloat16().cuda()
    scaling = 16 / rank
    y0, (a_pad, b_pad, t) = ops.lora_linear(x, W, None, A, B, scaling, return_saved=True)
    y1 = ops.lora_linear(x, W, None, A, B, scaling, addend=r)
    ref = y0.float() + r.float()
    assert (y1.float() - ref).abs().max().item() <= 2 ** -7 * max(1.0, ref.abs().max().item())
    assert O.rel_l2(y1.float().cpu(), ref.cpu()) <= 4e-3
    wt = W.t().contiguous()
    dx_a, dA_a, _ = ops.lora_linear_backward(x, dy, wt, a_pad, b_pad, t, rank)
    dx_b, _, _ = ops.lora_linear_backward(x, dy * 0.5, wt, a_pad, b_pad, t, rank)
    acc = dx_a.clone()
    out, dA_c, _ = ops.lora_linear_backward(x, dy * 0.5, wt, a_pad, b_pad, t, rank, dx_accumulate=acc)
    assert out.data_ptr() == acc.data_ptr()                 # accumulated in place
    ref = dx_a.float() + dx_b.float()
    assert O.rel_l2(out.float().cpu(), ref.cpu()) <= 6e-3
    assert O.rel_l2(dA_c.cpu(), 0.5 * dA_a.cpu()) <= 1e-2


def test_lora_pack_kernel_is_bit_identical_to_the_torch_formula():
    """al_lora_pack (one launch) == zero-padded bf16(A), bf16(scaling * B) built with torch ops; non-fp32 parameters keep
    the torch path."""
    g = torch.Generator().manual_seed(9)
    for rank, in_dim, out_dim in ((64, 3072, 1024), (12, 256, 520)):
        A = (torch.randn(rank, in_dim, generator=g) * 0.05).cuda()
        B = (torch.randn(out_dim, rank, generator=g) * 0.05).cuda()
        scaling = 16 / rank
        a, b = ops.pack_lora(A, B, scaling)
        r_pad = (rank + 7) // 8 * 8
        a_ref = torch.zeros(r_pad, in_dim, dtype=torch.bfloat16, device="cuda")
        a_ref[:rank] = A.to(torch.bfloat16)
        b_ref = torch.zeros(out_dim, r_pad, dtype=torch.bfloat16, device="cuda")
        b_ref[:, :rank] = (B.float() * scaling).to(torch.bfloat16)
        assert torch.equal(a, a_ref) and torch.equal(b, b_ref)
        a16, b16 = ops.pack_lora(A.half(), B.half(), scaling)
        assert a16.shape == a_ref.shape and b16.shape == b_ref.shape


def test_fused_and_native_paths_are_the_default_for_bf16_cuda():
    """AudioLLM.to("cuda") with bf16 LLaMA weights switches to the fused frozen+LoRA GEMMs and the native row kernels
    by itself; fp32 weights keep the reference-style hooks. Loss of the default path == loss of the hook path."""
    from unittest.mock import patch
    from transformers import LlamaConfig, LlamaForCausalLM
    from audio_llama_b200.models import base as B
    from audio_llama_b200.models.allm import AudioLLM
    from audio_llama_b200.config import EncoderConfig
    from audio_llama_b200.encoder import WhisperEncoderModule
    from audio_llama_b200 import llama_native, synth

    def fake(dtype):
        def f(lp, wp):
            torch.manual_seed(0)
            lc = LlamaConfig(vocab_size=320, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                             num_attention_heads=4, num_key_value_heads=4)
            ec = EncoderConfig(d_model=128, n_layers=1, n_heads=2, ffn_dim=256, n_mels=80)
            return (B.FrozenModelWrapper(LlamaForCausalLM(lc).to(dtype)),
                    B.FrozenModelWrapper(WhisperEncoderModule(ec, synth.init_encoder_weights(ec), max_batch=2)))
        return f

    def loss_of(m):
        g = torch.Generator().manual_seed(3)
        for l in m.lora_layers.values():
            with torch.no_grad():
                l.lora_A.copy_(torch.randn(l.lora_A.shape, generator=g) * 0.05)
                l.lora_B.copy_(torch.randn(l.lora_B.shape, generator=g) * 0.05)
        ids, mask, labels = (t.cuda() for t in synth.synth_text(2, 16, 320))
        return float(m(input_ids=ids, attention_mask=mask, labels=labels).loss)

    try:
        with patch.object(B, "load_base_models", fake(torch.bfloat16)):
            m_default = AudioLLM("x", "y", lora_rank=8).to("cuda")
            m_hooks = _manual(AudioLLM("x", "y", lora_rank=8))
        assert getattr(m_default, "fused_lora", False) and getattr(m_default, "native_llama", False) and not m_default.hooks
        assert not getattr(m_hooks, "fused_lora", False) and len(m_hooks.hooks) == len(m_hooks.lora_layers)
        a, b = loss_of(m_default), loss_of(m_hooks)
        assert abs(a - b) <= 2e-2 * abs(b)
        with patch.object(B, "load_base_models", fake(torch.float32)):
            m32 = AudioLLM("x", "y", lora_rank=8).to("cuda")
        assert not getattr(m32, "fused_lora", False) and len(m32.hooks) == len(m32.lora_layers)
    finally:
        llama_native.disable_rope_patch()
