"""GPU parity of the LLaMA-side native ops (first slice) against the HF modules they replace
(transformers LlamaRMSNorm / LlamaMLP's act*up / apply_rotary_pos_emb / ForCausalLMLoss), forward and backward.
bf16 kernels vs an fp32 evaluation of the same formulas: rel-L2 <= 1e-2 (measured ~3e-3)."""
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

from audio_llama_b200 import llama_native as LN
from oracle import encoder as O
from oracle import llama as OL


def rel(a, b):
    return O.rel_l2(a.float().cpu(), b.float().cpu())


@pytest.mark.parametrize("rows,d", [(517, 3072), (64, 2048), (33, 256), (5, 8192)])
def test_rmsnorm(rows, d):
    from transformers.models.llama.modeling_llama import LlamaRMSNorm
    g = torch.Generator().manual_seed(d)
    x = (torch.randn(rows, d, generator=g) * 2).bfloat16().cuda()
    w = (1 + 0.1 * torch.randn(d, generator=g)).bfloat16().cuda()
    dy = torch.randn(rows, d, generator=g).bfloat16().cuda()
    m = LlamaRMSNorm(d, eps=1e-5).cuda().bfloat16()
    with torch.no_grad():
        m.weight.copy_(w)
    y, yh = LN.rmsnorm(x, w, 1e-5), m(x)
    # same rounding points as HF; only the summation order of mean(x^2) differs, which moves a few values by one bf16 ulp
    assert (y != yh).float().mean() <= 0.02 and rel(y, yh) <= 2e-3
    xr = x.float().requires_grad_(True)
    yr = OL.rmsnorm(xr, w.float(), 1e-5)
    yr.backward(dy.float())
    assert rel(y, yr) <= 5e-3
    xn = x.clone().requires_grad_(True)
    LN.rmsnorm(xn.view(1, rows, d), w, 1e-5).backward(dy.view(1, rows, d))
    assert rel(xn.grad, xr.grad) <= 1e-2


def test_residual_rmsnorm_and_linear_add():
    """The fused decoder layer's pieces: residual_rmsnorm returns (x, rmsnorm(x)) and its backward adds the gradient of
    the pass-through output inside the rmsnorm backward kernel; frozen_linear_add folds the residual add into the GEMM."""
    g = torch.Generator().manual_seed(11)
    rows, d, out_dim = 517, 3072, 1024
    x = torch.randn(rows, d, generator=g).bfloat16().cuda().requires_grad_(True)
    w = (1 + 0.1 * torch.randn(d, generator=g)).bfloat16().cuda()
    gx = torch.randn(rows, d, generator=g).bfloat16().cuda()
    gy = torch.randn(rows, d, generator=g).bfloat16().cuda()
    xa, y = LN.residual_rmsnorm(x, w, 1e-5)
    assert torch.equal(xa, x.detach()) and torch.equal(y, LN.rmsnorm(x.detach(), w, 1e-5))
    torch.autograd.backward([xa, y], [gx, gy])
    x2 = x.detach().clone().requires_grad_(True)
    LN.rmsnorm(x2, w, 1e-5).backward(gy)
    ref = x2.grad.float() + gx.float()
    assert (x.grad.float() - ref).abs().max().item() <= 2 ** -7 * max(1.0, ref.abs().max().item())
    # pass-through only / norm only
    x3 = x.detach().clone().requires_grad_(True)
    xa3, _ = LN.residual_rmsnorm(x3, w, 1e-5)
    xa3.backward(gx)
    assert torch.equal(x3.grad, gx)
    # o_proj + residual
    W = (torch.randn(out_dim, d, generator=g) * 0.02).bfloat16().cuda()
    r = torch.randn(rows, out_dim, generator=g).bfloat16().cuda().requires_grad_(True)
    a = torch.randn(rows, d, generator=g).bfloat16().cuda().requires_grad_(True)
    dy = torch.randn(rows, out_dim, generator=g).bfloat16().cuda()
    out = LN.frozen_linear_add(a, W, None, r)
    ref = (a.detach().float() @ W.float().t()) + r.detach().float()
    assert O.rel_l2(out.float().cpu(), ref.cpu()) <= 4e-3
    out.backward(dy)
    assert torch.equal(r.grad, dy)
    assert O.rel_l2(a.grad.float().cpu(), (dy.float() @ W.float()).cpu()) <= 6e-3


def test_swiglu():
    g = torch.Generator().manual_seed(1)
    a = (torch.randn(300, 8192, generator=g) * 2).bfloat16().cuda()
    b = torch.randn(300, 8192, generator=g).bfloat16().cuda()
    dh = torch.randn(300, 8192, generator=g).bfloat16().cuda()
    ar, br = a.float().requires_grad_(True), b.float().requires_grad_(True)
    hr = OL.swiglu(ar, br)
    hr.backward(dh.float())
    an, bn = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    h = LN.swiglu(an, bn)
    h.backward(dh)
    assert rel(h, hr) <= 5e-3 and rel(an.grad, ar.grad) <= 1e-2 and rel(bn.grad, br.grad) <= 1e-2


@pytest.mark.parametrize("B,S,Hq,Hkv,hd,cb", [(2, 300, 24, 8, 128, 1), (3, 65, 32, 8, 64, 3)])
def test_rope(B, S, Hq, Hkv, hd, cb):
    from transformers.models.llama import modeling_llama as ML
    orig = LN._ORIG.get("rope", ML.apply_rotary_pos_emb)
    g = torch.Generator().manual_seed(S)
    q = torch.randn(B, S, Hq, hd, generator=g).bfloat16().cuda().transpose(1, 2)        # [B, H, S, hd] view, as in HF
    k = torch.randn(B, S, Hkv, hd, generator=g).bfloat16().cuda().transpose(1, 2)
    ang = torch.rand(cb, S, hd // 2, generator=g) * 6.28
    cos = torch.cat([ang.cos(), ang.cos()], -1).bfloat16().cuda()
    sin = torch.cat([ang.sin(), ang.sin()], -1).bfloat16().cuda()
    dq = torch.randn(B, Hq, S, hd, generator=g).bfloat16().cuda()
    qr, kr = q.float().requires_grad_(True), k.float().requires_grad_(True)
    qe, ke = OL.rope(qr, kr, cos.float(), sin.float())
    qh, kh = orig(q.float(), k.float(), cos.float(), sin.float())               # the HF function itself
    assert torch.equal(qe.detach(), qh) and torch.equal(ke.detach(), kh)
    (qe * dq.float()).sum().backward()
    qn, kn = q.clone().requires_grad_(True), k.clone().requires_grad_(True)
    qo, ko = LN.apply_rotary_pos_emb(qn, kn, cos, sin)
    assert qo.shape == qe.shape and ko.shape == ke.shape
    (qo.float() * dq.float()).sum().backward()
    assert rel(qo, qe) <= 5e-3 and rel(ko, ke) <= 5e-3
    assert rel(qn.grad, qr.grad) <= 1e-2


@pytest.mark.parametrize("rows,d,V,chunk", [(300, 256, 5003, 128), (64, 512, 128258, 64), (7, 64, 40, 16)])
def test_linear_cross_entropy(rows, d, V, chunk):
    g = torch.Generator().manual_seed(V)
    h = torch.randn(rows, d, generator=g).bfloat16().cuda()
    W = (torch.randn(V, d, generator=g) * 0.05).bfloat16().cuda()
    labels = torch.randint(0, V, (rows,), generator=g)
    labels[::5] = -100
    labels = labels.cuda()
    hr = h.float().requires_grad_(True)
    logits = (hr @ W.float().t()).bfloat16().float()              # HF's lm_head output is bf16 before the fp32 upcast
    ref = torch.nn.functional.cross_entropy(hr @ W.float().t(), labels, ignore_index=-100)
    ref.backward()
    hn = h.clone().requires_grad_(True)
    loss = LN.linear_cross_entropy(hn, W, labels, chunk_rows=chunk)
    loss.backward()
    assert abs(float(loss) - float(ref)) <= 3e-3 * max(1.0, abs(float(ref)))
    assert rel(hn.grad, hr.grad) <= 2e-2
    # every label ignored: loss 0, gradient 0, no NaN
    hz = h.clone().requires_grad_(True)
    lz = LN.linear_cross_entropy(hz, W, torch.full_like(labels, -100), chunk_rows=chunk)
    lz.backward()
    assert float(lz) == 0.0 and float(hz.grad.abs().max()) == 0.0
    del logits
    # a label >= vocab (torch asserts on it): nothing outside the logits row is read and the loss turns NaN
    bad = labels.clone()
    bad[1] = V + 5
    hb = h.clone().requires_grad_(True)
    lb = LN.linear_cross_entropy(hb, W, bad, chunk_rows=chunk)
    assert torch.isnan(lb)


def test_causal_lm_loss_shift():
    """llama_native.causal_lm_loss([B, S, d] hidden, lm_head, labels) == the oracle's ForCausalLMLoss restatement."""
    g = torch.Generator().manual_seed(9)
    h = torch.randn(3, 41, 128, generator=g).bfloat16().cuda()
    W = (torch.randn(777, 128, generator=g) * 0.05).bfloat16().cuda()
    labels = torch.randint(0, 777, (3, 41), generator=g)
    labels[1, 30:] = -100
    labels[:, :5] = -100
    ref = OL.causal_lm_loss(h.float().cpu() @ W.float().cpu().t(), labels)
    got = LN.causal_lm_loss(h, W, labels.cuda(), chunk_rows=32)
    assert abs(float(got) - float(ref)) <= 3e-3 * abs(float(ref))


def test_native_llama_in_audio_llm_matches_hf():
    """AudioLLM.enable_native_llama_ops(): loss and LoRA gradients of a bf16 LLaMA equal the stock HF path's."""
    from unittest.mock import Mock, patch
    from transformers import LlamaConfig, LlamaForCausalLM
    from audio_llama_b200.models import base as B
    from audio_llama_b200.models.allm import AudioLLM
    from audio_llama_b200.config import EncoderConfig
    from audio_llama_b200.encoder import WhisperEncoderModule
    from audio_llama_b200 import synth

    def fake(lp, wp):
        torch.manual_seed(0)
        lc = LlamaConfig(vocab_size=322, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                         num_attention_heads=4, num_key_value_heads=2)
        ec = EncoderConfig(d_model=128, n_layers=1, n_heads=2, ffn_dim=256, n_mels=80)
        return (B.FrozenModelWrapper(LlamaForCausalLM(lc).to(torch.bfloat16)),
                B.FrozenModelWrapper(WhisperEncoderModule(ec, synth.init_encoder_weights(ec), max_batch=2)))

    def run(native):
        with patch.object(B, "load_base_models", fake):
            m = _manual(AudioLLM("x", "y", lora_rank=8))
        g = torch.Generator().manual_seed(3)
        for l in m.lora_layers.values():
            with torch.no_grad():
                l.lora_A.copy_(torch.randn(l.lora_A.shape, generator=g) * 0.05)
                l.lora_B.copy_(torch.randn(l.lora_B.shape, generator=g) * 0.05)
        if native:
            m.enable_native_llama_ops()
        ids, mask, labels = (t.cuda() for t in synth.synth_text(2, 24, 320))
        out = m(input_ids=ids, attention_mask=mask, labels=labels)
        out.loss.backward()
        l0 = m.lora_layers["model.layers.1.mlp.down_proj"]
        l1 = m.lora_layers["model.layers.0.self_attn.q_proj"]
        return float(out.loss), l0.lora_A.grad.float().cpu(), l1.lora_B.grad.float().cpu()

    try:
        a, b = run(False), run(True)
    finally:
        LN.disable_rope_patch()
    assert abs(a[0] - b[0]) <= 2e-2 * abs(a[0])
    assert O.rel_l2(b[1], a[1]) <= 1e-1 and O.rel_l2(b[2], a[2]) <= 1e-1


def test_frozen_linear_and_inference_logits():
    """o_proj / lm_head on the tcgen05 GEMM: forward (ragged vocabulary) and dgrad vs torch, and the no-grad logits of an
    AudioLLM in native mode vs the stock HF path (the generate() prefill)."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 70, 256, generator=g).bfloat16().cuda()
    W = (torch.randn(1003, 256, generator=g) * 0.05).bfloat16().cuda()        # 1003 % 8 != 0: forward only
    with torch.no_grad():
        y = LN.frozen_linear(x, W)
    assert y.shape == (3, 70, 1003) and rel(y, x.float() @ W.float().t()) <= 5e-3
    W2 = (torch.randn(512, 256, generator=g) * 0.05).bfloat16().cuda()
    b2 = torch.randn(512, generator=g).bfloat16().cuda()
    xr = x.float().requires_grad_(True)
    (xr @ W2.float().t() + b2.float()).backward(torch.ones(3, 70, 512, device="cuda"))
    xn = x.clone().requires_grad_(True)
    yn = LN.frozen_linear(xn, W2, b2)
    yn.backward(torch.ones_like(yn))
    assert rel(yn, xr.detach() @ W2.float().t() + b2.float()) <= 5e-3 and rel(xn.grad, xr.grad) <= 1e-2

    from unittest.mock import patch
    from transformers import LlamaConfig, LlamaForCausalLM
    from audio_llama_b200.models import base as B
    from audio_llama_b200.models.allm import AudioLLM
    from audio_llama_b200.config import EncoderConfig
    from audio_llama_b200.encoder import WhisperEncoderModule
    from audio_llama_b200 import synth

    def fake(lp, wp):
        torch.manual_seed(0)
        lc = LlamaConfig(vocab_size=322, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                         num_attention_heads=4, num_key_value_heads=2)
        ec = EncoderConfig(d_model=128, n_layers=1, n_heads=2, ffn_dim=256, n_mels=80)
        return (B.FrozenModelWrapper(LlamaForCausalLM(lc).to(torch.bfloat16)),
                B.FrozenModelWrapper(WhisperEncoderModule(ec, synth.init_encoder_weights(ec), max_batch=2)))

    def logits(native):
        with patch.object(B, "load_base_models", fake):
            m = _manual(AudioLLM("x", "y", lora_rank=8)).eval()
        if native:
            m.enable_fused_lora()
            m.enable_native_llama_ops()
        ids, mask, _ = (t.cuda() for t in synth.synth_text(2, 24, 320))
        with torch.no_grad():
            return m(input_ids=ids, attention_mask=mask).logits.float().cpu(), mask.bool().cpu()

    try:
        (a, keep), (b, _) = logits(False), logits(True)
    finally:
        LN.disable_rope_patch()
    assert O.rel_l2(b[keep], a[keep]) <= 3e-2           # real positions (padded ones see a different mask on purpose)


def test_generate_in_native_mode():
    """AudioLLM.generate() with the native LLaMA ops and the fused LoRA linears: HF's KV-cache decode loop runs on the
    patched modules (RMSNorm / RoPE at one token per step, lm_head on the tcgen05 GEMM) and stays deterministic."""
    from unittest.mock import Mock, patch
    from transformers import LlamaConfig, LlamaForCausalLM
    from audio_llama_b200.models import base as B
    from audio_llama_b200.models.allm import AudioLLM
    from audio_llama_b200.config import EncoderConfig
    from audio_llama_b200.encoder import WhisperEncoderModule
    from audio_llama_b200.features import LogMelExtractor
    from audio_llama_b200 import synth

    def fake(lp, wp):
        torch.manual_seed(0)
        lc = LlamaConfig(vocab_size=322, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                         num_attention_heads=4, num_key_value_heads=2)
        ec = EncoderConfig(d_model=128, n_layers=1, n_heads=2, ffn_dim=256, n_mels=80)
        return (B.FrozenModelWrapper(LlamaForCausalLM(lc).to(torch.bfloat16)),
                B.FrozenModelWrapper(WhisperEncoderModule(ec, synth.init_encoder_weights(ec), max_batch=1, out_dtype=torch.bfloat16)))

    with patch.object(B, "load_base_models", fake):
        m = _manual(AudioLLM("x", "y", lora_rank=8))
    m.projector.to(torch.bfloat16)
    tok = Mock()
    tok.convert_tokens_to_ids = lambda t: {"<audio>": 320, "</audio>": 321}[t]
    tok.pad_token_id, tok.bos_token_id, tok.eos_token_id = 0, 1, None
    tok.decode = lambda t, skip_special_tokens=True: " ".join(str(int(x)) for x in t)
    m.tokenizer = tok
    m.enable_fused_lora()
    m.enable_native_llama_ops()
    ids, mask, _ = (t.cuda() for t in synth.synth_text(1, 8, 320))
    feats = LogMelExtractor(80)([synth.synth_clip(0)], sampling_rate=16000).input_features.unsqueeze(1)
    try:
        a = m.generate(input_ids=ids, attention_mask=mask, audio_features=feats, max_new_tokens=6, do_sample=False,
                       temperature=None, top_p=None)
        b = m.generate(input_ids=ids, attention_mask=mask, audio_features=feats, max_new_tokens=6, do_sample=False,
                       temperature=None, top_p=None)
        # the decode loop itself: 6 new tokens from the conditioned embeddings, identical on a second run
        with torch.no_grad():
            emb, amask, _ = m._conditioned_inputs(ids, mask, feats, None)
            kw = dict(inputs_embeds=emb, attention_mask=amask, max_new_tokens=6, do_sample=False, temperature=None,
                      top_p=None, pad_token_id=0)
            t1 = m.llama.model.generate(**kw)
            t2 = m.llama.model.generate(**kw)
    finally:
        LN.disable_rope_patch()
    assert isinstance(a, str) and a == b          # (the reference slices `outputs[0, input_length:]`, allm.py:333-346)
    assert t1.shape == (1, 6) and torch.equal(t1, t2)
