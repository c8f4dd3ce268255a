"""LLaMA-side native ops: RMSNorm, SwiGLU, rotary embedding, causal GQA attention (forward and
backward), the frozen linears and the lm_head + cross-entropy tail of the HF LlamaForCausalLM that the reference drives
(/root/reference/src/models/allm.py:99-104), as hand-written sm_100a kernels wired in through torch.autograd.Function.

`enable(audio_llm)` patches the LLaMA inside an AudioLLM in place (the module API stays the reference's):
  * every LlamaRMSNorm.forward          -> al_rmsnorm_forward / _backward         (weight frozen: dx only)
  * every LlamaMLP.forward              -> gate / up / down linears unchanged (fused LoRA when enabled), the
                                           `act_fn(gate) * up` between them -> al_swiglu_forward / _backward
  * modeling_llama.apply_rotary_pos_emb -> al_rope (forward and transposed rotation for the gradient)
  * every LlamaAttention.forward        -> q/k/v projections -> al_rope -> al_gqa_attention_forward / _backward -> o_proj
                                           (head_dim 128, causal + right-padding masks: `attention_plan`; anything else
                                           keeps HF's forward)
  * every LlamaDecoderLayer.forward     -> `_layer_forward`: the same pieces with the residual adds inside the o_proj /
                                           down_proj GEMM epilogues and the gradient sums inside the dgrad GEMMs / the RMSNorm
                                           backward (no elementwise kernels, no autograd accumulation passes)
  * nn.Linear without LoRA (o_proj, lm_head) -> the tcgen05 GEMM forward and dgrad (frozen weights)
  * the loss (labels given)             -> al_linear_ce: lm_head + cross-entropy per chunk of rows, never forming the
                                           [tokens, vocab] logits (HF upcasts them to fp32: 8 GB at the README batch);
                                           `outputs.logits` is None in that mode.
There is no CPU fallback: inputs that are not bf16 CUDA tensors go to the module's original forward.
"""
from __future__ import annotations

import os
import types

import torch

from ._lib import check, lib, ptr, stream_ptr


def _ok(t: torch.Tensor) -> bool:
    return t.is_cuda and t.dtype == torch.bfloat16


# ----------------------------------------------------------------------------- RMSNorm
class _RMSNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, eps):
        xc = x.contiguous()
        d = xc.shape[-1]
        rows = xc.numel() // d
        y = torch.empty_like(xc)
        rstd = torch.empty(rows, dtype=torch.float32, device=xc.device)
        check(lib().al_rmsnorm_forward(ptr(xc), ptr(weight), ptr(y), ptr(rstd), rows, d, float(eps), stream_ptr()),
              "al_rmsnorm_forward")
        ctx.save_for_backward(xc, weight, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, weight, rstd = ctx.saved_tensors
        d = xc.shape[-1]
        dyc = dy.contiguous()
        dx = torch.empty_like(xc)
        check(lib().al_rmsnorm_backward(ptr(xc), ptr(weight), ptr(rstd), ptr(dyc), ptr(dx), xc.numel() // d, d, stream_ptr()),
              "al_rmsnorm_backward")
        return dx, None, None


def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    return _RMSNormFn.apply(x, weight, eps)


class _ResidualRMSNormFn(torch.autograd.Function):
    """(x, rmsnorm(x)): the residual stream and its normalised copy from one node, so that the backward receives both
    gradients (the one arriving over the residual connection and the one through the norm) and adds them inside the
    rmsnorm backward kernel (`al_rmsnorm_backward_ex`) — autograd would otherwise run a separate elementwise add."""

    @staticmethod
    def forward(ctx, x, weight, eps):
        xc = x.contiguous()
        d = xc.shape[-1]
        rows = xc.numel() // d
        y = torch.empty_like(xc)
        rstd = torch.empty(rows, dtype=torch.float32, device=xc.device)
        check(lib().al_rmsnorm_forward(ptr(xc), ptr(weight), ptr(y), ptr(rstd), rows, d, float(eps), stream_ptr()),
              "al_rmsnorm_forward")
        ctx.save_for_backward(xc, weight, rstd)
        return xc.view_as(xc), y

    @staticmethod
    def backward(ctx, g_x, g_y):
        xc, weight, rstd = ctx.saved_tensors
        if g_y is None:
            return g_x, None, None
        d = xc.shape[-1]
        dyc = g_y.contiguous()
        add = g_x.contiguous() if g_x is not None else None
        dx = torch.empty_like(xc)
        check(lib().al_rmsnorm_backward_ex(ptr(xc), ptr(weight), ptr(rstd), ptr(dyc), ptr(add), ptr(dx), xc.numel() // d, d,
                                           stream_ptr()), "al_rmsnorm_backward_ex")
        return dx, None, None


def residual_rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float):
    return _ResidualRMSNormFn.apply(x, weight, eps)


# ----------------------------------------------------------------------------- SwiGLU
class _SwiGLUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gate, up):
        g, u = gate.contiguous(), up.contiguous()
        h = torch.empty_like(g)
        check(lib().al_swiglu_forward(ptr(g), ptr(u), ptr(h), g.numel(), stream_ptr()), "al_swiglu_forward")
        ctx.save_for_backward(g, u)
        return h

    @staticmethod
    def backward(ctx, dh):
        g, u = ctx.saved_tensors
        dhc = dh.contiguous()
        dg, du = torch.empty_like(g), torch.empty_like(u)
        check(lib().al_swiglu_backward(ptr(g), ptr(u), ptr(dhc), ptr(dg), ptr(du), g.numel(), stream_ptr()), "al_swiglu_backward")
        return dg, du


def swiglu(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    return _SwiGLUFn.apply(gate, up)


# ----------------------------------------------------------------------------- frozen linear (no LoRA on it)
def _weight_t(weight: torch.Tensor) -> torch.Tensor:
    """[in, round8(out)] transpose of a frozen weight, built once per weight tensor."""
    from . import ops

    def build():
        out_dim, in_dim = weight.shape
        ld = (out_dim + 7) // 8 * 8
        wt = torch.zeros(in_dim, ld, dtype=weight.dtype, device=weight.device)
        wt[:, :out_dim] = weight.detach().t()
        return wt
    return _CACHE().get((weight,), build)


_cache = None


def _CACHE():
    global _cache
    if _cache is None:
        from . import ops
        _cache = ops.TensorDerivedCache()
    return _cache


class _FrozenLinearFn(torch.autograd.Function):
    """y = x W^T (+ b) with W frozen: forward and dgrad on the tcgen05 GEMM (o_proj, lm_head: the linears the reference
    puts no LoRA on). Ragged out_features (the reference's vocabulary of 128 258) go through TMA tails."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        from . import ops
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        b = bias.detach().float().contiguous() if bias is not None else None
        out_dim = weight.shape[0]
        ld = (out_dim + 7) // 8 * 8
        if ld == out_dim:
            y = ops.gemm_bf16(x2, weight, b)
        else:                                   # row pitch of the output must be a multiple of 16 bytes for TMA
            ypad = torch.empty(x2.shape[0], ld, dtype=torch.bfloat16, device=x.device)
            check(lib().al_gemm_bf16(ptr(x2), x2.shape[1], x2.numel(), x2.shape[0], 1, ptr(weight), out_dim, x2.shape[1],
                                     ptr(b), ptr(ypad), ld, ypad.numel(), 0, None, 0, None, stream_ptr()), "al_gemm_bf16")
            y = ypad[:, :out_dim]
        ctx.save_for_backward(weight)
        ctx.in_shape = x.shape
        return y.reshape(*x.shape[:-1], out_dim)

    @staticmethod
    def backward(ctx, dy):
        (weight,) = ctx.saved_tensors
        out_dim, in_dim = weight.shape
        ld = (out_dim + 7) // 8 * 8
        if ld != out_dim:
            raise RuntimeError("frozen_linear: backward needs out_features % 8 == 0")
        dy2 = dy.reshape(-1, out_dim).contiguous()
        wt = _weight_t(weight)                  # [in, out]
        dx = torch.empty(dy2.shape[0], in_dim, dtype=torch.bfloat16, device=dy.device)
        check(lib().al_gemm_bf16(ptr(dy2), ld, dy2.numel(), dy2.shape[0], 1, ptr(wt), in_dim, out_dim, None, ptr(dx), in_dim,
                                 dx.numel(), 0, None, 0, None, stream_ptr()), "al_gemm_bf16")
        return dx.reshape(ctx.in_shape), None, None


def frozen_linear(x, weight, bias=None):
    return _FrozenLinearFn.apply(x, weight, bias)


class _FrozenLinearAddFn(torch.autograd.Function):
    """y = addend + x W^T (+ b), W frozen, the add in the GEMM epilogue (`al_linear_add_bf16`): o_proj and the residual
    connection around the attention block. d addend = dy (the same tensor, no copy)."""

    @staticmethod
    def forward(ctx, x, weight, bias, addend):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        out_dim, in_dim = weight.shape
        b = bias.detach().float().contiguous() if bias is not None else None
        a2 = addend.reshape(-1, out_dim).contiguous()
        y = torch.empty(x2.shape[0], out_dim, dtype=torch.bfloat16, device=x.device)
        check(lib().al_linear_add_bf16(ptr(x2), x2.shape[0], in_dim, out_dim, ptr(weight), ptr(b), ptr(a2), ptr(y), stream_ptr()),
              "al_linear_add_bf16")
        ctx.save_for_backward(weight)
        ctx.in_shape = x.shape
        return y.reshape(*x.shape[:-1], out_dim)

    @staticmethod
    def backward(ctx, dy):
        (weight,) = ctx.saved_tensors
        out_dim, in_dim = weight.shape
        dx = None
        if ctx.needs_input_grad[0]:
            dy2 = dy.reshape(-1, out_dim).contiguous()
            wt = _weight_t(weight)                  # [in, out]
            dx = torch.empty(dy2.shape[0], in_dim, dtype=torch.bfloat16, device=dy.device)
            check(lib().al_gemm_bf16(ptr(dy2), out_dim, dy2.numel(), dy2.shape[0], 1, ptr(wt), in_dim, out_dim, None, ptr(dx),
                                     in_dim, dx.numel(), 0, None, 0, None, stream_ptr()), "al_gemm_bf16")
            dx = dx.reshape(ctx.in_shape)
        return dx, None, None, (dy if ctx.needs_input_grad[3] else None)


def frozen_linear_add(x, weight, bias, addend):
    return _FrozenLinearAddFn.apply(x, weight, bias, addend)


# ----------------------------------------------------------------------------- rotary embedding
class _RoPEFn(torch.autograd.Function):
    """x [B, S, H, hd] contiguous; cos / sin [Bc, S, hd] contiguous (Bc = 1 or B)."""

    @staticmethod
    def forward(ctx, x, cos, sin):
        B, S, H, hd = x.shape
        out = torch.empty_like(x)
        check(lib().al_rope(ptr(x), ptr(cos), ptr(sin), ptr(out), B, S, H, hd, cos.shape[0], 0, stream_ptr()), "al_rope")
        ctx.save_for_backward(cos, sin)
        return out

    @staticmethod
    def backward(ctx, dy):
        cos, sin = ctx.saved_tensors
        dyc = dy.contiguous()
        B, S, H, hd = dyc.shape
        dx = torch.empty_like(dyc)
        check(lib().al_rope(ptr(dyc), ptr(cos), ptr(sin), ptr(dx), B, S, H, hd, cos.shape[0], 1, stream_ptr()), "al_rope")
        return dx, None, None


def apply_rotary_pos_emb(q, k, cos, sin, unsqueeze_dim=1):
    """Drop-in for transformers.models.llama.modeling_llama.apply_rotary_pos_emb: q, k [B, H, S, hd] (views of the
    [B, S, H, hd] projection outputs), cos / sin [B or 1, S, hd]."""
    if not (_ok(q) and _ok(k) and _ok(cos) and unsqueeze_dim == 1 and q.shape[-1] % 16 == 0):
        hf = _ORIG.get("rope")
        if hf is None:                       # called directly, without enable(): HF's own function is still in place
            from transformers.models.llama import modeling_llama as ML
            hf = ML.apply_rotary_pos_emb
        return hf(q, k, cos, sin, unsqueeze_dim=unsqueeze_dim)
    cc, ss = cos.contiguous(), sin.contiguous()
    qo = _RoPEFn.apply(q.transpose(1, 2).contiguous(), cc, ss).transpose(1, 2)
    ko = _RoPEFn.apply(k.transpose(1, 2).contiguous(), cc, ss).transpose(1, 2)
    return qo, ko


# ----------------------------------------------------------------------------- lm_head + cross-entropy
def _lm_head_t(weight: torch.Tensor) -> torch.Tensor:
    return _weight_t(weight)


class _LinearCEFn(torch.autograd.Function):
    """loss = mean over labels != -100 of cross_entropy(h W^T, labels); W frozen. The gradient of h is produced in the
    forward (chunk by chunk, right after the chunk's logits) and only scaled in the backward."""

    @staticmethod
    def forward(ctx, h, weight, labels, chunk_rows):
        hc = h.contiguous()
        rows, d = hc.shape
        V = weight.shape[0]
        lab = labels.contiguous()
        n_valid = (lab != -100).sum().clamp(min=1).to(torch.float32)         # stays on the device: no host sync
        nbytes = int(lib().al_linear_ce_workspace_bytes(chunk_rows, V))
        ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=hc.device)
        off = (-ws.data_ptr()) % 1024
        loss_sum = torch.empty(1, dtype=torch.float32, device=hc.device)
        need_dh = ctx.needs_input_grad[0]
        dh = torch.empty_like(hc) if need_dh else None
        check(lib().al_linear_ce(ptr(hc), ptr(weight), ptr(_lm_head_t(weight)) if need_dh else None, ptr(lab), rows, d, V,
                                 1.0, int(chunk_rows), ptr(ws[off:]), ptr(loss_sum), ptr(dh), stream_ptr()), "al_linear_ce")
        if need_dh:
            ctx.save_for_backward(dh, n_valid)
        return (loss_sum / n_valid).squeeze(0)

    @staticmethod
    def backward(ctx, g):
        dh, n_valid = ctx.saved_tensors
        return dh * (g / n_valid).to(dh.dtype), None, None, None


def linear_cross_entropy(h: torch.Tensor, weight: torch.Tensor, shifted_labels: torch.Tensor, chunk_rows: int = 2048):
    """h [tokens, d] bf16, weight [vocab, d] bf16 (frozen lm_head), shifted_labels [tokens] int64 (-100 = ignore)."""
    return _LinearCEFn.apply(h, weight, shifted_labels, chunk_rows)


def causal_lm_loss(h: torch.Tensor, weight: torch.Tensor, labels: torch.Tensor, chunk_rows: int = 2048):
    """The loss of LlamaForCausalLM.forward (HF loss_utils.ForCausalLMLoss): position t predicts labels[t + 1]."""
    shifted = torch.nn.functional.pad(labels, (0, 1), value=-100)[..., 1:]
    return linear_cross_entropy(h.reshape(-1, h.shape[-1]), weight, shifted.reshape(-1), chunk_rows)


def causal_only_mask(attention_mask, labels=None):
    """None when `attention_mask` only right-pads (ones then zeros in every row) AND no padded position enters the
    loss: under causal attention the padded keys lie after every real query, so the real positions see exactly the
    same keys with or without the explicit mask, and HF's SDPA path can then run `is_causal=True` (half the score
    tiles) instead of a dense [B, 1, S, S] bias. The PADDED positions themselves do see other keys without the mask,
    so the mask is only dropped when their labels are all -100 (the reference's dataset pads `labels` with the pad
    token id, dataset.py:82-92, so its batches keep the mask and the stock numerics). Rows padded on the left or in
    the middle keep their mask. One tiny device->host read."""
    if attention_mask is None:
        return None
    m = attention_mask
    ok = (m[:, 1:] <= m[:, :-1]).all() if m.shape[1] > 1 else torch.ones((), dtype=torch.bool, device=m.device)
    if labels is not None:
        ok = ok & ((labels == -100) | (m != 0)).all()
    return None if bool(ok) else attention_mask


# ----------------------------------------------------------------------------- causal GQA attention (head_dim 128)
def gqa_attention_forward(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, kv_len, scale: float):
    """q [B, S, Hq, 128], k / v [B, S, Hkv, 128] bf16 contiguous (q, k after RoPE); kv_len int32 [B] or None.
    Returns (out [B, S, Hq, 128] bf16, lse [B, Hq, S] f32, base-2)."""
    B, S, Hq, D = q.shape
    Hkv = k.shape[2]
    if D != 128 or q.dtype != torch.bfloat16 or not q.is_cuda:
        raise ValueError("gqa_attention: bf16 CUDA tensors with head_dim 128 only")
    q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
    out = torch.empty_like(q)
    lse = torch.empty(B, Hq, S, dtype=torch.float32, device=q.device)
    check(lib().al_gqa_attention_forward(ptr(q), ptr(k), ptr(v), ptr(out), ptr(lse), ptr(kv_len), B, S, Hq, Hkv, D,
                                         float(scale), stream_ptr()), "al_gqa_attention_forward")
    return out, lse


class _GqaAttentionFn(torch.autograd.Function):
    """softmax(scale q k^T + causal + key-padding mask) v with grouped kv heads; backward on the native kernels too
    (scores recomputed from q, k and the saved log-sum-exp)."""

    @staticmethod
    def forward(ctx, q, k, v, kv_len, scale):
        out, lse = gqa_attention_forward(q, k, v, kv_len, scale)
        ctx.save_for_backward(q.contiguous(), k.contiguous(), v.contiguous(), out, lse)
        ctx.kv_len, ctx.scale = kv_len, float(scale)
        return out

    @staticmethod
    def backward(ctx, d_out):
        q, k, v, out, lse = ctx.saved_tensors
        B, S, Hq, D = q.shape
        Hkv = k.shape[2]
        d_out = d_out.contiguous()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        dsum = torch.empty(B, Hq, S, dtype=torch.float32, device=q.device)
        check(lib().al_gqa_attention_backward(ptr(q), ptr(k), ptr(v), ptr(out), ptr(lse), ptr(d_out), ptr(ctx.kv_len), ptr(dq),
                                              ptr(dk), ptr(dv), ptr(dsum), B, S, Hq, Hkv, D, ctx.scale, stream_ptr()),
              "al_gqa_attention_backward")
        return dq, dk, dv, None, None


def gqa_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, kv_len=None, scale: float = None) -> torch.Tensor:
    """q [B, S, Hq, 128], k / v [B, S, Hkv, 128] bf16 on the GPU -> [B, S, Hq, 128]; differentiable."""
    if scale is None:
        scale = q.shape[-1] ** -0.5
    return _GqaAttentionFn.apply(q, k, v, kv_len, scale)


# What AudioLLM.forward hands the patched LlamaAttention.forward for the duration of one model call: whether the native
# attention may be used (the 2-D attention mask was None or purely right-padded) and the per-sample key count.
_ATTN_STATE = {"active": False, "kv_len": None}


_STATIC_PLAN = {"on": False}


class static_attention_plan:
    """Context: attention_plan() takes every 2-D mask as right padding WITHOUT reading it on the host (kv_len = count of
    non-zeros, computed on the device). For callers that have checked their masks once (a fixed synthetic batch, a
    collate_fn that only pads on the right) and replay the step as a CUDA graph."""

    def __enter__(self):
        self.prev = _STATIC_PLAN["on"]
        _STATIC_PLAN["on"] = True

    def __exit__(self, *a):
        _STATIC_PLAN["on"] = self.prev


def attention_plan(attention_mask):
    """(usable, kv_len): the native attention covers causal + right-padding masks. `attention_mask` is the 2-D [B, S]
    mask the reference passes (float after _extend_attention_mask). None -> no padding. A mask that is ones-then-zeros
    in every row -> kv_len = number of ones (padded queries still see every real key before them, exactly as HF's
    causal-and-key-padding bias lets them). Anything else (left padding, holes) -> not usable, stock path with the
    mask. One tiny device->host read."""
    if attention_mask is None:
        return True, None
    if _STATIC_PLAN["on"]:                              # the caller vouches for right padding (no host read: CUDA-graph capture)
        return True, (attention_mask != 0).sum(dim=1).to(torch.int32).contiguous()
    m = attention_mask
    if m.dim() != 2:
        return False, None
    ok = (m[:, 1:] <= m[:, :-1]).all() if m.shape[1] > 1 else torch.ones((), dtype=torch.bool, device=m.device)
    ok = ok & (m[:, 0] != 0).all()                      # every row has at least one real key (HF would hit an all-masked row otherwise)
    if not bool(ok):
        return False, None
    return True, (m != 0).sum(dim=1).to(torch.int32).contiguous()


def _attention_forward(m, hidden_states, position_embeddings=None, attention_mask=None, past_key_values=None, _orig=None,
                       **kwargs):
    """LlamaAttention.forward with the rotary embedding and the attention product on the native kernels: q / k / v stay
    in the [B, S, H, hd] layout the projections write (no transposes), al_rope rotates q and k, al_gqa_attention_* does
    softmax(scale q k^T + causal + key padding) v and its backward. Falls back to HF's forward whenever the case is not
    covered (KV cache in use, head_dim != 128, non-bf16, a mask that is not right padding)."""
    st = _ATTN_STATE
    hd = m.head_dim
    usable = (past_key_values is None and hd == 128 and _ok(hidden_states)
              and m.q_proj.weight.dtype == torch.bfloat16 and position_embeddings is not None)
    if st["active"] and not usable:
        # AudioLLM.forward has replaced the padding mask by kv_len for this call: silently running HF's path without it
        # would attend to the padding
        raise RuntimeError("native attention was planned for this forward but the layer cannot run it "
                           "(KV cache in use, head_dim != 128 or non-bf16 tensors)")
    if not st["active"]:
        return _orig(m, hidden_states, position_embeddings=position_embeddings, attention_mask=attention_mask,
                     past_key_values=past_key_values, **kwargs)
    B, S, _ = hidden_states.shape
    q = m.q_proj(hidden_states).view(B, S, -1, hd)
    k = m.k_proj(hidden_states).view(B, S, -1, hd)
    v = m.v_proj(hidden_states).view(B, S, -1, hd)
    cos, sin = position_embeddings
    cc, ss = cos.contiguous(), sin.contiguous()
    q = _RoPEFn.apply(q.contiguous(), cc, ss)
    k = _RoPEFn.apply(k.contiguous(), cc, ss)
    out = gqa_attention(q, k, v.contiguous(), st["kv_len"], m.scaling)
    return m.o_proj(out.reshape(B, S, -1)), None


def _layer_forward(layer, hidden_states, attention_mask=None, position_ids=None, past_key_values=None, use_cache=False,
                   position_embeddings=None, _orig=None, **kwargs):
    """LlamaDecoderLayer.forward with nothing left to elementwise kernels or to autograd's gradient accumulation:
      x, h   = residual_rmsnorm(x)                      backward: d x(residual) + rmsnorm' in one kernel
      q,k,v  = fused_lora_multi(h)                      backward: dh summed in the dgrad GEMMs' epilogues
      x      = x + o_proj(attention(rope(q), rope(k), v))    the add in o_proj's epilogue
      x, h   = residual_rmsnorm(x);  g, u = fused_lora_multi(h)
      x      = x + down_proj(swiglu(g, u))              the add in down_proj's (fused LoRA) epilogue
    Taken when the native attention is planned for this forward and q/k/v/gate/up/down carry LoRA layers while o_proj
    does not (the reference's default target list); everything else goes through HF's forward over the patched
    sub-modules (same arithmetic, separate add kernels)."""
    st = _ATTN_STATE
    lo = getattr(layer, "_al_lora", None)
    attn, mlp = layer.self_attn, layer.mlp
    if not (st["active"] and lo is not None and past_key_values is None and _ok(hidden_states)
            and position_embeddings is not None and attn.head_dim == 128 and mlp.config.hidden_act == "silu"
            and attn.q_proj.weight.dtype == torch.bfloat16 and hidden_states.shape[-1] % 8 == 0):
        return _orig(layer, hidden_states, attention_mask=attention_mask, position_ids=position_ids,
                     past_key_values=past_key_values, use_cache=use_cache, position_embeddings=position_embeddings, **kwargs)
    from .models.lora import fused_lora_forward_add, fused_lora_multi
    hd = attn.head_dim
    B, S, _ = hidden_states.shape
    x, h = residual_rmsnorm(hidden_states, layer.input_layernorm.weight, layer.input_layernorm.variance_epsilon)
    q, k, v = fused_lora_multi(h, [(attn.q_proj, lo["q_proj"]), (attn.k_proj, lo["k_proj"]), (attn.v_proj, lo["v_proj"])])
    cos, sin = position_embeddings
    cc, ss = cos.contiguous(), sin.contiguous()
    q = _RoPEFn.apply(q.view(B, S, -1, hd), cc, ss)
    k = _RoPEFn.apply(k.view(B, S, -1, hd), cc, ss)
    a = gqa_attention(q, k, v.view(B, S, -1, hd), st["kv_len"], attn.scaling)
    x = frozen_linear_add(a.reshape(B, S, -1), attn.o_proj.weight, attn.o_proj.bias, x)
    x, h = residual_rmsnorm(x, layer.post_attention_layernorm.weight, layer.post_attention_layernorm.variance_epsilon)
    g, u = fused_lora_multi(h, [(mlp.gate_proj, lo["gate_proj"]), (mlp.up_proj, lo["up_proj"])])
    return fused_lora_forward_add(mlp.down_proj, lo["down_proj"], swiglu(g, u), x)


def _layer_lora_table(audio_llm, llama, layer):
    """{projection: LoRALayer} of one decoder layer when it has the reference's default LoRA layout, else None."""
    names = {id(m): n for n, m in llama.named_modules()}
    lora = getattr(audio_llm, "lora_layers", None) or {}
    table = {}
    for proj, mod in (("q_proj", layer.self_attn.q_proj), ("k_proj", layer.self_attn.k_proj), ("v_proj", layer.self_attn.v_proj),
                      ("gate_proj", layer.mlp.gate_proj), ("up_proj", layer.mlp.up_proj), ("down_proj", layer.mlp.down_proj)):
        ll = lora.get(names.get(id(mod)))
        if ll is None or mod.weight.requires_grad or mod.weight.shape[0] % 8 or mod.weight.shape[1] % 8:
            return None
        table[proj] = ll
    o = layer.self_attn.o_proj
    if names.get(id(o)) in lora or o.weight.requires_grad or o.weight.shape[0] % 8:
        return None
    return table


# ----------------------------------------------------------------------------- wiring
_ORIG = {}


def enable(audio_llm, rmsnorm_=True, mlp=True, rope=True, fused_ce=True, causal_only=True, frozen_linears=True,
           attention=True, fused_layers=True):
    """Patch the HF LLaMA inside `audio_llm` (an audio_llama_b200.models.allm.AudioLLM) to the native ops."""
    from transformers.models.llama import modeling_llama as ML
    llama = audio_llm.llama.model
    if frozen_linears:
        # the linears the reference puts no LoRA on (o_proj, lm_head): frozen GEMM + dgrad on the tcgen05 kernel
        for name, mod in llama.named_modules():
            if isinstance(mod, torch.nn.Linear) and name not in audio_llm.lora_layers and not mod.weight.requires_grad:
                def lin_fwd(m, x, _orig=type(mod).forward):
                    # (a ragged out_features — lm_head at the reference's vocabulary — is only taken without autograd: its
                    #  dgrad would need a padded W^T pitch; training goes through al_linear_ce instead)
                    if _ok(x) and m.weight.dtype == torch.bfloat16 and x.shape[-1] % 8 == 0 and \
                            (m.weight.shape[0] % 8 == 0 or not (torch.is_grad_enabled() and x.requires_grad)):
                        return frozen_linear(x, m.weight, m.bias)
                    return _orig(m, x)
                mod.forward = types.MethodType(lin_fwd, mod)
    for mod in llama.modules():
        if rmsnorm_ and isinstance(mod, ML.LlamaRMSNorm):
            def norm_fwd(m, hidden_states, _orig=type(mod).forward):
                if _ok(hidden_states) and m.weight.dtype == torch.bfloat16 and hidden_states.shape[-1] % 8 == 0:
                    return rmsnorm(hidden_states, m.weight, m.variance_epsilon)
                return _orig(m, hidden_states)
            mod.forward = types.MethodType(norm_fwd, mod)
        if attention and isinstance(mod, ML.LlamaAttention) and mod.head_dim == 128:
            def attn_fwd(m, hidden_states, position_embeddings=None, attention_mask=None, past_key_values=None,
                         _orig=type(mod).forward, **kw):
                return _attention_forward(m, hidden_states, position_embeddings, attention_mask, past_key_values,
                                          _orig=_orig, **kw)
            mod.forward = types.MethodType(attn_fwd, mod)
        if mlp and isinstance(mod, ML.LlamaMLP):
            def mlp_fwd(m, x, _orig=type(mod).forward):
                if _ok(x) and m.config.hidden_act == "silu":
                    return m.down_proj(swiglu(m.gate_proj(x), m.up_proj(x)))
                return _orig(m, x)
            mod.forward = types.MethodType(mlp_fwd, mod)
    if fused_layers and attention and rmsnorm_ and mlp and frozen_linears and getattr(audio_llm, "fused_lora", False) \
            and os.environ.get("AUDIOLLM_B200_FUSED_LAYERS", "1") != "0":
        for mod in llama.modules():
            if isinstance(mod, ML.LlamaDecoderLayer):
                mod._al_lora = _layer_lora_table(audio_llm, llama, mod)

                def layer_fwd(m, hidden_states, attention_mask=None, position_ids=None, past_key_values=None, use_cache=False,
                              position_embeddings=None, _orig=type(mod).forward, **kw):
                    return _layer_forward(m, hidden_states, attention_mask, position_ids, past_key_values, use_cache,
                                          position_embeddings, _orig=_orig, **kw)
                mod.forward = types.MethodType(layer_fwd, mod)
    if rope and "rope" not in _ORIG:
        _ORIG["rope"] = ML.apply_rotary_pos_emb
        ML.apply_rotary_pos_emb = apply_rotary_pos_emb
    # the plan (mask -> kv_len) is only taken when EVERY attention layer runs the native kernels
    attn_mods = [mod for mod in llama.modules() if isinstance(mod, ML.LlamaAttention)]
    audio_llm.native_attention = bool(attention) and bool(attn_mods) and all(
        mod.head_dim == 128 and mod.q_proj.weight.dtype == torch.bfloat16 for mod in attn_mods)
    audio_llm.native_ce = bool(fused_ce)
    audio_llm.native_causal_only = bool(causal_only)
    audio_llm.native_llama = True
    return audio_llm


def disable_rope_patch():
    from transformers.models.llama import modeling_llama as ML
    if "rope" in _ORIG:
        ML.apply_rotary_pos_emb = _ORIG.pop("rope")
