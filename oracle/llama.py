"""ORACLE (test infrastructure, not product code) — CPU restatement of the LLaMA-side row operations that
audio_llama_b200.llama_native replaces. Only tests/ may import this.

The algorithms live in the reference's third-party dependency `transformers` (lock 4.49.0, /root/reference/uv.lock:
1143-1145; installed 5.5.0), reached from /root/reference/src/models/allm.py:99-104 (`self.llama.model(inputs_embeds=...,
labels=...)`). Each function restates one HF definition in fp32 torch ops:
  rmsnorm            HF models/llama/modeling_llama.py:52-70   (LlamaRMSNorm.forward)
  swiglu             HF models/llama/modeling_llama.py:171-186 (LlamaMLP.forward: act_fn(gate_proj(x)) * up_proj(x), SiLU)
  rotate_half / rope HF models/llama/modeling_llama.py:138-168 (rotate_half, apply_rotary_pos_emb)
  causal_lm_loss     HF loss/loss_utils.py:28-67               (fixed_cross_entropy, ForCausalLMLoss: shift, ignore -100, mean)
Parity pin: tests/test_oracle_golden.py::test_llama_oracle_matches_hf runs the HF modules themselves (CPU, fp32) against
these restatements; the reference's own tests hold no vectors for them.
"""
from __future__ import annotations

import torch


def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    xf = x.to(torch.float32)
    var = xf.pow(2).mean(-1, keepdim=True)
    return weight * (xf * torch.rsqrt(var + eps)).to(x.dtype)


def swiglu(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.silu(gate) * up


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def rope(q: torch.Tensor, k: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor):
    """q, k [B, H, S, hd]; cos, sin [B or 1, S, hd] (unsqueeze_dim = 1)."""
    c, s = cos.unsqueeze(1), sin.unsqueeze(1)
    return q * c + rotate_half(q) * s, k * c + rotate_half(k) * s


def causal_lm_loss(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
    """logits [B, S, V], labels [B, S]: position t predicts labels[t + 1]; mean over the labels != ignore_index."""
    logits = logits.float()
    shifted = torch.nn.functional.pad(labels, (0, 1), value=ignore_index)[..., 1:].contiguous()
    return torch.nn.functional.cross_entropy(logits.view(-1, logits.shape[-1]), shifted.view(-1), ignore_index=ignore_index,
                                             reduction="mean")


def gqa_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float, kv_len=None) -> torch.Tensor:
    """What HF LlamaAttention computes between the rotary embedding and o_proj (modeling_llama.py eager_attention_forward
    / the sdpa interface with a causal + key-padding mask), restated in float32: q [B, S, Hq, D], k / v [B, S, Hkv, D]
    -> [B, S, Hq, D]. kv_len [B]: keys at positions >= kv_len[b] are masked (right padding)."""
    B, S, Hq, D = q.shape
    Hkv = k.shape[2]
    g = Hq // Hkv
    qf = q.float().permute(0, 2, 1, 3)                                    # [B, Hq, S, D]
    kf = k.float().permute(0, 2, 1, 3).repeat_interleave(g, dim=1)        # repeat_kv
    vf = v.float().permute(0, 2, 1, 3).repeat_interleave(g, dim=1)
    s = qf @ kf.transpose(2, 3) * scale
    mask = torch.ones(S, S, dtype=torch.bool, device=q.device).tril()
    mask = mask[None, None].expand(B, 1, S, S)
    if kv_len is not None:
        key_ok = torch.arange(S, device=q.device)[None, :] < torch.as_tensor(kv_len, device=q.device)[:, None]
        mask = mask & key_ok[:, None, None, :]
    s = s.masked_fill(~mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    return (p @ vf).permute(0, 2, 1, 3).contiguous()
