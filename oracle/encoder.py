"""ORACLE (test infrastructure, not product code) — CPU fp32 restatement of the frozen Whisper
encoder forward, the AudioProjector, the LoRA update and the splice.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this.

Parity pin: `encoder_forward` is checked against HF `WhisperEncoder` 5.5.0 and `projector_forward`,
`combine`, `lora_linear` against the reference's own modules imported from /root/reference/src
(tests/golden/make_golden.py -> tests/golden/*.npz; tests/test_oracle_golden.py). The reference's
own tests pin shapes and ordering only — those pins are restated in
tests/test_splice_semantics.py.

E2 follows HF models/whisper/modeling_whisper.py:613-647 (forward), :380-414 (layer),
   :279-282,:310 (q scaled together with its bias, k without bias), :215-238 (softmax(QK^T)V, no mask,
   scaling 1.0 because q is pre-scaled), :55-65 (sinusoid table, part of the weights here).
P1 follows /root/reference/src/models/projector.py:11-19.
S1/S2 follow /root/reference/src/models/allm.py:74-89 (labels), :143-170 (concat order), :184-196 (mask).
L1 follows /root/reference/src/models/lora.py:9-21, 41-43.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def encoder_forward(w: Dict[str, torch.Tensor], cfg, mel: torch.Tensor,
                    return_layers: bool = False):
    """mel [B, n_mels, 3000] fp32 -> [B, 1500, d] fp32."""
    d, H = cfg.d_model, cfg.n_heads
    hd = d // H
    if mel.shape[-1] != 2 * cfg.n_ctx:
        raise ValueError(f"Whisper expects the mel input features to be of length {2 * cfg.n_ctx}, "
                         f"but found {mel.shape[-1]}.")
    x = F.gelu(F.conv1d(mel, w["conv1.weight"], w["conv1.bias"], padding=1))
    x = F.gelu(F.conv1d(x, w["conv2.weight"], w["conv2.bias"], stride=2, padding=1))
    x = x.permute(0, 2, 1) + w["embed_positions.weight"][: cfg.n_ctx]
    B, T, _ = x.shape
    taps = [x.clone()] if return_layers else None
    for l in range(cfg.n_layers):
        p = f"layers.{l}."
        h = F.layer_norm(x, (d,), w[p + "self_attn_layer_norm.weight"], w[p + "self_attn_layer_norm.bias"], 1e-5)
        q = F.linear(h, w[p + "self_attn.q_proj.weight"], w[p + "self_attn.q_proj.bias"]) * (hd ** -0.5)
        k = F.linear(h, w[p + "self_attn.k_proj.weight"])
        v = F.linear(h, w[p + "self_attn.v_proj.weight"], w[p + "self_attn.v_proj.bias"])
        q = q.view(B, T, H, hd).transpose(1, 2)
        k = k.view(B, T, H, hd).transpose(1, 2)
        v = v.view(B, T, H, hd).transpose(1, 2)
        a = torch.softmax(q @ k.transpose(2, 3), dim=-1) @ v
        a = a.transpose(1, 2).reshape(B, T, d)
        x = x + F.linear(a, w[p + "self_attn.out_proj.weight"], w[p + "self_attn.out_proj.bias"])
        h = F.layer_norm(x, (d,), w[p + "final_layer_norm.weight"], w[p + "final_layer_norm.bias"], 1e-5)
        h = F.gelu(F.linear(h, w[p + "fc1.weight"], w[p + "fc1.bias"]))
        x = x + F.linear(h, w[p + "fc2.weight"], w[p + "fc2.bias"])
        if return_layers:
            taps.append(x.clone())
    out = F.layer_norm(x, (d,), w["layer_norm.weight"], w["layer_norm.bias"], 1e-5)
    return (out, taps) if return_layers else out


def projector_forward(w: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """LN(W2 gelu_erf(W1 x + b1) + b2), eps 1e-5 (projector.py:11-19)."""
    h = F.gelu(F.linear(x, w["layers.0.weight"], w["layers.0.bias"]))
    y = F.linear(h, w["layers.2.weight"], w["layers.2.bias"])
    return F.layer_norm(y, (y.shape[-1],), w["layers.3.weight"], w["layers.3.bias"], 1e-5)


def lora_linear(x: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor],
                lora_A: torch.Tensor, lora_B: torch.Tensor, scaling: float) -> torch.Tensor:
    """Frozen linear + hook: out + (x @ (B @ A).T) * scaling (lora.py:20-21, 41-43)."""
    return F.linear(x, W, bias) + (x @ (lora_B @ lora_A).T) * scaling


# ----------------------------------------------------------------------------- splice (S1/S2)
def splice_index_map(t_txt: int, n_audio: int = 1500) -> np.ndarray:
    """int64 [n_audio + 2 + t_txt]: source code per output row —
    -1 = <audio> row, -2 = </audio> row, 0..n_audio-1 encoded as (1<<40)+r = audio row r,
    j >= 0 (small) = text row j. Row 0 <audio>; 1..n_audio audio; n_audio+1 </audio>; then text
    (allm.py:165-170)."""
    m = np.empty(n_audio + 2 + t_txt, dtype=np.int64)
    m[0] = -1
    m[1:1 + n_audio] = (1 << 40) + np.arange(n_audio)
    m[1 + n_audio] = -2
    m[2 + n_audio:] = np.arange(t_txt)
    return m


def combine(embed_table: torch.Tensor, input_ids: torch.Tensor, projected: torch.Tensor,
            start_id: int, end_id: int) -> torch.Tensor:
    """S1: cat([E[<audio>], proj, E[</audio>], E[input_ids]], dim=1); ValueError on out-of-vocab
    delimiter ids (allm.py:140-141)."""
    vocab = embed_table.shape[0]
    if start_id >= vocab or end_id >= vocab:
        raise ValueError(f"Token IDs {start_id}, {end_id} are outside vocabulary size {vocab}")
    B = input_ids.shape[0]
    s = embed_table[torch.full((B, 1), start_id, dtype=torch.long)]
    e = embed_table[torch.full((B, 1), end_id, dtype=torch.long)]
    return torch.cat([s, projected.to(embed_table.dtype), e, embed_table[input_ids]], dim=1)


def extend_mask(attention_mask: torch.Tensor, audio_seq_len: int, has_special_tokens: bool = True) -> torch.Tensor:
    """S2: cat([ones(B, A(+2)) float32, attention_mask]) — result promotes to float32 (allm.py:184-196)."""
    n = audio_seq_len + 2 if has_special_tokens else audio_seq_len
    ones = torch.ones(attention_mask.shape[0], n)
    return torch.cat([ones, attention_mask], dim=1)


def extend_labels(labels: torch.Tensor, audio_embed_len: int) -> torch.Tensor:
    """S2: cat([full(-100, (B, A+2)), labels]) (allm.py:81-89)."""
    pad = torch.full((labels.shape[0], audio_embed_len), -100, dtype=labels.dtype)
    return torch.cat([pad, labels], dim=1)


# ----------------------------------------------------------------------------- ragged extension (config 5)
def ragged_layout(n_audio_rows: Sequence[Sequence[int]], t_txt: int):
    """Extension row of SURVEY.md §8 (NOT in the reference — parity unpinned by it).

    n_audio_rows[b] = rows kept per clip of sample b. Returns (span_offsets, text_offset, total, S_max):
    span_offsets[b][i] = output row of the i-th span's <audio> token (exclusive prefix sum over a_i+2);
    text_offset[b] = first text row; total[b] = text_offset + t_txt; S_max = max total.
    """
    span_offsets: List[List[int]] = []
    text_offset: List[int] = []
    for rows in n_audio_rows:
        off, cur = [], 0
        for a in rows:
            off.append(cur)
            cur += int(a) + 2
        span_offsets.append(off)
        text_offset.append(cur)
    total = [t + t_txt for t in text_offset]
    return span_offsets, text_offset, total, max(total)


def combine_ragged(embed_table: torch.Tensor, input_ids: torch.Tensor, attention_mask: torch.Tensor,
                   labels: Optional[torch.Tensor], projected: Sequence[Sequence[torch.Tensor]],
                   start_id: int, end_id: int):
    """projected[b][i]: [a_i, d] rows kept for span i of sample b. Output right-padded with zero rows to
    the batch max; mask 1.0 over real rows / 0 over pad; labels -100 over audio, delimiters and pad."""
    B, T = input_ids.shape
    d = embed_table.shape[1]
    rows = [[p.shape[0] for p in ps] for ps in projected]
    span_off, text_off, total, S = ragged_layout(rows, T)
    out = torch.zeros(B, S, d, dtype=embed_table.dtype)
    mask = torch.zeros(B, S, dtype=torch.float32)
    lab = torch.full((B, S), -100, dtype=torch.int64)
    for b in range(B):
        for i, p in enumerate(projected[b]):
            o = span_off[b][i]
            out[b, o] = embed_table[start_id]
            out[b, o + 1:o + 1 + p.shape[0]] = p.to(embed_table.dtype)
            out[b, o + 1 + p.shape[0]] = embed_table[end_id]
        t0 = text_off[b]
        out[b, t0:t0 + T] = embed_table[input_ids[b]]
        mask[b, :t0] = 1.0
        mask[b, t0:t0 + T] = attention_mask[b].to(torch.float32)
        if labels is not None:
            lab[b, t0:t0 + T] = labels[b]
    return out, mask, (lab if labels is not None else None)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a - b||_2 / ||b||_2 in float64 — the metric the bf16 tolerance (2e-2) is stated in."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
