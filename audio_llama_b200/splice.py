"""Ragged splice — the config-5 EXTENSION of SURVEY.md §8 (the reference supports neither variable-length clips
nor several audio spans per sample: dataset.py:106-112, allm.py:165-170, so its parity is pinned by the repo's
own oracle, not by the reference).

Semantics, chosen so that one full-length clip per sample degenerates to S1/S2 exactly: sample b carries
k_b >= 1 clips; clip i keeps its first a_i = ((n_i // 160) - 1) // 2 + 1 encoder rows (n_i samples, conv2
stride-2 rule). Output rows: [<audio>, a_1 rows, </audio>, <audio>, a_2 rows, </audio>, ..., text..., zero pad]
to the batch max; mask 1 over real rows / 0 over pad; labels -100 over audio, delimiter and pad rows. The span
start offsets are exclusive prefix sums over (a_i + 2), computed on the device inside the kernel.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops
from ._lib import check, lib, ptr, stream_ptr
from .config import HOP, N_CTX, N_SAMPLES


def encoder_rows_for_samples(n: int) -> int:
    n = min(int(n), N_SAMPLES)
    return (n // HOP - 1) // 2 + 1


def splice_ragged(table: torch.Tensor, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor],
                  labels: Optional[torch.Tensor], audio: torch.Tensor, span_rows: Sequence[Sequence[int]],
                  start_id: int, end_id: int):
    """table [V, d]; input_ids/attention_mask/labels [B, T] int64; audio [n_clips, 1500, d] (projected rows of every
    clip in sample order, table dtype); span_rows[b] = rows kept for each clip of sample b.
    Returns (inputs_embeds [B, S, d], mask fp32 [B, S], labels int64 [B, S] | None, span_start int32 [B, max_spans])."""
    vocab, d = table.shape
    if start_id >= vocab or end_id >= vocab:
        raise ValueError(f"Token IDs {start_id}, {end_id} are outside vocabulary size {vocab}")
    B, T = input_ids.shape
    if len(span_rows) != B:
        raise ValueError("span_rows must have one list per sample")
    if audio.dtype != table.dtype or not audio.is_contiguous() or not audio.is_cuda:
        raise TypeError("audio must be a contiguous CUDA tensor of the table's dtype")
    max_spans = max(len(r) for r in span_rows)
    n_clips = sum(len(r) for r in span_rows)
    if n_clips != audio.shape[0]:
        raise ValueError(f"{n_clips} spans but {audio.shape[0]} clips of audio rows")
    rows_t = torch.zeros(B, max_spans, dtype=torch.int32)
    src_t = torch.zeros(B, max_spans, dtype=torch.int32)
    ns_t = torch.zeros(B, dtype=torch.int32)
    clip = 0
    S = 0
    for b, rs in enumerate(span_rows):
        if not rs:
            raise ValueError("every sample needs at least one span")
        ns_t[b] = len(rs)
        for i, a in enumerate(rs):
            if not (1 <= a <= N_CTX):
                raise ValueError(f"span of {a} rows outside 1..{N_CTX}")
            rows_t[b, i] = a
            src_t[b, i] = clip * audio.shape[1]
            clip += 1
        S = max(S, sum(a + 2 for a in rs) + T)
    dev = table.device
    rows_d, src_d, ns_d = rows_t.to(dev), src_t.to(dev), ns_t.to(dev)
    out = torch.empty(B, S, d, dtype=table.dtype, device=dev)
    mask_out = torch.empty(B, S, dtype=torch.float32, device=dev)
    labels_out = torch.empty(B, S, dtype=torch.int64, device=dev) if labels is not None else None
    starts = torch.zeros(B, max_spans, dtype=torch.int32, device=dev)
    check(lib().al_splice_ragged(ptr(table), table.element_size(), d, ptr(input_ids), ptr(attention_mask), ptr(labels),
                                 B, T, S, ptr(rows_d), ptr(src_d), ptr(ns_d), max_spans, ptr(audio), start_id, end_id,
                                 ptr(out), ptr(mask_out), ptr(labels_out), ptr(starts), table.shape[0],
                                 ptr(ops.bad_id_flag(dev)), stream_ptr()),
          "al_splice_ragged")
    ops.raise_if_bad_ids(dev, table.shape[0])
    return out, mask_out, labels_out, starts
