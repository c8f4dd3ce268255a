"""Waveform ingest on the GPU: what the reference does on the CPU between torchaudio.load and
the feature extractor — channel mean, torchaudio sinc resampling to 16 kHz, and the 30 s pad / truncate in the order
each caller uses (/root/reference/src/inference.py:84-98 truncates AFTER resampling;
/root/reference/src/dataset.py:105-123 pads / truncates to 480 000 INPUT samples BEFORE mono-mix and resampling).
Output is the [clips, 480000] zero-padded float32 layout `ops.mel_forward` reads, plus the valid lengths.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

from ._lib import check, lib, ptr, stream_ptr
from .config import N_SAMPLES, SAMPLE_RATE


def ingest(waves: Sequence[torch.Tensor], orig_sr: int, mode: str = "inference", device="cuda",
           target_sr: int = SAMPLE_RATE) -> Tuple[torch.Tensor, torch.Tensor]:
    """waves: list of [C, n] (or [n]) float32 tensors sampled at orig_sr (same channel count).
    Returns (wave [B, 480000] float32 on `device`, n_samples int32 [B])."""
    if mode not in ("inference", "train"):
        raise ValueError("mode must be 'inference' or 'train'")
    clips = [w if w.dim() == 2 else w.unsqueeze(0) for w in waves]
    if not clips:
        raise ValueError("no clips")
    C = clips[0].shape[0]
    if any(c.shape[0] != C for c in clips):
        raise ValueError("all clips of one call must have the same channel count")
    B = len(clips)
    dev = torch.device(device)
    n_max = max(c.shape[1] for c in clips)
    buf = torch.zeros(B, C, max(n_max, 1), dtype=torch.float32)
    lens = torch.zeros(B, dtype=torch.int32)
    for i, c in enumerate(clips):
        buf[i, :, : c.shape[1]] = c.to(torch.float32)
        lens[i] = c.shape[1]
    if mode == "train":
        # dataset.py:106-112: cut / zero-pad every clip to 480 000 samples at the FILE's rate first
        n_in_cap = N_SAMPLES
        lens = torch.full_like(lens, N_SAMPLES)
        if buf.shape[2] < N_SAMPLES:
            buf = torch.nn.functional.pad(buf, (0, N_SAMPLES - buf.shape[2]))
    else:
        n_in_cap = buf.shape[2]
    x = buf.to(dev, non_blocking=True)
    n_in = lens.to(dev, non_blocking=True)
    out = torch.empty(B, N_SAMPLES, dtype=torch.float32, device=dev)
    n_out = torch.empty(B, dtype=torch.int32, device=dev)
    check(lib().al_ingest_forward(ptr(x), x.stride(0), x.stride(1), C, ptr(n_in), n_in_cap, int(orig_sr), int(target_sr),
                                  ptr(out), N_SAMPLES, N_SAMPLES, ptr(n_out), B, stream_ptr()), "al_ingest_forward")
    return out, n_out
