"""Synthetic inputs and random-init weights of the named architectures.

There are no checkpoints or audio files on the box, so every test and
bench uses this fixed recipe:

* clip i of a run with base seed s: 0.1*N(0,1) + 0.25*sin(2*pi*f_i*n/16000), f_i = 110*2^(i mod 6),
  clipped to [-1, 1], float32, 16 kHz;
* encoder / projector / LoRA weights: seeded normal init with the same statistics HF uses
  (`_init_weights`: N(0, 0.02) linears and convs, LayerNorm = (1, 0), sinusoid positions —
  HF modeling_whisper.py:55-65, 523-527), held in a plain dict keyed by the HF parameter names
  so that a real `WhisperEncoder.state_dict()` can be dropped in instead.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from .config import EncoderConfig, N_SAMPLES, SAMPLE_RATE, projector_hidden


def synth_clip(i: int, n_samples: int = N_SAMPLES, seed: int = 1234) -> np.ndarray:
    g = np.random.default_rng(seed + i)
    f = 110.0 * (2 ** (i % 6))
    n = np.arange(n_samples, dtype=np.float64)
    x = 0.1 * g.standard_normal(n_samples) + 0.25 * np.sin(2.0 * np.pi * f * n / SAMPLE_RATE)
    return np.clip(x, -1.0, 1.0).astype(np.float32)


def synth_batch(n_clips: int, n_samples: int = N_SAMPLES, seed: int = 1234, first: int = 0) -> np.ndarray:
    """[n_clips, n_samples] float32; clip index = first + row (so ranks can shard one global batch)."""
    return np.stack([synth_clip(first + i, n_samples, seed) for i in range(n_clips)])


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    """Whisper's fixed position table: cat([sin, cos]) — HF modeling_whisper.py:55-65."""
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2, dtype=torch.float32))
    t = torch.arange(length, dtype=torch.float32).view(-1, 1) * inv.view(1, -1)
    return torch.cat([t.sin(), t.cos()], dim=1)


def init_encoder_weights(cfg: EncoderConfig, seed: int = 0, std: float = 0.02,
                         ln_jitter: float = 0.0) -> Dict[str, torch.Tensor]:
    """Random-init Whisper encoder weights, fp32, HF parameter names.

    `ln_jitter` > 0 perturbs LayerNorm gains/biases and linear biases (HF init leaves them at
    1/0, which would hide a wrong bias or gain in a parity test).
    """
    g = torch.Generator().manual_seed(seed)
    d, f = cfg.d_model, cfg.ffn_dim

    def nrm(*shape, s=std):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * s

    def bias(n):
        return nrm(n, s=ln_jitter) if ln_jitter > 0 else torch.zeros(n)

    def gain(n):
        return 1.0 + nrm(n, s=ln_jitter) if ln_jitter > 0 else torch.ones(n)

    w: Dict[str, torch.Tensor] = {}
    w["conv1.weight"] = nrm(d, cfg.n_mels, 3)
    w["conv1.bias"] = bias(d)
    w["conv2.weight"] = nrm(d, d, 3)
    w["conv2.bias"] = bias(d)
    w["embed_positions.weight"] = sinusoids(cfg.n_ctx, d)
    for l in range(cfg.n_layers):
        p = f"layers.{l}."
        w[p + "self_attn_layer_norm.weight"] = gain(d)
        w[p + "self_attn_layer_norm.bias"] = bias(d)
        w[p + "self_attn.q_proj.weight"] = nrm(d, d)
        w[p + "self_attn.q_proj.bias"] = bias(d)
        w[p + "self_attn.k_proj.weight"] = nrm(d, d)           # k_proj has no bias (HF :279)
        w[p + "self_attn.v_proj.weight"] = nrm(d, d)
        w[p + "self_attn.v_proj.bias"] = bias(d)
        w[p + "self_attn.out_proj.weight"] = nrm(d, d)
        w[p + "self_attn.out_proj.bias"] = bias(d)
        w[p + "final_layer_norm.weight"] = gain(d)
        w[p + "final_layer_norm.bias"] = bias(d)
        w[p + "fc1.weight"] = nrm(f, d)
        w[p + "fc1.bias"] = bias(f)
        w[p + "fc2.weight"] = nrm(d, f)
        w[p + "fc2.bias"] = bias(d)
    w["layer_norm.weight"] = gain(d)
    w["layer_norm.bias"] = bias(d)
    return w


def init_projector_weights(d_in: int, d_out: int, hidden: Optional[int] = None,
                           seed: int = 1, ln_jitter: float = 0.0) -> Dict[str, torch.Tensor]:
    """nn.Linear default init (U(-1/sqrt(in), 1/sqrt(in))) for `layers.0`, `layers.2`; LayerNorm `layers.3`.

    Key names = the reference's checkpoint keys (/root/reference/src/models/projector.py:11-16,
    /root/reference/src/train.py:115).
    """
    h = hidden if hidden is not None else projector_hidden(d_in, d_out)
    g = torch.Generator().manual_seed(seed)

    def uni(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(*shape, generator=g, dtype=torch.float32) * 2 - 1) * b

    w = {
        "layers.0.weight": uni((h, d_in), d_in), "layers.0.bias": uni((h,), d_in),
        "layers.2.weight": uni((d_out, h), h), "layers.2.bias": uni((d_out,), h),
        "layers.3.weight": torch.ones(d_out), "layers.3.bias": torch.zeros(d_out),
    }
    if ln_jitter > 0:
        w["layers.3.weight"] = 1.0 + torch.randn(d_out, generator=g) * ln_jitter
        w["layers.3.bias"] = torch.randn(d_out, generator=g) * ln_jitter
    return w


def synth_text(batch: int, t_txt: int, vocab: int, seed: int = 7):
    """input_ids in [0, vocab-2), right-padded attention mask, labels (= ids, -100 on pad)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab - 2, (batch, t_txt), generator=g, dtype=torch.int64)
    lens = torch.randint(t_txt // 2, t_txt + 1, (batch,), generator=g)
    mask = (torch.arange(t_txt).view(1, -1) < lens.view(-1, 1)).to(torch.int64)
    labels = torch.where(mask.bool(), ids, torch.full_like(ids, -100))
    return ids, mask, labels
