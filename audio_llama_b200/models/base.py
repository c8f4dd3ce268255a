"""load_base_models / FrozenModelWrapper — same names as the reference (/root/reference/src/models/base.py:5-26).

The LLaMA side is untouched HF (out of scope). The Whisper side loads the HF checkpoint for its
weights only and hands them to the B200 encoder (`WhisperEncoderModule`), which stands where
`WhisperModel.from_pretrained(path).encoder` stands in the reference.
"""
from __future__ import annotations

import torch

from ..config import EncoderConfig
from ..encoder import WhisperEncoderModule


class FrozenModelWrapper:
    def __init__(self, model):
        self.model = model
        for param in self.model.parameters():
            param.requires_grad = False

    def forward(self, *args, **kwargs):
        with torch.no_grad():
            return self.model(*args, **kwargs)

    def to(self, device):
        self.model = self.model.to(device)
        return self


def encoder_config_from_hf(hf_cfg) -> EncoderConfig:
    return EncoderConfig(d_model=hf_cfg.d_model, n_layers=hf_cfg.encoder_layers,
                         n_heads=hf_cfg.encoder_attention_heads, ffn_dim=hf_cfg.encoder_ffn_dim,
                         n_mels=hf_cfg.num_mel_bins, n_ctx=hf_cfg.max_source_positions)


def wrap_hf_encoder(hf_encoder, max_batch: int = 32) -> WhisperEncoderModule:
    """HF WhisperEncoder (any init) -> B200 encoder module with the same weights."""
    cfg = encoder_config_from_hf(hf_encoder.config)
    sd = {k: v.detach().float() for k, v in hf_encoder.state_dict().items()}
    return WhisperEncoderModule(cfg, sd, hf_config=hf_encoder.config, max_batch=max_batch)


def load_base_models(llama_model_path, whisper_model_path):
    from transformers import LlamaForCausalLM, WhisperModel
    llama = LlamaForCausalLM.from_pretrained(llama_model_path)
    whisper_encoder = wrap_hf_encoder(WhisperModel.from_pretrained(whisper_model_path).encoder)
    return FrozenModelWrapper(llama), FrozenModelWrapper(whisper_encoder)
