"""The audio-conditioning hot path end to end on one GPU:

    waveform --K-mel--> log-mel --encoder--> [B,1500,d] --projector (+LN store in place)--> inputs_embeds rows 1..1500
                                                      splice gather: delimiters + text rows, mask, labels

i.e. what the reference does across `process_audio` (inference.py:79-111), `AudioLLM._process_audio_features`
(allm.py:198-221), `AudioProjector.forward` (projector.py:18-19) and
`AudioLLM._combine_text_and_audio_embeddings` / `_extend_attention_mask` / label extension
(allm.py:74-89, 109-196). All buffers are preallocated so a step launches kernels only (CUDA-graph friendly).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops
from .config import EncoderConfig, N_CTX, N_FRAMES
from .encoder import WhisperEncoderB200
from .models.projector import projector_forward_raw


class AudioConditioner:
    def __init__(self, cfg: EncoderConfig, encoder_weights: Dict[str, torch.Tensor],
                 projector_weights: Dict[str, torch.Tensor], embed_table: torch.Tensor, start_id: int, end_id: int,
                 max_batch: int, device="cuda"):
        self.cfg = cfg
        self.device = torch.device(device)
        self.max_batch = max_batch
        vocab = embed_table.shape[0]
        if start_id >= vocab or end_id >= vocab:       # allm.py:140-141
            raise ValueError(f"Token IDs {start_id}, {end_id} are outside vocabulary size {vocab}")
        self.start_id, self.end_id = int(start_id), int(end_id)
        self.table = embed_table.to(self.device).contiguous()
        self.encoder = WhisperEncoderB200(cfg, encoder_weights, max_batch, device=self.device, out_dtype=torch.bfloat16)
        self.pw = {k: v.detach().to(self.device, torch.float32).contiguous() for k, v in projector_weights.items()}
        self._pcache: dict = {}
        self.d_out = self.pw["layers.2.weight"].shape[0]
        self._mel = torch.empty(max_batch, cfg.n_mels, N_FRAMES, dtype=torch.float32, device=self.device)
        self._enc = torch.empty(max_batch, N_CTX, cfg.d_model, dtype=torch.bfloat16, device=self.device)

    @torch.no_grad()
    def mel(self, wave: torch.Tensor, n_samples: Optional[torch.Tensor] = None) -> torch.Tensor:
        return ops.mel_forward(wave, n_samples, n_mels=self.cfg.n_mels, out=self._mel[: wave.shape[0]])

    @torch.no_grad()
    def __call__(self, wave: torch.Tensor, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                 labels: Optional[torch.Tensor] = None, n_samples: Optional[torch.Tensor] = None,
                 out: Optional[torch.Tensor] = None):
        """wave [B, n] fp32, input_ids/attention_mask/labels [B, T] int64 (all on the GPU) ->
        (inputs_embeds [B, 1502+T, d_l] in the table's dtype, mask fp32 [B, 1502+T], labels int64 | None)."""
        B, T = input_ids.shape
        if B > self.max_batch:
            raise ValueError(f"batch {B} > max_batch {self.max_batch}")
        mel = self.mel(wave, n_samples)
        enc = self.encoder(mel, out=self._enc[:B])
        S = N_CTX + 2 + T
        if out is None:
            out = torch.empty(B, S, self.d_out, dtype=self.table.dtype, device=self.device)
        # projector: GEMM+GELU, GEMM, LayerNorm stored straight into rows 1..1500 of every sample
        projector_forward_raw(self.pw, enc.view(B * N_CTX, self.cfg.d_model), out=out, rows_per_group=N_CTX,
                              out_group_stride=S, out_row_offset=1, cache=self._pcache)
        # splice: delimiter + text rows, mask, labels (audio rows already in place)
        emb, mask, lab = ops.splice(self.table, input_ids, attention_mask, labels, N_CTX, self.start_id, self.end_id,
                                    audio_rows=None, out=out)
        return emb, mask, lab
