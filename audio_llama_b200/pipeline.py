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


class HostBatch:
    """Pinned host buffers of one batch in the reference dataloader's layout (waveforms instead of CPU-made mel):
    wave [B, 480000] f32, input_ids / attention_mask / labels [B, T] i64, plus the output buffers the step's
    result is copied back into."""

    def __init__(self, B: int, T: int, d_out: int, dtype, n_samples: int = 480000):
        pin = lambda *shape, dt: torch.empty(*shape, dtype=dt).pin_memory()
        S = N_CTX + 2 + T
        self.wave = pin(B, n_samples, dt=torch.float32)
        self.ids = pin(B, T, dt=torch.int64)
        self.mask = pin(B, T, dt=torch.int64)
        self.labels = pin(B, T, dt=torch.int64)
        self.out_embeds = pin(B, S, d_out, dt=dtype)
        self.out_mask = pin(B, S, dt=torch.float32)
        self.out_labels = pin(B, S, dt=torch.int64)

    def h2d_bytes(self):
        return sum(t.numel() * t.element_size() for t in (self.wave, self.ids, self.mask, self.labels))

    def d2h_bytes(self):
        return sum(t.numel() * t.element_size() for t in (self.out_embeds, self.out_mask, self.out_labels))


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
        if embed_table.dtype not in (torch.bfloat16, torch.float32):
            # inputs_embeds takes the table's dtype and the projector's LayerNorm stores bf16 or fp32 only
            raise TypeError(f"embedding table must be bfloat16 or float32 for the audio-conditioning path, got {embed_table.dtype}")
        self.table = embed_table.to(self.device).contiguous()
        self.encoder = WhisperEncoderB200(cfg, encoder_weights, max_batch, device=self.device, out_dtype=torch.bfloat16)
        self.pw = {k: v.detach().to(self.device, torch.float32).contiguous() for k, v in projector_weights.items()}
        self._pcache: dict = {}
        self.d_out = self.pw["layers.2.weight"].shape[0]
        self._mel = torch.empty(max_batch, cfg.n_mels, N_FRAMES, dtype=torch.float32, device=self.device)
        self._enc = torch.empty(max_batch, N_CTX, cfg.d_model, dtype=torch.bfloat16, device=self.device)
        self._clip_max = torch.empty(max_batch, dtype=torch.int32, device=self.device)   # per-clip log-mel maxima

    # ------------------------------------------------------------------ host-buffer entry point
    def _host_state(self, B, T):
        key = (B, T)
        st = getattr(self, "_hs", None)
        if st is None or st["key"] != key:
            dev = self.device
            S = N_CTX + 2 + T
            mk = lambda: dict(
                wave=torch.empty(B, 480000, dtype=torch.float32, device=dev), ids=torch.empty(B, T, dtype=torch.int64, device=dev),
                mask=torch.empty(B, T, dtype=torch.int64, device=dev), labels=torch.empty(B, T, dtype=torch.int64, device=dev),
                emb=torch.empty(B, S, self.d_out, dtype=self.table.dtype, device=dev),
                mask_out=torch.empty(B, S, dtype=torch.float32, device=dev),
                labels_out=torch.empty(B, S, dtype=torch.int64, device=dev),
                h2d_done=torch.cuda.Event(), compute_done=torch.cuda.Event(), d2h_done=torch.cuda.Event())
            st = dict(key=key, slots=[mk(), mk()], i=0, h2d=torch.cuda.Stream(dev), d2h=torch.cuda.Stream(dev), primed=None)
            self._hs = st
        return st

    @torch.no_grad()
    def run_host(self, batches):
        """Host buffers in, host buffers out: for each HostBatch, copy its inputs to the GPU, run the path, copy
        inputs_embeds / mask / labels back into the batch's pinned output buffers. Copies run on their own
        streams with two device slots, so the H2D of batch i+1 and the D2H of batch i-1 overlap the kernels of
        batch i. Returns when every output is on the host."""
        batches = list(batches)
        if not batches:
            return
        B, T = batches[0].ids.shape
        st = self._host_state(B, T)
        cur = torch.cuda.current_stream(self.device)

        def upload(hb, slot):
            with torch.cuda.stream(st["h2d"]):
                st["h2d"].wait_event(slot["compute_done"])        # the slot's previous kernels have read its inputs
                slot["wave"].copy_(hb.wave, non_blocking=True)
                slot["ids"].copy_(hb.ids, non_blocking=True)
                slot["mask"].copy_(hb.mask, non_blocking=True)
                slot["labels"].copy_(hb.labels, non_blocking=True)
                slot["h2d_done"].record(st["h2d"])

        slots = st["slots"]
        slots[0]["compute_done"].record(cur)
        slots[1]["compute_done"].record(cur)
        upload(batches[0], slots[0])
        for i, hb in enumerate(batches):
            slot = slots[i & 1]
            if i + 1 < len(batches):
                upload(batches[i + 1], slots[(i + 1) & 1])
            cur.wait_event(slot["h2d_done"])
            cur.wait_event(slot["d2h_done"])                      # the slot's previous result has left `emb`
            emb, mo, lo = self(slot["wave"], slot["ids"], slot["mask"], slot["labels"], out=slot["emb"],
                               mask_out=slot["mask_out"], labels_out=slot["labels_out"])
            slot["compute_done"].record(cur)
            with torch.cuda.stream(st["d2h"]):
                st["d2h"].wait_event(slot["compute_done"])
                hb.out_embeds.copy_(emb, non_blocking=True)
                hb.out_mask.copy_(mo, non_blocking=True)
                hb.out_labels.copy_(lo, non_blocking=True)
                slot["d2h_done"].record(st["d2h"])
        cur.wait_stream(st["d2h"])
        cur.wait_stream(st["h2d"])
        self.raise_if_bad_ids()              # (synchronises: every output is on the host when this returns)

    @torch.no_grad()
    def mel(self, wave: torch.Tensor, n_samples: Optional[torch.Tensor] = None, raw: bool = False) -> torch.Tensor:
        """The extractor's features [B, n_mels, 3000] f32 (raw=True: before the per-clip floor, which the encoder's
        first kernel then applies from self._clip_max -- the form __call__ uses)."""
        B = wave.shape[0]
        return ops.mel_forward(wave, n_samples, n_mels=self.cfg.n_mels, out=self._mel[:B], ws=self._clip_max[:B], raw=raw)

    @torch.no_grad()
    def __call__(self, wave: torch.Tensor, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                 labels: Optional[torch.Tensor] = None, n_samples: Optional[torch.Tensor] = None,
                 out: Optional[torch.Tensor] = None, mask_out: Optional[torch.Tensor] = None,
                 labels_out: Optional[torch.Tensor] = None):
        """wave [B, n] fp32, input_ids/attention_mask/labels [B, T] int64 (all on the GPU) ->
        (inputs_embeds [B, 1502+T, d_l] in the table's dtype, mask fp32 [B, 1502+T], labels int64 | None).
        With `out`, `mask_out` and `labels_out` given the call launches kernels only (no allocation, no host sync);
        an input id outside the table is reported by raise_if_bad_ids() (run_host calls it; the kernels never read
        outside the table)."""
        B, T = input_ids.shape
        if B > self.max_batch:
            raise ValueError(f"batch {B} > max_batch {self.max_batch}")
        mel = self.mel(wave, n_samples, raw=True)
        enc = self.encoder(mel, out=self._enc[:B], clip_max=self._clip_max[:B])
        S = N_CTX + 2 + T
        if out is None:
            out = torch.empty(B, S, self.d_out, dtype=self.table.dtype, device=self.device)
        # projector: GEMM+GELU, GEMM, LayerNorm stored straight into rows 1..1500 of every sample
        projector_forward_raw(self.pw, enc.view(B * N_CTX, self.cfg.d_model), out=out, rows_per_group=N_CTX,
                              out_group_stride=S, out_row_offset=1, cache=self._pcache)
        # splice: delimiter + text rows, mask, labels (audio rows already in place)
        emb, mask, lab = ops.splice(self.table, input_ids, attention_mask, labels, N_CTX, self.start_id, self.end_id,
                                    audio_rows=None, out=out, mask_out=mask_out, labels_out=labels_out, check_ids=False)
        return emb, mask, lab

    def raise_if_bad_ids(self):
        """IndexError if any splice launched since the last check met an input id outside the embedding table (what
        the reference's embed_tokens(input_ids) raises on, allm.py:64). Synchronises the device."""
        ops.raise_if_bad_ids(self.device, self.table.shape[0])

    # ------------------------------------------------------------------ config 5: ragged clips, several spans per sample
    @torch.no_grad()
    def forward_ragged(self, wave: torch.Tensor, n_samples: torch.Tensor, spans_per_sample, input_ids: torch.Tensor,
                       attention_mask: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None,
                       n_samples_host=None):
        """EXTENSION (not in the reference, SURVEY.md §8 extension row). wave [n_clips, <=480000] fp32 with n_samples
        [n_clips] int32 valid samples (all clips of all samples, in sample order); spans_per_sample[b] = number of
        clips of sample b. Every clip is zero-padded to 30 s and encoded exactly as the reference would encode it
        (the encoder is never masked, HF modeling_whisper.py:607-610); only its first
        ((n // 160) - 1) // 2 + 1 encoder rows are kept and spliced as <audio> rows </audio> spans before the text.
        n_samples_host: the same lengths as a host sequence (a dataloader has them); without it they are read back from
        the device, which synchronises (the span layout and S_max are host decisions).
        Returns (inputs_embeds [B, S_max, d_l], mask fp32, labels | None, span_start int32 [B, max_spans])."""
        from .splice import encoder_rows_for_samples, splice_ragged
        n_clips = wave.shape[0]
        if sum(spans_per_sample) != n_clips:
            raise ValueError("spans_per_sample must add up to the number of clips")
        mel = ops.mel_forward(wave, n_samples, n_mels=self.cfg.n_mels)
        enc = self.encoder(mel)                                             # chunks of max_batch internally
        proj = torch.empty(n_clips, N_CTX, self.d_out, dtype=self.table.dtype, device=self.device)
        projector_forward_raw(self.pw, enc.view(n_clips * N_CTX, self.cfg.d_model), out=proj.view(n_clips * N_CTX, self.d_out),
                              rows_per_group=n_clips * N_CTX, out_group_stride=0, out_row_offset=0, cache=self._pcache)
        lens = list(n_samples_host) if n_samples_host is not None else n_samples.tolist()
        if len(lens) != n_clips:
            raise ValueError("n_samples_host must hold one length per clip")
        rows, i = [], 0
        for k in spans_per_sample:
            rows.append([encoder_rows_for_samples(n) for n in lens[i:i + k]])
            i += k
        return splice_ragged(self.table, input_ids, attention_mask, labels, proj, rows, self.start_id, self.end_id)
