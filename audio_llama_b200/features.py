"""Feature side of the reference API on the GPU:

* `LogMelExtractor` stands where `WhisperFeatureExtractor` / `AutoProcessor` stand in
  /root/reference/src/inference.py:38,100-105 — call it with a waveform (numpy / list / tensor) and
  `sampling_rate`, read `.input_features` ([B, n_mels, 3000] float32) from the result;
* `process_audio(...)` keeps the signature of /root/reference/src/inference.py:79-111;
* `train_log_mel(...)` is the MelSpectrogram + log of AudioLLMDataset._process_audio
  (/root/reference/src/dataset.py:101-143) for already-loaded waveforms.

The arithmetic is `al_mel_forward` (one fused kernel + the per-clip floor pass). Decoding / resampling audio
files is not on the hot path and stays with torchaudio, as in the reference.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Sequence, Union

import numpy as np
import torch

from . import ops
from .config import N_SAMPLES, SAMPLE_RATE


def _to_batch(raw, device) -> tuple:
    """list / array / tensor of mono clips -> (wave [B, n] float32 on device, n_samples int32 [B])."""
    if isinstance(raw, torch.Tensor):
        raw = raw.detach()
        clips = [raw] if raw.dim() == 1 else list(raw)
    elif isinstance(raw, np.ndarray):
        clips = [raw] if raw.ndim == 1 else list(raw)
    else:
        clips = list(raw)
        if len(clips) and np.isscalar(clips[0]):
            clips = [np.asarray(clips, dtype=np.float32)]
    lens = [min(int(len(c)), N_SAMPLES) for c in clips]
    n = max(max(lens, default=1), 1)
    buf = torch.zeros(len(clips), n, dtype=torch.float32)
    for i, c in enumerate(clips):
        t = c if isinstance(c, torch.Tensor) else torch.as_tensor(np.asarray(c), dtype=torch.float32)
        buf[i, :lens[i]] = t[:lens[i]].to(torch.float32).cpu()
    return buf.to(device, non_blocking=True), torch.tensor(lens, dtype=torch.int32, device=device)


class LogMelExtractor:
    """Whisper log-mel features on the GPU (M1). `feature_size` = number of mel bins (80 or 128)."""

    def __init__(self, feature_size: int = 128, sampling_rate: int = SAMPLE_RATE, device="cuda"):
        self.feature_size = feature_size
        self.sampling_rate = sampling_rate
        self.device = torch.device(device)

    def __call__(self, raw_speech, sampling_rate: int = None, return_tensors: str = "pt", **kwargs):
        if sampling_rate is not None and sampling_rate != self.sampling_rate:
            # same check as HF feature_extraction_whisper.py:264-270
            raise ValueError(f"The model corresponding to this feature extractor was trained using a sampling rate of "
                             f"{self.sampling_rate}. Please make sure that the provided `raw_speech` input was sampled "
                             f"with {self.sampling_rate} and not {sampling_rate}.")
        wave, n = _to_batch(raw_speech, self.device)
        feats = ops.mel_forward(wave, n, n_mels=self.feature_size, mode=ops.MEL_WHISPER)
        return SimpleNamespace(input_features=feats)


def process_audio(audio_path, processor, max_length=30, sample_rate=16000, device="cuda"):
    """Same contract as the reference's process_audio: file -> input_features [1, n_mels, 3000] on `device`."""
    import torchaudio
    waveform, sr = torchaudio.load(audio_path)
    if waveform.shape[0] > 1:
        waveform = torch.mean(waveform, dim=0, keepdim=True)
    if sr != sample_rate:
        waveform = torchaudio.transforms.Resample(orig_freq=sr, new_freq=sample_rate)(waveform)
    max_samples = sample_rate * max_length
    if waveform.shape[1] > max_samples:
        waveform = waveform[:, :max_samples]
    return processor(waveform.squeeze(0), sampling_rate=sample_rate, return_tensors="pt").input_features.to(device)


def train_log_mel(waveform: Union[torch.Tensor, Sequence], n_mels: int = 128, device="cuda") -> torch.Tensor:
    """M2: [B, n] (or [n]) waveform -> [B, 1, n_mels, 3000] = log(MelSpectrogram(...) + 1e-9)[..., :3000]
    (dataset.py:125-143; the channel axis is kept as the dataset keeps it)."""
    wave, n = _to_batch(waveform, torch.device(device))
    return ops.mel_forward(wave, n, n_mels=n_mels, mode=ops.MEL_TRAIN).unsqueeze(1)


def collate_fn(batch):
    """Same batch layout as the reference's collate_fn (/root/reference/src/dataset.py:186-204): drops items whose
    audio_features is None, stacks each key, keeps `metadata` as a list."""
    valid_items = [item for item in batch if item["audio_features"] is not None]
    if not valid_items:
        raise ValueError("No valid audio features found in batch. Check audio file paths and processing.")
    return {
        "audio_features": torch.stack([item["audio_features"] for item in valid_items]),
        "input_ids": torch.stack([item["input_ids"] for item in valid_items]),
        "attention_mask": torch.stack([item["attention_mask"] for item in valid_items]),
        "labels": torch.stack([item["labels"] for item in valid_items]),
        "metadata": [item.get("metadata", {}) for item in valid_items],
    }
