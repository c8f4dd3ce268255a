"""Feature side of the reference API on the GPU:

* `LogMelExtractor` stands where `WhisperFeatureExtractor` / `AutoProcessor` stand in
  /root/reference/src/inference.py:38,100-105 — call it with a waveform (numpy / list / tensor) and
  `sampling_rate`, read `.input_features` ([B, n_mels, 3000] float32) from the result;
* `process_audio(...)` keeps the signature of /root/reference/src/inference.py:79-111;
* `train_log_mel(...)` is the MelSpectrogram + log of AudioLLMDataset._process_audio
  (/root/reference/src/dataset.py:101-143) for already-loaded waveforms.

The arithmetic is `al_mel_forward` (one fused kernel + the per-clip floor pass). Only DECODING the file stays on
the host (torchaudio.load, or the stdlib `wave` reader when torchaudio has no backend): the mono mix, the sinc
resampling and the 30 s pad / truncate run on the GPU through `al_ingest_forward` (audio_llama_b200/ingest.py), in
the order each reference caller uses.
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
        n_samples = kwargs.get("n_samples")
        if isinstance(raw_speech, torch.Tensor) and raw_speech.is_cuda and raw_speech.dtype == torch.float32:
            # already on the GPU (e.g. the output of ingest.ingest): no host round trip
            wave = raw_speech if raw_speech.dim() == 2 else raw_speech.unsqueeze(0)
            n = n_samples
        else:
            wave, n = _to_batch(raw_speech, self.device)
        feats = ops.mel_forward(wave.contiguous(), n, n_mels=self.feature_size, mode=ops.MEL_WHISPER)
        return SimpleNamespace(input_features=feats)


def load_audio(audio_path):
    """Decode an audio file on the host -> (waveform [channels, n] float32 in [-1, 1], sample_rate): torchaudio.load as
    the reference calls it (inference.py:84, dataset.py:105), or the stdlib WAV reader (PCM 8/16/24/32-bit) when this
    torchaudio build has no decoding backend."""
    import os
    if not os.path.exists(audio_path):
        raise FileNotFoundError(f"Audio file not found: {audio_path}")          # dataset.py:102-103
    try:
        import torchaudio
        waveform, sr = torchaudio.load(audio_path)
        return waveform.to(torch.float32), int(sr)
    except (ImportError, RuntimeError, OSError):
        import wave
        with wave.open(audio_path, "rb") as f:
            ch, width, sr, n = f.getnchannels(), f.getsampwidth(), f.getframerate(), f.getnframes()
            raw = f.readframes(n)
        if width == 1:
            x = (np.frombuffer(raw, np.uint8).astype(np.float32) - 128.0) / 128.0
        elif width == 2:
            x = np.frombuffer(raw, "<i2").astype(np.float32) / 32768.0
        elif width == 3:
            b = np.frombuffer(raw, np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            x = ((v ^ 0x800000) - 0x800000).astype(np.float32) / 8388608.0
        elif width == 4:
            x = np.frombuffer(raw, "<i4").astype(np.float32) / 2147483648.0
        else:
            raise RuntimeError(f"unsupported WAV sample width {width}")
        return torch.from_numpy(x.reshape(-1, ch).T.copy()), int(sr)


def process_audio(audio_path, processor, max_length=30, sample_rate=16000, device="cuda"):
    """Same contract as the reference's process_audio (/root/reference/src/inference.py:79-111): file ->
    input_features [1, n_mels, 3000] on `device`. The file is decoded on the host; mono mix, resampling to
    `sample_rate` and the truncation to `max_length` seconds (in that order, inference.py:88-98) run on the GPU
    (al_ingest_forward). With a LogMelExtractor as `processor` the waveform never leaves the GPU; any other processor
    (the HF one) is handed the waveform on the host exactly as the reference does."""
    from .ingest import ingest
    waveform, sr = load_audio(audio_path)
    if sample_rate != SAMPLE_RATE or max_length != 30:
        raise ValueError("the GPU ingest path is fixed to 16 kHz / 30 s clips (the reference's defaults)")
    wave, n = ingest([waveform], sr, mode="inference", device=device, target_sr=sample_rate)
    if isinstance(processor, LogMelExtractor):
        return processor(wave, sampling_rate=sample_rate, n_samples=n).input_features
    mono = wave[0, : int(n.item())].cpu()
    return processor(mono, sampling_rate=sample_rate, return_tensors="pt").input_features.to(device)


def dataset_process_audio(audio_path, max_audio_length: int = 30, sample_rate: int = SAMPLE_RATE, n_mels: int = 128,
                          device="cuda") -> torch.Tensor:
    """AudioLLMDataset._process_audio(path) (/root/reference/src/dataset.py:101-143) on the GPU: file ->
    log(MelSpectrogram + 1e-9) [1, 128, 3000] float32 (the TRAINING feature variant, M2). Order as the dataset: pad /
    truncate the decoded file to max_audio_length * sample_rate INPUT samples first (106-112), then mono mix and
    resampling (114-123), then the mel transform (125-133) and the crop to 3000 frames (136-137).
    Bug-compatibility: for a file sampled ABOVE 16 kHz the reference's resampled clip is shorter than 3000 frames and
    its 80-row padding branch (138-140) fails inside torch.cat, so __getitem__ drops the sample (69-72); the same
    RuntimeError is raised here instead of inventing features the reference never produced. (For a file BELOW 16 kHz
    that is longer than 15 s the reference's frame 2999 reads 40 resampled samples past 480 000; here the clip ends at
    480 000 and that one frame sees the reflected edge instead.)"""
    from .ingest import ingest
    waveform, sr = load_audio(audio_path)
    if sample_rate != SAMPLE_RATE or max_audio_length != 30:
        raise ValueError("the GPU ingest path is fixed to 16 kHz / 30 s clips (the reference's defaults)")
    if sr > sample_rate:
        raise RuntimeError(f"Sizes of tensors must match except in dimension 2. Expected size {n_mels} but got size 80 "
                           f"(the reference pads short clips with 80 rows, dataset.py:138-140: a {sr} Hz file cannot be "
                           f"processed by AudioLLMDataset._process_audio)")
    wave, n = ingest([waveform], sr, mode="train", device=device, target_sr=sample_rate)
    return ops.mel_forward(wave, n, n_mels=n_mels, mode=ops.MEL_TRAIN)


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
