"""ORACLE (test infrastructure, not product code) — CPU restatement of the reference's waveform ingest:
mono-mix + torchaudio sinc resampling + the two pad / truncate orders.

Only tests/ (and bench.py's CPU legs) may import this.

Parity pin: tests/golden/resample.npz holds outputs of torchaudio.transforms.Resample itself (made by
tests/golden/make_golden.py) — the call sites are /root/reference/src/inference.py:87-98 and
/root/reference/src/dataset.py:105-123. Algorithm: torchaudio functional._get_sinc_resample_kernel /
_apply_sinc_resample_kernel (TA functional.py:1305-1432), method sinc_interp_hann, lowpass_filter_width 6,
rolloff 0.99, kernel built in float64 and cast to float32.
"""
from __future__ import annotations

import math

import numpy as np


def sinc_kernel_bank(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """(kernels float32 [new, 2*width + orig], width, orig, new) after gcd reduction."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    t = np.arange(0, -new, -1, dtype=np.float64)[:, None] / new + idx
    t = np.clip(t * base, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    k = k * window * (base / orig)
    return k.astype(np.float32), width, orig, new


def resample(wave: np.ndarray, orig_freq: int, new_freq: int) -> np.ndarray:
    """wave [..., n] float32 -> [..., ceil(new * n / orig)] float32."""
    if int(orig_freq) == int(new_freq):
        return np.asarray(wave, np.float32)
    k, width, orig, new = sinc_kernel_bank(orig_freq, new_freq)
    x = np.asarray(wave, np.float32)
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    n = x2.shape[1]
    xp = np.pad(x2, ((0, 0), (width, width + orig)))
    n_blocks = (xp.shape[1] - k.shape[1]) // orig + 1
    idx = np.arange(k.shape[1])[None, :] + orig * np.arange(n_blocks)[:, None]
    frames = xp[:, idx]                                        # [B, blocks, L]
    y = np.einsum("bnl,pl->bnp", frames.astype(np.float64), k.astype(np.float64)).astype(np.float32)
    y = y.reshape(x2.shape[0], -1)[:, : math.ceil(new * n / orig)]
    return y.reshape(*lead, -1)


def ingest_inference(wave: np.ndarray, sr: int, target_sr: int = 16000, max_seconds: int = 30) -> np.ndarray:
    """process_audio order (inference.py:84-98): mono mean -> resample -> truncate to max_seconds. Returns [n]."""
    w = np.asarray(wave, np.float32)
    if w.ndim == 2:
        w = w.mean(axis=0, dtype=np.float32) if w.shape[0] > 1 else w[0]
    w = resample(w, sr, target_sr) if sr != target_sr else w
    return w[: target_sr * max_seconds]


def ingest_train(wave: np.ndarray, sr: int, target_sr: int = 16000, max_seconds: int = 30) -> np.ndarray:
    """AudioLLMDataset._process_audio order (dataset.py:105-123): pad / truncate to max_seconds*target_sr INPUT
    samples first (whatever the file's rate is), then mono mean, then resample. Returns [n]."""
    w = np.asarray(wave, np.float32)
    if w.ndim == 1:
        w = w[None]
    m = max_seconds * target_sr
    w = w[:, :m] if w.shape[1] > m else np.pad(w, ((0, 0), (0, m - w.shape[1])))
    w = w.mean(axis=0, dtype=np.float32) if w.shape[0] > 1 else w[0]
    return resample(w, sr, target_sr) if sr != target_sr else w
