"""ORACLE (test infrastructure, not product code) — CPU restatement of the log-mel front end.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this. The product path (audio_llama_b200.features) never does.

Parity pin: checked against HF `WhisperFeatureExtractor` 5.5.0 (the reference's third-party
dependency, lock pins 4.49.0 — /root/reference/uv.lock:1143-1145) through the reference's own call
site /root/reference/src/inference.py:100-105, via the fixtures in tests/golden/mel_*.npz made by
tests/golden/make_golden.py. The reference's tests hold no numeric vectors for this path.

M1 (inference variant, the parity target) follows
  HF models/whisper/feature_extraction_whisper.py:135-164 (_torch_extract_fbank_features),
  :95-103 (filter bank arguments), :296-303 (pad / truncate to 480 000 samples),
  HF audio_utils.py:263-297 (hertz_to_mel slaney), :299-333 (mel_to_hertz), :356-375 (triangles),
  :527-546 (mel_filter_bank body).
M2 (training variant) follows /root/reference/src/dataset.py:101-143 and torchaudio
  functional.melscale_fbanks (HTK scale, norm=None).
"""
from __future__ import annotations

import numpy as np

N_FFT = 400
HOP = 160
N_SAMPLES = 480000
N_FRAMES = 3000
N_FREQ = 201


# ----------------------------------------------------------------------------- filter banks
def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    mels = 3.0 * f / 200.0
    logstep = 27.0 / np.log(6.4)
    log_region = f >= 1000.0
    return np.where(log_region, 15.0 + np.log(np.maximum(f, 1e-30) / 1000.0) * logstep, mels)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f = 200.0 * m / 3.0
    logstep = np.log(6.4) / 27.0
    return np.where(m >= 15.0, 1000.0 * np.exp(logstep * (m - 15.0)), f)


def _triangles(fft_freqs, filter_freqs):
    """HF audio_utils.py:356-375."""
    diff = np.diff(filter_freqs)
    slopes = filter_freqs[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / diff[:-1]
    up = slopes[:, 2:] / diff[1:]
    return np.maximum(0.0, np.minimum(down, up))


def mel_filter_bank_slaney(n_mels: int, n_freq: int = N_FREQ, fmin: float = 0.0, fmax: float = 8000.0,
                           sr: int = 16000) -> np.ndarray:
    """float64 [n_freq, n_mels]; slaney scale + slaney area norm (HF audio_utils.py:527-546)."""
    mel_pts = np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2)
    hz_pts = _mel_to_hz_slaney(mel_pts)
    fft_freqs = np.linspace(0, sr // 2, n_freq)
    fb = _triangles(fft_freqs, hz_pts)
    enorm = 2.0 / (hz_pts[2:n_mels + 2] - hz_pts[:n_mels])
    return fb * enorm[None, :]


def mel_filter_bank_htk(n_mels: int, n_freq: int = N_FREQ, fmin: float = 0.0, fmax: float = 8000.0) -> np.ndarray:
    """float32 [n_freq, n_mels]; torchaudio functional.melscale_fbanks(mel_scale='htk', norm=None)
    (TA functional.py:518-580). torchaudio does this in float32 torch ops with python-float end points;
    restated with the same torch ops so the zero / non-zero pattern (4 dead filters at 128 mels) matches."""
    import math
    import torch
    all_freqs = torch.linspace(0, fmax, n_freq)
    m_min = 2595.0 * math.log10(1.0 + fmin / 700.0)
    m_max = 2595.0 * math.log10(1.0 + fmax / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up)).numpy().astype(np.float32)


# ----------------------------------------------------------------------------- STFT
def hann_periodic(n: int = N_FFT, dtype=np.float64) -> np.ndarray:
    """torch.hann_window(n) (periodic=True): 0.5 - 0.5 cos(2 pi k / n)."""
    k = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)).astype(dtype)


def pad_or_trim(wave: np.ndarray, n: int = N_SAMPLES) -> np.ndarray:
    """Right-pad with 0.0 / truncate to n samples (HF feature_extraction_whisper.py:296-303)."""
    wave = np.asarray(wave)
    if wave.shape[-1] >= n:
        return wave[..., :n]
    pad = [(0, 0)] * (wave.ndim - 1) + [(0, n - wave.shape[-1])]
    return np.pad(wave, pad)


def stft_power(wave: np.ndarray, dtype=np.float32) -> np.ndarray:
    """|STFT|^2 of one padded clip: torch.stft(n_fft=400, hop=160, hann, center=True, reflect) then
    `[..., :-1].abs() ** 2` (HF :149-150). Returns [201, 3000]."""
    x = np.asarray(wave, dtype=dtype)
    xp = np.pad(x, (N_FFT // 2, N_FFT // 2), mode="reflect")
    n_frames = 1 + (xp.shape[0] - N_FFT) // HOP            # 3001
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(n_frames)[:, None]
    frames = xp[idx] * hann_periodic(N_FFT, dtype)[None, :]
    spec = np.fft.rfft(frames, axis=1)                     # numpy >= 2 keeps float32 -> complex64
    power = (spec.real.astype(dtype) ** 2 + spec.imag.astype(dtype) ** 2)
    return power[:-1].T.astype(dtype)                      # drop last frame -> [201, 3000]


# ----------------------------------------------------------------------------- M1 / M2
def log_mel_whisper(waves, n_mels: int = 128, dtype=np.float32) -> np.ndarray:
    """M1. waves: [B, n] array or list of 1-D arrays (any length) -> [B, n_mels, 3000].

    dtype=float32 restates the reference's arithmetic type; dtype=float64 is the exact value of the
    same formula (used to bound how far two float32 implementations may legitimately differ).
    """
    fb = mel_filter_bank_slaney(n_mels).astype(dtype)      # built in f64, cast (HF :151)
    out = []
    for w in waves:
        p = stft_power(pad_or_trim(np.asarray(w, dtype=np.float32)), dtype)
        mel = fb.T @ p                                      # [n_mels, 3000]
        # log10 evaluated in float64 then rounded: numpy's float32 log10 is 1 ulp off at 1e-10 (gives
        # -10.000001, torch gives -10.0), and the zeros-clip known answer is exactly -1.5.
        logs = np.log10(np.maximum(mel, dtype(1e-10)).astype(np.float64)).astype(dtype)
        logs = np.maximum(logs, logs.max() - dtype(8.0))    # per-clip max (HF :156-158)
        out.append(((logs + dtype(4.0)) / dtype(4.0)).astype(dtype))
    return np.stack(out)


def log_mel_train(waves, n_mels: int = 128, dtype=np.float32) -> np.ndarray:
    """M2. MelSpectrogram(16000, 400, hop 160, n_mels, power 2) -> ln(x + 1e-9) -> first 3000 frames.
    Returns [B, 1, n_mels, 3000] (the dataset keeps the channel axis: dataset.py:125-143)."""
    fb = mel_filter_bank_htk(n_mels).astype(dtype)
    out = []
    for w in waves:
        x = pad_or_trim(np.asarray(w, dtype=np.float32))
        xp = np.pad(x.astype(dtype), (N_FFT // 2, N_FFT // 2), mode="reflect")
        n_frames = 1 + (xp.shape[0] - N_FFT) // HOP
        idx = np.arange(N_FFT)[None, :] + HOP * np.arange(n_frames)[:, None]
        spec = np.fft.rfft(xp[idx] * hann_periodic(N_FFT, dtype)[None, :], axis=1)
        power = (spec.real.astype(dtype) ** 2 + spec.imag.astype(dtype) ** 2).T   # [201, 3001]
        mel = fb.T @ power
        out.append(np.log(mel + dtype(1e-9))[None, :, :N_FRAMES].astype(dtype))
    return np.stack(out)


def encoder_frames_for_samples(n: int) -> int:
    """Config-5 extension (SURVEY.md §8 extension row): encoder rows kept for a clip of n samples —
    mel frames n//160, conv2 stride-2 length rule (HF modeling_whisper.py:532-538)."""
    n = min(int(n), N_SAMPLES)
    return (n // HOP - 1) // 2 + 1
