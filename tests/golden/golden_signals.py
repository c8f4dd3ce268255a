"""Deterministic test signals shared by make_golden.py (run here, against the reference) and the tests
(run anywhere). Pure numpy — no reference import."""
import numpy as np

from audio_llama_b200 import synth


def kat_signals():
    n = np.arange(480000)
    return {
        "zeros": np.zeros(480000, np.float32),
        "sine440": (0.5 * np.sin(2 * np.pi * 440 * n / 16000)).astype(np.float32),
        "noise0": (0.1 * np.random.default_rng(0).standard_normal(480000)).astype(np.float32),
        "noise1_5s": (0.1 * np.random.default_rng(1).standard_normal(80000)).astype(np.float32),
        "synth0": synth.synth_clip(0),
        "synth3": synth.synth_clip(3),
        "synth7_12s": synth.synth_clip(7, n_samples=200000),
    }


def encoder_input(cfg, B):
    t = np.arange(3000, dtype=np.float64)[None, None, :]
    m = np.arange(cfg.n_mels, dtype=np.float64)[None, :, None]
    b = np.arange(B, dtype=np.float64)[:, None, None]
    x = 0.6 * np.sin(0.013 * t + 0.31 * m + b) + 0.4 * np.cos(0.0007 * t * m + 0.5 * b) - 0.2
    return x.astype(np.float32)


def resample_input(sr: int) -> np.ndarray:
    """Stereo test signal [2, ~0.2 s] at `sr` Hz (noise + a 330 Hz tone on channel 0), float32."""
    n = int(sr * 0.2) + 13
    rng = np.random.default_rng(5 + sr)
    x = (0.3 * rng.standard_normal((2, n))).astype(np.float32)
    x[0] += (0.4 * np.sin(2 * np.pi * 330.0 * np.arange(n) / sr)).astype(np.float32)
    return x
