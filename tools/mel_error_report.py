#!/usr/bin/env python
"""Measured error of both log-mel kernel forms (tensor-core DFT, CUDA-core FFT) against the HF fixtures and the float64
oracle, per test signal — the numbers the gates in tests/test_gpu_mel.py are derived from (2x the measured value on
the pure tone, where two fp32 transforms legitimately differ near the max - 8 floor). Writes one JSON object."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from golden_signals import kat_signals  # noqa: E402

from audio_llama_b200 import _lib, ops  # noqa: E402
from oracle import mel as M  # noqa: E402


def gpu_mel(x, n_mels=128, mode=ops.MEL_WHISPER):
    w = torch.zeros(1, 480000)
    w[0, :len(x)] = torch.from_numpy(x)
    n = torch.tensor([len(x)], dtype=torch.int32)
    return ops.mel_forward(w.cuda(), n.cuda(), n_mels=n_mels, mode=mode)[0].cpu().numpy()


def main():
    gm = np.load(os.path.join(ROOT, "tests", "golden", "mel_whisper.npz"))
    gt = np.load(os.path.join(ROOT, "tests", "golden", "mel_train.npz"))
    sig = kat_signals()
    out = {}
    for form, name in ((1, "tensor_core"), (0, "fft")):
        _lib.lib().al_mel_set_mode(form)
        rec = {}
        for k in ("sine440", "noise0", "synth0", "synth3"):
            f = gpu_mel(sig[k])
            ref64 = M.log_mel_whisper([sig[k]], 128, dtype=np.float64)[0]
            grid = gm[f"{k}_128_grid"]
            rec[k] = {"max_abs_vs_hf_grid": float(np.abs(f[::8, ::50] - grid).max()),
                      "max_abs_vs_f64": float(np.abs(f - ref64).max()),
                      "rel_l2_vs_f64": float(np.linalg.norm(f - ref64) / np.linalg.norm(ref64)),
                      "hf_vs_f64_grid": float(np.abs(ref64[::8, ::50] - grid).max())}
        # training variant (ln domain, no floor)
        for k in ("noise0", "synth0"):
            f = gpu_mel(sig[k], mode=ops.MEL_TRAIN)
            ref64 = M.log_mel_train([sig[k]], dtype=np.float64)[0, 0]
            sel = ref64 > -15.0
            rec[k + "_train"] = {"max_abs_vs_f64_above_-15": float(np.abs(f - ref64)[sel].max()),
                                 "rel_l2_vs_f64_above_-15": float(np.linalg.norm((f - ref64)[sel]) / np.linalg.norm(ref64[sel]))}
        out[name] = rec
    _lib.lib().al_mel_set_mode(1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
