"""GPU parity: waveform ingest (channel mean + sinc resample + pad / truncate) vs the oracle and vs torchaudio's own
outputs (tests/golden/resample.npz). Tolerance 2e-5 absolute on unit-scale audio (torchaudio's float32 dense
convolution vs sparse fp32 accumulation; see test_oracle_golden.test_resample_oracle_matches_torchaudio)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from audio_llama_b200 import ops
from audio_llama_b200.ingest import ingest
from golden_signals import resample_input
from oracle import mel as M
from oracle import resample as R


@pytest.mark.parametrize("sr", [44100, 48000, 8000, 22050, 32000, 16000])
def test_ingest_matches_torchaudio_and_oracle(golden_dir, sr):
    g = np.load(os.path.join(golden_dir, "resample.npz"))
    x = resample_input(sr)
    wave, n = ingest([torch.from_numpy(x)], sr)
    ref = R.ingest_inference(x, sr)
    assert wave.shape == (1, 480000) and int(n[0]) == ref.shape[0]
    y = wave[0, : ref.shape[0]].cpu().numpy()
    assert np.abs(y - ref).max() <= 2e-5
    if sr != 16000:
        assert np.abs(y - g[f"mono_{sr}"]).max() <= 2e-5             # torchaudio itself
    assert (wave[0, ref.shape[0]:] == 0).all()


def test_ingest_batch_ragged_mono_and_train_order():
    rng = np.random.default_rng(0)
    clips = [(0.2 * rng.standard_normal(n)).astype(np.float32) for n in (3000, 12345, 700)]
    wave, n = ingest([torch.from_numpy(c) for c in clips], 22050)
    for i, c in enumerate(clips):
        ref = R.ingest_inference(c, 22050)
        assert int(n[i]) == len(ref)
        assert np.abs(wave[i, : len(ref)].cpu().numpy() - ref).max() <= 2e-5
    # training order: 480 000 INPUT samples kept, then resampled (48 kHz -> 160 000 output samples)
    long = (0.2 * rng.standard_normal((2, 500000))).astype(np.float32)
    wt, nt = ingest([torch.from_numpy(long)], 48000, mode="train")
    ref = R.ingest_train(long, 48000)
    assert int(nt[0]) == 160000 == len(ref)
    assert np.abs(wt[0, :160000].cpu().numpy() - ref).max() <= 2e-5 and (wt[0, 160000:] == 0).all()
    # inference order on the same clip: resample everything, then cut to 30 s
    wi, ni = ingest([torch.from_numpy(long[:, :200000])], 48000)
    assert int(ni[0]) == len(R.ingest_inference(long[:, :200000], 48000))


def test_ingest_feeds_mel():
    """file-rate stereo clip -> ingest -> mel == oracle ingest -> oracle mel."""
    x = resample_input(44100)
    wave, n = ingest([torch.from_numpy(x)], 44100)
    f = ops.mel_forward(wave, n).cpu().numpy()[0]
    ref = M.log_mel_whisper([R.ingest_inference(x, 44100)], 128)[0]
    assert np.abs(f - ref).max() <= 1e-4
