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


def _write_wav(path, x, sr):
    """[channels, n] float32 in [-1, 1] -> 16-bit PCM WAV (stdlib)."""
    import wave
    pcm = np.clip(np.round(x.T * 32768.0), -32768, 32767).astype("<i2")
    with wave.open(str(path), "wb") as f:
        f.setnchannels(x.shape[0])
        f.setsampwidth(2)
        f.setframerate(sr)
        f.writeframes(pcm.tobytes())
    return pcm.astype(np.float32).T / 32768.0              # what a decoder returns


def test_process_audio_file_api(tmp_path):
    """features.process_audio(path, processor) -- the reference's inference.process_audio (inference.py:79-111) -- on a
    stereo 22.05 kHz file: decode on the host, mono mix + resample + truncate + log-mel on the GPU."""
    from audio_llama_b200.features import LogMelExtractor, process_audio
    x = resample_input(22050)
    dec = _write_wav(tmp_path / "a.wav", x, 22050)
    f = process_audio(str(tmp_path / "a.wav"), LogMelExtractor(128))
    assert f.shape == (1, 128, 3000) and f.is_cuda
    ref = M.log_mel_whisper([R.ingest_inference(dec, 22050)], 128)
    assert np.abs(f.cpu().numpy() - ref).max() <= 1e-4
    # any other processor (the HF extractor's call convention) gets the mono waveform on the host
    seen = {}

    class HostProcessor:
        def __call__(self, wave, sampling_rate=None, return_tensors=None):
            from types import SimpleNamespace
            seen["wave"], seen["sr"] = wave, sampling_rate
            return SimpleNamespace(input_features=torch.zeros(1, 128, 3000))

    process_audio(str(tmp_path / "a.wav"), HostProcessor())
    want = R.ingest_inference(dec, 22050)
    assert seen["sr"] == 16000 and not seen["wave"].is_cuda and seen["wave"].shape == (len(want),)
    assert np.abs(seen["wave"].numpy() - want).max() <= 2e-5
    with pytest.raises(FileNotFoundError):
        process_audio(str(tmp_path / "missing.wav"), LogMelExtractor(128))


def test_dataset_process_audio_file_api(tmp_path):
    """features.dataset_process_audio(path) -- AudioLLMDataset._process_audio (dataset.py:101-143): pad / truncate the
    decoded file to 480 000 input samples, mono mix, resample, ln(MelSpectrogram + 1e-9), crop to 3000 frames."""
    from audio_llama_b200.features import dataset_process_audio
    rng = np.random.default_rng(3)
    x = (0.2 * rng.standard_normal((2, 30000))).astype(np.float32)          # stereo, 8 kHz, 3.75 s
    dec = _write_wav(tmp_path / "b.wav", x, 8000)
    f = dataset_process_audio(str(tmp_path / "b.wav"))
    assert f.shape == (1, 128, 3000)
    mono16k = R.ingest_train(dec, 8000)[:480000]
    ref = M.log_mel_train([mono16k], 128)[0]
    # (the resampler agrees with torchaudio to 2e-5 absolute on the waveform: bins that hold real signal energy are
    #  compared; the empty upper half of an upsampled 8 kHz file is round-off in both implementations)
    live = ref > -6.0
    assert live.mean() > 0.03
    assert np.abs(f.cpu().numpy() - ref)[live].max() <= 1e-3
    # a file above 16 kHz: the reference's own padding branch fails (80 rows vs 128) and the sample is dropped
    _write_wav(tmp_path / "c.wav", x, 22050)
    with pytest.raises(RuntimeError):
        dataset_process_audio(str(tmp_path / "c.wav"))
