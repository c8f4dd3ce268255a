"""Named shapes of the audio-conditioning path.

The reference never states these itself: it reads them off the HuggingFace configs of
whatever checkpoints it is pointed at (`/root/reference/src/models/allm.py:16-17`).
The presets below are the architectures BASELINE.json names (SURVEY.md §8 header).
"""
from dataclasses import dataclass

N_FFT = 400            # HF feature_extraction_whisper.py:60 default, dataset.py:127
HOP = 160              # HF feature_extraction_whisper.py:59 default, dataset.py:128
SAMPLE_RATE = 16000
CLIP_SECONDS = 30
N_SAMPLES = SAMPLE_RATE * CLIP_SECONDS      # 480 000
N_FRAMES = N_SAMPLES // HOP                 # 3000 mel frames per clip
N_FREQ = N_FFT // 2 + 1                     # 201
N_CTX = N_FRAMES // 2                       # 1500 encoder positions (conv2 stride 2)


@dataclass(frozen=True)
class EncoderConfig:
    """Whisper encoder shape (HF modeling_whisper.py:555-580)."""
    d_model: int
    n_layers: int
    n_heads: int
    ffn_dim: int
    n_mels: int = 128
    n_ctx: int = N_CTX

    @property
    def head_dim(self) -> int:
        return self.d_model // self.n_heads


WHISPER_TINY = EncoderConfig(d_model=384, n_layers=4, n_heads=6, ffn_dim=1536, n_mels=80)
WHISPER_TINY_128 = EncoderConfig(d_model=384, n_layers=4, n_heads=6, ffn_dim=1536, n_mels=128)
WHISPER_LARGE_V3_TURBO = EncoderConfig(d_model=1280, n_layers=32, n_heads=20, ffn_dim=5120, n_mels=128)

ENCODERS = {
    "whisper-tiny": WHISPER_TINY,
    "whisper-tiny-128": WHISPER_TINY_128,
    "whisper-large-v3-turbo": WHISPER_LARGE_V3_TURBO,
}

# LLaMA hidden sizes of the configs BASELINE.json names (only the embedding width and the
# LoRA-targeted linear shapes matter to this path).
LLAMA_DIMS = {
    "cfg1-2layer-d256": 256,
    "llama-3.2-1b": 2048,
    "llama-3.2-3b": 3072,
}


def projector_hidden(d_in: int, d_out: int) -> int:
    """`hidden_dim = (input_dim + output_dim) // 2` — /root/reference/src/models/projector.py:8-9."""
    return (d_in + d_out) // 2
