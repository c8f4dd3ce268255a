"""audio_llama_b200 — the audio-conditioning path of cdreetz/audio-llama (log-mel -> frozen Whisper encoder ->
AudioProjector -> splice into LLaMA's inputs_embeds) on hand-written sm_100a kernels behind a C ABI.

Layout: csrc/ (CUDA kernels + C ABI), _lib.py / ops.py (ctypes binding), encoder.py, features.py, pipeline.py,
parallel.py, models/ (the reference's module API: allm, projector, lora, base).
"""
__version__ = "0.1.0"
