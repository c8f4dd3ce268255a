"""Drop-in shim: put `audio_llama_b200/compat` on sys.path (ahead of the reference's `src/`) and the reference's
own import lines — `from models.allm import AudioLLM`, `from models.projector import AudioProjector`, ... —
resolve to the B200 implementations."""
