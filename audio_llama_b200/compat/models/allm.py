from audio_llama_b200.models.allm import *  # noqa: F401,F403
from audio_llama_b200.models.allm import AudioLLM  # noqa: F401
