from audio_llama_b200.models.projector import *  # noqa: F401,F403
from audio_llama_b200.models.projector import AudioProjector  # noqa: F401
