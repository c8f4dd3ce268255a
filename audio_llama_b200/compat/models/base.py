from audio_llama_b200.models.base import *  # noqa: F401,F403
from audio_llama_b200.models.base import FrozenModelWrapper, load_base_models  # noqa: F401
