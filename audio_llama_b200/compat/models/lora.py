from audio_llama_b200.models.lora import *  # noqa: F401,F403
from audio_llama_b200.models.lora import LoRALayer, apply_lora_to_llama, lora_forward_hook  # noqa: F401
