#!/usr/bin/env python
"""Inference side (/root/reference/src/inference.py -> AudioLLM.generate): one 30 s clip + a 64-token prompt through the
path, then HF's KV-cache decode loop on the LLaMA (random init, Llama-3.2-3B or -1B shape, bf16). Times the prefill +
N greedy new tokens with the stock HF modules and with enable_fused_lora() + enable_native_llama_ops()."""
import argparse, json, os, sys, time
from unittest.mock import Mock, patch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_llama_b200 import synth, llama_native
from audio_llama_b200.config import WHISPER_LARGE_V3_TURBO
from audio_llama_b200.features import LogMelExtractor
from audio_llama_b200.models import base as B
from audio_llama_b200.models.allm import AudioLLM

LLAMAS = {
    "3b": dict(hidden_size=3072, intermediate_size=8192, num_hidden_layers=28, num_attention_heads=24, num_key_value_heads=8, vocab_size=128258),
    "1b": dict(hidden_size=2048, intermediate_size=8192, num_hidden_layers=16, num_attention_heads=32, num_key_value_heads=8, vocab_size=128258),
}
ap = argparse.ArgumentParser()
ap.add_argument("--llama", default="3b")
ap.add_argument("--new-tokens", type=int, default=32)
args = ap.parse_args()
dev = torch.device("cuda", 0)


def build(native):
    def fake(lp, wp):
        from transformers import LlamaConfig, LlamaForCausalLM
        from audio_llama_b200.encoder import WhisperEncoderModule
        torch.manual_seed(0)
        with torch.device(dev):
            llama = LlamaForCausalLM(LlamaConfig(max_position_embeddings=4096, **LLAMAS[args.llama])).to(torch.bfloat16)
        enc = WhisperEncoderModule(WHISPER_LARGE_V3_TURBO, synth.init_encoder_weights(WHISPER_LARGE_V3_TURBO, seed=0), max_batch=1,
                                   out_dtype=torch.bfloat16)
        return B.FrozenModelWrapper(llama), B.FrozenModelWrapper(enc)
    # native=False keeps the reference-style hooks (the default for bf16 CUDA weights is the fused / native path)
    with patch.object(B, "load_base_models", fake), patch.dict(os.environ, {"AUDIOLLM_B200_NATIVE": "1" if native else "0"}):
        m = AudioLLM("x", "y", lora_rank=64).to(dev)
    m.projector.to(torch.bfloat16)
    v = LLAMAS[args.llama]["vocab_size"]
    tok = Mock()
    tok.convert_tokens_to_ids = lambda t: {"<audio>": v - 2, "</audio>": v - 1}[t]
    tok.pad_token_id, tok.bos_token_id, tok.eos_token_id = 0, 1, None
    tok.decode = lambda t, skip_special_tokens=True: ""
    m.tokenizer = tok
    if native:
        m.enable_fused_lora()
        m.enable_native_llama_ops()
    return m


res = {}
ids = torch.randint(0, 1000, (1, 64), device=dev)
mask = torch.ones(1, 64, dtype=torch.int64, device=dev)
feats = LogMelExtractor(128, device=dev)([synth.synth_clip(0)], sampling_rate=16000).input_features.unsqueeze(1)
for native in (False, True):
    m = build(native)
    ts = []
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.generate(input_ids=ids, attention_mask=mask, audio_features=feats, max_new_tokens=args.new_tokens, do_sample=False,
                   temperature=None, top_p=None, min_new_tokens=args.new_tokens)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    res["native" if native else "hf"] = {"generate_s": ts, "best_s": min(ts[1:])}
    del m
    llama_native.disable_rope_patch()
    torch.cuda.empty_cache()
res["speedup"] = res["hf"]["best_s"] / res["native"]["best_s"]
res["config"] = {"llama": args.llama, "prompt_tokens": 1502 + 64, "new_tokens": args.new_tokens}
print(json.dumps(res))
