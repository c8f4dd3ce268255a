"""AudioLLM — the reference's model-module API (/root/reference/src/models/allm.py:8-348) with the
audio-conditioning path running on hand-written sm_100a kernels:

  _process_audio_features  -> al_encoder_forward      (frozen Whisper encoder, E1/E2)
  projector                -> al_projector_forward    (P1)
  _combine_text_and_audio_embeddings / _extend_attention_mask / label extension -> al_splice (S1/S2)

Same constructor, attributes (.llama, .whisper_encoder, .projector, .lora_layers, .hooks, .tokenizer,
.audio_start_token/.audio_end_token), method names, argument meaning and error behaviour. The LLaMA forward /
generate tail stays stock HF (out of scope). Differences, on purpose: none of the reference's print() calls
(allm.py:18,83,88-89,121-122,137,172,217), and the dead conv1/2/3 subsampler (allm.py:39-43) is not built.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from . import base as _base
from .lora import apply_lora_to_llama, lora_forward_hook
from .projector import AudioProjector


class _PlaceAudioRows(torch.autograd.Function):
    """inputs_embeds[:, 1:1+A] = projected_audio, in place, differentiable w.r.t. the audio rows only: the splice
    kernel has already gathered the delimiter and text rows from the frozen table into `emb` (no gradient there), so
    the backward is the slice of the incoming gradient (cast to the projector's dtype)."""

    @staticmethod
    def forward(ctx, emb, projected_audio, A):
        ctx.A = A
        ctx.src_dtype = projected_audio.dtype
        emb[:, 1:1 + A].copy_(projected_audio)
        ctx.mark_dirty(emb)
        return emb

    @staticmethod
    def backward(ctx, grad_out):
        return None, grad_out[:, 1:1 + ctx.A].to(ctx.src_dtype), None


class AudioLLM(nn.Module):
    def __init__(self, llama_path, whisper_path, lora_rank=64):
        super().__init__()
        # looked up through the module so tests can patch `models.base.load_base_models` like the reference's do
        self.llama, self.whisper_encoder = _base.load_base_models(llama_path, whisper_path)

        whisper_dim = self.whisper_encoder.model.config.d_model
        llama_dim = self.llama.model.config.hidden_size
        self.projector = AudioProjector(whisper_dim, llama_dim)

        self.lora_layers = apply_lora_to_llama(self.llama.model, rank=lora_rank)
        self.hooks = []
        for name, module in self.llama.model.named_modules():
            if name in self.lora_layers:
                hook = module.register_forward_hook(
                    lambda mod, inp, out, n=name: lora_forward_hook(mod, inp, out, self.lora_layers[n]))
                self.hooks.append(hook)

        self.audio_start_token = "<audio>"
        self.audio_end_token = "</audio>"
        self.tokenizer = None

    # ------------------------------------------------------------------ fused LoRA (B200 extension)
    def enable_fused_lora(self):
        """Swap every hooked frozen linear's forward for the fused GEMM (frozen product + rank-r update in one
        accumulator, `al_lora_linear_forward`) and drop the forward hooks. Needs the LLaMA weights in bf16 on the
        GPU; other inputs fall back to the module's own forward + the reference-style hook arithmetic."""
        import types
        from .lora import fused_lora_forward
        for h in self.hooks:
            h.remove()
        self.hooks = []
        for name, module in self.llama.model.named_modules():
            if name in self.lora_layers:
                lora = self.lora_layers[name]

                def fwd(mod, x, _lora=lora, _orig=type(module).forward):
                    if x.is_cuda and x.dtype == torch.bfloat16 and mod.weight.dtype == torch.bfloat16:
                        return fused_lora_forward(mod, _lora, x)
                    return _orig(mod, x) + _lora(x)

                module.forward = types.MethodType(fwd, module)
        self.fused_lora = True
        return self

    def enable_native_llama_ops(self, **which):
        """RMSNorm / SwiGLU / rotary embedding / lm_head + cross-entropy of the HF LLaMA on hand-written kernels
        (audio_llama_b200.llama_native). Opt-in: the reference's module API is unchanged."""
        from .. import llama_native
        return llama_native.enable(self, **which)

    # ------------------------------------------------------------------ forward (allm.py:47-106)
    def forward(self, input_ids=None, attention_mask=None, audio_features=None, labels=None, **kwargs):
        device = input_ids.device
        if next(self.llama.model.parameters()).device != device:
            self.llama.model = self.llama.model.to(device)

        if audio_features is not None:
            combined_embeddings, combined_attention_mask, adjusted_labels = self._conditioned_inputs(
                input_ids, attention_mask, audio_features, labels)
        else:
            combined_embeddings = self.llama.model.model.embed_tokens(input_ids)
            combined_attention_mask = attention_mask
            adjusted_labels = labels

        if getattr(self, "native_attention", False):
            # causal + right-padding masks run on the native GQA attention (forward and backward): the key count per
            # sample replaces the mask, HF builds no [B, 1, S, S] bias, padded positions keep HF's exact semantics
            from .. import llama_native
            usable, kv_len = llama_native.attention_plan(combined_attention_mask)
            if usable and combined_embeddings.is_cuda and combined_embeddings.dtype == torch.bfloat16:
                kwargs["use_cache"] = False        # (this forward returns a loss / logits, never a cache: generate() goes elsewhere)
                llama_native._ATTN_STATE.update(active=True, kv_len=kv_len)
                try:
                    return self._llama_forward(combined_embeddings, None, adjusted_labels, kwargs)
                finally:
                    llama_native._ATTN_STATE.update(active=False, kv_len=None)
        if getattr(self, "native_causal_only", False):
            from .. import llama_native
            combined_attention_mask = llama_native.causal_only_mask(combined_attention_mask, adjusted_labels)
        return self._llama_forward(combined_embeddings, combined_attention_mask, adjusted_labels, kwargs)

    def _llama_forward(self, combined_embeddings, combined_attention_mask, adjusted_labels, kwargs):
        if getattr(self, "native_ce", False) and adjusted_labels is not None and combined_embeddings.dtype == torch.bfloat16:
            # lm_head + cross-entropy fused per chunk of rows (no [tokens, vocab] logits; `logits` is None in this mode)
            from transformers.modeling_outputs import CausalLMOutputWithPast
            from .. import llama_native
            kwargs.setdefault("use_cache", False)      # a loss-only forward: no KV cache (HF's default builds and cats one)
            hidden = self.llama.model.model(inputs_embeds=combined_embeddings, attention_mask=combined_attention_mask,
                                            **kwargs).last_hidden_state
            loss = llama_native.causal_lm_loss(hidden, self.llama.model.lm_head.weight, adjusted_labels)
            return CausalLMOutputWithPast(loss=loss, logits=None)
        return self.llama.model(inputs_embeds=combined_embeddings, attention_mask=combined_attention_mask,
                                labels=adjusted_labels, **kwargs)

    def _delimiter_ids(self):
        audio_start_id = self.tokenizer.convert_tokens_to_ids(self.audio_start_token)
        audio_end_id = self.tokenizer.convert_tokens_to_ids(self.audio_end_token)
        vocab_size = self.llama.model.model.embed_tokens.weight.shape[0]
        if audio_start_id >= vocab_size or audio_end_id >= vocab_size:
            raise ValueError(f"Token IDs {audio_start_id}, {audio_end_id} are outside vocabulary size {vocab_size}")
        return audio_start_id, audio_end_id

    def _conditioned_inputs(self, input_ids, attention_mask, audio_features, labels):
        """Encoder -> projector -> one splice launch producing embeds, fp32 mask and labels together."""
        start_id, end_id = self._delimiter_ids()
        table = self.llama.model.model.embed_tokens.weight
        processed_audio = self._process_audio_features(audio_features)
        projected_audio = self.projector(processed_audio)                       # autograd-tracked (trainable)
        B, A, _ = projected_audio.shape
        emb, mask, lab = ops.splice(table.detach(), input_ids, attention_mask, labels, A, start_id, end_id,
                                    audio_rows=None)
        return self._place(emb, projected_audio, A), mask, lab

    @staticmethod
    def _place(emb, projected_audio, A):
        """Rows 1..A of the spliced buffer <- the projected audio, in place (the reference builds a second tensor with a
        4-way torch.cat, allm.py:165-170). When the projector is being trained the copy goes through _PlaceAudioRows
        so that it receives the gradient of exactly those rows."""
        if projected_audio.requires_grad and torch.is_grad_enabled():
            return _PlaceAudioRows.apply(emb, projected_audio, A)
        emb[:, 1:1 + A] = projected_audio.to(emb.dtype)
        return emb

    # ------------------------------------------------------------------ allm.py:109-174
    def _combine_text_and_audio_embeddings(self, text_embeddings, audio_features, input_ids):
        if audio_features is None:
            return text_embeddings
        start_id, end_id = self._delimiter_ids()
        processed_audio = self._process_audio_features(audio_features)
        projected_audio = self.projector(processed_audio)
        table = self.llama.model.model.embed_tokens.weight
        A = projected_audio.shape[1]
        emb, _, _ = ops.splice(table.detach(), input_ids, None, None, A, start_id, end_id, audio_rows=None,
                               want_mask=False, want_labels=False)
        return self._place(emb, projected_audio, A)

    # ------------------------------------------------------------------ allm.py:176-196
    def _extend_attention_mask(self, attention_mask, audio_seq_len, has_special_tokens=True):
        batch_size, text_seq_len = attention_mask.shape
        total_audio_len = audio_seq_len + 2 if has_special_tokens else audio_seq_len
        audio_attention = torch.ones(batch_size, total_audio_len, device=attention_mask.device)
        return torch.cat([audio_attention, attention_mask], dim=1)

    # ------------------------------------------------------------------ allm.py:198-221
    def _process_audio_features(self, audio_features):
        device = audio_features.device
        if next(self.whisper_encoder.model.parameters()).device != device:
            self.whisper_encoder.model = self.whisper_encoder.model.to(device)
        audio_features = audio_features.squeeze(1)
        with torch.no_grad():
            whisper_output = self.whisper_encoder.model(audio_features)
            return whisper_output.last_hidden_state

    def get_trainable_params(self):
        """Return only trainable parameters (projector + LoRA) — allm.py:244-249."""
        params = list(self.projector.parameters())
        for lora in self.lora_layers.values():
            params.extend(list(lora.parameters()))
        return params

    def to(self, device):
        self.llama.to(device)
        self.whisper_encoder.to(device)
        self.projector = self.projector.to(device)
        for layer_name in self.lora_layers:
            self.lora_layers[layer_name] = self.lora_layers[layer_name].to(device)
        out = super().to(device)
        self._auto_enable_native(device)
        return out

    def _auto_enable_native(self, device):
        """On a CUDA device with bf16 LLaMA weights the fused frozen + LoRA GEMM and the native row kernels are the
        DEFAULT (the hook path runs two cuBLAS GEMMs and a dense [out, in] delta per linear). Anything
        else (fp32 / fp16 LLaMA weights, CPU) keeps the reference's hook arithmetic, which stays the tested fallback;
        which path is active is logged once. AUDIOLLM_B200_NATIVE=0 keeps the hooks everywhere."""
        import logging
        import os
        log = logging.getLogger("audio_llama_b200")
        dev = torch.device(device) if not isinstance(device, torch.device) else device
        w = next(self.llama.model.parameters())
        want = dev.type == "cuda" and w.dtype == torch.bfloat16 and os.environ.get("AUDIOLLM_B200_NATIVE", "1") != "0"
        if want and not getattr(self, "fused_lora", False):
            self.enable_fused_lora()
            self.enable_native_llama_ops()
            log.info("AudioLLM: fused frozen+LoRA GEMMs and native LLaMA row kernels enabled (bf16 weights on %s)", dev)
        elif not want and not getattr(self, "_logged_fallback", False):
            self._logged_fallback = True
            log.info("AudioLLM: LoRA runs through the reference-style forward hooks (LLaMA weights %s on %s: the fused "
                     "kernels need bf16 on CUDA)", w.dtype, dev)

    # ------------------------------------------------------------------ allm.py:263-348
    def generate(self, input_ids=None, attention_mask=None, audio_features=None, max_new_tokens=256,
                 temperature=0.7, top_p=0.9, do_sample=True, **kwargs):
        self.eval()
        text_len = input_ids.shape[1]
        with torch.no_grad():
            if audio_features is not None:
                combined_embeddings, combined_attention_mask, _ = self._conditioned_inputs(
                    input_ids, attention_mask, audio_features, None)
                audio_seq_len = combined_embeddings.shape[1] - text_len
            else:
                combined_embeddings = self.llama.model.model.embed_tokens(input_ids)
                combined_attention_mask = attention_mask
        tok = self.tokenizer
        generation_config = {
            "max_new_tokens": max_new_tokens, "temperature": temperature, "top_p": top_p, "do_sample": do_sample,
            "pad_token_id": tok.pad_token_id if tok is not None else None,
            "bos_token_id": tok.bos_token_id if tok is not None else None,
            "eos_token_id": tok.eos_token_id if tok is not None else None,
        }
        generation_config.update(kwargs)
        with torch.no_grad():
            outputs = self.llama.model.generate(inputs_embeds=combined_embeddings,
                                                attention_mask=combined_attention_mask, **generation_config)
        input_length = text_len + (audio_seq_len if audio_features is not None else 0)
        generated_tokens = outputs[0, input_length:]
        if tok is not None:
            return tok.decode(generated_tokens, skip_special_tokens=True)
        return generated_tokens
