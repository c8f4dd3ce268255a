"""Frozen Whisper encoder forward on B200 (E1/E2).

Host side of `al_encoder_*`: packs the HF-named fp32 weights once into the layouts the kernels want and
owns the workspace. Mirrors what the reference reaches through
`AudioLLM._process_audio_features` -> `self.whisper_encoder.model(x).last_hidden_state`
(/root/reference/src/models/allm.py:198-221; HF modeling_whisper.py:593-647).
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Dict

import torch

from ._lib import check, lib, ptr, stream_ptr
from .config import EncoderConfig


def _aligned_empty(nbytes: int, device, align: int = 1024) -> torch.Tensor:
    buf = torch.empty(nbytes + align, dtype=torch.uint8, device=device)
    off = (-buf.data_ptr()) % align
    return buf[off:off + nbytes]


class WhisperEncoderB200:
    """mel [B, n_mels, 3000] fp32 -> last_hidden_state [B, 1500, d]. Inference only (the reference keeps the
    encoder frozen and under no_grad: base.py:8-9, allm.py:216)."""

    def __init__(self, cfg: EncoderConfig, weights: Dict[str, torch.Tensor], max_batch: int,
                 device="cuda", out_dtype=torch.bfloat16):
        if cfg.head_dim != 64:
            raise ValueError("Whisper encoders have head_dim 64; got %d" % cfg.head_dim)
        self.cfg = cfg
        self.device = torch.device(device)
        self.max_batch = int(max_batch)
        self.out_dtype = out_dtype
        self._w = self._pack(weights)
        L = lib()
        nbytes = L.al_encoder_workspace_bytes(cfg.d_model, cfg.n_layers, cfg.n_heads, cfg.ffn_dim, cfg.n_mels, self.max_batch)
        self._ws = _aligned_empty(nbytes, self.device)
        h = C.c_void_p()
        check(L.al_encoder_create(C.byref(h), cfg.d_model, cfg.n_layers, cfg.n_heads, cfg.ffn_dim, cfg.n_mels,
                                  self.max_batch, ptr(self._ws), nbytes), "al_encoder_create")
        self._h = h
        check(L.al_encoder_set_options(h, 1), "al_encoder_set_options")   # AL_ATT_Q_LOG2: see _pack
        w = self._w
        check(L.al_encoder_set_stem(h, ptr(w["conv1_w"]), ptr(w["conv1_b"]), ptr(w["conv2_w"]), ptr(w["conv2_b"]),
                                    ptr(w["pos"]), ptr(w["lnf_g"]), ptr(w["lnf_b"])), "al_encoder_set_stem")
        for l in range(cfg.n_layers):
            lw = w["layers"][l]
            check(L.al_encoder_set_layer(h, l, ptr(lw["ln1_g"]), ptr(lw["ln1_b"]), ptr(lw["wqkv"]), ptr(lw["bqkv"]),
                                         ptr(lw["wo"]), ptr(lw["bo"]), ptr(lw["ln2_g"]), ptr(lw["ln2_b"]),
                                         ptr(lw["w1"]), ptr(lw["b1"]), ptr(lw["w2"]), ptr(lw["b2"])),
                  "al_encoder_set_layer")

    # ------------------------------------------------------------------ weight packing (one-off, not hot path)
    def _pack(self, sd: Dict[str, torch.Tensor]):
        cfg, dev = self.cfg, self.device
        d, hd = cfg.d_model, cfg.head_dim
        c_pad = (cfg.n_mels + 63) // 64 * 64
        f = lambda k: sd[k].detach().to(dev, torch.float32)
        bf = lambda t: t.to(torch.bfloat16).contiguous()
        # conv weights [out, in, k] -> [out, k, in(pad)] -> [out, k*in]: column kk*c_in + ci
        c1 = torch.zeros(d, 3, c_pad, device=dev)
        c1[:, :, :cfg.n_mels] = f("conv1.weight").permute(0, 2, 1)
        c2 = f("conv2.weight").permute(0, 2, 1).contiguous()
        out = {
            "conv1_w": bf(c1.view(d, 3 * c_pad)), "conv1_b": f("conv1.bias").contiguous(),
            "conv2_w": bf(c2.view(d, 3 * d)), "conv2_b": f("conv2.bias").contiguous(),
            "pos": f("embed_positions.weight")[: cfg.n_ctx].contiguous(),
            "lnf_g": f("layer_norm.weight").contiguous(), "lnf_b": f("layer_norm.bias").contiguous(),
            "layers": [],
        }
        # q is scaled together with its bias (HF :310): hd^-0.5 (0.125 for head_dim 64), and log2(e) rides along so
        # that the attention kernel's scores arrive in log2 units (AL_ATT_Q_LOG2: exp2 straight from the accumulator).
        # The product is formed in fp32 before the one bf16 rounding of the packed weight.
        scale = hd ** -0.5 * 1.4426950408889634
        for l in range(cfg.n_layers):
            p = f"layers.{l}."
            wq, bq = f(p + "self_attn.q_proj.weight") * scale, f(p + "self_attn.q_proj.bias") * scale
            wk = f(p + "self_attn.k_proj.weight")                       # k_proj has no bias (HF :279)
            wv, bv = f(p + "self_attn.v_proj.weight"), f(p + "self_attn.v_proj.bias")
            out["layers"].append({
                "ln1_g": f(p + "self_attn_layer_norm.weight").contiguous(), "ln1_b": f(p + "self_attn_layer_norm.bias").contiguous(),
                "wqkv": bf(torch.cat([wq, wk, wv], 0)), "bqkv": torch.cat([bq, torch.zeros_like(bq), bv]).contiguous(),
                "wo": bf(f(p + "self_attn.out_proj.weight")), "bo": f(p + "self_attn.out_proj.bias").contiguous(),
                "ln2_g": f(p + "final_layer_norm.weight").contiguous(), "ln2_b": f(p + "final_layer_norm.bias").contiguous(),
                "w1": bf(f(p + "fc1.weight")), "b1": f(p + "fc1.bias").contiguous(),
                "w2": bf(f(p + "fc2.weight")), "b2": f(p + "fc2.bias").contiguous(),
            })
        return out

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, mel: torch.Tensor, n_layers_run: int = -1, out: torch.Tensor = None,
                clip_max: torch.Tensor = None) -> torch.Tensor:
        """clip_max: `mel` is the raw (pre-floor) output of ops.mel_forward(raw=True) and this is its `ws`; the
        per-clip floor and the (x + 4) / 4 step then run inside the plan's first kernel."""
        if mel.dim() == 4:                       # [B, 1, n_mels, 3000] as the dataloader hands it (allm.py:214)
            mel = mel.squeeze(1)
        if mel.shape[-1] != 2 * self.cfg.n_ctx:  # same error as HF modeling_whisper.py:613-617
            raise ValueError(f"Whisper expects the mel input features to be of length {2 * self.cfg.n_ctx}, "
                             f"but found {mel.shape[-1]}. Make sure to pad the input mel features to {2 * self.cfg.n_ctx}.")
        if mel.shape[1] != self.cfg.n_mels:
            raise ValueError(f"expected {self.cfg.n_mels} mel bins, got {mel.shape[1]}")
        if not mel.is_cuda:
            raise ValueError("mel must be on the GPU (no CPU fallback)")
        mel = mel.to(torch.float32).contiguous()
        B = mel.shape[0]
        if out is None:
            out = torch.empty(B, self.cfg.n_ctx, self.cfg.d_model, dtype=self.out_dtype, device=mel.device)
        done = 0
        while done < B:                          # larger batches run in max_batch chunks
            n = min(self.max_batch, B - done)
            check(lib().al_encoder_forward_ex(self._h, ptr(mel[done:]),
                                              ptr(clip_max[done:]) if clip_max is not None else None, n,
                                              ptr(out[done:]), 1 if out.dtype == torch.float32 else 0, n_layers_run,
                                              stream_ptr()), "al_encoder_forward_ex")
            done += n
        return out

    __call__ = forward

    KINDS = ("pack_mel", "conv1", "conv2", "layernorm", "qkv", "attention", "out_proj", "fc1", "fc2")

    def set_profiling(self, on: bool):
        check(lib().al_encoder_set_profiling(self._h, 1 if on else 0), "al_encoder_set_profiling")

    def read_profile(self):
        """{kind: (milliseconds, launches)} accumulated since the last read (synchronises)."""
        ms = (C.c_float * len(self.KINDS))()
        n = (C.c_int * len(self.KINDS))()
        check(lib().al_encoder_profile_read(self._h, ms, n), "al_encoder_profile_read")
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(self.KINDS)}

    def hidden_state(self, B: int) -> torch.Tensor:
        """Copy of the fp32 residual stream [B, 1500, d] after the last forward (tests only)."""
        n = B * self.cfg.n_ctx * self.cfg.d_model
        base = lib().al_encoder_hidden(self._h)
        off = base - self._ws.data_ptr()
        return self._ws[off:off + 4 * n].view(torch.float32).view(B, self.cfg.n_ctx, self.cfg.d_model).clone()

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().al_encoder_destroy(self._h)
                self._h = None
        except Exception:
            pass


class WhisperEncoderModule(torch.nn.Module):
    """Stands where HF `WhisperModel(...).encoder` stands in the reference (`FrozenModelWrapper.model`):
    callable on input_features, returns an object with `.last_hidden_state`, exposes `.config.d_model`."""

    def __init__(self, cfg: EncoderConfig, weights: Dict[str, torch.Tensor], hf_config=None, max_batch: int = 32,
                 out_dtype=torch.float32):
        super().__init__()
        self.enc_cfg = cfg
        self.config = hf_config if hf_config is not None else SimpleNamespace(
            d_model=cfg.d_model, num_mel_bins=cfg.n_mels, encoder_layers=cfg.n_layers,
            encoder_attention_heads=cfg.n_heads, encoder_ffn_dim=cfg.ffn_dim, max_source_positions=cfg.n_ctx)
        # frozen parameters kept as buffers-with-parameter-semantics so `.parameters()` / `.to()` behave
        self._names = list(weights.keys())
        for i, k in enumerate(self._names):
            self.register_parameter(f"w{i}", torch.nn.Parameter(weights[k].detach().clone(), requires_grad=False))
        self._max_batch = max_batch
        self._out_dtype = out_dtype
        self._impl = None
        self._impl_device = None

    def state_dict_hf(self) -> Dict[str, torch.Tensor]:
        return {k: getattr(self, f"w{i}") for i, k in enumerate(self._names)}

    def forward(self, input_features, attention_mask=None, **kwargs):
        dev = input_features.device
        if self._impl is None or self._impl_device != dev:
            self._impl = WhisperEncoderB200(self.enc_cfg, self.state_dict_hf(), self._max_batch, device=dev,
                                            out_dtype=self._out_dtype)
            self._impl_device = dev
        return SimpleNamespace(last_hidden_state=self._impl(input_features))
