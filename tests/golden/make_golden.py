"""Generates tests/golden/*.npz by running the REFERENCE itself in the build container.

Run once, here (needs /root/reference and the installed transformers; neither is read at test
time):   python tests/golden/make_golden.py

What is exercised, unmodified:
  * HF WhisperFeatureExtractor(feature_size=n_mels) exactly as /root/reference/src/inference.py:100-105
    calls it (the M1 parity target) and its mel_filter_bank;
  * torchaudio MelSpectrogram + log exactly as /root/reference/src/dataset.py:125-133 (M2);
  * HF WhisperEncoder (E2) with the seeded weights of audio_llama_b200.synth loaded by state_dict;
  * /root/reference/src/models/projector.py AudioProjector (P1), lora.py LoRALayer + hook (L1),
    allm.py AudioLLM._combine_text_and_audio_embeddings / _extend_attention_mask / forward (S1/S2/F1)
    with load_base_models patched to random-init named shapes (no checkpoints exist offline).
Outputs are small: sub-sampled grids, sums and selected elements, a few hundred KB in total.
"""
import contextlib
import io
import os
import sys
from unittest.mock import Mock, patch

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from audio_llama_b200 import synth  # noqa: E402
from audio_llama_b200.config import WHISPER_TINY_128, EncoderConfig  # noqa: E402
sys.path.insert(0, HERE)
from golden_signals import encoder_input, kat_signals  # noqa: E402


def mel_golden():
    from transformers import WhisperFeatureExtractor
    from transformers.audio_utils import mel_filter_bank
    out = {}
    for n_mels in (128, 80):
        fe = WhisperFeatureExtractor(feature_size=n_mels)
        out[f"fbank{n_mels}"] = np.asarray(fe.mel_filters, dtype=np.float64)
        for name, x in kat_signals().items():
            if n_mels == 80 and name not in ("noise0", "synth0"):
                continue
            f = fe(x, sampling_rate=16000, return_tensors="pt").input_features[0].numpy()
            assert f.shape == (n_mels, 3000) and f.dtype == np.float32
            k = f"{name}_{n_mels}"
            out[k + "_grid"] = f[::8, ::50].copy()
            out[k + "_stats"] = np.array([f.astype(np.float64).sum(), f.min(), f.max()], np.float64)
            out[k + "_col1500"] = f[:, 1500].copy()
            out[k + "_row10"] = f[10, :].copy()
    # batched == per clip
    fe = WhisperFeatureExtractor(feature_size=128)
    sig = kat_signals()
    fb = fe([sig["noise0"], sig["sine440"]], sampling_rate=16000, return_tensors="pt").input_features.numpy()
    out["batched_grid"] = fb[:, ::8, ::50].copy()
    np.savez_compressed(os.path.join(HERE, "mel_whisper.npz"), **out)

    # M2: dataset.py:125-133 verbatim
    import torchaudio
    out2 = {}
    for name in ("zeros", "noise0", "synth0"):
        w = torch.from_numpy(sig[name]).unsqueeze(0)
        mel = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=400, hop_length=160,
                                                   n_mels=128, power=2.0)(w)
        lm = torch.log(mel + 1e-9)[:, :, :3000].numpy()
        out2[name + "_grid"] = lm[0, ::8, ::50].copy()
        out2[name + "_stats"] = np.array([lm.astype(np.float64).sum(), lm.min(), lm.max()], np.float64)
    fbt = torchaudio.functional.melscale_fbanks(201, 0.0, 8000.0, 128, 16000, norm=None, mel_scale="htk")
    out2["fbank_htk128"] = fbt.numpy()
    np.savez_compressed(os.path.join(HERE, "mel_train.npz"), **out2)


def encoder_golden():
    from transformers import WhisperConfig, WhisperModel
    out = {}
    for tag, cfg, seed in (("tiny128", WHISPER_TINY_128, 0),
                           ("small2", EncoderConfig(d_model=256, n_layers=2, n_heads=4, ffn_dim=512, n_mels=80), 3)):
        hf_cfg = WhisperConfig(vocab_size=51865, num_mel_bins=cfg.n_mels, d_model=cfg.d_model,
                               encoder_layers=cfg.n_layers, encoder_attention_heads=cfg.n_heads,
                               encoder_ffn_dim=cfg.ffn_dim, decoder_layers=1, decoder_attention_heads=cfg.n_heads,
                               decoder_ffn_dim=cfg.ffn_dim, max_source_positions=1500)
        enc = WhisperModel(hf_cfg).eval().encoder
        w = synth.init_encoder_weights(cfg, seed=seed, ln_jitter=0.1)
        missing = enc.load_state_dict(w, strict=True)
        mel = torch.from_numpy(encoder_input(cfg, 2))
        with torch.no_grad():
            y = enc(mel).last_hidden_state
        out[tag + "_grid"] = y[:, ::25, ::16].numpy().copy()
        out[tag + "_row7"] = y[:, 7, :].numpy().copy()
        out[tag + "_norm"] = np.array([float(y.double().norm()), float(y.double().sum())])
    np.savez_compressed(os.path.join(HERE, "encoder_hf.npz"), **out)


def encoder_turbo_golden():
    """The named architecture itself (whisper-large-v3-turbo encoder: d=1280, 32 layers, 20 heads, 128 mel) with the
    seeded random-init weights of audio_llama_b200.synth, ONE clip, through the installed HF WhisperEncoder in fp32 on
    the CPU (about a minute). Written to its own file so that the small fixtures need not be regenerated with it."""
    from transformers import WhisperConfig, WhisperModel
    from audio_llama_b200.config import WHISPER_LARGE_V3_TURBO as cfg
    hf_cfg = WhisperConfig(vocab_size=51866, num_mel_bins=cfg.n_mels, d_model=cfg.d_model,
                           encoder_layers=cfg.n_layers, encoder_attention_heads=cfg.n_heads,
                           encoder_ffn_dim=cfg.ffn_dim, decoder_layers=1, decoder_attention_heads=cfg.n_heads,
                           decoder_ffn_dim=cfg.ffn_dim, max_source_positions=1500)
    enc = WhisperModel(hf_cfg).eval().encoder
    enc.load_state_dict(synth.init_encoder_weights(cfg, seed=0, ln_jitter=0.1), strict=True)
    mel = torch.from_numpy(encoder_input(cfg, 1))
    with torch.no_grad():
        y = enc(mel).last_hidden_state
    out = {"turbo_grid": y[:, ::25, ::16].numpy().copy(), "turbo_row7": y[:, 7, :].numpy().copy(),
           "turbo_row1499": y[:, 1499, :].numpy().copy(),
           "turbo_norm": np.array([float(y.double().norm()), float(y.double().sum())])}
    np.savez_compressed(os.path.join(HERE, "encoder_hf_turbo.npz"), **out)


def reference_modules_golden():
    from models.projector import AudioProjector
    from models.lora import LoRALayer, lora_forward_hook
    out = {}
    # P1
    pw = synth.init_projector_weights(384, 256, seed=1, ln_jitter=0.1)
    proj = AudioProjector(384, 256)
    proj.load_state_dict(pw, strict=True)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 40, 384, generator=g)
    with torch.no_grad():
        out["proj_x"] = x.numpy()
        out["proj_y"] = proj(x).numpy()
    # L1
    lin = torch.nn.Linear(96, 160)
    lora = LoRALayer(96, 160, rank=8, alpha=16)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(160, 96, generator=g) * 0.05)
        lin.bias.copy_(torch.randn(160, generator=g) * 0.05)
        lora.lora_A.copy_(torch.randn(8, 96, generator=g) * 0.01)
        lora.lora_B.copy_(torch.randn(160, 8, generator=g) * 0.01)
    xl = torch.randn(3, 17, 96, generator=g)
    with torch.no_grad():
        y = lora_forward_hook(lin, (xl,), lin(xl), lora)
    for k, v in dict(lora_W=lin.weight, lora_b=lin.bias, lora_A=lora.lora_A, lora_B=lora.lora_B,
                     lora_x=xl, lora_y=y).items():
        out[k] = v.detach().numpy().copy()
    out["lora_scaling"] = np.array([lora.scaling])
    np.savez_compressed(os.path.join(HERE, "reference_modules.npz"), **out)


def allm_golden():
    """Config 1: tiny-128-shaped encoder + 2-layer d=256 LLaMA, 4 clips (runs on CPU)."""
    from transformers import LlamaConfig, LlamaForCausalLM, WhisperConfig, WhisperModel
    from transformers import WhisperFeatureExtractor
    import models.allm as allm
    from models.base import FrozenModelWrapper
    cfg = WHISPER_TINY_128
    vocab = 320

    def fake(llama_path, whisper_path):
        torch.manual_seed(0)
        lc = LlamaConfig(vocab_size=vocab, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                         num_attention_heads=4, num_key_value_heads=4, max_position_embeddings=4096)
        wc = WhisperConfig(vocab_size=51865, num_mel_bins=128, d_model=cfg.d_model, encoder_layers=cfg.n_layers,
                           encoder_attention_heads=cfg.n_heads, encoder_ffn_dim=cfg.ffn_dim, decoder_layers=1,
                           decoder_attention_heads=cfg.n_heads, decoder_ffn_dim=cfg.ffn_dim)
        enc = WhisperModel(wc).eval().encoder
        enc.load_state_dict(synth.init_encoder_weights(cfg, seed=0, ln_jitter=0.1))
        return FrozenModelWrapper(LlamaForCausalLM(lc).eval()), FrozenModelWrapper(enc)

    with patch.object(allm, "load_base_models", fake), contextlib.redirect_stdout(io.StringIO()):
        model = allm.AudioLLM("x", "y", lora_rank=8)
        model.projector.load_state_dict(synth.init_projector_weights(384, 256, seed=1, ln_jitter=0.1))
        tok = Mock()
        tok.convert_tokens_to_ids = lambda t: {"<audio>": vocab - 2, "</audio>": vocab - 1}[t]
        model.tokenizer = tok
        B, T = 4, 16
        ids, mask, labels = synth.synth_text(B, T, vocab, seed=7)
        fe = WhisperFeatureExtractor(feature_size=128)
        clips = [synth.synth_clip(i) for i in range(B)]
        feats = fe(clips, sampling_rate=16000, return_tensors="pt").input_features.unsqueeze(1)
        E = model.llama.model.model.embed_tokens.weight.detach()
        with torch.no_grad():
            text_emb = model.llama.model.model.embed_tokens(ids)
            enc_out = model._process_audio_features(feats)
            combined = model._combine_text_and_audio_embeddings(text_emb, feats, ids)
            ext = model._extend_attention_mask(mask, 1500)
            ext_ns = model._extend_attention_mask(mask, 1500, has_special_tokens=False)
            res = model(input_ids=ids, attention_mask=mask, audio_features=feats, labels=labels)
        n_train = sum(p.numel() for p in model.get_trainable_params())
    out = dict(
        embed_table=E.numpy().copy(), ids=ids.numpy(), mask=mask.numpy(), labels=labels.numpy(),
        enc_grid=enc_out[:, ::25, ::16].numpy().copy(),
        combined_shape=np.array(combined.shape), combined_grid=combined[:, ::53, ::8].numpy().copy(),
        combined_rows=combined[:, [0, 1, 2, 750, 1500, 1501, 1502, 1503, 1517], :].numpy().copy(),
        ext_mask=ext.numpy().copy(), ext_mask_dtype=np.array([str(ext.dtype)]),
        ext_mask_ns_shape=np.array(ext_ns.shape),
        loss=np.array([float(res.loss)]), logits_shape=np.array(res.logits.shape),
        n_trainable=np.array([n_train]),
    )
    np.savez_compressed(os.path.join(HERE, "allm_config1.npz"), **out)


def resample_golden():
    """torchaudio.transforms.Resample exactly as inference.py:91-93 / dataset.py:119-123 construct it."""
    import torchaudio
    from golden_signals import resample_input
    out = {}
    for sr in (44100, 48000, 8000, 22050, 32000):
        x = resample_input(sr)
        res = torchaudio.transforms.Resample(orig_freq=sr, new_freq=16000)
        y = res(torch.from_numpy(x)).numpy()
        mono = res(torch.mean(torch.from_numpy(x), dim=0, keepdim=True)).numpy()[0]
        out[f"y_{sr}"] = y
        out[f"mono_{sr}"] = mono
    np.savez_compressed(os.path.join(HERE, "resample.npz"), **out)


def checkpoint_golden():
    """A checkpoint.pt written by the REFERENCE's own save_checkpoint (train.py:102-131) for a tiny model."""
    import argparse
    import shutil
    import tempfile
    from types import SimpleNamespace
    from models.projector import AudioProjector
    from models.lora import LoRALayer
    import importlib.util
    # train.py configures logging / imports wandb at import time; load only save_checkpoint's source
    src = open("/root/reference/src/train.py").read()
    start = src.index("def save_checkpoint(")
    end = src.index("def evaluate(")
    ns = {"os": os, "torch": torch, "logger": SimpleNamespace(info=lambda *a, **k: None)}
    exec(src[start:end], ns)
    torch.manual_seed(3)
    model = SimpleNamespace(projector=AudioProjector(16, 24),
                            lora_layers={"model.layers.0.self_attn.q_proj": LoRALayer(24, 24, rank=4),
                                         "model.layers.0.mlp.down_proj": LoRALayer(48, 24, rank=4)})
    for l in model.lora_layers.values():
        torch.nn.init.normal_(l.lora_A, std=0.02)
    params = list(model.projector.parameters()) + [p for l in model.lora_layers.values() for p in l.parameters()]
    opt = torch.optim.AdamW(params, lr=1e-3)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (1 + s))
    x = torch.randn(5, 16)
    (model.projector(x).sum()).backward()
    opt.step(); sched.step()
    tmp = tempfile.mkdtemp()
    args = argparse.Namespace(output_dir=tmp, learning_rate=1e-3, lora_rank=4)
    ns["save_checkpoint"](model, opt, sched, 7, 1, args, dataset_config={"text_key": "text"})
    shutil.copy(os.path.join(tmp, "checkpoint-7", "checkpoint.pt"), os.path.join(HERE, "reference_checkpoint.pt"))
    with torch.no_grad():
        y = model.projector(x)
    np.savez_compressed(os.path.join(HERE, "reference_checkpoint_io.npz"), x=x.numpy(), y=y.numpy(),
                        lora_A0=model.lora_layers["model.layers.0.self_attn.q_proj"].lora_A.detach().numpy())
    shutil.rmtree(tmp)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    if len(sys.argv) > 1 and sys.argv[1] == "turbo":      # python tests/golden/make_golden.py turbo
        encoder_turbo_golden()
        print("encoder_hf_turbo.npz", os.path.getsize(os.path.join(HERE, "encoder_hf_turbo.npz")))
        sys.exit(0)
    mel_golden()
    encoder_golden()
    reference_modules_golden()
    allm_golden()
    checkpoint_golden()
    resample_golden()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
