"""The README training step (BASELINE.json configs[2]) on the B200 path: AudioLLM forward (mel -> encoder -> projector
-> splice -> LLaMA with LoRA) + backward + the data-parallel exchange + AdamW, one process per GPU.

Mirrors the reference's hot loop (/root/reference/src/train.py:261-300: forward, loss.backward(), clip_grad_norm_,
optimizer.step(), zero_grad) with the one thing the reference does not have: the all-reduce of the trainable set
(projector + LoRA, 95 726 720 parameters at the README shapes) over NCCL. The exchange is OVERLAPPED with the backward
pass (parallel.FlatGradBucket.arm_overlap: chunked, launched from gradient hooks). LLaMA itself is stock HF with
random-init weights of the named shape (no checkpoints offline); tools/train_step_dp.py and bench.py's `config3`
record both run this module."""
from __future__ import annotations

import time
from typing import Dict, Optional
from unittest.mock import Mock, patch

import torch
import torch.distributed as dist

from . import parallel, synth
from .config import WHISPER_LARGE_V3_TURBO, EncoderConfig

LLAMAS = {
    "3b": dict(hidden_size=3072, intermediate_size=8192, num_hidden_layers=28, num_attention_heads=24, num_key_value_heads=8, vocab_size=128258),
    "1b": dict(hidden_size=2048, intermediate_size=8192, num_hidden_layers=16, num_attention_heads=32, num_key_value_heads=8, vocab_size=128258),
    "tiny": dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=4, vocab_size=320),
}


def build_model(llama: str, ecfg: EncoderConfig, batch: int, dev, lora_rank: int = 64, fused_lora: bool = True,
                native_llama: bool = True, encoder_weights: Optional[Dict[str, torch.Tensor]] = None):
    """AudioLLM with random-init weights of the named shapes on `dev` (bf16 LLaMA, fp32 trainable projector)."""
    from .models import base as B
    from .models.allm import AudioLLM
    lcfg = LLAMAS[llama]

    def fake_load(llama_path, whisper_path):
        from transformers import LlamaConfig, LlamaForCausalLM
        from .encoder import WhisperEncoderModule
        torch.manual_seed(0)
        with torch.device(dev):
            lm = LlamaForCausalLM(LlamaConfig(max_position_embeddings=4096, **lcfg)).to(torch.bfloat16)
        ew = encoder_weights if encoder_weights is not None else synth.init_encoder_weights(ecfg, seed=0)
        enc = WhisperEncoderModule(ecfg, ew, max_batch=batch, out_dtype=torch.bfloat16)
        return B.FrozenModelWrapper(lm), B.FrozenModelWrapper(enc)

    with patch.object(B, "load_base_models", fake_load):
        model = AudioLLM("x", "y", lora_rank=lora_rank)
    vocab = lcfg["vocab_size"]
    tok = Mock()
    tok.convert_tokens_to_ids = lambda t: {"<audio>": vocab - 2, "</audio>": vocab - 1}[t]
    model.tokenizer = tok
    model = model.to(dev)
    if fused_lora:
        model.enable_fused_lora()
    if native_llama:
        model.enable_native_llama_ops()
    model.projector.to(torch.float32)
    for l in model.lora_layers.values():          # the reference's init makes the LoRA update identically zero
        torch.nn.init.normal_(l.lora_A, std=0.01)
    return model


class TrainStep:
    """One rank's training step. step() returns device-timed milliseconds: total, and the part of the exchange that
    the backward pass did not hide (the wait inside finish_overlap)."""

    def __init__(self, model, ecfg: EncoderConfig, batch: int, dev, rank: int = 0, world: int = 1, t_txt: int = 512,
                 overlap: bool = True, n_chunks: int = 4, lr: float = 1e-4, graph: bool = False):
        from .features import LogMelExtractor
        self.model, self.dev, self.rank, self.world, self.batch = model, dev, rank, world, batch
        vocab = model.llama.model.model.embed_tokens.weight.shape[0]
        self.params = model.get_trainable_params()
        self.bucket = parallel.FlatGradBucket(self.params)
        self.overlap = overlap and world > 1
        self.n_chunks = 1
        if self.overlap:
            self.bucket.arm_overlap(n_chunks)
            self.n_chunks = len(self.bucket._ov["chunks"])
        self.opt = torch.optim.AdamW(self.params, lr=lr)
        self.ids, self.mask, self.labels = (t.to(dev) for t in synth.synth_text(batch, t_txt, vocab, seed=7 + rank))
        self.fe = LogMelExtractor(ecfg.n_mels, device=dev)
        self.clips = [synth.synth_clip(rank * batch + i) for i in range(batch)]
        self.loss = None
        # graph=True (single GPU): zero + forward + backward are captured ONCE as a CUDA graph and replayed (the ~3000
        # launches of the step then follow each other without the ~2 us launch gaps); the feature extraction, the gradient
        # clipping and AdamW stay eager. The two host reads of the forward (attention_plan, the splice id check) are
        # deferred: the mask is right padding by construction here and the id flag is read after the step.
        self.graph = None
        self.want_graph = bool(graph) and world == 1
        self.graph_error = None
        self.eager_steps = 0

    def _fwd_bwd(self, feats):
        self.bucket.zero()
        out = self.model(input_ids=self.ids, attention_mask=self.mask, audio_features=feats, labels=self.labels)
        out.loss.backward()
        return out.loss.detach()

    def _capture(self, feats):
        """Capture zero + forward + backward on the static inputs. Any failure leaves the eager path in place."""
        from . import llama_native, ops
        from .models import lora as lora_mod
        self.static_feats = feats.clone()
        g = torch.cuda.CUDAGraph()
        try:
            ops.DEFER_ID_CHECKS = True
            lora_mod.REPACK_ALWAYS = True                # the pack launches must be part of the graph
            with llama_native.static_attention_plan():
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):               # one more warm-up on the side stream the capture will use
                    self._fwd_bwd(self.static_feats)
                torch.cuda.current_stream().wait_stream(s)
                torch.cuda.synchronize()
                with torch.cuda.graph(g, stream=s):
                    self.static_loss = self._fwd_bwd(self.static_feats)
            self.graph = g
        except Exception as e:                           # noqa: BLE001
            self.graph_error = f"{type(e).__name__}: {e}"[:300]
            self.graph = None
            self.want_graph = False
            torch.cuda.synchronize()
        finally:
            ops.DEFER_ID_CHECKS = False
            lora_mod.REPACK_ALWAYS = False

    def step(self) -> Dict[str, float]:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        feats = self.fe(self.clips, sampling_rate=16000).input_features.unsqueeze(1)
        if self.want_graph and self.graph is None and self.eager_steps >= 2:
            self._capture(feats)
        if self.graph is not None:
            self.static_feats.copy_(feats)
            self.graph.replay()
            loss = self.static_loss
        else:
            loss = self._fwd_bwd(feats)
            self.eager_steps += 1
        ev[1].record()
        if self.overlap:
            self.bucket.finish_overlap()
            self.chunks_in_backward = self.bucket._ov.get("launched_in_backward")
        else:
            self.bucket.allreduce_mean()
        ev[2].record()
        torch.nn.utils.clip_grad_norm_(self.params, 1.0)     # after the exchange: no second collective (train.py:294)
        self.opt.step()
        ev[3].record()
        torch.cuda.synchronize()
        self.loss = float(loss)
        if self.graph is not None:
            from . import ops
            ops.raise_if_bad_ids(self.dev)               # the id check the captured forward could not do on the host
        return {"step_ms": ev[0].elapsed_time(ev[3]), "fwd_bwd_ms": ev[0].elapsed_time(ev[1]),
                "exchange_exposed_ms": ev[1].elapsed_time(ev[2]), "optimizer_ms": ev[2].elapsed_time(ev[3])}

    def allreduce_alone_ms(self, reps: int = 3) -> Optional[float]:
        """The same bucket reduced with nothing else running (the bus-bandwidth figure)."""
        if self.world <= 1:
            return None
        self.bucket.disarm_overlap()
        best = None
        for _ in range(reps):
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self.bucket.allreduce_mean()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        return best


def run_config3(dev, rank: int, world: int, llama: str = "3b", batch: int = 8, steps: int = 3, warmup: int = 1,
                overlap: bool = True, encoder_weights=None, ecfg: EncoderConfig = WHISPER_LARGE_V3_TURBO,
                graph: bool = False) -> Optional[dict]:
    """The config-3 record: max-over-ranks step time of `steps` training steps, the exposed and stand-alone exchange
    times and the bus bandwidth. Every rank calls it; rank 0 gets the dict."""
    model = build_model(llama, ecfg, batch, dev, encoder_weights=encoder_weights)
    ts = TrainStep(model, ecfg, batch, dev, rank, world, overlap=overlap, graph=graph)
    for _ in range(warmup + (2 if ts.want_graph else 0)):     # (graph mode: two eager steps, then the capture step)
        ts.step()
    recs = []
    for _ in range(steps):
        if world > 1:
            dist.barrier()
        recs.append(ts.step())
    keys = list(recs[0].keys())
    t = torch.tensor([[r[k] for k in keys] for r in recs], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    alone = ts.allreduce_alone_ms()
    if alone is not None:
        a = torch.tensor([alone], device=dev, dtype=torch.float64)
        dist.all_reduce(a, op=dist.ReduceOp.MAX)
        alone = float(a.item())
    # median over the timed steps of the max over ranks (the first steps after model construction still grow the
    # allocator's pools and autotune nothing here, but one slow outlier should not be the reported step)
    mean = {k: float(t[:, i].median().item()) for i, k in enumerate(keys)}
    all_steps = [float(x) for x in t[:, keys.index("step_ms")].tolist()]
    nbytes = ts.bucket.numel * 4
    rec = {
        "workload": f"configs[2]: README training step -- whisper-large-v3-turbo encoder + Llama-3.2-{llama.upper()} shape, LoRA r=64 on "
                    f"q/k/v/gate/up/down, batch {batch} x 30 s clips per GPU, bf16, T_txt 512 (S = 2014), data-parallel dp{world}",
        "steps": steps, "warmup": warmup, "step_ms": mean["step_ms"], "step_ms_each": all_steps, "fwd_bwd_ms": mean["fwd_bwd_ms"],
        "optimizer_ms": mean["optimizer_ms"], "audio_s_per_s": world * batch * 30.0 / (mean["step_ms"] / 1e3),
        "trainable_params": ts.bucket.numel, "loss": ts.loss,
        "cuda_graph": {"requested": bool(graph), "replayed": ts.graph is not None, "error": ts.graph_error},
        "allreduce": {"bytes": nbytes, "overlapped_with_backward": bool(ts.overlap),
                      "chunks": ts.n_chunks,
                      "chunks_launched_during_backward": getattr(ts, "chunks_in_backward", None),
                      "exposed_ms": mean["exchange_exposed_ms"], "alone_ms": alone,
                      "bus_gbs": (2 * (world - 1) / world * nbytes / 1e9 / (alone / 1e3)) if alone else None},
        "llama": "HF LlamaForCausalLM module tree, random init; every heavy op on this repo's kernels: fused frozen+LoRA GEMMs, "
                 "causal GQA attention forward + backward (key padding by kv_len), RMSNorm / SwiGLU / RoPE / lm_head+CE",
    }
    del ts, model
    torch.cuda.empty_cache()
    return rec if rank == 0 else None
