#!/usr/bin/env python
"""Config-3 slice: one data-parallel TRAINING step of AudioLLM through the B200 path, one process per GPU
(torchrun), NCCL all-reduce of the projector + LoRA gradients through one flat bucket.

    torchrun --nproc-per-node N tools/train_step_dp.py [--llama 3b|1b|tiny] [--batch 1] [--steps 3]

LLaMA itself is stock HF (out of scope); what this exercises is the path + autograd glue + parallel.FlatGradBucket
over NVLink. Rank 0 prints one JSON line with the step time, the all-reduce time and the bucket size."""
import argparse
import json
import os
import sys
import time
from unittest.mock import Mock, patch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from audio_llama_b200 import parallel, synth
from audio_llama_b200.config import WHISPER_LARGE_V3_TURBO, WHISPER_TINY_128
from audio_llama_b200.features import LogMelExtractor
from audio_llama_b200.models import base as B
from audio_llama_b200.models.allm import AudioLLM

LLAMAS = {
    "3b": dict(hidden_size=3072, intermediate_size=8192, num_hidden_layers=28, num_attention_heads=24, num_key_value_heads=8, vocab_size=128258),
    "1b": dict(hidden_size=2048, intermediate_size=8192, num_hidden_layers=16, num_attention_heads=32, num_key_value_heads=8, vocab_size=128258),
    "tiny": dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=4, vocab_size=320),
}


def main():
    sys.stdout.flush()
    real_stdout = os.dup(1)          # stdout carries only the JSON line (NCCL prints its banner to fd 1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--llama", default="1b")
    ap.add_argument("--encoder", default="turbo")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--rank_lora", type=int, default=64)
    ap.add_argument("--fused-lora", action="store_true", help="AudioLLM.enable_fused_lora(): fused frozen+LoRA GEMM")
    ap.add_argument("--native-llama", action="store_true", help="AudioLLM.enable_native_llama_ops(): RMSNorm / SwiGLU / RoPE / lm_head+CE kernels")
    ap.add_argument("--profile", action="store_true", help="torch.profiler over the last step: GPU time by kernel (rank 0)")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ecfg = WHISPER_LARGE_V3_TURBO if args.encoder == "turbo" else WHISPER_TINY_128
    lcfg = LLAMAS[args.llama]

    def fake_load(llama_path, whisper_path):
        from transformers import LlamaConfig, LlamaForCausalLM
        from audio_llama_b200.encoder import WhisperEncoderModule
        torch.manual_seed(0)
        with torch.device(dev):
            llama = LlamaForCausalLM(LlamaConfig(max_position_embeddings=4096, **lcfg)).to(torch.bfloat16)
        enc = WhisperEncoderModule(ecfg, synth.init_encoder_weights(ecfg, seed=0), max_batch=args.batch, out_dtype=torch.bfloat16)
        return B.FrozenModelWrapper(llama), B.FrozenModelWrapper(enc)

    with patch.object(B, "load_base_models", fake_load):
        model = AudioLLM("x", "y", lora_rank=args.rank_lora)
    vocab = lcfg["vocab_size"]
    tok = Mock()
    tok.convert_tokens_to_ids = lambda t: {"<audio>": vocab - 2, "</audio>": vocab - 1}[t]
    model.tokenizer = tok
    model = model.to(dev)
    if args.fused_lora:
        model.enable_fused_lora()
    if args.native_llama:
        model.enable_native_llama_ops()
    model.projector.to(torch.float32)
    for l in model.lora_layers.values():
        torch.nn.init.normal_(l.lora_A, std=0.01)
    params = model.get_trainable_params()
    bucket = parallel.FlatGradBucket(params)
    opt = torch.optim.AdamW(params, lr=1e-4)

    T = 512
    ids, mask, labels = (t.to(dev) for t in synth.synth_text(args.batch, T, vocab, seed=7 + rank))
    fe = LogMelExtractor(ecfg.n_mels, device=dev)
    clips = [synth.synth_clip(rank * args.batch + i) for i in range(args.batch)]
    t_step, t_ar = [], []
    prof = None
    for step in range(args.steps):
        if args.profile and rank == 0 and step == args.steps - 1:
            prof = torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA])
            prof.__enter__()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        feats = fe(clips, sampling_rate=16000).input_features.unsqueeze(1)
        bucket.zero()
        out = model(input_ids=ids, attention_mask=mask, audio_features=feats, labels=labels)
        out.loss.backward()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        bucket.allreduce_mean()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        t_step.append(t3 - t0)
        t_ar.append(t2 - t1)
    if prof is not None:
        prof.__exit__(None, None, None)
        rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
        tot = sum(e.device_time_total for e in rows)
        print(f"# GPU time of the profiled step: {tot / 1e3:.1f} ms", file=sys.stderr)
        for e in rows[:40]:
            print(f"# {e.device_time_total / 1e3:9.2f} ms {100 * e.device_time_total / tot:5.1f}% x{e.count:<5d} {e.key[:110]}", file=sys.stderr)
    if rank == 0:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps({"world": world, "llama": args.llama, "encoder": args.encoder, "batch_per_gpu": args.batch, "fused_lora": bool(args.fused_lora), "native_llama": bool(args.native_llama),
                          "trainable_params": bucket.numel, "bucket_mb": bucket.numel * 4 / 1e6,
                          "step_s": t_step, "allreduce_s": t_ar, "loss": float(out.loss.detach()),
                          "allreduce_bus_gbs": (2 * (world - 1) / world * bucket.numel * 4 / 1e9 / min(t_ar[1:] or t_ar)) if world > 1 else None}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
