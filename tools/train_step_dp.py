#!/usr/bin/env python
"""Config-3 slice: data-parallel TRAINING steps of AudioLLM through the B200 path, one process per GPU (torchrun), the
projector + LoRA gradients exchanged through one flat bucket whose chunked NCCL all-reduce overlaps the backward pass.

    torchrun --nproc-per-node N tools/train_step_dp.py [--llama 3b|1b|tiny] [--batch 8] [--steps 3] [--no-overlap]

The step itself lives in audio_llama_b200/train_step.py (bench.py's `config3` record runs the same code). Rank 0
prints one JSON line."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from audio_llama_b200 import train_step
from audio_llama_b200.config import WHISPER_LARGE_V3_TURBO, WHISPER_TINY_128


def main():
    sys.stdout.flush()
    real_stdout = os.dup(1)          # stdout carries only the JSON line (NCCL prints its banner to fd 1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--llama", default="3b")
    ap.add_argument("--encoder", default="turbo")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--no-overlap", action="store_true", help="one all-reduce after backward instead of the overlapped chunks")
    ap.add_argument("--graph", action="store_true", help="single GPU: replay zero + forward + backward as one CUDA graph")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        from audio_llama_b200 import parallel
        parallel.init_nccl(dev)
    ecfg = WHISPER_LARGE_V3_TURBO if args.encoder == "turbo" else WHISPER_TINY_128
    rec = train_step.run_config3(dev, rank, world, llama=args.llama, batch=args.batch, steps=args.steps,
                                 overlap=not args.no_overlap, ecfg=ecfg, graph=args.graph)
    if rank == 0:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
