"""Checkpoint I/O in the reference's own format (§8f row 4).

`save_checkpoint` writes exactly what /root/reference/src/train.py:102-131 writes —
`{model: {projector: sd, lora_layers: {name: sd}}, optimizer, scheduler, step, epoch, args, dataset_config}` to
`<output_dir>/checkpoint-<step>/checkpoint.pt` (or `final_checkpoint/`) — and `load_checkpoint` accepts both
layouts the reference's loader accepts (/root/reference/src/inference.py:51-68: the full format and the flat
`{projector, lora_layers}` one). The reference has no resume path; `load_checkpoint(..., optimizer, scheduler)`
adds it (restores optimizer / scheduler state and returns step and epoch). Frozen base weights are never saved,
as in the reference.
"""
from __future__ import annotations

import os
from typing import Optional

import torch


def save_checkpoint(model, optimizer, scheduler, step: int, epoch: int, args, dataset_config=None, final: bool = False) -> str:
    output_dir = args.output_dir if hasattr(args, "output_dir") else args["output_dir"]
    path = os.path.join(output_dir, "final_checkpoint" if final else f"checkpoint-{step}")
    os.makedirs(path, exist_ok=True)
    checkpoint = {
        "model": {
            "projector": model.projector.state_dict(),
            "lora_layers": {name: layer.state_dict() for name, layer in model.lora_layers.items()},
        },
        "optimizer": optimizer.state_dict() if optimizer is not None else None,
        "scheduler": scheduler.state_dict() if scheduler else None,
        "step": step,
        "epoch": epoch,
        "args": vars(args) if hasattr(args, "__dict__") else dict(args),
        "dataset_config": dataset_config,
    }
    file = os.path.join(path, "checkpoint.pt")
    torch.save(checkpoint, file)
    return file


def load_checkpoint(model, checkpoint_path: str, optimizer=None, scheduler=None, map_location="cpu", strict_lora: bool = False):
    """Loads projector + LoRA weights into `model`; with optimizer / scheduler also resumes them.
    Returns {'step', 'epoch', 'args', 'dataset_config'} (None where the file has none)."""
    ckpt = torch.load(checkpoint_path, map_location=map_location, weights_only=False)
    body = ckpt["model"] if isinstance(ckpt, dict) and "model" in ckpt else ckpt
    model.projector.load_state_dict(body["projector"])
    missing = []
    for name, sd in body["lora_layers"].items():
        if name in model.lora_layers:                    # the reference silently skips unknown names (inference.py:58-60)
            model.lora_layers[name].load_state_dict(sd)
        else:
            missing.append(name)
    if strict_lora and missing:
        raise KeyError(f"checkpoint has LoRA layers the model lacks: {missing[:4]}...")
    if optimizer is not None and isinstance(ckpt, dict) and ckpt.get("optimizer") is not None:
        optimizer.load_state_dict(ckpt["optimizer"])
    if scheduler is not None and isinstance(ckpt, dict) and ckpt.get("scheduler") is not None:
        scheduler.load_state_dict(ckpt["scheduler"])
    get = (lambda k: ckpt.get(k)) if isinstance(ckpt, dict) else (lambda k: None)
    return {"step": get("step"), "epoch": get("epoch"), "args": get("args"), "dataset_config": get("dataset_config")}
