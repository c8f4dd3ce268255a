#!/usr/bin/env python
"""Times al_layernorm at the encoder's shape (48000 x 1280 fp32 -> bf16; 369 MB per call, larger than L2): the
persistent packed-arithmetic kernel, or the generic one with AUDIOLLM_B200_LN=generic."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from audio_llama_b200 import ops

rows, d = int(os.environ.get("ROWS", 48000)), int(os.environ.get("D", 1280))
x = torch.randn(rows, d, device="cuda")
out = torch.empty(rows, d, dtype=torch.bfloat16, device="cuda")
g, b = torch.randn(d, device="cuda"), torch.randn(d, device="cuda")
name = "generic kernel" if os.environ.get("AUDIOLLM_B200_LN") == "generic" else "persistent rows kernel"
if True:
    for _ in range(3):
        ops.layernorm(x, g, b, out=out, rows_per_group=rows)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        ops.layernorm(x, g, b, out=out, rows_per_group=rows)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 100
    nbytes = rows * d * 6
    print(f"{name}: {ms * 1e3:.1f} us  {nbytes / ms / 1e6:.0f} GB/s = {nbytes / ms / 1e6 / 6550.4:.3f} of HBM")
