#!/usr/bin/env python
"""Times al_attention alone at the bench shape (B=32, T=1500, H=20) with CUDA events; prints TFLOP/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_llama_b200 import ops

B, T, H = int(os.environ.get("B", 32)), 1500, 20
iters = int(os.environ.get("ITERS", 20))
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B, T, 3 * H * 64, generator=g)
qkv[..., :H * 64] *= 0.125 * 3
qkv = qkv.bfloat16().cuda()
for _ in range(3):
    ops.attention(qkv, H)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ops.attention(qkv, H)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
fl = 4.0 * T * T * H * 64 * B
print(f"{os.environ.get('AUDIOLLM_B200_LIB', 'default')}: attention B={B} {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
