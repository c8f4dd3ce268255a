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
# the form the encoder uses: log2(e) folded into q (AL_ATT_Q_LOG2); same probabilities
qkv2 = qkv.clone()
qkv2[..., :H * 64] = (qkv[..., :H * 64].float() * 1.4426950408889634).bfloat16()
for _ in range(3):
    ops.attention(qkv2, H, True)
torch.cuda.synchronize()
e0.record()
for _ in range(iters):
    ops.attention(qkv2, H, True)
e1.record()
torch.cuda.synchronize()
ms2 = e0.elapsed_time(e1) / iters
print(f"  q in log2 units (encoder form): {ms2:.3f} ms  {fl / ms2 / 1e9:.1f} TFLOP/s")

if os.environ.get("COMPARE", "1") != "0":          # library reference point on the same tensors (not on the product path)
    import torch.nn.functional as F
    d = H * 64
    q, k, v = (qkv[..., i * d:(i + 1) * d].reshape(B, T, H, 64).transpose(1, 2) for i in range(3))
    for backend in ("CUDNN_ATTENTION", "FLASH_ATTENTION"):
        try:
            from torch.nn.attention import SDPBackend, sdpa_kernel
            with sdpa_kernel(getattr(SDPBackend, backend)):
                for _ in range(3):
                    F.scaled_dot_product_attention(q, k, v, scale=1.0)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(iters):
                    F.scaled_dot_product_attention(q, k, v, scale=1.0)
                e1.record()
                torch.cuda.synchronize()
            ms_l = e0.elapsed_time(e1) / iters
            print(f"torch SDPA {backend}: {ms_l:.3f} ms  {fl / ms_l / 1e9:.1f} TFLOP/s  (strided q/k/v views of the same qkv)")
        except Exception as e:                     # noqa: BLE001
            print(f"torch SDPA {backend}: unavailable ({type(e).__name__}: {str(e)[:80]})")
