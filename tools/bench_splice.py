#!/usr/bin/env python
"""Times al_splice alone, COLD: rotating (ids, output) sets whose footprint exceeds the L2 several times over, at the
bench shape (B=32, T=512, d=2048, bf16, vocab 128258) -- the same measurement bench.py reports as
kernels["splice (1 launch, cold)"]."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from audio_llama_b200 import ops, synth

B, T, D, V, A = int(os.environ.get("B", 32)), 512, int(os.environ.get("D", 2048)), 128258, 1500
SETS = 6
dev = torch.device("cuda")
table = (torch.randn(V, D, generator=torch.Generator().manual_seed(2)) * 0.02).to(torch.bfloat16).to(dev)
S = A + 2 + T
ids = [synth.synth_text(B, T, V, seed=1000 + 17 * k)[0].to(dev) for k in range(SETS)]
_, mask, labels = synth.synth_text(B, T, V)
mask, labels = mask.to(dev), labels.to(dev)
out = [torch.empty(B, S, D, dtype=torch.bfloat16, device=dev) for _ in range(SETS)]
mo, lo = torch.empty(B, S, dtype=torch.float32, device=dev), torch.empty(B, S, dtype=torch.int64, device=dev)


def call(k):
    ops.splice(table, ids[k], mask, labels, A, V - 2, V - 1, audio_rows=None, out=out[k], mask_out=mo, labels_out=lo, check_ids=False)


for k in range(SETS):
    call(k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10 * SETS
e0.record()
for i in range(n):
    call(i % SETS)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
nbytes = B * (2 * (T + 2) * D * 2 + S * 12 + T * 24)
print(f"{os.environ.get('AUDIOLLM_B200_LIB', 'default')}: splice cold B={B} d={D}: {ms * 1e3:.1f} us  {nbytes / ms / 1e6:.0f} GB/s algorithmic "
      f"= {nbytes / ms / 1e6 / 6550.4:.3f} of HBM (6550 GB/s)")
# warm repeat of one set, for contrast
e0.record()
for i in range(20):
    call(0)
e1.record()
torch.cuda.synchronize()
print(f"  L2-warm repeat: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
