#!/usr/bin/env python
"""Times al_mel_forward alone (32 clips x 30 s) with CUDA events; prints us/clip and the HBM-roofline fraction."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_llama_b200 import ops, synth
B = int(os.environ.get("B", 32))
x = torch.from_numpy(synth.synth_batch(B)).cuda()
out = torch.empty(B, 128, 3000, device="cuda")
for _ in range(3):
    ops.mel_forward(x, out=out)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(10):
    flush.zero_()                      # evict L2 between timed launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.mel_forward(x, out=out); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
bytes_alg = B * (480000 * 4 + 128 * 3000 * 4)
print(f"{os.environ.get('AUDIOLLM_B200_LIB', 'default')}: mel B={B} {ms*1e3:.1f} us  {ms*1e3/B:.2f} us/clip  {bytes_alg/ms/1e6:.0f} GB/s algorithmic = {bytes_alg/ms/1e6/6550.4:.3f} of HBM")
ref = None
