#!/usr/bin/env python
"""GPU time by kernel of one README training step (configs[2]) on one GPU: torch.profiler over TrainStep.step()."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from audio_llama_b200 import train_step
from audio_llama_b200.config import WHISPER_LARGE_V3_TURBO

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
model = train_step.build_model(os.environ.get("LLAMA", "3b"), WHISPER_LARGE_V3_TURBO, 8, dev)
ts = train_step.TrainStep(model, WHISPER_LARGE_V3_TURBO, 8, dev)
for _ in range(2):
    print(ts.step())
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    r = ts.step()
print(r)
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"# GPU time of the profiled step: {tot / 1e3:.1f} ms")
for e in rows[:32]:
    print(f"# {e.device_time_total / 1e3:9.2f} ms {100 * e.device_time_total / tot:5.1f}% x{e.count:<5d} {e.key[:120]}")
