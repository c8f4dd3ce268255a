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

print('# most-launched device activities:')
for e in sorted(prof.key_averages(), key=lambda e: -e.count)[:14]:
    print(f"#   x{e.count:<5d} {e.device_time_total / 1e3:7.2f} ms  {e.key[:110]}")

# --- where the GPU idles inside the step: gaps between consecutive device activities (kernels / memcpys / memsets)
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
evs.sort(key=lambda e: e.time_range.start)
gaps = []
last_end, last_name = None, None
for e in evs:
    st, en = e.time_range.start, e.time_range.end
    if last_end is not None and st > last_end:
        gaps.append((st - last_end, last_name, e.name))
    if last_end is None or en > last_end:
        last_end, last_name = en, e.name
tot_gap = sum(g for g, _, _ in gaps)
print(f"# idle between device activities: {tot_gap / 1e3:.2f} ms in {len(gaps)} gaps "
      f"({sum(1 for g, _, _ in gaps if g > 20)} gaps > 20 us = {sum(g for g, _, _ in gaps if g > 20) / 1e3:.2f} ms)")
by = {}
for g, a, b in gaps:
    k = (a[:40], b[:40])
    by[k] = (by.get(k, (0, 0))[0] + g, by.get(k, (0, 0))[1] + 1)
for (a, b), (g, n) in sorted(by.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"#   {g / 1e3:7.2f} ms x{n:<4d} after {a:40s} before {b}")
