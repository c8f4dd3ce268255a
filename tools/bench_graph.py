#!/usr/bin/env python
"""How much of the step is launch overhead? Runs the resident step of bench.py (32 clips, configs[1]) eagerly and as a
captured CUDA graph (torch.cuda.CUDAGraph around AudioConditioner.__call__) and prints both times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_llama_b200 import synth
from audio_llama_b200.config import WHISPER_LARGE_V3_TURBO as cfg
from audio_llama_b200.pipeline import AudioConditioner

B, T, D, V = 32, 512, 2048, 128258
dev = torch.device("cuda", 0)
ew = synth.init_encoder_weights(cfg, seed=0)
pw = synth.init_projector_weights(cfg.d_model, D, seed=1)
table = (torch.randn(V, D, generator=torch.Generator().manual_seed(2)) * 0.02).to(torch.bfloat16).to(dev)
cond = AudioConditioner(cfg, ew, pw, table, V - 2, V - 1, max_batch=B, device=dev)
wave = torch.from_numpy(synth.synth_batch(B)).to(dev)
ids, mask, labels = (t.to(dev) for t in synth.synth_text(B, T, V))
emb = torch.empty(B, 1502 + T, D, dtype=torch.bfloat16, device=dev)


def step():
    return cond(wave, ids, mask, labels, out=emb)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


t_eager = timed(step)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        step()
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    step()
t_graph = timed(g.replay)
t_eager2 = timed(step)
print(f"eager {t_eager:.3f} ms  graph {t_graph:.3f} ms  eager again {t_eager2:.3f} ms  ({(t_eager2 / t_graph - 1) * 100:.2f} % launch overhead)")
