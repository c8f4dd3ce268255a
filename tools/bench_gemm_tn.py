#!/usr/bin/env python
"""Times the weight-gradient GEMM form (al_gemm_tn_accumulate: out[M][N] += A_src^T W_src, contraction over the token rows)
on the LoRA shapes of the README training step (rank 64, 8 x 2014 rows) against the bytes it has to read."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_llama_b200._lib import check, lib, ptr, stream_ptr

K, R = 8 * 2014, 64
g = torch.Generator(device="cuda").manual_seed(0)


def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, (M, N) in {"dA in=3072 (U^T x)": (R, 3072), "dA in=8192": (R, 8192), "dB out=1024 (dy^T T)": (1024, R),
                     "dB out=3072": (3072, R), "dB out=8192": (8192, R)}.items():
    a = torch.randn(K, M, device="cuda", generator=g).bfloat16()
    w = torch.randn(K, N, device="cuda", generator=g).bfloat16()
    out = torch.zeros(M, N, device="cuda", dtype=torch.float32)
    ms = timed(lambda: check(lib().al_gemm_tn_accumulate(ptr(a), M, M, ptr(w), N, N, K, ptr(out), N, stream_ptr()), "tn"))
    mb = (a.numel() + w.numel()) * 2 / 1e6
    print(f"{name:24s} {ms * 1e3:7.1f} us   reads {mb:6.1f} MB -> {mb / ms / 1e3:5.2f} TB/s")
