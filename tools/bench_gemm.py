#!/usr/bin/env python
"""Times the tcgen05 GEMM alone on the encoder's shapes (M = 48000 = 32 clips x 1500) for each epilogue; CUDA events,
L2 flushed between launches. Prints TFLOP/s per (shape, epilogue)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_llama_b200 import ops
from audio_llama_b200.ops import EPI_GELU, EPI_OUT_F32, EPI_REDUCE_ADD, EPI_RESIDUAL

M = int(os.environ.get("M", 48000))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(name, N, K, flags):
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    b = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if flags & EPI_OUT_F32 else torch.bfloat16)
    for _ in range(3):
        ops.gemm_bf16(a, w, b, flags=flags, out=out, resid=out if flags & EPI_RESIDUAL else None)
    ts = []
    for _ in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm_bf16(a, w, b, flags=flags, out=out, resid=out if flags & EPI_RESIDUAL else None); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print(f"{name:34s} N={N:5d} K={K:5d} flags={flags:2d}  {ms*1e3:7.1f} us  {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)


run("qkv  (bf16 out)", 3840, 1280, 0)
run("fc1  (bf16 out, no GELU)", 5120, 1280, 0)
run("fc1  (bf16 out, GELU)", 5120, 1280, EPI_GELU)
run("fc2  (fp32 reduce-add)", 1280, 5120, EPI_OUT_F32 | EPI_REDUCE_ADD)
run("fc2  (fp32 store)", 1280, 5120, EPI_OUT_F32)
run("fc2  (bf16 out)", 1280, 5120, 0)
run("out  (fp32 reduce-add)", 1280, 1280, EPI_OUT_F32 | EPI_REDUCE_ADD)
run("out  (fp32 store)", 1280, 1280, EPI_OUT_F32)
run("out  (fp32, residual read in epilogue)", 1280, 1280, EPI_OUT_F32 | EPI_RESIDUAL)
run("fc2  (fp32, residual read in epilogue)", 1280, 5120, EPI_OUT_F32 | EPI_RESIDUAL)
run("out  (bf16 out)", 1280, 1280, 0)
