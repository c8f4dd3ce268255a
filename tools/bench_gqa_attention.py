#!/usr/bin/env python
"""Times the native causal GQA attention (forward, backward) at the README training shape (8 x 2014 tokens, 24 query /
8 kv heads, head_dim 128) against torch's scaled_dot_product_attention (cuDNN / flash through SDPA) on the same tensors."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from audio_llama_b200 import llama_native as LN

B, S, Hq, Hkv, D = int(os.environ.get("B", 8)), int(os.environ.get("S", 2014)), 24, 8, 128
g = torch.Generator().manual_seed(0)
q = torch.randn(B, S, Hq, D, generator=g).bfloat16().cuda().requires_grad_(True)
k = torch.randn(B, S, Hkv, D, generator=g).bfloat16().cuda().requires_grad_(True)
v = torch.randn(B, S, Hkv, D, generator=g).bfloat16().cuda().requires_grad_(True)
do = torch.randn(B, S, Hq, D, generator=g).bfloat16().cuda()
flops_fwd = 4.0 * S * S * D * Hq * B / 2          # causal: half of the score tiles
flops_bwd = 2.5 * flops_fwd


def timeit(fn, n=10):
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


def ours_fwd():
    return LN.gqa_attention(q, k, v, None, D ** -0.5)


def ours_fb():
    q.grad = k.grad = v.grad = None
    LN.gqa_attention(q, k, v, None, D ** -0.5).backward(do)


def sdpa(qq, kk, vv):
    return F.scaled_dot_product_attention(qq.transpose(1, 2), kk.transpose(1, 2), vv.transpose(1, 2), is_causal=True,
                                          enable_gqa=True).transpose(1, 2)


def ref_fwd():
    return sdpa(q, k, v)


def ref_fb():
    q.grad = k.grad = v.grad = None
    sdpa(q, k, v).backward(do)


with torch.no_grad():
    t_of = timeit(ours_fwd)
    t_rf = timeit(ref_fwd)
t_ob = timeit(ours_fb)
t_rb = timeit(ref_fb)
print(f"native  fwd {t_of:.3f} ms ({flops_fwd / t_of / 1e9:.0f} TFLOP/s)   fwd+bwd {t_ob:.3f} ms (bwd alone ~{flops_bwd / max(t_ob - t_of, 1e-6) / 1e9:.0f} TFLOP/s)")
print(f"SDPA    fwd {t_rf:.3f} ms ({flops_fwd / t_rf / 1e9:.0f} TFLOP/s)   fwd+bwd {t_rb:.3f} ms (bwd alone ~{flops_bwd / max(t_rb - t_rf, 1e-6) / 1e9:.0f} TFLOP/s)")

if os.environ.get("KERNELS"):                      # per-kernel device times of the native path
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            ours_fb()
        torch.cuda.synchronize()
    for ev in prof.key_averages():
        if "gqa_" in ev.key:
            print(f"  {ev.key[:48]:48s} {ev.device_time_total / ev.count:9.1f} us x {ev.count}")

if os.environ.get("CYCLES"):                       # a -DGQ_CYCLES build (tools/build_gqa_cycles.sh): cycle accounts of one compute warp
    import ctypes
    lib = ctypes.CDLL(os.environ["AUDIOLLM_B200_LIB"])
    buf = (ctypes.c_ulonglong * 64)()
    lib.al_debug_gqa_cycles(buf, 1)
    ours_fb()
    lib.al_debug_gqa_cycles(buf, 0)
    v = list(buf)
    for base, nm in ((0, "forward"), (16, "dQ"), (32, "dK/dV")):
        steps, in_loop, between, items = v[base:base + 4]
        if steps:
            print(f"  {nm:8s} {items} items, {steps} kv-tile steps: {in_loop / steps:7.0f} cycles per step inside the tile loops, "
                  f"{between / max(items, 1):7.0f} cycles per item between them")
