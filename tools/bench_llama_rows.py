#!/usr/bin/env python
"""Times the LLaMA-side row kernels (llama_rows.cu) and the MN-major weight-gradient GEMM alone at the config-3 shapes
(8 x 2014 tokens, Llama-3.2-3B: d 3072, ffn 8192, 24 / 8 heads of 128, vocab 128 258) with CUDA events; prints the
algorithmic GB/s against the measured HBM copy bandwidth."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_llama_b200 import llama_native as LN
from audio_llama_b200._lib import check, lib, ptr, stream_ptr

M, D, F, V, HBM = 8 * 2014, 3072, 8192, 128258, 6550.4
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, n=10):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()                                   # evict L2 between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s: torch.randn(*s, device="cuda", generator=g).bfloat16()
out = {}


def rec(name, ms, nbytes):
    out[name] = {"ms": ms, "gbs": nbytes / ms / 1e6, "frac_of_hbm": nbytes / ms / 1e6 / HBM}
    print(f"{name:28s} {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.0f} GB/s = {nbytes / ms / 1e6 / HBM:.2f} of HBM")


x, w, dy = rnd(M, D), rnd(D), rnd(M, D)
y, rstd, dx = torch.empty_like(x), torch.empty(M, device="cuda"), torch.empty_like(x)
rec("rmsnorm_fwd", timed(lambda: check(lib().al_rmsnorm_forward(ptr(x), ptr(w), ptr(y), ptr(rstd), M, D, 1e-5, stream_ptr()))), M * D * 4)
rec("rmsnorm_bwd", timed(lambda: check(lib().al_rmsnorm_backward(ptr(x), ptr(w), ptr(rstd), ptr(dy), ptr(dx), M, D, stream_ptr()))), M * D * 6)
ga, up, dh = rnd(M, F), rnd(M, F), rnd(M, F)
h, dg, du = torch.empty_like(ga), torch.empty_like(ga), torch.empty_like(ga)
rec("swiglu_fwd", timed(lambda: check(lib().al_swiglu_forward(ptr(ga), ptr(up), ptr(h), ga.numel(), stream_ptr()))), M * F * 6)
rec("swiglu_bwd", timed(lambda: check(lib().al_swiglu_backward(ptr(ga), ptr(up), ptr(dh), ptr(dg), ptr(du), ga.numel(), stream_ptr()))), M * F * 10)
q = rnd(8, 2014, 24, 128)
cs, sn = rnd(1, 2014, 128), rnd(1, 2014, 128)
qo = torch.empty_like(q)
rec("rope (q, 24 heads)", timed(lambda: check(lib().al_rope(ptr(q), ptr(cs), ptr(sn), ptr(qo), 8, 2014, 24, 128, 1, 0, stream_ptr()))), q.numel() * 4)
R = 2048
ldv = (V + 7) // 8 * 8
logits = torch.randn(R, ldv, device="cuda", generator=g).bfloat16()
labels = torch.randint(0, V, (R,), device="cuda", generator=g)
loss = torch.zeros(1, device="cuda")
rec("ce_inplace (2048 rows)", timed(lambda: check(lib().al_cross_entropy_inplace(ptr(logits), ptr(labels), R, V, ldv, 1.0, ptr(loss), stream_ptr()))), R * V * 4)
# weight-gradient GEMM (MN-major operands): dA = U^T x and a projector-sized dW
U, xx = rnd(M, 64), rnd(M, D)
dA = torch.zeros(64, D, device="cuda")
ms = timed(lambda: check(lib().al_gemm_tn_accumulate(ptr(U), 64, 64, ptr(xx), D, D, M, ptr(dA), D, stream_ptr())))
rec("gemm_tn dA [64 x 3072], K=16112", ms, M * (64 + D) * 2)
dyp, hp = rnd(48000, 2048), rnd(48000, 1664)
dW = torch.zeros(2048, 1664, device="cuda")
ms = timed(lambda: check(lib().al_gemm_tn_accumulate(ptr(dyp), 2048, 2048, ptr(hp), 1664, 1664, 48000, ptr(dW), 1664, stream_ptr())))
out["gemm_tn dW2 projector [2048 x 1664], K=48000"] = {"ms": ms, "tflops": 2.0 * 48000 * 2048 * 1664 / ms / 1e9}
print(f"gemm_tn dW2 projector        {ms * 1e3:8.1f} us  {2.0 * 48000 * 2048 * 1664 / ms / 1e9:7.0f} TFLOP/s")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_llama_rows.json", "w"), indent=1)
