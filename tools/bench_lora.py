#!/usr/bin/env python
"""Times the fused frozen-linear + LoRA GEMM (al_lora_linear_forward) against the frozen GEMM alone, and the native
backward (al_lora_linear_backward), on Llama-3.2-3B's four LoRA-targeted shapes at the README batch (8 x 2014 rows,
rank 64). Prints TFLOP/s (frozen + low-rank FLOPs) and the overhead of the low-rank path."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_llama_b200 import ops

M, R = 8 * 2014, 64
SHAPES = {"q_proj": (3072, 3072), "kv_proj": (3072, 1024), "gate_up": (3072, 8192), "down": (8192, 3072)}


def timed(fn, n=20):
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


out = {}
g = torch.Generator(device="cuda").manual_seed(0)
for name, (i, o) in SHAPES.items():
    x = torch.randn(M, i, device="cuda", generator=g).bfloat16()
    W = (torch.randn(o, i, device="cuda", generator=g) * 0.02).bfloat16()
    A = torch.randn(R, i, device="cuda", generator=g) * 0.05
    B = torch.randn(o, R, device="cuda", generator=g) * 0.05
    dy = torch.randn(M, o, device="cuda", generator=g).bfloat16()
    Wt = W.t().contiguous()
    t_plain = timed(lambda: ops.gemm_bf16(x, W))
    packed = ops.pack_lora(A, B, 0.25)       # (the parameters change once per optimizer step, not per call)
    t_fused = timed(lambda: ops.lora_linear(x, W, None, A, B, 0.25, packed=packed))
    t_fused_pack = timed(lambda: ops.lora_linear(x, W, None, A, B, 0.25))
    _, (a_pad, b_pad, t) = ops.lora_linear(x, W, None, A, B, 0.25, return_saved=True)
    t_bwd = timed(lambda: ops.lora_linear_backward(x, dy, Wt, a_pad, b_pad, t, R))
    t_dgrad = timed(lambda: ops.gemm_bf16(dy, Wt))
    Af, Bf = A.bfloat16(), B.bfloat16()

    def torch_bwd():                                   # the same rank-r formulas on cuBLAS
        u = dy @ Bf
        dx = dy @ W + (u @ Af) * 0.25
        dA = (u.T @ x) * 0.25
        dB = (dy.T @ (x @ Af.T)) * 0.25
        return dx, dA, dB
    t_torch = timed(torch_bwd)
    f_frozen = 2.0 * M * i * o
    f_lora = 2.0 * M * R * (i + o)
    out[name] = {"in": i, "out": o, "frozen_gemm_ms": t_plain, "fused_fwd_ms": t_fused,
                 "fwd_overhead": t_fused / t_plain - 1.0, "fused_fwd_with_pack_ms": t_fused_pack, "fused_fwd_tflops": (f_frozen + f_lora) / t_fused / 1e9,
                 "bwd_ms": t_bwd, "frozen_dgrad_ms": t_dgrad, "bwd_over_dgrad": t_bwd / t_dgrad, "torch_cublas_bwd_ms": t_torch,
                 "bwd_tflops": (f_frozen + 3 * f_lora) / t_bwd / 1e9}
    print(name, json.dumps(out[name]))
json.dump(out, open("gpurun_out/bench_lora.json", "w"), indent=1)
