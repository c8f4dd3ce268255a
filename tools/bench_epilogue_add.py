import os, sys
sys.path.insert(0, "/root/repo")
import torch
from audio_llama_b200 import ops
M, R = 8 * 2014, 64
def timed(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
g = torch.Generator(device="cuda").manual_seed(0)
for name, (i, o) in {"q_proj": (3072, 3072), "kv_proj": (3072, 1024), "gate_up": (3072, 8192), "down": (8192, 3072)}.items():
    x = torch.randn(M, i, device="cuda", generator=g).bfloat16()
    W = (torch.randn(o, i, device="cuda", generator=g) * 0.02).bfloat16()
    A = torch.randn(R, i, device="cuda", generator=g) * 0.05
    B = torch.randn(o, R, device="cuda", generator=g) * 0.05
    dy = torch.randn(M, o, device="cuda", generator=g).bfloat16()
    r = torch.randn(M, o, device="cuda", generator=g).bfloat16()
    _, (a_pad, b_pad, t) = ops.lora_linear(x, W, None, A, B, 0.25, return_saved=True)
    wt = W.t().contiguous()
    acc = torch.zeros(M, i, device="cuda", dtype=torch.bfloat16)
    f0 = timed(lambda: ops.lora_linear(x, W, None, A, B, 0.25, packed=(a_pad, b_pad)))
    f1 = timed(lambda: ops.lora_linear(x, W, None, A, B, 0.25, packed=(a_pad, b_pad), addend=r))
    b0 = timed(lambda: ops.lora_linear_backward(x, dy, wt, a_pad, b_pad, t, R))
    b1 = timed(lambda: ops.lora_linear_backward(x, dy, wt, a_pad, b_pad, t, R, dx_accumulate=acc))
    print(f"{name:8s} fwd {f0:.3f} -> +addend {f1:.3f} ms   bwd {b0:.3f} -> +accumulate {b1:.3f} ms")

if os.environ.get("KERNELS"):
    from torch.profiler import ProfilerActivity, profile
    for label, kw in (("plain", {}), ("accumulate", {"dx_accumulate": acc})):
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                ops.lora_linear_backward(x, dy, wt, a_pad, b_pad, t, R, **kw)
            torch.cuda.synchronize()
        print(f"-- {name} backward, {label}")
        for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:5]:
            print(f"   {ev.device_time_total / ev.count:8.1f} us x{ev.count}  {ev.key[:70]}")
