"""LoRA — same names and semantics as the reference (/root/reference/src/models/lora.py:6-43):
`LoRALayer(in_dim, out_dim, rank=8, alpha=16)` with `lora_A [rank, in]` (zeros), `lora_B [out, rank]`
(N(0, 0.01)), `scaling = alpha / rank`; `apply_lora_to_llama` matches q/k/v/gate/up/down by substring;
`lora_forward_hook` returns `output + lora(x)`.

What changes is the arithmetic order: the reference materialises the dense [out, in] delta `B @ A` on every
call and runs a second full-size GEMM (lora.py:20-21). Here the update is the rank-r side path
`(x @ A^T) @ B^T * scaling` — 2*M*r*(in+out) FLOPs instead of 2*out*in*r + 2*M*in*out — identical in exact
arithmetic, within fp32 round-off in practice (tests/test_oracle_golden.py pins it to the reference's own output, tests/test_gpu_lora.py checks the fused kernel).
"""
from __future__ import annotations

import torch
import torch.nn as nn


class LoRALayer(nn.Module):
    def __init__(self, in_dim, out_dim, rank=8, alpha=16):
        super().__init__()
        self.lora_A = nn.Parameter(torch.zeros(rank, in_dim))
        self.lora_B = nn.Parameter(torch.randn(out_dim, rank) * 0.01)
        self.rank = rank
        self.alpha = alpha
        self.scaling = alpha / rank
        nn.init.zeros_(self.lora_A)
        nn.init.normal_(self.lora_B, std=0.01)

    def forward(self, x):
        # rank-r side path; never forms the [out, in] matrix
        a = self.lora_A.to(x.dtype)
        b = self.lora_B.to(x.dtype)
        return ((x @ a.T) @ b.T) * self.scaling


def apply_lora_to_llama(llama_model, rank=8, alpha=16, target_modules=None):
    """Returns {qualified module name: LoRALayer} for every nn.Linear whose name contains a target
    (lora.py:23-39)."""
    if target_modules is None:
        target_modules = ['q_proj', 'k_proj', 'v_proj', 'gate_proj', 'up_proj', 'down_proj']
    lora_layers = {}
    for name, module in llama_model.named_modules():
        if isinstance(module, nn.Linear):
            if any(target_name in name for target_name in target_modules):
                lora_layers[name] = LoRALayer(module.in_features, module.out_features, rank, alpha)
    return lora_layers


def lora_forward_hook(module, input, output, lora_layer):
    """Add LoRA output to the original linear layer output (lora.py:41-43)."""
    return output + lora_layer(input[0])


# ----------------------------------------------------------------------------- fused path (B200)
def _frozen_weight_t(weight: torch.Tensor) -> torch.Tensor:
    """The frozen weight transposed, [in, out], built once per weight tensor."""
    from .. import ops
    global _WTS
    if _WTS is None:
        _WTS = ops.TensorDerivedCache()
    return _WTS.get((weight,), lambda: weight.detach().t().contiguous())


def _packed_operands(lora_A, lora_B, scaling):
    """pack_lora(A, B, scaling), repacked after optimizer steps (the parameters' version counters move)."""
    from .. import ops
    global _PACKS
    if REPACK_ALWAYS:
        return ops.pack_lora(lora_A, lora_B, scaling)
    if _PACKS is None:
        _PACKS = ops.TensorDerivedCache()
    return _PACKS.get((lora_A, lora_B), lambda: ops.pack_lora(lora_A, lora_B, scaling), extra=scaling)


_PACKS = None
_WTS = None
# True: the bf16 operands are rebuilt from the parameters on every forward instead of once per optimizer step. A step that
# is captured as a CUDA graph needs the pack launches INSIDE the graph (a cache hit at capture time would freeze the
# LoRA weights of every replay at their captured values).
REPACK_ALWAYS = False


class _FusedLoRALinearFn(torch.autograd.Function):
    """y = x W^T + b + s (x A^T) B^T through `al_lora_linear_forward` (the rank-r product rides in the frozen GEMM's
    TMEM accumulator). Backward keeps W frozen and is native too (`al_lora_linear_backward`): dx = dy W + s (dy B) A in
    one GEMM with the low-rank pair in the K loop, dA = s (dy B)^T x and dB = s dy^T (x A^T) as split-K GEMMs."""

    @staticmethod
    def forward(ctx, x, weight, bias, lora_A, lora_B, scaling):
        from .. import ops
        y, (a_pad, b_pad, t) = ops.lora_linear(x.contiguous(), weight, bias, lora_A, lora_B, scaling, out_dtype=x.dtype,
                                              return_saved=True, packed=_packed_operands(lora_A, lora_B, scaling))
        ctx.save_for_backward(x, weight, a_pad, b_pad, t)
        ctx.scaling = scaling
        ctx.rank = lora_A.shape[0]
        ctx.dtypes = (lora_A.dtype, lora_B.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        from .. import ops
        x, weight, a_pad, b_pad, t = ctx.saved_tensors
        need_dx = ctx.needs_input_grad[0]
        dx, dA, dB_raw = ops.lora_linear_backward(x.contiguous(), dy.to(torch.bfloat16), _frozen_weight_t(weight) if need_dx else None,
                                                  a_pad, b_pad, t, ctx.rank, need_dx=need_dx)
        return dx, None, None, dA.to(ctx.dtypes[0]), (dB_raw * ctx.scaling).to(ctx.dtypes[1]), None


def fused_lora_forward(module: nn.Linear, lora_layer: LoRALayer, x: torch.Tensor) -> torch.Tensor:
    """Replacement for `module(x)` + lora_forward_hook when x / W are bf16 CUDA tensors."""
    return _FusedLoRALinearFn.apply(x, module.weight, module.bias, lora_layer.lora_A, lora_layer.lora_B,
                                    lora_layer.scaling)


class _FusedLoRALinearAddFn(torch.autograd.Function):
    """y = addend + x W^T + b + s (x A^T) B^T: _FusedLoRALinearFn with the residual connection folded into the GEMM
    epilogue (`al_lora_linear_forward_ex`). The addend's gradient is dy itself."""

    @staticmethod
    def forward(ctx, x, weight, bias, lora_A, lora_B, scaling, addend):
        from .. import ops
        y, (a_pad, b_pad, t) = ops.lora_linear(x.contiguous(), weight, bias, lora_A, lora_B, scaling, out_dtype=x.dtype,
                                              return_saved=True, packed=_packed_operands(lora_A, lora_B, scaling),
                                              addend=addend)
        ctx.save_for_backward(x, weight, a_pad, b_pad, t)
        ctx.scaling = scaling
        ctx.rank = lora_A.shape[0]
        ctx.dtypes = (lora_A.dtype, lora_B.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        from .. import ops
        x, weight, a_pad, b_pad, t = ctx.saved_tensors
        need_dx = ctx.needs_input_grad[0]
        dyb = dy.to(torch.bfloat16)
        dx, dA, dB_raw = ops.lora_linear_backward(x.contiguous(), dyb, _frozen_weight_t(weight) if need_dx else None,
                                                  a_pad, b_pad, t, ctx.rank, need_dx=need_dx)
        return dx, None, None, dA.to(ctx.dtypes[0]), (dB_raw * ctx.scaling).to(ctx.dtypes[1]), None, \
            (dy if ctx.needs_input_grad[6] else None)


def fused_lora_forward_add(module: nn.Linear, lora_layer: LoRALayer, x: torch.Tensor, addend: torch.Tensor) -> torch.Tensor:
    """addend + module(x) + lora(x) in one GEMM (down_proj and the residual connection around the MLP)."""
    return _FusedLoRALinearAddFn.apply(x, module.weight, module.bias, lora_layer.lora_A, lora_layer.lora_B,
                                       lora_layer.scaling, addend)


class _FusedLoRAMultiFn(torch.autograd.Function):
    """Several LoRA-carrying projections of ONE input (q / k / v, gate / up): forward = one fused GEMM each; backward sums
    their input gradients inside the GEMM epilogues (dx = dy_0 W_0 + ...; the next product adds into the same buffer,
    `al_lora_linear_backward_ex`) instead of leaving the sum to autograd's elementwise adds.
    Arguments: x, then (weight, bias, lora_A, lora_B, scaling) per projection."""

    @staticmethod
    def forward(ctx, x, *args):
        from .. import ops
        n = len(args) // 5
        xc = x.contiguous()
        outs, saved, meta = [], [xc], []
        for i in range(n):
            weight, bias, lora_A, lora_B, scaling = args[5 * i:5 * i + 5]
            y, (a_pad, b_pad, t) = ops.lora_linear(xc, weight, bias, lora_A, lora_B, scaling, out_dtype=x.dtype,
                                                  return_saved=True, packed=_packed_operands(lora_A, lora_B, scaling))
            outs.append(y)
            saved += [weight, a_pad, b_pad, t]
            meta.append((scaling, lora_A.shape[0], lora_A.dtype, lora_B.dtype))
        ctx.save_for_backward(*saved)
        ctx.meta = meta
        return tuple(outs)

    @staticmethod
    def backward(ctx, *dys):
        from .. import ops
        saved = ctx.saved_tensors
        x = saved[0]
        need_dx = ctx.needs_input_grad[0]
        dx = None
        grads = []
        for i, dy in enumerate(dys):
            weight, a_pad, b_pad, t = saved[1 + 4 * i:5 + 4 * i]
            scaling, rank, dt_a, dt_b = ctx.meta[i]
            if dy is None:
                grads += [None, None, None, None, None]
                continue
            dxi, dA, dB_raw = ops.lora_linear_backward(x, dy.to(torch.bfloat16), _frozen_weight_t(weight) if need_dx else None,
                                                       a_pad, b_pad, t, rank, need_dx=need_dx, dx_accumulate=dx)
            if need_dx:
                dx = dxi
            grads += [None, None, dA.to(dt_a), (dB_raw * scaling).to(dt_b), None]
        return (dx, *grads)


def fused_lora_multi(x: torch.Tensor, pairs):
    """pairs: [(nn.Linear, LoRALayer), ...] sharing the input x -> tuple of outputs."""
    args = []
    for module, lora_layer in pairs:
        args += [module.weight, module.bias, lora_layer.lora_A, lora_layer.lora_B, lora_layer.scaling]
    return _FusedLoRAMultiFn.apply(x, *args)

