"""AudioProjector — same constructor, attribute names and checkpoint keys as the reference
(/root/reference/src/models/projector.py:5-19: `layers` = Sequential(Linear, GELU, Linear, LayerNorm), keys
`layers.{0,2,3}.{weight,bias}`), forward on B200 through `al_projector_forward`:
GEMM1 + bias + erf-GELU epilogue -> GEMM2 + bias (fp32 out) -> LayerNorm, optionally storing straight into
`inputs_embeds[b, 1 + t]` (the splice for the audio rows).

Forward and backward (the projector is the trainable part of the path) both run on the hand-written sm_100a
kernels: `al_projector_forward` / `al_projector_backward`.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from .._lib import check, lib, ptr, stream_ptr


def projector_forward_raw(w: Dict[str, torch.Tensor], x_bf16: torch.Tensor, out_dtype=torch.bfloat16,
                          out: Optional[torch.Tensor] = None, rows_per_group: Optional[int] = None,
                          out_group_stride: int = 0, out_row_offset: int = 0, cache: Optional[dict] = None,
                          keep: Optional[dict] = None) -> torch.Tensor:
    """x [rows, d_in] bf16 on the GPU; w = fp32 (or pre-cast) parameter dict with the reference's key names."""
    if not x_bf16.is_cuda:
        raise ValueError("projector input must be on the GPU (no CPU fallback)")
    x2 = x_bf16.reshape(-1, x_bf16.shape[-1]).contiguous()
    rows, d_in = x2.shape
    W1, W2 = w["layers.0.weight"], w["layers.2.weight"]
    hidden, d_out = W1.shape[0], W2.shape[0]
    c = cache if cache is not None else {}
    if "w1" not in c:                     # bf16 copies of the trainable fp32 weights, refreshed by the caller
        c["w1"] = W1.detach().to(torch.bfloat16).contiguous()
        c["w2"] = W2.detach().to(torch.bfloat16).contiguous()
    f32 = lambda k: w[k].detach().to(torch.float32).contiguous()
    if keep is None and cache is not None:
        # inference path: the two scratch matrices live in the caller's cache (no allocation in steady state); the
        # training path (keep) hands them to autograd, so it gets fresh ones
        if c.get("ws_rows") != rows or c["h_ws"].device != x2.device:
            c["h_ws"] = torch.empty(rows, hidden, dtype=torch.bfloat16, device=x2.device)
            c["y_ws"] = torch.empty(rows, d_out, dtype=torch.float32, device=x2.device)
            c["ws_rows"] = rows
        h_ws, y_ws = c["h_ws"], c["y_ws"]
    else:
        h_ws = torch.empty(rows, hidden, dtype=torch.bfloat16, device=x2.device)
        y_ws = torch.empty(rows, d_out, dtype=torch.float32, device=x2.device)
    if out is not None and out.dtype not in (torch.bfloat16, torch.float32):
        # the LayerNorm kernel stores bf16 or fp32 only: anything else (an fp16 inputs_embeds buffer, say) would be
        # filled with bf16 bit patterns
        raise TypeError(f"projector output buffer must be bfloat16 or float32, got {out.dtype}")
    if out is None:
        out = torch.empty(*x_bf16.shape[:-1], d_out, dtype=out_dtype, device=x2.device)
        rows_per_group, out_group_stride, out_row_offset = max(rows, 1), 0, 0
    check(lib().al_projector_forward(
        ptr(x2), rows, d_in, hidden, d_out, ptr(c["w1"]), ptr(f32("layers.0.bias")), ptr(c["w2"]),
        ptr(f32("layers.2.bias")), ptr(f32("layers.3.weight")), ptr(f32("layers.3.bias")), ptr(h_ws), ptr(y_ws),
        ptr(out), 1 if out.dtype == torch.float32 else 0, out.shape[-1], rows_per_group, out_group_stride,
        out_row_offset, stream_ptr()), "al_projector_forward")
    if keep is not None:
        keep["h"], keep["y"] = h_ws, y_ws
    return out


class _ProjectorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W1, b1, W2, b2, g, beta):
        w = {"layers.0.weight": W1, "layers.0.bias": b1, "layers.2.weight": W2, "layers.2.bias": b2,
             "layers.3.weight": g, "layers.3.bias": beta}
        keep = {}
        xb = x.detach().to(torch.bfloat16)
        out = projector_forward_raw(w, xb, out_dtype=torch.float32, keep=keep)
        ctx.save_for_backward(xb, W1, b1, W2, g, keep["h"], keep["y"])
        return out.to(x.dtype) if x.dtype != torch.bfloat16 else out

    @staticmethod
    def backward(ctx, dout):
        """Native backward (`al_projector_backward`): LayerNorm-backward kernel, split-K tcgen05 GEMMs for dW1 / dW2,
        dh = dy W2, and the GELU derivative applied in the epilogue of the recomputed x W1^T GEMM."""
        import ctypes as C
        xb, W1, b1, W2, g, h, y = ctx.saved_tensors
        dev = xb.device
        x2 = xb.reshape(-1, xb.shape[-1]).contiguous()
        rows, d_in = x2.shape
        hidden, d_out = W1.shape[0], W2.shape[0]
        do = dout.reshape(rows, d_out).to(torch.float32).contiguous()
        L = lib()
        nbytes = L.al_projector_backward_workspace_bytes(rows, d_in, hidden, d_out)
        buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
        ws = buf[(-buf.data_ptr()) % 1024:]
        f32 = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
        dW1, db1, dW2, db2, dg, dbeta = f32(hidden, d_in), f32(hidden), f32(d_out, hidden), f32(d_out), f32(d_out), f32(d_out)
        w1b = W1.detach().to(torch.bfloat16).contiguous()
        w2b = W2.detach().to(torch.bfloat16).contiguous()
        check(L.al_projector_backward(ptr(x2), rows, d_in, hidden, d_out, ptr(w1b), ptr(b1.detach().float().contiguous()),
                                      ptr(w2b), ptr(g.detach().float().contiguous()), ptr(h), ptr(y), ptr(do), ptr(ws),
                                      ptr(dW1), ptr(db1), ptr(dW2), ptr(db2), ptr(dg), ptr(dbeta), stream_ptr()),
              "al_projector_backward")
        cast = lambda t, like: t.to(like.dtype)
        return None, cast(dW1, W1), cast(db1, b1), cast(dW2, W2), cast(db2, W2), cast(dg, g), cast(dbeta, g)


class AudioProjector(nn.Module):
    def __init__(self, input_dim, output_dim, hidden_dim=None):
        super().__init__()
        if hidden_dim is None:
            hidden_dim = (input_dim + output_dim) // 2
        # kept as the reference's Sequential so state_dict keys / parameter init are identical
        self.layers = nn.Sequential(
            nn.Linear(input_dim, hidden_dim),
            nn.GELU(),
            nn.Linear(hidden_dim, output_dim),
            nn.LayerNorm(output_dim),
        )

    def forward(self, x):
        l0, l2, l3 = self.layers[0], self.layers[2], self.layers[3]
        if not x.is_cuda:
            raise RuntimeError("AudioProjector (B200) needs CUDA tensors: there is no CPU fallback")
        return _ProjectorFn.apply(x, l0.weight, l0.bias, l2.weight, l2.bias, l3.weight, l3.bias)

    def forward_into(self, x_bf16: torch.Tensor, inputs_embeds: torch.Tensor, n_audio: int, row_offset: int = 1):
        """Inference fast path: LayerNorm stores into inputs_embeds[b, row_offset + t] (no autograd)."""
        w = {k: v for k, v in self.layers.state_dict().items()}
        S = inputs_embeds.shape[1]
        with torch.no_grad():
            projector_forward_raw(w, x_bf16, out=inputs_embeds, rows_per_group=n_audio, out_group_stride=S,
                                  out_row_offset=row_offset)
        return inputs_embeds
