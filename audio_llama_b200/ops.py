"""Torch-tensor wrappers over the C ABI. PyTorch is plumbing here (device memory, streams); every
function launches hand-written sm_100a kernels from libaudiollm_sm100.so and raises if that fails."""
from __future__ import annotations

from typing import Optional

import torch

from ._lib import check, lib, ptr, stream_ptr

EPI_GELU, EPI_OUT_F32, EPI_REDUCE_ADD, EPI_ROWAUX, EPI_RESIDUAL = 1, 2, 4, 8, 16
MEL_WHISPER, MEL_TRAIN = 0, 1
N_FRAMES = 3000


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def mel_forward(wave: torch.Tensor, n_samples: Optional[torch.Tensor] = None, n_mels: int = 128,
                mode: int = MEL_WHISPER, out: Optional[torch.Tensor] = None, ws: Optional[torch.Tensor] = None,
                raw: bool = False) -> torch.Tensor:
    """wave [B, n] float32 (n >= 480000 unless n_samples is given) -> [B, n_mels, 3000] float32.
    `ws` = a caller-owned int32 [>= B] scratch (per-clip maxima), so a steady-state caller allocates nothing.
    raw=True (Whisper mode): stop before the per-clip floor; `out` then holds log10 values and `ws` the clip maxima,
    to be finished by WhisperEncoderB200.forward(mel, clip_max=ws) / ops.pack_mel(..., clip_max=ws)."""
    _req(wave, torch.float32, "wave")
    B, n = wave.shape
    if n_samples is None and n < 480000:
        n_samples = torch.full((B,), n, dtype=torch.int32, device=wave.device)
    if n_samples is not None:
        _req(n_samples, torch.int32, "n_samples")
    if out is None:
        out = torch.empty(B, n_mels, N_FRAMES, dtype=torch.float32, device=wave.device)
    if ws is None:
        ws = torch.empty(max(B, 1), dtype=torch.int32, device=wave.device)
    elif ws.dtype != torch.int32 or ws.numel() < B or not ws.is_cuda:
        raise ValueError("mel_forward: ws must be a CUDA int32 tensor with at least B elements")
    if raw and mode != MEL_WHISPER:
        raise ValueError("mel_forward: raw=True only applies to the Whisper mode")
    if mode == MEL_TRAIN:
        _ensure_train_bank(n_mels)
    check(lib().al_mel_forward_ex(ptr(wave), ptr(n_samples), B, n, n_mels, mode, 1 if raw else 0, ptr(out), ptr(ws),
                                  stream_ptr()), "al_mel_forward_ex")
    return out


_train_bank_installed = set()


def htk_filterbank_torch(n_mels: int) -> torch.Tensor:
    """The HTK bank exactly as torchaudio.functional.melscale_fbanks(201, 0, 8000, n_mels, 16000, norm=None,
    mel_scale="htk") builds it (float32 torch ops; TA functional.py:518-580) — what
    /root/reference/src/dataset.py:125-131 ends up using. Returns float32 [201, n_mels]."""
    import math
    all_freqs = torch.linspace(0, 8000, 201)
    m_min = 2595.0 * math.log10(1.0 + 0.0 / 700.0)
    m_max = 2595.0 * math.log10(1.0 + 8000.0 / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def _ensure_train_bank(n_mels: int):
    if n_mels in _train_bank_installed:
        return
    fb = htk_filterbank_torch(n_mels).double().contiguous().numpy()
    check(lib().al_mel_set_filterbank_host(n_mels, MEL_TRAIN, fb.ctypes.data), "al_mel_set_filterbank_host")
    _train_bank_installed.add(n_mels)


def mel_filterbank(n_mels: int, mode: int = MEL_WHISPER):
    if mode == MEL_TRAIN:
        _ensure_train_bank(n_mels)
    import numpy as np
    fb = np.empty((201, n_mels), dtype=np.float64)
    check(lib().al_mel_filterbank_host(n_mels, mode, fb.ctypes.data), "al_mel_filterbank_host")
    return fb


def gemm_bf16(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, flags: int = 0,
              out: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None,
              resid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """epilogue(a @ w.T + bias): a [M, K] or [batch, M, K] bf16, w [N, K] bf16. With EPI_REDUCE_ADD `out` (fp32)
    is accumulated into; with EPI_RESIDUAL `resid` (fp32, same shape as out, may be out itself) is added."""
    _req(a, torch.bfloat16, "a")
    _req(w, torch.bfloat16, "w")
    a3 = a if a.dim() == 3 else a.unsqueeze(0)
    batch, M, K = a3.shape
    N = w.shape[0]
    odt = torch.float32 if flags & EPI_OUT_F32 else torch.bfloat16
    if out is None:
        if flags & EPI_REDUCE_ADD:
            raise ValueError("EPI_REDUCE_ADD needs an existing `out`")
        out = torch.empty(*a.shape[:-1], N, dtype=odt, device=a.device)
    _req(out, odt, "out")
    if bias is not None:
        _req(bias, torch.float32, "bias")
    if aux is not None:
        _req(aux, torch.float32, "aux")
    if resid is not None:
        _req(resid, torch.float32, "resid")
    check(lib().al_gemm_bf16(ptr(a), K, M * K, M, batch, ptr(w), N, K, ptr(bias), ptr(out), N, M * N, flags,
                             ptr(aux), aux.shape[-1] if aux is not None else 0, ptr(resid), stream_ptr()),
          "al_gemm_bf16")
    return out


def gemm_bf16_strided(a_base: torch.Tensor, a_row_stride: int, a_batch_stride: int, m_per_batch: int, batch: int,
                      w: torch.Tensor, bias, out_base: torch.Tensor, o_row_stride: int, o_batch_stride: int,
                      flags: int = 0, aux=None):
    """Raw-stride form (overlapping rows allowed): what the conv stem uses."""
    N, K = w.shape
    check(lib().al_gemm_bf16(ptr(a_base), a_row_stride, a_batch_stride, m_per_batch, batch, ptr(w), N, K, ptr(bias),
                             ptr(out_base), o_row_stride, o_batch_stride, flags, ptr(aux),
                             aux.shape[-1] if aux is not None else 0, None, stream_ptr()), "al_gemm_bf16")
    return out_base


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
              out_dtype=torch.bfloat16, out: Optional[torch.Tensor] = None, rows_per_group: Optional[int] = None,
              out_group_stride: int = 0, out_row_offset: int = 0) -> torch.Tensor:
    _req(x, torch.float32, "x")
    d = x.shape[-1]
    rows = x.numel() // d
    if (out.dtype if out is not None else out_dtype) not in (torch.bfloat16, torch.float32):
        raise TypeError("layernorm stores bfloat16 or float32 only")
    if out is None:
        out = torch.empty(x.shape, dtype=out_dtype, device=x.device)
        rows_per_group, out_group_stride, out_row_offset = max(rows, 1), 0, 0
    elif rows_per_group is None:
        rows_per_group, out_group_stride, out_row_offset = max(rows, 1), 0, 0
    check(lib().al_layernorm(ptr(x), ptr(_req(gamma, torch.float32, "gamma")), ptr(_req(beta, torch.float32, "beta")),
                             ptr(out), rows, d, eps, 1 if out.dtype == torch.float32 else 0, out.shape[-1],
                             rows_per_group, out_group_stride, out_row_offset, stream_ptr()), "al_layernorm")
    return out


def attention(qkv: torch.Tensor, n_heads: int, q_log2: bool = False) -> torch.Tensor:
    """qkv [B, T, 3*H*64] bf16 (q pre-scaled by head_dim^-1/2; q_log2: and by log2 e) -> [B, T, H*64] bf16."""
    _req(qkv, torch.bfloat16, "qkv")
    B, T, d3 = qkv.shape
    if d3 != 3 * n_heads * 64:
        raise ValueError("attention kernel is head_dim 64 only")
    out = torch.empty(B, T, n_heads * 64, dtype=torch.bfloat16, device=qkv.device)
    check(lib().al_attention_ex(ptr(qkv), ptr(out), B, T, n_heads, 1 if q_log2 else 0, stream_ptr()), "al_attention_ex")
    return out


def pack_mel(mel: torch.Tensor, c_pad: int, clip_max: Optional[torch.Tensor] = None) -> torch.Tensor:
    """clip_max = the `ws` of mel_forward(raw=True): the per-clip floor and (x + 4) / 4 are applied while packing."""
    _req(mel, torch.float32, "mel")
    B, n_mels, T = mel.shape
    out = torch.empty(B, T + 2, c_pad, dtype=torch.bfloat16, device=mel.device)
    check(lib().al_pack_mel_ex(ptr(mel), ptr(clip_max), ptr(out), B, n_mels, T, c_pad, stream_ptr()), "al_pack_mel_ex")
    return out


def f32_to_bf16(x: torch.Tensor) -> torch.Tensor:
    _req(x, torch.float32, "x")
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(lib().al_f32_to_bf16(ptr(x), ptr(out), x.numel(), stream_ptr()), "al_f32_to_bf16")
    return out


def splice(table: torch.Tensor, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor],
           labels: Optional[torch.Tensor], n_audio: int, start_id: int, end_id: int,
           audio_rows: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
           want_mask: bool = True, want_labels: bool = True, mask_out: Optional[torch.Tensor] = None,
           labels_out: Optional[torch.Tensor] = None, check_ids: bool = True):
    """S1/S2. Returns (inputs_embeds [B, n_audio+2+T, d], mask float32 | None, labels int64 | None).
    With audio_rows=None rows 1..n_audio of `out` are left for the projector to fill in place.
    `out` / `mask_out` / `labels_out` may be caller-owned buffers (a steady-state caller then allocates nothing).
    An input id outside the table raises IndexError, as the reference's embed_tokens(input_ids) does (allm.py:64);
    check_ids=False defers that to raise_if_bad_ids() (no host sync on the hot path; the kernel never reads outside
    the table either way)."""
    if not table.is_cuda or not table.is_contiguous():
        raise ValueError("table must be a contiguous CUDA tensor")
    if table.dtype not in (torch.bfloat16, torch.float32, torch.float16):
        raise TypeError("embedding table must be bf16 / fp16 / fp32")
    _req(input_ids, torch.int64, "input_ids")
    vocab, d = table.shape
    if start_id >= vocab or end_id >= vocab:
        # same check and message as the reference (allm.py:140-141)
        raise ValueError(f"Token IDs {start_id}, {end_id} are outside vocabulary size {vocab}")
    B, T = input_ids.shape
    S = n_audio + 2 + T
    if out is None:
        out = torch.empty(B, S, d, dtype=table.dtype, device=table.device)
    if audio_rows is not None:
        if audio_rows.dtype != table.dtype or not audio_rows.is_contiguous():
            raise TypeError("audio_rows must be contiguous and of the table's dtype")
    if attention_mask is not None:
        _req(attention_mask, torch.int64, "attention_mask")
    if labels is not None:
        _req(labels, torch.int64, "labels")
    if mask_out is None:
        mask_out = torch.empty(B, S, dtype=torch.float32, device=table.device) if want_mask else None
    else:
        _req(mask_out, torch.float32, "mask_out")
    if labels is None:
        labels_out = None
    elif labels_out is None:
        labels_out = torch.empty(B, S, dtype=torch.int64, device=table.device) if want_labels else None
    else:
        _req(labels_out, torch.int64, "labels_out")
    flag = bad_id_flag(table.device)
    check(lib().al_splice(ptr(table), table.element_size(), d, ptr(input_ids), ptr(attention_mask), ptr(labels), B, T,
                          n_audio, start_id, end_id, ptr(audio_rows), ptr(out), ptr(mask_out), ptr(labels_out),
                          vocab, ptr(flag), stream_ptr()), "al_splice")
    if check_ids and not DEFER_ID_CHECKS:
        raise_if_bad_ids(table.device, vocab)
    return out, mask_out, labels_out


_bad_id_flags = {}


def bad_id_flag(device) -> torch.Tensor:
    """The per-device int32 word the splice kernels OR with 1 when they meet an input id outside the table."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    f = _bad_id_flags.get(key)
    if f is None:
        f = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", key))
        _bad_id_flags[key] = f
    return f


# True: splice() never synchronises to check its ids, whatever check_ids says; the caller reads the flag later through
# raise_if_bad_ids() (a training step replayed as a CUDA graph cannot read the host inside the step)
DEFER_ID_CHECKS = False


def raise_if_bad_ids(device, vocab=None):
    """Synchronises; raises IndexError if a splice launch on this device met an out-of-range input id since the last
    check (and clears the flag)."""
    f = bad_id_flag(device)
    if int(f.item()) != 0:
        f.zero_()
        raise IndexError("index out of range in self: input_ids outside the embedding table"
                         + (f" (vocabulary size {vocab})" if vocab is not None else ""))


def pack_lora(lora_a: torch.Tensor, lora_b: torch.Tensor, scaling: float):
    """The kernel's operands: A zero-padded to a multiple of 8 rows, bf16 [r_pad, in]; scaling * B, bf16 [out, r_pad]."""
    r, in_dim = lora_a.shape
    out_dim = lora_b.shape[0]
    r_pad = (r + 7) // 8 * 8
    if lora_a.is_cuda and lora_a.dtype == torch.float32 and lora_b.dtype == torch.float32 and lora_a.is_contiguous() \
            and lora_b.is_contiguous():
        a = torch.empty(r_pad, in_dim, dtype=torch.bfloat16, device=lora_a.device)
        b = torch.empty(out_dim, r_pad, dtype=torch.bfloat16, device=lora_b.device)
        check(lib().al_lora_pack(ptr(lora_a.detach()), ptr(lora_b.detach()), r, in_dim, out_dim, float(scaling), ptr(a), ptr(b),
                                 stream_ptr()), "al_lora_pack")
        return a, b
    a = torch.zeros(r_pad, in_dim, dtype=torch.bfloat16, device=lora_a.device)
    a[:r] = lora_a.detach().to(torch.bfloat16)
    b = torch.zeros(out_dim, r_pad, dtype=torch.bfloat16, device=lora_b.device)
    b[:, :r] = (lora_b.detach().float() * scaling).to(torch.bfloat16)
    return a, b


def lora_linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], lora_a: torch.Tensor,
                lora_b: torch.Tensor, scaling: float, out_dtype=torch.bfloat16, return_saved: bool = False,
                packed=None, addend: Optional[torch.Tensor] = None):
    """Frozen linear + LoRA (L1): x [..., in] bf16, w [out, in] bf16, lora_a [r, in], lora_b [out, r] (any float dtype).
    out = x w^T + bias + scaling * (x a^T) b^T with the rank-r product accumulated inside the frozen GEMM.
    `packed` = pack_lora(lora_a, lora_b, scaling) when the caller keeps it across calls (the parameters only change at
    optimizer steps); return_saved=True also returns (a_pad, b_scaled_pad, t) for lora_linear_backward. `addend`
    (bf16, the output's shape) is added in the GEMM epilogue (the residual connection around down_proj)."""
    _req(x, torch.bfloat16, "x")
    _req(w, torch.bfloat16, "w")
    out_dim, in_dim = w.shape
    a, b = packed if packed is not None else pack_lora(lora_a, lora_b, scaling)
    r_pad = a.shape[0]
    x2 = x.reshape(-1, in_dim)
    rows = x2.shape[0]
    t_ws = torch.empty(rows, r_pad, dtype=torch.bfloat16, device=x.device)
    out = torch.empty(*x.shape[:-1], out_dim, dtype=out_dtype, device=x.device)
    if bias is not None:
        bias = _req(bias.detach().float().contiguous(), torch.float32, "bias")
    if addend is not None:
        if out_dtype != torch.bfloat16 or addend.shape != out.shape:
            raise ValueError("lora_linear: an addend needs a bf16 output of the same shape")
        addend = _req(addend.contiguous(), torch.bfloat16, "addend")
    check(lib().al_lora_linear_forward_ex(ptr(x2), rows, in_dim, out_dim, r_pad, ptr(w), ptr(bias), ptr(a), ptr(b),
                                          ptr(t_ws), ptr(addend), ptr(out), 1 if out_dtype == torch.float32 else 0,
                                          stream_ptr()), "al_lora_linear_forward")
    if return_saved:
        return out, (a, b, t_ws)
    return out


def lora_linear_backward(x: torch.Tensor, dy: torch.Tensor, w_t: Optional[torch.Tensor], a_pad: torch.Tensor,
                         b_scaled_pad: torch.Tensor, t_saved: torch.Tensor, rank: int, need_dx: bool = True,
                         dx_accumulate: Optional[torch.Tensor] = None):
    """Backward of lora_linear with the weight frozen. x [..., in], dy [..., out] bf16; w_t = w.t().contiguous()
    ([in, out] bf16, needed only for dx). Returns (dx or None, dA [rank, in] f32, dB_raw [out, rank] f32) where the
    gradient of the unscaled lora_B is scaling * dB_raw (dA already carries the scaling). `dx_accumulate` (bf16,
    x's shape, contiguous): dx is added INTO it in the GEMM epilogue and it is returned as dx (several projections of one
    input: their input gradients sum without separate add passes)."""
    _req(x, torch.bfloat16, "x")
    dy = _req(dy.contiguous(), torch.bfloat16, "dy")
    r_pad, in_dim = a_pad.shape
    out_dim = b_scaled_pad.shape[0]
    x2 = x.reshape(-1, in_dim)
    dy2 = dy.reshape(-1, out_dim)
    rows = x2.shape[0]
    nbytes = int(lib().al_lora_linear_backward_workspace_bytes(rows, in_dim, out_dim, r_pad))
    ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=x.device)
    off = (-ws.data_ptr()) % 1024
    ws = ws[off:off + nbytes]
    dx = None
    if need_dx:
        _req(w_t, torch.bfloat16, "w_t")
        if dx_accumulate is not None:
            if not dx_accumulate.is_contiguous() or dx_accumulate.numel() != x2.numel():
                raise ValueError("lora_linear_backward: dx_accumulate must be contiguous with x's element count")
            dx = _req(dx_accumulate, torch.bfloat16, "dx_accumulate").view(rows, in_dim)
        else:
            dx = torch.empty_like(x2)
    dAB = torch.empty(r_pad * in_dim + out_dim * r_pad, dtype=torch.float32, device=x.device)   # (one memset for both)
    dA = dAB[:r_pad * in_dim].view(r_pad, in_dim)
    dB = dAB[r_pad * in_dim:].view(out_dim, r_pad)
    check(lib().al_lora_linear_backward_ex(ptr(x2), ptr(dy2), rows, in_dim, out_dim, r_pad, ptr(w_t) if need_dx else None,
                                           ptr(a_pad), ptr(b_scaled_pad), ptr(t_saved), ptr(ws),
                                           ptr(dx) if (need_dx and dx_accumulate is not None) else None, ptr(dx), ptr(dA),
                                           ptr(dB), stream_ptr()), "al_lora_linear_backward")
    return (dx.view_as(x) if need_dx else None), dA[:rank], dB[:, :rank]


class TensorDerivedCache:
    """value = build(tensor), rebuilt when the tensor object, its storage address or its version counter changes (an
    in-place optimizer step or load_state_dict bumps `_version`). Entries die with their tensor (weak references)."""

    def __init__(self):
        self._d = {}

    def get(self, tensors, build, extra=None):
        import weakref
        tensors = tuple(tensors)
        key = tuple(id(t) for t in tensors)
        stamp = tuple((t.data_ptr(), t._version, t.device) for t in tensors) + (extra,)
        hit = self._d.get(key)
        if hit is not None and hit[1] == stamp and all(r() is t for r, t in zip(hit[0], tensors)):
            return hit[2]
        if len(self._d) > 4096:                       # dead entries of freed models
            self._d = {k: v for k, v in self._d.items() if all(r() is not None for r in v[0])}
        val = build()
        self._d[key] = (tuple(weakref.ref(t) for t in tensors), stamp, val)
        return val


def launch_count() -> int:
    return int(lib().al_launch_count())
