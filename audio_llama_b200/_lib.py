"""ctypes binding of libaudiollm_sm100.so (the C ABI declared in include/audiollm_b200.h).

There is NO fallback: if the shared library is missing or a call fails, this raises. Build it with
`python -c "import __graft_entry__ as g; g.build()"` (or `make -C audio_llama_b200/csrc`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AUDIOLLM_B200_LIB lets kernel experiments (tools/) load an alternative build of the same library
LIB_PATH = os.environ.get("AUDIOLLM_B200_LIB") or os.path.join(_HERE, "libaudiollm_sm100.so")

_lib = None

vp = C.c_void_p
i32 = C.c_int
i64 = C.c_longlong
f32 = C.c_float
sz = C.c_size_t

# name -> (restype, argtypes); mirrors include/audiollm_b200.h one to one
SIGNATURES = {
    "al_version": (i32, []),
    "al_last_error": (C.c_char_p, []),
    "al_launch_count": (i64, []),
    "al_mel_forward": (i32, [vp, vp, i32, i64, i32, i32, vp, vp, vp]),
    "al_mel_forward_ex": (i32, [vp, vp, i32, i64, i32, i32, i32, vp, vp, vp]),
    "al_pack_mel_ex": (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
    "al_encoder_forward_ex": (i32, [vp, vp, vp, i32, vp, i32, i32, vp]),
    "al_mel_filterbank_host": (i32, [i32, i32, vp]),
    "al_mel_set_filterbank_host": (i32, [i32, i32, vp]),
    "al_mel_set_mode": (i32, [i32]),
    "al_ingest_forward": (i32, [vp, i64, i64, i32, vp, i32, i32, i32, vp, i64, i32, vp, i32, vp]),
    "al_gemm_bf16": (i32, [vp, i64, i64, i32, i32, vp, i32, i32, vp, vp, i64, i64, i32, vp, i32, vp, vp]),
    "al_gemm_set_mode": (i32, [i32]),
    "al_gemm_tn_accumulate": (i32, [vp, i64, i32, vp, i64, i32, i32, vp, i64, vp]),
    "al_layernorm": (i32, [vp, vp, vp, vp, i32, i32, f32, i32, i64, i32, i64, i64, vp]),
    "al_attention": (i32, [vp, vp, i32, i32, i32, vp]),
    "al_attention_ex": (i32, [vp, vp, i32, i32, i32, i32, vp]),
    "al_gqa_attention_forward": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, vp]),
    "al_gqa_attention_backward": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, vp]),
    "al_encoder_set_options": (i32, [vp, i32]),
    "al_pack_mel": (i32, [vp, vp, i32, i32, i32, i32, vp]),
    "al_f32_to_bf16": (i32, [vp, vp, i64, vp]),
    "al_encoder_workspace_bytes": (sz, [i32, i32, i32, i32, i32, i32]),
    "al_encoder_create": (i32, [C.POINTER(vp), i32, i32, i32, i32, i32, i32, vp, sz]),
    "al_encoder_set_stem": (i32, [vp] * 8),
    "al_encoder_set_layer": (i32, [vp, i32] + [vp] * 12),
    "al_encoder_forward": (i32, [vp, vp, i32, vp, i32, i32, vp]),
    "al_encoder_hidden": (vp, [vp]),
    "al_encoder_set_profiling": (i32, [vp, i32]),
    "al_encoder_profile_read": (i32, [vp, vp, vp]),
    "al_encoder_destroy": (i32, [vp]),
    "al_projector_forward": (i32, [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i64, i32, i64, i64, vp]),
    "al_projector_backward_workspace_bytes": (sz, [i32, i32, i32, i32]),
    "al_projector_backward": (i32, [vp, i32, i32, i32, i32] + [vp] * 15),
    "al_lora_linear_forward": (i32, [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp]),
    "al_lora_linear_backward_workspace_bytes": (sz, [i32, i32, i32, i32]),
    "al_lora_linear_backward": (i32, [vp, vp, i32, i32, i32, i32] + [vp] * 9),
    "al_lora_pack": (i32, [vp, vp, i32, i32, i32, f32, vp, vp, vp]),
    "al_lora_linear_forward_ex": (i32, [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, i32, vp]),
    "al_lora_linear_backward_ex": (i32, [vp, vp, i32, i32, i32, i32] + [vp] * 10),
    "al_linear_add_bf16": (i32, [vp, i32, i32, i32, vp, vp, vp, vp, vp]),
    "al_rmsnorm_forward": (i32, [vp, vp, vp, vp, i32, i32, f32, vp]),
    "al_rmsnorm_backward": (i32, [vp, vp, vp, vp, vp, i32, i32, vp]),
    "al_rmsnorm_backward_ex": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, vp]),
    "al_swiglu_forward": (i32, [vp, vp, vp, i64, vp]),
    "al_swiglu_backward": (i32, [vp, vp, vp, vp, vp, i64, vp]),
    "al_rope": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "al_cross_entropy_inplace": (i32, [vp, vp, i32, i32, i64, f32, vp, vp]),
    "al_linear_ce_workspace_bytes": (sz, [i32, i32]),
    "al_linear_ce": (i32, [vp, vp, vp, vp, i32, i32, i32, f32, i32, vp, vp, vp, vp]),
    "al_splice": (i32, [vp, i32, i32, vp, vp, vp, i32, i32, i32, i64, i64, vp, vp, vp, vp, i64, vp, vp]),
    "al_splice_ragged": (i32, [vp, i32, i32, vp, vp, vp, i32, i32, i32, vp, vp, vp, i32, vp, i64, i64, vp, vp, vp, vp, i64, vp, vp]),
}


class AudioLLMLibError(RuntimeError):
    pass


def lib():
    """The loaded library; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AudioLLMLibError(
                f"{LIB_PATH} not found: the sm_100a extension is not built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)          # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().al_last_error().decode("utf-8", "replace")
        raise AudioLLMLibError(f"{what or 'audiollm_b200'} failed (rc={rc}): {msg}")


def ptr(t):
    """Device (or host) pointer of a torch tensor / None."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
