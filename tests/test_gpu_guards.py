"""GPU: out-of-bounds guards. compute-sanitizer is closed on this pool, so every kernel that writes through tails
(TMA clipping, masked rows, partial tiles) is run with its output placed between sentinel regions that must
come back untouched."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from audio_llama_b200 import ops, synth
from audio_llama_b200.ops import EPI_GELU, EPI_OUT_F32, EPI_REDUCE_ADD

PAD = 4096


def guarded(shape, dtype, fill):
    n = 1
    for s in shape:
        n *= s
    buf = torch.full((n + 2 * PAD,), fill, dtype=dtype, device="cuda")
    return buf, buf[PAD:PAD + n].view(*shape)


def check_guards(buf, fill, n):
    assert (buf[:PAD] == fill).all() and (buf[PAD + n:] == fill).all()


@pytest.mark.parametrize("M,N,K", [(77, 1664, 1280), (129, 320, 384), (300, 200, 72), (1, 8, 8), (255, 257, 64)])
def test_gemm_tails_do_not_write_outside(M, N, K):
    g = torch.Generator().manual_seed(0)
    a = torch.randn(M, K, generator=g).bfloat16().cuda()
    w = (torch.randn(N, K, generator=g) * 0.05).bfloat16().cuda()
    b = torch.randn(N, generator=g).cuda()
    for flags, dt, fill in ((0, torch.bfloat16, 7.0), (EPI_GELU, torch.bfloat16, 7.0), (EPI_OUT_F32, torch.float32, 7.0)):
        if (dt == torch.bfloat16 and N % 8) or (dt == torch.float32 and N % 4):
            # output rows not 16-byte aligned: the library refuses (argument error), it never writes
            from audio_llama_b200._lib import AudioLLMLibError
            buf, out = guarded((M, N), dt, fill)
            with pytest.raises(AudioLLMLibError, match="multiple of 16"):
                ops.gemm_bf16(a, w, b, flags=flags, out=out)
            check_guards(buf, fill, M * N)
            continue
        buf, out = guarded((M, N), dt, fill)
        ops.gemm_bf16(a, w, b, flags=flags, out=out)
        torch.cuda.synchronize()
        check_guards(buf, fill, M * N)
        ref = a.float() @ w.float().T + b
        if flags == EPI_GELU:
            ref = torch.nn.functional.gelu(ref)
        assert (out.float() - ref).abs().max() <= (2 ** -7) * max(1.0, ref.abs().max().item())
    if N % 4 == 0:
        buf, out = guarded((M, N), torch.float32, 0.0)
        ops.gemm_bf16(a, w, None, flags=EPI_OUT_F32 | EPI_REDUCE_ADD, out=out)
        torch.cuda.synchronize()
        check_guards(buf, 0.0, M * N)


@pytest.mark.parametrize("B,T,H", [(1, 92, 1), (2, 300, 2), (1, 129, 3)])
def test_attention_tail_rows(B, T, H):
    qkv = (torch.randn(B, T, 3 * H * 64, generator=torch.Generator().manual_seed(1)) * 0.5).bfloat16().cuda()
    # ops.attention allocates its own output; run it against a guarded copy through the raw entry point
    from audio_llama_b200._lib import check, lib, ptr, stream_ptr
    buf, out = guarded((B, T, H * 64), torch.bfloat16, 3.0)
    check(lib().al_attention(ptr(qkv), ptr(out), B, T, H, stream_ptr()), "al_attention")
    torch.cuda.synchronize()
    check_guards(buf, 3.0, B * T * H * 64)
    assert torch.isfinite(out.float()).all() and not (out == 3.0).all()


def test_layernorm_splice_mel_guards():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(37, 384, generator=g).cuda()
    buf, out = guarded((37, 384), torch.bfloat16, 5.0)
    ops.layernorm(x, torch.ones(384).cuda(), torch.zeros(384).cuda(), out=out)
    torch.cuda.synchronize()
    check_guards(buf, 5.0, 37 * 384)
    E = torch.randn(50, 64, generator=g).cuda()
    ids, mask, labels = (t.cuda() for t in synth.synth_text(3, 5, 50))
    buf, out = guarded((3, 7 + 5, 64), torch.float32, 9.0)
    ops.splice(E, ids, mask, labels, 5, 48, 49, audio_rows=torch.zeros(3, 5, 64).cuda(), out=out)
    torch.cuda.synchronize()
    check_guards(buf, 9.0, 3 * 12 * 64)
    wave = torch.from_numpy(synth.synth_batch(2, n_samples=20000)).cuda()
    buf, out = guarded((2, 80, 3000), torch.float32, -7.0)
    ops.mel_forward(wave, n_mels=80, out=out)
    torch.cuda.synchronize()
    check_guards(buf, -7.0, 2 * 80 * 3000)
    assert (out != -7.0).all()
