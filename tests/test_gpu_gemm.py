"""GPU parity: tcgen05 GEMM (through the C ABI) vs a plain PyTorch fp32 reference of the same op on the same
bf16-rounded operands. Tolerance: fp32-accumulate-order noise only (1e-3 relative to the output scale) for fp32
outputs, plus one bf16 rounding (2^-8) for bf16 outputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from audio_llama_b200 import ops
from audio_llama_b200.ops import EPI_GELU, EPI_OUT_F32, EPI_REDUCE_ADD, EPI_RESIDUAL, EPI_ROWAUX


def rnd(*shape, seed=0, s=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * s


def check(y, ref, bf16_out):
    scale = ref.abs().max().item()
    err = (y.float().cpu() - ref).abs().max().item()
    tol = (2 ** -8 + 1e-3) if bf16_out else 1e-3
    assert err <= tol * scale, (err, scale)


SHAPES = [(128, 256, 64), (128, 256, 256), (256, 512, 1280), (1500, 1280, 1280), (3000, 3840, 1280),
          (1500, 5120, 1280), (1500, 1280, 5120), (77, 1664, 1280), (300, 384, 1152), (129, 320, 384), (64, 64, 72)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_bias_bf16(M, N, K):
    a, w, b = rnd(M, K, seed=1).bfloat16(), rnd(N, K, seed=2, s=0.05).bfloat16(), rnd(N, seed=3)
    ref = a.float() @ w.float().T + b
    y = ops.gemm_bf16(a.cuda(), w.cuda(), b.cuda())
    check(y, ref, True)


@pytest.mark.parametrize("M,N,K", [(1500, 5120, 1280), (200, 1664, 1280), (128, 256, 64)])
def test_gemm_gelu(M, N, K):
    a, w, b = rnd(M, K, seed=4).bfloat16(), rnd(N, K, seed=5, s=0.05).bfloat16(), rnd(N, seed=6)
    ref = F.gelu(a.float() @ w.float().T + b)
    check(ops.gemm_bf16(a.cuda(), w.cuda(), b.cuda(), flags=EPI_GELU), ref, True)


@pytest.mark.parametrize("M,N,K", [(1500, 1280, 1280), (130, 2048, 1664), (128, 256, 64)])
def test_gemm_f32_and_reduce_add(M, N, K):
    a, w, b = rnd(M, K, seed=7).bfloat16(), rnd(N, K, seed=8, s=0.05).bfloat16(), rnd(N, seed=9)
    ref = a.float() @ w.float().T + b
    y = ops.gemm_bf16(a.cuda(), w.cuda(), b.cuda(), flags=EPI_OUT_F32)
    check(y, ref, False)
    x0 = rnd(M, N, seed=10)
    x = x0.clone().cuda()
    ops.gemm_bf16(a.cuda(), w.cuda(), b.cuda(), flags=EPI_OUT_F32 | EPI_REDUCE_ADD, out=x)
    check(x, x0 + ref, False)
    ops.gemm_bf16(a.cuda(), w.cuda(), None, flags=EPI_OUT_F32 | EPI_REDUCE_ADD, out=x)       # no bias
    check(x, x0 + 2 * ref - b, False)
    # in-place residual add (what out_proj / fc2 use): x <- x + a w^T + b, no atomics
    x = x0.clone().cuda()
    ops.gemm_bf16(a.cuda(), w.cuda(), b.cuda(), flags=EPI_OUT_F32 | EPI_RESIDUAL, out=x, resid=x)
    check(x, x0 + ref, False)
    y2 = ops.gemm_bf16(a.cuda(), w.cuda(), b.cuda(), flags=EPI_OUT_F32 | EPI_RESIDUAL, resid=x0.cuda())   # out of place
    check(y2, x0 + ref, False)


def test_gemm_batched_rowaux():
    B, M, N, K = 3, 200, 384, 128
    a, w, b = rnd(B, M, K, seed=11).bfloat16(), rnd(N, K, seed=12, s=0.1).bfloat16(), rnd(N, seed=13)
    pos = rnd(M, N, seed=14)
    ref = F.gelu(a.float() @ w.float().T + b) + pos
    y = ops.gemm_bf16(a.cuda(), w.cuda(), b.cuda(), flags=EPI_OUT_F32 | EPI_GELU | EPI_ROWAUX, aux=pos.cuda())
    check(y, ref, False)


@pytest.mark.parametrize("cin,cout,T,stride", [(128, 384, 3000, 1), (384, 384, 3000, 2), (128, 1280, 600, 1)])
def test_conv_as_strided_gemm(cin, cout, T, stride):
    """conv1d(k=3, pad=1, stride s) as one GEMM over overlapping rows of the time-major padded input."""
    B = 2
    x = rnd(B, cin, T, seed=15)
    w = rnd(cout, cin, 3, seed=16, s=0.05)
    b = rnd(cout, seed=17)
    xb, wb = x.bfloat16(), w.bfloat16()
    ref = F.gelu(F.conv1d(xb.float(), wb.float(), b, stride=stride, padding=1)).permute(0, 2, 1)   # [B, T/s, cout]
    xt = torch.zeros(B, T + 2, cin, dtype=torch.bfloat16)
    xt[:, 1:T + 1] = xb.permute(0, 2, 1)
    wp = wb.permute(0, 2, 1).contiguous().view(cout, 3 * cin)
    To = T // stride
    out = torch.empty(B, To, cout, dtype=torch.bfloat16).cuda()
    ops.gemm_bf16_strided(xt.cuda(), stride * cin, (T + 2) * cin, To, B, wp.cuda(), b.cuda(), out, cout, To * cout,
                          flags=EPI_GELU)
    check(out, ref, True)


@pytest.mark.parametrize("M,N,K", [(64, 3072, 16112), (3072, 64, 16112), (72, 200, 1000), (8, 256, 77), (300, 520, 4096)])
@pytest.mark.parametrize("pair", [1, 0])
def test_gemm_tn_accumulate(M, N, K, pair):
    """Weight-gradient form (MN-major operand staging): out += A^T W for row-major A [K, M], W [K, N]; ragged M / N / K
    tails go through TMA zero-fill. fp32 reference of the bf16 operands; the existing contents of `out` are kept."""
    from audio_llama_b200._lib import check, lib, ptr, stream_ptr
    g = torch.Generator().manual_seed(M * 7 + N)
    lda, ldw = (M + 7) // 8 * 8, (N + 7) // 8 * 8
    A = torch.zeros(K, lda).bfloat16()
    W = torch.zeros(K, ldw).bfloat16()
    A[:, :M] = torch.randn(K, M, generator=g).bfloat16()
    W[:, :N] = torch.randn(K, N, generator=g).bfloat16()
    out0 = torch.randn(M, N, generator=g)
    ref = out0 + A[:, :M].float().t() @ W[:, :N].float()
    ldo = (N + 3) // 4 * 4
    out = torch.zeros(M, ldo)
    out[:, :N] = out0
    Ad, Wd, od = A.cuda(), W.cuda(), out.cuda()
    lib().al_gemm_set_mode(pair)
    try:
        check(lib().al_gemm_tn_accumulate(ptr(Ad), lda, M, ptr(Wd), ldw, N, K, ptr(od), ldo, stream_ptr()), "al_gemm_tn_accumulate")
    finally:
        lib().al_gemm_set_mode(1)
    got = od.cpu()[:, :N]
    assert (got - ref).norm() / ref.norm() <= 2e-3
    if ldo > N:
        assert (od.cpu()[:, N:] == 0).all()
