"""GPU parity: LayerNorm, mel repack, splice gather (bit-exact), casts."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from audio_llama_b200 import ops, synth
from oracle import encoder as O


@pytest.mark.parametrize("rows,d", [(1, 256), (37, 384), (1500, 1280), (3000, 2048), (513, 3072)])
@pytest.mark.parametrize("odt", [torch.bfloat16, torch.float32])
def test_layernorm(rows, d, odt):
    g = torch.Generator().manual_seed(rows + d)
    x = (torch.randn(rows, d, generator=g) * 3 + 0.5)
    gamma = 1 + 0.1 * torch.randn(d, generator=g)
    beta = 0.1 * torch.randn(d, generator=g)
    ref = F.layer_norm(x, (d,), gamma, beta, 1e-5)
    y = ops.layernorm(x.cuda(), gamma.cuda(), beta.cuda(), out_dtype=odt).float().cpu()
    tol = 1e-5 if odt == torch.float32 else 1.6e-2
    assert (y - ref).abs().max() <= tol * max(1.0, ref.abs().max())
    if odt == torch.bfloat16:   # exactly the bf16 rounding of the fp32 result, up to 1 ulp of fp32 noise
        assert (y - ref.bfloat16().float()).abs().max() <= 2 ** -7 * ref.abs().max()


def test_layernorm_row_mapping():
    """Projector LN stores straight into inputs_embeds[b, 1 + t]."""
    B, A, T, d = 3, 50, 7, 256
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B * A, d, generator=g).cuda()
    gamma = torch.ones(d).cuda(); beta = torch.zeros(d).cuda()
    S = A + 2 + T
    out = torch.full((B, S, d), 7.0, dtype=torch.float32).cuda()
    ops.layernorm(x, gamma, beta, out=out, rows_per_group=A, out_group_stride=S, out_row_offset=1)
    ref = F.layer_norm(x.cpu(), (d,)).view(B, A, d)
    assert (out[:, 1:1 + A].cpu() - ref).abs().max() < 1e-5
    assert (out[:, 0] == 7.0).all() and (out[:, 1 + A:] == 7.0).all()


@pytest.mark.parametrize("n_mels,c_pad", [(128, 128), (80, 128)])
def test_pack_mel(n_mels, c_pad):
    g = torch.Generator().manual_seed(1)
    mel = torch.randn(2, n_mels, 3000, generator=g)
    out = ops.pack_mel(mel.cuda(), c_pad).cpu()
    assert out.shape == (2, 3002, c_pad)
    assert (out[:, 0] == 0).all() and (out[:, 3001] == 0).all()
    assert (out[:, 1:3001, :n_mels] == mel.permute(0, 2, 1).bfloat16()).all()
    assert (out[:, :, n_mels:] == 0).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,T,A,d", [(4, 16, 1500, 256), (2, 512, 1500, 2048), (1, 1, 3, 8)])
def test_splice_bit_exact(dtype, B, T, A, d):
    vocab = 300
    g = torch.Generator().manual_seed(2)
    E = torch.randn(vocab, d, generator=g).to(dtype)
    ids, mask, labels = synth.synth_text(B, T, vocab)
    proj = torch.randn(B, A, d, generator=g).to(dtype)
    ref = O.combine(E, ids, proj, vocab - 2, vocab - 1)
    out, m, lab = ops.splice(E.cuda(), ids.cuda(), mask.cuda(), labels.cuda(), A, vocab - 2, vocab - 1,
                             audio_rows=proj.cuda())
    assert torch.equal(out.cpu().view(torch.uint8), ref.contiguous().view(torch.uint8))      # every byte
    assert m.dtype == torch.float32 and torch.equal(m.cpu(), O.extend_mask(mask, A))
    assert torch.equal(lab.cpu(), O.extend_labels(labels, A + 2))
    # in-place form: audio rows untouched
    pre = torch.full((B, A + 2 + T, d), 3.0).to(dtype).cuda()
    out2, _, _ = ops.splice(E.cuda(), ids.cuda(), mask.cuda(), None, A, vocab - 2, vocab - 1, out=pre)
    assert (out2[:, 1:1 + A] == 3.0).all()
    assert torch.equal(out2[:, 0].cpu(), ref[:, 0]) and torch.equal(out2[:, A + 1:].cpu(), ref[:, A + 1:])


def test_splice_bad_delimiter_raises():
    E = torch.zeros(10, 8).cuda()
    with pytest.raises(ValueError):
        ops.splice(E, torch.zeros(1, 3, dtype=torch.long).cuda(), None, None, 5, 10, 9)


@pytest.mark.parametrize("bad", [-1, 300, 2 ** 40])
def test_splice_out_of_range_input_id_raises(bad):
    """An input id outside the table: the reference's embed_tokens(input_ids) raises (allm.py:64). The kernel must not
    read outside the table (guard rows around it stay intact), zeroes that output row, and the host side raises."""
    vocab, d, B, T, A = 300, 64, 2, 9, 4
    big = torch.full((vocab + 2, d), 7.0).cuda()                     # guard row before and after the real table
    E = big[1:1 + vocab]
    E.copy_(torch.randn(vocab, d, generator=torch.Generator().manual_seed(0)))
    ids, mask, labels = synth.synth_text(B, T, vocab)
    ids[1, 3] = bad
    with pytest.raises(IndexError):
        ops.splice(E, ids.cuda(), mask.cuda(), labels.cuda(), A, vocab - 2, vocab - 1, audio_rows=None)
    # deferred form: no exception at the call, rows of good ids exact, the bad row zeroed, the flag raised later
    out, _, _ = ops.splice(E, ids.cuda(), mask.cuda(), labels.cuda(), A, vocab - 2, vocab - 1, audio_rows=None, check_ids=False)
    good = ids.clone()
    good[1, 3] = 0
    ref = E.cpu()[good]
    ref[1, 3] = 0
    assert torch.equal(out[:, A + 2:].cpu(), ref)
    assert (out[:, A + 2:] != 7.0).all()                             # nothing came from the guard rows
    with pytest.raises(IndexError):
        ops.raise_if_bad_ids(E.device, vocab)
    ops.raise_if_bad_ids(E.device, vocab)                            # the flag was cleared
    # ragged form
    from audio_llama_b200.splice import splice_ragged
    audio = torch.randn(B, 1500, d).cuda()
    with pytest.raises(IndexError):
        splice_ragged(E, ids.cuda(), mask.cuda(), labels.cuda(), audio, [[10], [20]], vocab - 2, vocab - 1)


def test_layernorm_and_projector_refuse_fp16_output():
    x = torch.randn(8, 64).cuda()
    with pytest.raises(TypeError):
        ops.layernorm(x, torch.ones(64).cuda(), torch.zeros(64).cuda(), out=torch.empty(8, 64, dtype=torch.float16).cuda())
    from audio_llama_b200.models.projector import projector_forward_raw
    pw = synth.init_projector_weights(64, 32, seed=0)
    pw = {k: v.cuda() for k, v in pw.items()}
    with pytest.raises(TypeError):
        projector_forward_raw(pw, torch.randn(8, 64).bfloat16().cuda(), out=torch.empty(8, 32, dtype=torch.float16).cuda(),
                              rows_per_group=8, out_group_stride=0, out_row_offset=0)
    from audio_llama_b200.config import EncoderConfig
    from audio_llama_b200.pipeline import AudioConditioner
    cfg = EncoderConfig(d_model=128, n_layers=1, n_heads=2, ffn_dim=256, n_mels=80)
    with pytest.raises(TypeError):
        AudioConditioner(cfg, synth.init_encoder_weights(cfg, seed=0), synth.init_projector_weights(128, 64, seed=1),
                         torch.zeros(10, 64, dtype=torch.float16), 8, 9, max_batch=1)


def test_f32_to_bf16():
    x = torch.randn(4096 * 3)
    assert torch.equal(ops.f32_to_bf16(x.cuda()).cpu(), x.bfloat16())
