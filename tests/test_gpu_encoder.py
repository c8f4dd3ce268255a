"""GPU parity: encoder forward (conv stem + layers + final LN), projector, against the fp32 oracle.
Tolerance: 2e-2 relative (rel-L2, bf16 compute vs fp32 oracle) — BASELINE.md §4."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from audio_llama_b200 import ops, synth
from audio_llama_b200.config import WHISPER_TINY, WHISPER_TINY_128, WHISPER_LARGE_V3_TURBO, EncoderConfig
from audio_llama_b200.encoder import WhisperEncoderB200
from oracle import encoder as O
from golden_signals import encoder_input

TOL = 2e-2


def run_case(cfg, B, seed, layers=None):
    torch.set_num_threads(os.cpu_count())
    w = synth.init_encoder_weights(cfg, seed=seed, ln_jitter=0.1)
    mel = torch.from_numpy(encoder_input(cfg, B))
    ref, taps = O.encoder_forward(w, cfg, mel, return_layers=True)
    enc = WhisperEncoderB200(cfg, w, max_batch=B, out_dtype=torch.float32)
    y = enc(mel.cuda()).cpu()
    return y, ref, taps, enc, mel


def test_stem_only():
    """conv1 + GELU + conv2 + GELU + positions = fp32 residual stream before layer 0."""
    cfg = EncoderConfig(d_model=256, n_layers=0, n_heads=4, ffn_dim=512, n_mels=80)
    y, ref, taps, enc, mel = run_case(cfg, 2, 3)
    x0 = enc.hidden_state(2).cpu()
    assert O.rel_l2(x0, taps[0]) <= 1e-2
    assert O.rel_l2(y, ref) <= TOL


@pytest.mark.parametrize("cfg,B,seed", [
    (EncoderConfig(d_model=256, n_layers=2, n_heads=4, ffn_dim=512, n_mels=80), 2, 3),
    (WHISPER_TINY_128, 2, 0),
    (WHISPER_TINY, 3, 1),
])
def test_encoder_small(cfg, B, seed):
    y, ref, taps, enc, mel = run_case(cfg, B, seed)
    assert y.shape == ref.shape == (B, 1500, cfg.d_model)
    assert torch.isfinite(y).all()
    rel = O.rel_l2(y, ref)
    assert rel <= TOL, rel
    xl = enc.hidden_state(B).cpu()
    assert O.rel_l2(xl, taps[-1]) <= TOL


def test_encoder_matches_hf_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "encoder_hf.npz"))
    cfg = WHISPER_TINY_128
    w = synth.init_encoder_weights(cfg, seed=0, ln_jitter=0.1)
    enc = WhisperEncoderB200(cfg, w, max_batch=2, out_dtype=torch.float32)
    y = enc(torch.from_numpy(encoder_input(cfg, 2)).cuda()).cpu()
    ref = torch.from_numpy(g["tiny128_grid"])
    assert O.rel_l2(y[:, ::25, ::16], ref) <= TOL


def test_encoder_chunks_and_bf16_out():
    cfg = EncoderConfig(d_model=128, n_layers=1, n_heads=2, ffn_dim=256, n_mels=80)
    w = synth.init_encoder_weights(cfg, seed=5, ln_jitter=0.1)
    mel = torch.from_numpy(encoder_input(cfg, 3)).cuda()
    e2 = WhisperEncoderB200(cfg, w, max_batch=2, out_dtype=torch.bfloat16)    # 3 clips through a 2-clip plan
    e3 = WhisperEncoderB200(cfg, w, max_batch=3, out_dtype=torch.float32)
    y2, y3 = e2(mel), e3(mel)
    assert y2.dtype == torch.bfloat16
    assert O.rel_l2(y2.float(), y3) <= 5e-3
    assert torch.equal(e3(mel[1:2]), y3[1:2])            # clips are independent: same bits alone or in a batch
    with pytest.raises(ValueError):
        e3(mel[..., :2999])


def test_encoder_turbo_one_clip():
    """Full-size architecture (whisper-large-v3-turbo shape, 32 layers), 1 clip, vs the fp32 oracle."""
    cfg = WHISPER_LARGE_V3_TURBO
    y, ref, taps, enc, mel = run_case(cfg, 1, 0)
    rel = O.rel_l2(y, ref)
    assert rel <= TOL, rel


def test_encoder_turbo_matches_hf_golden(golden_dir):
    """The named architecture (whisper-large-v3-turbo encoder) against the installed HF WhisperEncoder itself (fp32 CPU
    run of tests/golden/make_golden.py turbo): sub-sampled grid and two full rows, rel-L2 <= 2e-2 (north-star bound)."""
    g = np.load(os.path.join(golden_dir, "encoder_hf_turbo.npz"))
    cfg = WHISPER_LARGE_V3_TURBO
    w = synth.init_encoder_weights(cfg, seed=0, ln_jitter=0.1)
    enc = WhisperEncoderB200(cfg, w, max_batch=1, out_dtype=torch.float32)
    y = enc(torch.from_numpy(encoder_input(cfg, 1)).cuda()).cpu()
    assert O.rel_l2(y[:, ::25, ::16], torch.from_numpy(g["turbo_grid"])) <= TOL
    assert O.rel_l2(y[:, 7, :], torch.from_numpy(g["turbo_row7"])) <= TOL
    assert O.rel_l2(y[:, 1499, :], torch.from_numpy(g["turbo_row1499"])) <= TOL
    nrm = float(y.double().norm())
    assert abs(nrm - g["turbo_norm"][0]) <= 1e-2 * g["turbo_norm"][0]


def test_encoder_turbo_batch_independence():
    """Turbo shape, B = 3 (M = 4500 rows: the 256-row GEMM tiles and the attention tiles straddle clip boundaries):
    clip 1 alone is bit-identical to clip 1 inside the batch, and so is the last clip."""
    cfg = WHISPER_LARGE_V3_TURBO
    w = synth.init_encoder_weights(cfg, seed=0, ln_jitter=0.1)
    enc = WhisperEncoderB200(cfg, w, max_batch=3, out_dtype=torch.bfloat16)
    mel = torch.from_numpy(encoder_input(cfg, 3)).cuda()
    y3 = enc(mel).clone()
    assert torch.equal(enc(mel[1:2]), y3[1:2])
    assert torch.equal(enc(mel[2:3]), y3[2:3])
    assert not torch.equal(y3[0], y3[1])


@pytest.mark.parametrize("d_in,d_out,rows", [(384, 256, 300), (1280, 2048, 1500), (1280, 3072, 257)])
def test_projector(d_in, d_out, rows):
    from audio_llama_b200.models.projector import projector_forward_raw
    pw = synth.init_projector_weights(d_in, d_out, seed=1, ln_jitter=0.1)
    x = torch.randn(rows, d_in, generator=torch.Generator().manual_seed(9))
    ref = O.projector_forward(pw, x.bfloat16().float())
    y = projector_forward_raw({k: v.cuda() for k, v in pw.items()}, x.cuda().bfloat16(), out_dtype=torch.float32).cpu()
    assert O.rel_l2(y, ref) <= TOL
    assert (y - ref).abs().max() <= 0.1


@pytest.mark.parametrize("d_in,d_out,rows", [(384, 256, 300), (1280, 2048, 3000), (1280, 3072, 1500)])
def test_projector_backward_native(d_in, d_out, rows):
    """Gradients of the native projector backward vs torch autograd of the fp32 oracle (on bf16-rounded input)."""
    from audio_llama_b200.models.projector import AudioProjector
    pw = synth.init_projector_weights(d_in, d_out, seed=1, ln_jitter=0.1)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(rows, d_in, generator=g).bfloat16()
    dout = torch.randn(rows, d_out, generator=g) * 0.1
    # oracle gradients
    ref_w = {k: v.clone().requires_grad_(True) for k, v in pw.items()}
    O.projector_forward(ref_w, x.float()).backward(dout)
    # native
    proj = AudioProjector(d_in, d_out).cuda()
    proj.load_state_dict(pw)
    y = proj(x.cuda())
    y.backward(dout.cuda().to(y.dtype))
    names = {"layers.0.weight": proj.layers[0].weight, "layers.0.bias": proj.layers[0].bias,
             "layers.2.weight": proj.layers[2].weight, "layers.2.bias": proj.layers[2].bias,
             "layers.3.weight": proj.layers[3].weight, "layers.3.bias": proj.layers[3].bias}
    for k, p in names.items():
        rel = O.rel_l2(p.grad.float().cpu(), ref_w[k].grad)
        assert rel <= 3e-2, (k, rel)
