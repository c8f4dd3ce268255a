"""GPU parity through the reference-facing API: AudioLLM (config 1) against the fixture the REFERENCE's own
AudioLLM produced (tests/golden/allm_config1.npz), the feature extractor, the full-size pipeline properties and
the ragged extension."""
import contextlib
import io
import os
from unittest.mock import Mock, patch

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from audio_llama_b200 import ops, synth
from audio_llama_b200.config import WHISPER_TINY_128, WHISPER_LARGE_V3_TURBO, EncoderConfig
from audio_llama_b200.features import LogMelExtractor, train_log_mel
from audio_llama_b200.models import base as B
from audio_llama_b200.models.allm import AudioLLM
from audio_llama_b200.pipeline import AudioConditioner
from oracle import encoder as O
from oracle import mel as M

VOCAB = 320


def fake_load(llama_path, whisper_path):
    """Same random-init recipe tests/golden/make_golden.py used for the reference run."""
    from transformers import LlamaConfig, LlamaForCausalLM, WhisperConfig, WhisperModel
    cfg = WHISPER_TINY_128
    torch.manual_seed(0)
    lc = LlamaConfig(vocab_size=VOCAB, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                     num_attention_heads=4, num_key_value_heads=4, max_position_embeddings=4096)
    wc = WhisperConfig(vocab_size=51865, num_mel_bins=128, d_model=cfg.d_model, encoder_layers=cfg.n_layers,
                       encoder_attention_heads=cfg.n_heads, encoder_ffn_dim=cfg.ffn_dim, decoder_layers=1,
                       decoder_attention_heads=cfg.n_heads, decoder_ffn_dim=cfg.ffn_dim)
    enc = WhisperModel(wc).eval().encoder                 # (same construction order as the reference run: RNG state)
    enc.load_state_dict(synth.init_encoder_weights(cfg, seed=0, ln_jitter=0.1))
    llama = LlamaForCausalLM(lc).eval()
    return B.FrozenModelWrapper(llama), B.FrozenModelWrapper(B.wrap_hf_encoder(enc, max_batch=4))


@pytest.fixture(scope="module")
def model():
    with patch.object(B, "load_base_models", fake_load):
        m = AudioLLM("x", "y", lora_rank=8)
    m.projector.load_state_dict(synth.init_projector_weights(384, 256, seed=1, ln_jitter=0.1))
    tok = Mock()
    tok.convert_tokens_to_ids = lambda t: {"<audio>": VOCAB - 2, "</audio>": VOCAB - 1}[t]
    m.tokenizer = tok
    return m.to("cuda")


def test_audiollm_config1_matches_reference_run(model, golden_dir):
    g = np.load(os.path.join(golden_dir, "allm_config1.npz"))
    E = model.llama.model.model.embed_tokens.weight.detach().cpu()
    assert torch.equal(E, torch.from_numpy(g["embed_table"]))             # same random-init LLaMA as the reference run
    ids, mask, labels = (torch.from_numpy(g[k]).cuda() for k in ("ids", "mask", "labels"))
    Bsz, T = ids.shape
    clips = [synth.synth_clip(i) for i in range(Bsz)]
    feats = LogMelExtractor(128)(clips, sampling_rate=16000).input_features.unsqueeze(1)    # [B,1,128,3000]
    assert feats.shape == (Bsz, 1, 128, 3000) and feats.dtype == torch.float32
    with torch.no_grad():
        enc = model._process_audio_features(feats)
        text_emb = model.llama.model.model.embed_tokens(ids)
        comb = model._combine_text_and_audio_embeddings(text_emb, feats, ids)
        ext = model._extend_attention_mask(mask, 1500)
        out = model(input_ids=ids, attention_mask=mask, audio_features=feats, labels=labels)
    assert O.rel_l2(enc[:, ::25, ::16].float().cpu(), torch.from_numpy(g["enc_grid"])) <= 2e-2
    assert tuple(comb.shape) == tuple(g["combined_shape"]) == (Bsz, 1502 + T, 256)
    rows = [0, 1, 2, 750, 1500, 1501, 1502, 1503, 1517]
    ref_rows = torch.from_numpy(g["combined_rows"])
    c = comb.float().cpu()
    for i, r in enumerate(rows):
        if r in (0, 1501) or r >= 1502:
            assert torch.equal(c[:, r], ref_rows[:, i])                   # gathers: bit-exact
    audio_idx = [i for i, r in enumerate(rows) if 1 <= r <= 1500]
    assert O.rel_l2(c[:, [rows[i] for i in audio_idx]], ref_rows[:, audio_idx]) <= 2e-2
    assert O.rel_l2(c[:, ::53, ::8], torch.from_numpy(g["combined_grid"])) <= 2e-2
    assert str(ext.dtype) == g["ext_mask_dtype"][0] and torch.equal(ext.cpu(), torch.from_numpy(g["ext_mask"]))
    assert tuple(out.logits.shape) == tuple(g["logits_shape"])
    assert abs(float(out.loss) - float(g["loss"][0])) <= 2e-2 * abs(float(g["loss"][0]))
    assert sum(p.numel() for p in model.get_trainable_params()) == int(g["n_trainable"][0])


def test_audiollm_trains_projector_and_lora(model):
    ids, mask, labels = (t.cuda() for t in synth.synth_text(2, 16, VOCAB))
    feats = LogMelExtractor(128)([synth.synth_clip(i) for i in range(2)], sampling_rate=16000).input_features.unsqueeze(1)
    for l in model.lora_layers.values():
        torch.nn.init.normal_(l.lora_A, std=0.01)
    for p in model.get_trainable_params():
        p.grad = None
    out = model(input_ids=ids, attention_mask=mask, audio_features=feats, labels=labels)
    out.loss.backward()
    grads = [p.grad for p in model.get_trainable_params()]
    assert all(g is not None and torch.isfinite(g).all() for g in grads)
    assert any(g.abs().sum() > 0 for g in grads[:6])                      # projector receives gradient through LLaMA
    assert all(p.grad is None for p in model.llama.model.parameters())   # frozen


def test_audiollm_error_and_text_only(model):
    ids, mask, _ = (t.cuda() for t in synth.synth_text(1, 8, VOCAB))
    out = model(input_ids=ids, attention_mask=mask)                       # audio_features=None -> text-only path
    assert out.logits.shape[1] == 8
    bad = Mock()
    bad.convert_tokens_to_ids = lambda t: VOCAB + 5
    good = model.tokenizer
    model.tokenizer = bad
    try:
        with pytest.raises(ValueError, match="outside vocabulary size"):
            model._combine_text_and_audio_embeddings(None, torch.zeros(1, 1, 128, 3000).cuda(), ids)
    finally:
        model.tokenizer = good
    with pytest.raises(ValueError):
        model._process_audio_features(torch.zeros(1, 1, 128, 2999).cuda())


def test_audiollm_generate(model):
    """generate(): same conditioning, stock HF sampling on inputs_embeds; returns only the new tokens (allm.py:333-346)."""
    ids, mask, _ = (t.cuda() for t in synth.synth_text(1, 8, VOCAB))
    feats = LogMelExtractor(128)([synth.synth_clip(0)], sampling_rate=16000).input_features.unsqueeze(1)
    tok = model.tokenizer
    tok.pad_token_id, tok.bos_token_id, tok.eos_token_id = 0, 1, None
    tok.decode = lambda t, skip_special_tokens=True: " ".join(str(int(x)) for x in t)
    out = model.generate(input_ids=ids, attention_mask=mask, audio_features=feats, max_new_tokens=5, do_sample=False,
                         temperature=None, top_p=None)
    assert isinstance(out, str)
    out2 = model.generate(input_ids=ids, attention_mask=mask, audio_features=feats, max_new_tokens=5, do_sample=False,
                          temperature=None, top_p=None)
    assert out == out2                                         # greedy + deterministic kernels


def test_feature_extractor_api():
    fe = LogMelExtractor(80)
    x = synth.synth_clip(1, n_samples=50000)
    f = fe(x, sampling_rate=16000).input_features
    assert f.shape == (1, 80, 3000)
    ref = M.log_mel_whisper([x], 80)
    assert np.abs(f.cpu().numpy() - ref).max() <= 3e-5
    with pytest.raises(ValueError):
        fe(x, sampling_rate=8000)
    t = train_log_mel(torch.from_numpy(x))
    assert t.shape == (1, 1, 128, 3000)


def test_pipeline_full_size_properties():
    """BASELINE config 2 shapes (turbo -> d 2048, T_txt 512), 4 clips: size-independent properties."""
    cfg = WHISPER_LARGE_V3_TURBO
    Bsz, T, d_l, vocab = 4, 512, 2048, 4096
    ew = synth.init_encoder_weights(cfg, seed=0)
    pw = synth.init_projector_weights(cfg.d_model, d_l, seed=1)
    table = (torch.randn(vocab, d_l, generator=torch.Generator().manual_seed(2)) * 0.02).bfloat16().cuda()
    cond = AudioConditioner(cfg, ew, pw, table, vocab - 2, vocab - 1, max_batch=Bsz)
    ids, mask, labels = (t.cuda() for t in synth.synth_text(Bsz, T, vocab))
    wave = torch.from_numpy(synth.synth_batch(Bsz)).cuda()
    emb, m, lab = cond(wave, ids, mask, labels)
    assert emb.shape == (Bsz, 2014, d_l) and torch.isfinite(emb.float()).all()
    assert torch.equal(emb[:, 0], table[vocab - 2].expand(Bsz, -1)) and torch.equal(emb[:, 1501], table[vocab - 1].expand(Bsz, -1))
    assert torch.equal(emb[:, 1502:], table[ids])                          # bit-exact gather
    a = emb[:, 1:1501].float()                                             # LayerNorm output (gamma 1, beta 0): rows ~ N(0,1)
    assert a.mean(-1).abs().max() < 2e-2 and (a.var(-1, unbiased=False) - 1).abs().max() < 5e-2
    assert torch.equal(m[:, :1502], torch.ones(Bsz, 1502, device="cuda")) and torch.equal(m[:, 1502:], mask.float())
    assert (lab[:, :1502] == -100).all() and torch.equal(lab[:, 1502:], labels)
    # batch independence: clip 2 alone gives the same bits
    emb1, _, _ = cond(wave[2:3], ids[2:3], mask[2:3], labels[2:3])
    assert torch.equal(emb1[0], emb[2])


def test_ragged_splice_extension():
    """Config 5 semantics (not in the reference): device prefix sums bit-exact vs the oracle; k=1 full clip == S1."""
    from audio_llama_b200.splice import splice_ragged
    g = torch.Generator().manual_seed(5)
    vocab, d, T = 64, 32, 6
    E = torch.randn(vocab, d, generator=g)
    ids, mask, labels = synth.synth_text(3, T, vocab)
    n_samples = [[480000], [16000, 200000, 90000], [333333, 480000]]
    rows = [[M.encoder_frames_for_samples(n) for n in s] for s in n_samples]
    enc_rows = [[torch.randn(1500, d, generator=g) for _ in s] for s in n_samples]       # per clip [1500, d], first a rows kept
    proj = [[e[:a] for e, a in zip(es, rs)] for es, rs in zip(enc_rows, rows)]
    ref, ref_mask, ref_lab = O.combine_ragged(E, ids, mask, labels, proj, vocab - 2, vocab - 1)
    span_off, text_off, total, S = O.ragged_layout(rows, T)
    flat = torch.stack([e for es in enc_rows for e in es]).cuda()                        # [n_clips, 1500, d]
    out, m, lab, starts = splice_ragged(E.cuda(), ids.cuda(), mask.cuda(), labels.cuda(), flat, rows, vocab - 2, vocab - 1)
    assert out.shape == ref.shape
    assert torch.equal(out.cpu(), ref) and torch.equal(m.cpu(), ref_mask) and torch.equal(lab.cpu(), ref_lab)
    for b, offs in enumerate(span_off):
        assert starts[b, :len(offs)].cpu().tolist() == offs                              # int32 prefix sums, bit-exact
    # degenerate case reproduces S1/S2 byte for byte
    o1, m1, l1, _ = splice_ragged(E.cuda(), ids.cuda(), mask.cuda(), labels.cuda(), flat[:3], [[1500]] * 3, vocab - 2, vocab - 1)
    s1 = O.combine(E, ids, flat[:3].cpu(), vocab - 2, vocab - 1)
    assert torch.equal(o1.cpu(), s1) and torch.equal(m1.cpu(), O.extend_mask(mask, 1500))


def test_pipeline_ragged_config5():
    """Config 5 end to end (variable-length clips, 1-3 spans per sample) against the oracle chain."""
    cfg = EncoderConfig(d_model=128, n_layers=2, n_heads=2, ffn_dim=256, n_mels=80)
    d_l, vocab, T = 64, 100, 10
    ew = synth.init_encoder_weights(cfg, seed=0, ln_jitter=0.1)
    pw = synth.init_projector_weights(cfg.d_model, d_l, seed=1, ln_jitter=0.1)
    table = torch.randn(vocab, d_l, generator=torch.Generator().manual_seed(2))
    spans = [1, 3, 2]
    rng = np.random.default_rng(11)
    lens = [480000] + [int(rng.integers(16000, 480001)) for _ in range(5)]
    clips = [synth.synth_clip(i, n_samples=n) for i, n in enumerate(lens)]
    buf = np.zeros((6, 480000), np.float32)
    for i, c in enumerate(clips):
        buf[i, : len(c)] = c
    ids, mask, labels = synth.synth_text(3, T, vocab)
    cond = AudioConditioner(cfg, ew, pw, table.cuda(), vocab - 2, vocab - 1, max_batch=4)      # 6 clips through a 4-clip plan
    out, m, lab, starts = cond.forward_ragged(torch.from_numpy(buf).cuda(), torch.tensor(lens, dtype=torch.int32).cuda(),
                                              spans, ids.cuda(), mask.cuda(), labels.cuda())
    # oracle chain
    mel = torch.from_numpy(M.log_mel_whisper(clips, cfg.n_mels))
    enc = O.encoder_forward(ew, cfg, mel)
    proj = O.projector_forward(pw, enc)
    per_sample, i = [], 0
    for k in spans:
        per_sample.append([proj[c][: M.encoder_frames_for_samples(lens[c])] for c in range(i, i + k)])
        i += k
    ref, ref_m, ref_lab = O.combine_ragged(table, ids, mask, labels, per_sample, vocab - 2, vocab - 1)
    assert out.shape == ref.shape
    assert torch.equal(m.cpu(), ref_m) and torch.equal(lab.cpu(), ref_lab)                # layout bit-exact
    audio = ref_m.bool() & (ref_lab == -100)
    assert O.rel_l2(out.cpu()[audio], ref[audio]) <= 2e-2                                   # bf16 encoder / projector rows
    text = ref_lab != -100
    assert torch.equal(out.cpu()[text], ref[text])                                          # gathered rows bit-exact
    _, _, total, S = O.ragged_layout([[p.shape[0] for p in ps] for ps in per_sample], T)
    for b in range(3):
        assert (out[b, total[b]:].cpu() == 0).all() and (m[b, total[b]:].cpu() == 0).all()   # right padding rows
    # host-side lengths given: no device read-back, same result
    out2, m2, lab2, starts2 = cond.forward_ragged(torch.from_numpy(buf).cuda(), torch.tensor(lens, dtype=torch.int32).cuda(),
                                                  spans, ids.cuda(), mask.cuda(), labels.cuda(), n_samples_host=lens)
    assert torch.equal(out2, out) and torch.equal(m2, m) and torch.equal(lab2, lab) and torch.equal(starts2, starts)
