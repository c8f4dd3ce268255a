"""CPU: host-side logic — API surface mirrors the reference, LoRA low-rank path == the reference's dense path,
checkpoint key names, sharding, error behaviour, and the world_size-2 gradient exchange over gloo."""
import inspect
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from audio_llama_b200 import parallel, synth
from audio_llama_b200.config import WHISPER_LARGE_V3_TURBO, projector_hidden
from audio_llama_b200.models.lora import LoRALayer, apply_lora_to_llama, lora_forward_hook
from audio_llama_b200.models.projector import AudioProjector


def test_api_surface_matches_reference():
    """Names / signatures of SURVEY.md §8b."""
    from audio_llama_b200.models import allm, base, lora, projector
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(allm.AudioLLM.__init__) == ["self", "llama_path", "whisper_path", "lora_rank"]
    assert inspect.signature(allm.AudioLLM.__init__).parameters["lora_rank"].default == 64
    assert sig(allm.AudioLLM.forward)[:5] == ["self", "input_ids", "attention_mask", "audio_features", "labels"]
    assert sig(allm.AudioLLM.generate)[:8] == ["self", "input_ids", "attention_mask", "audio_features",
                                               "max_new_tokens", "temperature", "top_p", "do_sample"]
    g = inspect.signature(allm.AudioLLM.generate).parameters
    assert (g["max_new_tokens"].default, g["temperature"].default, g["top_p"].default, g["do_sample"].default) == (256, 0.7, 0.9, True)
    for m in ("_combine_text_and_audio_embeddings", "_extend_attention_mask", "_process_audio_features",
              "get_trainable_params", "to"):
        assert hasattr(allm.AudioLLM, m)
    assert sig(allm.AudioLLM._combine_text_and_audio_embeddings) == ["self", "text_embeddings", "audio_features", "input_ids"]
    assert sig(allm.AudioLLM._extend_attention_mask) == ["self", "attention_mask", "audio_seq_len", "has_special_tokens"]
    assert sig(projector.AudioProjector.__init__) == ["self", "input_dim", "output_dim", "hidden_dim"]
    assert sig(lora.LoRALayer.__init__) == ["self", "in_dim", "out_dim", "rank", "alpha"]
    assert sig(lora.apply_lora_to_llama) == ["llama_model", "rank", "alpha", "target_modules"]
    assert sig(lora.lora_forward_hook) == ["module", "input", "output", "lora_layer"]
    assert sig(base.load_base_models) == ["llama_model_path", "whisper_model_path"]
    assert {"forward", "to"} <= set(dir(base.FrozenModelWrapper))
    w = base.FrozenModelWrapper(nn.Linear(2, 2))
    assert hasattr(w, "model") and not any(q.requires_grad for q in w.model.parameters())


def test_compat_shim_resolves_reference_import_lines():
    """`from models.allm import AudioLLM` etc. (R/src/train.py:13-16) resolve to the B200 modules."""
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "audio_llama_b200", "compat"))
    try:
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        allm = importlib.import_module("models.allm")
        proj = importlib.import_module("models.projector")
        lora = importlib.import_module("models.lora")
        base = importlib.import_module("models.base")
        assert allm.AudioLLM.__module__ == "audio_llama_b200.models.allm"
        assert proj.AudioProjector is AudioProjector and lora.LoRALayer is LoRALayer
        assert callable(base.load_base_models) and callable(lora.lora_forward_hook)
    finally:
        sys.path.pop(0)
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]


def test_projector_state_dict_keys_and_param_count():
    p = AudioProjector(1280, 3072)
    assert sorted(p.state_dict()) == ["layers.0.bias", "layers.0.weight", "layers.2.bias", "layers.2.weight",
                                      "layers.3.bias", "layers.3.weight"]
    assert p.layers[0].out_features == projector_hidden(1280, 3072) == 2176
    assert sum(x.numel() for x in p.parameters()) == 9481344          # /root/reference/src/training.log:243 (part)
    assert sum(x.numel() for x in AudioProjector(1280, 2048).parameters()) == 5545600
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        p(torch.zeros(1, 2, 1280))


def test_trainable_count_readme_config():
    """95 726 720 = projector + LoRA r=64 on q,k,v,gate,up,down x 28 layers (SURVEY.md §4 vi)."""
    shapes = {"q_proj": (3072, 3072), "k_proj": (3072, 1024), "v_proj": (3072, 1024),
              "gate_proj": (3072, 8192), "up_proj": (3072, 8192), "down_proj": (8192, 3072)}
    lora = sum(64 * (i + o) for i, o in shapes.values()) * 28
    assert lora == 86245376 and lora + 9481344 == 95726720


def test_lora_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "reference_modules.npz"))
    lin = nn.Linear(96, 160)
    lora = LoRALayer(96, 160, rank=8, alpha=16)
    with torch.no_grad():
        lin.weight.copy_(torch.from_numpy(g["lora_W"])); lin.bias.copy_(torch.from_numpy(g["lora_b"]))
        lora.lora_A.copy_(torch.from_numpy(g["lora_A"])); lora.lora_B.copy_(torch.from_numpy(g["lora_B"]))
    x = torch.from_numpy(g["lora_x"])
    y = lora_forward_hook(lin, (x,), lin(x), lora)
    np.testing.assert_allclose(y.detach().numpy(), g["lora_y"], rtol=0, atol=1e-6)
    assert lora.scaling == 2.0
    # init: A zeros, B ~ N(0, 0.01) -> the update is identically zero at init (lora.py:9-18)
    fresh = LoRALayer(32, 48, rank=4)
    assert (fresh.lora_A == 0).all() and 0.003 < fresh.lora_B.std() < 0.03
    assert (fresh(torch.randn(2, 32)) == 0).all()


def test_apply_lora_targets_by_substring():
    class Blk(nn.Module):
        def __init__(self):
            super().__init__()
            self.q_proj, self.k_proj, self.v_proj, self.o_proj = (nn.Linear(8, 8) for _ in range(4))
            self.gate_proj, self.up_proj, self.down_proj = nn.Linear(8, 16), nn.Linear(8, 16), nn.Linear(16, 8)
            self.norm = nn.LayerNorm(8)
    m = nn.ModuleDict({"layers": nn.ModuleList([Blk(), Blk()])})
    ll = apply_lora_to_llama(m, rank=2)
    assert len(ll) == 12 and all("o_proj" not in k for k in ll)
    assert ll["layers.0.down_proj"].lora_A.shape == (2, 16) and ll["layers.0.down_proj"].lora_B.shape == (8, 2)


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 32, 255, 256):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)


def test_synth_recipe_is_deterministic():
    a, b = synth.synth_clip(3), synth.synth_clip(3)
    assert a.dtype == np.float32 and a.shape == (480000,) and (a == b).all() and np.abs(a).max() <= 1.0
    assert (synth.synth_batch(2, first=3)[0] == a).all()
    w = synth.init_encoder_weights(WHISPER_LARGE_V3_TURBO.__class__(64, 1, 1, 128, 80))
    assert "layers.0.self_attn.k_proj.weight" in w and "layers.0.self_attn.k_proj.bias" not in w


def _ddp_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                   # identical trainable weights on every rank
    proj = nn.Sequential(nn.Linear(6, 5), nn.GELU(), nn.Linear(5, 4), nn.LayerNorm(4))
    lora = LoRALayer(4, 3, rank=2)
    with torch.no_grad():
        lora.lora_A.normal_(0, 0.1)
    params = list(proj.parameters()) + list(lora.parameters())
    bucket = parallel.FlatGradBucket(params)
    g = torch.Generator().manual_seed(123)
    x_all, y_all = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
    lo, hi = parallel.shard_range(8, rank, world)
    bucket.zero()
    loss = ((lora(proj(x_all[lo:hi])) - y_all[lo:hi]) ** 2).mean()
    loss.backward()
    flat = bucket.allreduce_mean().clone()
    if rank == 0:
        # single-process reference on the full batch
        for p in params:
            p.grad = None
        ((lora(proj(x_all)) - y_all) ** 2).mean().backward()
        ref = torch.cat([p.grad.flatten() for p in params])
        out.put((flat, ref))
    dist.barrier()
    dist.destroy_process_group()


def _ddp_overlap_worker(rank, world, port, out):
    """Two steps with the overlapped (chunked, hook-driven) exchange; the second step also checks the re-arming."""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    proj = nn.Sequential(nn.Linear(6, 5), nn.GELU(), nn.Linear(5, 4), nn.LayerNorm(4))
    lora = LoRALayer(4, 3, rank=2)
    unused = nn.Parameter(torch.zeros(3))                  # a trainable parameter that receives no gradient
    with torch.no_grad():
        lora.lora_A.normal_(0, 0.1)
    params = list(proj.parameters()) + list(lora.parameters()) + [unused]
    bucket = parallel.FlatGradBucket(params).arm_overlap(n_chunks=3)
    g = torch.Generator().manual_seed(123)
    x_all, y_all = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
    lo, hi = parallel.shard_range(8, rank, world)
    flats = []
    for step in range(2):
        bucket.zero()
        loss = ((lora(proj(x_all[lo:hi] * (step + 1))) - y_all[lo:hi]) ** 2).mean()
        loss.backward()
        flats.append(bucket.finish_overlap().clone())
    n_chunks = len(bucket._ov["chunks"])
    bucket.disarm_overlap()                                # the reference pass below must not launch collectives
    if rank == 0:
        refs = []
        for step in range(2):
            for p in params:
                p.grad = None
            ((lora(proj(x_all * (step + 1))) - y_all) ** 2).mean().backward()
            refs.append(torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).flatten() for p in params]))
        out.put((flats, refs, n_chunks))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_overlapped_world2_gloo():
    """FlatGradBucket.arm_overlap / finish_overlap: chunk all-reduces launched from the gradient hooks during backward
    give the gradient of the full batch, step after step, including a chunk whose parameter got no gradient."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_ddp_overlap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    flats, refs, n_chunks = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert n_chunks == 3
    for f, r in zip(flats, refs):
        assert torch.allclose(f, r, atol=1e-6)


def test_gradient_allreduce_world2_gloo():
    """One flat-bucket allreduce of projector + LoRA grads over 2 ranks == the gradient of the full batch."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    flat, ref = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert torch.allclose(flat, ref, atol=1e-6)


def test_checkpoint_reference_format_roundtrip(golden_dir, tmp_path):
    """A checkpoint.pt written by the reference's own save_checkpoint loads; ours has the same layout; resume works."""
    import argparse
    from types import SimpleNamespace
    from audio_llama_b200 import checkpoint as C
    ref_file = os.path.join(golden_dir, "reference_checkpoint.pt")
    g = np.load(os.path.join(golden_dir, "reference_checkpoint_io.npz"))

    def fresh():
        torch.manual_seed(99)
        m = SimpleNamespace(projector=nn.Sequential(), lora_layers={})
        m.projector = _RefShapedProjector(16, 24)
        m.lora_layers = {"model.layers.0.self_attn.q_proj": LoRALayer(24, 24, rank=4),
                         "model.layers.0.mlp.down_proj": LoRALayer(48, 24, rank=4)}
        return m

    m = fresh()
    params = list(m.projector.parameters()) + [p for l in m.lora_layers.values() for p in l.parameters()]
    opt = torch.optim.AdamW(params, lr=1e-3)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (1 + s))
    info = C.load_checkpoint(m, ref_file, optimizer=opt, scheduler=sched)
    assert info["step"] == 7 and info["epoch"] == 1 and info["dataset_config"] == {"text_key": "text"}
    assert info["args"]["lora_rank"] == 4
    with torch.no_grad():
        np.testing.assert_allclose(m.projector.layers(torch.from_numpy(g["x"])).numpy(), g["y"], atol=1e-6)
    np.testing.assert_array_equal(m.lora_layers["model.layers.0.self_attn.q_proj"].lora_A.detach().numpy(), g["lora_A0"])
    assert opt.state_dict()["state"] and sched.last_epoch == 1                 # resumed
    # our writer produces the same top-level / nested keys as the reference's file
    args = argparse.Namespace(output_dir=str(tmp_path), learning_rate=1e-3, lora_rank=4)
    mine = C.save_checkpoint(m, opt, sched, 7, 1, args, dataset_config={"text_key": "text"})
    assert mine.endswith(os.path.join("checkpoint-7", "checkpoint.pt"))
    a = torch.load(mine, weights_only=False)
    b = torch.load(ref_file, weights_only=False)
    assert sorted(a) == sorted(b) and sorted(a["model"]) == sorted(b["model"])
    assert sorted(a["model"]["projector"]) == sorted(b["model"]["projector"])
    assert sorted(a["model"]["lora_layers"]) == sorted(b["model"]["lora_layers"])
    for k in b["model"]["projector"]:
        assert torch.equal(a["model"]["projector"][k], b["model"]["projector"][k])
    # flat format (inference.py:62-67) also loads
    flat = str(tmp_path / "flat.pt")
    torch.save(b["model"], flat)
    m2 = fresh()
    assert C.load_checkpoint(m2, flat)["step"] is None
    assert torch.equal(m2.projector.layers[0].weight, m.projector.layers[0].weight)
    assert C.save_checkpoint(m, None, None, 9, 2, args, final=True).endswith(os.path.join("final_checkpoint", "checkpoint.pt"))


class _RefShapedProjector(nn.Module):
    """CPU stand-in with AudioProjector's parameter layout (the real one refuses CPU tensors in forward)."""

    def __init__(self, i, o):
        super().__init__()
        h = (i + o) // 2
        self.layers = nn.Sequential(nn.Linear(i, h), nn.GELU(), nn.Linear(h, o), nn.LayerNorm(o))


def test_causal_only_mask_rule():
    """llama_native.causal_only_mask: only purely right-padded masks may be dropped in favour of causal attention."""
    from audio_llama_b200.llama_native import causal_only_mask
    right = torch.tensor([[1., 1., 1., 0., 0.], [1., 1., 1., 1., 1.]])
    left = torch.tensor([[0., 1., 1., 1., 1.], [1., 1., 1., 1., 1.]])
    hole = torch.tensor([[1., 0., 1., 1., 0.]])
    assert causal_only_mask(right) is None and causal_only_mask(None) is None
    assert causal_only_mask(left) is left and causal_only_mask(hole) is hole
    assert causal_only_mask(torch.ones(3, 1)) is None


def test_tensor_derived_cache_invalidation():
    """Packed LoRA operands / transposed frozen weights are cached per tensor and rebuilt when the tensor is updated in
    place (optimizer step) or replaced."""
    from audio_llama_b200.ops import TensorDerivedCache
    c = TensorDerivedCache()
    w = torch.randn(4, 3)
    calls = []

    def build():
        calls.append(1)
        return w.t().contiguous()
    a = c.get((w,), build)
    assert c.get((w,), build) is a and len(calls) == 1
    w.mul_(2.0)                                   # in-place update bumps the version counter
    b = c.get((w,), build)
    assert len(calls) == 2 and torch.equal(b, w.t())
    assert c.get((w,), build, extra=0.5) is not b and len(calls) == 3      # a different scaling is a different entry


def test_native_llama_wiring_falls_back_off_gpu():
    """llama_native.enable() patches the HF modules; inputs that are not bf16 CUDA tensors must take the modules' own
    forward (same results bit for bit), never a CPU re-implementation."""
    import types
    from transformers import LlamaConfig, LlamaForCausalLM
    from transformers.models.llama import modeling_llama as ML
    from audio_llama_b200 import llama_native as LN
    torch.manual_seed(0)
    m = LlamaForCausalLM(LlamaConfig(vocab_size=50, hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                                     num_attention_heads=4, num_key_value_heads=2)).eval()
    for p in m.parameters():
        p.requires_grad = False
    ids = torch.randint(0, 50, (2, 9))
    with torch.no_grad():
        before = m(input_ids=ids).logits
    fake = types.SimpleNamespace(llama=types.SimpleNamespace(model=m), lora_layers={})
    orig_rope = ML.apply_rotary_pos_emb
    try:
        LN.enable(fake)
        assert ML.apply_rotary_pos_emb is LN.apply_rotary_pos_emb and fake.native_ce and fake.native_causal_only
        with torch.no_grad():
            after = m(input_ids=ids).logits
    finally:
        LN.disable_rope_patch()
    assert ML.apply_rotary_pos_emb is orig_rope
    assert torch.equal(before, after)


def test_pack_lora_operands():
    """ops.pack_lora: A zero-padded to a multiple of 8 rows, scaling folded into B, both bf16 (pure torch: CPU-checkable)."""
    from audio_llama_b200 import ops
    g = torch.Generator().manual_seed(0)
    A, Bm = torch.randn(5, 16, generator=g), torch.randn(24, 5, generator=g)
    a, b = ops.pack_lora(A, Bm, 0.25)
    assert a.shape == (8, 16) and b.shape == (24, 8) and a.dtype == b.dtype == torch.bfloat16
    assert torch.equal(a[:5], A.bfloat16()) and (a[5:] == 0).all()
    assert torch.equal(b[:, :5], (Bm * 0.25).bfloat16()) and (b[:, 5:] == 0).all()
    # the packed operands reproduce the hook's update: (x a^T)(s b)^T == scaling * x (B A)^T
    x = torch.randn(7, 16, generator=g)
    ref = (x @ (Bm @ A).T) * 0.25
    got = (x @ a.float().T) @ b.float().T
    assert torch.allclose(got, ref, rtol=0, atol=0.15)          # bf16 operand rounding


def test_causal_only_mask_needs_ignored_pad_labels():
    """The mask is only dropped when every padded position is out of the loss (labels -100 there): the reference's
    dataset pads labels with token ids (dataset.py:82-92), and a padded query DOES see other keys without the mask."""
    from audio_llama_b200.llama_native import causal_only_mask
    m = torch.tensor([[1, 1, 1, 0, 0], [1, 1, 1, 1, 1]])
    ignored = torch.tensor([[5, 6, 7, -100, -100], [1, 2, 3, 4, 5]])
    padded_with_ids = torch.tensor([[5, 6, 7, 0, 0], [1, 2, 3, 4, 5]])
    assert causal_only_mask(m, ignored) is None
    assert causal_only_mask(m, padded_with_ids) is m
    assert causal_only_mask(m, None) is None


def test_flat_bucket_survives_zero_grad_set_to_none():
    """optimizer.zero_grad() (set_to_none=True, what the reference loop calls, train.py:300) detaches the gradients from
    the bucket; the next allreduce_mean() must gather the fresh gradients back instead of reducing stale zeros."""
    from audio_llama_b200.parallel import FlatGradBucket
    torch.manual_seed(0)
    lin = nn.Linear(4, 3)
    params = list(lin.parameters())
    bucket = FlatGradBucket(params)
    opt = torch.optim.SGD(params, lr=0.1)
    x = torch.randn(5, 4)
    for step in range(3):
        lin(x).pow(2).sum().backward()
        want = torch.cat([p.grad.reshape(-1).clone() for p in params])
        flat = bucket.allreduce_mean()                      # world size 1: gathers, no communication
        assert torch.equal(flat, want), step
        assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() and
                   p.grad.data_ptr() < bucket.flat.data_ptr() + bucket.flat.numel() * 4 for p in params)
        opt.step()
        opt.zero_grad()                                     # default set_to_none=True
        assert all(p.grad is None for p in params)
    # and the documented alternative keeps the views alive
    lin(x).sum().backward()
    bucket.allreduce_mean()
    opt.zero_grad(set_to_none=False)
    assert bucket.rebind() == 0 and float(bucket.flat.abs().sum()) == 0.0


def test_place_audio_rows_autograd():
    """_PlaceAudioRows: in-place row placement whose backward is the slice of the gradient (replaces the reference's
    4-way torch.cat, allm.py:165-170, on the training path)."""
    from audio_llama_b200.models.allm import AudioLLM, _PlaceAudioRows
    torch.manual_seed(1)
    B, A, T, d = 2, 5, 3, 8
    emb = torch.randn(B, A + 2 + T, d)
    proj = torch.randn(B, A, d, dtype=torch.float64, requires_grad=True)
    ref = torch.cat([emb[:, :1], proj.to(emb.dtype), emb[:, A + 1:]], dim=1)
    w = torch.randn_like(ref)
    (g_ref,) = torch.autograd.grad((ref * w).sum(), proj)
    out = AudioLLM._place(emb.clone(), proj, A)
    assert torch.equal(out, ref.detach())
    (g,) = torch.autograd.grad((out * w).sum(), proj)
    assert g.dtype == torch.float64 and torch.equal(g, g_ref)
    # without gradients it is a plain in-place store
    with torch.no_grad():
        e2 = emb.clone()
        assert AudioLLM._place(e2, proj, A) is e2 and torch.equal(e2, ref.detach())


def test_static_attention_plan_reads_nothing_on_the_host():
    """attention_plan under static_attention_plan(): kv_len straight from the mask (the caller vouches for right padding),
    no validity read; outside the context a left-padded mask is refused as before."""
    import torch
    from audio_llama_b200 import llama_native as LN
    right = torch.tensor([[1, 1, 1, 0], [1, 1, 1, 1]], dtype=torch.float32)
    left = torch.tensor([[0, 1, 1, 1], [1, 1, 1, 1]], dtype=torch.float32)
    ok, kv = LN.attention_plan(right)
    assert ok and kv.tolist() == [3, 4]
    ok, kv = LN.attention_plan(left)
    assert not ok and kv is None
    with LN.static_attention_plan():
        ok, kv = LN.attention_plan(right)
        assert ok and kv.dtype == torch.int32 and kv.tolist() == [3, 4]
        assert LN.attention_plan(None) == (True, None)
    assert not LN._STATIC_PLAN["on"]

