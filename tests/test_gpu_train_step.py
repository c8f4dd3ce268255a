"""GPU: the README training step module (audio_llama_b200/train_step.py) — eager vs replayed as a CUDA graph."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from audio_llama_b200 import llama_native, train_step
from audio_llama_b200.config import EncoderConfig


def _run(graph: bool, steps: int):
    ecfg = EncoderConfig(d_model=128, n_layers=1, n_heads=2, ffn_dim=256, n_mels=80)
    dev = torch.device("cuda", 0)
    # head_dim 128 (native attention + the fused decoder layer), 2 query heads per kv head
    train_step.LLAMAS["t128"] = dict(hidden_size=512, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                                     num_key_value_heads=2, vocab_size=320)
    try:
        model = train_step.build_model("t128", ecfg, 2, dev, lora_rank=8)
        ts = train_step.TrainStep(model, ecfg, 2, dev, t_txt=40, graph=graph)
        losses = []
        for _ in range(steps):
            ts.step()
            losses.append(ts.loss)
        return losses, ts
    finally:
        train_step.LLAMAS.pop("t128", None)
        llama_native.disable_rope_patch()


def test_train_step_as_cuda_graph_matches_eager():
    """Zero + forward + backward replayed as ONE CUDA graph (steps 3..): same loss trajectory as the eager step (split-K
    reduce-adds are not bit-reproducible: compared at 1e-3), the graph really is replayed, and the deferred id check runs."""
    eager, _ = _run(False, 6)
    graphed, ts = _run(True, 6)
    assert ts.graph is not None and ts.graph_error is None
    assert all(abs(a - b) <= 1e-3 * abs(a) for a, b in zip(eager, graphed)), (eager, graphed)
    assert graphed[-1] < graphed[0]                         # it trains
