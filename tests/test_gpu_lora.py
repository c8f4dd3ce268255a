"""GPU parity: fused frozen-linear + LoRA GEMM (L1) vs the oracle (= the reference's hook, pinned by
tests/golden/reference_modules.npz) on LLaMA-3.2-3B's four LoRA-targeted shapes."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from audio_llama_b200 import ops
from oracle import encoder as O


@pytest.mark.parametrize("in_dim,out_dim,rank,rows", [(3072, 3072, 64, 300), (3072, 1024, 64, 2014), (3072, 8192, 64, 257),
                                                      (8192, 3072, 64, 130), (256, 256, 8, 77)])
def test_lora_linear(in_dim, out_dim, rank, rows):
    g = torch.Generator().manual_seed(in_dim + out_dim)
    x = torch.randn(rows, in_dim, generator=g).bfloat16()
    W = (torch.randn(out_dim, in_dim, generator=g) * 0.02).bfloat16()
    bias = torch.randn(out_dim, generator=g) * 0.1
    A = torch.randn(rank, in_dim, generator=g) * 0.05          # reference init is zeros (update == 0): randomise
    B = torch.randn(out_dim, rank, generator=g) * 0.05
    scaling = 16 / rank
    ref = O.lora_linear(x.float(), W.float(), bias, A, B, scaling)
    base = O.lora_linear(x.float(), W.float(), bias, torch.zeros_like(A), B, scaling)
    y = ops.lora_linear(x.cuda(), W.cuda(), bias.cuda(), A.cuda(), B.cuda(), scaling, out_dtype=torch.float32).cpu()
    assert O.rel_l2(y, ref) <= 1e-2
    # the low-rank part itself (what the hook adds) is reproduced, not just the frozen product
    assert O.rel_l2(y - base, ref - base) <= 3e-2
    # zero A (the reference's init) -> exactly the frozen linear
    y0 = ops.lora_linear(x.cuda(), W.cuda(), bias.cuda(), torch.zeros_like(A).cuda(), B.cuda(), scaling, out_dtype=torch.float32).cpu()
    assert O.rel_l2(y0, base) <= 2e-3


def test_lora_linear_matches_reference_hook_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "reference_modules.npz"))
    t = lambda k: torch.from_numpy(g[k])
    y = ops.lora_linear(t("lora_x").cuda().bfloat16(), t("lora_W").cuda().bfloat16(), t("lora_b").cuda(), t("lora_A").cuda(),
                        t("lora_B").cuda(), float(g["lora_scaling"][0]), out_dtype=torch.float32).cpu()
    assert O.rel_l2(y, t("lora_y")) <= 2e-2
