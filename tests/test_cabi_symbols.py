"""CPU: the C-ABI library builds, loads, and exports every symbol include/audiollm_b200.h declares; the ctypes
table mirrors the header one to one. No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "audiollm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(al_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    return g.LIB


def test_header_declares_the_path():
    fns = header_functions()
    for need in ("al_mel_forward", "al_encoder_forward", "al_projector_forward", "al_splice", "al_splice_ragged",
                 "al_gemm_bf16", "al_attention", "al_layernorm", "al_last_error", "al_version"):
        assert need in fns


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built)
    for fn in header_functions():
        assert hasattr(lib, fn), f"{fn} declared in the header but not exported"
    lib.al_version.restype = ctypes.c_int
    assert lib.al_version() >= 100


def test_ctypes_table_matches_header(built):
    from audio_llama_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_functions()
    _lib.lib()          # resolves every symbol, raises otherwise


def test_argument_errors_are_reported_not_thrown(built):
    """Error contract: negative return + al_last_error text, no exception / abort (host-only calls)."""
    from audio_llama_b200 import _lib
    L = _lib.lib()
    assert L.al_mel_filterbank_host(0, 0, None) == -1
    assert b"al_mel_filterbank_host" in L.al_last_error()
    h = ctypes.c_void_p()
    assert L.al_encoder_create(ctypes.byref(h), 100, 1, 2, 64, 80, 1, None, 0) == -1      # d != heads * 64
    assert b"head_dim" in L.al_last_error()
    with pytest.raises(_lib.AudioLLMLibError):
        _lib.check(-1, "demo")


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from audio_llama_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.AudioLLMLibError, match="no CPU fallback"):
        _lib.lib()


def test_filterbank_host_matches_oracle(built):
    """The one host-side computation of the library: the mel filter bank (fp64), vs the oracle / HF golden."""
    import numpy as np
    from audio_llama_b200 import ops
    from oracle import mel as M
    for n in (80, 128):
        np.testing.assert_allclose(ops.mel_filterbank(n, 0), M.mel_filter_bank_slaney(n), rtol=1e-12, atol=1e-15)
    assert (ops.mel_filterbank(128, 1) == M.mel_filter_bank_htk(128).astype(np.float64)).all()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under audio_llama_b200/ may import it."""
    pkg = os.path.join(ROOT, "audio_llama_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)
