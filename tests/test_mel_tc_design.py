"""CPU check of the ARITHMETIC of the tensor-core log-mel kernel (csrc/mel_tc.cu), restated in numpy by
tests/proto_mel_tc.py: two-level folded DFT-400, fp16 hi + lo operands, three products with fp32 accumulation, banded
mel projection. It pins the design (operand layout, the 0.5 rows, the per-slot power-of-two scale, the stream table of
the filter bank) against the oracle without a GPU; the kernel itself is tested in test_gpu_mel.py."""
import numpy as np
import pytest

from oracle import mel as M
from golden_signals import kat_signals
import proto_mel_tc as P


@pytest.mark.parametrize("name", ["noise0", "synth0", "sine440"])
def test_folded_split_dft_matches_oracle(name):
    x = kat_signals()[name]
    sel = np.arange(3, 3000, 97)
    mel, power = P.mel_tc(x, frames_sel=sel)
    p64 = M.stft_power(M.pad_or_trim(x), np.float64)[:, sel]
    assert np.abs(power - p64).max() / p64.max() <= 2e-6
    mel64 = M.mel_filter_bank_slaney(128).T @ p64
    full = M.log_mel_whisper([x], 128, dtype=np.float64)[0]
    floor = (full.max() * 4 - 4) - 8.0
    a = np.maximum(np.log10(np.maximum(mel.astype(np.float64), 1e-10)), floor)
    b = np.maximum(np.log10(np.maximum(mel64, 1e-10)), floor)
    assert (np.abs(a - b) / 4).max() <= 1e-5


@pytest.mark.parametrize("n_mels,mode", [(128, 0), (80, 0), (128, 1)])
def test_banks_are_banded(n_mels, mode):
    """Every frequency bin of the reference's three banks feeds at most two consecutive mel bins (what the kernel's
    compile-time structure tables and the prototype's stream table rely on)."""
    fb = (M.mel_filter_bank_slaney(n_mels) if mode == 0 else M.mel_filter_bank_htk(n_mels)).astype(np.float32)
    assert P.mel_stream_table(fb) is not None
    nz = [np.nonzero(fb[k])[0] for k in range(201)]
    assert all(len(z) <= 2 and (len(z) < 2 or z[1] == z[0] + 1) for z in nz)


def test_twiddle_rows_carry_the_half():
    tw = P.twiddle_mats()
    assert tw.shape == (4, 112, 112)
    assert np.allclose(tw[0][:101, 0], 0.5) and np.allclose(tw[2][:100, 0], 0.5)       # i = 0: x[200] counted twice
    assert np.all(tw[:, :, 101:] == 0) and np.all(tw[0][101:] == 0) and np.all(tw[2][100:] == 0)
