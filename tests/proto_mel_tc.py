#!/usr/bin/env python
"""Numerical prototype (numpy, CPU) of the tensor-core log-mel kernel (csrc/mel_tc.cu): two-level folded real DFT-400
as four ~100x100 products, operands split into fp16 hi + lo, three products (hi*hi, lo*hi, hi*lo) accumulated in
fp32, then the streaming banded mel projection. Run it to see the error of this arithmetic against the float64 value
of the reference formula (oracle.mel) before spending GPU time; it also documents the operand layout the kernel uses.

Not product code and not imported by the package: a design check that lives with the tests (it uses the oracle as its
checker); tests/test_mel_tc_design.py runs a reduced version on CPU."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mel as M  # noqa: E402

KP = 112     # padded contraction length (K index i = 0..100 used)
NP = 112     # padded output count per parity (even k: 101 used, odd k: 100 used)


def twiddle_mats():
    """B operands [4][NP][KP] float64: Ce, Se (even k = 2 jn), Co, So (odd k = 2 jn + 1). Rows i = 0 and i = 100 of
    the contraction carry the factor 0.5 (the folded sequence counts x[200] and x[100] +- x[300] twice)."""
    i = np.arange(KP)[None, :].astype(np.float64)
    jn = np.arange(NP)[:, None].astype(np.float64)
    half = np.ones(KP)
    half[0] = 0.5
    half[100] = 0.5
    use_i = (np.arange(KP) <= 100)[None, :]
    out = []
    for parity, fn, nvalid in ((0, np.cos, 101), (0, np.sin, 101), (1, np.cos, 100), (1, np.sin, 100)):
        k = 2 * jn + parity
        m = fn(2.0 * np.pi * i * k / 400.0) * half[None, :]
        m = np.where(use_i & (np.arange(NP)[:, None] < nvalid), m, 0.0)
        out.append(m)
    return np.stack(out)


def split_f16_trunc(v):
    """hi = v with the low 13 mantissa bits cleared (exact in fp16 when normal), lo = fp16(v - hi)."""
    v = v.astype(np.float32)
    hi32 = (v.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    hi = hi32.astype(np.float16)
    lo = (v - hi32).astype(np.float16)
    return hi, lo


def split_f16_round(v):
    hi = v.astype(np.float16)
    lo = (v - hi.astype(np.float64)).astype(np.float16)
    return hi, lo


def folded_operands(frames, scale):
    """frames [F, 400] float32 (unwindowed) -> ce, se, co, so [F, KP] float32, scaled."""
    w = M.hann_periodic(400, np.float64)
    x = frames.astype(np.float32)
    F = x.shape[0]
    out = np.zeros((4, F, KP), np.float32)
    for i in range(101):
        ws = np.float32(np.float32(w[i]) * np.float32(scale))
        ws2 = np.float32(np.float32(w[200 - i]) * np.float32(scale))
        a = x[:, i] + x[:, 400 - i] if i > 0 else x[:, 0] + 0 * x[:, 0]      # (kernel reads x[400]: w[0] = 0 kills it)
        c = x[:, i] - x[:, 400 - i] if i > 0 else x[:, 0]
        b = x[:, 200 - i] + x[:, 200 + i]
        d = x[:, 200 - i] - x[:, 200 + i]
        r = ws2 * b
        s = ws2 * d
        out[0, :, i] = ws * a + r       # ce
        out[1, :, i] = ws * c - s       # se
        out[2, :, i] = ws * a - r       # co
        out[3, :, i] = ws * c + s       # so
    return out


def mel_stream_table(fb):
    """fb [201, n_mels] float32 -> per-bin (ml, w_lo, w_hi): bin k adds w_lo*P to mel ml and w_hi*P to mel ml+1.
    Returns None when the bank is not banded that way."""
    n_mels = fb.shape[1]
    ml = np.zeros(201, np.int32)
    wl = np.zeros(201, np.float32)
    wh = np.zeros(201, np.float32)
    prev = -1
    for k in range(201):
        nz = np.nonzero(fb[k])[0]
        if len(nz) == 0:
            ml[k] = prev
        elif len(nz) == 1:
            m = int(nz[0])
            # keep ml non-decreasing: attach as the upper mel of (m-1) when possible
            if m - 1 >= prev:
                ml[k] = m - 1
                wh[k] = fb[k, m]
            elif m >= prev:
                ml[k] = m
                wl[k] = fb[k, m]
            else:
                return None
        elif len(nz) == 2 and nz[1] == nz[0] + 1 and nz[0] >= prev:
            ml[k] = nz[0]
            wl[k] = fb[k, nz[0]]
            wh[k] = fb[k, nz[1]]
        else:
            return None
        prev = ml[k]
    return ml, wl, wh


def mel_stream(power, table, n_mels):
    """power [F, 201] float32 -> mel [F, n_mels] float32 with the kernel's two running accumulators."""
    ml, wl, wh = table
    F = power.shape[0]
    out = np.zeros((F, n_mels), np.float32)
    j = 0
    A = np.zeros(F, np.float32)
    Bc = np.zeros(F, np.float32)

    def emit(m, v):
        if 0 <= m < n_mels:
            out[:, m] = v

    for k in range(201):
        d = int(ml[k]) + 1 - j
        assert d >= 0
        if d == 1:
            emit(j - 1, A)
            A, Bc = Bc, np.zeros(F, np.float32)
            j += 1
        elif d >= 2:
            emit(j - 1, A)
            emit(j, Bc)
            A = np.zeros(F, np.float32)
            Bc = np.zeros(F, np.float32)
            j += d
        A = (wl[k] * power[:, k] + A).astype(np.float32)
        Bc = (wh[k] * power[:, k] + Bc).astype(np.float32)
    emit(j - 1, A)
    emit(j, Bc)
    return out


def mel_tc(wave, n_mels=128, mode=0, split=split_f16_trunc, frames_sel=None):
    x = M.pad_or_trim(np.asarray(wave, np.float32))
    xp = np.pad(x, (200, 200), mode="reflect")
    idx = np.arange(401)[None, :] + 160 * np.arange(3000)[:, None]     # x[400] is read (times w[0] = 0)
    xp = np.concatenate([xp, np.zeros(8, np.float32)])
    if frames_sel is not None:
        idx = idx[frames_sel]
    fr = xp[idx]
    # per 32-frame slot scale (a power of two that brings 4*max|x| under 2^15)
    F = fr.shape[0]
    mx = np.abs(fr).reshape(-1, 1).max() if F else 0.0
    ex = ((np.float32(mx).view(np.uint32) >> 23) & 0xFF).astype(np.int64)
    sh = int(np.clip(139 - ex, -60, 60)) if mx > 0 else 0
    scale = 2.0 ** sh
    ops = folded_operands(fr, scale)
    tw = twiddle_mats()
    D = np.zeros((4, F, NP), np.float32)
    for g in range(4):
        ah, al = split(ops[g])
        bh, bl = split_f16_round(tw[g])
        acc = ah.astype(np.float32) @ bh.astype(np.float32).T
        acc = acc + al.astype(np.float32) @ bh.astype(np.float32).T
        acc = acc + ah.astype(np.float32) @ bl.astype(np.float32).T
        D[g] = acc
    power = np.zeros((F, 201), np.float32)
    power[:, 0::2] = (D[0] ** 2 + D[1] ** 2)[:, :101]
    power[:, 1::2] = (D[2] ** 2 + D[3] ** 2)[:, :100]
    fb = (M.mel_filter_bank_slaney(n_mels) if mode == 0 else M.mel_filter_bank_htk(n_mels)).astype(np.float32)
    table = mel_stream_table(fb)
    assert table is not None
    mel = mel_stream(power, table, n_mels) * np.float32(2.0 ** (-2 * sh))
    return mel.T, power.T * np.float32(2.0 ** (-2 * sh))


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from golden_signals import kat_signals
    sig = kat_signals()
    sel = np.arange(0, 3000, 7)
    for name in ("noise0", "synth0", "sine440", "noise1_5s"):
        x = sig[name]
        mel, power = mel_tc(x, frames_sel=sel)
        p64 = M.stft_power(M.pad_or_trim(x), np.float64)[:, sel]
        fb = M.mel_filter_bank_slaney(128)
        mel64 = fb.T @ p64
        full = M.log_mel_whisper([x], 128, dtype=np.float64)[0]
        floor = (full.max() * 4 - 4) - 8.0
        l = np.maximum(np.log10(np.maximum(mel.astype(np.float64), 1e-10)), floor)
        l64 = np.maximum(np.log10(np.maximum(mel64, 1e-10)), floor)
        err = np.abs(l - l64) / 4
        rel_p = np.abs(power - p64).max() / p64.max()
        print(f"{name:10s} max |dlogmel|/4 = {err.max():.2e}   frac > 1e-5: {(err > 1e-5).mean():.2e}   "
              f"max |dP|/max P = {rel_p:.2e}")
