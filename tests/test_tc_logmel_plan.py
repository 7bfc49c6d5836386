"""The arithmetic plan of the tensor-core log-mel (tools/probes/logmel_tc_probe.*), checked on the CPU:
the folded integer operands, the f16-split basis bank in its shared-memory layout and the two-accumulator mel table must
reproduce the float64 oracle before any of it runs on a GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "probes"))
import logmel_tc_probe as P  # noqa: E402


def _bank_planes():
    bank = P.bank_bytes().view(np.uint16)
    k = np.arange(112)[:, None]
    j = np.arange(112)[None, :]
    off = ((k // 8) * 1792 + (j // 8) * 128 + (j % 8) * 16 + (k % 8) * 2) // 2
    planes = np.zeros((4, 2, 112, 112))
    for p in range(4):
        for h in range(2):
            planes[p, h] = bank[(p * 2 + h) * (P.PLANE // 2) + off].view(np.float16).astype(np.float64)
    return planes


def test_basis_bank_layout_and_split():
    planes = _bank_planes()
    B = P.basis_products()
    assert np.abs(planes[:, 0] + planes[:, 1] - B).max() < 2.0 ** -23          # hi + unscaled residual (f16 subnormals)
    assert np.all(planes[:, :, 101:, :] == 0) and np.all(planes[0, :, :, 101:] == 0) and np.all(planes[2, :, :, 100:] == 0)


def test_mel_table_reproduces_the_filterbank():
    F, w0, w1, emit, m_emit, last = P.mel_table(80)
    rng = np.random.default_rng(0)
    pw = rng.random(201)
    out = np.zeros(80)
    a0 = a1 = 0.0
    for k in range(201):
        for r in range(emit[k]):
            out[m_emit[k] + r] = a0
            a0, a1 = a1, 0.0
        a0 += float(w0[k]) * pw[k]
        a1 += float(w1[k]) * pw[k]
    out[last] = a0
    if last + 1 < 80:
        out[last + 1] = a1
    assert emit.max() <= 2 and np.allclose(out, F.astype(np.float64) @ pw, rtol=1e-12, atol=1e-15)


def test_folded_products_give_the_whisper_log_mel():
    rng = np.random.default_rng(1)
    nf = 200
    t = np.arange(160 * nf + 400) / 16000.0
    y = 0.9 * np.sin(2 * np.pi * 1234.5 * t) + rng.normal(0, 1e-4, t.size)
    y[16000:] = rng.normal(0, 0.1, t.size - 16000)
    pcm = np.clip(np.rint(y * 32768), -32768, 32767).astype(np.int16)
    F = P.mel_table(80)[0]
    ref, _ = P.reference_log10_mel(pcm, nf, F)
    planes = _bank_planes()
    x = pcm[np.arange(nf)[:, None] * 160 + np.arange(400)[None, :]].astype(np.int64)
    n = np.arange(112)
    x4 = np.where((400 - n)[None, :] == 400, x[:, [0]], x[:, np.minimum(400 - n, 399)])      # x[400] := x[0]
    A, Bv, Dn, Dr = x[:, n] + x[:, n + 200], x[:, 200 - n] + x4, x[:, n] - x[:, n + 200], x[:, 200 - n] - x4
    acc = []
    for p, v in enumerate((A + Bv, A - Bv, Dn - Dr, Dn + Dr)):
        assert np.abs(v).max() < 1 << 18
        hi = (v + 64) >> 7
        lo = (v - (hi << 7)) / 128.0
        assert np.abs(hi).max() <= 1024 and np.array_equal(hi.astype(np.float16).astype(np.int64), hi)      # exact in f16
        assert np.array_equal(lo.astype(np.float16).astype(np.float64), lo)
        acc.append((hi + lo) @ (planes[p, 0] + planes[p, 1]))
    X = np.zeros((nf, 203), complex)                                   # slot k + 1; slots 0 and 202 are the mirrored bins
    X[:, 1:202:2] = acc[0][:, :101] + 1j * acc[1][:, :101]
    X[:, 2:201:2] = acc[2][:, :100] + 1j * acc[3][:, :100]
    X[:, 0], X[:, 202] = np.conj(X[:, 2]), np.conj(X[:, 200])
    Xw = 0.5 * X[:, 1:202] - 0.25 * (X[:, 0:201] + X[:, 2:203])        # periodic Hann as a 3-tap on the bins
    power = (Xw.real ** 2 + Xw.imag ** 2) * 2.0 ** -16
    got = np.log10(np.maximum(power @ F.astype(np.float64).T, 1e-10)).T
    floor = ref.max() - 8.0
    assert np.abs(np.maximum(got, floor) - np.maximum(ref, floor)).max() / 4.0 < 1e-5       # basis quantisation only; the bar is 1e-4
