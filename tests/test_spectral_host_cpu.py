"""CPU: the product's spectral constants (sparse mel matrix, banded pseudo-inverse, windows) against the
oracle's dense ones, and the kernel algorithm restated in numpy (index math of the warp FFT)."""
import math

import numpy as np
import torch

from common import SPECTRAL
from emu_backend import EmuBackend
from oracle import spectral_ops as osp


def test_constants_match_oracle():
    from gansynth_b200 import spectral_ops as sp
    h = sp.host_constants(16000)
    m, p = osp.mel_constants(1024, 16000)
    assert np.array_equal(h["mel"], m) and np.array_equal(h["pinv"], p)
    assert np.array_equal(h["hann"], osp.hann_window(2048).numpy())
    assert np.array_equal(h["synth_window"], osp.inverse_stft_window(2048, 512).numpy())
    # sparse forms reproduce the dense matrices
    dense = np.zeros_like(m)
    for j in range(1024):
        for i in range(sp.MEL_TAPS):
            if h["mel_w"][i, j] != 0:
                dense[h["mel_k0"][j] + i, j] = h["mel_w"][i, j]
    assert np.array_equal(dense, m)
    band = np.zeros_like(p)
    for k in range(1024):
        c = h["pb_cnt"][k]
        band[h["pb_j0"][k]:h["pb_j0"][k] + c, k] = h["pb_w"][:c, k]
    assert h["band"] <= 48
    dropped = np.abs(p - band)
    assert dropped.max() <= sp.PINV_REL_THRESHOLD * np.abs(p).max()
    assert dropped.sum(0).max() < 1e-6      # L1 mass dropped per linear bin: phase error < 1e-6 * |phase|


def test_sparse_spectral_path_matches_oracle_on_cpu():
    """Emulated kernels (same sparse constants, IF from wrapped differences instead of the cumsum
    formulation) against the oracle."""
    from gansynth_b200 import spectral_ops as sp
    h = sp.host_constants(16000)
    consts = {k: torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v for k, v in h.items()}
    g = torch.Generator().manual_seed(0)
    t = torch.arange(64000) / 16000.0
    w = torch.stack([0.3 * torch.sin(2 * math.pi * 440 * t) * torch.exp(-3 * t) + 0.01 * torch.randn(64000, generator=g),
                     0.1 * torch.randn(64000, generator=g)])
    lm, inst = EmuBackend().spectrogram_fwd(w, consts, 128, 16)
    olm, oinst = osp.convert_to_spectrogram(w, **SPECTRAL)
    assert float((lm - olm).abs().max()) < 1e-4
    d = (inst - oinst).abs() % 2.0
    d = torch.minimum(d, 2.0 - d)
    assert float((d > 1e-3).float().mean()) < 1e-4
    back = EmuBackend().waveform_fwd(olm, oinst, consts, 64000)
    want = osp.convert_to_waveform(olm, oinst, **SPECTRAL)
    assert float((back - want).abs().max() / want.abs().max()) < 1e-3


def _brev5(k):
    return ((k & 1) << 4) | ((k & 2) << 2) | (k & 4) | ((k & 8) >> 2) | ((k & 16) >> 4)


def _fft32(re, im):
    r = np.float32(0.70710678118654752)
    for lg in range(5):
        s = 16 >> lg
        for i in range(32):
            if (i & s) == 0:
                j, q = i + s, (i & (s - 1)) * (16 // s)
                ar, ai, br, bi = re[i].copy(), im[i].copy(), re[j].copy(), im[j].copy()
                re[i], im[i] = ar + br, ai + bi
                dr, di = ar - br, ai - bi
                if q == 0:
                    re[j], im[j] = dr, di
                elif q == 8:
                    re[j], im[j] = di, -dr
                elif q == 4:
                    re[j], im[j] = (dr + di) * r, (di - dr) * r
                elif q == 12:
                    re[j], im[j] = (di - dr) * r, -(dr + di) * r
                else:
                    c, sn = np.float32(math.cos(2 * math.pi * q / 32)), np.float32(math.sin(2 * math.pi * q / 32))
                    re[j], im[j] = dr * c + di * sn, di * c - dr * sn


def _warp_fft1024(re, im):
    tw = np.exp(-2j * np.pi * np.outer(np.arange(32), np.arange(32)) / 1024).astype(np.complex64)
    _fft32(re, im)
    br, bi = np.zeros((32, 32), np.float32), np.zeros((32, 32), np.float32)
    for k1 in range(32):
        yr, yi = re[_brev5(k1)], im[_brev5(k1)]
        br[k1], bi[k1] = yr * tw[k1].real - yi * tw[k1].imag, yr * tw[k1].imag + yi * tw[k1].real
    for n2 in range(32):
        re[n2], im[n2] = br[:, n2], bi[:, n2]
    _fft32(re, im)


def test_warp_fft_algorithm_restated_in_numpy():
    """The register/lane index math of csrc/spectral.cu (radix-32 x radix-32 with a transposing
    exchange, real-FFT split step, inverse by re/im swap) reproduces numpy's rfft / irfft."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal(2048).astype(np.float32)
    tw2 = np.exp(-2j * np.pi * np.arange(1024) / 2048).astype(np.complex64)
    re, im = np.zeros((32, 32), np.float32), np.zeros((32, 32), np.float32)
    for n1 in range(32):
        re[n1], im[n1] = x[64 * n1 + 2 * np.arange(32)], x[64 * n1 + 2 * np.arange(32) + 1]
    _warp_fft1024(re, im)
    z = np.zeros(1024, np.complex64)
    for k2 in range(32):
        z[np.arange(32) + 32 * k2] = re[_brev5(k2)] + 1j * im[_brev5(k2)]
    big = np.zeros(1025, np.complex64)
    for k in range(513):
        zc = np.conj(z[(1024 - k) & 1023])
        e, o = 0.5 * (z[k] + zc), -0.5j * (z[k] - zc)
        t = tw2[k] * o
        big[k], big[1024 - k] = e + t, np.conj(e - t)
    ref = np.fft.rfft(x.astype(np.float64))
    assert np.abs(big - ref).max() < 1e-6 * np.abs(ref).max()
    # inverse
    spec = ref.astype(np.complex64)
    spec[0] = 0
    for n1 in range(32):
        for lane in range(32):
            k = 32 * n1 + lane
            a = spec[k]
            c = spec[1024 - k] if k else complex(spec[1024].real, 0.0)
            e, d = 0.5 * (a + np.conj(c)), 0.5 * (a - np.conj(c))
            o = np.conj(tw2[k & 1023]) * d
            zz = e + 1j * o
            re[n1, lane], im[n1, lane] = zz.real, zz.imag
    _warp_fft1024(im, re)
    y = np.zeros(2048, np.float32)
    for k2 in range(32):
        m = np.arange(32) + 32 * k2
        y[2 * m], y[2 * m + 1] = re[_brev5(k2)] / 1024, im[_brev5(k2)] / 1024
    want = np.fft.irfft(np.concatenate([[0], ref[1:]]), n=2048)
    assert np.abs(y - want).max() < 1e-6 * np.abs(want).max()
