"""GPU: spectral kernels against the oracle (reference spectral_ops.py:45-149 restatement).

Tolerances (fp32): log-mel magnitude 1e-3 absolute on the normalised scale (the reference's output
units); instantaneous frequency compared modulo 2 (units of pi: a 1-ulp change next to the +-pi branch
cut of atan2 / floor-mod legitimately moves a value by 2) with 1e-3 on all but a 1e-4 fraction of
ill-conditioned bins (phase of a near-zero STFT bin is noise in any fp32 implementation); waveform 1e-3
of the clip's peak."""
import math

import pytest
import torch

from common import SPECTRAL

pytestmark = pytest.mark.gpu


def _signals(batch, seed=0):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(64000) / 16000.0
    out = []
    for b in range(batch):
        f0 = 110.0 * 2 ** (b % 5)
        tone = sum(torch.sin(2 * math.pi * f0 * h * t) / h for h in range(1, 6)) * torch.exp(-2.0 * t) * 0.3
        out.append(tone + 0.01 * torch.randn(64000, generator=g) if b % 2 == 0 else 0.1 * torch.randn(64000, generator=g))
    return torch.stack(out)


def _if_diff(a, b):
    d = (a - b).abs() % 2.0
    return torch.minimum(d, 2.0 - d)


def test_spectrogram_matches_oracle():
    from gansynth_b200 import spectral_ops as sp
    from oracle import spectral_ops as osp
    w = _signals(6)
    lm, inst = sp.convert_to_spectrogram(w.cuda(), **SPECTRAL)
    olm, oinst = osp.convert_to_spectrogram(w, **SPECTRAL)
    assert lm.shape == (6, 128, 1024) and inst.shape == (6, 128, 1024)
    assert float((lm.cpu() - olm).abs().max()) < 1e-3
    d = _if_diff(inst.cpu(), oinst)
    assert float((d > 1e-3).float().mean()) < 1e-4
    assert float(d.median()) < 1e-5


def test_spectrogram_chunking_is_invariant():
    from gansynth_b200 import functional as F
    from gansynth_b200 import spectral_ops as sp
    w = _signals(2, seed=3).cuda()
    consts = sp.device_constants(16000, w.device)
    ref = F.K.spectrogram_fwd(w, consts, 128, 128)
    for L in (1, 4, 16, 64):
        lm, inst = F.K.spectrogram_fwd(w, consts, 128, L)
        assert torch.equal(lm, ref[0])
        assert float(_if_diff(inst, ref[1]).max()) < 1e-6


def test_waveform_segmentation_is_invariant():
    """Cutting a clip into time segments (phase prefix + atomic overlap exchange) only re-associates the
    overlap-add of the 1536 boundary samples; repeated runs are bit-identical (two addends per sample)."""
    from gansynth_b200 import functional as F
    from gansynth_b200 import spectral_ops as sp
    g = torch.Generator().manual_seed(11)
    lm = (torch.rand(3, 128, 1024, generator=g) * 1.6 - 1.0).cuda()
    inst = (torch.randn(3, 128, 1024, generator=g) * 0.3).cuda()
    consts = sp.device_constants(16000, lm.device)
    ref = F.K.waveform_fwd(lm, inst, consts, 64000)
    peak = float(ref.abs().max())
    for fps in (8, 16, 24, 64):
        got = F.K.waveform_fwd(lm, inst, consts, 64000, fps)
        assert float((got - ref).abs().max()) <= 2e-6 * peak, fps
        assert torch.equal(got, F.K.waveform_fwd(lm, inst, consts, 64000, fps))


def test_ragged_time_steps_match_oracle():
    """T = 20 frames (not a multiple of the 16-frame rounds / 8-frame groups), 10000 samples: partial rounds,
    a 4-frame last run and a 4-frame last segment."""
    from gansynth_b200 import functional as F
    from gansynth_b200 import spectral_ops as sp
    from oracle import spectral_ops as osp
    cfg = dict(waveform_length=10000, sample_rate=16000, spectrogram_shape=[20, 1024], overlap=0.75)
    w = _signals(3, seed=7)[:, :10000].contiguous()
    lm, inst = sp.convert_to_spectrogram(w.cuda(), **cfg)
    olm, oinst = osp.convert_to_spectrogram(w, **cfg)
    assert lm.shape == (3, 20, 1024)
    assert float((lm.cpu() - olm).abs().max()) < 1e-3
    d = _if_diff(inst.cpu(), oinst)
    assert float((d > 1e-3).float().mean()) < 1e-3 and float(d.median()) < 1e-5
    want = osp.convert_to_waveform(olm, oinst, **cfg)
    peak = want.abs().amax(dim=1, keepdim=True)
    got = sp.convert_to_waveform(olm.cuda(), oinst.cuda(), **cfg).cpu()
    assert float(((got - want).abs() / peak).max()) < 1e-3
    consts = sp.device_constants(16000, torch.device("cuda"))
    whole = F.K.waveform_fwd(olm.cuda(), oinst.cuda(), consts, 10000).cpu()
    assert float(((whole - want).abs() / peak).max()) < 1e-3


def test_silence_known_answer():
    """silence -> every magnitude channel (ln 1e-6 + 3.76)/10.05 and IF 0."""
    from gansynth_b200 import spectral_ops as sp
    lm, inst = sp.convert_to_spectrogram(torch.zeros(2, 64000, device="cuda"), **SPECTRAL)
    assert float((lm - (math.log(1e-6) + 3.76) / 10.05).abs().max()) < 1e-6
    assert float(inst.abs().max()) == 0.0


def test_waveform_matches_oracle():
    from gansynth_b200 import spectral_ops as sp
    from oracle import spectral_ops as osp
    w = _signals(4, seed=1)
    olm, oinst = osp.convert_to_spectrogram(w, **SPECTRAL)
    want = osp.convert_to_waveform(olm, oinst, **SPECTRAL)
    got = sp.convert_to_waveform(olm.cuda(), oinst.cuda(), **SPECTRAL).cpu()
    assert got.shape == (4, 64000)
    peak = want.abs().amax(dim=1, keepdim=True)
    assert float(((got - want).abs() / peak).max()) < 1e-3
    # generator-like input: tanh-range random images
    g = torch.Generator().manual_seed(5)
    lm2 = torch.rand(2, 128, 1024, generator=g) * 1.6 - 1.0
    if2 = torch.randn(2, 128, 1024, generator=g) * 0.3
    want = osp.convert_to_waveform(lm2, if2, **SPECTRAL)
    got = sp.convert_to_waveform(lm2.cuda(), if2.cuda(), **SPECTRAL).cpu()
    peak = want.abs().amax(dim=1, keepdim=True)
    assert float(((got - want).abs() / peak).max()) < 1e-3


def test_round_trip_property_full_batch():
    """BASELINE config 3 size (batch 256): size-independent properties instead of the oracle --
    batch-order invariance (clip b of the big batch == the same clip processed alone), finite outputs,
    constant mel columns."""
    from gansynth_b200 import spectral_ops as sp
    g = torch.Generator().manual_seed(0)
    w = (0.1 * torch.randn(256, 64000, generator=g)).cuda()
    lm, inst = sp.convert_to_spectrogram(w, **SPECTRAL)
    back = sp.convert_to_waveform(lm, inst, **SPECTRAL)
    assert back.shape == (256, 64000) and bool(torch.isfinite(back).all())
    lm1, inst1 = sp.convert_to_spectrogram(w[17:18].contiguous(), **SPECTRAL)
    assert torch.equal(lm1[0], lm[17])
    assert float(_if_diff(inst1[0], inst[17]).max()) < 1e-6
    # a single clip is cut into time segments (one CTA each): same samples up to the association of the overlap-add
    back1 = sp.convert_to_waveform(lm[17:18].contiguous(), inst[17:18].contiguous(), **SPECTRAL)
    assert float((back1[0] - back[17]).abs().max()) <= 1e-6 * float(back[17].abs().max())
    # the 107 all-zero mel columns are constant whatever the input
    zero_cols = torch.from_numpy((sp.host_constants(16000)["mel"] != 0).sum(0) == 0).cuda()
    assert int(zero_cols.sum()) == 107
    assert float((lm[..., zero_cols] - (math.log(1e-6) + 3.76) / 10.05).abs().max()) < 1e-6
    assert float(inst[..., zero_cols].abs().max()) == 0.0


def test_empty_batch_and_argument_errors():
    """Edge cases of the C ABI: batch 0 is a no-op, inconsistent framing is refused with a message."""
    from gansynth_b200 import _lib
    from gansynth_b200 import functional as F
    from gansynth_b200 import spectral_ops as sp
    consts = sp.device_constants(16000, torch.device("cuda"))
    lm, inst = F.K.spectrogram_fwd(torch.zeros(0, 64000, device="cuda"), consts, 128, 32)
    assert lm.shape == (0, 128, 1024) and inst.shape == (0, 128, 1024)
    assert F.K.waveform_fwd(lm, inst, consts, 64000).shape == (0, 64000)
    with pytest.raises(_lib.GansynthLibraryError):
        F.K.spectrogram_fwd(torch.zeros(1, 70000, device="cuda"), consts, 128, 32)      # 128 frames cover 67072 samples
    with pytest.raises(_lib.GansynthLibraryError):
        F.K.waveform_fwd(torch.zeros(1, 128, 1024, device="cuda"), torch.zeros(1, 128, 1024, device="cuda"), consts, 64000, 12)
    with pytest.raises(NotImplementedError):      # frame step 2048 * (1 - 0.3) does not divide the frame
        sp.convert_to_spectrogram(torch.zeros(1, 64000, device="cuda"), 64000, 16000, [128, 1024], 0.3)


GENERIC_CONFIGS = [
    # waveform_length, spectrogram_shape, overlap          (BASELINE config 1: [16, 16] at 75 %: frame 32, hop 8, 152 samples)
    (152, [16, 16], 0.75), (140, [16, 16], 0.75), (800, [12, 64], 0.5), (1400, [20, 128], 0.75), (4500, [8, 512], 0.5),
]


@pytest.mark.parametrize("wave_len,shape,overlap", GENERIC_CONFIGS)
def test_generic_configurations_match_oracle(wave_len, shape, overlap):
    """spectral_ops.py:45-149 is generic in spectrogram_shape / overlap: every configuration other than the production
    one (1024 bins, 75 %) runs on the generic kernels (csrc/spectral_generic.cu) -- forward and inverse against the
    oracle, same tolerances as the fast kernels."""
    from gansynth_b200 import spectral_ops as sp
    from oracle import spectral_ops as osp
    cfg = dict(waveform_length=wave_len, sample_rate=16000, spectrogram_shape=shape, overlap=overlap)
    g = torch.Generator().manual_seed(3)
    t = torch.arange(wave_len) / 16000.0
    w = torch.stack([0.1 * torch.randn(wave_len, generator=g), 0.3 * torch.sin(2 * math.pi * 1500.0 * t) + 0.01 * torch.randn(wave_len, generator=g),
                     torch.zeros(wave_len)])
    lm, inst = sp.convert_to_spectrogram(w.cuda(), **cfg)
    olm, oinst = osp.convert_to_spectrogram(w, **cfg)
    assert tuple(lm.shape) == (3, shape[0], shape[1])
    assert float((lm.cpu() - olm).abs().max()) < 1e-3
    d = _if_diff(inst.cpu(), oinst)
    assert float((d > 1e-3).float().mean()) < 2e-3, float((d > 1e-3).float().mean())
    assert float(d[2].max()) == 0.0                                  # silence: phase 0 everywhere
    back = sp.convert_to_waveform(olm.cuda(), oinst.cuda(), **cfg).cpu()
    oback = osp.convert_to_waveform(olm, oinst, **cfg)
    assert back.shape == (3, wave_len)
    assert float((back - oback).abs().max()) < 1e-3 * max(1e-3, float(oback.abs().max()))
