"""CPU: known-answer and cross-implementation tests that pin the oracle (the reference ships no tests or
golden vectors, SURVEY section 4 -- these replace them) plus the committed golden fixtures."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as TF

from common import HYPER, SMALL, SPECTRAL
from oracle import models as omodels
from oracle import networks as onet
from oracle import ops as oops
from oracle import spectral_ops as osp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ----------------------------------------------------------------------------- ops identities
def test_same_padding_strided_conv_is_asymmetric():
    x = torch.randn(2, 3, 8, 10, dtype=torch.float64)
    w = torch.randn(3, 3, 3, 5, dtype=torch.float64)
    got = oops.conv2d(x, w, None, (2, 2), 2.0)
    c = oops.he_constant(w.shape, 2.0)
    want = TF.conv2d(TF.pad(x, (0, 1, 0, 1)), (w * c).permute(3, 2, 0, 1), stride=2)
    assert torch.allclose(got, want, atol=1e-12)
    wrong = TF.conv2d(x, (w * c).permute(3, 2, 0, 1), stride=2, padding=1)
    assert not torch.allclose(got, wrong, atol=1e-3)


def test_conv2d_transpose_is_gradient_of_strided_conv():
    """tf.nn.conv2d_transpose == conv2d_backprop_input of the SAME strided conv."""
    x = torch.randn(2, 4, 5, 6, dtype=torch.float64)
    var = torch.randn(3, 3, 4, 7, dtype=torch.float64)        # [kh, kw, Cin, filters]
    got = oops.conv2d_transpose(x, var, None, (2, 2), 2.0)
    assert got.shape == (2, 7, 10, 12)
    # forward conv whose filter is var transposed to [kh, kw, filters, Cin]; its input-gradient at x
    y_in = torch.zeros(2, 7, 10, 12, dtype=torch.float64, requires_grad=True)
    c = oops.he_constant(var.shape, 2.0)
    filt = (var * c).permute(0, 1, 3, 2)
    y = TF.conv2d(TF.pad(y_in, (0, 1, 0, 1)), filt.permute(3, 2, 0, 1), stride=2)
    (grad,) = torch.autograd.grad(y, y_in, grad_outputs=x)
    assert torch.allclose(got, grad, atol=1e-12)
    bad = TF.conv_transpose2d(x, (var * c).permute(2, 3, 0, 1), stride=2, padding=1, output_padding=1)
    assert not torch.allclose(got, bad, atol=1e-3)


def test_batch_stddev_grouping():
    x = torch.randn(8, 3, 2, 4, dtype=torch.float64)
    s = oops.batch_stddev(x)
    assert s.shape == (8, 1, 2, 4)
    assert torch.equal(s[0], s[2]) and torch.equal(s[0], s[4]) and torch.equal(s[0], s[6])
    assert not torch.equal(s[0], s[1])
    manual = torch.sqrt(x[[0, 2, 4, 6]].var(0, unbiased=False) + 1e-12).mean()
    assert torch.allclose(s[0, 0, 0, 0], manual)


def test_pixel_norm_and_resampling():
    x = torch.randn(2, 5, 3, 4, dtype=torch.float64)
    y = oops.pixel_normalization(x)
    assert torch.allclose((y * y).mean(1), torch.ones(2, 3, 4, dtype=torch.float64), atol=1e-9)
    up = oops.upscale2d(x, (2, 3))
    assert up.shape == (2, 5, 6, 12) and torch.equal(up[:, :, 1, 2], x[:, :, 0, 0])
    assert torch.allclose(oops.downscale2d(up, (2, 3)), x)
    assert oops.upscale2d(x, (1, 1)) is x


def test_lerp_weights_first_argument():
    assert onet.lerp(torch.tensor(2.0), torch.tensor(10.0), 0.25) == 0.25 * 2.0 + 0.75 * 10.0


def test_growth_schedule_and_param_counts():
    pg = onet.PGGAN([2, 16], [128, 1024], 32, 256, 31.0 / 127.0)
    assert pg.max_depth == 6 and abs(pg.growing_depth - 5.0) < 1e-12
    assert [pg.channels(d) for d in range(7)] == [256, 256, 256, 256, 128, 64, 32]
    g, d = pg.variable_shapes()
    assert sum(int(np.prod(s)) for s, _ in g.values()) == 8932238
    assert sum(int(np.prod(s)) for s, _ in d.values()) == 6830973


def test_tf_adam_epsilon_outside_bias_correction():
    p = {"x/w": torch.tensor([1.0], requires_grad=True)}
    opt = omodels.TFAdam(["x/w"], p, lr=0.1, beta1=0.0, beta2=0.99)
    g = torch.tensor([0.5])
    opt.apply(p, {"x/w": g})
    lr_t = 0.1 * math.sqrt(1 - 0.99)
    want = 1.0 - lr_t * 0.5 / (math.sqrt(0.01 * 0.25) + 1e-8)
    assert abs(float(p["x/w"]) - want) < 1e-7


# ----------------------------------------------------------------------------- spectral KATs
def test_hann_overlap_add_is_1p5():
    w = osp.hann_window(2048, torch.float64)
    s = (w * w).reshape(4, 512).sum(0)
    assert torch.allclose(s, torch.full((512,), 1.5, dtype=torch.float64), atol=1e-12)
    assert torch.allclose(osp.inverse_stft_window(2048, 512, torch.float64), w / 1.5, atol=1e-12)


def test_mel_matrix_structure():
    m, p = osp.mel_constants(1024, 16000)
    assert m.shape == (1024, 1024) and m.dtype == np.float32
    assert int((m != 0).sum()) == 2042 and int((m != 0).sum(1).max()) <= 2 and int((m != 0).sum(0).max()) <= 6
    assert int(((m != 0).sum(0) == 0).sum()) == 107          # all-zero mel columns
    assert np.linalg.matrix_rank(m.astype(np.float64)) == 726
    # pseudo-inverse properties on the kept subspace: M P M == M
    assert np.abs(m.astype(np.float64) @ p.astype(np.float64) @ m.astype(np.float64) - m).max() < 1e-4
    # independent float64 HTK construction agrees to fp32 resolution
    def mel(f):
        return 1127.0 * np.log1p(f / 700.0)
    lin = mel(np.linspace(0.0, 8000.0, 1024)[1:])[:, None]
    e = np.linspace(mel(0.0), mel(8000.0), 1026)
    ref = np.maximum(0, np.minimum((lin - e[None, :-2]) / (e[None, 1:-1] - e[None, :-2]),
                                   (e[None, 2:] - lin) / (e[None, 2:] - e[None, 1:-1])))
    assert np.abs(m[1:] - ref).max() < 5e-4


def test_mel_matrix_against_torchaudio_htk():
    """Third-party corroboration of the HTK filterbank: torchaudio's melscale_fbanks(mel_scale="htk", norm=None)
    builds its triangles in Hz instead of mel (a different but equally standard reading of the same centres), so the
    two agree to ~4e-4 with IDENTICAL supports -- which pins the edge frequencies, the bin-to-frequency mapping
    (linspace(0, nyquist, 1024), the reference's off-by-one quirk included) and the zero first row."""
    torchaudio = pytest.importorskip("torchaudio")
    m, _ = osp.mel_constants(1024, 16000)
    fb = torchaudio.functional.melscale_fbanks(n_freqs=1024, f_min=0.0, f_max=8000.0, n_mels=1024, sample_rate=16000,
                                               norm=None, mel_scale="htk").numpy()
    assert fb.shape == m.shape
    assert np.array_equal(m != 0, fb != 0)
    assert np.abs(m - fb).max() < 1e-3
    from gansynth_b200 import spectral_ops as sp                     # the product's host constants are the same matrix
    assert np.array_equal(sp.host_constants(16000)["mel"], m)


def test_pinv_against_numpy_with_tfp_rcond():
    """tfp.math.pinv(a) = SVD pseudo-inverse with singular values below rcond * s_max dropped, rcond = 10 * max(shape)
    * eps(float32) (App. B.12): numpy's pinv with that rcond is the same definition from an independent code path.
    The cut matters here: the mel matrix has rank 726 of 1024, and 726 is what survives."""
    m, p = osp.mel_constants(1024, 16000)
    rcond = 10.0 * 1024 * np.finfo(np.float32).eps
    ref = np.linalg.pinv(m.astype(np.float64), rcond=rcond)
    assert p.shape == ref.shape == (1024, 1024)
    assert np.abs(p - ref).max() < 2e-3 * np.abs(ref).max()
    s = np.linalg.svd(m.astype(np.float64), compute_uv=False)
    assert int((s > rcond * s.max()).sum()) == 726


def test_inverse_stft_against_torch_istft():
    """Third-party corroboration of tf.signal.inverse_stft as restated (irfft -> window / 1.5 -> overlap-add -> crop of
    the front padding): torch.istft divides by the true overlap-added window envelope, which equals 1.5 wherever four
    frames overlap, so the two agree on every sample except the first and last 1536 of the padded signal."""
    g = torch.Generator().manual_seed(3)
    log_mel = (torch.rand(2, 128, 1024, generator=g, dtype=torch.float64) * 1.6 - 1.0)
    mel_if = torch.randn(2, 128, 1024, generator=g, dtype=torch.float64) * 0.3
    got = osp.convert_to_waveform(log_mel, mel_if, **SPECTRAL)
    p = torch.from_numpy(osp.mel_constants(1024, 16000)[1]).double()
    mag = torch.exp(log_mel * 10.05 - 3.76) @ p
    phase = torch.cumsum(mel_if * math.pi, dim=-2) @ p
    spec = torch.nn.functional.pad(torch.polar(mag, phase), (1, 0))              # [B, T, 1025], DC = 0
    spec[..., -1] = spec[..., -1].real + 0j                                      # irfft ignores Im of the Nyquist bin
    n = 512 * 127 + 2048
    # center=True only trims 1024 samples at each end (the untrimmed form trips torch's check that the window
    # envelope is non-zero: hann[0] = 0): want[i] is padded sample i + 1024
    want = torch.istft(spec.transpose(1, 2), n_fft=2048, hop_length=512, win_length=2048,
                       window=osp.hann_window(2048, torch.float64), center=True)
    pad = n - 64000
    assert got.shape == (2, 64000) and want.shape == (2, n - 2048)
    lo, hi = 1536, n - 1536                                                      # padded samples with four frames
    a = got[:, lo - pad:hi - pad] if lo >= pad else got[:, :hi - pad]            # got[i] is padded sample i + pad
    b = want[:, max(lo, pad) - 1024:hi - 1024]
    assert a.shape == b.shape
    assert float((a - b).abs().max()) < 1e-9 * float(want.abs().max())


def test_stft_against_scipy():
    import scipy.signal
    x = torch.randn(1, 4096 + 2048, dtype=torch.float64)
    got = osp.stft(x, 2048, 512)[0].numpy()
    _, _, z = scipy.signal.stft(x[0].numpy(), window=scipy.signal.get_window("hann", 2048, fftbins=True), nperseg=2048,
                                noverlap=1536, boundary=None, padded=False, return_onesided=True)
    want = z.T * scipy.signal.get_window("hann", 2048, fftbins=True).sum()
    assert np.abs(got - want).max() < 1e-9 * np.abs(want).max()


def test_pure_tone_phase_advance_before_mel():
    """a tone at a bin centre advances 2 pi f hop / sr per frame (mod 2 pi)."""
    k = 100
    f = k * 16000.0 / 2048.0
    t = torch.arange(2048 + 512 * 9, dtype=torch.float64) / 16000.0
    s = osp.stft(torch.sin(2 * math.pi * f * t)[None], 2048, 512)[0]
    ph = torch.angle(s[:, k])
    adv = torch.remainder(ph[1:] - ph[:-1], 2 * math.pi)
    want = (2 * math.pi * f * 512 / 16000.0) % (2 * math.pi)
    assert float((torch.remainder(adv - want + math.pi, 2 * math.pi) - math.pi).abs().max()) < 1e-6


def test_unwrap_matches_numpy():
    ph = torch.from_numpy(np.cumsum(np.random.default_rng(0).uniform(-3, 3, (4, 50)), axis=1))
    wrapped = torch.remainder(ph + math.pi, 2 * math.pi) - math.pi
    got = osp.unwrap(wrapped, axis=-1)
    assert np.abs(got.numpy() - np.unwrap(wrapped.numpy(), axis=-1)).max() < 1e-9


def test_silence_known_answer():
    lm, inst = osp.convert_to_spectrogram(torch.zeros(1, 64000), **SPECTRAL)
    assert float((lm - (math.log(1e-6) + 3.76) / 10.05).abs().max()) < 1e-6
    assert float(inst.abs().max()) == 0.0
    assert lm.shape == (1, 128, 1024)


def test_round_trip_lengths_and_tonal_similarity():
    t = torch.arange(64000) / 16000.0
    w = (torch.sin(2 * math.pi * 2000.0 * t) * torch.exp(-3 * t))[None] * 0.5
    lm, inst = osp.convert_to_spectrogram(w, **SPECTRAL)
    back = osp.convert_to_waveform(lm, inst, **SPECTRAL)
    assert back.shape == (1, 64000)
    # lossy by design (rank-726 projector): 0.87 at 440 Hz ... 0.998 at 2 kHz, cf. SURVEY App. D
    assert float(TF.cosine_similarity(w, back, dim=1)) > 0.99


# ----------------------------------------------------------------------------- step plumbing (BASELINE config 1)
@pytest.mark.parametrize("level", [0.0, 0.3, 1.0])
def test_config1_plumbing(level):
    pg = onet.PGGAN(growing_level=level, **SMALL)
    params = pg.init_variables(seed=3)
    g = torch.Generator().manual_seed(0)
    z = torch.randn(4, 256, generator=g)
    lab = TF.one_hot(torch.randint(0, 61, (4,), generator=g), 61).float()
    step = omodels.GANSynthStep(pg, params, HYPER)
    for _ in range(2):
        d, _ = step.discriminator_update(torch.randn(4, 2, 16, 16, generator=g), lab, z)
        gl, _ = step.generator_update(lab, z)
        assert torch.isfinite(d) and torch.isfinite(gl)
    assert step.global_step == 2


# ----------------------------------------------------------------------------- golden fixtures
def test_golden_fixtures():
    """tests/golden/*.npz were written by tests/tools/make_golden.py from this oracle; they freeze its
    behaviour so later edits cannot drift silently."""
    path = os.path.join(GOLDEN, "small_step.npz")
    if not os.path.exists(path):
        pytest.skip("golden fixtures not generated")
    z = np.load(path)
    pg = onet.PGGAN(growing_level=float(z["level"]), **SMALL)
    params = pg.init_variables(seed=int(z["seed"]), bias_std=0.1)
    img = pg.generator(params, torch.from_numpy(z["latents"]), torch.from_numpy(z["labels"]))
    assert np.abs(img.detach().numpy() - z["fake_images"]).max() < 1e-5
    _, logits = pg.discriminator(params, torch.from_numpy(z["images"]), torch.from_numpy(z["labels"]))
    assert np.abs(logits.detach().numpy() - z["logits"]).max() < 1e-4 * np.abs(z["logits"]).max()
    step = omodels.GANSynthStep(pg, params, HYPER)
    d, _ = step.discriminator_update(torch.from_numpy(z["images"]), torch.from_numpy(z["labels"]), torch.from_numpy(z["latents"]))
    gl, _ = step.generator_update(torch.from_numpy(z["labels"]), torch.from_numpy(z["latents"]))
    assert abs(float(d) - float(z["d_loss"])) < 1e-4 and abs(float(gl) - float(z["g_loss"])) < 1e-4
    s = np.load(os.path.join(GOLDEN, "spectral.npz"))
    lm, inst = osp.convert_to_spectrogram(torch.from_numpy(s["wave"]), **SPECTRAL)
    assert np.abs(lm.numpy()[:, ::16, ::16] - s["logmel_sub"]).max() < 1e-4
    back = osp.convert_to_waveform(lm, inst, **SPECTRAL)
    assert np.abs(back.numpy()[:, ::64] - s["back_sub"]).max() < 1e-4 * np.abs(s["back_sub"]).max()
