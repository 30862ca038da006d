"""Oracle restatement of the spectral front-end (reference spectral_ops.py:8-149), PyTorch CPU.
Test infrastructure only (see oracle/__init__.py).

TF-1.13 semantics: periodic Hann, un-normalised rfft of 2048-sample frames at
hop 512, HTK mel matrix built in float32 with one zero leading row, tfp.math.pinv by SVD with
rcond = 10 * max(rows, cols) * eps, inverse_stft window = hann / sum_k hann^2[n + k * step],
floor-mod in unwrap.
"""
import math

import numpy as np
import torch


# ----------------------------------------------------------------------------- constants
def linear_to_mel_weight_matrix(num_mel_bins, num_spectrogram_bins, sample_rate,
                                lower_edge_hertz, upper_edge_hertz, dtype=np.float32):
    """tf.signal.linear_to_mel_weight_matrix as TF 1.13 evaluates it: every step in `dtype`
    (SURVEY App. B-11).  mel(f) = 1127 ln(1 + f / 700); triangular filters on the mel scale; the DC
    row is zero."""
    t = dtype

    def hertz_to_mel(f):
        return (t(1127.0) * np.log(t(1.0) + f / t(700.0))).astype(dtype)

    nyquist = t(sample_rate / 2.0)
    lin = np.linspace(t(0.0), nyquist, num_spectrogram_bins, dtype=dtype)[1:]
    spec_mel = hertz_to_mel(lin)[:, None]
    edges = np.linspace(hertz_to_mel(np.asarray(lower_edge_hertz, dtype)),
                        hertz_to_mel(np.asarray(upper_edge_hertz, dtype)),
                        num_mel_bins + 2, dtype=dtype)
    lower, center, upper = edges[None, :-2], edges[None, 1:-1], edges[None, 2:]
    lower_slopes = (spec_mel - lower) / (center - lower)
    upper_slopes = (upper - spec_mel) / (upper - center)
    weights = np.maximum(t(0.0), np.minimum(lower_slopes, upper_slopes))
    return np.pad(weights, [[1, 0], [0, 0]]).astype(dtype)


def pinv(a, rcond=None):
    """tfp.math.pinv: SVD, singular values <= rcond * max are dropped,
    rcond defaults to 10 * max(rows, cols) * eps(dtype)."""
    a = np.asarray(a)
    if rcond is None:
        rcond = 10.0 * max(a.shape) * np.finfo(a.dtype).eps
    u, s, vt = np.linalg.svd(a, full_matrices=False)
    keep = s > rcond * s.max()
    s_inv = np.where(keep, 1.0 / np.where(keep, s, 1.0), 0.0).astype(a.dtype)
    return ((vt.T * s_inv) @ u.T).astype(a.dtype)


def hann_window(n, dtype=torch.float32):
    """tf.signal.hann_window(periodic=True): 0.5 - 0.5 cos(2 pi k / n)."""
    k = torch.arange(n, dtype=torch.float64)
    return (0.5 - 0.5 * torch.cos(2.0 * math.pi * k / n)).to(dtype)


def inverse_stft_window(frame_length, frame_step, dtype=torch.float32):
    """tf.signal.inverse_stft_window_fn(frame_step, hann): hann / sum over overlapping hops of
    hann^2 -- hann / 1.5 at 75 % overlap."""
    w = hann_window(frame_length, dtype)
    denom = (w * w).reshape(-1, frame_step).sum(0).repeat(frame_length // frame_step)
    return w / denom


def frame_params(waveform_length, spectrogram_shape, overlap):
    """spectral_ops.py:50-53."""
    time_steps, num_freq_bins = spectrogram_shape
    frame_length = num_freq_bins * 2
    frame_step = int((1.0 - overlap) * frame_length)
    num_samples = frame_step * (time_steps - 1) + frame_length
    return time_steps, num_freq_bins, frame_length, frame_step, num_samples


_CONST_CACHE = {}


def mel_constants(num_freq_bins, sample_rate):
    """(linear->mel matrix [bins, mel], its pseudo-inverse [mel, bins]) as float32 numpy arrays.
    spectral_ops.py:76-82 and :115-122 (num_spectrogram_bins = bins AFTER the DC drop: a quirk the
    reference has and this keeps)."""
    key = (num_freq_bins, sample_rate)
    if key not in _CONST_CACHE:
        m = linear_to_mel_weight_matrix(num_freq_bins, num_freq_bins, sample_rate, 0.0, sample_rate / 2.0)
        _CONST_CACHE[key] = (m, pinv(m))
    return _CONST_CACHE[key]


# ----------------------------------------------------------------------------- phase helpers
def diff(x, axis=-1):
    """spectral_ops.py:8-17."""
    n = x.shape[axis]
    return x.narrow(axis, 1, n - 1) - x.narrow(axis, 0, n - 1)


def unwrap(phases, axis=-1):
    """spectral_ops.py:20-31 (tf.mod is floor-mod; pi and 2 pi are rounded to the tensor dtype)."""
    pi = torch.tensor(math.pi, dtype=phases.dtype)
    two_pi = torch.tensor(math.pi * 2.0, dtype=phases.dtype)
    d = diff(phases, axis=axis)
    mods = torch.remainder(d + pi, two_pi) - pi
    mods = torch.where((mods == -pi) & (d > 0.0), pi.expand_as(mods), mods)
    corrects = mods - d
    cums = torch.cumsum(corrects, dim=axis)
    shape = list(phases.shape)
    shape[axis] = 1
    cums = torch.cat([torch.zeros(shape, dtype=phases.dtype), cums], dim=axis)
    return phases + cums


def instantaneous_frequency(phases, axis=-2):
    """spectral_ops.py:34-42."""
    pi = torch.tensor(math.pi, dtype=phases.dtype)
    unwrapped = unwrap(phases, axis=axis)
    d = diff(unwrapped, axis=axis)
    first = unwrapped.narrow(axis, 0, 1)
    return torch.cat([first, d], dim=axis) / pi


# ----------------------------------------------------------------------------- forward / inverse
def stft(waveforms, frame_length, frame_step):
    """tf.signal.stft with a periodic Hann window, pad_end=False."""
    frames = waveforms.unfold(-1, frame_length, frame_step)
    return torch.fft.rfft(frames * hann_window(frame_length, waveforms.dtype), n=frame_length, dim=-1)


def convert_to_spectrogram(waveforms, waveform_length, sample_rate, spectrogram_shape, overlap,
                           return_intermediates=False):
    """spectral_ops.py:45-94.  waveforms [B, waveform_length] -> (log-mel magnitude, mel IF), each
    [B, time_steps, num_freq_bins]."""
    dt = waveforms.dtype
    time_steps, bins, frame_length, frame_step, num_samples = frame_params(
        waveform_length, spectrogram_shape, overlap)
    x = torch.nn.functional.pad(waveforms, (num_samples - waveform_length, 0))
    s = stft(x, frame_length, frame_step)[..., 1:]
    mag = torch.abs(s)
    # zeros are normalised to +0 so silence has phase 0 whatever the FFT's signed zeros are
    phase = torch.atan2(s.imag + 0.0, s.real + 0.0)
    m = torch.from_numpy(mel_constants(bins, sample_rate)[0]).to(dt)
    mel_mag = mag @ m
    mel_phase = phase @ m
    log_mel = torch.log(mel_mag + 1.0e-6)
    mel_if = instantaneous_frequency(mel_phase, axis=-2)
    log_mel = (log_mel - (-3.76)) / 10.05
    mel_if = (mel_if - 0.0) / 1.0
    if return_intermediates:
        return log_mel, mel_if, dict(mag=mag, phase=phase, mel_mag=mel_mag, mel_phase=mel_phase)
    return log_mel, mel_if


def convert_to_waveform(log_mel, mel_if, waveform_length, sample_rate, spectrogram_shape, overlap):
    """spectral_ops.py:97-149.  (log-mel magnitude, mel IF) [B, T, bins] -> waveforms
    [B, waveform_length]."""
    dt = log_mel.dtype
    time_steps, bins, frame_length, frame_step, num_samples = frame_params(
        waveform_length, spectrogram_shape, overlap)
    log_mel = log_mel * 10.05 + (-3.76)
    mel_if = mel_if * 1.0 + 0.0
    mel_mag = torch.exp(log_mel)
    mel_phase = torch.cumsum(mel_if * torch.tensor(math.pi, dtype=dt), dim=-2)
    p = torch.from_numpy(mel_constants(bins, sample_rate)[1]).to(dt)
    mag = mel_mag @ p
    phase = mel_phase @ p
    s = torch.complex(mag * torch.cos(phase), mag * torch.sin(phase))
    s = torch.nn.functional.pad(s, (1, 0))
    frames = torch.fft.irfft(s, n=frame_length, dim=-1) * inverse_stft_window(frame_length, frame_step, dt)
    # overlap_and_add
    out = torch.zeros(*frames.shape[:-2], num_samples, dtype=dt)
    for t in range(time_steps):
        out[..., t * frame_step:t * frame_step + frame_length] += frames[..., t, :]
    return out[..., num_samples - waveform_length:]
