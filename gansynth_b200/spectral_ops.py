"""Spectral front-end on the CUDA kernels (host mirror of reference spectral_ops.py:45-149: same
function names and arguments).

    convert_to_spectrogram(waveforms [B, L], waveform_length, sample_rate, spectrogram_shape, overlap)
        -> (log-mel magnitude, mel instantaneous frequency), each [B, T, 1024]
    convert_to_waveform(log_mel, mel_if, waveform_length, sample_rate, spectrogram_shape, overlap)
        -> waveforms [B, L]

The fast kernels are specialised to the reference's production configuration (gan_synth_main.py:70-75):
1024 frequency bins (frame 2048) and 75 % overlap (hop 512); time_steps and waveform_length are free.  Any other
(bins, overlap) -- e.g. BASELINE config 1's [16, 16] spectrogram -- runs on the generic kernels (direct DFT, dense mel
products; csrc/spectral_generic.cu).
The constant matrices the reference lets TF constant-fold (tf.signal.linear_to_mel_weight_matrix,
tfp.math.pinv; spectral_ops.py:76-82,115-122) are built once on the host here and handed to the
kernels in sparse form: the mel matrix has <= 6 non-zeros per column, its pseudo-inverse decays
exponentially away from a band and is truncated at 1e-8 of its largest entry (<= 48 rows per column).
"""
import math

import numpy as np
import torch

from . import functional as F

NUM_BINS = 1024
OVERLAP = 0.75
MEL_TAPS = 6
PINV_REL_THRESHOLD = 1.0e-8


def _hertz_to_mel(f, dtype):
    return (dtype(1127.0) * np.log(dtype(1.0) + f / dtype(700.0))).astype(dtype)


def linear_to_mel_weight_matrix(num_mel_bins, num_spectrogram_bins, sample_rate, lower_edge_hertz,
                                upper_edge_hertz, dtype=np.float32):
    """HTK mel filterbank with TF-1.13 float32 evaluation (tf.signal.linear_to_mel_weight_matrix):
    triangles on the mel axis, DC row zero.  Returns [num_spectrogram_bins, num_mel_bins]."""
    t = dtype
    lin = np.linspace(t(0.0), t(sample_rate / 2.0), num_spectrogram_bins, dtype=dtype)[1:]
    spec_mel = _hertz_to_mel(lin, dtype)[:, None]
    edges = np.linspace(_hertz_to_mel(np.asarray(lower_edge_hertz, dtype), dtype),
                        _hertz_to_mel(np.asarray(upper_edge_hertz, dtype), dtype), num_mel_bins + 2, dtype=dtype)
    lo, ce, up = edges[None, :-2], edges[None, 1:-1], edges[None, 2:]
    w = np.maximum(t(0.0), np.minimum((spec_mel - lo) / (ce - lo), (up - spec_mel) / (up - ce)))
    return np.pad(w, [[1, 0], [0, 0]]).astype(dtype)


def pseudo_inverse(a):
    """tfp.math.pinv: SVD with singular values below 10 * max(shape) * eps * s_max dropped."""
    rcond = 10.0 * max(a.shape) * np.finfo(a.dtype).eps
    u, s, vt = np.linalg.svd(a, full_matrices=False)
    keep = s > rcond * s.max()
    s_inv = np.where(keep, 1.0 / np.where(keep, s, 1.0), 0.0).astype(a.dtype)
    return ((vt.T * s_inv) @ u.T).astype(a.dtype)


def _column_sparse_mel(m):
    """[bins, mel] -> first non-zero row per mel column and MEL_TAPS zero-padded weights."""
    k0 = np.zeros(m.shape[1], np.int32)
    w = np.zeros((MEL_TAPS, m.shape[1]), np.float32)
    for j in range(m.shape[1]):
        nz = np.nonzero(m[:, j])[0]
        if nz.size == 0:
            continue
        lo, hi = int(nz.min()), int(nz.max())
        if hi - lo + 1 > MEL_TAPS:
            raise ValueError("mel column %d spans %d rows (> %d)" % (j, hi - lo + 1, MEL_TAPS))
        k0[j] = lo
        w[:hi - lo + 1, j] = m[lo:hi + 1, j]
    return k0, w


def _banded_pinv(p):
    """[mel, bins] -> per linear bin: first kept mel row, kept row count, zero-padded weights."""
    thr = PINV_REL_THRESHOLD * np.abs(p).max()
    keep = np.abs(p) > thr
    j0 = np.zeros(p.shape[1], np.int32)
    cnt = np.zeros(p.shape[1], np.int32)
    for k in range(p.shape[1]):
        nz = np.nonzero(keep[:, k])[0]
        if nz.size:
            j0[k], cnt[k] = int(nz.min()), int(nz.max() - nz.min() + 1)
    band = int(max(1, cnt.max()))
    w = np.zeros((band, p.shape[1]), np.float32)
    for k in range(p.shape[1]):
        w[:cnt[k], k] = p[j0[k]:j0[k] + cnt[k], k]
    return j0, cnt, w, band


_HOST_CONSTS = {}
_DEVICE_CONSTS = {}


def host_constants(sample_rate):
    """Dense fp32 constants (mel matrix, its pseudo-inverse, analysis / synthesis windows) and their
    sparse forms.  Cached per sample rate."""
    if sample_rate not in _HOST_CONSTS:
        m = linear_to_mel_weight_matrix(NUM_BINS, NUM_BINS, sample_rate, 0.0, sample_rate / 2.0)
        p = pseudo_inverse(m)
        n = torch.arange(2 * NUM_BINS, dtype=torch.float64)
        hann = (0.5 - 0.5 * torch.cos(2.0 * math.pi * n / (2 * NUM_BINS))).to(torch.float32)
        hop = int((1.0 - OVERLAP) * 2 * NUM_BINS)
        denom = (hann * hann).reshape(-1, hop).sum(0).repeat(2 * NUM_BINS // hop)
        mel_k0, mel_w = _column_sparse_mel(m)
        pb_j0, pb_cnt, pb_w, band = _banded_pinv(p)
        _HOST_CONSTS[sample_rate] = dict(mel=m, pinv=p, hann=hann.numpy(), synth_window=(hann / denom).numpy(),
                                         mel_k0=mel_k0, mel_w=mel_w, pb_j0=pb_j0, pb_cnt=pb_cnt, pb_w=pb_w, band=band)
    return _HOST_CONSTS[sample_rate]


def device_constants(sample_rate, device):
    key = (sample_rate, str(device))
    if key not in _DEVICE_CONSTS:
        h = host_constants(sample_rate)
        d = {k: torch.from_numpy(np.ascontiguousarray(h[k])).to(device)
             for k in ("hann", "synth_window", "mel_k0", "mel_w", "pb_j0", "pb_cnt", "pb_w")}
        d["band"] = h["band"]
        _DEVICE_CONSTS[key] = d
    return _DEVICE_CONSTS[key]


def _check_config(spectrogram_shape, overlap):
    """-> (time_steps, bins, fast): `fast` = the reference's production configuration (1024 bins, 75 % overlap), which
    the specialised kernels of csrc/spectral.cu serve; every other configuration (spectral_ops.py:50-53 is generic;
    BASELINE config 1 uses [16, 16]) runs on the generic kernels of csrc/spectral_generic.cu."""
    time_steps, num_freq_bins = (int(s) for s in spectrogram_shape)
    fast = num_freq_bins == NUM_BINS and abs(float(overlap) - OVERLAP) <= 1e-12
    return time_steps, num_freq_bins, fast


_GENERIC_CONSTS = {}


def generic_constants(num_freq_bins, sample_rate, overlap, device):
    """Dense constants of the generic path: analysis / synthesis windows, linear->mel matrix and its pseudo-inverse
    (spectral_ops.py:50-53, 58-62, 76-82, 115-122, 133-139) -> (dict of device tensors, frame_step)."""
    key = (int(num_freq_bins), sample_rate, float(overlap), str(device))
    if key not in _GENERIC_CONSTS:
        bins = int(num_freq_bins)
        frame_length = 2 * bins
        frame_step = int((1.0 - overlap) * frame_length)
        if frame_step <= 0 or frame_length % frame_step:
            raise NotImplementedError("overlap %r: the frame step %d must divide the frame length %d" % (overlap, frame_step, frame_length))
        m = linear_to_mel_weight_matrix(bins, bins, sample_rate, 0.0, sample_rate / 2.0)
        p = pseudo_inverse(m)
        n = torch.arange(frame_length, dtype=torch.float64)
        hann = (0.5 - 0.5 * torch.cos(2.0 * math.pi * n / frame_length)).to(torch.float32)
        denom = (hann * hann).reshape(-1, frame_step).sum(0).repeat(frame_length // frame_step)
        host = dict(hann=hann, synth_window=hann / denom, mel=torch.from_numpy(np.ascontiguousarray(m)),
                    pinv=torch.from_numpy(np.ascontiguousarray(p)))
        _GENERIC_CONSTS[key] = ({k: v.contiguous().to(device) for k, v in host.items()}, frame_step)
    return _GENERIC_CONSTS[key]


_SM_COUNT = None


def sm_count():
    """Streaming multiprocessors of the current device (148 on a B200); the launch policies below size their grids from it."""
    global _SM_COUNT
    if _SM_COUNT is None:
        _SM_COUNT = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count \
            if torch.cuda.is_available() else 148
    return _SM_COUNT


def frames_per_run(batch, time_steps):
    """Consecutive frames one CTA (16 warps, one frame per warp per round) walks.  Longer runs mean fewer
    run-boundary rows for the fix-up pass; shorter runs give the SMs (148 on a B200) more CTAs to balance: 32 when that
    still leaves >= 4 CTAs per SM (batch 256: 1024 CTAs = 6.9 waves), else one round of 16."""
    if batch * -(-time_steps // 32) >= 4 * sm_count():
        return 32
    return 16


def frames_per_segment(batch, time_steps):
    """Inverse kernel: one CTA per clip when the batch fills the SMs; smaller batches cut every clip into
    segments of a multiple of 8 frames so that about one CTA per SM exists (batch 8: 16 segments of 8 frames)."""
    segs = min(max(1, sm_count() // max(1, batch)), max(1, time_steps // 8))
    if segs <= 1:
        return time_steps
    return 8 * -(-time_steps // (8 * segs))


def convert_to_spectrogram(waveforms, waveform_length, sample_rate, spectrogram_shape, overlap):
    """spectral_ops.py:45-94."""
    time_steps, bins, fast = _check_config(spectrogram_shape, overlap)
    if waveforms.shape[1] != waveform_length:
        raise ValueError("waveforms have %d samples, waveform_length is %d" % (waveforms.shape[1], waveform_length))
    if not fast:
        consts, frame_step = generic_constants(bins, sample_rate, overlap, waveforms.device)
        return F.K.spectrogram_generic(waveforms.detach(), consts, time_steps, bins, frame_step)
    consts = device_constants(sample_rate, waveforms.device)
    return F.K.spectrogram_fwd(waveforms.detach(), consts, time_steps,
                               frames_per_run(waveforms.shape[0], time_steps))


def convert_to_waveform(log_mel_magnitude_spectrograms, mel_instantaneous_frequencies, waveform_length, sample_rate,
                        spectrogram_shape, overlap):
    """spectral_ops.py:97-149."""
    time_steps, bins, fast = _check_config(spectrogram_shape, overlap)
    if tuple(log_mel_magnitude_spectrograms.shape[1:]) != (time_steps, bins):
        raise ValueError("spectrograms must be [B, %d, %d]" % (time_steps, bins))
    if not fast:
        consts, frame_step = generic_constants(bins, sample_rate, overlap, log_mel_magnitude_spectrograms.device)
        return F.K.waveform_generic(log_mel_magnitude_spectrograms.detach(), mel_instantaneous_frequencies.detach(), consts,
                                    waveform_length, bins, frame_step)
    consts = device_constants(sample_rate, log_mel_magnitude_spectrograms.device)
    return F.K.waveform_fwd(log_mel_magnitude_spectrograms.detach(), mel_instantaneous_frequencies.detach(), consts,
                            waveform_length, frames_per_segment(log_mel_magnitude_spectrograms.shape[0], time_steps))
