"""Shared helpers for the parity tests: seeded inputs, hyper-parameters, oracle <-> product glue."""
import torch

HYPER = dict(generator_learning_rate=8e-4, generator_beta1=0.0, generator_beta2=0.99,
             discriminator_learning_rate=8e-4, discriminator_beta1=0.0, discriminator_beta2=0.99,
             mode_seeking_loss_weight=0.1, real_gradient_penalty_weight=5.0, fake_gradient_penalty_weight=0.0)

SPECTRAL = dict(waveform_length=64000, sample_rate=16000, spectrogram_shape=[128, 1024], overlap=0.75)

SMALL = dict(min_resolution=[4, 4], max_resolution=[16, 16], min_channels=32, max_channels=256)   # BASELINE config 1
FULL = dict(min_resolution=[2, 16], max_resolution=[128, 1024], min_channels=32, max_channels=256)


def seeded_inputs(batch, res, seed=0, latent_dim=256, num_labels=61):
    g = torch.Generator().manual_seed(seed)
    latents = torch.randn(batch, latent_dim, generator=g)
    labels = torch.nn.functional.one_hot(torch.randint(0, num_labels, (batch,), generator=g), num_labels).float()
    images = torch.randn(batch, 2, *res, generator=g) * 0.5
    return latents, labels, images


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def grad_close(got, want, tol):
    """max |got - want| <= tol * max |want| per variable; an all-zero reference (inactive colour block)
    requires an all-zero result."""
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    scale = float(want.abs().max())
    if scale == 0.0:
        return float(got.abs().max()) == 0.0
    return float((got - want).abs().max()) <= tol * scale
