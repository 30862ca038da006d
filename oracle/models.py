"""Oracle restatement of the GANSynth step (reference models.py:22-89), PyTorch CPU autograd.
Test infrastructure only (see oracle/__init__.py).

Loss (models.py:39-65): non-saturating GAN loss on the logit of the true pitch class, R1
zero-centred gradient penalty on reals, gradient-based mode-seeking loss on the generator.
Optimiser: tf.train.AdamOptimizer semantics.  One iteration = D update then G
update, each on a fresh batch (models.py:189-192).
"""
import math

import torch
import torch.nn.functional as F

from . import spectral_ops


def select_logits(logits, labels):
    """models.py:39-40: gather_nd(logits, where(labels)) == logits[b, argmax(labels[b])]."""
    idx = torch.argmax(labels, dim=1)
    return logits.gather(1, idx[:, None])[:, 0]


def real_images_from_waveforms(waveforms, spectral_params):
    """models.py:27-28."""
    mag, inst = spectral_ops.convert_to_spectrogram(waveforms, **spectral_params)
    return torch.stack([mag, inst], dim=1)


def discriminator_loss(pggan, params, real_images, labels, latents, hp):
    """models.py:25-54, 65 for the D update: G forward, D on real and fake, softplus terms, R1."""
    fake_images = pggan.generator(params, latents, labels).detach()
    if hp.get("fake_gradient_penalty_weight"):
        fake_images.requires_grad_(True)
    real_images = real_images.detach().requires_grad_(True)
    _, real_logits = pggan.discriminator(params, real_images, labels)
    _, fake_logits = pggan.discriminator(params, fake_images, labels)
    real_logits = select_logits(real_logits, labels)
    fake_logits = select_logits(fake_logits, labels)
    losses = F.softplus(-real_logits) + F.softplus(fake_logits)
    if hp["real_gradient_penalty_weight"]:
        grads = torch.autograd.grad(real_logits.sum(), real_images, create_graph=True)[0]
        losses = losses + grads.pow(2).sum(dim=(1, 2, 3)) * hp["real_gradient_penalty_weight"]
    if hp.get("fake_gradient_penalty_weight"):
        # models.py:50-54 (off on the reference's command line, gan_synth_main.py:87)
        grads = torch.autograd.grad(fake_logits.sum(), fake_images, create_graph=True)[0]
        losses = losses + grads.pow(2).sum(dim=(1, 2, 3)) * hp["fake_gradient_penalty_weight"]
    return losses.mean()


def generator_loss(pggan, params, labels, latents, hp):
    """models.py:25, 34, 57-64 for the G update."""
    latents = latents.detach().requires_grad_(True)
    fake_images = pggan.generator(params, latents, labels)
    _, fake_logits = pggan.discriminator(params, fake_images, labels)
    fake_logits = select_logits(fake_logits, labels)
    losses = F.softplus(-fake_logits)
    if hp["mode_seeking_loss_weight"]:
        grads = torch.autograd.grad(fake_images.sum(), latents, create_graph=True)[0]
        losses = losses + hp["mode_seeking_loss_weight"] / (grads.pow(2).sum(dim=1) + 1.0e-6)
    return losses.mean()


def network_gradients(loss, params, prefix):
    names = [n for n in params if n.startswith(prefix + "/")]
    grads = torch.autograd.grad(loss, [params[n] for n in names], allow_unused=True)
    return {n: (torch.zeros_like(params[n]) if g is None else g) for n, g in zip(names, grads)}


class TFAdam(object):
    """tf.train.AdamOptimizer: epsilon is added OUTSIDE the bias correction."""

    def __init__(self, names, params, lr, beta1, beta2, epsilon=1.0e-8):
        self.lr, self.b1, self.b2, self.eps, self.t = lr, beta1, beta2, epsilon, 0
        self.m = {n: torch.zeros_like(params[n]) for n in names}
        self.v = {n: torch.zeros_like(params[n]) for n in names}

    def apply(self, params, grads):
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        with torch.no_grad():
            for n, g in grads.items():
                self.m[n] = self.b1 * self.m[n] + (1.0 - self.b1) * g
                self.v[n] = self.b2 * self.v[n] + (1.0 - self.b2) * g * g
                params[n] -= lr_t * self.m[n] / (torch.sqrt(self.v[n]) + self.eps)


class GANSynthStep(object):
    """One D update + one G update on leaf parameter tensors (models.py:67-89, 189-192)."""

    def __init__(self, pggan, params, hp):
        self.pggan, self.hp = pggan, hp
        self.params = {n: p.detach().clone().requires_grad_(True) for n, p in params.items()}
        g_names = [n for n in self.params if n.startswith("generator/")]
        d_names = [n for n in self.params if n.startswith("discriminator/")]
        self.g_opt = TFAdam(g_names, self.params, hp["generator_learning_rate"],
                            hp["generator_beta1"], hp["generator_beta2"])
        self.d_opt = TFAdam(d_names, self.params, hp["discriminator_learning_rate"],
                            hp["discriminator_beta1"], hp["discriminator_beta2"])
        self.global_step = 0

    def discriminator_update(self, real_images, labels, latents, apply=True):
        loss = discriminator_loss(self.pggan, self.params, real_images, labels, latents, self.hp)
        grads = network_gradients(loss, self.params, "discriminator")
        if apply:
            self.d_opt.apply(self.params, grads)
        return loss.detach(), grads

    def generator_update(self, labels, latents, apply=True):
        loss = generator_loss(self.pggan, self.params, labels, latents, self.hp)
        grads = network_gradients(loss, self.params, "generator")
        if apply:
            self.g_opt.apply(self.params, grads)
            self.global_step += 1
        return loss.detach(), grads


class PitchClassifierStep(object):
    """models.py:253-304 restated: softmax cross-entropy (mean over the batch) + weight_decay * sum of tf.nn.l2_loss over the
    variables whose name lacks "normalization", tf.train.MomentumOptimizer with use_nesterov (accum = momentum * accum + g;
    var -= lr * (g + momentum * accum))."""

    def __init__(self, resnet, params, weight_decay, momentum, use_nesterov):
        self.resnet = resnet
        self.params = {n: p.detach().clone().requires_grad_(True) for n, p in params.items()}
        self.accum = {n: torch.zeros_like(p) for n, p in params.items()}
        self.weight_decay, self.momentum, self.use_nesterov = weight_decay, momentum, use_nesterov

    def loss(self, images, labels):
        _, logits = self.resnet(self.params, images)
        ce = -(torch.log_softmax(logits, dim=1) * labels).sum(dim=1).mean()
        l2 = sum((p * p).sum() / 2 for n, p in self.params.items() if "normalization" not in n)
        return ce + self.weight_decay * l2, ce, logits

    def update(self, images, labels, lr):
        total, ce, logits = self.loss(images, labels)
        names = list(self.params)
        grads = torch.autograd.grad(total, [self.params[n] for n in names])
        with torch.no_grad():
            for n, g in zip(names, grads):
                self.accum[n].mul_(self.momentum).add_(g)
                step = g + self.momentum * self.accum[n] if self.use_nesterov else self.accum[n]
                self.params[n].sub_(lr * step)
        return float(total.detach()), float(ce.detach()), logits.detach()
