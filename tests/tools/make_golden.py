"""Writes tests/golden/*.npz from the oracle (the reference itself cannot run here: no TensorFlow).
Run from the repo root:  python tests/tools/make_golden.py"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import HYPER, SMALL, SPECTRAL, seeded_inputs  # noqa: E402
from oracle import models as omodels  # noqa: E402
from oracle import networks as onet  # noqa: E402
from oracle import spectral_ops as osp  # noqa: E402

out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)
level, seed = 0.3, 3
pg = onet.PGGAN(growing_level=level, **SMALL)
params = pg.init_variables(seed=seed, bias_std=0.1)
latents, labels, images = seeded_inputs(4, [16, 16])
fake = pg.generator(params, latents, labels)
_, logits = pg.discriminator(params, images, labels)
step = omodels.GANSynthStep(pg, params, HYPER)
d, _ = step.discriminator_update(images, labels, latents)
g, _ = step.generator_update(labels, latents)
np.savez_compressed(os.path.join(out, "small_step.npz"), level=level, seed=seed, latents=latents.numpy(),
                    labels=labels.numpy(), images=images.numpy(), fake_images=fake.detach().numpy(),
                    logits=logits.detach().numpy(), d_loss=float(d), g_loss=float(g))
t = torch.arange(64000) / 16000.0
gen = torch.Generator().manual_seed(0)
wave = torch.stack([0.3 * torch.sin(2 * math.pi * 440.0 * t) * torch.exp(-3 * t) + 0.01 * torch.randn(64000, generator=gen),
                    0.1 * torch.randn(64000, generator=gen)])
lm, inst = osp.convert_to_spectrogram(wave, **SPECTRAL)
back = osp.convert_to_waveform(lm, inst, **SPECTRAL)
np.savez_compressed(os.path.join(out, "spectral.npz"), wave=wave.numpy(), logmel_sub=lm.numpy()[:, ::16, ::16],
                    inst_sub=inst.numpy()[:, ::16, ::16], back_sub=back.numpy()[:, ::64])
print("wrote", os.listdir(out))
