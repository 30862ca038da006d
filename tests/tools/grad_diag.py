"""Per-variable error of the full-size D / G sub-step gradients: CUDA path and fp32 oracle, both against
the fp64 oracle.  Run on the GPU box: python tests/tools/grad_diag.py [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import FULL, HYPER, seeded_inputs  # noqa: E402
import gansynth_b200.models as pmodels  # noqa: E402
import gansynth_b200.networks as pnet  # noqa: E402
import gansynth_b200.ops as ops  # noqa: E402
from oracle import models as omodels  # noqa: E402
from oracle import networks as onet  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
store = ops.set_default_store(ops.VariableStore(device="cuda", seed=0))
opg = onet.PGGAN(growing_level=1.0, **FULL)
params = opg.init_variables(seed=3, bias_std=0.1)
ppg = pnet.PGGAN(growing_level=1.0, **FULL)
ppg._ensure_variables("generator", 256, 61)
ppg._ensure_variables("discriminator", 0, 61)
store.load(params)
latents, labels, images = seeded_inputs(batch, [128, 1024])
model = pmodels.GANSynth(ppg.generator, ppg.discriminator, None, None, {}, HYPER)
lc, zc, ic = labels.cuda(), latents.cuda(), images.cuda()
model._ensure_optimizers(lc, zc)
o32 = omodels.GANSynthStep(opg, params, HYPER)
o64 = omodels.GANSynthStep(opg, {n: p.double() for n, p in params.items()}, HYPER)


def err(a, ref):
    ref = ref.double()
    s = float(ref.abs().max())
    return float((a.double().cpu() - ref).abs().max()) / s if s else float(a.abs().max())


for scope in ("discriminator", "generator"):
    if scope == "discriminator":
        l32, g32 = o32.discriminator_update(images, labels, latents, apply=False)
        l64, g64 = o64.discriminator_update(images.double(), labels.double(), latents.double(), apply=False)
        model._set_trainable(scope)
        loss = model.discriminator_loss_fn(ic, lc, zc)
    else:
        l32, g32 = o32.generator_update(labels, latents, apply=False)
        l64, g64 = o64.generator_update(labels.double(), latents.double(), apply=False)
        model._set_trainable(scope)
        loss = model.generator_loss_fn(lc, zc)
    names = list(store.trainable_variables(scope))
    grads = torch.autograd.grad(loss, [store.vars[n] for n in names], allow_unused=True)
    print("%s loss: cuda %.8f  oracle32 %.8f  oracle64 %.8f" % (scope, float(loss.detach()), float(l32), float(l64)))
    print("%-60s %10s %10s %10s %10s %10s" % ("variable", "cuda/64", "ora32/64", "|g|max", "relL2", "1-cos"))
    for n, g in zip(names, grads):
        if g is None:
            continue
        gg, ww = g.double().cpu().reshape(-1), g64[n].reshape(-1)
        rel = float((gg - ww).norm() / ww.norm())
        cos = float(torch.dot(gg, ww) / (gg.norm() * ww.norm()))
        print("%-60s %10.2e %10.2e %10.2e %10.2e %10.2e" % (n, err(g, g64[n]), err(g32[n], g64[n]), float(g64[n].abs().max()), rel, 1 - cos))
