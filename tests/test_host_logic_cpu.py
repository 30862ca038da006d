"""CPU: the product's host logic (autograd composition, wiring, loss, Adam bookkeeping) on the emulated
kernel API must reproduce the oracle (reference networks.py / models.py restatement)."""
import pytest
import torch

from common import HYPER, SMALL, grad_close, rel_err, seeded_inputs
from oracle import models as omodels
from oracle import networks as onet

TOL = 2e-4  # fp32, different summation orders; second-order terms amplify rounding


def _pair(level, emu_store, batch=4, bias_std=0.1):
    import gansynth_b200.networks as pnet
    opg = onet.PGGAN(growing_level=level, **SMALL)
    params = opg.init_variables(seed=3, bias_std=bias_std)
    ppg = pnet.PGGAN(growing_level=level, **SMALL)
    latents, labels, images = seeded_inputs(batch, [16, 16])
    ppg._ensure_variables("generator", 256, 61)
    ppg._ensure_variables("discriminator", 0, 61)
    emu_store.load(params)
    return opg, params, ppg, latents, labels, images


def test_variable_names_and_shapes_match_oracle(emu):
    opg, params, ppg, *_ = _pair(1.0, emu)
    assert list(emu.vars.keys()) == list(params.keys())
    for n, v in emu.vars.items():
        assert tuple(v.shape) == tuple(params[n].shape), n


@pytest.mark.parametrize("level", [0.0, 0.1, 0.3, 0.6, 1.0])
def test_forward_parity(emu, level):
    opg, params, ppg, latents, labels, images = _pair(level, emu)
    want = opg.generator(params, latents, labels)
    got = ppg.generator(latents, labels)
    assert got.shape == want.shape
    assert rel_err(got, want) < TOL
    wf, wl = opg.discriminator(params, images, labels)
    gf, gl = ppg.discriminator(images, labels)
    assert rel_err(gf, wf) < TOL and rel_err(gl, wl) < TOL


@pytest.mark.parametrize("level", [0.3, 1.0])
def test_step_gradients_and_adam_parity(emu, level):
    import gansynth_b200.models as pmodels
    opg, params, ppg, latents, labels, images = _pair(level, emu)
    ostep = omodels.GANSynthStep(opg, params, HYPER)
    model = pmodels.GANSynth(ppg.generator, ppg.discriminator, None, None, {}, HYPER, device="cpu")
    model._ensure_optimizers(labels, latents)
    for it in range(2):
        lat2 = torch.randn(4, 256, generator=torch.Generator().manual_seed(10 + it))
        # D sub-step
        want_loss, want_grads = ostep.discriminator_update(images, labels, latents)
        model._set_trainable("discriminator")
        loss = model.discriminator_loss_fn(images, labels, latents)
        assert abs(float(loss) - float(want_loss)) < 1e-4 * max(1.0, abs(float(want_loss)))
        model._apply("discriminator", loss)
        for n, g in emu.unflatten("discriminator", model._opt["discriminator"]["grad"]).items():
            assert grad_close(g, want_grads[n], 5e-4), n
        # G sub-step
        want_loss, want_grads = ostep.generator_update(labels, lat2)
        model._set_trainable("generator")
        loss = model.generator_loss_fn(labels, lat2)
        assert abs(float(loss) - float(want_loss)) < 1e-4 * max(1.0, abs(float(want_loss)))
        model._apply("generator", loss)
        for n, g in emu.unflatten("generator", model._opt["generator"]["grad"]).items():
            assert grad_close(g, want_grads[n], 5e-4), n
        # weights after the TF-Adam update
        for n, v in emu.vars.items():
            assert rel_err(v, ostep.params[n]) < 5e-4, n
