"""GPU: networks and the training step on the real kernels against the oracle
(reference networks.py / models.py restatement), fp32, tolerance 1e-3 relative (north_star)."""
import pytest
import torch

from common import FULL, HYPER, SMALL, SPECTRAL, grad_close, rel_err, seeded_inputs
from oracle import models as omodels
from oracle import networks as onet

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _pair(cfg, level, store, bias_std=0.1):
    import gansynth_b200.networks as pnet
    opg = onet.PGGAN(growing_level=level, **cfg)
    params = opg.init_variables(seed=3, bias_std=bias_std)
    ppg = pnet.PGGAN(growing_level=level, **cfg)
    ppg._ensure_variables("generator", 256, 61)
    ppg._ensure_variables("discriminator", 0, 61)
    store.load(params)
    return opg, params, ppg


@pytest.mark.parametrize("level", [0.0, 0.1, 0.3, 0.6, 1.0])
def test_small_forward_parity(cuda_store, level):
    opg, params, ppg = _pair(SMALL, level, cuda_store)
    latents, labels, images = seeded_inputs(4, [16, 16])
    with torch.no_grad():
        got = ppg.generator(latents.cuda(), labels.cuda())
        gf, gl = ppg.discriminator(images.cuda(), labels.cuda())
        want = opg.generator(params, latents, labels)
        wf, wl = opg.discriminator(params, images, labels)
    assert rel_err(got, want) < TOL and rel_err(gf, wf) < TOL and rel_err(gl, wl) < TOL


@pytest.mark.parametrize("level", [0.3, 1.0])
def test_small_step_parity(cuda_store, level):
    """Two full iterations (D update + G update): losses, flat gradients and updated weights."""
    import gansynth_b200.models as pmodels
    opg, params, ppg = _pair(SMALL, level, cuda_store)
    latents, labels, images = seeded_inputs(4, [16, 16])
    ostep = omodels.GANSynthStep(opg, params, HYPER)
    model = pmodels.GANSynth(ppg.generator, ppg.discriminator, None, None, {}, HYPER)
    lc, zc, ic = labels.cuda(), latents.cuda(), images.cuda()
    model._ensure_optimizers(lc, zc)
    for it in range(2):
        lat2 = torch.randn(4, 256, generator=torch.Generator().manual_seed(10 + it))
        want_loss, want_grads = ostep.discriminator_update(images, labels, latents)
        model._set_trainable("discriminator")
        loss = model.discriminator_loss_fn(ic, lc, zc)
        assert abs(float(loss.detach()) - float(want_loss)) < TOL * max(1.0, abs(float(want_loss)))
        model._apply("discriminator", loss)
        for n, g in cuda_store.unflatten("discriminator", model._opt["discriminator"]["grad"]).items():
            assert grad_close(g, want_grads[n], TOL), n
        want_loss, want_grads = ostep.generator_update(labels, lat2)
        model._set_trainable("generator")
        loss = model.generator_loss_fn(lc, lat2.cuda())
        assert abs(float(loss.detach()) - float(want_loss)) < TOL * max(1.0, abs(float(want_loss)))
        model._apply("generator", loss)
        for n, g in cuda_store.unflatten("generator", model._opt["generator"]["grad"]).items():
            assert grad_close(g, want_grads[n], TOL), n
        for n, v in cuda_store.vars.items():
            assert rel_err(v, ostep.params[n]) < TOL, n


def test_full_forward_parity(cuda_store):
    """BASELINE config 2 architecture (2x16 -> 128x1024, fully grown), batch 4."""
    opg, params, ppg = _pair(FULL, 1.0, cuda_store)
    latents, labels, images = seeded_inputs(4, [128, 1024])
    with torch.no_grad():
        got = ppg.generator(latents.cuda(), labels.cuda())
        gf, gl = ppg.discriminator(images.cuda(), labels.cuda())
        want = opg.generator(params, latents, labels)
        wf, wl = opg.discriminator(params, images, labels)
    assert got.shape == (4, 2, 128, 1024)
    assert rel_err(got, want) < TOL and rel_err(gf, wf) < TOL and rel_err(gl, wl) < TOL


def test_full_step_gradient_parity(cuda_store):
    """Full-size D and G sub-step gradients (R1 and mode-seeking double backward) at batch 4.

    These second-order gradients are ill-conditioned in fp32: the fp32 ORACLE itself sits 1e-3..8e-3
    (max-norm, per variable) from the fp64 oracle (profiles/grad_diag_r1.txt).  The criterion is therefore
    stated against the fp64 oracle: the CUDA path must be within 1e-3, or within 5x of the error the fp32
    oracle makes on the same variable.  Losses (first-order quantities) must meet 1e-3 outright."""
    import gansynth_b200.models as pmodels
    opg, params, ppg = _pair(FULL, 1.0, cuda_store)
    latents, labels, images = seeded_inputs(4, [128, 1024])
    o32 = omodels.GANSynthStep(opg, params, HYPER)
    o64 = omodels.GANSynthStep(opg, {n: p.double() for n, p in params.items()}, HYPER)
    model = pmodels.GANSynth(ppg.generator, ppg.discriminator, None, None, {}, HYPER)
    lc, zc, ic = labels.cuda(), latents.cuda(), images.cuda()
    model._ensure_optimizers(lc, zc)

    def err(a, ref):
        ref = ref.double()
        return float((a.double().cpu() - ref).abs().max()) / max(float(ref.abs().max()), 1e-30)

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
        assert abs(float(loss.detach()) - float(l64)) < TOL * max(1.0, abs(float(l64)))
        names = list(cuda_store.trainable_variables(scope))
        grads = torch.autograd.grad(loss, [cuda_store.vars[n] for n in names], allow_unused=True)
        for n, g in zip(names, grads):
            if g is None:
                continue
            e_cuda, e_ora = err(g, g64[n]), err(g32[n], g64[n])
            assert e_cuda < max(TOL, 5.0 * e_ora), (n, e_cuda, e_ora)


def test_train_and_generate_entry_points(cuda_store, tmp_path):
    """GANSynth.train / generate with the reference's call signatures on synthetic inputs."""
    import numpy as np
    import gansynth_b200.models as pmodels
    import gansynth_b200.networks as pnet
    gs = pmodels.get_or_create_global_step()
    ppg = pnet.PGGAN(growing_level=gs / 4, **FULL)
    rng = np.random.default_rng(0)
    batches = [(0.1 * rng.standard_normal((4, 64000)).astype(np.float32),
                np.eye(61, dtype=np.float32)[rng.integers(0, 61, 4)]) for _ in range(8)]
    it = iter(batches)
    gen = torch.Generator().manual_seed(1)
    model = pmodels.GANSynth(ppg.generator, ppg.discriminator, lambda: next(it),
                             lambda: torch.randn(4, 256, generator=gen), SPECTRAL, HYPER)
    model.train(str(tmp_path), None, total_steps=2, save_checkpoint_steps=1, save_summary_steps=1, log_tensor_steps=1)
    assert int(gs.value) == 2 and torch.isfinite(model.generator_loss) and torch.isfinite(model.discriminator_loss)
    it2 = iter(batches[:2])
    model.real_input_fn = lambda: next(it2)
    outs = list(model.generate(str(tmp_path)))
    assert len(outs) == 2 and outs[0].shape == (4, 64000) and outs[0].dtype == np.float32
    assert np.isfinite(outs[0]).all()
