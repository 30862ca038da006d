"""GPU: networks and the training step on the real kernels against the oracle
(reference networks.py / models.py restatement), fp32, tolerance 1e-3 relative (north_star)."""
import pytest
import torch

from common import FULL, HYPER, SMALL, SPECTRAL, check_substep as _check_substep, rel_err, seeded_inputs
from oracle import models as omodels
from oracle import networks as onet

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(params=["fp32", "tc"])
def conv_mode(request):
    """fp32: every convolution on the exact-fp32 kernels (impl 4) -- the strict parity mode.
    tc: the default, tcgen05 bf16x3 kernels wherever the shape allows (~1e-5 per convolution)."""
    import gansynth_b200.functional as F
    prev = F.K.impl
    F.K.impl = 4 if request.param == "fp32" else 0
    yield request.param
    F.K.impl = prev


def grad_ok(mode, got, want64, want32=None):
    """Direction / size criterion against the fp64 oracle run WITHOUT the product's leaky-relu masks (kept as a coarse sanity bound only; the
    gradient parity claim itself is `check_substep` below): cosine >= 0.9995 / 0.995, relative L2 <= 2e-2 / 1e-1."""
    g, w = got.detach().double().cpu().reshape(-1), want64.detach().double().reshape(-1)
    if float(w.abs().max()) == 0.0:
        return float(g.abs().max()) == 0.0, "zero reference"
    cos = float(torch.dot(g, w) / (g.norm() * w.norm() + 1e-300))
    rel = float((g - w).norm() / w.norm())
    lim = (0.9995, 2e-2) if mode == "fp32" else (0.995, 1e-1)
    return (cos >= lim[0] and rel <= lim[1]), "cos %.6f relL2 %.3e" % (cos, rel)


def check_substep(model, store, ostep, scope, images, labels, latents, dtype, mode, plain=True, retain_graph=False):
    return _check_substep(model, store, ostep, scope, images, labels, latents, dtype, mode, plain, grad_ok, retain_graph)


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
def test_small_forward_parity(cuda_store, conv_mode, level):
    opg, params, ppg = _pair(SMALL, level, cuda_store)
    latents, labels, images = seeded_inputs(4, [16, 16])
    with torch.no_grad():
        got = ppg.generator(latents.cuda(), labels.cuda())
        gf, gl = ppg.discriminator(images.cuda(), labels.cuda())
        want = opg.generator(params, latents, labels)
        wf, wl = opg.discriminator(params, images, labels)
    assert rel_err(got, want) < TOL and rel_err(gf, wf) < TOL and rel_err(gl, wl) < TOL


@pytest.mark.parametrize("level", [0.3, 1.0])
def test_small_step_parity(cuda_store, conv_mode, level):
    """Two full iterations (D update + G update): losses and flat gradients by `check_substep` (masks pinned, 1e-3
    on every element), then the fused TF-Adam update.  With beta1 = 0 the first Adam steps move every element by
    about lr * sign(g), so the updated weights are compared where the gradient is RESOLVED (|g| above 1e-2 of the
    variable's largest): there they must meet 1e-3 outright; elements with a gradient inside the 1e-3 noise floor
    may legitimately step the other way (2 * lr apart).  Each iteration starts from the oracle's weights."""
    import gansynth_b200.models as pmodels
    opg, params, ppg = _pair(SMALL, level, cuda_store)
    latents, labels, images = seeded_inputs(4, [16, 16])
    ostep = omodels.GANSynthStep(opg, {n: p.double() for n, p in params.items()}, HYPER)
    model = pmodels.GANSynth(ppg.generator, ppg.discriminator, None, None, {}, HYPER)
    model._ensure_optimizers(labels.cuda(), latents.cuda())
    lr = HYPER["generator_learning_rate"]
    for it in range(2):
        lat2 = torch.randn(4, 256, generator=torch.Generator().manual_seed(10 + it))
        for scope, z in (("discriminator", latents), ("generator", lat2)):
            loss, got, want = check_substep(model, cuda_store, ostep, scope, images, labels, z, torch.float64, conv_mode,
                                            retain_graph=True)
            before = {n: ostep.params[n].detach().clone() for n in got}
            model._apply(scope, loss)
            if scope == "discriminator":
                ostep.discriminator_update(images.double(), labels.double(), z.double())
            else:
                ostep.generator_update(labels.double(), z.double())
            for n in got:
                ref = ostep.params[n].detach()
                diff = (cuda_store.vars[n].detach().double().cpu() - ref).abs()
                g = want[n].detach().abs()
                resolved = g > 1e-2 * float(g.max())
                bound = TOL * float(ref.abs().max())
                if bool(resolved.any()):
                    assert float(diff[resolved].max()) <= bound, (n, float(diff[resolved].max()))
                assert float(diff.max()) <= bound + 2.5 * lr, (n, float(diff.max()))
                assert float((ref - before[n]).abs().max()) > 0.0 or float(g.max()) == 0.0
            # next sub-step from the oracle's state (the un-resolved elements would otherwise drift apart)
            cuda_store.load({n: ostep.params[n].detach() for n in got})


def test_fake_gradient_penalty_branch(cuda_store, conv_mode):
    """models.py:50-54: the zero-centred penalty on the generator distribution (weight 0.0 on the reference's command
    line, so off the benchmarked path): D sub-step loss and gradients by `check_substep`."""
    import gansynth_b200.models as pmodels
    hp = dict(HYPER, fake_gradient_penalty_weight=2.5)
    opg, params, ppg = _pair(SMALL, 1.0, cuda_store)
    latents, labels, images = seeded_inputs(4, [16, 16])
    ostep = omodels.GANSynthStep(opg, {n: p.double() for n, p in params.items()}, hp)
    model = pmodels.GANSynth(ppg.generator, ppg.discriminator, None, None, {}, hp)
    model._ensure_optimizers(labels.cuda(), latents.cuda())
    check_substep(model, cuda_store, ostep, "discriminator", images, labels, latents, torch.float64, conv_mode)


def test_full_forward_parity(cuda_store, conv_mode):
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


@pytest.mark.parametrize("level", [0.1, 0.3])
def test_full_forward_parity_growing(cuda_store, level):
    """Full-size networks in the growth phase (blend at 32x256 for level 0.1, at 128x1024 for 0.3), batch 4,
    default (tensor-core) mode: the lerp / upscale2d / downscale2d branches of networks.py:109-152, 244-287."""
    opg, params, ppg = _pair(FULL, level, cuda_store)
    latents, labels, images = seeded_inputs(4, [128, 1024])
    with torch.no_grad():
        got = ppg.generator(latents.cuda(), labels.cuda())
        gf, gl = ppg.discriminator(images.cuda(), labels.cuda())
        want = opg.generator(params, latents, labels)
        wf, wl = opg.discriminator(params, images, labels)
    assert got.shape == (4, 2, 128, 1024)
    assert rel_err(got, want) < TOL and rel_err(gf, wf) < TOL and rel_err(gl, wl) < TOL


def test_full_step_gradient_parity(cuda_store, conv_mode):
    """Full-size D and G sub-step gradients (R1 and mode-seeking double backward) at batch 4 against the fp64
    oracle: `check_substep` (masks pinned: 1e-3 on every gradient element; mask flips counted; losses 1e-3)."""
    import gansynth_b200.models as pmodels
    opg, params, ppg = _pair(FULL, 1.0, cuda_store)
    latents, labels, images = seeded_inputs(4, [128, 1024])
    o64 = omodels.GANSynthStep(opg, {n: p.double() for n, p in params.items()}, HYPER)
    model = pmodels.GANSynth(ppg.generator, ppg.discriminator, None, None, {}, HYPER)
    model._ensure_optimizers(labels.cuda(), latents.cuda())
    for scope in ("discriminator", "generator"):
        check_substep(model, cuda_store, o64, scope, images, labels, latents, torch.float64, conv_mode)


def test_full_step_parity_batch8(cuda_store):
    """The metric's own configuration (BASELINE configs[1]: full size, batch 8 -- batch_stddev groups {0,2,4,6} /
    {1,3,5,7}), default tensor-core mode, against the fp32 oracle: losses 1e-3, gradients by `check_substep`."""
    import gansynth_b200.models as pmodels
    opg, params, ppg = _pair(FULL, 1.0, cuda_store)
    latents, labels, images = seeded_inputs(8, [128, 1024])
    o32 = omodels.GANSynthStep(opg, params, HYPER)
    model = pmodels.GANSynth(ppg.generator, ppg.discriminator, None, None, {}, HYPER)
    model._ensure_optimizers(labels.cuda(), latents.cuda())
    for scope in ("discriminator", "generator"):
        check_substep(model, cuda_store, o32, scope, images, labels, latents, torch.float32, "tc")


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
    # evaluate (models.py:196-230) with a callable standing in for the frozen pitch-classifier graph
    from gansynth_b200 import metrics
    proj = torch.randn(2 * 128, 6, generator=torch.Generator().manual_seed(4)).cuda()
    seen = {"real": [], "fake": []}

    def classifier(images):
        feats = images.mean(dim=3).reshape(images.shape[0], -1) @ proj
        seen["real" if images is model.real_images else "fake"].append(feats.cpu().numpy())
        return feats, feats[:, :3]

    it3 = iter(batches[:6])
    model.real_input_fn = lambda: next(it3)
    out = model.evaluate(str(tmp_path), None, classifier, "images:0", ["features:0", "logits:0"])
    assert set(out) == {"frechet_inception_distance"} and np.isfinite(out["frechet_inception_distance"])
    want = metrics.frechet_inception_distance(np.concatenate(seen["real"]), np.concatenate(seen["fake"]))
    assert len(seen["real"]) == 6 and abs(out["frechet_inception_distance"] - want) <= 1e-9 * max(1.0, abs(want))
    with pytest.raises(NotImplementedError):
        model.evaluate(str(tmp_path), None, b"frozen-graph-bytes", "images:0", ["features:0", "logits:0"])
    # TensorFlow-1 Saver files out and back in: variables, global step, Adam slots and step counts
    prefix = model.export_tf_checkpoint(str(tmp_path / "tf" / ("model.ckpt-%d" % int(gs.value))))
    before = {n: v.clone() for n, v in model.store.state().items()}
    slots = {s: (o["m"].clone(), o["v"].clone(), o["t"]) for s, o in model._opt.items()}
    model.store.load({n: torch.zeros_like(v) for n, v in before.items()})
    for o in model._opt.values():
        o["m"].zero_()
        o["v"].zero_()
        o["t"] = 0
    gs.value = 0
    assert model.import_tf_checkpoint(str(tmp_path / "tf")) == prefix
    assert int(gs.value) == 2
    for n, v in model.store.state().items():
        assert torch.equal(v, before[n]), n
    for s_, o in model._opt.items():
        assert torch.equal(o["m"], slots[s_][0]) and torch.equal(o["v"], slots[s_][1]) and o["t"] == slots[s_][2]


def test_cuda_graph_substeps_match_eager(cuda_store):
    """One iteration replayed as CUDA graphs against the same iteration run eagerly FROM THE SAME STATE (weights,
    Adam slots) after two warm-up iterations.  Losses must agree tightly; the flat gradients agree in direction
    and size (the filter-gradient / bias-gradient atomics reorder sums, and TF-Adam with beta1 = 0 turns a sign
    change of a near-zero gradient element into a full-size step, so longer trajectories legitimately diverge)."""
    import gansynth_b200.functional as F
    import gansynth_b200.models as pmodels
    g = torch.Generator().manual_seed(5)
    batches = [(0.5 * torch.randn(4, 512, generator=g),
                torch.nn.functional.one_hot(torch.randint(0, 61, (4,), generator=g), 61).float(),
                torch.randn(4, 256, generator=g), torch.randn(4, 256, generator=g)) for _ in range(4)]
    prev = F.K.impl
    F.K.impl = 4
    try:
        store = cuda_store
        _, params, ppg = _pair(SMALL, 1.0, store)
        model = pmodels.GANSynth(ppg.generator, ppg.discriminator, None, None, {}, HYPER)
        # the spectral kernels are built for 1024 bins: here the "waveforms" are the 2x16x16 images themselves
        model.real_images_from_waveforms = lambda w: w.reshape(4, 2, 16, 16)

        def iteration(batch):
            w, lab, z1, z2 = (t.cuda() for t in batch)
            d = float(model.discriminator_step(w, lab, z1))
            gd = model._opt["discriminator"]["grad"].clone()
            gl = float(model.generator_step(lab, z2))
            gg = model._opt["generator"]["grad"].clone()
            return d, gl, gd, gg

        model.use_cuda_graphs = True
        for b in batches[:2]:
            iteration(b)                     # calls 1 and 2 run eagerly and create every variable / workspace
        snap = dict(vars={n: v.detach().clone() for n, v in store.vars.items()},
                    opt={s: (o["m"].clone(), o["v"].clone(), o["t"]) for s, o in model._opt.items()},
                    step=model.global_step.value)

        def restore():
            with torch.no_grad():
                for n, v in store.vars.items():
                    v.copy_(snap["vars"][n])
            F.K.weight_cache_refresh()       # parameter values changed behind the store's back (VariableStore.load does this)
            for s, o in model._opt.items():
                o["m"].copy_(snap["opt"][s][0]); o["v"].copy_(snap["opt"][s][1]); o["t"] = snap["opt"][s][2]
            model.global_step.value = snap["step"]

        model.use_cuda_graphs = False
        d_e, g_e, gd_e, gg_e = iteration(batches[2])
        restore()
        model.use_cuda_graphs = True
        d_c, g_c, gd_c, gg_c = iteration(batches[2])          # third call of each sub-step: capture + replay
        assert len(model._graphs) == 2 and all("graph" in e for e in model._graphs.values())
        restore()
        d_r, g_r, gd_r, gg_r = iteration(batches[2])          # pure replay
    finally:
        F.K.impl = prev
    for d, gl, gd, gg in ((d_c, g_c, gd_c, gg_c), (d_r, g_r, gd_r, gg_r)):
        assert abs(d - d_e) < 1e-5 * max(1.0, abs(d_e)), (d, d_e)
        assert abs(gl - g_e) < 2e-3 * max(1.0, abs(g_e)), (gl, g_e)
        for got, want in ((gd, gd_e), (gg, gg_e)):
            cos = float(torch.dot(got.double(), want.double()) / (got.double().norm() * want.double().norm()))
            assert cos > 0.999 and abs(float(got.norm() / want.norm()) - 1.0) < 2e-2, cos


def test_generate_batch_graph_replay_matches_eager(cuda_store):
    """Inference replayed as a CUDA graph (third call on) against the eager chain, including after the weights
    changed in place (the weight split is part of the graph).  The generated IMAGES are compared with the eager
    run (1e-4 on the tanh range: the split-K reduction order differs run to run); the waveform is compared with
    the eager inverse transform OF THE SAME IMAGES, because the phase cumsum + sin/cos of the inverse amplify
    1e-7 image differences far beyond any fixed waveform tolerance."""
    import gansynth_b200.models as pmodels
    import gansynth_b200.networks as pnet
    from gansynth_b200 import spectral_ops as sp
    ppg = pnet.PGGAN(growing_level=1.0, **FULL)
    model = pmodels.GANSynth(ppg.generator, ppg.discriminator, None, None, SPECTRAL, HYPER)
    g = torch.Generator().manual_seed(9)
    lab = torch.nn.functional.one_hot(torch.arange(2) % 61, 61).float().cuda()

    def check(z):
        model.use_cuda_graphs = False
        model.generate_batch(lab, z)
        want_img = model.fake_images.clone()
        model.use_cuda_graphs = True
        got = model.generate_batch(lab, z)
        img = model.fake_images.clone()
        assert got.shape == (2, 64000) and bool(torch.isfinite(got).all())
        assert float((img - want_img).abs().max()) <= 1e-4
        wave = sp.convert_to_waveform(img[:, 0].contiguous(), img[:, 1].contiguous(), **SPECTRAL)
        assert float((got - wave).abs().max()) <= 1e-6 * float(wave.abs().max())
        return got, img

    z = torch.randn(2, 256, generator=g).cuda()
    outs = [check(z) for _ in range(4)]
    assert any("graph" in e for k, e in model._graphs.items() if k[0] == "generate")
    assert outs[2][0].data_ptr() != outs[3][0].data_ptr()                 # copies, not the graph's buffer
    with torch.no_grad():
        for n, v in cuda_store.vars.items():
            if n.endswith("dense/weight") and n.startswith("generator/"):
                v.mul_(0.5)
    z2 = torch.randn(2, 256, generator=g).cuda()
    _, img2 = check(z2)                                                   # replay with new inputs and new weights
    assert float((img2 - outs[3][1]).abs().max()) > 1e-3


def test_growth_phase_substeps_replay_as_graphs(cuda_store):
    """Progressive growing (growing_depth <= max_depth): one graph per blend depth, the blend weight in device memory.
    Three consecutive iterations with a different weight each, every one against the same iteration run eagerly from
    the same state; the step crosses no depth boundary (global steps 4..6 of 64: growing_depth 0.52 .. 0.73)."""
    import gansynth_b200.functional as F
    import gansynth_b200.models as pmodels
    g = torch.Generator().manual_seed(6)
    batches = [(0.5 * torch.randn(4, 512, generator=g),
                torch.nn.functional.one_hot(torch.randint(0, 61, (4,), generator=g), 61).float(),
                torch.randn(4, 256, generator=g), torch.randn(4, 256, generator=g)) for _ in range(6)]
    # the device-coefficient blend is the host-coefficient blend, forward and both gradients, bit for bit
    a = torch.randn(3, 5, 7, generator=g).cuda().requires_grad_()
    b = torch.randn(3, 5, 7, generator=g).cuda().requires_grad_()
    coef = torch.tensor([0.3, 0.7], device="cuda")
    y_dev, y_host = F.AxpbyDev.apply(a, b, coef), F.Axpby.apply(a, b, float(coef[0]), float(coef[1]))
    assert torch.equal(y_dev, y_host)
    up = torch.randn(3, 5, 7, generator=g).cuda()
    for yy in (y_dev, y_host):
        ga, gb = torch.autograd.grad(yy, (a, b), up, retain_graph=True)
        assert torch.equal(ga, float(coef[0]) * up) and torch.equal(gb, float(coef[1]) * up)
    prev = F.K.impl
    F.K.impl = 4
    try:
        store = cuda_store
        gs = pmodels.get_or_create_global_step()
        _, params, ppg = _pair(SMALL, gs / 64, store)
        model = pmodels.GANSynth(ppg.generator, ppg.discriminator, None, None, {}, HYPER)
        model.real_images_from_waveforms = lambda w: w.reshape(4, 2, 16, 16)

        def iteration(batch):
            w, lab, z1, z2 = (t.cuda() for t in batch)
            d = float(model.discriminator_step(w, lab, z1))
            gl = float(model.generator_step(lab, z2))
            return d, gl, model._opt["generator"]["grad"].clone()

        def snapshot():
            return dict(vars={n: v.detach().clone() for n, v in store.vars.items()},
                        opt={s: (o["m"].clone(), o["v"].clone(), o["t"]) for s, o in model._opt.items()},
                        step=model.global_step.value)

        def restore(snap):
            with torch.no_grad():
                for n, v in store.vars.items():
                    v.copy_(snap["vars"][n])
            F.K.weight_cache_refresh()       # parameter values changed behind the store's back (VariableStore.load does this)
            for s, o in model._opt.items():
                o["m"].copy_(snap["opt"][s][0]); o["v"].copy_(snap["opt"][s][1]); o["t"] = snap["opt"][s][2]
            model.global_step.value = snap["step"]

        gs.value = 2
        model.use_cuda_graphs = True
        for b in batches[:2]:
            iteration(b)                     # global steps 2 and 3: eager calls 1 and 2 of the depth-1 structure
        assert ppg.structure_key() == ("grow", 1) and int(gs.value) == 4
        results = []
        for b in batches[2:5]:
            snap = snapshot()
            model.use_cuda_graphs = False
            want = iteration(b)
            restore(snap)
            model.use_cuda_graphs = True
            got = iteration(b)               # capture on the first pass through here, pure replays afterwards
            results.append((got, want))
        keys = [k for k in model._graphs if "graph" in model._graphs[k]]
        assert len(keys) == 2 and all(k[1] == (("grow", 1),) for k in keys), keys
        coefs = ppg.lerp_coef.cpu()
        assert abs(float(coefs.sum()) - 1.0) < 1e-6 and 0.0 < float(coefs[0]) < 1.0
    finally:
        F.K.impl = prev
    for (d, gl, gg), (d_e, g_e, gg_e) in results:
        assert abs(d - d_e) < 1e-5 * max(1.0, abs(d_e)), (d, d_e)
        assert abs(gl - g_e) < 2e-3 * max(1.0, abs(g_e)), (gl, g_e)
        cos = float(torch.dot(gg.double(), gg_e.double()) / (gg.double().norm() * gg_e.double().norm()))
        assert cos > 0.999 and abs(float(gg.norm() / gg_e.norm()) - 1.0) < 2e-2, cos
    # the three iterations really used different blend weights
    assert len({round(w[0], 6) for _, w in results}) == 3


# ----------------------------------------------------------------------------- pitch classifier (networks.py:293-413)
@pytest.mark.parametrize("cfg,shape", [
    (dict(conv_param=dict(filters=8, kernel_size=[7, 7], strides=[2, 2]), pool_param=dict(kernel_size=[3, 3], strides=[2, 2]),
          residual_params=[dict(filters=8, strides=[1, 1], blocks=2), dict(filters=16, strides=[2, 2], blocks=2)],
          groups=4, classes=11), (3, 2, 32, 64)),
    # tensor-core sized blocks (64 / 128 / 256 channels) and a 512-channel block on the fp32 kernels
    (dict(conv_param=dict(filters=64, kernel_size=[7, 7], strides=[2, 2]), pool_param=dict(kernel_size=[3, 3], strides=[2, 2]),
          residual_params=[dict(filters=64, strides=[1, 1], blocks=2), dict(filters=128, strides=[2, 2], blocks=1),
                           dict(filters=256, strides=[2, 2], blocks=1), dict(filters=512, strides=[2, 2], blocks=1)],
          groups=32, classes=61), (2, 2, 128, 256)),
])
def test_resnet_classifier_forward_parity(cuda_store, cfg, shape):
    """networks.ResNet on the CUDA kernels (7x7 stem, max pool, group norm + relu, weight-standardised 3x3 / 1x1 stride-2
    convolutions, spatial mean, logits) against the oracle restatement, 1e-3."""
    import gansynth_b200.networks as pnet
    o = onet.ResNet(**cfg)
    params = o.init_variables(seed=5)
    images = torch.randn(*shape, generator=torch.Generator().manual_seed(1))
    net = pnet.ResNet(**cfg)
    net(images.cuda())
    cuda_store.load(params)
    gf, gl = net(images.cuda())
    wf, wl = o({n: p.double() for n, p in params.items()}, images.double())
    assert rel_err(gf, wf) < TOL and rel_err(gl, wl) < TOL


def test_baseline_config1_trains_on_the_cuda_path_with_media_summaries(cuda_store, tmp_path):
    """BASELINE configs[0]: the 2-stage PGGAN (4x4 -> 16x16), batch 4, spectrogram_shape [16, 16] (frame 32, hop 8,
    152-sample clips) -- the spectral front-end runs on the generic kernels -- through GANSynth.train on the GPU with the
    audio / image summaries of models.py:131-161 switched on; losses finite, the event file carries every summary, the
    first D loss equals the oracle's."""
    pytest.importorskip("tensorboard")
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    import gansynth_b200.models as pmodels
    import gansynth_b200.networks as pnet
    from oracle import spectral_ops as osp
    spectral = dict(waveform_length=152, sample_rate=16000, spectrogram_shape=[16, 16], overlap=0.75)
    gs = pmodels.get_or_create_global_step()
    opg, params, ppg = _pair(SMALL, gs / 8, cuda_store)
    g = torch.Generator().manual_seed(0)
    batches = [(0.1 * torch.randn(4, 152, generator=g), torch.nn.functional.one_hot(torch.randint(0, 61, (4,), generator=g), 61).float())
               for _ in range(8)]
    lats = [torch.randn(4, 256, generator=g) for _ in range(8)]
    it, itz = iter(batches), iter(lats)
    model = pmodels.GANSynth(ppg.generator, ppg.discriminator, lambda: next(it), lambda: next(itz), spectral, HYPER)
    model.media_summaries = True
    model.train(str(tmp_path), None, total_steps=3, save_checkpoint_steps=0, save_summary_steps=1, log_tensor_steps=1)
    assert int(gs.value) == 3 and torch.isfinite(model.generator_loss) and torch.isfinite(model.discriminator_loss)
    acc = EventAccumulator(str(tmp_path), size_guidance={"audio": 0, "images": 0, "scalars": 0})
    acc.Reload()
    tags = acc.Tags()
    assert set(tags["scalars"]) >= {"generator_loss", "discriminator_loss"}
    assert sorted(tags["audio"]) == sorted("%s/%d" % (n, i) for n in ("real_waveforms", "fake_waveforms") for i in range(4))
    assert len(tags["images"]) == 16
    assert acc.Audio("fake_waveforms/0")[0].length_frames == 152
    im = acc.Images("real_magnitude_spectrograms/0")[0]
    assert (im.height, im.width) == (16, 16)
    # the very first D loss against the oracle (same weights, first batch, first latents; growing_level 0)
    ostep = omodels.GANSynthStep(onet.PGGAN(growing_level=0.0, **SMALL), params, HYPER)
    real_images = torch.stack(osp.convert_to_spectrogram(batches[0][0], **spectral), dim=1)
    want, _ = ostep.discriminator_update(real_images, batches[0][1], lats[0], apply=False)
    first = [e.value for e in acc.Scalars("discriminator_loss")][0]
    assert abs(first - float(want)) < 1e-3 * max(1.0, abs(float(want))), (first, float(want))


def test_pitch_classifier_training_parity(cuda_store):
    """models.PitchClassifier (models.py:253-410) on the GPU -- generic spectral front-end ([32, 64] spectrogram), ResNet
    forward and backward kernels, weight-decayed Nesterov momentum -- against the oracle restatement: three steps, loss 1e-3,
    every variable 1e-3; then evaluate() returns the accuracy over an input."""
    import gansynth_b200.models as pmodels
    import gansynth_b200.networks as pnet
    from oracle import spectral_ops as osp
    cfg = dict(conv_param=dict(filters=8, kernel_size=[7, 7], strides=[2, 2]), pool_param=dict(kernel_size=[3, 3], strides=[2, 2]),
               residual_params=[dict(filters=8, strides=[1, 1], blocks=2), dict(filters=16, strides=[2, 2], blocks=2)],
               groups=4, classes=11)
    o = onet.ResNet(**cfg)
    params = o.init_variables(seed=5)
    net = pnet.ResNet(**cfg)
    spectral = dict(waveform_length=600, sample_rate=16000, spectrogram_shape=[32, 64], overlap=0.75)
    g = torch.Generator().manual_seed(4)
    batches = [(0.1 * torch.randn(4, 600, generator=g), torch.nn.functional.one_hot(torch.randint(0, 11, (4,), generator=g), 11).float())
               for _ in range(3)]
    hp = dict(weight_decay=1e-3, momentum=0.9, use_nesterov=True,
              learning_rate=lambda step: pmodels.exponential_decay(0.05, step, 2, 0.5))
    it = iter(batches)
    clf = pmodels.PitchClassifier(net, lambda: next(it), spectral, hp)
    ostep = omodels.PitchClassifierStep(o, {n: p.double() for n, p in params.items()}, 1e-3, 0.9, True)
    cuda_images = clf._images                   # the CUDA front-end (generic kernels): used once, for the shapes
    for i, (w, lab) in enumerate(batches):
        images = torch.stack(osp.convert_to_spectrogram(w, **spectral), dim=1)
        if i == 0:
            assert tuple(cuda_images(w.cuda()).shape) == tuple(images.shape)
            clf._ensure_optimizer(cuda_images(w.cuda()))
            cuda_store.load(params)
        # both sides see the ORACLE's images: an instantaneous frequency next to the +-pi branch cut may legitimately come
        # out 2 apart (test_spectral_gpu compares it modulo 2), which a classifier input cannot tolerate
        clf._images = lambda _w, _im=images: _im.cuda()
        lr = hp["learning_rate"](clf.global_step)
        want_total, want_ce, _ = ostep.update(images.double(), lab.double(), lr)
        ce = clf.train_step(w, lab)
        assert abs(float(ce) - want_ce) < TOL * max(1.0, abs(want_ce)), (float(ce), want_ce)
        for n, v in cuda_store.vars.items():
            assert rel_err(v, ostep.params[n]) < TOL, (i, n, rel_err(v, ostep.params[n]))
    it2 = iter(batches)
    clf.input_fn = lambda: next(it2)
    clf._images = lambda wv: torch.stack(osp.convert_to_spectrogram(wv.cpu(), **spectral), dim=1).cuda()
    clf.save_checkpoint("/tmp/gs_clf_ckpt")
    acc = clf.evaluate("/tmp/gs_clf_ckpt")["accuracy"]
    with torch.no_grad():
        want = sum(int((o(ostep.params, torch.stack(osp.convert_to_spectrogram(w, **spectral), dim=1).double())[1].argmax(1)
                        == lab.argmax(1)).sum()) for w, lab in batches) / 12.0
    assert abs(acc - want) <= 1.0 / 12.0 + 1e-9      # (a borderline argmax may differ between the fp32 and the fp64 weights)
