"""The product (its real kernels on a GPU, or its host logic over tests/emu_backend.py on CPU) against the vectors the
reference's own code produced (tests/golden/reference_*.npz, see tests/golden/make_reference_vectors.py): fp32 results
against float64 reference values, 1e-3 relative (north_star's tolerance)."""
import importlib.util
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_spec = importlib.util.spec_from_file_location("make_reference_vectors", os.path.join(GOLDEN, "make_reference_vectors.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)

TOL = 1e-3


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def variables(z, prefix):
    return {k[len(prefix):]: torch.from_numpy(z[k]).float() for k in z.files if k.startswith(prefix)}


def rel(got, want):
    got = got.detach().double().cpu().numpy()
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    return float(np.abs(got - want).max() / (np.abs(want).max() + 1e-30))


def product_pggan(store, level, device):
    import gansynth_b200.networks as pnet
    pg = pnet.PGGAN(growing_level=level, **gen.TINY)
    pg._ensure_variables("generator", gen.LATENT, gen.LABELS)
    pg._ensure_variables("discriminator", 0, gen.LABELS)
    return pg


def check_forward(store, device):
    """networks.py generator / discriminator at the five growth levels of the fixture."""
    z = load("reference_pggan")
    dev = lambda k: torch.from_numpy(z[k]).float().to(device)
    for k, level in enumerate(z["levels"]):
        pg = product_pggan(store, float(level), device)
        store.load(variables(z, "var:"))
        with torch.no_grad():
            fake = pg.generator(dev("latents"), dev("labels"))
            features, logits = pg.discriminator(dev("images"), dev("labels"))
        assert rel(fake, z["fake_images_%d" % k]) < TOL, (level, rel(fake, z["fake_images_%d" % k]))
        assert rel(features, z["features_%d" % k]) < TOL and rel(logits, z["logits_%d" % k]) < TOL, level


def check_spectral(device):
    """spectral_ops.py both ways at the fixture size (generic kernels) and the reference's own size (fast kernels)."""
    import gansynth_b200.spectral_ops as psp
    z = load("reference_spectral")
    for tag, params, sub in (("tiny", gen.TINY_SPECTRAL, None), ("full", gen.FULL_SPECTRAL, (8, 16, 32))):
        wave = torch.from_numpy(z[tag + "_wave"]).float().to(device)
        logmel, inst = psp.convert_to_spectrogram(wave, **params)
        want_lm, want_if = (z["tiny_logmel"], z["tiny_inst"]) if sub is None else (z["full_logmel_sub"], z["full_inst_sub"])
        got_lm, got_if = (logmel, inst) if sub is None else (logmel[:, ::sub[0], ::sub[1]], inst[:, ::sub[0], ::sub[1]])
        assert rel(got_lm, want_lm) < TOL, (tag, rel(got_lm, want_lm))
        # instantaneous frequency: a phase difference within rounding of +-pi may unwrap to the other side (2.0 apart,
        # spectral_ops.py:23-25); everywhere else 1e-3 of the +-1 range.  Silent, rounding-dominated bins are the only
        # place that happens: at most 0.5 % of them.
        d = np.abs(got_if.detach().double().cpu().numpy() - want_if)
        wrapped = np.abs(d - 2.0) < 2e-3
        assert float(wrapped.mean()) < 5e-3 and float(d[~wrapped].max()) < 2e-3, (tag, float(wrapped.mean()), float(d[~wrapped].max()))
        # the inverse from the REFERENCE's spectrogram, so both sides start from the same phases
        if sub is None:
            lm64, if64 = torch.from_numpy(z["tiny_logmel"]), torch.from_numpy(z["tiny_inst"])
            back = psp.convert_to_waveform(lm64.float().to(device), if64.float().to(device), **params)
            assert rel(back, z["tiny_back"]) < TOL, (tag, rel(back, z["tiny_back"]))


def check_training_sequence(store, device, fixture="reference_step"):
    """models.py GANSynth: D run, G run, ... under gan_synth_main.py's growth schedule, through the product's public
    sub-step calls.  Losses and the applied gradient of every run against the reference's; after each run the
    reference's variables are loaded, so every run starts from the reference's state (with beta1 = 0 an fp32 gradient
    inside rounding of zero steps the other way, see tests/test_model_gpu.py::test_small_step_parity)."""
    import gansynth_b200.models as pmodels
    z = load(fixture)
    hyper = dict(gen.HYPER, fake_gradient_penalty_weight=float(z["fake_penalty"]))
    pmodels.reset_global_step()
    level = pmodels.get_or_create_global_step() / gen.GROWING_STEPS                       # gan_synth_main.py:48-53
    pg = product_pggan(store, level, device)
    store.load(variables(z, "var0:"))
    model = pmodels.GANSynth(pg.generator, pg.discriminator, None, None, gen.TINY_SPECTRAL, hyper, device=device)
    model.use_cuda_graphs = False
    dev = lambda a: torch.from_numpy(a).float().to(device)
    for run in range(2 * int(z["iterations"])):
        tag = "run%d:" % run
        which = str(z[tag + "which"])
        assert int(model.global_step.value) == int(z[tag + "global_step"])
        waves, labels, latents = dev(z[tag + "waveforms"]), dev(z[tag + "labels"]), dev(z[tag + "latents"])
        if which == "discriminator":
            loss = model.discriminator_step(waves, labels, latents)
        else:
            loss = model.generator_step(labels, latents)
        want = float(z[tag + which + "_loss"])
        assert abs(float(loss.detach()) - want) < TOL * max(1.0, abs(want)), (run, float(loss.detach()), want)
        flat = model._opt[which]["grad"]
        for name, (a, k) in store.offsets[which].items():
            want_g = z[tag + "grad:" + name]
            got_g = flat[a:a + k].detach().double().cpu().numpy().reshape(want_g.shape)
            scale = float(np.abs(want_g).max())
            if scale == 0.0:
                assert float(np.abs(got_g).max()) == 0.0, name
            else:
                assert float(np.abs(got_g - want_g).max()) <= TOL * scale, (run, name, float(np.abs(got_g - want_g).max()) / scale)
        # tf.train.AdamOptimizer: where the gradient is resolved (above 1e-2 of the variable's largest) the updated
        # weights meet the tolerance outright; elsewhere both sides moved by at most ~1.7 lr, possibly opposite ways
        lr = hyper[which + "_learning_rate"]
        for name, ref in variables(z, tag + "var:").items():
            got = store.vars[name].detach().float().cpu()
            g = np.abs(z[tag + "grad:" + name])
            resolved = torch.from_numpy(g > 1e-2 * float(g.max())) if float(g.max()) > 0 else torch.zeros_like(ref, dtype=torch.bool)
            diff = (got - ref).abs()
            bound = TOL * float(ref.abs().max())
            if bool(resolved.any()):
                assert float(diff[resolved].max()) <= bound, (run, name, float(diff[resolved].max()), bound)
            assert float(diff.max()) <= bound + 4.0 * lr, (run, name, float(diff.max()))
        store.load(variables(z, tag + "var:"))
    assert int(model.global_step.value) == int(z["final_global_step"])


def check_classifier(store, device):
    """networks.py ResNet + models.py PitchClassifier: two train-op runs (cross entropy + L2, Nesterov momentum under the
    decaying learning rate of pitch_classifier_main.py:69-74), then the exported features / logits head."""
    import gansynth_b200.models as pmodels
    import gansynth_b200.networks as pnet
    z = load("reference_classifier")
    h = gen.CLASSIFIER_HYPER
    pmodels.reset_global_step()
    hyper = dict(weight_decay=h["weight_decay"], momentum=h["momentum"], use_nesterov=h["use_nesterov"],
                 learning_rate=lambda step: pmodels.exponential_decay(h["base_learning_rate"], step, h["decay_steps"], h["decay_rate"]))
    net = pnet.ResNet(**gen.TINY_RESNET)
    model = pmodels.PitchClassifier(net, None, gen.TINY_SPECTRAL, hyper, device=device)
    dev = lambda a: torch.from_numpy(a).float().to(device)
    shaped = lambda d: {n: v.reshape(tuple(store.vars[n].shape)) for n, v in d.items()}    # gamma / beta: [1, C, 1, 1] there
    for run in range(int(z["iterations"])):
        tag = "run%d:" % run
        waves, labels = dev(z[tag + "waveforms"]), dev(z[tag + "labels"])
        if run == 0:
            model._ensure_optimizer(model._images(waves))
            store.load(shaped(variables(z, "var0:")))
        before = {n: v.detach().clone() for n, v in store.vars.items()}
        decay_loss = model.weight_decay_loss()
        loss = float(model.train_step(waves, labels)) + decay_loss
        want = float(z[tag + "loss"])
        assert abs(loss - want) < TOL * max(1.0, abs(want)), (run, loss, want)
        assert abs(model.accuracy - float(z[tag + "accuracy"])) < 1e-12
        flat = model._opt["grad"]
        for name, (a, k) in store.offsets[model.SCOPE].items():
            want_g = z[tag + "grad:" + name].reshape(-1)
            got_g = flat[a:a + k].detach().double().cpu().numpy()
            if "normalization" not in name:                                              # models.py:267-271
                got_g = got_g + h["weight_decay"] * before[name].detach().double().cpu().numpy().reshape(-1)
            scale = float(np.abs(want_g).max())
            assert float(np.abs(got_g - want_g).max()) <= TOL * scale, (run, name, float(np.abs(got_g - want_g).max()) / scale)
        after = shaped(variables(z, tag + "var:"))
        for name, ref in after.items():
            assert rel(store.vars[name], ref.numpy()) < TOL, (run, name)
        store.load(after)
    with torch.no_grad():
        features, logits = net(dev(z["images"]))
    assert rel(features, z["features"]) < TOL and rel(logits, z["logits"]) < TOL


def check_full_step(store, device, sample_tol, norm_tol=5e-3):
    """The benchmark's own configuration -- fully grown 128x1024 networks, batch 8, gan_synth_main.py's hyper-parameters --
    against ONE session.run of the reference's models.GANSynth (tests/golden/reference_full_step.npz): images, features,
    the selected logits, both losses at 1e-3; of every gradient the L2 norm (`norm_tol`) and 16 evenly spaced elements
    (`sample_tol` of the variable's largest magnitude).  The weights are rebuilt from their names on both sides."""
    import json
    import gansynth_b200.models as pmodels
    import gansynth_b200.networks as pnet
    z = load("reference_full_step")
    with open(os.path.join(GOLDEN, "reference_variables.json")) as f:
        shapes = json.load(f)["gan_synth"]
    pmodels.reset_global_step()
    pg = pnet.PGGAN(growing_level=1.0, **gen.FULL)
    pg._ensure_variables("generator", gen.FULL_LATENT, gen.FULL_LABELS)
    pg._ensure_variables("discriminator", 0, gen.FULL_LABELS)
    assert set(store.vars) == set(shapes)
    store.load({n: gen.named_value(n, s).float() for n, s in shapes.items()})
    model = pmodels.GANSynth(pg.generator, pg.discriminator, None, None, gen.FULL_SPECTRAL, gen.HYPER, device=device)
    model.use_cuda_graphs = False
    waves, labels, latents = (t.float().to(device) for t in gen.full_inputs())
    real = model.real_images_from_waveforms(waves)
    assert rel(real[:, 0, ::4, ::16], z["real_images_sub"][:, 0]) < TOL
    d = np.abs(real[:, 1, ::4, ::16].detach().double().cpu().numpy() - z["real_images_sub"][:, 1])
    wrapped = np.abs(d - 2.0) < 2e-3                         # a phase step within rounding of +-pi unwraps either way
    assert float(wrapped.mean()) < 5e-3 and float(d[~wrapped].max()) < 2e-3, (float(wrapped.mean()), float(d[~wrapped].max()))
    with torch.no_grad():
        fake = pg.generator(latents, labels)
        assert rel(fake[:, :, ::4, ::16], z["fake_images_sub"]) < TOL
        features, logits = pg.discriminator(fake, labels)
        assert rel(features, z["fake_features"]) < TOL
        assert rel(pmodels._select_logits(logits, labels), z["fake_logits"]) < TOL
        features, logits = pg.discriminator(real, labels)
        assert rel(features, z["real_features"]) < TOL
        assert rel(pmodels._select_logits(logits, labels), z["real_logits"]) < TOL
    model._ensure_optimizers(labels, latents)
    for which in ("discriminator", "generator"):
        model._set_trainable(which)
        if which == "discriminator":
            loss = model.discriminator_loss_fn(real, labels, latents)
        else:
            loss = model.generator_loss_fn(labels, latents)
        want = float(z[which + "_loss"])
        assert abs(float(loss.detach()) - want) < TOL * max(1.0, abs(want)), (which, float(loss.detach()), want)
        model._backward(which, loss)
        flat = model._opt[which]["grad"]
        worst, worst_norm, failures = (0.0, ""), (0.0, ""), []
        for name, (a, k) in store.offsets[which].items():
            if "grad:" + name not in z.files:                  # colour blocks of the lower resolutions: not on the grown path
                assert float(flat[a:a + k].abs().max()) == 0.0, name
                continue
            want_s = z["grad:" + name]
            got_s = gen.grad_summary(flat[a:a + k].detach().cpu())
            nerr = abs(got_s[0] - want_s[0]) / want_s[0]
            err = float(np.abs(got_s[2:] - want_s[2:]).max()) / want_s[1]
            worst, worst_norm = max(worst, (err, name)), max(worst_norm, (nerr, name))
            failures += [(name, nerr, err)] if (nerr > norm_tol or err > sample_tol) else []
        print("%s: loss %.6f (reference %.6f); worst gradient norm error %.2e (%s); worst sampled element %.2e of its "
              "variable's maximum (%s)" % (which, float(loss.detach()), want, worst_norm[0], worst_norm[1], worst[0], worst[1]))
        assert not failures, failures[:6]


def fixture_architectures():
    """tests/golden/reference_architectures.npz in the shape of gen.random_architectures()."""
    z = load("reference_architectures")
    for case in z["cases"]:
        tag = "case%d:" % int(case)
        c = [int(v) for v in z[tag + "cfg"]]
        cfg = dict(min_resolution=c[0:2], max_resolution=c[2:4], min_channels=c[4], max_channels=c[5])
        t = lambda k: torch.from_numpy(z[tag + k])
        params = {k[len(tag) + 4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(tag + "var:")}
        yield (cfg, c[6], c[7], float(z[tag + "level"]), t("latents"), t("labels"), t("images"), params, t("fake_images"),
               t("features"), t("logits"))


def check_architectures(device):
    """Architectures off the benchmark's beaten path (non-square 4x5 and 1x1 seeds, channel counts 6 / 12, 2 or 7 classes)
    through the product's networks.py: variable names and shapes, then values at 1e-3."""
    import gansynth_b200.models as pmodels
    import gansynth_b200.networks as pnet
    import gansynth_b200.ops as pops
    for cfg, latent, classes, level, latents, labels, images, params, fake, features, logits in fixture_architectures():
        store = pops.set_default_store(pops.VariableStore(device=device, seed=0))
        pmodels.reset_global_step()
        pg = pnet.PGGAN(growing_level=level, **cfg)
        pg._ensure_variables("generator", latent, classes)
        pg._ensure_variables("discriminator", 0, classes)
        assert {n: tuple(v.shape) for n, v in store.vars.items()} == {n: tuple(v.shape) for n, v in params.items()}, cfg
        store.load({n: v.float() for n, v in params.items()})
        with torch.no_grad():
            got = pg.generator(latents.float().to(device), labels.float().to(device))
            got_features, got_logits = pg.discriminator(images.float().to(device), labels.float().to(device))
        assert rel(got, fake.numpy()) < TOL and rel(got_features, features.numpy()) < TOL and rel(got_logits, logits.numpy()) < TOL, cfg


def fixture_spectral_configs():
    z = load("reference_spectral_configs")
    for case in range(int(z["count"])):
        tag = "case%d:" % case
        p = [int(v) for v in z[tag + "params"]]
        params = dict(waveform_length=p[0], sample_rate=p[1], spectrogram_shape=p[2:4], overlap=float(z[tag + "overlap"]))
        yield (params,) + tuple(torch.from_numpy(z[tag + k]) for k in ("waves", "logmel", "inst", "back"))


def check_spectral_configs(device):
    """The generic spectral kernels over eight configurations of the reference's spectral_ops.py (see the fixture)."""
    import gansynth_b200.spectral_ops as psp
    for params, waves, logmel, inst, back in fixture_spectral_configs():
        got_lm, got_if = psp.convert_to_spectrogram(waves.float().to(device), **params)
        assert rel(got_lm, logmel.numpy()) < TOL, (params, rel(got_lm, logmel.numpy()))
        d = np.abs(got_if.detach().double().cpu().numpy() - inst.numpy())
        wrapped = np.abs(d - 2.0) < 2e-3                     # a phase step within rounding of +-pi unwraps either way
        assert float(wrapped.mean()) < 5e-3 and float(d[~wrapped].max()) < 2e-3, (params, float(wrapped.mean()), float(d[~wrapped].max()))
        got_back = psp.convert_to_waveform(logmel.float().to(device), inst.float().to(device), **params)
        assert rel(got_back, back.numpy()) < TOL, (params, rel(got_back, back.numpy()))


def config1_oracle(z):
    """The oracle stepping through tests/golden/reference_config1.npz: yields (run, which, inputs, oracle step) BEFORE the
    run's update and applies the update afterwards, so callers can compare their own sub-step from the same state."""
    from oracle import models as omodels
    from oracle import networks as onet
    holder = {}
    pg = onet.PGGAN(growing_level=lambda: holder["step"].global_step / gen.CONFIG1_GROWING_STEPS, **gen.CONFIG1)
    names = [str(n) for n in z["variable_names"]]
    g_table, d_table = pg.variable_shapes(gen.FULL_LATENT, gen.FULL_LABELS)
    shapes = {n: s for n, (s, _) in {**g_table, **d_table}.items()}
    assert sorted(names) == sorted(shapes)
    step = holder["step"] = omodels.GANSynthStep(pg, {n: gen.named_value(n, shapes[n]) for n in names}, gen.HYPER)
    for run in range(2 * int(z["iterations"])):
        which = ("discriminator", "generator")[run % 2]
        assert step.global_step == int(z["run%d:global_step" % run])
        waves, labels, latents = gen.config1_inputs(run)
        yield run, which, waves, labels, latents, pg, step
        real = omodels.real_images_from_waveforms(waves, gen.CONFIG1_SPECTRAL)
        if which == "discriminator":
            step.discriminator_update(real, labels, latents)
        else:
            step.generator_update(labels, latents)


def check_config1(store, device, sample_tol=1e-2, norm_tol=5e-3):
    """BASELINE configs[0] (2-stage PGGAN 4x4 -> 16x16, batch 4, the [16, 16] spectrogram) through the product's public
    sub-step calls against the reference's own run of it: loss of every run at 1e-3, norm and samples of every applied
    gradient.  Every run starts from the reference's state, carried by the oracle (which test_reference_pin_cpu.py holds
    to the same file at 1e-10)."""
    import gansynth_b200.models as pmodels
    import gansynth_b200.networks as pnet
    z = load("reference_config1")
    pmodels.reset_global_step()
    level = pmodels.get_or_create_global_step() / gen.CONFIG1_GROWING_STEPS
    pg = pnet.PGGAN(growing_level=level, **gen.CONFIG1)
    pg._ensure_variables("generator", gen.FULL_LATENT, gen.FULL_LABELS)
    pg._ensure_variables("discriminator", 0, gen.FULL_LABELS)
    model = pmodels.GANSynth(pg.generator, pg.discriminator, None, None, gen.CONFIG1_SPECTRAL, gen.HYPER, device=device)
    model.use_cuda_graphs = False
    for run, which, waves, labels, latents, _, ostep in config1_oracle(z):
        store.load({n: p.detach().float() for n, p in ostep.params.items()})
        model.global_step.value = ostep.global_step
        tag = "run%d:" % run
        dev = lambda t: t.float().to(device)
        if which == "discriminator":
            loss = model.discriminator_step(dev(waves), dev(labels), dev(latents))
        else:
            loss = model.generator_step(dev(labels), dev(latents))
        want = float(z[tag + which + "_loss"])
        assert abs(float(loss) - want) < TOL * max(1.0, abs(want)), (run, float(loss), want)
        flat = model._opt[which]["grad"]
        for name, (a, k) in store.offsets[which].items():
            want_s = z[tag + "grad:" + name]
            got_s = gen.grad_summary(flat[a:a + k].detach().cpu())
            if want_s[1] == 0.0:                               # colour blocks the growth level has not reached
                assert got_s[1] == 0.0, (run, name)
                continue
            assert abs(got_s[0] - want_s[0]) <= norm_tol * want_s[0], (run, name, got_s[0], want_s[0])
            assert float(np.abs(got_s[2:] - want_s[2:]).max()) <= sample_tol * want_s[1], (run, name)
