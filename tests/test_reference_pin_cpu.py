"""CPU: the oracle against vectors the reference's OWN code produced.

`tests/golden/reference_*.npz` were written by `tests/golden/make_reference_vectors.py`: the unmodified networks.py /
spectral_ops.py / models.py of /root/reference executed op by op over `oracle/tf1_eager` (an eager restatement of the
TensorFlow-1.13 primitives they call).  Everything above the primitives -- scopes and variable shapes, He constants, block
order, the tf.cond growth arms and lerp weights, label conditioning, the three loss terms, both tf.gradients penalties,
Adam / Nesterov-momentum updates, the D-run / G-run sequence with the growth level following global_step -- is therefore
the reference's, and the oracle has to reproduce it to float64 round-off.  The last test re-runs the generator where
/root/reference exists, so the committed files cannot drift away from what the reference says.
"""
import importlib.util
import json
import os

import numpy as np
import pytest
import torch

from oracle import models as omodels
from oracle import networks as onet
from oracle import spectral_ops as osp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_spec = importlib.util.spec_from_file_location("make_reference_vectors", os.path.join(GOLDEN, "make_reference_vectors.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)

TIGHT = 1e-10      # float64 on both sides; the two differ only in the order of a few additions


def _load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _close(got, want, tol=TIGHT):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    scale = max(1.0, float(np.abs(want).max()))
    err = float(np.abs(got - want).max())
    assert err <= tol * scale, (err, scale)


def _grad_close(got, want, tol=1e-8):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64).reshape(np.shape(got))
    scale = float(np.abs(want).max())
    if scale == 0.0:
        assert float(np.abs(got).max()) == 0.0
    else:
        assert float(np.abs(got - want).max()) <= tol * scale, (float(np.abs(got - want).max()), scale)


def _variables(z, prefix):
    return {k[len(prefix):]: _t(z[k]) for k in z.files if k.startswith(prefix)}


# ------------------------------------------------------------------------------------------------ networks.py
def test_generator_and_discriminator_at_every_growth_regime():
    z = _load("reference_pggan")
    params = _variables(z, "var:")
    g_table, d_table = onet.PGGAN(growing_level=0.0, **gen.TINY).variable_shapes(gen.LATENT, gen.LABELS)
    assert set(params) == set(g_table) | set(d_table)            # the reference's scopes give exactly the oracle's names
    for name, (shape, _) in {**g_table, **d_table}.items():
        assert tuple(params[name].shape) == tuple(shape), name
    depths = []
    for k, level in enumerate(z["levels"]):
        pg = onet.PGGAN(growing_level=float(level), **gen.TINY)
        _close(pg.growing_depth, z["growing_depth_%d" % k])
        depths.append(pg.growing_depth)
        _close(pg.generator(params, _t(z["latents"]), _t(z["labels"])), z["fake_images_%d" % k])
        features, logits = pg.discriminator(params, _t(z["images"]), _t(z["labels"]))
        _close(features, z["features_%d" % k])
        _close(logits, z["logits_%d" % k])
    # the five levels walk through every arm: nothing grown, first blend, middle blends, fully grown
    assert depths[0] == 0.0 and 0.0 < depths[1] < 1.0 and 1.0 < depths[2] < 3.0 and depths[-1] > 3.0


def test_variable_tables_of_both_command_lines():
    with open(os.path.join(GOLDEN, "reference_variables.json")) as f:
        tables = json.load(f)
    g_table, d_table = onet.PGGAN(growing_level=0.0, min_resolution=[2, 16], max_resolution=[128, 1024], min_channels=32,
                                  max_channels=256).variable_shapes(256, 61)
    want = {n: list(s) for n, (s, _) in {**g_table, **d_table}.items()}
    assert tables["gan_synth"] == want
    count = lambda prefix: sum(int(np.prod(shape)) for name, shape in tables["gan_synth"].items() if name.startswith(prefix))
    # the sizes of the two flat gradient buffers the data-parallel exchange moves: 35.7 MB and 27.3 MB
    assert count("generator/") == 8932238 and count("discriminator/") == 6830973
    resnet = onet.ResNet(conv_param=dict(filters=64, kernel_size=[7, 7], strides=[2, 2]), pool_param=dict(kernel_size=[3, 3], strides=[2, 2]),
                         residual_params=[dict(filters=64, strides=[1, 1], blocks=3), dict(filters=128, strides=[2, 2], blocks=4),
                                          dict(filters=256, strides=[2, 2], blocks=6), dict(filters=512, strides=[2, 2], blocks=3)],
                         groups=32, classes=61)
    ours = resnet.variable_shapes()
    assert set(tables["pitch_classifier"]) == set(ours)
    for name, shape in tables["pitch_classifier"].items():
        if name.endswith(("/beta", "/gamma")):                   # the reference keeps these as [1, C, 1, 1] (ops.py:131-140)
            assert shape == [1, ours[name][0], 1, 1], name
        else:
            assert shape == list(ours[name]), name


def test_odd_architectures_from_fixture():
    """Seed resolutions 4x5 and 1x1, channel counts that are not multiples of four, 2 / 7 classes, 3 / 8 latents
    (tests/golden/reference_architectures.npz): the oracle on any box."""
    import reference_vectors as rv
    for cfg, latent, classes, level, latents, labels, images, params, fake, features, logits in rv.fixture_architectures():
        ours = onet.PGGAN(growing_level=level, **cfg)
        _close(ours.generator(params, latents, labels), fake)
        got_features, got_logits = ours.discriminator(params, images, labels)
        _close(got_features, features)
        _close(got_logits, logits)


def test_product_odd_architectures_from_fixture(emu):
    import reference_vectors as rv
    rv.check_architectures("cpu")


# ------------------------------------------------------------------------------------------------ spectral_ops.py
def test_spectral_round_trip():
    z = _load("reference_spectral")
    wave = _t(z["full_wave"]).double()
    logmel, inst = osp.convert_to_spectrogram(wave, **gen.FULL_SPECTRAL)
    _close(logmel[:, ::8, ::16], z["full_logmel_sub"])
    _close(inst[:, ::8, ::16], z["full_inst_sub"])
    _close(osp.convert_to_waveform(logmel, inst, **gen.FULL_SPECTRAL)[:, ::32], z["full_back_sub"])
    logmel, inst = osp.convert_to_spectrogram(_t(z["tiny_wave"]), **gen.TINY_SPECTRAL)
    _close(logmel, z["tiny_logmel"])
    _close(inst, z["tiny_inst"])
    _close(osp.convert_to_waveform(logmel, inst, **gen.TINY_SPECTRAL), z["tiny_back"])


# ------------------------------------------------------------------------------------------------ models.py GANSynth
@pytest.mark.parametrize("fixture", ["reference_step", "reference_step_fake_penalty"])
def test_training_sequence(fixture):
    z = _load(fixture)
    hyper = dict(gen.HYPER, fake_gradient_penalty_weight=float(z["fake_penalty"]))
    holder = {}
    pg = onet.PGGAN(growing_level=lambda: holder["step"].global_step / gen.GROWING_STEPS, **gen.TINY)   # gan_synth_main.py:48-53
    step = holder["step"] = omodels.GANSynthStep(pg, _variables(z, "var0:"), hyper)
    for run in range(2 * int(z["iterations"])):
        tag = "run%d:" % run
        which = str(z[tag + "which"])
        assert step.global_step == int(z[tag + "global_step"])
        waves, labels, latents = _t(z[tag + "waveforms"]), _t(z[tag + "labels"]), _t(z[tag + "latents"])
        real = omodels.real_images_from_waveforms(waves, gen.TINY_SPECTRAL)                 # models.py:27-28
        _close(real, z[tag + "real_images"])
        fake = pg.generator(step.params, latents, labels).detach()
        _close(fake, z[tag + "fake_images"])
        _close(osp.convert_to_waveform(fake[:, 0], fake[:, 1], **gen.TINY_SPECTRAL), z[tag + "fake_waveforms"])   # models.py:30-31
        # every session.run evaluates both losses on its batch
        _close(omodels.discriminator_loss(pg, step.params, real, labels, latents, hyper).detach(), z[tag + "discriminator_loss"])
        _close(omodels.generator_loss(pg, step.params, labels, latents, hyper).detach(), z[tag + "generator_loss"])
        if which == "discriminator":
            _, grads = step.discriminator_update(real, labels, latents)
        else:
            _, grads = step.generator_update(labels, latents)
        names = [k[len(tag) + 5:] for k in z.files if k.startswith(tag + "grad:")]
        assert set(names) == set(grads)
        for name in names:
            _grad_close(grads[name].numpy(), z[tag + "grad:" + name])
            _close(step.params[name].detach(), z[tag + "var:" + name])                      # tf.train.AdamOptimizer
    assert step.global_step == int(z["final_global_step"]) == int(z["iterations"])          # only the G op counts steps


# ------------------------------------------------------------------------------------------------ models.py PitchClassifier
def test_classifier_training_and_export_head():
    z = _load("reference_classifier")
    flat = lambda d: {n: (v.reshape(-1) if n.endswith(("/beta", "/gamma")) else v) for n, v in d.items()}
    resnet = onet.ResNet(**gen.TINY_RESNET)
    h = gen.CLASSIFIER_HYPER
    step = omodels.PitchClassifierStep(resnet, flat(_variables(z, "var0:")), h["weight_decay"], h["momentum"], h["use_nesterov"])
    hits = seen = 0
    for run in range(int(z["iterations"])):
        tag = "run%d:" % run
        waves, labels = _t(z[tag + "waveforms"]), _t(z[tag + "labels"])
        images = omodels.real_images_from_waveforms(waves, gen.TINY_SPECTRAL)
        before = {n: p.detach().clone() for n, p in step.params.items()}
        lr = h["base_learning_rate"] * h["decay_rate"] ** (run / h["decay_steps"])          # pitch_classifier_main.py:69-74
        total, _, logits = step.update(images, labels, lr)
        _close(total, z[tag + "loss"])
        hits += int((logits.argmax(dim=1) == labels.argmax(dim=1)).sum())
        seen += labels.shape[0]
        _close(hits / seen, z[tag + "accuracy"])                                            # tf.metrics.accuracy: running
        for name in before:
            _close(step.params[name].detach(), z[tag + "var:" + name].reshape(before[name].shape))
    features, logits = resnet(step.params, _t(z["images"]))
    _close(features.detach(), z["features"])
    _close(logits.detach(), z["logits"])


# ------------------------------------------------------------------------------------------------ the benchmark's own size
def _full_variables():
    with open(os.path.join(GOLDEN, "reference_variables.json")) as f:
        shapes = json.load(f)["gan_synth"]
    return {n: gen.named_value(n, s) for n, s in shapes.items()}


@pytest.mark.skipif(os.environ.get("GS_FULL_PIN", "0") != "1",
                    reason="about two minutes and 25 GB of float64 on the CPU: GS_FULL_PIN=1 (profiles/full_pin_r2.txt has a run)")
def test_full_size_step_oracle():
    """One session.run of the reference's GANSynth at gan_synth_main.py's own configuration (fully grown 128x1024, batch 8)
    against the oracle, float64: images, features, selected logits, both losses, every gradient's norm / maximum / samples."""
    z = _load("reference_full_step")
    params = _full_variables()
    assert sorted(str(n) for n in z["variable_names"] if str(n) != "global_step") == sorted(n for n in params if "color_block" not in n or "128x1024" in n)
    waves, labels, latents = gen.full_inputs()
    pg = onet.PGGAN(growing_level=1.0, **gen.FULL)
    step = omodels.GANSynthStep(pg, params, gen.HYPER)
    real = omodels.real_images_from_waveforms(waves, gen.FULL_SPECTRAL)
    _close(real[:, :, ::4, ::16], z["real_images_sub"])
    with torch.no_grad():
        fake = pg.generator(step.params, latents, labels)
        _close(fake[:, :, ::4, ::16], z["fake_images_sub"])
        _close(osp.convert_to_waveform(fake[:, 0], fake[:, 1], **gen.FULL_SPECTRAL)[:, ::64], z["fake_waveforms_sub"])
        features, logits = pg.discriminator(step.params, fake, labels)
        _close(features, z["fake_features"])
        _close(omodels.select_logits(logits, labels), z["fake_logits"])
        features, logits = pg.discriminator(step.params, real, labels)
        _close(features, z["real_features"])
        _close(omodels.select_logits(logits, labels), z["real_logits"])
    d_loss, d_grads = step.discriminator_update(real, labels, latents, apply=False)
    g_loss, g_grads = step.generator_update(labels, latents, apply=False)
    _close(d_loss, z["discriminator_loss"])
    _close(g_loss, z["generator_loss"])
    for name, grad in {**d_grads, **g_grads}.items():
        if "grad:" + name in z.files:
            want = z["grad:" + name]
            _grad_close(gen.grad_summary(grad.detach()), want, 1e-7 * max(1.0, want[1] and want[0] / want[1]))


# ------------------------------------------------------------------------------------------------ BASELINE configs[0]
def test_baseline_config1_sequence():
    """`gan_synth_main.py --train` with the 2-stage PGGAN at batch 4 as the reference runs it (three D-run / G-run
    iterations, growth levels 0, 1/4, 2/4): the oracle reproduces both losses, every applied gradient and every updated
    variable (norm, maximum, 16 samples each) from the same named weights."""
    import reference_vectors as rv
    z = _load("reference_config1")
    for run, which, waves, labels, latents, pg, step in rv.config1_oracle(z):
        tag = "run%d:" % run
        real = omodels.real_images_from_waveforms(waves, gen.CONFIG1_SPECTRAL)
        with torch.no_grad():
            _close(pg.generator(step.params, latents, labels), z[tag + "fake_images"])
        _close(omodels.discriminator_loss(pg, step.params, real, labels, latents, gen.HYPER).detach(), z[tag + "discriminator_loss"])
        _close(omodels.generator_loss(pg, step.params, labels, latents, gen.HYPER).detach(), z[tag + "generator_loss"])
        if run > 0:                                           # the variables the previous run's optimizer moved
            prev, scope = "run%d:" % (run - 1), ("generator", "discriminator")[run % 2]
            for name, p in step.params.items():
                if name.startswith(scope + "/"):
                    _grad_close(gen.grad_summary(p.detach()), z[prev + "var:" + name], 1e-9)
        grads = (step.discriminator_update(real, labels, latents, apply=False) if which == "discriminator"
                 else step.generator_update(labels, latents, apply=False))[1]
        for name, grad in grads.items():
            want = z[tag + "grad:" + name]
            if want[1] == 0.0:
                assert float(grad.abs().max()) == 0.0, name
            else:
                _grad_close(gen.grad_summary(grad), want, 1e-7 * want[0] / want[1])


def test_product_baseline_config1_sequence(emu):
    import reference_vectors as rv
    rv.check_config1(emu, "cpu")


# ------------------------------------------------------------------------------------------------ metrics.py
def test_evaluation_statistics():
    """gansynth_b200.metrics against values the reference's metrics.py (plain numpy / scipy) returned for the same inputs."""
    from gansynth_b200 import metrics
    z = _load("reference_metrics")
    _close(metrics.softmax(z["logits"]), z["softmax"], 1e-14)
    _close(metrics.kl_divergence(metrics.softmax(z["logits"][:8]), metrics.softmax(z["logits"][8:16])), z["kl"], 1e-13)
    _close(metrics.inception_score(z["logits"]), z["inception_score"], 1e-13)
    _close(metrics.frechet_inception_distance(z["real"], z["fake"]), z["frechet_inception_distance"], 1e-9)
    assert np.array_equal(metrics.binomial_proportion_test(z["props_p"], 400, z["props_q"], 300, 0.05), z["binomial"])


# ------------------------------------------------------------------------------------------------ the product's host logic
def test_product_forward_on_reference_vectors(emu):
    import reference_vectors as rv
    rv.check_forward(emu, "cpu")


def test_product_spectral_on_reference_vectors(emu):
    import reference_vectors as rv
    rv.check_spectral("cpu")


@pytest.mark.parametrize("fixture", ["reference_step", "reference_step_fake_penalty"])
def test_product_training_sequence_on_reference_vectors(emu, fixture):
    import reference_vectors as rv
    rv.check_training_sequence(emu, "cpu", fixture)


@pytest.mark.skipif(os.environ.get("GS_FULL_PIN", "0") != "1", reason="minutes of CPU: GS_FULL_PIN=1")
def test_product_full_size_step_on_reference_vectors(emu):
    import reference_vectors as rv
    rv.check_full_step(emu, "cpu", 1e-2)


def test_product_classifier_on_reference_vectors(emu):
    import reference_vectors as rv
    rv.check_classifier(emu, "cpu")


# ------------------------------------------------------------------------------------------------ provenance
@pytest.mark.skipif(not os.path.exists(gen.REFERENCE), reason="reference tree not present")
def test_committed_vectors_are_what_the_reference_says():
    """Re-runs the reference (every case of the generator) and compares with the committed files; then once more in
    float32 -- the precision TensorFlow itself would use -- against the float32 oracle at fp32 round-off."""
    for name, fn in gen.CASES.items():
        fresh, committed = fn(), _load(name)
        assert set(fresh) == set(committed.files), name
        for key in committed.files:
            if committed[key].dtype.kind in "fc":
                _close(fresh[key], committed[key], 1e-12)
            else:
                assert np.array_equal(fresh[key], committed[key]), (name, key)
    z32 = gen.pggan_forward(dtype=torch.float32)
    params = {k[4:]: _t(v) for k, v in z32.items() if k.startswith("var:")}
    for k, level in enumerate(z32["levels"]):
        pg = onet.PGGAN(growing_level=float(level), **gen.TINY)
        _close(pg.generator(params, _t(z32["latents"]), _t(z32["labels"])), z32["fake_images_%d" % k], 3e-5)
        _, logits = pg.discriminator(params, _t(z32["images"]), _t(z32["labels"]))
        _close(logits, z32["logits_%d" % k], 1e-4)


@pytest.mark.skipif(not os.path.exists(gen.REFERENCE), reason="reference tree not present")
def test_oracle_follows_the_reference_over_configurations():
    """Live, no fixture: seeded random architectures (seed resolution, number of doublings, channel clamp, latent size,
    label count, growth level, batch) through the reference's networks.py and through the oracle with the same variables."""
    for cfg, latent, classes, level, latents, labels, images, params, fake, features, logits in gen.random_architectures():
        ours = onet.PGGAN(growing_level=level, **cfg)
        g_table, d_table = ours.variable_shapes(latent, classes)
        assert {n: tuple(s) for n, (s, _) in {**g_table, **d_table}.items()} == {n: tuple(v.shape) for n, v in params.items()}, cfg
        _close(ours.generator(params, latents, labels), fake)
        got_features, got_logits = ours.discriminator(params, images, labels)
        _close(got_features, features)
        _close(got_logits, logits)


@pytest.mark.skipif(not os.path.exists(gen.REFERENCE), reason="reference tree not present")
def test_product_host_logic_follows_the_reference_over_configurations(emu):
    """The same architectures through the product's networks.py over emulated kernels (fp32): names, shapes, values."""
    import gansynth_b200.models as pmodels
    import gansynth_b200.networks as pnet
    import gansynth_b200.ops as pops
    for cfg, latent, classes, level, latents, labels, images, params, fake, features, logits in gen.random_architectures():
        store = pops.set_default_store(pops.VariableStore(device="cpu", seed=0))
        pmodels.reset_global_step()
        pg = pnet.PGGAN(growing_level=level, **cfg)
        pg._ensure_variables("generator", latent, classes)
        pg._ensure_variables("discriminator", 0, classes)
        assert {n: tuple(v.shape) for n, v in store.vars.items()} == {n: tuple(v.shape) for n, v in params.items()}, cfg
        store.load({n: v.float() for n, v in params.items()})
        with torch.no_grad():
            _close(pg.generator(latents.float(), labels.float()), fake, 2e-4)
            got_features, got_logits = pg.discriminator(images.float(), labels.float())
        _close(got_features, features, 2e-4)
        _close(got_logits, logits, 2e-4)


def test_spectral_configurations_from_fixture():
    """Spectrogram shapes, overlaps 0.5 / 0.75 / 0.875, three sample rates, waveform lengths with front padding from 0 to
    most of a frame (tests/golden/reference_spectral_configs.npz): the oracle both ways."""
    import reference_vectors as rv
    for params, waves, logmel, inst, back in rv.fixture_spectral_configs():
        got_lm, got_if = osp.convert_to_spectrogram(waves, **params)
        _close(got_lm, logmel)
        _close(got_if, inst)
        _close(osp.convert_to_waveform(got_lm, got_if, **params), back)


def test_product_spectral_configurations_from_fixture(emu):
    import reference_vectors as rv
    rv.check_spectral_configs("cpu")
