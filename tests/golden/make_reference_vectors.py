"""Golden vectors produced by the reference's OWN Python, run here.

TensorFlow 1.13 cannot be installed in this image, but skmhrk1209/GANSynth is a composition of TensorFlow primitives.
This script puts `oracle/tf1_eager/` (an eager PyTorch-CPU restatement of exactly those primitives) and `/root/reference`
on sys.path and imports the reference's unmodified `networks.py`, `spectral_ops.py` and `models.py`: its scopes, variable
names and shapes, weight scaling, block order, tf.cond growth logic, loss terms, tf.gradients penalties and optimizer calls
execute as written.  What they compute goes into `tests/golden/reference_*.npz` / `.json` (float64 unless said otherwise,
a few hundred kB in all); `tests/test_reference_pin_cpu.py` holds the oracle to these files on any box and, where
`/root/reference` exists, re-runs this script's cases and checks that the committed files still say what the reference says.

    python tests/golden/make_reference_vectors.py            # rewrites the fixtures (needs /root/reference)

One construction of the reference's GANSynth / PitchClassifier object == one session.run of its graph (see the stand-in's
header), so the training sequence `session.run(discriminator_train_op); session.run(generator_train_op)`
(models.py:189-192) is: construct, run the D op; construct again on the next batch, run the G op.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"

# the fixture configurations: three doublings of a non-square 2x4 seed, like the reference's 2x16 -> 128x1024 in small
TINY = dict(min_resolution=[2, 4], max_resolution=[16, 32], min_channels=4, max_channels=16)
TINY_SPECTRAL = dict(waveform_length=300, sample_rate=16000, spectrogram_shape=[16, 32], overlap=0.75)
FULL_SPECTRAL = dict(waveform_length=64000, sample_rate=16000, spectrogram_shape=[128, 1024], overlap=0.75)
LATENT, LABELS, BATCH = 8, 5, 4
HYPER = dict(generator_learning_rate=8e-4, generator_beta1=0.0, generator_beta2=0.99,
             discriminator_learning_rate=8e-4, discriminator_beta1=0.0, discriminator_beta2=0.99,
             mode_seeking_loss_weight=0.1, real_gradient_penalty_weight=5.0, fake_gradient_penalty_weight=0.0)
TINY_RESNET = dict(conv_param=dict(filters=8, kernel_size=[7, 7], strides=[2, 2]),
                   pool_param=dict(kernel_size=[3, 3], strides=[2, 2]),
                   residual_params=[dict(filters=8, strides=[1, 1], blocks=2), dict(filters=16, strides=[2, 2], blocks=2)],
                   groups=4, classes=LABELS)
CLASSIFIER_HYPER = dict(weight_decay=1e-4, momentum=0.9, use_nesterov=True, base_learning_rate=0.05, decay_steps=3.0,
                        decay_rate=0.1)
GROWING_STEPS = 7                       # levels 0, 1/7, 2/7 over the three recorded iterations
LEVELS = [0.0, 0.01, 0.3, 0.62, 1.0]    # forward-only cases: every tf.cond arm of networks.py:125-152 / 260-287


def reference_modules():
    """(tf stand-in, reference networks, spectral_ops, models, utils.Dict), imported from /root/reference."""
    for p in (REFERENCE, os.path.join(ROOT, "oracle", "tf1_eager")):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    shadowed = {n: sys.modules.pop(n) for n in ("tensorflow", "tensorflow_probability", "tensorflow_hub", "networks", "ops",
                                                "spectral_ops", "models", "metrics", "utils") if n in sys.modules}
    try:
        import tensorflow as tf
        import networks
        import spectral_ops
        import models
        from utils import Dict
        for m in (networks, spectral_ops, models):
            assert os.path.dirname(os.path.abspath(m.__file__)) == REFERENCE, m.__file__
        return tf, networks, spectral_ops, models, Dict
    finally:
        for n in ("tensorflow", "tensorflow_probability", "tensorflow_hub", "networks", "ops", "spectral_ops", "models",
                  "metrics", "utils"):
            sys.modules.pop(n, None)
        sys.modules.update(shadowed)
        for p in (REFERENCE, os.path.join(ROOT, "oracle", "tf1_eager")):
            sys.path.remove(p)


def _np(t):
    return t.detach().cpu().numpy().copy()


def _values(tf, trainable_only=True):
    return {n: _np(v.t) for n, v in tf.variables().items() if v.trainable or not trainable_only}


def _perturb_biases(tf, seed):
    """The reference starts biases / beta at 0 and gamma at 1; the fixtures move them so that they matter."""
    gen = torch.Generator().manual_seed(seed)
    for name, var in tf.variables().items():
        if name.endswith(("/bias", "/beta")):
            var.assign(0.1 * torch.randn(var.t.shape, generator=gen, dtype=torch.float64).to(var.t.dtype))
        elif name.endswith("/gamma"):
            var.assign(1.0 + 0.2 * torch.randn(var.t.shape, generator=gen, dtype=torch.float64).to(var.t.dtype))


def _inputs(gen, waveform_length, dtype):
    waves = (0.3 * torch.randn(BATCH, waveform_length, generator=gen, dtype=torch.float64)).to(dtype)
    labels = torch.nn.functional.one_hot(torch.randint(0, LABELS, (BATCH,), generator=gen), LABELS).to(dtype)
    return waves, labels


# ------------------------------------------------------------------------------------------------ cases
def pggan_forward(dtype=torch.float64):
    """networks.py PGGAN.generator / .discriminator at every growth regime."""
    tf, networks, _, _, _ = reference_modules()
    tf.reset_default_graph()
    tf.set_float_dtype(dtype)
    tf.set_random_seed(0)
    gen = torch.Generator().manual_seed(11)
    latents = torch.randn(BATCH, LATENT, generator=gen, dtype=torch.float64).to(dtype)
    labels = torch.nn.functional.one_hot(torch.tensor([0, 3, 1, 3]), LABELS).to(dtype)
    images = (0.5 * torch.randn(BATCH, 2, *TINY["max_resolution"], generator=gen, dtype=torch.float64)).to(dtype)
    out = dict(latents=_np(latents), labels=_np(labels), images=_np(images), levels=np.asarray(LEVELS))
    level = tf.Tensor(torch.tensor(0.5, dtype=dtype))
    pggan = networks.PGGAN(growing_level=level, **TINY)
    tf.build_all_branches(True)                               # graph construction: every variable of every branch
    pggan.generator(tf.Tensor(latents), tf.Tensor(labels))
    pggan.discriminator(tf.Tensor(images), tf.Tensor(labels))
    tf.build_all_branches(False)
    _perturb_biases(tf, 12)
    for name, value in _values(tf).items():
        out["var:" + name] = value
    for k, lv in enumerate(LEVELS):
        level.t = torch.tensor(lv, dtype=dtype)
        pggan = networks.PGGAN(growing_level=level, **TINY)   # growing_depth is computed in the constructor (networks.py:29)
        fake = pggan.generator(tf.Tensor(latents), tf.Tensor(labels))
        features, logits = pggan.discriminator(tf.Tensor(images), tf.Tensor(labels))
        out["growing_depth_%d" % k] = _np(pggan.growing_depth.t)
        out["fake_images_%d" % k] = _np(fake.t)
        out["features_%d" % k] = _np(features.t)
        out["logits_%d" % k] = _np(logits.t)
    return out


def variable_tables():
    """Name -> shape of every variable the reference's two command lines create (gan_synth_main.py:43-54,
    pitch_classifier_main.py:42-53), from building the graphs at full size with all tf.cond branches traced."""
    tf, networks, _, _, D = reference_modules()
    tables = {}
    tf.reset_default_graph()
    tf.set_float_dtype(torch.float32)
    tf.build_all_branches(True)
    with torch.no_grad():
        pggan = networks.PGGAN(min_resolution=[2, 16], max_resolution=[128, 1024], min_channels=32, max_channels=256,
                               growing_level=tf.Tensor(torch.tensor(0.5)))
        labels = tf.Tensor(torch.nn.functional.one_hot(torch.tensor([0, 1, 2, 3]), 61).float())
        pggan.generator(tf.Tensor(torch.zeros(4, 256)), labels)
        pggan.discriminator(tf.Tensor(torch.zeros(4, 2, 128, 1024)), labels)
        tables["gan_synth"] = {n: list(v.t.shape) for n, v in tf.variables().items()}
        tf.reset_default_graph()
        resnet = networks.ResNet(conv_param=D(filters=64, kernel_size=[7, 7], strides=[2, 2]),
                                 pool_param=D(kernel_size=[3, 3], strides=[2, 2]),
                                 residual_params=[D(filters=64, strides=[1, 1], blocks=3), D(filters=128, strides=[2, 2], blocks=4),
                                                  D(filters=256, strides=[2, 2], blocks=6), D(filters=512, strides=[2, 2], blocks=3)],
                                 groups=32, classes=61)
        resnet(tf.Tensor(torch.zeros(1, 2, 64, 128)))
        tables["pitch_classifier"] = {n: list(v.t.shape) for n, v in tf.variables().items()}
    tf.build_all_branches(False)
    return tables


def spectral(dtype=torch.float64):
    """spectral_ops.convert_to_spectrogram / convert_to_waveform at the reference's own size (sub-sampled) and at the
    fixture size (whole)."""
    import math
    tf, _, spectral_ops, _, _ = reference_modules()
    tf.set_float_dtype(dtype)
    t = torch.arange(64000, dtype=torch.float64) / 16000.0
    gen = torch.Generator().manual_seed(0)
    wave = torch.stack([0.3 * torch.sin(2 * math.pi * 440.0 * t) * torch.exp(-3 * t)
                        + 0.01 * torch.randn(64000, generator=gen, dtype=torch.float64),
                        0.1 * torch.randn(64000, generator=gen, dtype=torch.float64)]).float().to(dtype)
    logmel, inst = spectral_ops.convert_to_spectrogram(tf.Tensor(wave), **FULL_SPECTRAL)
    back = spectral_ops.convert_to_waveform(logmel, inst, **FULL_SPECTRAL)
    out = dict(full_wave=_np(wave).astype(np.float32), full_logmel_sub=_np(logmel.t)[:, ::8, ::16],
               full_inst_sub=_np(inst.t)[:, ::8, ::16], full_back_sub=_np(back.t)[:, ::32])
    small, _ = _inputs(torch.Generator().manual_seed(21), TINY_SPECTRAL["waveform_length"], dtype)
    logmel, inst = spectral_ops.convert_to_spectrogram(tf.Tensor(small), **TINY_SPECTRAL)
    back = spectral_ops.convert_to_waveform(logmel, inst, **TINY_SPECTRAL)
    out.update(tiny_wave=_np(small), tiny_logmel=_np(logmel.t), tiny_inst=_np(inst.t), tiny_back=_np(back.t))
    return out


def gan_step(fake_penalty=0.0, iterations=3, dtype=torch.float64):
    """models.py GANSynth under gan_synth_main.py's growth schedule (level = global_step / growing_steps) for
    `iterations` x (D run, G run): inputs, both losses and the applied gradients of every run, variables after it."""
    tf, networks, _, models, Dict = reference_modules()
    tf.reset_default_graph()
    tf.set_float_dtype(dtype)
    tf.set_random_seed(0)
    hyper = dict(HYPER, fake_gradient_penalty_weight=fake_penalty)
    gen = torch.Generator().manual_seed(31 + int(fake_penalty * 10))
    feed = {}

    def real_input_fn():
        waves, labels = _inputs(gen, TINY_SPECTRAL["waveform_length"], dtype)
        feed["waveforms"], feed["labels"] = waves, labels
        return tf.Tensor(waves.clone().requires_grad_(True)), tf.Tensor(labels)   # tf.gradients reaches real_images

    def fake_input_fn():
        z = tf.random.normal([BATCH, LATENT])
        feed["latents"] = z.t.detach()
        return z

    def construct():
        pggan = networks.PGGAN(growing_level=tf.cast(tf.divide(x=tf.train.get_or_create_global_step(), y=GROWING_STEPS),
                                                     tf.float32), **TINY)       # gan_synth_main.py:48-53
        return models.GANSynth(generator=pggan.generator, discriminator=pggan.discriminator, real_input_fn=real_input_fn,
                               fake_input_fn=fake_input_fn, spectral_params=Dict(TINY_SPECTRAL), hyper_params=Dict(hyper))

    tf.build_all_branches(True)
    construct()                                               # graph construction; its draw of inputs is discarded
    tf.build_all_branches(False)
    _perturb_biases(tf, 32)
    out = dict(fake_penalty=np.asarray(fake_penalty), iterations=np.asarray(iterations))
    for name, value in _values(tf).items():
        out["var0:" + name] = value
    run = 0
    for _ in range(iterations):
        for which in ("discriminator", "generator"):
            model = construct()
            op = getattr(model, which + "_train_op")
            tag = "run%d:" % run
            out[tag + "which"] = np.asarray(which)
            out[tag + "global_step"] = _np(tf.train.get_or_create_global_step().t)
            for k in ("waveforms", "labels", "latents"):
                out[tag + k] = _np(feed[k])
            out[tag + "real_images"] = _np(model.real_images.t)
            out[tag + "fake_images"] = _np(model.fake_images.t)
            out[tag + "fake_waveforms"] = _np(model.fake_waveforms.t)
            out[tag + "generator_loss"] = _np(model.generator_loss.t)
            out[tag + "discriminator_loss"] = _np(model.discriminator_loss.t)
            for grad, var in op.grads_and_vars:
                out[tag + "grad:" + var.op.name] = np.zeros(var.t.shape) if grad is None else _np(grad)
            op.run()
            for name, value in _values(tf).items():
                if name.startswith(which):
                    out[tag + "var:" + name] = value
            run += 1
    out["final_global_step"] = _np(tf.train.get_or_create_global_step().t)
    return out


def classifier_step(iterations=2, dtype=torch.float64):
    """networks.py ResNet + models.py PitchClassifier: loss (cross entropy + L2), Nesterov momentum under an
    exponentially decaying learning rate, `iterations` train-op runs."""
    tf, networks, _, models, Dict = reference_modules()
    tf.reset_default_graph()
    tf.set_float_dtype(dtype)
    tf.set_random_seed(0)
    gen = torch.Generator().manual_seed(41)
    feed = {}

    def input_fn():
        waves, labels = _inputs(gen, TINY_SPECTRAL["waveform_length"], dtype)
        feed["waveforms"], feed["labels"] = waves, labels
        return tf.Tensor(waves), tf.Tensor(labels)

    resnet = networks.ResNet(conv_param=Dict(TINY_RESNET["conv_param"]), pool_param=Dict(TINY_RESNET["pool_param"]),
                             residual_params=[Dict(p) for p in TINY_RESNET["residual_params"]],
                             groups=TINY_RESNET["groups"], classes=TINY_RESNET["classes"])
    h = CLASSIFIER_HYPER
    hyper = Dict(weight_decay=h["weight_decay"], momentum=h["momentum"], use_nesterov=h["use_nesterov"],
                 learning_rate=lambda global_step: tf.train.exponential_decay(                   # pitch_classifier_main.py:69-74
                     learning_rate=h["base_learning_rate"], global_step=global_step, decay_steps=h["decay_steps"],
                     decay_rate=h["decay_rate"]))

    def construct():
        return models.PitchClassifier(network=resnet, input_fn=input_fn, spectral_params=Dict(TINY_SPECTRAL), hyper_params=hyper)

    construct()                                               # graph construction
    tf.local_variables_initializer()                          # the Scaffold's local_init_op (models.py:311-316)
    _perturb_biases(tf, 42)
    out = dict(iterations=np.asarray(iterations))
    for name, value in _values(tf).items():
        out["var0:" + name] = value
    for run in range(iterations):
        model = construct()
        tag = "run%d:" % run
        out[tag + "waveforms"], out[tag + "labels"] = _np(feed["waveforms"]), _np(feed["labels"])
        out[tag + "loss"] = _np(model.loss.t)
        out[tag + "accuracy"] = _np(model.update_op.t)       # what evaluate() returns (models.py:405-407): running
        for grad, var in model.train_op.grads_and_vars:
            out[tag + "grad:" + var.op.name] = _np(grad)
        model.train_op.run()
        for name, value in _values(tf).items():
            out[tag + "var:" + name] = value
    # the inference head the reference exports: features / logits of given images (models.py:300-304)
    images = 0.5 * torch.randn(BATCH, 2, *TINY_SPECTRAL["spectrogram_shape"], generator=gen, dtype=torch.float64).to(dtype)
    features, logits = resnet(tf.Tensor(images))
    out.update(images=_np(images), features=_np(features.t), logits=_np(logits.t))
    return out


# ------------------------------------------------------------------------------------------------ the benchmark's own size
FULL = dict(min_resolution=[2, 16], max_resolution=[128, 1024], min_channels=32, max_channels=256)   # gan_synth_main.py:43-47
FULL_BATCH, FULL_LATENT, FULL_LABELS = 8, 256, 61
GRAD_SAMPLES = 16


def named_value(name, shape):
    """Value of a full-size variable as a function of its NAME alone (64 MB of weights cannot be committed; both the
    generator below and the tests rebuild them from this): truncated normal for weights, 0.1 N(0, 1) for biases, rounded
    to float32 so that every precision starts from the same numbers."""
    seed = sum((i + 1) * b for i, b in enumerate(name.encode())) % (2 ** 31)
    rng = torch.Generator().manual_seed(seed)
    if name.endswith("/bias"):
        t = 0.1 * torch.randn(list(shape), generator=rng, dtype=torch.float64)
    else:
        t = torch.empty(list(shape), dtype=torch.float64)
        torch.nn.init.trunc_normal_(t, 0.0, 1.0, -2.0, 2.0, generator=rng)
    return t.float().double()


def full_inputs():
    rng = torch.Generator().manual_seed(61)
    t = torch.arange(64000, dtype=torch.float64) / 16000.0
    freqs = 110.0 * 2.0 ** (torch.arange(FULL_BATCH, dtype=torch.float64) / 2.0)
    waves = 0.4 * torch.sin(2 * np.pi * freqs[:, None] * t[None, :]) * torch.exp(-2.0 * t)[None, :] \
        + 0.02 * torch.randn(FULL_BATCH, 64000, generator=rng, dtype=torch.float64)
    labels = torch.nn.functional.one_hot(torch.randint(0, FULL_LABELS, (FULL_BATCH,), generator=rng), FULL_LABELS).double()
    latents = torch.randn(FULL_BATCH, FULL_LATENT, generator=rng, dtype=torch.float64)
    return waves.float().double(), labels, latents.float().double()


def grad_summary(grad):
    """(L2 norm, largest magnitude, GRAD_SAMPLES evenly spaced elements of the flattened TF-layout gradient)."""
    flat = grad.reshape(-1).double()
    idx = torch.linspace(0, flat.numel() - 1, GRAD_SAMPLES).long()
    return np.concatenate([[float(flat.norm()), float(flat.abs().max())], flat[idx].numpy()])


def full_step(dtype=torch.float64):
    """ONE session.run of models.GANSynth at the configuration of gan_synth_main.py (fully grown 128x1024 networks,
    batch 8, the command line's hyper-parameters): images, logits-derived losses and a summary of every gradient.
    About a minute and 20 GB in float64; not part of CASES (run with `--full`)."""
    tf, networks, _, models, Dict = reference_modules()
    tf.reset_default_graph()
    tf.set_float_dtype(dtype)
    waves, labels, latents = full_inputs()
    holder = {}

    def real_input_fn():
        return tf.Tensor(waves.to(dtype).requires_grad_(True)), tf.Tensor(labels.to(dtype))

    def fake_input_fn():
        holder["z"] = tf.Tensor(latents.to(dtype).requires_grad_(True))
        return holder["z"]

    def construct():
        pggan = networks.PGGAN(growing_level=tf.Tensor(torch.tensor(1.0, dtype=dtype)), **FULL)
        return models.GANSynth(generator=pggan.generator, discriminator=pggan.discriminator, real_input_fn=real_input_fn,
                               fake_input_fn=fake_input_fn, spectral_params=Dict(FULL_SPECTRAL), hyper_params=Dict(HYPER))

    # graph construction on a small stand-in batch is not possible (shapes are the graph's), so the variables of the grown
    # path are created by a first evaluation and then given their named values
    with torch.no_grad():
        pggan = networks.PGGAN(growing_level=tf.Tensor(torch.tensor(1.0, dtype=dtype)), **FULL)
        fake = pggan.generator(tf.Tensor(latents.to(dtype)), tf.Tensor(labels.to(dtype)))
        pggan.discriminator(fake, tf.Tensor(labels.to(dtype)))
    for name, var in tf.variables().items():
        var.assign(named_value(name, var.t.shape).to(dtype))
    model = construct()
    out = dict(variable_names=np.asarray(list(tf.variables())),
               real_images_sub=_np(model.real_images.t)[:, :, ::4, ::16], fake_images_sub=_np(model.fake_images.t)[:, :, ::4, ::16],
               fake_waveforms_sub=_np(model.fake_waveforms.t)[:, ::64],
               real_logits=_np(model.real_logits.t), fake_logits=_np(model.fake_logits.t),
               real_features=_np(model.real_features.t), fake_features=_np(model.fake_features.t),
               generator_loss=_np(model.generator_loss.t), discriminator_loss=_np(model.discriminator_loss.t))
    for which in ("discriminator", "generator"):
        for grad, var in getattr(model, which + "_train_op").grads_and_vars:
            out["grad:" + var.op.name] = grad_summary(torch.zeros_like(var.t) if grad is None else grad)
    return out


# ------------------------------------------------------------------------------------------------ BASELINE config 1
CONFIG1 = dict(min_resolution=[4, 4], max_resolution=[16, 16], min_channels=32, max_channels=256)
CONFIG1_SPECTRAL = dict(waveform_length=152, sample_rate=16000, spectrogram_shape=[16, 16], overlap=0.75)
CONFIG1_BATCH, CONFIG1_GROWING_STEPS = 4, 4


def config1_inputs(run):
    rng = torch.Generator().manual_seed(71 + run)
    waves = 0.3 * torch.randn(CONFIG1_BATCH, CONFIG1_SPECTRAL["waveform_length"], generator=rng, dtype=torch.float64)
    labels = torch.nn.functional.one_hot(torch.randint(0, FULL_LABELS, (CONFIG1_BATCH,), generator=rng), FULL_LABELS).double()
    latents = torch.randn(CONFIG1_BATCH, FULL_LATENT, generator=rng, dtype=torch.float64)
    return waves.float().double(), labels, latents.float().double()


def config1_sequence(iterations=3, dtype=torch.float64):
    """BASELINE.json configs[0] -- `gan_synth_main.py --train` with the 2-stage PGGAN (4x4 -> 16x16), batch 4 -- as the
    reference itself runs it: D run, G run, ... with the growth level following global_step (levels 0, 1/4, 2/4: nothing
    grown, first blend, second blend).  The 8 MB of weights are `named_value`s and evolve by the reference's own Adam
    updates, so the file keeps summaries only: both losses of every run, norm / maximum / samples of the applied
    gradients and of the updated variables."""
    tf, networks, _, models, Dict = reference_modules()
    tf.reset_default_graph()
    tf.set_float_dtype(dtype)
    tf.set_random_seed(0)
    state = dict(run=0)

    def real_input_fn():
        waves, labels, _ = config1_inputs(state["run"])
        return tf.Tensor(waves.to(dtype).requires_grad_(True)), tf.Tensor(labels.to(dtype))

    def fake_input_fn():
        return tf.Tensor(config1_inputs(state["run"])[2].to(dtype).requires_grad_(True))

    def construct():
        pggan = networks.PGGAN(growing_level=tf.cast(tf.divide(x=tf.train.get_or_create_global_step(), y=CONFIG1_GROWING_STEPS),
                                                     tf.float32), **CONFIG1)
        return models.GANSynth(generator=pggan.generator, discriminator=pggan.discriminator, real_input_fn=real_input_fn,
                               fake_input_fn=fake_input_fn, spectral_params=Dict(CONFIG1_SPECTRAL), hyper_params=Dict(HYPER))

    tf.build_all_branches(True)
    construct()
    tf.build_all_branches(False)
    for name, var in tf.variables().items():
        if var.trainable:
            var.assign(named_value(name, var.t.shape).to(dtype))
    out = dict(iterations=np.asarray(iterations),
               variable_names=np.asarray([n for n, v in tf.variables().items() if v.trainable]),
               variable_sizes=np.asarray([v.t.numel() for n, v in tf.variables().items() if v.trainable]))
    for run in range(2 * iterations):
        which = ("discriminator", "generator")[run % 2]
        state["run"] = run
        model = construct()
        op = getattr(model, which + "_train_op")
        tag = "run%d:" % run
        out[tag + "global_step"] = _np(tf.train.get_or_create_global_step().t)
        out[tag + "generator_loss"] = _np(model.generator_loss.t)
        out[tag + "discriminator_loss"] = _np(model.discriminator_loss.t)
        out[tag + "fake_images"] = _np(model.fake_images.t)
        for grad, var in op.grads_and_vars:
            out[tag + "grad:" + var.op.name] = grad_summary(torch.zeros_like(var.t) if grad is None else grad)
        op.run()
        for grad, var in op.grads_and_vars:
            out[tag + "var:" + var.op.name] = grad_summary(var.t.detach())
    return out


def random_architectures(count=8):
    """Seeded random architectures run through the reference's networks.py, live: yields (cfg, latent size, classes,
    level, latents, labels, images, variables by name, the reference's fake images / features / logits), float64."""
    tf, networks, _, _, _ = reference_modules()
    tf.set_float_dtype(torch.float64)
    rng = np.random.default_rng(7)
    for case in range(count):
        tf.reset_default_graph()
        tf.set_random_seed(case)
        seed_res = [int(rng.choice([1, 2, 3, 4])), int(rng.choice([1, 2, 4, 5]))]
        doublings = int(rng.integers(1, 4))      # the reference cannot build a graph without at least one doubling
        cfg = dict(min_resolution=seed_res, max_resolution=[r << doublings for r in seed_res],
                   min_channels=int(rng.choice([2, 4, 6])), max_channels=int(rng.choice([8, 12, 64])))
        latent, classes, batch = int(rng.choice([3, 8])), int(rng.choice([2, 7])), 4 * int(rng.integers(1, 3))
        level = float(rng.choice([0.0, rng.uniform(0.0, 1.0), 1.0]))
        g = torch.Generator().manual_seed(100 + case)
        latents = torch.randn(batch, latent, generator=g, dtype=torch.float64)
        labels = torch.nn.functional.one_hot(torch.randint(0, classes, (batch,), generator=g), classes).double()
        images = torch.randn(batch, 2, *cfg["max_resolution"], generator=g, dtype=torch.float64)
        ref = networks.PGGAN(growing_level=tf.Tensor(torch.tensor(level, dtype=torch.float64)), **cfg)
        tf.build_all_branches(True)
        ref.generator(tf.Tensor(latents), tf.Tensor(labels))
        ref.discriminator(tf.Tensor(images), tf.Tensor(labels))
        tf.build_all_branches(False)
        _perturb_biases(tf, 200 + case)
        fake = ref.generator(tf.Tensor(latents), tf.Tensor(labels))
        features, logits = ref.discriminator(tf.Tensor(images), tf.Tensor(labels))
        params = {n: v.t.detach().clone() for n, v in tf.variables().items()}
        yield cfg, latent, classes, level, latents, labels, images, params, fake.t.detach(), features.t.detach(), logits.t.detach()


ODD_CASES = (1, 3, 6)      # 4x5 seed with 4..64 channels; 4x5 -> 32x40 with 6..12 channels; a 1x1 seed


def odd_architectures():
    """Three of `random_architectures` as a fixture (the others only run live)."""
    out = dict(cases=np.asarray(ODD_CASES))
    for case, (cfg, latent, classes, level, latents, labels, images, params, fake, features, logits) in enumerate(random_architectures()):
        if case not in ODD_CASES:
            continue
        tag = "case%d:" % case
        out[tag + "cfg"] = np.asarray(cfg["min_resolution"] + cfg["max_resolution"] + [cfg["min_channels"], cfg["max_channels"], latent, classes])
        out[tag + "level"] = np.asarray(level)
        for key, value in dict(latents=latents, labels=labels, images=images, fake_images=fake, features=features, logits=logits).items():
            out[tag + key] = _np(value)
        for name, value in params.items():
            out[tag + "var:" + name] = _np(value)
    return out


def random_spectral_configs(count=8):
    """Seeded random spectral configurations through the reference's spectral_ops.py, live: yields (parameters, waveforms,
    log-mel, IF, reconstructed waveforms), float64."""
    tf, _, spectral_ops, _, _ = reference_modules()
    tf.set_float_dtype(torch.float64)
    rng = np.random.default_rng(9)
    for case in range(count):
        bins = int(rng.choice([16, 32, 64, 256]))
        steps = int(rng.choice([4, 9, 16, 33]))
        overlap = float(rng.choice([0.5, 0.75, 0.875]))
        frame_step = int((1.0 - overlap) * 2 * bins)
        covered = frame_step * (steps - 1) + 2 * bins
        params = dict(waveform_length=int(covered - rng.integers(0, 2 * bins - 1)), sample_rate=int(rng.choice([8000, 16000, 44100])),
                      spectrogram_shape=[steps, bins], overlap=overlap)
        g = torch.Generator().manual_seed(300 + case)
        waves = 0.3 * torch.randn(3, params["waveform_length"], generator=g, dtype=torch.float64)
        logmel, inst = spectral_ops.convert_to_spectrogram(tf.Tensor(waves), **params)
        back = spectral_ops.convert_to_waveform(logmel, inst, **params)
        yield params, waves, logmel.t, inst.t, back.t


def spectral_configs():
    """All eight of `random_spectral_configs` as a fixture (they are small)."""
    out = {}
    for case, (params, waves, logmel, inst, back) in enumerate(random_spectral_configs()):
        tag = "case%d:" % case
        out[tag + "params"] = np.asarray([params["waveform_length"], params["sample_rate"], *params["spectrogram_shape"]])
        out[tag + "overlap"] = np.asarray(params["overlap"])
        for key, value in dict(waves=waves, logmel=logmel, inst=inst, back=back).items():
            out[tag + key] = _np(value)
    out["count"] = np.asarray(case + 1)
    return out


def metrics_case():
    """metrics.py is plain numpy / scipy / sklearn: imported and called as it is."""
    for name in ("metrics",):
        sys.modules.pop(name, None)
    sys.path.insert(0, REFERENCE)
    try:
        import metrics
        assert os.path.dirname(os.path.abspath(metrics.__file__)) == REFERENCE
    finally:
        sys.path.remove(REFERENCE)
        sys.modules.pop("metrics", None)
    import scipy.linalg  # noqa: F401  (metrics.py reaches scipy.linalg / scipy.stats through the top-level package)
    import scipy.stats  # noqa: F401
    rng = np.random.default_rng(51)
    logits = rng.normal(size=(64, LABELS)) * 3.0
    real = rng.normal(size=(400, 6)) @ rng.normal(size=(6, 6))
    fake = rng.normal(size=(300, 6)) @ rng.normal(size=(6, 6)) + 0.5
    p, q = metrics.softmax(logits[:8]), metrics.softmax(logits[8:16])
    props_p, props_q = rng.dirichlet(np.ones(10)), rng.dirichlet(np.ones(10))
    return dict(logits=logits, real=real, fake=fake, softmax=metrics.softmax(logits), kl=metrics.kl_divergence(p, q),
                inception_score=np.asarray(metrics.inception_score(logits)),
                frechet_inception_distance=np.asarray(metrics.frechet_inception_distance(real, fake)),
                props_p=props_p, props_q=props_q,
                binomial=metrics.binomial_proportion_test(props_p, 400, props_q, 300, 0.05))


CASES = dict(reference_metrics=metrics_case, reference_config1=config1_sequence, reference_spectral_configs=spectral_configs, reference_architectures=odd_architectures, reference_pggan=pggan_forward, reference_spectral=spectral, reference_step=gan_step,
             reference_step_fake_penalty=lambda: gan_step(fake_penalty=2.0, iterations=1),
             reference_classifier=classifier_step)


def main():
    if "--full" in sys.argv:
        arrays = full_step()
        path = os.path.join(HERE, "reference_full_step.npz")
        np.savez_compressed(path, **arrays)
        print("reference_full_step %d arrays %.1f kB" % (len(arrays), os.path.getsize(path) / 1e3))
        return
    for name, fn in CASES.items():
        arrays = fn()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **arrays)
        print("%-32s %4d arrays %8.1f kB" % (name, len(arrays), os.path.getsize(path) / 1e3))
    with open(os.path.join(HERE, "reference_variables.json"), "w") as f:
        json.dump(variable_tables(), f, indent=0, sort_keys=False)
    print("reference_variables.json written")


if __name__ == "__main__":
    main()
