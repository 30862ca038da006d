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


def test_fake_gradient_penalty_branch(emu):
    """models.py:50-54 (fake_gradient_penalty_weight != 0; 0.0 on the reference's command line): loss and D gradients."""
    import gansynth_b200.models as pmodels
    hp = dict(HYPER, fake_gradient_penalty_weight=2.5)
    opg, params, ppg, latents, labels, images = _pair(1.0, emu)
    ostep = omodels.GANSynthStep(opg, params, hp)
    model = pmodels.GANSynth(ppg.generator, ppg.discriminator, None, None, {}, hp, device="cpu")
    model._ensure_optimizers(labels, latents)
    want_loss, want_grads = ostep.discriminator_update(images, labels, latents, apply=False)
    base_loss, _ = omodels.GANSynthStep(opg, params, HYPER).discriminator_update(images, labels, latents, apply=False)
    assert abs(float(want_loss) - float(base_loss)) > 1e-6            # the branch contributes
    model._set_trainable("discriminator")
    loss = model.discriminator_loss_fn(images, labels, latents)
    assert abs(float(loss.detach()) - float(want_loss)) < 1e-4 * max(1.0, abs(float(want_loss)))
    model._backward("discriminator", loss)
    for n, g in emu.unflatten("discriminator", model._opt["discriminator"]["grad"]).items():
        assert grad_close(g, want_grads[n], 5e-4), n


@pytest.mark.parametrize("level", [0.05, 0.1, 0.3])
def test_device_resident_blend_weight_matches_host_floats(emu, level):
    """Growth-phase graphs keep lerp's (t, 1 - t) in device memory (PGGAN.lerp_coef, AxpbyDev): forward values and the
    gradients w.r.t. latents and images must equal the host-float path, and both must match the oracle."""
    opg, params, ppg, latents, labels, images = _pair(level, emu)
    latents = latents.clone().requires_grad_()
    images = images.clone().requires_grad_()
    host_g = ppg.generator(latents, labels)
    host_f, host_l = ppg.discriminator(images, labels)
    g_host = torch.autograd.grad(host_g.sum() + host_l.sum(), (latents, images))
    ppg.update_lerp_coef("cpu")
    ppg.device_lerp_active = True
    try:
        dev_g = ppg.generator(latents, labels)
        dev_f, dev_l = ppg.discriminator(images, labels)
        g_dev = torch.autograd.grad(dev_g.sum() + dev_l.sum(), (latents, images))
    finally:
        ppg.device_lerp_active = False
    assert rel_err(dev_g, host_g) < 1e-6 and rel_err(dev_l, host_l) < 1e-6 and rel_err(dev_f, host_f) < 1e-6
    for a, b in zip(g_dev, g_host):
        assert rel_err(a, b) < 1e-6
    assert rel_err(dev_g, opg.generator(params, latents.detach(), labels)) < TOL
    gd = ppg.growing_depth
    import math
    assert ppg.structure_key() == ("grow", math.ceil(gd))
    t = float(ppg.lerp_coef[0])
    assert abs(t - (math.ceil(gd) - gd)) < 1e-6 and abs(float(ppg.lerp_coef.sum()) - 1.0) < 1e-6


def test_structure_key_follows_the_blend_depth(emu):
    import gansynth_b200.networks as pnet
    keys = []
    for level in (0.0, 0.1, 1.0 / 7.0, 0.3, 3.0 / 7.0, 0.5, 1.0):
        keys.append(pnet.PGGAN(growing_level=level, **SMALL).structure_key())
    # SMALL has max_depth 2: growing_depth = log2(1 + 7 level) -> 0, 0.77, 1, 1.63, 2, 2.17, 3
    assert keys == [("grow", 0), ("grow", 1), ("grow", 1), ("grow", 2), ("grow", 2), ("grown",), ("grown",)]


def test_train_loop_checkpoints_and_scalar_summaries(emu, tmp_path):
    """GANSynth.train (models.py:110-194) on the CPU emulation backend: stops at total_steps, restores from its own
    checkpoint on the next call, writes one scalar-summary line per save_summary_steps iterations."""
    import json
    import gansynth_b200.models as pmodels
    import gansynth_b200.networks as pnet
    gs = pmodels.get_or_create_global_step()
    ppg = pnet.PGGAN(growing_level=gs / 8, **SMALL)
    g = torch.Generator().manual_seed(0)

    def real():
        return (torch.randn(4, 512, generator=g) * 0.5,
                torch.nn.functional.one_hot(torch.randint(0, 61, (4,), generator=g), 61).float())

    def build():
        m = pmodels.GANSynth(ppg.generator, ppg.discriminator, real, lambda: torch.randn(4, 256, generator=g), {}, HYPER,
                             device="cpu")
        m.real_images_from_waveforms = lambda w: w.reshape(4, 2, 16, 16)
        return m

    model = build()
    model.train(str(tmp_path), None, total_steps=3, save_checkpoint_steps=2, save_summary_steps=1, log_tensor_steps=0)
    assert int(gs.value) == 3
    lines = [json.loads(l) for l in open(tmp_path / "summaries.jsonl")]
    assert [l["global_step"] for l in lines] == [1, 2, 3]
    assert all(set(l) >= {"generator_loss", "discriminator_loss", "growing_depth"} for l in lines)
    assert lines[0]["growing_depth"] < lines[2]["growing_depth"]
    assert sorted(p.name for p in tmp_path.glob("model.ckpt-*.pt")) == ["model.ckpt-2.pt", "model.ckpt-3.pt"]
    # the scalars are also TensorBoard events under the reference's tags
    try:
        from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    except ImportError:
        EventAccumulator = None
    if EventAccumulator is not None:
        acc = EventAccumulator(str(tmp_path))
        acc.Reload()
        assert set(acc.Tags()["scalars"]) >= {"generator_loss", "discriminator_loss"}
        ev = acc.Scalars("generator_loss")
        assert [e.step for e in ev] == [1, 2, 3]
        assert abs(ev[2].value - lines[2]["generator_loss"]) < 1e-6 * max(1.0, abs(lines[2]["generator_loss"]))
    weights = {n: v.clone() for n, v in emu.state().items()}
    # a second run: nothing to do (restored at step 3), the weights are the checkpoint's
    gs.value = 0
    model2 = build()
    model2.train(str(tmp_path), None, total_steps=3, save_checkpoint_steps=2, save_summary_steps=1, log_tensor_steps=0)
    assert int(gs.value) == 3
    for n, v in emu.state().items():
        assert torch.equal(v, weights[n]), n


def test_media_summaries_on_the_real_spectral_shapes(emu, tmp_path):
    """models.py:131-161 on the CPU emulation backend with a small 1024-bin model (1x128 -> 8x1024 images, 5000-sample
    clips): the event file carries the loss scalars, 4 + 4 audio clips and the four image groups."""
    pytest.importorskip("tensorboard")
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    import gansynth_b200.models as pmodels
    import gansynth_b200.networks as pnet
    cfg = dict(min_resolution=[1, 128], max_resolution=[8, 1024], min_channels=32, max_channels=64)
    spectral = dict(waveform_length=5000, sample_rate=16000, spectrogram_shape=[8, 1024], overlap=0.75)
    ppg = pnet.PGGAN(growing_level=1.0, **cfg)
    g = torch.Generator().manual_seed(0)

    def real():
        return (0.1 * torch.randn(4, 5000, generator=g),
                torch.nn.functional.one_hot(torch.randint(0, 61, (4,), generator=g), 61).float())

    model = pmodels.GANSynth(ppg.generator, ppg.discriminator, real, lambda: torch.randn(4, 256, generator=g), spectral,
                             HYPER, device="cpu")
    model.media_summaries = True
    model.train(str(tmp_path), None, total_steps=1, save_checkpoint_steps=0, save_summary_steps=1, log_tensor_steps=0)
    assert model._tb, "summaries were disabled"
    acc = EventAccumulator(str(tmp_path), size_guidance={"audio": 0, "images": 0, "scalars": 0})
    acc.Reload()
    tags = acc.Tags()
    assert set(tags["scalars"]) >= {"generator_loss", "discriminator_loss"}
    assert sorted(tags["audio"]) == sorted("%s/%d" % (n, i) for n in ("real_waveforms", "fake_waveforms") for i in range(4))
    assert len(tags["images"]) == 16 and "fake_instantaneous_frequencies/3" in tags["images"]
    a = acc.Audio("fake_waveforms/0")[0]
    assert a.sample_rate == 16000 and a.length_frames == 5000
    im = acc.Images("real_magnitude_spectrograms/0")[0]
    assert (im.height, im.width) == (8, 1024)


@pytest.mark.parametrize("level", [0.3, 1.0])
def test_mask_pinned_gradient_criterion(emu, level):
    """The GPU step tests' criterion (common.check_substep) on the emulated kernels: the product's leaky-relu masks
    are taped in the oracle's call order, the fp64 oracle replays them and every gradient element agrees to 1e-3;
    a tape from a DIFFERENT input makes the flip count explode (the instrument really pins the branch)."""
    import gansynth_b200.functional as F
    import gansynth_b200.models as pmodels
    from common import check_substep, record_masks
    from oracle import ops as oops
    opg, params, ppg, latents, labels, images = _pair(level, emu)
    o64 = omodels.GANSynthStep(opg, {n: p.double() for n, p in params.items()}, HYPER)
    model = pmodels.GANSynth(ppg.generator, ppg.discriminator, None, None, {}, HYPER, device="cpu")
    model._ensure_optimizers(labels, latents)

    def coarse(mode, got, want):
        return grad_close(got, want, 5e-3), "max-norm"

    for scope in ("discriminator", "generator"):
        check_substep(model, emu, o64, scope, images, labels, latents, torch.float64, "emu", True, coarse)
    # a tape recorded on other latents does not describe this run: most masks of the generator disagree
    with record_masks(F.K) as masks:
        model.generator_loss_fn(labels, torch.randn(4, 256, generator=torch.Generator().manual_seed(99)))
    with oops.MaskTape(masks) as tape:
        o64.generator_update(labels.double(), latents.double(), apply=False)
    assert tape.flips > 0.05 * tape.total


# ----------------------------------------------------------------------------- pitch classifier (networks.py:293-413)
RESNET_SMALL = dict(conv_param=dict(filters=8, kernel_size=[7, 7], strides=[2, 2]), pool_param=dict(kernel_size=[3, 3], strides=[2, 2]),
                    residual_params=[dict(filters=8, strides=[1, 1], blocks=2), dict(filters=16, strides=[2, 2], blocks=2)],
                    groups=4, classes=11)


def test_resnet_classifier_forward_parity(emu):
    """networks.ResNet (forward only) on the emulated kernels against the oracle restatement: variable names / shapes of
    the reference's scopes, features and logits."""
    import gansynth_b200.networks as pnet
    onet_ = onet.ResNet(**RESNET_SMALL)
    params = onet_.init_variables(seed=5)
    images = torch.randn(3, 2, 32, 64, generator=torch.Generator().manual_seed(1))
    pres = pnet.ResNet(**RESNET_SMALL)
    pres(images)                                   # creates the variables (tf.get_variable on first use)
    assert sorted(emu.vars.keys()) == sorted(params.keys())
    for n, v in emu.vars.items():
        assert tuple(v.shape) == tuple(params[n].shape), n
    emu.load(params)
    gf, gl = pres(images)
    wf, wl = onet_(params, images)
    assert gf.shape == wf.shape == (3, 16) and gl.shape == wl.shape == (3, 11)
    assert rel_err(gf, wf) < TOL and rel_err(gl, wl) < TOL


def test_evaluate_with_a_resnet_classifier(emu):
    """GANSynth.evaluate with a networks.ResNet as the classifier (the reference splices a frozen GraphDef of this
    network, models.py:196-230): the Frechet distance equals metrics.frechet_inception_distance of the oracle's features."""
    import numpy as np
    import gansynth_b200.metrics as metrics
    import gansynth_b200.models as pmodels
    import gansynth_b200.networks as pnet
    from oracle import spectral_ops as osp
    opg, params, ppg, latents, labels, images = _pair(1.0, emu, batch=4)
    cls_o = onet.ResNet(**RESNET_SMALL)
    cparams = cls_o.init_variables(seed=7)
    cls_p = pnet.ResNet(**RESNET_SMALL)
    spectral = dict(waveform_length=152, sample_rate=16000, spectrogram_shape=[16, 16], overlap=0.75)
    waves = [0.1 * torch.randn(4, 152, generator=torch.Generator().manual_seed(20 + i)) for i in range(3)]
    lats = [torch.randn(4, 256, generator=torch.Generator().manual_seed(30 + i)) for i in range(3)]
    it_w, it_z = iter(waves), iter(lats)
    model = pmodels.GANSynth(ppg.generator, ppg.discriminator, lambda: (next(it_w), labels), lambda: next(it_z), spectral, HYPER,
                             device="cpu")
    model.real_images_from_waveforms = lambda w: torch.stack(osp.convert_to_spectrogram(w, **spectral), dim=1)   # no CUDA here
    cls_p(images)
    emu.load(cparams)
    got = model.evaluate("/nonexistent", None, cls_p)
    real = np.concatenate([cls_o(cparams, torch.stack(osp.convert_to_spectrogram(w, **spectral), dim=1))[0].numpy() for w in waves])
    fake = np.concatenate([cls_o(cparams, opg.generator(params, z, labels))[0].numpy() for z in lats])
    want = metrics.frechet_inception_distance(real, fake)
    assert abs(got["frechet_inception_distance"] - want) < 1e-3 * max(1.0, abs(want))


def test_pitch_classifier_checkpoint_loads_by_tf_names(emu, tmp_path):
    """models.load_pitch_classifier: a TF-1 Saver checkpoint of the reference's classifier (variables `resnet/...`, gamma /
    beta stored [1, C, 1, 1] as ops.py:131-140 creates them, Momentum slots beside them) -> networks.ResNet in the
    pitch_classifier_main.py configuration; features equal the oracle's on the same values."""
    import numpy as np
    import gansynth_b200.models as pmodels
    import gansynth_b200.networks as pnet
    from gansynth_b200 import tf_checkpoint as tfc
    cfg = pnet.ResNet.pitch_classifier()
    o = onet.ResNet(cfg.conv_param, cfg.pool_param, cfg.residual_params, cfg.groups, cfg.classes)
    params = o.init_variables(seed=9)
    tensors = {}
    for n, v in params.items():
        a = v.numpy()
        tensors[n] = a.reshape(1, -1, 1, 1) if n.endswith(("gamma", "beta")) else a
        tensors[n + "/Momentum"] = np.zeros_like(tensors[n])
    tensors["global_step"] = np.asarray(50000, dtype=np.int64)
    prefix = str(tmp_path / "model.ckpt-50000")
    tfc.save_bundle(prefix, tensors)
    net = pmodels.load_pitch_classifier(str(tmp_path), device="cpu")
    images = torch.randn(2, 2, 32, 64, generator=torch.Generator().manual_seed(2))
    gf, gl = net(images)
    wf, wl = o(params, images)
    assert gf.shape == (2, 512) and gl.shape == (2, 61)
    assert rel_err(gf, wf) < TOL and rel_err(gl, wl) < TOL
    with pytest.raises(NotImplementedError):
        pmodels.load_pitch_classifier("pitch_classifier.pb", device="cpu")


def test_pitch_classifier_training_parity(emu):
    """models.PitchClassifier (models.py:253-410) on the emulated kernels against the oracle restatement: three Momentum
    (Nesterov) steps with weight decay and a decaying learning rate -- loss, running accuracy, every variable."""
    import gansynth_b200.models as pmodels
    import gansynth_b200.networks as pnet
    from oracle import spectral_ops as osp
    o = onet.ResNet(**RESNET_SMALL)
    params = o.init_variables(seed=5)
    net = pnet.ResNet(**RESNET_SMALL)
    spectral = dict(waveform_length=600, sample_rate=16000, spectrogram_shape=[32, 64], overlap=0.75)
    g = torch.Generator().manual_seed(4)
    batches = [(0.1 * torch.randn(4, 600, generator=g), torch.nn.functional.one_hot(torch.randint(0, 11, (4,), generator=g), 11).float())
               for _ in range(3)]
    hp = dict(weight_decay=1e-3, momentum=0.9, use_nesterov=True,
              learning_rate=lambda step: pmodels.exponential_decay(0.05, step, 2, 0.5))
    it = iter(batches)
    clf = pmodels.PitchClassifier(net, lambda: next(it), spectral, hp, device="cpu")
    clf._images = lambda w: torch.stack(osp.convert_to_spectrogram(w, **spectral), dim=1)          # no CUDA here
    ostep = omodels.PitchClassifierStep(o, {n: p.double() for n, p in params.items()}, 1e-3, 0.9, True)
    correct = seen = 0
    for i, (w, lab) in enumerate(batches):
        images = clf._images(w)
        if i == 0:
            clf._ensure_optimizer(images)
            emu.load(params)
        lr = hp["learning_rate"](clf.global_step)
        want_total, want_ce, want_logits = ostep.update(images.double(), lab.double(), lr)
        got_wd = clf.weight_decay_loss()
        ce = clf.train_step(w, lab)
        assert abs(float(ce) - want_ce) < 1e-4 * max(1.0, abs(want_ce))
        assert abs(float(ce) + got_wd - want_total) < 1e-4 * max(1.0, abs(want_total))
        correct += int((want_logits.argmax(1) == lab.argmax(1)).sum())
        seen += 4
        assert abs(clf.accuracy - correct / seen) < 1e-9
        for n, v in emu.vars.items():
            assert rel_err(v, ostep.params[n]) < 5e-4, (i, n, rel_err(v, ostep.params[n]))
    assert int(clf.global_step.value) == 3
