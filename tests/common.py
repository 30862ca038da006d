"""Shared helpers for the parity tests: seeded inputs, hyper-parameters, oracle <-> product glue."""
import torch

HYPER = dict(generator_learning_rate=8e-4, generator_beta1=0.0, generator_beta2=0.99,
             discriminator_learning_rate=8e-4, discriminator_beta1=0.0, discriminator_beta2=0.99,
             mode_seeking_loss_weight=0.1, real_gradient_penalty_weight=5.0, fake_gradient_penalty_weight=0.0)

SPECTRAL = dict(waveform_length=64000, sample_rate=16000, spectrogram_shape=[128, 1024], overlap=0.75)

SMALL = dict(min_resolution=[4, 4], max_resolution=[16, 16], min_channels=32, max_channels=256)   # BASELINE config 1
FULL = dict(min_resolution=[2, 16], max_resolution=[128, 1024], min_channels=32, max_channels=256)


def seeded_inputs(batch, res, seed=0, latent_dim=256, num_labels=61):
    g = torch.Generator().manual_seed(seed)
    latents = torch.randn(batch, latent_dim, generator=g)
    labels = torch.nn.functional.one_hot(torch.randint(0, num_labels, (batch,), generator=g), num_labels).float()
    images = torch.randn(batch, 2, *res, generator=g) * 0.5
    return latents, labels, images


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def grad_close(got, want, tol):
    """max |got - want| <= tol * max |want| per variable; an all-zero reference (inactive colour block)
    requires an all-zero result."""
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    scale = float(want.abs().max())
    if scale == 0.0:
        return float(got.abs().max()) == 0.0
    return float((got - want).abs().max()) <= tol * scale


# ----------------------------------------------------------------------------- leaky-relu mask instruments
import contextlib  # noqa: E402


def _mask_of(y):
    """Sign pattern of a layer output as the oracle sees it (NCHW; [B, E] stays flat: row-major NCHW order)."""
    m = y.detach() > 0
    if m.dim() == 4:
        m = m.permute(0, 3, 1, 2)
    return m.contiguous().cpu()


@contextlib.contextmanager
def record_masks(backend):
    """Records the leaky-relu mask of every activated layer output the kernel backend produces, in call order
    (the order oracle.ops.leaky_relu is called in: both sides walk the reference's graph in the same order).
    The mask is read from the sign of the layer OUTPUT, which is what the product's backward kernels use."""
    masks = []
    act_pos = {"conv_c": 7, "conv_t": 7, "bias_act": 2, "conv_pn": -1}      # conv_pn always activates
    originals = {}

    def wrap(name, pos):
        orig = getattr(backend, name)

        def f(*a, **k):
            out = orig(*a, **k)
            act = 1 if pos < 0 else k.get("act", a[pos] if len(a) > pos else 0)
            if act == 1:
                masks.append(_mask_of(out[0] if isinstance(out, tuple) else out))
            return out
        return orig, f

    for name, pos in act_pos.items():
        if hasattr(backend, name):
            originals[name], f = wrap(name, pos)
            setattr(backend, name, f)
    orig_lrelu = backend.lrelu

    def lrelu(x):
        y = orig_lrelu(x)
        masks.append(_mask_of(y))
        return y
    backend.lrelu = lrelu
    try:
        yield masks
    finally:
        for name in list(originals) + ["lrelu"]:
            delattr(backend, name)          # the instance attribute shadowed the class method


def masked_grad_report(got, want, tol):
    """Per-variable max-norm comparison -> list of (name, error) that exceed tol."""
    bad = []
    for n, g in got.items():
        w = want[n].detach().double().cpu()
        g = torch.zeros_like(w) if g is None else g.detach().double().cpu()
        scale = float(w.abs().max())
        err = float((g - w).abs().max())
        if (scale == 0.0 and err != 0.0) or err > tol * scale:
            bad.append((n, err / (scale + 1e-300)))
    return bad


MAX_FLIP_FRACTION = 1e-3


def check_substep(model, store, ostep, scope, images, labels, latents, dtype, mode, plain=True, grad_ok=None,
                  retain_graph=False):
    """One sub-step's loss and flat gradient against the oracle, with the leaky-relu discontinuity MEASURED
    instead of tolerated:

      1. the product runs the sub-step while `record_masks` tapes the sign pattern of every activated layer output
         (what its backward kernels use as the leaky-relu mask);
      2. the oracle runs the same sub-step REPLAYING that tape (oracle.ops.MaskTape): same piecewise-linear branch,
         so every gradient element must agree to 1e-3 of the variable's largest magnitude -- no cosine, no slack;
      3. the tape counts the elements whose mask differs from the oracle's own sign (pre-activations within
         rounding of zero): at most MAX_FLIP_FRACTION of all activations;
      4. (plain) the oracle also runs un-pinned: losses agree to 1e-3, gradients in direction and size (grad_ok).

    Returns (loss tensor of the product, product gradients by name, replayed oracle gradients by name)."""
    import gansynth_b200.functional as F
    from oracle import ops as oops
    cast = lambda t: t.to(dtype)
    dev = store.device
    model._set_trainable(scope)
    names = list(store.trainable_variables(scope))
    with record_masks(F.K) as masks:
        if scope == "discriminator":
            loss = model.discriminator_loss_fn(images.to(dev), labels.to(dev), latents.to(dev))
        else:
            loss = model.generator_loss_fn(labels.to(dev), latents.to(dev))
        # retain_graph: the caller differentiates `loss` again (model._apply)
        grads = torch.autograd.grad(loss, [store.vars[n] for n in names], allow_unused=True, retain_graph=retain_graph)
    got = dict(zip(names, grads))
    with oops.MaskTape(masks) as tape:
        if scope == "discriminator":
            l_m, g_m = ostep.discriminator_update(cast(images), cast(labels), cast(latents), apply=False)
        else:
            l_m, g_m = ostep.generator_update(cast(labels), cast(latents), apply=False)
    frac = tape.flips / max(1, tape.total)
    print("%s sub-step (%s): %d of %d leaky-relu masks differ from the oracle's sign (%.2e)" %
          (scope, mode, tape.flips, tape.total, frac))
    assert frac <= MAX_FLIP_FRACTION, (tape.flips, tape.total)
    assert abs(float(loss.detach()) - float(l_m)) < 1e-3 * max(1.0, abs(float(l_m))), (float(loss), float(l_m))
    bad = masked_grad_report(got, g_m, 1e-3)
    assert not bad, "gradient elements beyond 1e-3 with identical masks: %s" % bad[:4]
    if plain:
        assert grad_ok is not None
        if scope == "discriminator":
            l_p, g_p = ostep.discriminator_update(cast(images), cast(labels), cast(latents), apply=False)
        else:
            l_p, g_p = ostep.generator_update(cast(labels), cast(latents), apply=False)
        assert abs(float(loss.detach()) - float(l_p)) < 1e-3 * max(1.0, abs(float(l_p))), (float(loss), float(l_p))
        for n in names:
            if got[n] is not None:
                ok, why = grad_ok(mode, got[n], g_p[n])
                assert ok, (n, why)
    return loss, got, g_m


