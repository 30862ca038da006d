"""GANSynth model / training step on the CUDA kernels (host mirror of reference models.py:8-250:
same constructor, `train`, `evaluate`, `generate` signatures).

The reference builds one TF graph and alternates session.run(discriminator_train_op) /
session.run(generator_train_op) (models.py:189-192), each run pulling a fresh real batch and fresh
latents.  Here each run is an eager sub-step:

  D sub-step: G forward (no graph) -> spectral forward on the real batch -> D(real), D(fake) ->
              softplus losses + R1 penalty (double backward through D) -> grads of D vars -> [all-reduce]
              -> fused TF-Adam on the flat D buffer.
  G sub-step: G forward -> D(fake) -> softplus loss + mode-seeking term (double backward through G) ->
              grads of G vars -> [all-reduce] -> fused TF-Adam on the flat G buffer; global_step += 1.

Data parallel (new design): one process per GPU, identical replicas, rank-local batch (and
rank-local batch_stddev groups), one NCCL all-reduce of the flat gradient buffer per sub-step.
"""
import gc
import glob
import logging
import math
import os
import time

import numpy as np
import torch

from . import functional as F
from . import ops
from . import spectral_ops

logger = logging.getLogger("gansynth_b200")


class GlobalStep(object):
    """Host-side tf.train global step: an int that can be divided into a lazily evaluated level."""

    def __init__(self):
        self.value = 0

    def __int__(self):
        return int(self.value)

    def __float__(self):
        return float(self.value)

    def __truediv__(self, other):
        return lambda: float(self.value) / float(other)


_global_step = None


def get_or_create_global_step():
    global _global_step
    if _global_step is None:
        _global_step = GlobalStep()
    return _global_step


def reset_global_step():
    global _global_step
    _global_step = None


def _to_device(x, device, dtype=torch.float32):
    if torch.is_tensor(x):
        return x.to(device=device, dtype=dtype, non_blocking=True)
    return torch.as_tensor(np.asarray(x)).to(device=device, dtype=dtype, non_blocking=True)


def _select_logits(logits, labels):
    """models.py:39-40: tf.gather_nd(logits, tf.where(labels)) -> the logit of the true class, [B]."""
    return logits.gather(1, torch.argmax(labels, dim=1, keepdim=True))[:, 0]


def load_pitch_classifier(path, device="cuda"):
    """`networks.ResNet.pitch_classifier()` (pitch_classifier_main.py:42-53) with the variables of the checkpoint at
    `path`: a TF-1 Saver prefix or model_dir (variables `resnet/...`, Momentum slots ignored) or a `.pt` file holding
    {name: array}.  A frozen GraphDef (`*.pb`, the reference's --classifier default) is refused: it needs TensorFlow."""
    from . import networks
    from . import tf_checkpoint as tfc
    if path.endswith(".pb"):
        raise NotImplementedError("%s is a frozen TensorFlow GraphDef; pass the classifier's CHECKPOINT instead (the "
                                  "model_dir of pitch_classifier_main.py) -- it is run by networks.ResNet" % path)
    if path.endswith(".pt"):
        values = torch.load(path, map_location="cpu")
        values = values.get("variables", values)
    else:
        prefix = tfc.latest_checkpoint(path) if os.path.isdir(path) else path
        if prefix is None:
            raise FileNotFoundError("no checkpoint under %s" % path)
        values = tfc.load_bundle(prefix)
    values = {n: v for n, v in values.items() if n.startswith("resnet/") and "/Momentum" not in n}
    if not values:
        raise KeyError("%s holds no `resnet/...` variables" % path)
    net = networks.ResNet.pitch_classifier(classes=int(np.asarray(values["resnet/logits/bias"]).shape[0]))
    with torch.no_grad():
        net(torch.zeros(1, 2, 32, 64, device=device))        # creates the variables (tf.get_variable on first use)
    store = ops.default_store()
    missing = [n for n in store.vars if n.startswith("resnet/") and n not in values]
    if missing:
        raise KeyError("classifier checkpoint %s lacks %d variables, e.g. %s" % (path, len(missing), missing[0]))
    # the reference keeps gamma / beta as [1, C, 1, 1] (ops.py:131-140): same values, flat here
    store.load({n: np.asarray(v).reshape(tuple(store.vars[n].shape)) for n, v in values.items() if n in store.vars})
    return net


class GANSynth(object):

    def __init__(self, generator, discriminator, real_input_fn, fake_input_fn, spectral_params, hyper_params,
                 device="cuda", process_group=None):
        self.generator = generator
        self.discriminator = discriminator
        self.real_input_fn = real_input_fn
        self.fake_input_fn = fake_input_fn
        self.spectral_params = dict(spectral_params)
        self.hyper_params = hyper_params
        self.device = torch.device(device)
        self.process_group = process_group
        self.global_step = get_or_create_global_step()
        self.store = ops.default_store()
        self._opt = None
        self._graphs = {}
        self._graph_pool = None
        self._stream = None
        self.use_cuda_graphs = os.environ.get("GS_CUDA_GRAPHS", "1") != "0"
        # audio / image summaries next to the scalar ones (needs tensorboard): opt in with GS_MEDIA_SUMMARIES=1 or
        # model.media_summaries = True (exercised on the GPU by tests/test_model_gpu.py, BASELINE config 1)
        self.media_summaries = os.environ.get("GS_MEDIA_SUMMARIES", "0") == "1"
        self.generator_loss = None
        self.discriminator_loss = None
        # last evaluated tensors, named like the reference attributes (models.py:91-108)
        self.real_waveforms = self.fake_waveforms = None
        self.real_images = self.fake_images = None
        self.real_labels = self.fake_labels = None
        self.real_features = self.fake_features = None
        self.real_logits = self.fake_logits = None

    # ------------------------------------------------------------------ inputs
    def _next_real(self):
        waveforms, labels = self.real_input_fn()
        return _to_device(waveforms, self.device), _to_device(labels, self.device)

    def _next_latents(self):
        return _to_device(self.fake_input_fn(), self.device)

    def real_images_from_waveforms(self, waveforms):
        """models.py:27-28."""
        mag, inst = spectral_ops.convert_to_spectrogram(waveforms, **self.spectral_params)
        return torch.stack([mag, inst], dim=1)

    # ------------------------------------------------------------------ losses
    def discriminator_loss_fn(self, real_images, labels, latents):
        """models.py:25-54, 65."""
        hp = self.hyper_params
        with torch.no_grad():
            fake_images = self.generator(latents, labels)
        if hp.get("fake_gradient_penalty_weight"):
            fake_images = fake_images.detach().requires_grad_(True)
        real_images = real_images.detach().requires_grad_(True)
        real_features, real_logits = self.discriminator(real_images, labels)
        fake_features, fake_logits = self.discriminator(fake_images, labels)
        real_logits = _select_logits(real_logits, labels)
        fake_logits = _select_logits(fake_logits, labels)
        losses = torch.nn.functional.softplus(-real_logits) + torch.nn.functional.softplus(fake_logits)
        if hp["real_gradient_penalty_weight"]:
            with F.skip_weight_grads():
                (grads,) = torch.autograd.grad(real_logits, real_images, grad_outputs=torch.ones_like(real_logits),
                                               create_graph=True)
            flat = grads.reshape(grads.shape[0], -1)
            losses = losses + F.RowDot.apply(flat, flat) * hp["real_gradient_penalty_weight"]
        if hp.get("fake_gradient_penalty_weight"):
            # models.py:50-54: the same penalty on the generator distribution (0.0 on the reference's command line,
            # gan_synth_main.py:87, so off the benchmarked path)
            with F.skip_weight_grads():
                (grads,) = torch.autograd.grad(fake_logits, fake_images, grad_outputs=torch.ones_like(fake_logits),
                                               create_graph=True)
            flat = grads.reshape(grads.shape[0], -1)
            losses = losses + F.RowDot.apply(flat, flat) * hp["fake_gradient_penalty_weight"]
        self.real_images, self.fake_images = real_images.detach(), fake_images.detach()
        self.real_features, self.fake_features = real_features.detach(), fake_features.detach()
        self.real_logits, self.fake_logits = real_logits.detach(), fake_logits.detach()
        return losses.mean()

    def generator_loss_fn(self, labels, latents):
        """models.py:25, 34, 57-64."""
        hp = self.hyper_params
        latents = latents.detach().requires_grad_(True)
        fake_images = self.generator(latents, labels)
        fake_features, fake_logits = self.discriminator(fake_images, labels)
        fake_logits = _select_logits(fake_logits, labels)
        losses = torch.nn.functional.softplus(-fake_logits)
        if hp["mode_seeking_loss_weight"]:
            with F.skip_weight_grads():
                (lg,) = torch.autograd.grad(fake_images, latents, grad_outputs=self._ones_like(fake_images),
                                            create_graph=True)
            losses = losses + hp["mode_seeking_loss_weight"] / (F.RowDot.apply(lg, lg) + 1.0e-6)
        self.fake_images = fake_images.detach()
        self.fake_features, self.fake_logits = fake_features.detach(), fake_logits.detach()
        return losses.mean()

    def _ones_like(self, t):
        key = (tuple(t.shape), t.device)
        if getattr(self, "_ones_key", None) != key:
            self._ones_key, self._ones = key, torch.ones_like(t)
        return self._ones

    # ------------------------------------------------------------------ optimiser state
    def _ensure_optimizers(self, labels, latents):
        if self._opt is not None:
            return
        # touch both networks once so that every variable exists, then pack them into flat buffers
        with torch.no_grad():
            pg = getattr(self.generator, "__self__", None)
            if pg is not None and hasattr(pg, "_ensure_variables"):
                pg._ensure_variables("generator", latents.shape[1], labels.shape[1])
                pg._ensure_variables("discriminator", 0, labels.shape[1])
            else:
                self.discriminator(self.generator(latents, labels), labels)
        self._opt = {}
        symm = self._symmetric_allocator()
        for scope in ("generator", "discriminator"):
            flat = self.store.pack(scope, alloc=symm)
            grad = symm(flat.numel()) if symm is not None and getattr(flat, "_gs_symmetric", False) else torch.zeros_like(flat)
            self._opt[scope] = dict(flat=flat, grad=grad, m=torch.zeros_like(flat), v=torch.zeros_like(flat), t=0)
            if getattr(flat, "_gs_symmetric", False) and getattr(grad, "_gs_symmetric", False):
                self._opt[scope]["fused"] = self._rendezvous(flat, grad)
        F.K.register_parameters([o["flat"] for o in self._opt.values()])

    # ------------------------------------------------------------------ NVLink symmetric memory (data parallel)
    def _world(self):
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            return torch.distributed.get_world_size(self.process_group), torch.distributed.get_rank(self.process_group)
        return 1, 0

    def _symmetric_allocator(self):
        """numel -> zeroed fp32 tensor in NVLink symmetric memory, or None: single process, CPU, GS_FUSED_ALLREDUCE=0, or a
        PyTorch build / fabric without symmetric memory (then the update is ncclAllReduce + gs_adam_step)."""
        world, _ = self._world()
        if world <= 1 or self.device.type != "cuda" or os.environ.get("GS_FUSED_ALLREDUCE", "1") == "0":
            return None
        if torch.distributed.get_backend(self.process_group) != "nccl":
            return None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            group = self.process_group if self.process_group is not None else torch.distributed.group.WORLD
            if hasattr(symm_mem, "enable_symm_mem_for_group"):
                symm_mem.enable_symm_mem_for_group(group.group_name)
        except Exception as exc:                                   # no symmetric memory in this build
            logger.warning("symmetric memory unavailable (%s): gradient all-reduce through NCCL", exc)
            return None
        device = self.device if self.device.index is not None else torch.device("cuda", torch.cuda.current_device())

        def alloc(numel):
            t = symm_mem.empty(int(numel), dtype=torch.float32, device=device)
            t.zero_()
            t._gs_symmetric = True
            return t
        try:
            alloc(32)
        except Exception as exc:
            logger.warning("symmetric memory allocation failed (%s): gradient all-reduce through NCCL", exc)
            return None
        return alloc

    def _rendezvous(self, flat, grad):
        """Maps the parameter and gradient buffers of all ranks into each other (collective) -> what
        gs_adam_step_allreduce needs: multicast addresses when the NVSwitch offers them (GS_FUSED_ALLREDUCE=p2p forces the
        peer-pointer form), the peers' pointers, the two handles for the barriers."""
        import torch.distributed._symmetric_memory as symm_mem
        group = self.process_group if self.process_group is not None else torch.distributed.group.WORLD
        hp, hg = symm_mem.rendezvous(flat, group), symm_mem.rendezvous(grad, group)
        world, rank = self._world()

        def peers(h, t):
            ptrs = [int(p) for p in h.buffer_ptrs]
            off = int(t.data_ptr()) - ptrs[rank]                       # the tensor's offset inside the mapped allocation
            return torch.tensor([p + off for p in ptrs], dtype=torch.int64, device=flat.device), off
        pp, poff = peers(hp, flat)
        gp, goff = peers(hg, grad)
        use_mc = os.environ.get("GS_FUSED_ALLREDUCE", "1") != "p2p" and int(getattr(hp, "multicast_ptr", 0) or 0) != 0 \
            and int(getattr(hg, "multicast_ptr", 0) or 0) != 0
        return dict(hp=hp, hg=hg, param_peers=pp, grad_peers=gp, world=world, rank=rank,
                    param_mc=int(hp.multicast_ptr) + poff if use_mc else 0, grad_mc=int(hg.multicast_ptr) + goff if use_mc else 0,
                    mode="multimem" if use_mc else "p2p", synced=True)

    def _sync_optimizer_state(self):
        """The fused update shards the Adam slots: every rank holds only its slice of m / v.  Collective: after this call
        every rank holds the full slots (before a checkpoint is written)."""
        for scope, st in (self._opt or {}).items():
            f = st.get("fused")
            if f is None or f["synced"]:
                continue
            lo, hi = F.K.adam_slice(st["flat"].numel(), f["rank"], f["world"])
            for key in ("m", "v"):
                full = torch.zeros_like(st[key])
                full[lo:hi] = st[key][lo:hi]
                torch.distributed.all_reduce(full, group=self.process_group)
                st[key].copy_(full)
            f["synced"] = True

    def _set_trainable(self, scope):
        for n, v in self.store.vars.items():
            v.requires_grad_(n.startswith(scope + "/"))

    def _backward(self, scope, loss):
        """Gradients of `loss` w.r.t. the variables of `scope`, written into the flat gradient buffer."""
        st = self._opt[scope]
        names = list(self.store.trainable_variables(scope).keys())
        views = self.store.unflatten(scope, st["grad"])
        if "sinks" not in st:
            st["sinks"] = {self.store.vars[n].data_ptr(): views[n] for n in names}
        # the convolution layers add their filter / bias gradients straight into the flat buffer (functional.grad_sinks);
        # what autograd still returns (dense, embedding, 1x1 / to-RGB layers) is added on top
        st["grad"].zero_()
        with F.grad_sinks(st["sinks"]):
            grads = torch.autograd.grad(loss, [self.store.vars[n] for n in names], allow_unused=True)
        rest = [(views[n], g) for n, g in zip(names, grads) if g is not None]
        if rest:
            torch._foreach_add_([v for v, _ in rest], [g for _, g in rest])      # one launch for the lot

    def _update(self, scope):
        """[all-reduce] + fused TF-Adam on the flat buffers (models.py:67-89)."""
        hp = self.hyper_params
        st = self._opt[scope]
        gflat = st["grad"]
        f = st.get("fused")
        if f is not None:
            # ONE kernel reduces the gradients over NVLink (in the switch when it multicasts), applies Adam to this rank's
            # slice and broadcasts the new parameters; barriers: all gradients written / all parameters visible
            st["t"] += 1
            f["hg"].barrier(channel=0)
            F.K.adam_step_allreduce(st["flat"], st["m"], st["v"], f["grad_mc"], f["param_mc"], f["grad_peers"], f["param_peers"],
                                    f["rank"], f["world"], hp[scope + "_learning_rate"], hp[scope + "_beta1"],
                                    hp[scope + "_beta2"], 1.0e-8, st["t"], 1.0 / f["world"])
            f["hp"].barrier(channel=1)
            f["synced"] = False
            F.K.weight_cache_refresh(st["flat"])
            return
        scale = 1.0
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size(self.process_group)
            if world > 1:
                torch.distributed.all_reduce(gflat, group=self.process_group)
                scale = 1.0 / world
        st["t"] += 1
        F.K.adam_step(st["flat"], gflat, st["m"], st["v"], hp[scope + "_learning_rate"], hp[scope + "_beta1"],
                      hp[scope + "_beta2"], 1.0e-8, st["t"], scale)
        F.K.weight_cache_refresh(st["flat"])   # re-split the cached bf16 copies of these parameters, in place

    def _apply(self, scope, loss):
        """minimize(loss, var_list=scope variables) (models.py:81-89) with TF-Adam semantics."""
        self._backward(scope, loss)
        self._update(scope)

    # ------------------------------------------------------------------ sub-steps (one session.run each)
    def _discriminator_body(self, real_waveforms, labels, latents):
        """Everything of the D sub-step up to the flat gradient: spectral front-end, G forward, D(real),
        D(fake), R1 double backward, gradients of the D variables."""
        self._set_trainable("discriminator")
        real_images = self.real_images_from_waveforms(real_waveforms)
        loss = self.discriminator_loss_fn(real_images, labels, latents)
        self._backward("discriminator", loss)
        return loss.detach()

    def _generator_body(self, labels, latents):
        self._set_trainable("generator")
        loss = self.generator_loss_fn(labels, latents)
        self._backward("generator", loss)
        return loss.detach()

    def _pggans(self):
        """The PGGAN object(s) the two network callables are bound to ([] for foreign callables)."""
        owners = []
        for fn in (self.generator, self.discriminator):
            pg = getattr(fn, "__self__", None)
            if pg is None or not hasattr(pg, "structure_key"):
                return []
            if not any(pg is o for o in owners):
                owners.append(pg)
        return owners

    def _structure_key(self):
        """Key of the kernel sequence of a sub-step, or None when it cannot be known (foreign networks): fully
        grown, or the integer depth at which the progressive-growing blend happens.  Within one key only the blend
        weight changes with global_step, and that lives in device memory (PGGAN.lerp_coef), so the sub-step can be
        replayed as a CUDA graph."""
        owners = self._pggans()
        return tuple(pg.structure_key() for pg in owners) if owners else None

    def _run_body(self, scope, body, inputs):
        """Runs `body(*inputs)` on the model's own stream: eagerly the first two times (lazy variable creation,
        workspace growth, function attributes), then captured into a CUDA graph (on the same stream, so that
        every autograd node -- the variables' AccumulateGrad nodes included -- lives on the capturing stream)
        and replayed on static input buffers.  The all-reduce and the Adam launch stay outside the graph (the
        Adam step count is a kernel argument)."""
        if not all(t.is_cuda for t in inputs):
            return body(*inputs)
        if self._stream is None:
            self._stream = torch.cuda.Stream(inputs[0].device)
        outer = torch.cuda.current_stream()
        if outer == self._stream or torch.cuda.is_current_stream_capturing():
            return body(*inputs)
        self._stream.wait_stream(outer)
        owners = self._pggans()
        with torch.cuda.stream(self._stream):
            for pg in owners:
                pg.update_lerp_coef(inputs[0].device)       # this step's blend weight, outside any graph
                pg.device_lerp_active = True
            try:
                loss = self._run_body_on_stream(scope, body, inputs)
            finally:
                for pg in owners:
                    pg.device_lerp_active = False
        outer.wait_stream(self._stream)
        return loss

    def _run_body_on_stream(self, scope, body, inputs):
        from . import _lib
        structure = self._structure_key()
        if not self.use_cuda_graphs or structure is None:
            return body(*inputs)
        key = (scope, structure) + tuple(tuple(t.shape) for t in inputs)
        if key not in self._graphs:
            # growing_depth only increases: a blend depth that has been left never comes back, so the graphs of other
            # structures (and the output buffers they keep alive) are dropped when a new one appears
            for old in [k for k in self._graphs if k[0] == scope and k[1] != structure]:
                del self._graphs[old]
        entry = self._graphs.setdefault(key, dict(calls=0))
        entry["calls"] += 1
        if entry["calls"] <= 2:
            return body(*inputs)
        if "graph" not in entry:
            static_in = [t.clone() for t in inputs]
            gc.collect()            # drop autograd graphs of earlier iterations that only cycles keep alive
            graph = torch.cuda.CUDAGraph()
            before = _lib.launch_count
            # the two sub-step graphs never overlap in time: they share one memory pool.
            # thread_local: the NCCL watchdog thread may query events while this thread captures
            pool = None if os.environ.get("GS_GRAPH_NO_SHARED_POOL") else self._graph_pool
            with torch.cuda.graph(graph, pool=pool, stream=self._stream, capture_error_mode="thread_local"):
                loss = body(*static_in)
            if pool is None:
                self._graph_pool = graph.pool()
            entry.update(graph=graph, static_in=static_in, loss=loss, launches=_lib.launch_count - before)
            _lib.launch_count = before
        for dst, src in zip(entry["static_in"], inputs):
            dst.copy_(src, non_blocking=True)
        entry["graph"].replay()
        _lib.launch_count += entry["launches"]     # the captured kernels did launch
        return entry["loss"]

    def discriminator_step(self, real_waveforms=None, labels=None, latents=None):
        if real_waveforms is None:
            real_waveforms, labels = self._next_real()
        if latents is None:
            latents = self._next_latents()
        self._ensure_optimizers(labels, latents)
        self.real_waveforms, self.real_labels, self.fake_labels = real_waveforms, labels, labels
        loss = self._run_body("discriminator", self._discriminator_body, (real_waveforms, labels, latents))
        self._update("discriminator")            # on the caller's stream, after the join with the model stream
        self.discriminator_loss = loss
        return self.discriminator_loss

    def generator_step(self, labels=None, latents=None):
        if labels is None:
            _, labels = self._next_real()
        if latents is None:
            latents = self._next_latents()
        self._ensure_optimizers(labels, latents)
        self.real_labels = self.fake_labels = labels
        loss = self._run_body("generator", self._generator_body, (labels, latents))
        self._update("generator")
        self.generator_loss = loss
        self.global_step.value += 1
        return self.generator_loss

    def train_step(self):
        """One iteration of the reference hot loop (models.py:189-192)."""
        d = self.discriminator_step()
        g = self.generator_step()
        return d, g

    # ------------------------------------------------------------------ checkpoints
    def _checkpoint_state(self):
        state = dict(global_step=int(self.global_step.value), variables=self.store.state())
        if self._opt is not None:
            if any(o.get("fused") is not None and not o["fused"]["synced"] for o in self._opt.values()):
                raise RuntimeError("the fused data-parallel update shards the Adam slots over the ranks: call "
                                   "_sync_optimizer_state() on EVERY rank before a checkpoint is written (train() does)")
            state["optimizers"] = {s: dict(m=o["m"].cpu(), v=o["v"].cpu(), t=o["t"]) for s, o in self._opt.items()}
        return state

    def save_checkpoint(self, model_dir, max_to_keep=10):
        os.makedirs(model_dir, exist_ok=True)
        path = os.path.join(model_dir, "model.ckpt-%d.pt" % int(self.global_step.value))
        # written under a temporary name and renamed (as tf.train.Saver does): a crash mid-write never leaves a
        # truncated file that sorts as the newest checkpoint
        tmp = path + ".tmp"
        torch.save(self._checkpoint_state(), tmp)
        os.replace(tmp, path)
        kept = sorted(glob.glob(os.path.join(model_dir, "model.ckpt-*.pt")),
                      key=lambda p: int(p.rsplit("-", 1)[1][:-3]))
        for old in kept[:-max_to_keep]:
            os.remove(old)
        return path

    def restore_latest(self, model_dir, labels=None, latents=None):
        """Implicit restore of SingularMonitoredSession(checkpoint_dir=model_dir) (models.py:120)."""
        paths = sorted(glob.glob(os.path.join(model_dir, "model.ckpt-*.pt")),
                       key=lambda p: int(p.rsplit("-", 1)[1][:-3]))
        # a model_dir written by the reference itself (TensorFlow Saver files + `checkpoint` state file): used when there
        # is no .pt file, or when the TF checkpoint is NEWER than the newest .pt (step parsed from `...ckpt-<step>`)
        if os.path.exists(os.path.join(model_dir, "checkpoint")):
            from . import tf_checkpoint as tfc
            prefix = tfc.latest_checkpoint(model_dir)
            tf_step = -1
            if prefix is not None and os.path.exists(prefix + ".index"):
                tail = prefix.rsplit("-", 1)[-1]
                tf_step = int(tail) if tail.isdigit() else 0
            pt_step = int(paths[-1].rsplit("-", 1)[1][:-3]) if paths else -1
            if tf_step > pt_step:
                return self.import_tf_checkpoint(model_dir, labels, latents)
        if not paths:
            return None
        state = None
        while paths:
            try:
                state = torch.load(paths[-1], map_location="cpu")
                break
            except Exception as exc:          # unreadable (e.g. written by an interrupted older version): try the one before
                logger.warning("checkpoint %s is unreadable (%s); falling back", paths[-1], exc)
                paths.pop()
        if state is None:
            return None
        if labels is not None:
            self._ensure_optimizers(labels, latents)
        self.store.load(state["variables"])
        self.global_step.value = state["global_step"]
        if self._opt is not None and "optimizers" in state:
            for s, o in state["optimizers"].items():
                self._opt[s]["m"].copy_(o["m"])
                self._opt[s]["v"].copy_(o["v"])
                self._opt[s]["t"] = o["t"]
                if self._opt[s].get("fused") is not None:
                    self._opt[s]["fused"]["synced"] = True       # every rank holds the full slots again
        return paths[-1]

    def import_tf_checkpoint(self, prefix_or_dir, labels=None, latents=None):
        """Loads a TensorFlow-1 Saver checkpoint of the reference graph (variables by their TF names, global step,
        Adam slots when present): `prefix_or_dir` is a checkpoint prefix (`.../model.ckpt-1000`) or a model_dir
        holding a `checkpoint` state file (tf.train.latest_checkpoint)."""
        from . import tf_checkpoint as tfc
        prefix = tfc.latest_checkpoint(prefix_or_dir) if os.path.isdir(prefix_or_dir) else prefix_or_dir
        if prefix is None:
            return None
        hp = self.hyper_params
        state = tfc.split_training_state(tfc.load_bundle(prefix),
                                         (hp["generator_beta2"], hp["discriminator_beta2"]))
        if labels is not None:
            self._ensure_optimizers(labels, latents)
        if not self.store.vars and "generator/weight" in state["variables"]:
            # fresh model: create the variables first; the label embedding [num_labels, latent_dim] (networks.py:154-160)
            # carries the two sizes the constructors need
            num_labels, latent_dim = (int(d) for d in state["variables"]["generator/weight"].shape)
            for pg in self._pggans():
                pg._ensure_variables("generator", latent_dim, num_labels)
                pg._ensure_variables("discriminator", 0, num_labels)
        if not self.store.vars:
            raise RuntimeError("import_tf_checkpoint: the model has no variables yet (call it after the first step, or "
                               "bind PGGAN methods so they can be created from the checkpoint)")
        missing = [n for n in self.store.vars if n not in state["variables"]]
        if missing:
            raise KeyError("checkpoint %s lacks %d variables, e.g. %s" % (prefix, len(missing), missing[0]))
        self.store.load({n: state["variables"][n] for n in self.store.vars})
        if state["global_step"] is not None:
            self.global_step.value = state["global_step"]
        if self._opt is not None:
            for scope, st in state["optimizers"].items():
                for slot in ("m", "v"):
                    flat = self._opt[scope][slot]
                    for n, (o, k) in self.store.offsets[scope].items():
                        if n in st[slot]:
                            flat[o:o + k].copy_(torch.as_tensor(st[slot][n]).reshape(-1))
                if st["t"] is not None:
                    self._opt[scope]["t"] = st["t"]
        return prefix

    def export_tf_checkpoint(self, prefix):
        """Writes the current state as a TensorFlow-1 Saver checkpoint (`<prefix>.index`, `.data-00000-of-00001`,
        `checkpoint`) under the names the reference graph uses."""
        from . import tf_checkpoint as tfc
        hp = self.hyper_params
        opts = {}
        if self._opt is not None:
            for scope, o in self._opt.items():
                m, v = o["m"].detach().cpu(), o["v"].detach().cpu()
                shapes = {n: tuple(self.store.vars[n].shape) for n in self.store.offsets[scope]}
                opts[scope] = dict(m={n: m[a:a + k].reshape(shapes[n]).numpy() for n, (a, k) in self.store.offsets[scope].items()},
                                   v={n: v[a:a + k].reshape(shapes[n]).numpy() for n, (a, k) in self.store.offsets[scope].items()},
                                   t=int(o["t"]))
        tensors = tfc.join_training_state({n: t.numpy() for n, t in self.store.state().items()},
                                          int(self.global_step.value), opts,
                                          (hp["generator_beta1"], hp["discriminator_beta1"]),
                                          (hp["generator_beta2"], hp["discriminator_beta2"]))
        tfc.save_bundle(prefix, tensors)
        return prefix

    # ------------------------------------------------------------------ reference entry points
    def train(self, model_dir, config=None, total_steps=1000000, save_checkpoint_steps=1000, save_summary_steps=100,
              log_tensor_steps=100):
        """models.py:110-194: alternate D and G updates until global_step reaches total_steps or the
        input function is exhausted (StopIteration / IndexError play the role of OutOfRangeError)."""
        rank0 = not (torch.distributed.is_available() and torch.distributed.is_initialized()) or \
            torch.distributed.get_rank() == 0
        restored = False
        iteration = 0
        t0 = time.time()
        while int(self.global_step.value) < total_steps:
            try:
                waveforms, labels = self._next_real()
                latents = self._next_latents()
                if not restored:
                    self.restore_latest(model_dir, labels, latents)
                    restored = True
                    if int(self.global_step.value) >= total_steps:
                        break
                self.discriminator_step(waveforms, labels, latents)
                self.generator_step()
            except (StopIteration, IndexError):
                break
            iteration += 1
            step = int(self.global_step.value)
            if rank0 and log_tensor_steps and iteration % log_tensor_steps == 0:
                print("INFO:gansynth_b200:global_step = %d, generator_loss = %.6f, discriminator_loss = %.6f (%.3f sec)"
                      % (step, float(self.generator_loss), float(self.discriminator_loss), time.time() - t0), flush=True)
                t0 = time.time()
            if rank0 and save_summary_steps and iteration % save_summary_steps == 0:
                self._write_summary(model_dir, step)
            if save_checkpoint_steps and step % save_checkpoint_steps == 0:
                self._sync_optimizer_state()         # collective (sharded Adam slots of the fused update); every rank
                if rank0:
                    self.save_checkpoint(model_dir)
        self._sync_optimizer_state()
        if rank0:
            self.save_checkpoint(model_dir)
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.barrier(group=self.process_group)     # nobody reads model_dir before rank 0 has written it

    def _write_summary(self, model_dir, step):
        """SummarySaverHook stand-in (models.py:131-170): the scalar summaries (generator_loss, discriminator_loss) plus
        the growing depth, one JSON object per line in `<model_dir>/summaries.jsonl` and as TensorBoard scalar events,
        followed by the audio / image summaries (`_write_media_summaries`)."""
        import json
        os.makedirs(model_dir, exist_ok=True)
        depth = [float(pg.growing_depth) for pg in self._pggans()]
        rec = dict(global_step=int(step), generator_loss=float(self.generator_loss),
                   discriminator_loss=float(self.discriminator_loss), time=time.time())
        if depth:
            rec["growing_depth"] = depth[0]
        with open(os.path.join(model_dir, "summaries.jsonl"), "a") as f:
            f.write(json.dumps(rec) + "\n")
        # the same scalars as TensorBoard event files in model_dir (tags as in models.py:163-172), when the
        # tensorboard package is there
        if getattr(self, "_tb", None) is None or getattr(self, "_tb_dir", None) != model_dir:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self._tb, self._tb_dir = SummaryWriter(log_dir=model_dir), model_dir
            except Exception:
                self._tb, self._tb_dir = False, model_dir
        if self._tb:
            try:
                for tag in ("generator_loss", "discriminator_loss"):
                    self._tb.add_scalar(tag, rec[tag], global_step=int(step))
                if self.media_summaries:
                    self._write_media_summaries(int(step))
                self._tb.flush()
            except Exception as e:                 # never let an optional event file stop a training run
                print("WARNING:gansynth_b200:summaries disabled after %s: %s" % (type(e).__name__, e), flush=True)
                self._tb = False

    @torch.no_grad()
    def _write_media_summaries(self, step, max_outputs=4):
        """The audio and image summaries of models.py:131-161: up to four real and generated clips (16 kHz audio) and
        their log-mel magnitude / instantaneous-frequency images.  Generated clips come from a fresh draw of
        `fake_input_fn` through `generate_batch` (in the reference every `session.run` draws new latents too)."""
        if self.real_waveforms is None or self.real_labels is None:
            return
        real = self.real_waveforms
        fake = self.generate_batch(self.real_labels, self._next_latents())
        n = min(max_outputs, int(real.shape[0]))
        rate = int(self.spectral_params.get("sample_rate", 16000))
        for name, wave in (("real_waveforms", real), ("fake_waveforms", fake)):
            for i in range(n):
                self._tb.add_audio("%s/%d" % (name, i), wave[i:i + 1].detach().float().cpu().clamp(-1.0, 1.0), step,
                                   sample_rate=rate)
        real_images = self.real_images_from_waveforms(real)
        for name, img in (("real_magnitude_spectrograms", real_images[:, 0]),
                          ("fake_magnitude_spectrograms", self.fake_images[:, 0]),
                          ("real_instantaneous_frequencies", real_images[:, 1]),
                          ("fake_instantaneous_frequencies", self.fake_images[:, 1])):
            for i in range(n):
                x = img[i].detach().float()
                lo, hi = x.min(), x.max()
                self._tb.add_image("%s/%d" % (name, i), ((x - lo) / (hi - lo + 1.0e-12)).cpu(), step, dataformats="HW")

    def evaluate(self, model_dir, config, classifier, input_name="images:0", output_names=("features:0", "logits:0")):
        """models.py:196-230: Frechet distance between classifier features of real and generated images over
        the whole input.  The reference splices a frozen TF GraphDef of its pitch classifier (`classifier`) onto
        `real_images` / `fake_images`; a GraphDef cannot be executed without TensorFlow, so `classifier` is one of
          * a callable images [B, 2, H, W] (CUDA) -> (features [B, F], logits [B, C]) -- e.g. a `networks.ResNet`, which
            is that classifier on this repo's kernels;
          * the path of a CHECKPOINT of the reference's classifier (a TF-1 Saver prefix / model_dir written by
            pitch_classifier_main.py, or a `.pt` dict of name -> array): `networks.ResNet.pitch_classifier()` is built
            and the variables are loaded by their TF names."""
        if isinstance(classifier, (str, os.PathLike)):
            classifier = load_pitch_classifier(os.fspath(classifier), self.device)
        if not callable(classifier):
            raise NotImplementedError("evaluate() takes the pitch classifier as a callable images -> (features, logits) or "
                                      "as the path of a checkpoint of networks.ResNet; a frozen TensorFlow GraphDef cannot "
                                      "be run without TensorFlow")
        from . import metrics
        real_feats, fake_feats = [], []
        restored = False
        with torch.no_grad():
            while True:
                try:
                    waveforms, labels = self._next_real()
                except (StopIteration, IndexError):
                    break
                latents = self._next_latents()
                if not restored:
                    pg = getattr(self.generator, "__self__", None)
                    if pg is not None and hasattr(pg, "_ensure_variables"):
                        pg._ensure_variables("generator", latents.shape[1], labels.shape[1])
                        pg._ensure_variables("discriminator", 0, labels.shape[1])
                    self.restore_latest(model_dir)
                    restored = True
                self.real_images = self.real_images_from_waveforms(waveforms)
                self.fake_images = self.generator(latents, labels)
                real_feats.append(classifier(self.real_images)[0].detach().float().cpu().numpy())
                fake_feats.append(classifier(self.fake_images)[0].detach().float().cpu().numpy())
        if not real_feats:
            raise ValueError("evaluate(): the input function produced no batch")
        return dict(frechet_inception_distance=metrics.frechet_inception_distance(np.concatenate(real_feats),
                                                                                   np.concatenate(fake_feats)))

    def generate(self, model_dir, config=None):
        """models.py:232-250: yields float32 numpy [B, waveform_length] until the input function ends."""
        restored = False
        while True:
            try:
                _, labels = self._next_real()
            except (StopIteration, IndexError):
                break
            latents = self._next_latents()
            if not restored:
                pg = getattr(self.generator, "__self__", None)
                if pg is not None and hasattr(pg, "_ensure_variables"):
                    pg._ensure_variables("generator", latents.shape[1], labels.shape[1])
                    pg._ensure_variables("discriminator", 0, labels.shape[1])
                self.restore_latest(model_dir)
                restored = True
            yield self.generate_batch(labels, latents).cpu().numpy()

    def _generate_body(self, labels, latents):
        fake_images = self.generator(latents, labels)
        mag, inst = fake_images[:, 0].contiguous(), fake_images[:, 1].contiguous()
        return spectral_ops.convert_to_waveform(mag, inst, **self.spectral_params), fake_images

    @torch.no_grad()
    def generate_batch(self, labels, latents):
        """z + pitch -> images -> waveforms (models.py:25, 30-31).  Once the generator is fully grown the whole
        chain (weight split, ~60 kernels, inverse spectral transform) is replayed as one CUDA graph per batch size
        from the third call on: at small batches the host needs longer to enqueue it than the device to run it."""
        waveforms, images = self._run_body("generate", self._generate_body, (labels, latents))
        # the output buffers belong to the graph: hand out copies
        self.fake_images = images.clone()
        self.fake_waveforms = waveforms.clone()
        return self.fake_waveforms


def exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False):
    """tf.train.exponential_decay: learning_rate * decay_rate ^ (global_step / decay_steps) (pitch_classifier_main.py:69-74)."""
    p = float(global_step) / float(decay_steps)
    return float(learning_rate) * float(decay_rate) ** (math.floor(p) if staircase else p)


class PitchClassifier(object):
    """The reference's pitch classifier trainer (models.py:253-410): `network(images) -> (features, logits)` (a
    networks.ResNet) on the log-mel / IF images of real clips, softmax cross-entropy + L2 weight decay on every variable
    whose name does not contain "normalization" (:267-271), tf.train.MomentumOptimizer (Nesterov) with a learning rate that
    may be a callable of the global step (:283-290).  Same constructor and method signatures; hyper_params keys
    `weight_decay, learning_rate, momentum, use_nesterov`."""

    SCOPE = "resnet"

    def __init__(self, network, input_fn, spectral_params, hyper_params, device="cuda"):
        self.network = network
        self.input_fn = input_fn
        self.spectral_params = dict(spectral_params)
        self.hyper_params = hyper_params
        self.device = torch.device(device)
        self.global_step = get_or_create_global_step()
        self.store = ops.default_store()
        self._opt = None
        self.loss = self.accuracy = None
        self._correct = self._seen = 0

    def _hp(self, key):
        hp = self.hyper_params
        return hp[key] if isinstance(hp, dict) else getattr(hp, key)

    def _images(self, waveforms):
        mag, inst = spectral_ops.convert_to_spectrogram(waveforms, **self.spectral_params)
        return torch.stack([mag, inst], dim=1)

    def _ensure_optimizer(self, images):
        if self._opt is not None:
            return
        with torch.no_grad():
            self.network(images[:1])                 # tf.get_variable: every variable exists after one call
        flat = self.store.pack(self.SCOPE)
        wd = torch.zeros_like(flat)
        decay = float(self._hp("weight_decay"))
        for n, (o, k) in self.store.offsets[self.SCOPE].items():
            if "normalization" not in n:
                wd[o:o + k] = decay
        self._opt = dict(flat=flat, grad=torch.zeros_like(flat), accum=torch.zeros_like(flat), wd=wd)

    def learning_rate(self):
        lr = self._hp("learning_rate")
        return float(lr(self.global_step)) if callable(lr) else float(lr)

    def loss_fn(self, images, labels):
        """Cross-entropy part of models.py:264-266 (the weight decay of :267-271 is applied inside the optimiser kernel and
        added to the reported loss by `weight_decay_loss`)."""
        features, logits = self.network(images)
        ce = -(torch.log_softmax(logits, dim=1) * labels).sum(dim=1).mean()
        return ce, logits

    def weight_decay_loss(self):
        st = self._opt
        return 0.5 * float((st["flat"] * st["flat"] * st["wd"]).sum())

    def train_step(self, waveforms=None, labels=None):
        if waveforms is None:
            waveforms, labels = self.input_fn()
        waveforms, labels = _to_device(waveforms, self.device), _to_device(labels, self.device)
        images = self._images(waveforms)
        self._ensure_optimizer(images)
        st = self._opt
        names = list(self.store.trainable_variables(self.SCOPE).keys())
        for n in names:
            self.store.vars[n].requires_grad_(True)
        ce, logits = self.loss_fn(images, labels)
        grads = torch.autograd.grad(ce, [self.store.vars[n] for n in names], allow_unused=True)
        views = self.store.unflatten(self.SCOPE, st["grad"])
        st["grad"].zero_()
        pairs = [(views[n], g) for n, g in zip(names, grads) if g is not None]
        torch._foreach_add_([v for v, _ in pairs], [g for _, g in pairs])
        scale = 1.0
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            torch.distributed.all_reduce(st["grad"])
            scale = 1.0 / torch.distributed.get_world_size()
        self.loss = ce.detach()
        F.K.momentum_step(st["flat"], st["grad"], st["accum"], st["wd"], self.learning_rate(), self._hp("momentum"),
                          self._hp("use_nesterov"), scale)
        F.K.weight_cache_refresh(st["flat"])
        self.global_step.value += 1
        pred, true = torch.argmax(logits.detach(), dim=1), torch.argmax(labels, dim=1)
        self._correct += int((pred == true).sum())
        self._seen += int(true.numel())
        self.accuracy = self._correct / max(1, self._seen)       # tf.metrics.accuracy: running
        return self.loss

    # ------------------------------------------------------------------ checkpoints
    def save_checkpoint(self, model_dir):
        os.makedirs(model_dir, exist_ok=True)
        path = os.path.join(model_dir, "model.ckpt-%d.pt" % int(self.global_step.value))
        state = dict(global_step=int(self.global_step.value),
                     variables={n: v for n, v in self.store.state().items() if n.startswith(self.SCOPE + "/")})
        if self._opt is not None:
            state["momentum"] = self._opt["accum"].cpu()
        torch.save(state, path + ".tmp")
        os.replace(path + ".tmp", path)
        return path

    def restore_latest(self, model_dir, images=None):
        paths = sorted(glob.glob(os.path.join(model_dir, "model.ckpt-*.pt")), key=lambda p: int(p.rsplit("-", 1)[1][:-3]))
        if not paths:
            return None
        state = torch.load(paths[-1], map_location="cpu")
        if images is not None:
            self._ensure_optimizer(images)
        self.store.load({n: v for n, v in state["variables"].items() if n in self.store.vars})
        self.global_step.value = state["global_step"]
        if self._opt is not None and "momentum" in state:
            self._opt["accum"].copy_(state["momentum"])
        return paths[-1]

    def train(self, model_dir, config=None, total_steps=50000, save_checkpoint_steps=1000, save_summary_steps=100,
              log_tensor_steps=100):
        """models.py:306-386: train until `total_steps` or the end of the input; checkpoints, loss / accuracy log lines."""
        restored = False
        iteration = 0
        t0 = time.time()
        while int(self.global_step.value) < total_steps:
            try:
                waveforms, labels = self.input_fn()
            except (StopIteration, IndexError):
                break
            if not restored:
                self.restore_latest(model_dir, self._images(_to_device(waveforms, self.device)))
                restored = True
                if int(self.global_step.value) >= total_steps:
                    break
            self.train_step(waveforms, labels)
            iteration += 1
            step = int(self.global_step.value)
            if log_tensor_steps and iteration % log_tensor_steps == 0:
                print("INFO:gansynth_b200:global_step = %d, loss = %.6f, accuracy = %.4f (%.3f sec)" %
                      (step, float(self.loss) + self.weight_decay_loss(), self.accuracy, time.time() - t0), flush=True)
                t0 = time.time()
            if save_checkpoint_steps and step % save_checkpoint_steps == 0:
                self.save_checkpoint(model_dir)
        self.save_checkpoint(model_dir)

    def evaluate(self, model_dir, config=None):
        """models.py:388-410: accuracy over the whole input."""
        correct = seen = 0
        restored = False
        with torch.no_grad():
            while True:
                try:
                    waveforms, labels = self.input_fn()
                except (StopIteration, IndexError):
                    break
                waveforms, labels = _to_device(waveforms, self.device), _to_device(labels, self.device)
                images = self._images(waveforms)
                if not restored:
                    self.network(images[:1])
                    self.restore_latest(model_dir)
                    restored = True
                _, logits = self.network(images)
                correct += int((torch.argmax(logits, dim=1) == torch.argmax(labels, dim=1)).sum())
                seen += int(labels.shape[0])
        return dict(accuracy=correct / max(1, seen))
