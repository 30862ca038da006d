"""CPU, world_size 2 over gloo: the data-parallel sub-step (rank-local batch, one all-reduce of the flat
gradient buffer, 1/world folded into Adam) equals a single process that averages the two shards'
gradients.  Host logic only: the kernels are the torch-CPU emulation (tests/emu_backend.py)."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _setup_paths():
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)


def _build(rank_seed, level=1.0):
    _setup_paths()
    import gansynth_b200.functional as F
    import gansynth_b200.models as M
    import gansynth_b200.networks as N
    import gansynth_b200.ops as ops
    from common import HYPER, SMALL, seeded_inputs
    from emu_backend import EmuBackend
    from oracle import networks as onet
    F.set_backend(EmuBackend())
    store = ops.set_default_store(ops.VariableStore(device="cpu", seed=0))
    M.reset_global_step()
    opg = onet.PGGAN(growing_level=level, **SMALL)
    params = opg.init_variables(seed=3, bias_std=0.1)
    ppg = N.PGGAN(growing_level=level, **SMALL)
    ppg._ensure_variables("generator", 256, 61)
    ppg._ensure_variables("discriminator", 0, 61)
    store.load(params)
    model = M.GANSynth(ppg.generator, ppg.discriminator, None, None, {}, HYPER, device="cpu")
    latents, labels, images = seeded_inputs(4, [16, 16], seed=rank_seed)
    return model, store, latents, labels, images


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    model, store, latents, labels, images = _build(100 + rank)
    model._ensure_optimizers(labels, latents)
    model._set_trainable("discriminator")
    model._apply("discriminator", model.discriminator_loss_fn(images, labels, latents))
    model._set_trainable("generator")
    model._apply("generator", model.generator_loss_fn(labels, latents))
    torch.save({n: v.detach().clone() for n, v in store.vars.items()}, os.path.join(out_dir, "rank%d.pt" % rank))
    torch.distributed.destroy_process_group()


def test_two_rank_step_equals_averaged_single_process(tmp_path):
    world, port = 2, 29500 + os.getpid() % 500
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    for n in r0:
        assert torch.equal(r0[n], r1[n]), "replicas diverged at %s" % n      # identical replicas after the update

    # single process: gradients of the two shards computed separately, averaged, then one Adam step
    model, store, _, _, _ = _build(100)
    shards = [_build(100 + r)[2:] for r in range(world)]     # (latents, labels, images) of each rank
    import gansynth_b200.functional as F
    from emu_backend import EmuBackend
    F.set_backend(EmuBackend())
    import gansynth_b200.ops as ops
    ops.set_default_store(store)
    lat0, lab0, _ = shards[0]
    model._ensure_optimizers(lab0, lat0)
    for scope in ("discriminator", "generator"):
        model._set_trainable(scope)
        names = list(store.trainable_variables(scope))
        total = None
        for latents, labels, images in shards:
            loss = (model.discriminator_loss_fn(images, labels, latents) if scope == "discriminator"
                    else model.generator_loss_fn(labels, latents))
            grads = torch.autograd.grad(loss, [store.vars[n] for n in names], allow_unused=True)
            grads = [torch.zeros_like(store.vars[n]) if g is None else g for n, g in zip(names, grads)]
            total = grads if total is None else [a + b for a, b in zip(total, grads)]
        st = model._opt[scope]
        views = store.unflatten(scope, st["grad"])
        for n, g in zip(names, total):
            views[n].copy_(g / world)
        st["t"] += 1
        hp = model.hyper_params
        F.K.adam_step(st["flat"], st["grad"], st["m"], st["v"], hp[scope + "_learning_rate"], hp[scope + "_beta1"],
                      hp[scope + "_beta2"], 1.0e-8, st["t"], 1.0)
    for n, v in store.vars.items():
        assert float((v.detach() - r0[n]).abs().max()) <= 1e-6 * max(1.0, float(r0[n].abs().max())), n
    ops.set_default_store(None)
    from gansynth_b200.kernels import CudaBackend
    F.set_backend(CudaBackend())


def _sync_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    model, store, latents, labels, images = _build(100)
    model._ensure_optimizers(labels, latents)
    import gansynth_b200.functional as F
    from gansynth_b200.kernels import CudaBackend
    slices = {}
    for scope, st in model._opt.items():
        n = st["flat"].numel()
        lo, hi = CudaBackend().adam_slice(n, rank, world)            # host arithmetic of the C ABI (gs_adam_slice)
        slices[scope] = (lo, hi)
        # what the fused update leaves behind: only this rank's slice of the Adam slots is live, the rest is stale
        ref = torch.arange(n, dtype=torch.float32)
        for key, k in (("m", 1.0), ("v", 2.0)):
            st[key].fill_(-7.0)
            st[key][lo:hi] = k * ref[lo:hi]
        st["fused"] = dict(rank=rank, world=world, synced=False)
    try:
        model._checkpoint_state()
        guarded = False
    except RuntimeError:
        guarded = True
    model._sync_optimizer_state()
    ok = guarded
    for scope, st in model._opt.items():
        ref = torch.arange(st["flat"].numel(), dtype=torch.float32)
        ok &= torch.equal(st["m"], ref) and torch.equal(st["v"], 2.0 * ref) and st["fused"]["synced"]
    model._checkpoint_state()                                        # allowed again
    torch.save(dict(ok=ok, slices=slices), os.path.join(out_dir, "sync%d.pt" % rank))
    torch.distributed.destroy_process_group()


def test_sharded_adam_slots_are_gathered_before_a_checkpoint(tmp_path):
    """The fused all-reduce + Adam kernel leaves every rank with only its slice of m / v (models._update);
    `_sync_optimizer_state` (collective) rebuilds the full slots on every rank, and `_checkpoint_state` refuses to run
    before it.  World size 2 over gloo; the slices are those of gs_adam_slice."""
    world, port = 2, 29500 + (os.getpid() + 7) % 500
    mp.spawn(_sync_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(tmp_path, "sync%d.pt" % r)) for r in range(world)]
    assert all(o["ok"] for o in outs)
    for scope in outs[0]["slices"]:
        (a0, b0), (a1, b1) = outs[0]["slices"][scope], outs[1]["slices"][scope]
        assert a0 == 0 and b0 == a1 and b0 % 4 == 0 and b1 > a1
