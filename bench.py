#!/usr/bin/env python
"""bench.py -- GAN train steps/s of the GANSynth hot path on B200 (BASELINE.json metric, configs[1]).

    python bench.py --gpus 1 --steps K --warmup W
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU oracle (the reference cannot run here: no TensorFlow)

A "step" is one iteration of the reference hot loop (models.py:189-192): one discriminator update and
one generator update at batch 8 per GPU on the full 2x16 -> 128x1024 PGGAN (fully grown), synthetic
64000-sample waveforms -> spectral front-end -> 2x128x1024 images, R1 + mode-seeking double backward,
TF-Adam.  Rank 0 prints ONE JSON line.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 8
F_G = 14898199040.0   # conv+dense FLOP per sample, generator forward
F_D = 14894152192.0
ITER_FLOP_PER_SAMPLE = 7 * F_G + 11 * F_D   # SURVEY 3.1 nominal model: D-run F_G + 9 F_D, G-run 6 F_G + 2 F_D

# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
# (key = ConvProfiler key: form, n, h, w, ci, co, ksize, stride)
NCU_TRAFFIC = {
    ("conv_c", 8, 128, 1024, 32, 32, 3, 1): (134295552 + 86461184, "profiles/ncu_ck_32_summary.txt"),
    ("conv_t", 8, 128, 1024, 32, 32, 3, 1): (134295552 + 86461184, "profiles/ncu_ck_32_summary.txt (same kernel and shape)"),
    ("conv_w", 8, 128, 1024, 32, 32, 3, 1): (268535552 + 4204032, "profiles/ncu_w_32_v2_summary.txt"),
}

HYPER = dict(generator_learning_rate=8e-4, generator_beta1=0.0, generator_beta2=0.99,
             discriminator_learning_rate=8e-4, discriminator_beta1=0.0, discriminator_beta2=0.99,
             mode_seeking_loss_weight=0.1, real_gradient_penalty_weight=5.0, fake_gradient_penalty_weight=0.0)
SPECTRAL = dict(waveform_length=64000, sample_rate=16000, spectrogram_shape=[128, 1024], overlap=0.75)
FULL = dict(min_resolution=[2, 16], max_resolution=[128, 1024], min_channels=32, max_channels=256)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=smax, reasons=sorted(reasons),
                    samples=len(sm))


# ------------------------------------------------------------------------------------------------ ours
def build_model(device, seed=3, growing_level=1.0):
    import gansynth_b200.models as M
    import gansynth_b200.networks as N
    import gansynth_b200.ops as ops
    store = ops.set_default_store(ops.VariableStore(device=device, seed=seed))
    M.reset_global_step()
    pggan = N.PGGAN(growing_level=growing_level, **FULL)
    model = M.GANSynth(pggan.generator, pggan.discriminator, None, None, SPECTRAL, HYPER, device=device)
    return model, store


def host_batches(n, rank, pinned):
    """Synthetic inputs of SURVEY 8d config 2: 0.1*N(0,1) waveforms, random one-hot pitch, N(0,1) latents."""
    g = torch.Generator().manual_seed(1000 * rank)
    out = []
    for _ in range(n):
        w = 0.1 * torch.randn(BATCH, 64000, generator=g)
        lab = torch.nn.functional.one_hot(torch.randint(0, 61, (BATCH,), generator=g), 61).float()
        lab2 = torch.nn.functional.one_hot(torch.randint(0, 61, (BATCH,), generator=g), 61).float()
        z1, z2 = torch.randn(BATCH, 256, generator=g), torch.randn(BATCH, 256, generator=g)
        items = [w, lab, z1, lab2, z2]
        out.append([t.pin_memory() if pinned else t for t in items])
    return out


def run_steps(model, batches, device, from_host):
    """One iteration per batch.  from_host: inputs are pinned host tensors copied inside the region and
    the losses are read back (e2e); otherwise they are already device tensors."""
    h2d = d2h = 0
    for w, lab, z1, lab2, z2 in batches:
        if from_host:
            h2d += sum(t.numel() * t.element_size() for t in (w, lab, z1, lab2, z2))
            w, lab, z1, lab2, z2 = (t.to(device, non_blocking=True) for t in (w, lab, z1, lab2, z2))
        d = model.discriminator_step(w, lab, z1)
        g = model.generator_step(lab2, z2)
        if from_host:
            vals = torch.stack([d, g]).cpu()
            d2h += vals.numel() * vals.element_size()
    return h2d, d2h


class ConvProfiler(object):
    """CUDA-event pairs around every convolution-family ABI call inside the timed region (no syncs)."""

    def __init__(self, backend):
        self.backend, self.records, self.orig = backend, [], {}

    def __enter__(self):
        for name in ("conv_c", "conv_t", "conv_w"):
            fn = getattr(self.backend, name)
            self.orig[name] = fn
            setattr(self.backend, name, self._wrap(name, fn))
        return self

    def _wrap(self, name, fn):
        def wrapped(a, b, *rest, **kw):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = fn(a, b, *rest, **kw)
            e.record()
            if name == "conv_w":
                ksize, stride = rest[0], rest[1]
                n, h, w, ci = a.shape
                co = b.shape[3]
                big_pixels = n * h * w
            else:
                ksize, stride = rest[1], rest[2]
                big = a if name == "conv_c" else out
                n, h, w = big.shape[:3]
                ci = a.shape[3] if name == "conv_c" else out.shape[3]
                co = out.shape[3] if name == "conv_c" else a.shape[3]
                big_pixels = n * h * w
            flops = 2.0 * (big_pixels // (stride * stride)) * ksize * ksize * ci * co
            io_bytes = 4.0 * (a.numel() + out.numel() + b.numel())
            self.records.append(((name, n, h, w, ci, co, ksize, stride), s, e, flops, io_bytes))
            return out
        return wrapped

    def __exit__(self, *exc):
        for name, fn in self.orig.items():
            setattr(self.backend, name, fn)

    def summary(self):
        agg = {}
        for key, s, e, flops, io_bytes in self.records:
            ms = s.elapsed_time(e)
            a = agg.setdefault(key, dict(ms=0.0, n=0, flops=flops, bytes=io_bytes))
            a["ms"] += ms
            a["n"] += 1
        return agg


def spectral_secondary(device, pk):
    """BASELINE config 3: batch 256 round trip, GSamp/s and HBM roofline of the two spectral kernels."""
    import gansynth_b200.spectral_ops as sp
    g = torch.Generator().manual_seed(0)
    w = (0.1 * torch.randn(256, 64000, generator=g)).to(device)
    lm, inst = sp.convert_to_spectrogram(w, **SPECTRAL)
    for _ in range(3):
        lm, inst = sp.convert_to_spectrogram(w, **SPECTRAL)
        back = sp.convert_to_waveform(lm, inst, **SPECTRAL)
    reps = 10
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(reps)]
    torch.cuda.synchronize()
    for r in range(reps):          # queued back to back: the host runs ahead, the events bracket device time only
        ev[r][0].record()
        lm, inst = sp.convert_to_spectrogram(w, **SPECTRAL)
        ev[r][1].record()
        back = sp.convert_to_waveform(lm, inst, **SPECTRAL)
        ev[r][2].record()
    torch.cuda.synchronize()
    tf = sum(e[0].elapsed_time(e[1]) for e in ev)
    ti = sum(e[1].elapsed_time(e[2]) for e in ev)
    tf, ti = tf / reps, ti / reps
    nbytes = 256 * 1304576.0
    samp = 256 * 64000.0
    return dict(workload="spectral round trip batch 256 (config 3); working set 668 MB > L2",
                fwd_ms=tf, inv_ms=ti, fwd_gsamp_s=samp / tf / 1e6, inv_gsamp_s=samp / ti / 1e6,
                roundtrip_gsamp_s=samp / (tf + ti) / 1e6,
                fwd_hbm_gbs=nbytes / tf / 1e6, inv_hbm_gbs=nbytes / ti / 1e6,
                fwd_frac=nbytes / tf / 1e6 / pk["hbm"], inv_frac=nbytes / ti / 1e6 / pk["hbm"],
                algorithmic_bytes_per_clip=1304576)


def inference_secondary(model, device):
    """BASELINE config 5 at B = 8: labels + latents -> generator -> inverse spectral -> waveforms (models.py:232-250)."""
    g = torch.Generator().manual_seed(2)
    lab = torch.nn.functional.one_hot(torch.arange(BATCH) % 61, 61).float().to(device)
    z = torch.randn(BATCH, 256, generator=g).to(device)
    for _ in range(3):
        out = model.generate_batch(lab, z)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    torch.cuda.synchronize()
    s.record()
    for _ in range(reps):
        out = model.generate_batch(lab, z)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    return dict(workload="generate_batch at B=8 (CUDA-graph replay): G forward + convert_to_waveform -> [8, 64000]",
                ms_per_batch=ms, clips_per_s=BATCH / ms * 1e3, output_shape=list(out.shape))


def oracle_iteration_time(batch, iters, threads=None):
    """Times the CPU oracle's full iteration (D update + G update) at `batch` on the host cores."""
    from oracle import models as omodels
    from oracle import networks as onet
    from oracle import spectral_ops as osp
    if threads:
        torch.set_num_threads(threads)
    pg = onet.PGGAN(growing_level=1.0, **FULL)
    params = pg.init_variables(seed=3)
    step = omodels.GANSynthStep(pg, params, HYPER)
    g = torch.Generator().manual_seed(0)
    times = []
    for _ in range(iters):
        w = 0.1 * torch.randn(batch, 64000, generator=g)
        lab = torch.nn.functional.one_hot(torch.randint(0, 61, (batch,), generator=g), 61).float()
        z1, z2 = torch.randn(batch, 256, generator=g), torch.randn(batch, 256, generator=g)
        t0 = time.perf_counter()
        mag, inst = osp.convert_to_spectrogram(w, **SPECTRAL)
        real = torch.stack([mag, inst], 1)
        step.discriminator_update(real, lab, z1)
        step.generator_update(lab, z2)
        times.append(time.perf_counter() - t0)
    return times


def oracle_secondary_times(threads=None):
    """CPU restatement timed on the host cores for the two secondary workloads: the spectral
    round trip on a bounded sample of config 3 (two chunks of 8 clips) and pitch-conditional inference at B = 1."""
    from oracle import networks as onet
    from oracle import spectral_ops as osp
    if threads:
        torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    w = 0.1 * torch.randn(8, 64000, generator=g)
    osp.convert_to_waveform(*osp.convert_to_spectrogram(w, **SPECTRAL), **SPECTRAL)       # warm-up (constants)
    tf = ti = 0.0
    for _ in range(2):
        w = 0.1 * torch.randn(8, 64000, generator=g)
        t0 = time.perf_counter()
        mag, inst = osp.convert_to_spectrogram(w, **SPECTRAL)
        t1 = time.perf_counter()
        osp.convert_to_waveform(mag, inst, **SPECTRAL)
        t2 = time.perf_counter()
        tf, ti = tf + (t1 - t0), ti + (t2 - t1)
    samp = 16 * 64000.0
    pg = onet.PGGAN(growing_level=1.0, **FULL)
    params = pg.init_variables(seed=3)
    lab = torch.nn.functional.one_hot(torch.arange(1) % 61, 61).float()
    z = torch.randn(1, 256, generator=g)
    with torch.no_grad():
        t0 = time.perf_counter()
        img = pg.generator(params, z, lab)
        osp.convert_to_waveform(img[:, 0], img[:, 1], **SPECTRAL)
        t_inf = time.perf_counter() - t0
    return dict(kind="port", cores=threads or torch.get_num_threads(),
                sample="spectral: 2 chunks of 8 clips of the batch-256 workload; inference: one clip (B = 1)",
                fwd_gsamp_s=samp / tf / 1e9, inv_gsamp_s=samp / ti / 1e9, roundtrip_gsamp_s=samp / (tf + ti) / 1e9,
                inference_clips_per_s=1.0 / t_inf)


def bench_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU oracle")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)
    pk = peaks()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=device, dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            return float(t[0])
        return ms

    import gansynth_b200._lib as lib
    import gansynth_b200.functional as Fn
    model, store = build_model(device, growing_level=args.growing_level)
    host = host_batches(args.steps + args.warmup, rank, pinned=True)
    dev = [[t.to(device) for t in b] for b in host]
    torch.cuda.synchronize()

    # ---- device-resident run: `value` (sub-steps replay as CUDA graphs after the first warm-up steps)
    run_steps(model, dev[:args.warmup], device, False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.launch_count
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    run_steps(model, dev[args.warmup:], device, False)
    end.record()
    barrier()
    ms = max_over_ranks(start.elapsed_time(end))
    launches = lib.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end run through the public step API from pinned host buffers: `e2e`
    barrier()
    start.record()
    h2d, d2h = run_steps(model, host[args.warmup:], device, True)
    end.record()
    barrier()
    ms_e2e = max_over_ranks(start.elapsed_time(end))

    # ---- per-kernel timing: the graph replay cannot be bracketed kernel by kernel, so the same steps run once
    # more EAGERLY with a CUDA-event pair around every convolution-family ABI call (same inputs, same process)
    graphs_were = model.use_cuda_graphs
    model.use_cuda_graphs = False
    run_steps(model, dev[:1], device, False)
    barrier()
    with ConvProfiler(Fn.K) as prof:
        start.record()
        run_steps(model, dev[args.warmup:], device, False)
        end.record()
        barrier()
    ms_eager = start.elapsed_time(end)
    conv = prof.summary()
    model.use_cuda_graphs = graphs_were

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    iterations_s = args.steps / (ms / 1e3)
    value = iterations_s * world      # batch-8 steps of ALL ranks per second (every rank runs one per iteration)
    if args.conv_table:
        with open(args.conv_table, "w") as f:
            f.write("%-8s %3s %5s %5s %4s %4s %2s %2s %5s %10s %9s %9s %8s\n" % (
                "form", "n", "h", "w", "ci", "co", "k", "s", "calls", "total_ms", "avg_us", "TFLOP/s", "GB/s"))
            for key, v in sorted(conv.items(), key=lambda kv: -kv[1]["ms"]):
                avg = v["ms"] / v["n"]
                f.write("%-8s %3d %5d %5d %4d %4d %2d %2d %5d %10.3f %9.1f %9.1f %8.0f\n" % (
                    key + (v["n"], v["ms"], avg * 1e3, v["flops"] / (avg * 1e-3) / 1e12, v["bytes"] / (avg * 1e-3) / 1e9)))
    # one KERNEL per entry: the stride-1 3x3 input-gradient (conv_t) is the forward kernel on flipped weights, so the
    # two forms of a shape are the same launches as far as an ncu / CUPTI list can tell
    kernels = {}
    for key, v in conv.items():
        form = key[0]
        if form in ("conv_c", "conv_t") and key[6] == 3 and key[7] == 1 and key[4] == key[5]:
            key = ("conv_c",) + key[1:]
        a = kernels.setdefault(key, dict(ms=0.0, n=0, flops=v["flops"], bytes=v["bytes"]))
        a["ms"] += v["ms"]
        a["n"] += v["n"]
    top_key, top = max(kernels.items(), key=lambda kv: kv[1]["ms"])
    conv_ms = sum(v["ms"] for v in conv.values())
    avg_ms = top["ms"] / top["n"]
    achieved_tf = top["flops"] / (avg_ms * 1e-3) / 1e12
    line = dict(
        metric="GAN train steps/sec (batch 8/GPU, 128x1024 mel+IF)", value=value, unit="steps/s", n_gpus=world,
        steps=args.steps, warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak",
        vs_baseline=None, dtype="f32", data="synthetic",
        config=dict(workload="BASELINE configs[1]: full 2x16->128x1024 PGGAN G+D step (D update + G update, R1 + "
                             "mode-seeking double backward, TF-Adam), batch 8 per GPU, %s" %
                             ("fully grown" if args.growing_level >= 1.0 else "growing_level %.4f" % args.growing_level),
                    global_batch=BATCH * world, parallelism="dp%d" % world,
                    l2="per-step activation working set is several GB >> 126 MB L2; no flush needed",
                    cuda_graphs=bool(graphs_were), eager_ms_per_step=ms_eager / args.steps,
                    iterations_per_s=iterations_s,
                    value_counts="batch-8 steps over all ranks (N per lock-step data-parallel iteration)",
                    samples_per_s=value * BATCH),
        clocks=clocks,
        e2e=dict(value=args.steps * world / (ms_e2e / 1e3), unit="steps/s", h2d_bytes_per_step=h2d // args.steps,
                 d2h_bytes_per_step=d2h // args.steps),
        gpu_launches=launches,
        roofline=dict(bound="tensor", achieved=achieved_tf, peak=pk["tf_sustained"], unit="TFLOP/s",
                      frac=achieved_tf / pk["tf_sustained"], traffic=NCU_TRAFFIC.get(top_key, (None, None))[0],
                      traffic_source=NCU_TRAFFIC.get(top_key, (None, "no ncu --set full capture of this kernel/shape"))[1],
                      kernel="%s n=%d %dx%d ci=%d co=%d k=%d s=%d (tcgen05 bf16x3 implicit GEMM, TMA-fed; forward and "
                             "stride-1 input-gradient launches of this shape)" % top_key,
                      timing="CUDA-event pair around every launch of this kernel in an eager pass of the same steps "
                             "(the timed region replays CUDA graphs)",
                      launches_timed=top["n"], avg_launch_ms=avg_ms, algorithmic_flop_per_launch=top["flops"],
                      io_bytes_per_launch=top["bytes"], hbm_gbs_at_io_bytes=top["bytes"] / (avg_ms * 1e-3) / 1e9,
                      hbm_frac_at_io_bytes=top["bytes"] / (avg_ms * 1e-3) / 1e9 / pk["hbm"],
                      # launches per step x event-timed launch duration over the graph-replayed step time: comparable
                      # with the kernel's share in profiles/launches_*.csv / step_kernels_*.txt
                      share_of_step=top["ms"] / ms, share_of_eager_step=top["ms"] / ms_eager,
                      # the family total is dominated by launch-bound small layers whose event pairs include eager
                      # launch gaps: only meaningful against the eager pass it was measured in
                      conv_family_share_of_eager_step=conv_ms / ms_eager,
                      peak_source="%s bf16_tflops_sustained (kernel timed inside a long step)" % pk["source"],
                      step_tflops=ITER_FLOP_PER_SAMPLE * BATCH * value / 1e12,
                      step_frac=ITER_FLOP_PER_SAMPLE * BATCH * value / 1e12 / (pk["tf_sustained"] * world)),
    )
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count()
        times = oracle_iteration_time(4, 1, cores)
        line["cpu_baseline"] = dict(value=0.5 / times[0], unit="steps/s", cores=cores, kind="port",
                                    sample="one full iteration (D update + G update) of the oracle at batch 4 "
                                           "(%.1f s), scaled x0.5 to batch 8" % times[0])
    if world == 1 and not args.no_spectral:
        line["secondary"] = spectral_secondary(device, pk)
        line["secondary"]["inference"] = inference_secondary(model, device)
        if rank == 0 and not args.no_cpu_baseline:
            line["secondary"]["cpu_baseline"] = oracle_secondary_times(os.cpu_count())
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def bench_reference(args):
    """The reference's own implementation cannot run here (TensorFlow 1.13 absent, NCHW CPU kernels do
    not exist): this arm times the CPU oracle port of the same iteration on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = os.cpu_count()
    budget = float(os.environ.get("GS_REF_BUDGET_S", "240"))
    probe = oracle_iteration_time(4, 1, cores)[0]          # also serves as warm-up
    total = args.steps + max(0, args.warmup - 1)
    steps = args.steps if probe * total <= budget else max(1, int(budget / probe) - max(0, args.warmup - 1))
    times = oracle_iteration_time(4, max(0, args.warmup - 1) + steps, cores)[max(0, args.warmup - 1):]
    per = float(np.mean(times))
    value = 0.5 / per
    sample = ("each step = one full iteration (D update + G update) of the CPU oracle at batch 4, scaled x0.5 to "
              "batch 8; %d of the requested %d steps run to stay within %.0f s" % (steps, args.steps, budget))
    print(json.dumps(dict(
        impl="reference", metric="GAN train steps/sec (batch 8/GPU, 128x1024 mel+IF)", value=value, unit="steps/s",
        n_gpus=world, steps=steps, warmup=args.warmup, ms_per_step=1e3 / value, higher_is_better=True, scaling="weak",
        vs_baseline=None, dtype="f32", data="synthetic",
        config=dict(workload="BASELINE configs[1] on host CPU cores via the oracle port (PyTorch-CPU fp32)",
                    global_batch=BATCH, parallelism="cpu"),
        cpu_baseline=dict(value=value, unit="steps/s", cores=cores, kind="port", sample=sample),
        e2e=dict(value=value, unit="steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-spectral", action="store_true")
    ap.add_argument("--growing-level", type=float, default=1.0,
                    help="PGGAN growing_level held fixed for the run (default 1.0 = fully grown = BASELINE configs[1]; "
                         "< 63/127 exercises the progressive-growing blend path)")
    ap.add_argument("--conv-table", default=None, help="write per-shape convolution timings of the timed region here")
    args = ap.parse_args()
    # stdout carries ONE JSON line: everything libraries print there meanwhile (NCCL's version banner is a plain
    # printf to stdout at NCCL_DEBUG >= VERSION) is routed to stderr at the file-descriptor level
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line_out = io.StringIO()
        with contextlib.redirect_stdout(line_out):
            _dispatch(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    sys.stdout.write(line_out.getvalue())
    sys.stdout.flush()


def _dispatch(args):
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_ours(args)


if __name__ == "__main__":
    main()
