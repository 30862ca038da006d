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
    # dram__bytes_read.sum + dram__bytes_write.sum of one launch, `ncu --set full` captures of the round-2 binary
    ("conv_c", 8, 128, 1024, 32, 32, 3, 1): (134296064 + 85676800, "profiles/ncu_r2_ck_32_summary.txt"),
    ("conv_t", 8, 128, 1024, 32, 32, 3, 1): (134296064 + 85676800, "profiles/ncu_r2_ck_32_summary.txt (same kernel and shape)"),
    ("conv_w", 8, 128, 1024, 32, 32, 3, 1): (268608000 + 4684800, "profiles/ncu_r2_w_32_summary.txt"),
    ("conv_t", 8, 128, 1024, 32, 64, 3, 2): (67240448 + 74136064, "profiles/ncu_r2_t2_64_32_summary.txt"),
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


def conv_kernel_label(key):
    """Which kernel function serves a convolution-family call (the gates of csrc/conv.cu restated), for joining the
    event-timed algorithmic work with the CUPTI kernel names: (label, CUPTI name prefix)."""
    form, n, h, w, ci, co, ksize, stride = key
    if form == "conv_w":
        if ksize == 3 and ci % 32 == 0 and co % 32 == 0:
            return "conv_tcw_kernel (filter gradients)", "conv_tcw_kernel"
        return "filter gradient, 1x1 / 257-channel (SIMT)", None
    out_ch = co if form == "conv_c" else ci
    in_ch = ci if form == "conv_c" else co
    if ksize == 3 and in_ch % 32 == 0 and out_ch % 32 == 0 and out_ch <= 256:
        if stride == 1:
            if out_ch in (32, 64) and w >= 128 and h % 8 == 0:
                return "conv_tck_kernel (3x3 s1, thin layers, kw-stacked)", "conv_tck_kernel"
            return "conv_tc_kernel<C1> (3x3 s1)", "conv_tc_kernel<0"
        if form == "conv_c":
            return "conv_tc_kernel<C2> (3x3 s2 gather)", "conv_tc_kernel<1"
        return "conv_tc_kernel<T2> (3x3 s2 transposed)", "conv_tc_kernel<2"
    return "1x1 / 257-channel convolutions (SIMT)", None


# backend method -> CUPTI kernel-name prefix, for the HBM-bound elementwise passes whose algorithmic bytes are the
# tensors they read and write
ELEMENTWISE_KERNELS = {
    "mask_mul": "mask_mul4_kernel", "mask_mul_colsum": "mask_mul_colsum_kernel", "col_sum": "col_sum_kernel",
    "pn_fwd": "pixel_norm_vec_kernel<0", "pn_bwd": "pixel_norm_vec_kernel<1", "pn_bwd_mask_y": "pixel_norm_vec_kernel<3",
    "pn_bwd_mask": "pixel_norm_vec_kernel<3", "dense_fwd": "dense_fwd_kernel", "dense_dgrad": "dense_dgrad_kernel",
    "dense_wgrad": "dense_wgrad_kernel", "transpose_inner": "transpose_inner_kernel",
}


SPIN_CYCLES = 120000        # ~60 us at 1.9 GHz


class KernelProfiler(object):
    """CUDA-event pairs around every convolution-family ABI call of an eager pass (no syncs), and the algorithmic
    bytes (tensors read + written) of the elementwise backend calls."""

    def __init__(self, backend):
        self.backend, self.records, self.ew = backend, [], {}
        self.patched = []

    def __enter__(self):
        self._patch("_conv", self._wrap_conv(self.backend._conv))
        self._patch("conv_w", self._wrap_w(self.backend.conv_w))
        for name in ELEMENTWISE_KERNELS:
            if hasattr(self.backend, name):
                self._patch(name, self._wrap_ew(name, getattr(self.backend, name)))
        if hasattr(self.backend, "pn_bwd_mask_second_y"):
            self._patch("pn_bwd_mask_second_y", self._wrap_ew("pn_bwd_mask_second_y", self.backend.pn_bwd_mask_second_y))
        return self

    def _patch(self, name, fn):
        setattr(self.backend, name, fn)          # instance attribute shadows the class method
        self.patched.append(name)

    def _timed(self, key, flops, call):
        # The eager pass is host-bound: between recording the start event and the kernel launch the host spends ~25 us
        # (output allocation, ctypes, tensor-map encoding) during which an idle device would already have passed the
        # event.  A short spin kernel in front keeps the device busy until event and kernel are both queued, so the pair
        # holds device time only.
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(SPIN_CYCLES)
        s.record()
        out = call()
        e.record()
        outs = out if isinstance(out, tuple) else (out,)
        return out, s, e, outs

    def _wrap_conv(self, fn):
        def wrapped(name, x, w, bias, ksize, stride, wswap, alpha, act, precise, epi, aux, want_r, eps):
            out, s, e, outs = self._timed(None, None, lambda: fn(name, x, w, bias, ksize, stride, wswap, alpha, act, precise,
                                                                 epi, aux, want_r, eps))
            y = outs[0]
            form = "conv_c" if name == "gs_conv2d_fwd_ex" else "conv_t"
            big = x if form == "conv_c" else y
            n, h, wd = big.shape[:3]
            ci = x.shape[3] if form == "conv_c" else y.shape[3]
            co = y.shape[3] if form == "conv_c" else x.shape[3]
            flops = 2.0 * (n * h * wd // (stride * stride)) * ksize * ksize * ci * co
            io = 4.0 * (x.numel() + y.numel() + w.numel() + (aux.numel() if aux is not None else 0))
            self.records.append(((form, n, h, wd, ci, co, ksize, stride), s, e, flops, io))
            return out
        return wrapped

    def _wrap_w(self, fn):
        def wrapped(x, dy, ksize, stride, wswap, alpha, bias_of=None, out=None):
            res, s, e, outs = self._timed(None, None, lambda: fn(x, dy, ksize, stride, wswap, alpha, bias_of=bias_of, out=out))
            n, h, wd, ci = x.shape
            co = dy.shape[3]
            flops = 2.0 * (n * h * wd // (stride * stride)) * ksize * ksize * ci * co
            written = [o for o in (out if out is not None else outs) if torch.is_tensor(o)]
            self.records.append((("conv_w", n, h, wd, ci, co, ksize, stride), s, e, flops,
                                 4.0 * (x.numel() + dy.numel() + sum(o.numel() for o in written))))
            return res
        return wrapped

    def _wrap_ew(self, name, fn):
        def wrapped(*a, **k):
            out = fn(*a, **k)
            outs = out if isinstance(out, tuple) else (out,)
            nbytes = 4.0 * sum(t.numel() for t in list(a) + list(outs) if torch.is_tensor(t))
            rec = self.ew.setdefault(name, dict(bytes=0.0, n=0))
            rec["bytes"] += nbytes
            rec["n"] += 1
            return out
        return wrapped

    def __exit__(self, *exc):
        for name in self.patched:
            delattr(self.backend, name)

    def summary(self):
        agg = {}
        for key, s, e, flops, io_bytes in self.records:
            ms = s.elapsed_time(e)
            a = agg.setdefault(key, dict(ms=0.0, n=0, flops=flops, bytes=io_bytes))
            a["ms"] += ms
            a["n"] += 1
        return agg


def cupti_kernel_times(model, batches, device):
    """Kernel time per function name over graph-replayed iterations (CUPTI activity records through torch.profiler;
    outside the timed region) -> ({name: [launches, microseconds]}, total microseconds)."""
    import collections
    import re
    from torch.profiler import ProfilerActivity, profile
    agg = collections.defaultdict(lambda: [0, 0.0])
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run_steps(model, batches, device, False)
        torch.cuda.synchronize()
    tot = 0.0
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        name = re.sub(r"\(.*", "", ev.name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", ""))
        us = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        agg[name][0] += 1
        agg[name][1] += us
        tot += us
    return agg, tot


def kernel_table(conv, ew, cupti, cupti_total_us, steps_profiled, steps_eager, pk):
    """One row per kernel FUNCTION: CUPTI time per replayed step joined with the algorithmic FLOP / bytes of the calls
    it serves (from the eager pass), each against the roofline that bounds it (arithmetic intensity vs the ridge)."""
    ridge = pk["tf_sustained"] * 1e12 / (pk["hbm"] * 1e9)
    groups = {}
    for key, v in conv.items():
        label, prefix = conv_kernel_label(key)
        g = groups.setdefault(label, dict(prefix=prefix, flops=0.0, bytes=0.0, calls=0, event_ms=0.0))
        g["flops"] += v["flops"] * v["n"] / steps_eager
        g["bytes"] += v["bytes"] * v["n"] / steps_eager
        g["calls"] += v["n"] / steps_eager
        g["event_ms"] += v["ms"] / steps_eager
    for name, v in ew.items():
        prefix = ELEMENTWISE_KERNELS.get(name, "pixel_norm_vec_kernel<2" if name == "pn_bwd_mask_second_y" else None)
        g = groups.setdefault(prefix or name, dict(prefix=prefix, flops=0.0, bytes=0.0, calls=0, event_ms=None))
        g["bytes"] += v["bytes"] / steps_eager
        g["calls"] += v["n"] / steps_eager
    rows = []
    for label, g in groups.items():
        us = n = None
        if g["prefix"]:
            hits = [(k, v) for k, v in cupti.items() if k.startswith(g["prefix"])]
            if hits:
                us = sum(v[1] for _, v in hits) / steps_profiled
                n = sum(v[0] for _, v in hits) / steps_profiled
        if us is None and g["event_ms"] is not None:
            us = g["event_ms"] * 1e3          # SIMT convolutions: several kernel names, event time of the eager pass
        if not us:
            continue
        ai = g["flops"] / g["bytes"] if g["bytes"] else 0.0
        bound = "tensor" if ai >= ridge else "hbm"
        tf, gbs = g["flops"] / us / 1e6, g["bytes"] / us / 1e3
        rows.append(dict(kernel=label, launches_per_step=n if n is not None else g["calls"], us_per_step=us,
                         share_of_kernel_time=us * steps_profiled / cupti_total_us,
                         gflop_per_step=g["flops"] / 1e9, mbytes_per_step=g["bytes"] / 1e6, flop_per_byte=ai, bound=bound,
                         tflops=tf, gbs=gbs, frac=(tf / pk["tf_sustained"]) if bound == "tensor" else (gbs / pk["hbm"]),
                         tensor_frac=tf / pk["tf_sustained"], hbm_frac=gbs / pk["hbm"]))
    rows.sort(key=lambda r: -r["us_per_step"])
    return rows


def cpu_model():
    """Model name of the host CPU (reported beside every CPU-timed number)."""
    try:
        with open("/proc/cpuinfo") as f:
            for row in f:
                if row.lower().startswith("model name"):
                    return row.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


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


def inference_secondary(model, device, pk):
    """BASELINE config 5: labels + latents -> generator -> inverse spectral -> waveforms (models.py:232-250) at B = 8 and
    B = 64, with the roofline of the path: F_G = 14.9 GFLOP of convolution / dense work per clip against the sustained bf16
    peak (bf16x3 issues three MMA passes per algorithmic FLOP: 0.33 is a saturated pipe) and the bytes a clip must move
    at least (its 1 304 576-byte spectrogram -> waveform transform plus the fp32 activations of the generator's top two
    resolutions written once and read once) against the copy bandwidth."""
    def timed(batch):
        g = torch.Generator().manual_seed(2)
        lab = torch.nn.functional.one_hot(torch.arange(batch) % 61, 61).float().to(device)
        z = torch.randn(batch, 256, generator=g).to(device)
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
        return s.elapsed_time(e) / reps, out
    ms, out = timed(BATCH)
    res = dict(workload="generate_batch at B=8 (CUDA-graph replay): G forward + convert_to_waveform -> [8, 64000]",
               ms_per_batch=ms, clips_per_s=BATCH / ms * 1e3, output_shape=list(out.shape))
    try:
        ms64, _ = timed(64)
    except Exception as exc:                     # the B = 8 line above stands on its own
        res["roofline"] = dict(error="%s: %s" % (type(exc).__name__, exc))
        return res
    clips64 = 64 / ms64 * 1e3
    # per clip: top two resolutions (128x1024x32 and 64x512x64 fp32, two layers each, written + read) + the inverse transform
    act_bytes = 2 * 2 * 4.0 * (128 * 1024 * 32 + 64 * 512 * 64) + 1304576.0
    res["roofline"] = dict(batch=64, ms_per_batch=ms64, clips_per_s=clips64, flop_per_clip=F_G,
                           tflops=F_G * clips64 / 1e12, tensor_frac=F_G * clips64 / 1e12 / pk["tf_sustained"],
                           min_bytes_per_clip=act_bytes, hbm_gbs_at_min_bytes=act_bytes * clips64 / 1e9,
                           hbm_frac_at_min_bytes=act_bytes * clips64 / 1e9 / pk["hbm"],
                           bound="hbm" if F_G / act_bytes < pk["tf_sustained"] * 1e12 / (pk["hbm"] * 1e9) else "tensor",
                           note="B = 8 is launch / latency bound (about 60 kernels in under a millisecond); batches above 64 run as "
                                "64-clip chunks, so B = 64 is the saturated rate of config 5")
    return res


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

    # ---- per-kernel timing: the graph replay cannot be bracketed kernel by kernel, so (1) the same steps run once
    # more EAGERLY with a CUDA-event pair around every convolution-family ABI call (same inputs, same process), and
    # (2) two replayed iterations run under CUPTI (torch.profiler) for the time per kernel FUNCTION.  Both are outside
    # the timed regions above.
    cupti, cupti_total, cupti_steps = {}, 0.0, min(2, args.steps)
    if rank == 0:
        cupti, cupti_total = cupti_kernel_times(model, dev[args.warmup:args.warmup + cupti_steps], device)
    else:
        run_steps(model, dev[args.warmup:args.warmup + cupti_steps], device, False)     # the steps hold collectives: every rank runs them
    graphs_were = model.use_cuda_graphs
    model.use_cuda_graphs = False
    run_steps(model, dev[:1], device, False)
    barrier()
    with KernelProfiler(Fn.K) as prof:
        start.record()
        t_host = time.perf_counter()
        run_steps(model, dev[args.warmup:], device, False)
        host_ms_eager = (time.perf_counter() - t_host) * 1e3        # what the host needs to enqueue the steps without graphs
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
    # per kernel FUNCTION: CUPTI time of the replayed step joined with the algorithmic work of the eager pass
    table = kernel_table(conv, prof.ew, cupti, cupti_total, cupti_steps, args.steps, pk)
    conv_rows = [r for r in table if r["gflop_per_step"] > 0]
    dominant = max(conv_rows, key=lambda r: r["us_per_step"]) if conv_rows else None
    # the dominant function's heaviest (kernel, shape): per-launch numbers from the event pairs
    shapes = {}
    for key, v in conv.items():
        if dominant is None or conv_kernel_label(key)[0] != dominant["kernel"]:
            continue
        a = shapes.setdefault(key[1:] if key[0] != "conv_w" else key, dict(ms=0.0, n=0, flops=v["flops"], bytes=v["bytes"], key=key))
        a["ms"] += v["ms"]
        a["n"] += v["n"]
    top = max(shapes.values(), key=lambda v: v["ms"])
    top_key = top["key"]
    conv_ms = sum(v["ms"] for v in conv.values())
    avg_ms = top["ms"] / top["n"]
    achieved_tf = top["flops"] / (avg_ms * 1e-3) / 1e12
    achieved_gbs = top["bytes"] / (avg_ms * 1e-3) / 1e9
    ridge = pk["tf_sustained"] * 1e12 / (pk["hbm"] * 1e9)
    top_bound = "tensor" if top["flops"] / top["bytes"] >= ridge else "hbm"
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
                    eager_host_enqueue_ms_per_step=host_ms_eager / args.steps,
                    iterations_per_s=iterations_s,
                    value_counts="batch-8 steps over all ranks (N per lock-step data-parallel iteration)",
                    samples_per_s=value * BATCH),
        clocks=clocks,
        e2e=dict(value=args.steps * world / (ms_e2e / 1e3), unit="steps/s", h2d_bytes_per_step=h2d // args.steps,
                 d2h_bytes_per_step=d2h // args.steps),
        gpu_launches=launches,
        roofline=dict(bound=top_bound,
                      achieved=achieved_tf if top_bound == "tensor" else achieved_gbs,
                      peak=pk["tf_sustained"] if top_bound == "tensor" else pk["hbm"],
                      unit="TFLOP/s" if top_bound == "tensor" else "GB/s",
                      frac=(achieved_tf / pk["tf_sustained"]) if top_bound == "tensor" else (achieved_gbs / pk["hbm"]),
                      traffic=NCU_TRAFFIC.get(top_key, (None, None))[0],
                      traffic_source=NCU_TRAFFIC.get(top_key, (None, "no ncu --set full capture of this kernel/shape"))[1],
                      kernel="%s; heaviest shape: %s n=%d %dx%d ci=%d co=%d k=%d s=%d" % ((dominant["kernel"],) + top_key),
                      chosen_by="largest CUPTI time per replayed step among the kernel FUNCTIONS (see `kernels`); the "
                                "per-launch figures are the heaviest shape of that function",
                      bound_rule="arithmetic intensity %.0f FLOP/B %s ridge %.0f (sustained bf16 peak / copy bandwidth)" %
                                 (top["flops"] / top["bytes"], ">=" if top_bound == "tensor" else "<", ridge),
                      timing="CUDA-event pair around every launch of this kernel in an eager pass of the same steps, each pair "
                             "queued behind a short spin kernel so that it holds device time only, not the host's launch "
                             "latency (the timed region replays CUDA graphs)",
                      launches_timed=top["n"], avg_launch_ms=avg_ms, algorithmic_flop_per_launch=top["flops"],
                      algorithmic_bytes_per_launch=top["bytes"],
                      tensor_tflops=achieved_tf, tensor_frac=achieved_tf / pk["tf_sustained"],
                      hbm_gbs=achieved_gbs, hbm_frac=achieved_gbs / pk["hbm"],
                      function_us_per_step=dominant["us_per_step"], function_share_of_step=dominant["us_per_step"] / (ms / args.steps * 1e3),
                      share_of_eager_step=top["ms"] / ms_eager,
                      conv_family_share_of_eager_step=conv_ms / ms_eager,
                      peak_source="%s: bf16_tflops_sustained / hbm_gbs (kernel timed inside a long step)" % pk["source"],
                      step_tflops=ITER_FLOP_PER_SAMPLE * BATCH * value / 1e12,
                      step_frac=ITER_FLOP_PER_SAMPLE * BATCH * value / 1e12 / (pk["tf_sustained"] * world)),
        kernels=dict(note="per kernel function, per graph-replayed step: CUPTI time (outside the timed region) joined with "
                          "the algorithmic FLOP and bytes (tensors read + written) of the calls it serves; bf16x3 issues "
                          "3 MMA passes per algorithmic FLOP, so tensor_frac 0.33 is a saturated pipe",
                     kernel_time_us_per_step=cupti_total / max(1, cupti_steps), rows=table[:14]),
    )
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count()
        times = oracle_iteration_time(BATCH, 3, cores)[1:]          # first iteration = warm-up (constants, allocator)
        line["cpu_baseline"] = dict(value=1.0 / float(np.mean(times)), unit="steps/s", cores=cores, cpu=cpu_model(), kind="port",
                                    sample="two full iterations (D update + G update) of the oracle port at batch 8 after "
                                           "one warm-up iteration (%.1f s each)" % float(np.mean(times)))
    if args.kernel_table:
        with open(args.kernel_table, "w") as f:
            f.write("%-52s %6s %9s %6s %9s %9s %7s %6s %8s %8s %6s\n" % ("kernel function", "n/step", "us/step", "share", "GFLOP", "MB", "FLOP/B", "bound",
                                                                          "TFLOP/s", "GB/s", "frac"))
            for r in table:
                f.write("%-52s %6.0f %9.1f %6.3f %9.1f %9.1f %7.1f %6s %8.1f %8.0f %6.3f\n" % (
                    r["kernel"][:52], r["launches_per_step"], r["us_per_step"], r["share_of_kernel_time"], r["gflop_per_step"],
                    r["mbytes_per_step"], r["flop_per_byte"], r["bound"], r["tflops"], r["gbs"], r["frac"]))
    if world == 1 and not args.no_spectral:
        line["secondary"] = spectral_secondary(device, pk)
        line["secondary"]["inference"] = inference_secondary(model, device, pk)
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
    # the configuration itself: batch 8, every step a full iteration.  One warm-up iteration always runs (constants,
    # allocator); the rest of the requested warm-up and as many of the requested steps as the budget allows follow.
    probe = oracle_iteration_time(BATCH, 1, cores)[0]
    extra_warm = max(0, min(args.warmup - 1, 1))
    steps = max(1, min(args.steps, int(budget / probe) - extra_warm))
    times = oracle_iteration_time(BATCH, extra_warm + steps, cores)[extra_warm:]
    per = float(np.mean(times))
    value = 1.0 / per
    sample = ("each step = one full iteration (D update + G update) of the CPU oracle port at batch 8 on %d threads; "
              "%d of the requested %d steps timed (%.0f s budget), %d warm-up iteration(s)" %
              (cores, steps, args.steps, budget, 1 + extra_warm))
    print(json.dumps(dict(
        impl="reference", metric="GAN train steps/sec (batch 8/GPU, 128x1024 mel+IF)", value=value, unit="steps/s",
        n_gpus=world, steps=steps, warmup=args.warmup, ms_per_step=1e3 / value, higher_is_better=True, scaling="weak",
        vs_baseline=None, dtype="f32", data="synthetic",
        config=dict(workload="BASELINE configs[1] on host CPU cores via the oracle port (PyTorch-CPU fp32)",
                    global_batch=BATCH, parallelism="cpu"),
        cpu_baseline=dict(value=value, unit="steps/s", cores=cores, cpu=cpu_model(), kind="port", sample=sample),
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
    ap.add_argument("--kernel-table", default=None, help="write the per-kernel-function table (also in the JSON line) here")
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
