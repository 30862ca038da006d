"""CPU: the C-ABI library loads and exports every symbol include/gansynth_b200.h declares, with the
argument counts the ctypes binding uses (no compute calls: there is no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "gansynth_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|size_t|const char\*)\s+(gs_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    from gansynth_b200 import _lib
    ge.build()
    lib = _lib.load()
    decl = _header_functions()
    assert len(decl) >= 37
    for name, nargs in decl.items():
        assert hasattr(lib, name), "missing export %s" % name
        if name in _lib.SIGNATURES:
            assert len(_lib.SIGNATURES[name]) == nargs, "%s: header has %d args, binding %d" % (
                name, nargs, len(_lib.SIGNATURES[name]))
    for name in _lib.SIGNATURES:
        assert name in decl, "binding %s is not declared in the header" % name
    assert lib.gs_version() >= 100
    assert lib.gs_last_error() is not None


def test_product_refuses_to_run_without_cuda_tensors():
    """No CPU fallback: a CPU tensor must raise, not compute."""
    import torch
    from gansynth_b200 import _lib
    from gansynth_b200.kernels import CudaBackend
    with pytest.raises(_lib.GansynthLibraryError):
        CudaBackend().lrelu(torch.zeros(8))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing in the package (sub-packages included) or under tools/ imports it, and
    at the repo root only bench.py (cpu_baseline / --impl reference legs) and __graft_entry__.py (smoke) do."""
    pattern = re.compile(r"^\s*(from|import)\s+oracle\b", flags=re.M)
    for top in ("gansynth_b200", "tools"):
        for folder, _, files in os.walk(os.path.join(ROOT, top)):
            for fn in files:
                if fn.endswith(".py"):
                    assert not pattern.search(open(os.path.join(folder, fn)).read()), os.path.join(folder, fn)
    at_root = sorted(fn for fn in os.listdir(ROOT) if fn.endswith(".py") and pattern.search(open(os.path.join(ROOT, fn)).read()))
    assert at_root == ["__graft_entry__.py", "bench.py"], at_root


def test_argument_errors_are_reported_before_any_device_work():
    """The entry points validate shapes first: a refused call returns a negative code and a message without touching
    CUDA (so this runs on the CPU box)."""
    import __graft_entry__ as ge
    from gansynth_b200 import _lib
    ge.build()
    lib = _lib.load()
    # 128 frames of hop 512 / length 2048 cover 67072 samples: 70000 is refused
    assert lib.gs_spectrogram_fwd(None, None, None, None, None, None, None, 1, 70000, 128, 32, None) < 0
    assert b"exceeds" in lib.gs_last_error()
    # segments must be a multiple of 8 frames
    assert lib.gs_waveform_fwd(None, None, None, None, None, None, 46, None, None, 1, 64000, 128, 12, None) < 0
    assert b"multiple of 8" in lib.gs_last_error()
    # several runs per clip need the caller's scratch buffer
    assert lib.gs_spectrogram_fwd(None, None, None, None, None, None, None, 1, 64000, 128, 32, None) < 0
    assert b"scratch" in lib.gs_last_error()
    # an empty batch is a no-op
    assert lib.gs_spectrogram_fwd(None, None, None, None, None, None, None, 0, 64000, 128, 32, None) == 0
    assert lib.gs_pcm16_to_float(None, None, 0, None) == 0 and lib.gs_pcm16_to_float(None, None, -1, None) < 0
    with __import__("pytest").raises(_lib.GansynthLibraryError):
        _lib.host_call("gs_wav_read_batch", None, 1, None, 0, 1, None)


def test_context_api_without_a_device():
    """gs_context_*: host-side bookkeeping only (no device work), so it runs here: sizes, argument checks, and the
    entry points that need a workspace refuse to run without a bound context instead of allocating one."""
    import ctypes
    import __graft_entry__ as ge
    from gansynth_b200 import _lib
    ge.build()
    lib = _lib.load()
    assert lib.gs_workspace_bytes() > lib.gs_workspace_min_bytes() >= (8 << 20)
    handle = ctypes.c_void_p()
    assert lib.gs_context_create(None, lib.gs_workspace_bytes(), ctypes.byref(handle)) < 0        # null workspace
    assert lib.gs_context_create(0x1000, 1024, ctypes.byref(handle)) < 0                           # too small
    assert b"gs_workspace_min_bytes" in lib.gs_last_error()
    assert lib.gs_context_create(0x1001, lib.gs_workspace_bytes(), ctypes.byref(handle)) < 0      # misaligned
    lib.gs_context_bind(None)
    # a tensor-core-shaped convolution without a context: refused before any device work
    rc = lib.gs_conv2d_fwd(None, None, None, None, 1, 16, 16, 32, 32, 3, 1, 0, 1.0, 0, 3, None)
    assert rc < 0 and b"no context bound" in lib.gs_last_error()
    rc = lib.gs_spectrogram_fwd(None, None, None, None, None, None, 1, 1, 64000, 128, 32, None)
    assert rc < 0 and b"no context bound" in lib.gs_last_error()
    assert lib.gs_context_create(0x1000, lib.gs_workspace_min_bytes(), ctypes.byref(handle)) == 0  # records the pointer only
    assert lib.gs_context_bind(handle) == 0 and lib.gs_conv_weight_cache_reset() == 0
    assert lib.gs_context_bind(None) == 0 and lib.gs_context_destroy(handle) == 0


def test_spectral_launch_policies():
    from gansynth_b200 import spectral_ops as sp
    # forward: runs of 32 frames once that still gives >= 4 CTAs per SM, else one 16-frame round per CTA
    assert sp.frames_per_run(256, 128) == 32 and sp.frames_per_run(8, 128) == 16 and sp.frames_per_run(148, 128) == 32
    # inverse: whole clips when the batch fills the SMs, 8-frame multiples otherwise
    assert sp.frames_per_segment(256, 128) == 128 and sp.frames_per_segment(148, 128) == 128
    assert sp.frames_per_segment(8, 128) == 8 and sp.frames_per_segment(64, 128) == 64 and sp.frames_per_segment(1, 128) == 8
    assert sp.frames_per_segment(3, 20) == 16 and sp.frames_per_segment(3, 4) == 4
    for b in (1, 2, 5, 8, 33, 100, 300):
        for t in (4, 8, 20, 128):
            f = sp.frames_per_segment(b, t)
            assert f >= t or f % 8 == 0


def test_adam_slices_partition_the_flat_buffer():
    """gs_adam_slice: the per-rank slices of the fused all-reduce + Adam kernel are 16-byte aligned, disjoint and cover
    the buffer for every world size (host arithmetic only)."""
    import ctypes
    import __graft_entry__ as ge
    from gansynth_b200 import _lib
    ge.build()
    lib = _lib.load()
    for n in (0, 4, 32, 6830976, 8932256, 1000):
        for world in (1, 2, 3, 4, 8):
            edges = []
            for rank in range(world):
                lo, hi = ctypes.c_longlong(), ctypes.c_longlong()
                assert lib.gs_adam_slice(n, rank, world, ctypes.byref(lo), ctypes.byref(hi)) == 0
                assert lo.value % 4 == 0 and hi.value % 4 == 0 and lo.value <= hi.value
                edges.append((lo.value, hi.value))
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
    lo, hi = ctypes.c_longlong(), ctypes.c_longlong()
    assert lib.gs_adam_slice(10, 0, 2, ctypes.byref(lo), ctypes.byref(hi)) < 0          # not a multiple of 4
    assert lib.gs_adam_step_allreduce(None, None, None, None, None, None, None, 8, 0, 2, 1e-3, 0.0, 0.99, 1e-8, 1, 0.5, None) < 0
    assert b"multicast addresses or the arrays of peer pointers" in lib.gs_last_error()
