"""ctypes binding of libgansynth_b200.so (include/gansynth_b200.h).

There is NO fallback: if the shared library is missing or a call fails, this raises.  Build it with
``python -m gansynth_b200.build`` (or ``__graft_entry__.build()``).
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgansynth_b200.so")
if os.environ.get("GS_LIB", "").startswith("prof"):      # development: a stage-profiling build (gansynth_b200.build --prof)
    LIB_PATH = os.path.join(_HERE, "libgansynth_b200_%s.so" % os.environ["GS_LIB"])

_P = ctypes.c_void_p
_I = ctypes.c_int
_L = ctypes.c_longlong
_F = ctypes.c_float

# name -> argument types (all functions return int); mirrors include/gansynth_b200.h
SIGNATURES = {
    "gs_conv_weight_cache_reset": [],
    "gs_conv_weight_cache_refresh": [_P, _P, _P],
    "gs_conv2d_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _I, _I, _P],
    "gs_conv2d_dgrad": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _I, _I, _P],
    "gs_conv2d_wgrad": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P],
    "gs_conv2d_wgrad_ex": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P],
    "gs_conv2d_fwd_ex": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _I, _I, _P, _P, _F, _I, _P],
    "gs_conv2d_dgrad_ex": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _I, _I, _P, _P, _F, _I, _P],
    "gs_conv2d_transpose_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _I, _I, _P],
    "gs_conv2d_transpose_dgrad": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P],
    "gs_conv2d_transpose_wgrad": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P],
    "gs_dense_fwd": [_P, _P, _P, _I, _I, _I, _F, _P],
    "gs_dense_dgrad": [_P, _P, _P, _I, _I, _I, _F, _P],
    "gs_dense_wgrad": [_P, _P, _P, _I, _I, _I, _F, _P],
    "gs_embedding_fwd": [_P, _P, _P, _I, _I, _F, _P],
    "gs_embedding_bwd": [_P, _P, _P, _I, _I, _I, _F, _P],
    "gs_lrelu": [_P, _P, _L, _P],
    "gs_lrelu_mask_mul": [_P, _P, _P, _L, _P],
    "gs_tanh_fwd": [_P, _P, _L, _P],
    "gs_tanh_bwd": [_P, _P, _P, _L, _P],
    "gs_tanh_bwd2": [_P, _P, _P, _P, _L, _P],
    "gs_bias_act": [_P, _P, _P, _L, _I, _I, _P],
    "gs_row_broadcast": [_P, _P, _L, _I, _P],
    "gs_col_sum": [_P, _P, _L, _I, _P],
    "gs_lrelu_mask_mul_colsum": [_P, _P, _P, _P, _L, _I, _P],
    "gs_axpby": [_P, _P, _P, _F, _F, _L, _P],
    "gs_axpby_dev": [_P, _P, _P, _P, _I, _I, _L, _P],
    "gs_mul": [_P, _P, _P, _F, _L, _P],
    "gs_pixel_norm_fwd": [_P, _P, _P, _L, _I, _F, _P],
    "gs_pixel_norm_bwd": [_P, _P, _P, _P, _L, _I, _P],
    "gs_pixel_norm_bwd2": [_P, _P, _P, _P, _P, _L, _I, _P],
    "gs_pixel_norm_bwd_mask": [_P, _P, _P, _P, _P, _L, _I, _P],
    "gs_pixel_norm_bwd_premask": [_P, _P, _P, _P, _L, _I, _P],
    "gs_pixel_norm_bwd2_masked": [_P, _P, _P, _P, _P, _L, _I, _P],
    "gs_pixel_norm_bwd_mask_y": [_P, _P, _P, _P, _P, _L, _I, _P],
    "gs_pixel_norm_bwd_premask_y": [_P, _P, _P, _P, _L, _I, _P],
    "gs_pixel_norm_bwd2_masked_y": [_P, _P, _P, _P, _P, _L, _I, _P],
    "gs_pixel_norm_bwd2_pair_y": [_P, _P, _P, _P, _P, _P, _L, _I, _P],
    "gs_batch_stddev_fwd": [_P, _P, _I, _L, _I, _F, _P],
    "gs_batch_stddev_bwd": [_P, _P, _P, _I, _L, _I, _F, _P],
    "gs_batch_stddev_bwd2": [_P, _P, _P, _P, _P, _I, _L, _I, _F, _P],
    "gs_upscale2d": [_P, _P, _I, _I, _I, _I, _I, _I, _F, _P],
    "gs_pool2d": [_P, _P, _I, _I, _I, _I, _I, _I, _F, _P],
    "gs_transpose_inner": [_P, _P, _I, _I, _I, _P],
    "gs_row_dot": [_P, _P, _P, _I, _L, _P],
    "gs_row_scale": [_P, _P, _P, _I, _L, _F, _P],
    "gs_adam_step": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _L, _F, _P],
    "gs_adam_slice": [_L, _I, _I, _P, _P],
    "gs_adam_step_allreduce": [_P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _F, _F, _F, _F, _L, _F, _P],
    "gs_spectrogram_fwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "gs_waveform_fwd": [_P, _P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _P],
    "gs_spectrogram_generic": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "gs_waveform_generic": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "gs_group_norm_fwd": [_P, _P, _P, _P, _P, _I, _L, _I, _I, _F, _I, _P],
    "gs_max_pool2d": [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    "gs_spatial_mean": [_P, _P, _I, _L, _I, _P],
    "gs_group_norm_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _L, _I, _I, _F, _I, _P],
    "gs_max_pool2d_bwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "gs_spatial_mean_bwd": [_P, _P, _I, _L, _I, _P],
    "gs_momentum_step": [_P, _P, _P, _P, _L, _F, _F, _I, _F, _P],
    "gs_crc32c": [_P, _L, _P],
    "gs_wav_decode_pcm16": [_P, _L, _P, _I, _P, _P],
    "gs_wav_read_batch": [_P, _I, _P, _I, _I, _P],
    "gs_pcm16_to_float": [_P, _P, _L, _P],
}

_lib = None


class GansynthLibraryError(RuntimeError):
    pass


def load():
    """Loads the shared library (once) and declares every entry point.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GansynthLibraryError(
            "libgansynth_b200.so is not built (%s); run `python -m gansynth_b200.build`. "
            "There is no CPU or PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.gs_last_error.restype = ctypes.c_char_p
    lib.gs_last_error.argtypes = []
    lib.gs_version.restype = _I
    lib.gs_version.argtypes = []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _I
    lib.gs_workspace_bytes.restype = ctypes.c_size_t
    lib.gs_workspace_bytes.argtypes = []
    lib.gs_workspace_min_bytes.restype = ctypes.c_size_t
    lib.gs_workspace_min_bytes.argtypes = []
    lib.gs_context_create.restype = _I
    lib.gs_context_create.argtypes = [_P, ctypes.c_size_t, ctypes.POINTER(_P)]
    lib.gs_context_destroy.restype = _I
    lib.gs_context_destroy.argtypes = [_P]
    lib.gs_context_bind.restype = _I
    lib.gs_context_bind.argtypes = [_P]
    _lib = lib
    return lib


PROBE_SIGNATURES = {
    "gs_tc_probe": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "gs_tc_probe_time": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P],
}
_probe = None


def probe_call(name, *args):
    """Development self-test library (csrc/tc_probe.cu -> libgansynth_b200_probe.so; not part of the product ABI)."""
    global _probe
    if _probe is None:
        lib = ctypes.CDLL(os.path.join(_HERE, "libgansynth_b200_probe.so"))
        lib.gs_last_error.restype = ctypes.c_char_p
        for fname, argtypes in PROBE_SIGNATURES.items():
            getattr(lib, fname).argtypes = argtypes
            getattr(lib, fname).restype = _I
        _probe = lib
    rc = getattr(_probe, name)(*args)
    if rc != 0:
        raise GansynthLibraryError("%s failed (%d): %s" % (name, rc, _probe.gs_last_error().decode()))
    return rc


def is_loaded():
    return _lib is not None


launch_count = 0


def host_call(name, *args):
    """Calls a HOST-side entry point (input-pipeline natives: no kernel launch is counted)."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise GansynthLibraryError("%s failed (%d): %s" % (name, rc, lib.gs_last_error().decode()))
    return rc


# ---- library context: the device workspace is allocated HERE (by the host, through PyTorch) and handed to the library,
# which neither allocates nor frees device memory (include/gansynth_b200.h, gs_context_create).  One context per DEVICE,
# shared by the host threads of this process: the library binds contexts per thread, and PyTorch runs every backward
# function on its autograd worker thread, so each thread binds the device's context on its first call.  (The model's
# threads never run kernels concurrently: backward is synchronous with the caller.)
_tls = threading.local()
_ctx_lock = threading.Lock()
_contexts = {}          # device index -> (workspace tensor, handle): kept alive for the life of the process


def workspace_bytes():
    return int(load().gs_workspace_bytes())


def ensure_context():
    """Creates (once per device) the library context and binds it to the calling host thread."""
    import torch
    dev = torch.cuda.current_device()
    if getattr(_tls, "bound", None) == dev:
        return _contexts[dev][1]
    lib = load()
    with _ctx_lock:
        entry = _contexts.get(dev)
        if entry is None:
            nbytes = int(os.environ.get("GS_WORKSPACE_BYTES", workspace_bytes()))
            ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda:%d" % dev)
            handle = ctypes.c_void_p()
            rc = lib.gs_context_create(ws.data_ptr(), nbytes, ctypes.byref(handle))
            if rc != 0:
                raise GansynthLibraryError("gs_context_create failed (%d): %s" % (rc, lib.gs_last_error().decode()))
            entry = _contexts[dev] = (ws, handle)
    lib.gs_context_bind(entry[1])
    _tls.bound = dev
    return entry[1]


def reset_weight_cache():
    """gs_conv_weight_cache_reset on every context of this process (the cache belongs to the context, not to the thread
    that happens to call)."""
    if _lib is None:
        return
    for dev, (_, handle) in list(_contexts.items()):
        prev = getattr(_tls, "bound", None)
        _lib.gs_context_bind(handle)
        _lib.gs_conv_weight_cache_reset()
        _tls.bound = dev
        if prev is not None and prev != dev:
            _lib.gs_context_bind(_contexts[prev][1])
            _tls.bound = prev


def call(name, *args):
    """Calls a C-ABI entry point; raises GansynthLibraryError with gs_last_error() on failure."""
    global launch_count
    lib = load()
    ensure_context()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise GansynthLibraryError("%s failed (%d): %s" % (name, rc, lib.gs_last_error().decode()))
    launch_count += 1
    return rc
