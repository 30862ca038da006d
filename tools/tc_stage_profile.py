"""Where the warp-specialised convolution kernels wait.  Needs the stage-profiling build
(`python -m gansynth_b200.build --prof`; GS_LIB=prof is set here): every role sums the cycles it spends blocked on each
of its mbarriers and its total run time (csrc/tc_common.cuh, TC_WAIT / TC_PROF_FLUSH).
usage: python tools/tc_stage_profile.py            # the representative layers of the full-size step
       python tools/tc_stage_profile.py c 8 128 1024 32 32 1 [mask|pn]"""
import ctypes
import os
import sys

os.environ.setdefault("GS_LIB", "prof")
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gansynth_b200 import _lib  # noqa: E402
from gansynth_b200.kernels import CudaBackend  # noqa: E402

ROLES = {
    0: ("tck converters", "raw_full", "a_empty", "-"), 1: ("tck TMA loads", "raw_empty", "-", "-"),
    2: ("tck MMA issue", "acc_empty", "a_full", "b_full"), 3: ("tck epilogue 0", "acc_full", "store/bar", "aux_full"),
    4: ("tck epilogue 1", "acc_full", "store/bar", "aux_full"), 5: ("tck epi 0 parts", "(TMEM loads + shift-add)", "(release acc)", "(fence.proxy.async)"),
    8: ("tcw converters", "raw_full", "empty", "-"), 9: ("tcw TMA loads", "raw_empty", "-", "-"),
    10: ("tcw MMA issue", "full", "-", "-"), 11: ("tcw drain", "last MMA", "-", "-"),
    16: ("tc converters", "raw_full", "a_empty", "-"), 17: ("tc TMA loads", "raw_empty", "-", "-"),
    18: ("tc MMA issue", "acc_empty", "a_full", "b_full"), 19: ("tc epilogue", "acc_full", "store/bar", "aux_full"),
}


def read_table():
    lib = _lib.load()
    buf = (ctypes.c_ulonglong * 256)()
    lib.gs_tc_prof_read.argtypes = [ctypes.c_void_p]
    lib.gs_tc_prof_read.restype = ctypes.c_int
    assert lib.gs_tc_prof_read(buf) == 0
    return [[buf[4 * r + i] for i in range(4)] for r in range(64)]


def run(form, n, h, w, ci, co, st, epi=None, reps=3):
    k = CudaBackend()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, h, w, ci, generator=g).cuda()
    dy = torch.randn(n, h // st, w // st, co, generator=g).cuda()
    wt = torch.randn(3, 3, ci, co, generator=g).cuda()
    b = torch.randn(co, generator=g).cuda()

    def call():
        if form == "c":
            if epi == "pn":
                return k.conv_pn(x, wt, b, "c", 3, st, 0, 0.05, 1e-8)
            return k.conv_c(x, wt, None if epi == "mask" else b, 3, st, 0, 0.05, 0 if epi == "mask" else 1,
                            mask_src=dy if epi == "mask" else None)
        if form == "t":
            return k.conv_t(dy, wt, None, 3, st, 0, 0.05, 0, mask_src=x if epi == "mask" else None)
        return k.conv_w(x, dy, 3, st, 0, 0.05, bias_of="dy" if epi == "bias" else None)

    call()
    read_table()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        call()
    ev[1].record()
    torch.cuda.synchronize()
    us = ev[0].elapsed_time(ev[1]) * 1e3 / reps
    tab = read_table()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    print("%s n=%d %dx%d ci=%d co=%d s=%d %s: %.1f us per call (profiling build)" % (form, n, h, w, ci, co, st, epi or "", us))
    for r, (name, w0, w1, w2) in ROLES.items():
        t = tab[r]
        if t[3] == 0:
            continue
        tot = t[3]
        print("   %-16s run %7.1f us/CTA  waits: %s %4.1f%%  %s %4.1f%%  %s %4.1f%%  -> busy %4.1f%%" % (
            name, tot / reps / sms / 1.965e3, w0, 100.0 * t[0] / tot, w1, 100.0 * t[1] / tot, w2, 100.0 * t[2] / tot,
            100.0 * (tot - t[0] - t[1] - t[2]) / tot))


if len(sys.argv) > 7:
    run(sys.argv[1], *(int(v) for v in sys.argv[2:8]), epi=sys.argv[8] if len(sys.argv) > 8 else None)
else:
    for case in [("c", 8, 128, 1024, 32, 32, 1, None), ("t", 8, 128, 1024, 32, 32, 1, "mask"), ("c", 8, 128, 1024, 32, 32, 1, "pn"),
                 ("c", 8, 64, 512, 64, 64, 1, None), ("c", 8, 32, 256, 128, 128, 1, None), ("c", 8, 16, 128, 256, 256, 1, None),
                 ("c", 8, 128, 1024, 32, 64, 2, None), ("t", 8, 128, 1024, 32, 64, 2, None),
                 ("w", 8, 128, 1024, 32, 32, 1, None), ("w", 8, 128, 1024, 32, 32, 1, "bias"), ("w", 8, 64, 512, 64, 64, 1, None),
                 ("w", 8, 128, 1024, 32, 64, 2, None), ("w", 8, 32, 256, 128, 128, 1, None), ("w", 8, 16, 128, 256, 256, 1, None),
                 ("w", 8, 8, 64, 256, 256, 1, None), ("w", 8, 2, 16, 256, 256, 1, None)]:
        run(*case)
