"""Builds libgansynth_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the
repo snapshot to the GPU box)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgansynth_b200.so")
SOURCES = ["abi.cu", "conv.cu", "elementwise.cu", "dense.cu", "spectral.cu", "spectral_generic.cu", "io.cu", "classifier.cu"]
PROBE_LIB = os.path.join(HERE, "libgansynth_b200_probe.so")     # development self-test, not the product ABI
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_prof(verbose=False, defs=(), suffix="prof"):
    """The stage-profiling variant (conv.cu with -DGS_TC_PROF) -> libgansynth_b200_prof.so; a development tool
    (tools/tc_stage_profile.py, GS_LIB=prof), never loaded by default.  `defs` / `suffix`: further experiment builds
    (e.g. -DTCW_CW=16 -> libgansynth_b200_prof16.so, GS_LIB=prof16)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    build()
    o = os.path.join(CSRC, "conv_%s.o" % suffix)
    cmd = [nvcc] + NVCC_FLAGS + ["-DGS_TC_PROF"] + list(defs) + ["-c", os.path.join(CSRC, "conv.cu"), "-o", o]
    subprocess.check_call(cmd)
    objs = [os.path.join(CSRC, s.replace(".cu", ".o")) for s in SOURCES if s != "conv.cu"] + [o]
    lib = os.path.join(HERE, "libgansynth_b200_%s.so" % suffix)
    subprocess.check_call([nvcc, "-shared", "-o", lib] + objs + ["-lcudart"])
    return lib


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "gansynth_b200.h"))
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write("== nvcc %s ==\n%s\n" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lcudart"]
        subprocess.check_call(cmd)
    probe_src = os.path.join(CSRC, "tc_probe.cu")
    if force or _stale(PROBE_LIB, [probe_src, os.path.join(CSRC, "abi.o")] + headers):
        subprocess.check_call([nvcc] + NVCC_FLAGS + ["-shared", "-o", PROBE_LIB, probe_src, os.path.join(CSRC, "abi.o"), "-lcudart"])
    return LIB


if __name__ == "__main__":
    if "--prof" in sys.argv:
        print(build_prof())
    elif "--prof16" in sys.argv:
        print(build_prof(defs=["-DTCW_CW=16"], suffix="prof16"))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
