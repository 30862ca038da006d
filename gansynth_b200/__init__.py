"""gansynth_b200: B200-native (sm_100a) GANSynth hot path -- hand-written CUDA kernels behind a C ABI
(include/gansynth_b200.h), driven by a PyTorch host that mirrors the reference's module surface:

    gansynth_b200.ops           reference ops.py
    gansynth_b200.networks      reference networks.py (PGGAN)
    gansynth_b200.spectral_ops  reference spectral_ops.py
    gansynth_b200.models        reference models.py (GANSynth)
    gansynth_b200.utils         reference utils.py (Dict)

Importing the package does not load the shared library; the first kernel call does and raises if it
has not been built (no CPU / PyTorch fallback).
"""
__version__ = "0.1.0"
