"""CPU oracle for the GANSynth hot path (TEST INFRASTRUCTURE -- not product code).

This package is a plain PyTorch-CPU (fp32, optional fp64) restatement of the arithmetic of
skmhrk1209/GANSynth @ d135d40 for the path BASELINE.json's north_star names:

    ops.py:149-348, networks.py:1-290, spectral_ops.py:8-149, models.py:22-89

with the TensorFlow-1.13 op semantics listed in SURVEY.md Appendix B.

PARITY: pinned to the reference's own Python, NOT to TensorFlow's kernels.  The reference ships no tests, golden
vectors or fixtures, and TensorFlow 1.13 / tensorflow_probability cannot be imported or built in this image.  The
reference is, however, a pure-Python composition of TensorFlow primitives: oracle/tf1_eager/ restates those primitives
eagerly (from the TF-1.13 API semantics), tests/golden/make_reference_vectors.py imports the
UNMODIFIED networks.py / spectral_ops.py / models.py from /root/reference over it and commits what they produce
(tests/golden/reference_*.npz), and tests/test_reference_pin_cpu.py holds this restatement to those vectors in float64
at 1e-10 (forward at every growth regime, spectral both ways, the D-run / G-run training sequence with every gradient and
Adam update, the pitch classifier) and re-runs the generator wherever /root/reference exists.  What stays unpinned is the
behaviour of the TensorFlow kernels under those primitives; that layer is covered by (a) analytic known-answer tests and
independent numpy / scipy / torchaudio / torch.istft cross-checks in tests/test_oracle_cpu.py and (b) nothing else.
tests/golden/small_step.npz and spectral.npz are older fixtures this package generated itself (tests/tools/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product package gansynth_b200/ never does.
"""
