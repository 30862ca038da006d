"""CPU oracle for the GANSynth hot path (TEST INFRASTRUCTURE -- not product code).

This package is a plain PyTorch-CPU (fp32, optional fp64) restatement of the arithmetic of
skmhrk1209/GANSynth @ d135d40 for the path BASELINE.json's north_star names:

    ops.py:149-348, networks.py:1-290, spectral_ops.py:8-149, models.py:22-89

with the TensorFlow-1.13 op semantics listed in SURVEY.md Appendix B.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures, and TensorFlow 1.13 /
tensorflow_probability cannot be imported or built in this image, so this restatement cannot be
checked against outputs of the reference itself.  It is pinned only by (a) analytic known-answer
tests and independent numpy/scipy cross-checks in tests/test_oracle_*.py and (b) the committed
fixtures under tests/golden/ that it generated itself (tools/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product package gansynth_b200/ never does.
"""
