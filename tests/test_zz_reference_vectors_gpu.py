"""GPU: the CUDA path against the vectors the reference's own code produced (tests/golden/reference_*.npz, written by
tests/golden/make_reference_vectors.py from the unmodified networks.py / spectral_ops.py / models.py of the reference):
fp32 kernels against float64 reference values at north_star's 1e-3 relative.  The other GPU tests compare the kernels
with the oracle at the benchmark's sizes; tests/test_reference_pin_cpu.py pins the oracle to these same files."""
import pytest
import torch

import reference_vectors as rv

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["fp32", "tc"])
def conv_mode(request):
    import gansynth_b200.functional as F
    prev = F.K.impl
    F.K.impl = 4 if request.param == "fp32" else 0
    yield request.param
    F.K.impl = prev


def test_generator_and_discriminator_at_every_growth_regime(cuda_store, conv_mode):
    rv.check_forward(cuda_store, "cuda")


def test_spectral_both_ways(cuda_store):
    rv.check_spectral("cuda")


@pytest.mark.parametrize("fixture", ["reference_step", "reference_step_fake_penalty"])
def test_training_sequence(cuda_store, conv_mode, fixture):
    rv.check_training_sequence(cuda_store, "cuda", fixture)


def test_classifier_training_and_export_head(cuda_store):
    rv.check_classifier(cuda_store, "cuda")


def test_full_size_batch8_step(cuda_store, conv_mode):
    """BASELINE config 2 itself against the reference's own code: see reference_vectors.check_full_step.  Forward values
    and losses at 1e-3.  The sampled gradient elements are compared WITHOUT pinning the leaky-relu masks, so their bound is
    the spread fp32 arithmetic itself shows against the float64 reference at this size (torch-CPU fp32: 3.9e-3 of a
    variable's maximum), not north_star's 1e-3 -- that claim is test_model_gpu.py's mask-pinned criterion."""
    # measured on the B200 (profiles/pytest_gpu_r2_refvec_full.txt): norms within 1.4e-3 / 2.9e-3 (6.5e-3 in another run:
    # bias gradients are sums of a million cancelling terms accumulated by atomics), samples within 5.1e-3 / 1.4e-2;
    # the bounds leave a factor of four to seven for that run-to-run spread
    if conv_mode == "fp32":
        rv.check_full_step(cuda_store, "cuda", sample_tol=2e-2, norm_tol=1e-2)
    else:
        rv.check_full_step(cuda_store, "cuda", sample_tol=6e-2, norm_tol=3e-2)


def test_odd_architectures(cuda_store, conv_mode):
    rv.check_architectures("cuda")


def test_spectral_configurations(cuda_store):
    rv.check_spectral_configs("cuda")


def test_baseline_config1_sequence(cuda_store, conv_mode):
    """BASELINE configs[0] as the reference itself runs it: reference_vectors.check_config1."""
    if conv_mode == "fp32":
        rv.check_config1(cuda_store, "cuda", sample_tol=2e-2, norm_tol=1e-2)
    else:
        rv.check_config1(cuda_store, "cuda", sample_tol=6e-2, norm_tol=3e-2)
