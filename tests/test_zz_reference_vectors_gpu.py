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
