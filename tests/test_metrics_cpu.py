"""CPU: evaluation statistics (reference metrics.py:6-63) against analytic answers."""
import numpy as np

from gansynth_b200 import metrics


def test_softmax_and_kl():
    x = np.array([[1.0, 2.0, 3.0], [1000.0, 1000.0, 1000.0]])
    p = metrics.softmax(x)
    assert np.allclose(p.sum(1), 1.0) and np.allclose(p[1], 1.0 / 3.0)
    assert np.allclose(p[0], np.exp(x[0]) / np.exp(x[0]).sum())
    assert metrics.kl_divergence(np.array([0.5, 0.5, 0.0]), np.array([0.25, 0.25, 0.5])) == np.log(2.0)


def test_inception_score_limits():
    assert abs(metrics.inception_score(np.zeros((100, 61))) - 1.0) < 1e-12           # uniform posteriors
    conf = np.full((61 * 4, 61), -1e4)
    conf[np.arange(61 * 4), np.arange(61 * 4) % 61] = 1e4                             # confident and balanced
    assert abs(metrics.inception_score(conf) - 61.0) < 1e-6


def test_frechet_distance_of_gaussians():
    g = np.random.default_rng(0)
    a = g.normal(size=(20000, 4))
    assert metrics.frechet_inception_distance(a, a) < 1e-8
    # b = 2 a + d: cov_b = 4 cov_a, (cov_a cov_b)^(1/2) = 2 cov_a  =>  FID = |mu_a - mu_b|^2 + tr(cov_a)
    b = a * 2.0 + np.array([1.0, 0.0, -2.0, 0.0])
    want = ((a.mean(0) - b.mean(0)) ** 2).sum() + np.trace(np.cov(a, rowvar=False))
    assert abs(metrics.frechet_inception_distance(a, b) - want) < 1e-6 and 8.0 < want < 10.0


def test_binomial_test_and_ndb():
    same = metrics.binomial_proportion_test(np.array([0.2, 0.2]), 1000, np.array([0.2, 0.9]), 1000, 0.05)
    assert list(same) == [False, True]
    g = np.random.default_rng(1)
    real = g.normal(size=(2000, 3))
    assert metrics.num_different_bins(real, g.normal(size=(2000, 3)), num_bins=10, random_state=0) <= 2
    assert metrics.num_different_bins(real, g.normal(size=(2000, 3)) * 0.05, num_bins=10, random_state=0) >= 7
