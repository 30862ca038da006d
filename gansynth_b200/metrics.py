"""Evaluation statistics of the reference metrics.py:6-63 (host-side numpy / scipy post-processing of classifier
features and logits -- not on the device hot path): inception score, Frechet distance, number of statistically
different bins."""
import numpy as np
import scipy.linalg
import scipy.stats
from sklearn import cluster


def softmax(logits, axis=-1):
    """metrics.py:6-8 (max-shifted)."""
    z = np.asarray(logits) - np.max(logits, axis=axis, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=axis, keepdims=True)


def kl_divergence(p, q, axis=-1):
    """metrics.py:11-12: sum p log(p / q) with 0 log 0 = 0."""
    p, q = np.asarray(p), np.asarray(q)
    with np.errstate(divide="ignore", invalid="ignore"):
        terms = p * np.log(p / q)
    return np.where(p == 0.0, 0.0, terms).sum(axis=axis)


def inception_score(logits):
    """metrics.py:15-18: exp(E_x KL(p(y|x) || p(y)))."""
    cond = softmax(logits)
    marginal = cond.mean(axis=0, keepdims=True)
    return float(np.exp(kl_divergence(cond, marginal).mean()))


def frechet_inception_distance(real_features, fake_features):
    """metrics.py:21-32: |mu_r - mu_f|^2 + tr(C_r + C_f - 2 (C_r C_f)^(1/2)), unbiased covariances (np.cov)."""
    real, fake = np.asarray(real_features, np.float64), np.asarray(fake_features, np.float64)
    mu_r, mu_f = real.mean(axis=0), fake.mean(axis=0)
    cov_r, cov_f = np.cov(real, rowvar=False), np.cov(fake, rowvar=False)
    root = scipy.linalg.sqrtm(cov_r @ cov_f)
    if np.iscomplexobj(root):
        if not np.allclose(np.diagonal(root).imag, 0.0, atol=1.0e-3):
            raise ValueError("Imaginary component %g" % np.abs(root.imag).max())
        root = root.real
    return float(((mu_r - mu_f) ** 2).sum() + np.trace(cov_r + cov_f - 2.0 * root))


def binomial_proportion_test(p, m, q, n, significance_level):
    """metrics.py:35-40: two-sided z-test of two proportions with the pooled standard error."""
    pooled = (p * m + q * n) / (m + n)
    se = np.sqrt(pooled * (1.0 - pooled) * (1.0 / m + 1.0 / n))
    with np.errstate(divide="ignore", invalid="ignore"):
        z = (pooled - q) / se          # the reference tests the POOLED proportion against q (metrics.py:36-38)
    return 2.0 * scipy.stats.norm.cdf(-np.abs(z)) < significance_level


def num_different_bins(real_features, fake_features, num_bins=50, significance_level=0.05, random_state=None):
    """metrics.py:43-63: k-means bins of the real features; count bins whose fake occupancy differs."""
    real, fake = np.asarray(real_features), np.asarray(fake_features)
    km = cluster.KMeans(n_clusters=num_bins, n_init=10, random_state=random_state).fit(real)
    real_prop = np.bincount(km.labels_, minlength=num_bins) / float(len(real))
    d2 = ((fake[:, None, :] - km.cluster_centers_[None, :, :]) ** 2).sum(axis=2)
    fake_prop = np.bincount(d2.argmin(axis=1), minlength=num_bins) / float(len(fake))
    return int(np.count_nonzero(binomial_proportion_test(real_prop, len(real), fake_prop, len(fake), significance_level)))
