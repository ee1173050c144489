"""Oracle (test infrastructure): FID / Inception Score / KID in NumPy float64.

FID follows tfgan.eval.frechet_classifier_distance_from_activations as called at
metrics/fid_score.py:49-51 (TF-GAN, third-party `tensorflow-gan==0.0.0.dev0`, not
vendored in /root/reference): float64, UNBIASED covariance, trace of the matrix
square root of sigma_r*sigma_g via the symmetric form sqrt(sigma)^T sigma_g sqrt(sigma).
Pinned by metrics/fid_score_test.py:31-40 (89.091 +- 1e-4).
IS follows tfgan.eval.classifier_score_from_logits (metrics/inception_score.py:44):
exp(mean_x KL(p(y|x) || p(y))) — PARITY UNPINNED (no golden in the reference).
KID follows metrics/kid_score.py:44-149 literally, including its `n = r_e - r_s`
quirk — PARITY UNPINNED (the reference has no KID test).
"""
import math

import numpy as np
import scipy.linalg


def _sqrtm_psd(m):
  """Symmetric matrix square root through SVD (TF-GAN's _symmetric_matrix_square_root)."""
  u, s, vt = np.linalg.svd(m)
  eps = 1e-10
  si = np.where(s < eps, s, np.sqrt(s))
  return (u * si) @ vt


def trace_sqrt_product(sigma, sigma_v):
  sqrt_sigma = _sqrtm_psd(sigma)
  prod = sqrt_sigma @ sigma_v @ sqrt_sigma
  return np.trace(_sqrtm_psd(prod))


def fid_from_moments(mu_r, sigma_r, mu_g, sigma_g):
  mu_r, mu_g = np.asarray(mu_r, np.float64), np.asarray(mu_g, np.float64)
  tr = np.trace(sigma_r) + np.trace(sigma_g) - 2.0 * trace_sqrt_product(sigma_r, sigma_g)
  return float(tr + np.sum((mu_r - mu_g) ** 2))


def compute_fid_from_activations(real, fake):
  """metrics/fid_score.py:60-75."""
  real = np.asarray(real, np.float64)
  fake = np.asarray(fake, np.float64)
  mu_r, mu_g = real.mean(0), fake.mean(0)
  sr = np.cov(real, rowvar=False, ddof=1).reshape(real.shape[1], real.shape[1])
  sg = np.cov(fake, rowvar=False, ddof=1).reshape(fake.shape[1], fake.shape[1])
  return fid_from_moments(mu_r, sr, mu_g, sg)


def inception_score_from_logits(logits):
  """metrics/inception_score.py:44 -> tfgan classifier_score_from_logits."""
  logits = np.asarray(logits, np.float64)
  m = logits.max(1, keepdims=True)
  logp = logits - m - np.log(np.exp(logits - m).sum(1, keepdims=True))
  p = np.exp(logp)
  marg = p.mean(0, keepdims=True)
  log_marg = np.log(marg)
  kl = (p * (logp - log_marg)).sum(1)
  return float(np.exp(kl.mean()))


def kid(fake, real, max_batch_size=1024):
  """metrics/kid_score.py:44-149 (block estimator, cubic kernel)."""
  real = np.asarray(real, np.float64)
  fake = np.asarray(fake, np.float64)
  n_real, dim = real.shape
  n_gen, dim2 = fake.shape
  assert dim2 == dim
  n_bins = int(math.ceil(max(n_real, n_gen) / max_batch_size))
  bins_r = np.full(n_bins, int(math.ceil(n_real / n_bins)))
  bins_g = np.full(n_bins, int(math.ceil(n_gen / n_bins)))
  bins_r[:(n_bins * bins_r[0]) - n_real] -= 1
  bins_g[:(n_bins * bins_r[0]) - n_gen] -= 1
  assert bins_r.min() >= 2 and bins_g.min() >= 2
  inds_r = np.r_[0, np.cumsum(bins_r)]
  inds_g = np.r_[0, np.cumsum(bins_g)]
  ests = []
  for i in range(n_bins):
    r = real[inds_r[i]:inds_r[i + 1]]
    g = fake[inds_g[i]:inds_g[i + 1]]
    m = float(inds_r[i + 1] - inds_r[i])
    n = float(inds_r[i + 1] - inds_r[i])     # sic: kid_score.py:128 uses the REAL bin size
    k_rr = (r @ r.T / dim + 1) ** 3
    k_rg = (r @ g.T / dim + 1) ** 3
    k_gg = (g @ g.T / dim + 1) ** 3
    ests.append(-2 * k_rg.mean() + (k_rr.sum() - np.trace(k_rr)) / (m * (m - 1))
                + (k_gg.sum() - np.trace(k_gg)) / (n * (n - 1)))
  return float(np.mean(ests))
