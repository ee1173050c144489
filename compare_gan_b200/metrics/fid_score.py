"""Frechet Inception Distance (reference metrics/fid_score.py:39-75).

TF-GAN's frechet_classifier_distance_from_activations is replaced by: a float64 streaming
(sum, sum x x^T) accumulator on the GPU (cgan_cov_accumulate, all-reduced across ranks when eval is
sharded) and the 2048x2048 matrix square roots on the host in float64 (LAPACK via numpy), as the
reference does its final scalar arithmetic in float64 too."""
import numpy as np

from . import eval_task

FID_CODE_FAILED = 4242.0   # reference fid_score.py:36


def moments_from_sums(s, sxx, n):
  """mean and UNBIASED covariance from accumulated sums (tfgan uses N-1; pinned by fid_score_test.py:31-40)."""
  s, sxx = np.asarray(s, np.float64), np.asarray(sxx, np.float64)
  mu = s / n
  sigma = (sxx - n * np.outer(mu, mu)) / (n - 1)
  return mu, sigma


def _sqrtm_psd(m):
  u, s, vt = np.linalg.svd(m)
  si = np.where(s < 1e-10, s, np.sqrt(s))
  return (u * si) @ vt


def fid_from_moments(mu_r, sigma_r, mu_g, sigma_g):
  sq = _sqrtm_psd(sigma_r)
  tr_sqrt = np.trace(_sqrtm_psd(sq @ sigma_g @ sq))
  return float(np.trace(sigma_r) + np.trace(sigma_g) - 2.0 * tr_sqrt + np.sum((np.asarray(mu_r) - np.asarray(mu_g)) ** 2))


def compute_fid_from_activations(fake_activations, real_activations):
  """Returns the FID based on activations (reference fid_score.py:60-75)."""
  fake = np.asarray(fake_activations, np.float64)
  real = np.asarray(real_activations, np.float64)
  mu_g, mu_r = fake.mean(0), real.mean(0)
  d = fake.shape[1]
  sg = np.cov(fake, rowvar=False, ddof=1).reshape(d, d)
  sr = np.cov(real, rowvar=False, ddof=1).reshape(d, d)
  return fid_from_moments(mu_r, sr, mu_g, sg)


class FIDScoreTask(eval_task.EvalTask):
  """Evaluation task for the FID score (reference fid_score.py:39-57)."""
  _LABEL = "fid_score"

  def run_after_session(self, fake_dset, real_dset):
    if getattr(fake_dset, "moments", None) is not None and getattr(real_dset, "moments", None) is not None:
      (mg, sg), (mr, sr) = fake_dset.moments, real_dset.moments
      return {self._LABEL: fid_from_moments(mr, sr, mg, sg)}
    return {self._LABEL: compute_fid_from_activations(fake_dset.activations, real_dset.activations)}
