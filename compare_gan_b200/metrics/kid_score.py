"""Kernel Inception Distance (reference metrics/kid_score.py:34-149): block estimator with the cubic
kernel (x.y/d + 1)^3; the three Gram GEMMs of each block run on the GPU (cgan_gemm), the cubic /
diagonal-removal / means are finished on the host in float64.  Restates the reference literally,
including `n = r_e - r_s` at kid_score.py:128."""
import math

import numpy as np

from . import eval_task


def _gram(a, b):
  """a @ b.T on the device through the C-ABI GEMM."""
  from .. import kernels as K
  ad, bd = K.from_numpy(np.asarray(a, np.float32)), K.from_numpy(np.asarray(b, np.float32))
  return K.matmul(ad, bd, False, True).cpu().astype(np.float64)


def kid(fake_activations, real_activations, max_batch_size=1024, gram=None):
  gram = gram or _gram
  real, fake = np.asarray(real_activations), np.asarray(fake_activations)
  n_real, dim = real.shape
  n_gen, dim2 = fake.shape
  assert dim2 == dim
  n_bins = int(math.ceil(max(n_real, n_gen) / max_batch_size))
  bins_r = np.full(n_bins, int(math.ceil(n_real / n_bins)))
  bins_g = np.full(n_bins, int(math.ceil(n_gen / n_bins)))
  bins_r[:(n_bins * bins_r[0]) - n_real] -= 1
  bins_g[:(n_bins * bins_r[0]) - n_gen] -= 1
  assert bins_r.min() >= 2
  assert bins_g.min() >= 2
  inds_r = np.r_[0, np.cumsum(bins_r)]
  inds_g = np.r_[0, np.cumsum(bins_g)]
  ests = []
  for i in range(n_bins):
    r = real[inds_r[i]:inds_r[i + 1]]
    g = fake[inds_g[i]:inds_g[i + 1]]
    m = float(inds_r[i + 1] - inds_r[i])
    n = float(inds_r[i + 1] - inds_r[i])
    k_rr = (gram(r, r) / dim + 1) ** 3
    k_rg = (gram(r, g) / dim + 1) ** 3
    k_gg = (gram(g, g) / dim + 1) ** 3
    ests.append(-2 * k_rg.mean() + (k_rr.sum() - np.trace(k_rr)) / (m * (m - 1))
                + (k_gg.sum() - np.trace(k_gg)) / (n * (n - 1)))
  return float(np.mean(ests))


class KIDScoreTask(eval_task.EvalTask):
  _LABEL = "kid_score"

  def run_after_session(self, fake_dset, real_dset):
    return {self._LABEL: kid(fake_dset.activations, real_dset.activations)}
