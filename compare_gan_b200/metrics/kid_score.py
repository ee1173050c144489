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


def _block_sizes(count, n_bins, leading_real_bin=None):
  """Bins of ceil(count / n_bins), the leading ones one shorter so that the sizes add up.  Reference detail
  (kid_score.py:116-119): for the generated set the number of shortened bins is `n_bins * bins_r[0] - n_gen` with
  bins_r[0] read AFTER the real set's own shortening — with unequal set sizes this can leave the generated bins too long
  (the last block is then cut short by slicing).  Reproduced as is."""
  sizes = np.full(n_bins, int(math.ceil(count / n_bins)))
  lead = sizes[0] if leading_real_bin is None else leading_real_bin
  sizes[:(n_bins * lead) - count] -= 1
  assert sizes.min() >= 2
  return sizes


def _cubic(gram_matrix, dim):
  return (gram_matrix / dim + 1) ** 3


def _mean_off_diagonal(k, count):
  return (k.sum() - np.trace(k)) / (count * (count - 1))


def kid(fake_activations, real_activations, max_batch_size=1024, gram=None):
  """Mean over blocks of the unbiased MMD^2 estimate with the cubic kernel.  Both within-set terms are normalised with
  the REAL block's size, as the reference does (`n = r_e - r_s`, kid_score.py:128) — identical whenever the two sets
  have the same number of samples, which is how the evaluation calls it."""
  gram = gram or _gram
  real, fake = np.asarray(real_activations), np.asarray(fake_activations)
  (n_real, dim), (n_fake, dim_fake) = real.shape, fake.shape
  assert dim_fake == dim
  n_bins = int(math.ceil(max(n_real, n_fake) / max_batch_size))
  sizes_r = _block_sizes(n_real, n_bins)
  sizes_f = _block_sizes(n_fake, n_bins, leading_real_bin=int(sizes_r[0]))
  edges_r, edges_f = np.r_[0, np.cumsum(sizes_r)], np.r_[0, np.cumsum(sizes_f)]
  estimates = []
  for lo_r, hi_r, lo_f, hi_f in zip(edges_r[:-1], edges_r[1:], edges_f[:-1], edges_f[1:]):
    r, f = real[lo_r:hi_r], fake[lo_f:hi_f]
    count = float(hi_r - lo_r)
    cross = _cubic(gram(r, f), dim).mean()
    estimates.append(_mean_off_diagonal(_cubic(gram(r, r), dim), count) + _mean_off_diagonal(_cubic(gram(f, f), dim), count)
                     - 2 * cross)
  return float(np.mean(estimates))


class KIDScoreTask(eval_task.EvalTask):
  _LABEL = "kid_score"

  def run_after_session(self, fake_dset, real_dset):
    return {self._LABEL: kid(fake_dset.activations, real_dset.activations)}
