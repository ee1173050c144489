"""Gradient penalties (reference gans/penalty_lib.py:28-108)."""
from .. import gin_lite as gin
from .. import kernels as K
from .. import tape
from .. import utils


_ALPHA_RNG = {"seed": 0x5EEDA1FA, "offset": 0}     # stream of the un-fed interpolation coefficients (per process)


@gin.configurable
def no_penalty():
  """reference penalty_lib.py:28-30."""
  return K.zeros(1)


@gin.configurable(whitelist=[])
def wgangp_penalty(discriminator, x, x_fake, y, is_training, alpha=None):
  """WGAN gradient penalty (reference penalty_lib.py:59-82).  `alpha` [B,1,1,1] may be fed (parity tests,
  bench); otherwise it is drawn U[0,1) on the device."""
  if alpha is None:
    # tf.random.uniform (penalty_lib.py:72-73) through the library's counter-based generator (cgan_random_uniform): no
    # torch op on the product path; successive draws advance the offset
    from ..tpu import tpu_ops
    alpha = K.empty(x.shape[0], 1, 1, 1)
    # every replica draws its own coefficients (the reference's per-replica tf.random.uniform): the rank selects the stream
    K._call("random_uniform", alpha.ptr, alpha.numel, _ALPHA_RNG["seed"] + tpu_ops.replica_id(), _ALPHA_RNG["offset"])
    _ALPHA_RNG["offset"] += alpha.numel
  interpolates = K.interpolate(x, x_fake, alpha)
  interpolates.req = True                      # differentiate the logits wrt this leaf
  with tape.record(True):
    logits = discriminator(interpolates, y=y, is_training=is_training, reuse=True)[1]
    ones = K.fill_(K.empty(*logits.shape), 1.0)
    (gradients,) = tape.backward([(logits, ones)], [interpolates], K.add_grad, create_graph=True)
    return K.gp_penalty(gradients)


@gin.configurable("penalty", whitelist=["fn"])
def get_penalty_loss(fn=no_penalty, **kwargs):
  """Returns the penalty loss (reference penalty_lib.py:105-108)."""
  return utils.call_with_accepted_args(fn, **kwargs)
