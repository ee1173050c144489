"""GAN losses (reference gans/loss_lib.py:30-154): same functions, same argument names, same errors;
the arithmetic is one fused warp-shuffle reduction kernel (cgan_gan_loss)."""
from .. import gin_lite as gin
from .. import kernels as K
from .. import utils


def check_dimensions(d_real, d_fake, d_real_logits, d_fake_logits):
  """Checks the shapes and ranks of logits and prediction tensors (reference loss_lib.py:30-50)."""
  def _check_pair(a, b):
    if a != b:
      raise ValueError("Shape mismatch: %s vs %s." % (a, b))
    if len(a) != 2 or len(b) != 2:
      raise ValueError("Rank: expected 2, got %s and %s" % (len(a), len(b)))
  if (d_real is not None) and (d_fake is not None):
    _check_pair(list(d_real.shape), list(d_fake.shape))
  if (d_real_logits is not None) and (d_fake_logits is not None):
    _check_pair(list(d_real_logits.shape), list(d_fake_logits.shape))
  if (d_real is not None) and (d_real_logits is not None):
    _check_pair(list(d_real.shape), list(d_real_logits.shape))


def _fused(kind, d_real, d_fake, d_real_logits, d_fake_logits):
  check_dimensions(d_real, d_fake, d_real_logits, d_fake_logits)
  if d_real_logits is None or d_fake_logits is None:
    raise ValueError("The B200 loss kernel works from logits; pass d_real_logits and d_fake_logits.")
  return K.gan_losses(kind, d_real_logits, d_fake_logits)


@gin.configurable(whitelist=[])
def non_saturating(d_real_logits, d_fake_logits, d_real=None, d_fake=None):
  """reference loss_lib.py:53-78."""
  return _fused("non_saturating", d_real, d_fake, d_real_logits, d_fake_logits)


@gin.configurable(whitelist=[])
def wasserstein(d_real_logits, d_fake_logits, d_real=None, d_fake=None):
  """reference loss_lib.py:81-101."""
  return _fused("wasserstein", d_real, d_fake, d_real_logits, d_fake_logits)


@gin.configurable(whitelist=[])
def least_squares(d_real, d_fake, d_real_logits=None, d_fake_logits=None):
  """reference loss_lib.py:104-124 (d = sigmoid(logits) is recomputed inside the kernel)."""
  return _fused("least_squares", d_real, d_fake, d_real_logits, d_fake_logits)


@gin.configurable(whitelist=[])
def hinge(d_real_logits, d_fake_logits, d_real=None, d_fake=None):
  """reference loss_lib.py:127-148."""
  return _fused("hinge", d_real, d_fake, d_real_logits, d_fake_logits)


@gin.configurable("loss", whitelist=["fn"])
def get_losses(fn=non_saturating, **kwargs):
  """Returns the losses for the discriminator and generator (reference loss_lib.py:151-154)."""
  return utils.call_with_accepted_args(fn, **kwargs)
