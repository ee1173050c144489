"""Generator / Discriminator interfaces (reference architectures/abstract_arch.py:48-146)."""
from .. import gin_lite as gin
from .. import utils
from .. import variables as V


class _Module(object):
  def __init__(self, name):
    self._name = name

  @property
  def name(self):
    return self._name

  @property
  def trainable_variables(self):
    return list(V.current().trainable_under(self._name).values())


@gin.configurable("G", blacklist=["name", "image_shape"])
class AbstractGenerator(_Module):
  """Interface for generator architectures."""

  def __init__(self, name="generator", image_shape=None, batch_norm_fn=None, spectral_norm=False):
    super(AbstractGenerator, self).__init__(name=name)
    self._image_shape = image_shape
    self._batch_norm_fn = batch_norm_fn
    self._spectral_norm = spectral_norm

  def __call__(self, z, y, is_training, reuse=None):
    with V.variable_scope(self.name):
      return self.apply(z=z, y=y, is_training=is_training)

  def batch_norm(self, inputs, **kwargs):
    if self._batch_norm_fn is None:
      return inputs
    args = kwargs.copy()
    args["inputs"] = inputs
    if "use_sn" not in args:
      args["use_sn"] = self._spectral_norm
    return utils.call_with_accepted_args(self._batch_norm_fn, **args)

  def apply(self, z, y, is_training):
    raise NotImplementedError


@gin.configurable("D", blacklist=["name"])
class AbstractDiscriminator(_Module):
  """Interface for discriminator architectures."""

  def __init__(self, name="discriminator", batch_norm_fn=None, layer_norm=False, spectral_norm=False):
    super(AbstractDiscriminator, self).__init__(name=name)
    self._batch_norm_fn = batch_norm_fn
    self._layer_norm = layer_norm
    self._spectral_norm = spectral_norm
    if layer_norm:
      raise NotImplementedError("layer_norm is outside the accelerated hot path (SURVEY.md §2.1)")

  def __call__(self, x, y, is_training, reuse=None):
    with V.variable_scope(self.name):
      return self.apply(x=x, y=y, is_training=is_training)

  def batch_norm(self, inputs, **kwargs):
    if self._batch_norm_fn is None:
      return inputs
    args = kwargs.copy()
    args["inputs"] = inputs
    if "use_sn" not in args:
      args["use_sn"] = self._spectral_norm
    return utils.call_with_accepted_args(self._batch_norm_fn, **args)

  def apply(self, x, y, is_training):
    raise NotImplementedError
