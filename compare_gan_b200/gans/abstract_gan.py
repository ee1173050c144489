"""Interface for GAN models (reference gans/abstract_gan.py:28-92)."""


class AbstractGAN(object):
  """dataset: object with .name/.image_shape/.num_classes; parameters: legacy options dict; model_dir: str."""

  def __init__(self, dataset, parameters, model_dir):
    self._dataset = dataset
    self._parameters = parameters
    self._model_dir = model_dir

  @property
  def dataset(self):
    return self._dataset

  @property
  def parameters(self):
    return self._parameters

  @property
  def model_dir(self):
    return self._model_dir
