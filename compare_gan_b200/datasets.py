"""Dataset surface (reference datasets.py:66-329, 620-648), synthetic only: BASELINE measures on synthetic
data and the TFDS pipeline is out of scope.  Keeps the ImageDatasetV2 property surface and
the fake-data generator semantics of datasets.py:136-145 (uniform [0,1) images, seed 547)."""
import numpy as np

from . import gin_lite as gin

# name -> (resolution, colors, num_classes, eval_test_samples)   (datasets.py:370-512, 620-640)
DATASETS = {
    "cifar10": (32, 3, 10, 10000),
    "celeb_a": (64, 3, None, 10000),
    "celeb_a_hq_128": (128, 3, None, 10000),      # named by sndcgan_celebahq128.gin:4 (SURVEY App. C note)
    "lsun-bedroom": (128, 3, None, 30000),
    "imagenet_128": (128, 3, 1000, 50000),
}


class ImageDatasetV2(object):
  """Synthetic stand-in exposing name / image_shape / num_classes / eval_test_samples."""

  def __init__(self, name, resolution, colors, num_classes, eval_test_samples, seed=547):
    self._name, self._resolution, self._colors = name, resolution, colors
    self._num_classes, self._eval_test_samples, self._seed = num_classes, eval_test_samples, seed
    self._rng = np.random.RandomState(seed)

  @property
  def name(self):
    return self._name

  @property
  def num_classes(self):
    return self._num_classes

  @property
  def eval_test_samples(self):
    return self._eval_test_samples

  @property
  def image_shape(self):
    return (self._resolution, self._resolution, self._colors)

  def sample_images(self, n):
    """float32 U[0,1) NHWC (datasets.py:136-145)."""
    return self._rng.rand(n, self._resolution, self._resolution, self._colors).astype(np.float32)

  def sample_labels(self, n):
    if not self._num_classes:
      return None
    return self._rng.randint(0, self._num_classes, size=n).astype(np.int32)


@gin.configurable("dataset")
def get_dataset(name, seed=547):
  """Instantiates a data set and sets the random seed (reference datasets.py:643-648)."""
  if name not in DATASETS:
    raise ValueError("Dataset %s is not available." % name)
  res, colors, classes, n_eval = DATASETS[name]
  return ImageDatasetV2(name, res, colors, classes, n_eval, seed=seed)
