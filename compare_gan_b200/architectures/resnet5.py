"""Five-block ResNet pair for up to 128x128 (reference architectures/resnet5.py:36-145; the WGAN-GP ResNet of Gulrajani
et al. 2017): `channels` are width multipliers of `ch` along the network; the generator up-samples in as many leading
blocks as the image needs octaves above its 4x4 seed, the discriminator down-samples in all six of its blocks."""
import math

from .. import kernels as K
from . import netdef
from . import resnet_ops

SEED = 4
BLOCKS = 5


class Generator(resnet_ops.ResNetGenerator):

  def __init__(self, ch=64, channels=(8, 8, 4, 4, 2, 1), **kwargs):
    super(Generator, self).__init__(**kwargs)
    self._ch, self._channels = ch, channels

  def _octaves(self):
    side = self._image_shape[0]
    octaves = math.log2(float(side) / SEED)
    if not float(octaves).is_integer():
      raise ValueError("log2({}/{}) must be an integer.".format(side, SEED))
    if not 0 <= octaves <= BLOCKS:
      raise ValueError("Invalid image_size {}.".format(side))
    return int(octaves)

  def apply(self, z, y, is_training):
    widths = [self._ch * m for m in self._channels]
    flow = netdef.Flow(self, z, z=z, y=y, is_training=is_training)
    flow.linear(widths[0] * SEED * SEED, "fc_noise").reshape(-1, SEED, SEED, widths[0])
    octaves = self._octaves()
    for i in range(BLOCKS):
      block = self._resnet_block("B%d" % (i + 1), widths[i], widths[i + 1], "up" if i < octaves else "none")
      flow.x = block(flow.x, z=z, y=y, is_training=is_training)
    flow.norm("final_norm").relu().conv(self._image_shape[2], 3, 1, "final_conv")
    return K.sigmoid(flow.x)


class Discriminator(resnet_ops.ResNetDiscriminator):

  def __init__(self, ch=64, channels=(1, 2, 4, 4, 8, 8), **kwargs):
    super(Discriminator, self).__init__(**kwargs)
    self._ch, self._channels = ch, channels

  def apply(self, x, y, is_training):
    resnet_ops.validate_image_inputs(x)
    colors = x.shape[3]
    if colors not in (1, 3):
      raise ValueError("Number of color channels not supported: {}".format(colors))
    widths = [colors] + [self._ch * m for m in self._channels]        # B0: colors -> ch, then the multipliers
    net = x
    for i in range(BLOCKS + 1):
      net = self._resnet_block("B%d" % i, widths[i], widths[i + 1], "down")(net, z=None, y=y, is_training=is_training)
    features = K.globalpool(K.relu(net), mean=True)
    logit = netdef.Flow(self, features).linear(1, "disc_final_fc", use_sn=self._spectral_norm).x
    return K.sigmoid(logit), logit, features
