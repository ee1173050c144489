"""Five-block ResNet pair for up to 128x128 (reference architectures/resnet5.py:36-145; the WGAN-GP ResNet of Gulrajani
et al. 2017) as plans for `resnet_family`.  `channels` are width multipliers of `ch` along the network; the generator
up-samples in as many leading blocks as the image has octaves above its 4x4 seed, the discriminator down-samples in all
six of its blocks (B0 maps the colours to `ch`).  Neither network applies spectral norm outside its blocks' own flag."""
import math

from . import resnet_family as family

BLOCKS = 5


class Generator(family.PlainResNetGenerator):

  def __init__(self, ch=64, channels=(8, 8, 4, 4, 2, 1), **kwargs):
    super(Generator, self).__init__(**kwargs)
    self._ch, self._channels = ch, channels

  def _plan(self):
    side = self._image_shape[0]
    octaves = math.log2(float(side) / family.SEED)
    if not float(octaves).is_integer():
      raise ValueError("log2({}/{}) must be an integer.".format(side, family.SEED))
    if not 0 <= octaves <= BLOCKS:
      raise ValueError("Invalid image_size {}.".format(side))
    return family.GeneratorPlan(widths=[self._ch * m for m in self._channels],
                                scales=["up" if i < int(octaves) else "none" for i in range(BLOCKS)],
                                hierarchical_z=False, embed_z=False, embed_y=False, spectral_norm_outside_blocks=False)


class Discriminator(family.PlainResNetDiscriminator):

  def __init__(self, ch=64, channels=(1, 2, 4, 4, 8, 8), **kwargs):
    super(Discriminator, self).__init__(**kwargs)
    self._ch, self._channels = ch, channels

  def _plan(self, colors):
    return family.DiscriminatorPlan(first_block=0, widths=[self._ch * m for m in self._channels],
                                    scales=(BLOCKS + 1) * ["down"], project_y=False)
