"""32x32 ResNet pair (reference architectures/resnet_cifar.py:34-167; Miyato et al. 2018, table 3) as plans for
`resnet_family`: generator = 4x4x256 seed + three 256-wide up-sampling blocks (optionally with hierarchical z and
embedded z / y); discriminator = four 128-wide blocks, the first two down-sampling, optional class projection."""
from .. import gin_lite as gin
from . import resnet_family as family


@gin.configurable
class Generator(family.PlainResNetGenerator):

  def __init__(self, hierarchical_z=False, embed_z=False, embed_y=False, **kwargs):
    super(Generator, self).__init__(**kwargs)
    self._latent_options = (hierarchical_z, embed_z, embed_y)

  def _plan(self):
    assert tuple(self._image_shape[:2]) == (32, 32)
    hierarchical_z, embed_z, embed_y = self._latent_options
    return family.GeneratorPlan(widths=4 * [256], scales=3 * ["up"], hierarchical_z=hierarchical_z, embed_z=embed_z,
                                embed_y=embed_y, spectral_norm_outside_blocks=True)


@gin.configurable
class Discriminator(family.PlainResNetDiscriminator):

  def __init__(self, project_y=False, **kwargs):
    super(Discriminator, self).__init__(**kwargs)
    self._project_y = project_y

  def _plan(self, colors):
    return family.DiscriminatorPlan(first_block=1, widths=4 * [128], scales=["down", "down", "none", "none"],
                                    project_y=self._project_y)
