"""32x32 ResNet pair (reference architectures/resnet_cifar.py:34-167; Miyato et al. 2018, table 3): a 4x4x256 seed and
three up-sampling blocks in the generator, four 128-wide blocks (the first two down-sampling) in the discriminator,
optional class projection."""
from .. import gin_lite as gin
from .. import kernels as K
from . import arch_ops as ops
from . import netdef
from . import resnet_ops

G_BLOCKS, G_WIDTH = 3, 256
D_SCALES, D_WIDTH = ("down", "down", "none", "none"), 128


@gin.configurable
class Generator(resnet_ops.ResNetGenerator):

  def __init__(self, hierarchical_z=False, embed_z=False, embed_y=False, **kwargs):
    super(Generator, self).__init__(**kwargs)
    self._hierarchical_z, self._embed_z, self._embed_y = hierarchical_z, embed_z, embed_y

  def apply(self, z, y, is_training):
    assert tuple(self._image_shape[:2]) == (32, 32)
    sn = self._spectral_norm
    width = z.shape[1]
    if self._embed_z:
      z = ops.linear(z, width, scope="embed_z", use_sn=sn)
    if self._embed_y:
      y = ops.linear(y, width, scope="embed_y", use_sn=sn)
    z_seed, z_blocks, y_blocks = netdef.split_latent(z, y, G_BLOCKS, self._hierarchical_z)
    flow = netdef.Flow(self, z_seed, z=z, y=y, is_training=is_training)
    flow.linear(4 * 4 * G_WIDTH, "fc_noise", use_sn=sn).reshape(-1, 4, 4, G_WIDTH)
    for i in range(G_BLOCKS):
      block = self._resnet_block("B%d" % (i + 1), G_WIDTH, G_WIDTH, "up")
      flow.x = block(flow.x, z=z_blocks[i], y=y_blocks[i], is_training=is_training)
    flow.norm("final_norm").relu().conv(self._image_shape[2], 3, 1, "final_conv", use_sn=sn)
    return K.sigmoid(flow.x)


@gin.configurable
class Discriminator(resnet_ops.ResNetDiscriminator):

  def __init__(self, project_y=False, **kwargs):
    super(Discriminator, self).__init__(**kwargs)
    self._project_y = project_y

  def apply(self, x, y, is_training):
    resnet_ops.validate_image_inputs(x)
    colors = x.shape[3]
    if colors not in (1, 3):
      raise ValueError("Number of color channels not supported: {}".format(colors))
    net, cin = x, colors
    for i, scale in enumerate(D_SCALES):
      net = self._resnet_block("B%d" % (i + 1), cin, D_WIDTH, scale)(net, z=None, y=y, is_training=is_training)
      cin = D_WIDTH
    features = K.globalpool(K.relu(net), mean=True)
    logit = ops.linear(features, 1, scope="disc_final_fc", use_sn=self._spectral_norm)
    if self._project_y:
      if y is None:
        raise ValueError("You must provide class information y to project.")
      embedded = ops.linear(y, D_WIDTH, use_bias=False, scope="embedding_fc", use_sn=self._spectral_norm)
      logit = K.add(logit, netdef.projection_term(embedded, features))
    return K.sigmoid(logit), logit, features
