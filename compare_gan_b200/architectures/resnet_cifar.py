"""ResNet generator/discriminator for 32x32 (reference architectures/resnet_cifar.py:34-167)."""
from .. import gin_lite as gin
from .. import kernels as K
from . import arch_ops as ops
from . import resnet_ops


@gin.configurable
class Generator(resnet_ops.ResNetGenerator):
  """ResNet generator, 3 up-blocks, 32x32 (reference resnet_cifar.py:34-112)."""

  def __init__(self, hierarchical_z=False, embed_z=False, embed_y=False, **kwargs):
    super(Generator, self).__init__(**kwargs)
    self._hierarchical_z = hierarchical_z
    self._embed_z = embed_z
    self._embed_y = embed_y

  def apply(self, z, y, is_training):
    assert self._image_shape[0] == 32
    assert self._image_shape[1] == 32
    num_blocks = 3
    z_dim = z.shape[1]
    if self._embed_z:
      z = ops.linear(z, z_dim, scope="embed_z", use_sn=self._spectral_norm)
    if self._embed_y:
      y = ops.linear(y, z_dim, scope="embed_y", use_sn=self._spectral_norm)
    y_per_block = num_blocks * [y]
    if self._hierarchical_z:
      chunk = z_dim // (num_blocks + 1)
      zs = [K.slice_cols(z, i * chunk, (i + 1) * chunk) for i in range(num_blocks + 1)]
      z0, z_per_block = zs[0], zs[1:]
      if y is not None:
        y_per_block = [K.concat_cols(zi, y) for zi in z_per_block]
    else:
      z0 = z
      z_per_block = num_blocks * [z]
    output = ops.linear(z0, 4 * 4 * 256, scope="fc_noise", use_sn=self._spectral_norm)
    output = K.reshape(output, -1, 4, 4, 256)
    for block_idx in range(3):
      block = self._resnet_block(name="B{}".format(block_idx + 1), in_channels=256, out_channels=256, scale="up")
      output = block(output, z=z_per_block[block_idx], y=y_per_block[block_idx], is_training=is_training)
    output = self.batch_norm(output, z=z, y=y, is_training=is_training, name="final_norm")
    output = K.relu(output)
    output = ops.conv2d(output, output_dim=self._image_shape[2], k_h=3, k_w=3, d_h=1, d_w=1, name="final_conv",
                        use_sn=self._spectral_norm)
    return K.sigmoid(output)


@gin.configurable
class Discriminator(resnet_ops.ResNetDiscriminator):
  """ResNet discriminator, 4 blocks, 32x32 (reference resnet_cifar.py:115-167)."""

  def __init__(self, project_y=False, **kwargs):
    super(Discriminator, self).__init__(**kwargs)
    self._project_y = project_y

  def apply(self, x, y, is_training):
    resnet_ops.validate_image_inputs(x)
    colors = x.shape[3]
    if colors not in [1, 3]:
      raise ValueError("Number of color channels not supported: {}".format(colors))
    output = x
    for block_idx in range(4):
      block = self._resnet_block(name="B{}".format(block_idx + 1), in_channels=colors if block_idx == 0 else 128,
                                 out_channels=128, scale="down" if block_idx <= 1 else "none")
      output = block(output, z=None, y=y, is_training=is_training)
    output = K.relu(output)
    h = K.globalpool(output, mean=True)
    out_logit = ops.linear(h, 1, scope="disc_final_fc", use_sn=self._spectral_norm)
    if self._project_y:
      if y is None:
        raise ValueError("You must provide class information y to project.")
      embedded_y = ops.linear(y, 128, use_bias=False, scope="embedding_fc", use_sn=self._spectral_norm)
      out_logit = K.add(out_logit, K.rowdot(embedded_y, h))
    out = K.sigmoid(out_logit)
    return out, out_logit, h
