"""BigGAN generator/discriminator (reference architectures/resnet_biggan.py:80-425)."""
from .. import gin_lite as gin
from .. import kernels as K
from .. import variables as V
from . import abstract_arch
from . import arch_ops as ops
from . import resnet_ops


@gin.configurable
class BigGanResNetBlock(resnet_ops.ResNetBlock):
  """ResNet block with a 1x1 convolution for the (optional) shortcut (reference resnet_biggan.py:80-151)."""

  def __init__(self, add_shortcut=True, **kwargs):
    super(BigGanResNetBlock, self).__init__(**kwargs)
    self._add_shortcut = add_shortcut

  def apply(self, inputs, z, y, is_training):
    if inputs.shape[-1] != self._in_channels:
      raise ValueError("Unexpected number of input channels (expected {}, got {}).".format(
          self._in_channels, inputs.shape[-1]))
    with V.variable_scope(self._name):
      outputs = inputs
      outputs = ops.norm_relu(self.batch_norm, outputs, z=z, y=y, is_training=is_training, name="bn1")
      outputs = self._get_conv(outputs, self._in_channels, self._out_channels, self._scale1, suffix="conv1")
      outputs = ops.norm_relu(self.batch_norm, outputs, z=z, y=y, is_training=is_training, name="bn2")
      outputs = self._get_conv(outputs, self._out_channels, self._out_channels, self._scale2, suffix="conv2")
      if self._add_shortcut:
        shortcut = self._get_conv(inputs, self._in_channels, self._out_channels, self._scale, kernel_size=(1, 1),
                                  suffix="conv_shortcut")
        outputs = K.add(outputs, shortcut)
      return outputs


@gin.configurable
class Generator(abstract_arch.AbstractGenerator):
  """ResNet-based generator for 32..512 (reference resnet_biggan.py:154-302)."""

  def __init__(self, ch=96, blocks_with_attention="B4", hierarchical_z=True, embed_z=False, embed_y=True,
               embed_y_dim=128, embed_bias=False, **kwargs):
    super(Generator, self).__init__(**kwargs)
    self._ch = ch
    self._blocks_with_attention = set(blocks_with_attention.split(","))
    self._hierarchical_z = hierarchical_z
    self._embed_z = embed_z
    self._embed_y = embed_y
    self._embed_y_dim = embed_y_dim
    self._embed_bias = embed_bias

  def _resnet_block(self, name, in_channels, out_channels, scale):
    if scale not in ["up", "none"]:
      raise ValueError("Unknown generator ResNet block scaling: {}.".format(scale))
    return BigGanResNetBlock(name=name, in_channels=in_channels, out_channels=out_channels, scale=scale,
                             is_gen_block=True, spectral_norm=self._spectral_norm, batch_norm=self.batch_norm)

  def _get_in_out_channels(self):
    resolution = self._image_shape[0]
    table = {512: [16, 16, 8, 8, 4, 2, 1, 1], 256: [16, 16, 8, 8, 4, 2, 1], 128: [16, 16, 8, 4, 2, 1],
             64: [16, 16, 8, 4, 2], 32: [4, 4, 4, 4]}
    if resolution not in table:
      raise ValueError("Unsupported resolution: {}".format(resolution))
    m = table[resolution]
    return [self._ch * c for c in m[:-1]], [self._ch * c for c in m[1:]]

  def apply(self, z, y, is_training):
    seed_size = 4
    z_dim = z.shape[1]
    in_channels, out_channels = self._get_in_out_channels()
    num_blocks = len(in_channels)
    if self._embed_z:
      z = ops.linear(z, z_dim, scope="embed_z", use_sn=False, use_bias=self._embed_bias)
    if self._embed_y:
      y = ops.linear(y, self._embed_y_dim, scope="embed_y", use_sn=False, use_bias=self._embed_bias)
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
    net = ops.linear(z0, in_channels[0] * seed_size * seed_size, scope="fc_noise", use_sn=self._spectral_norm)
    net = K.reshape(net, -1, seed_size, seed_size, in_channels[0])
    for block_idx in range(num_blocks):
      name = "B{}".format(block_idx + 1)
      block = self._resnet_block(name=name, in_channels=in_channels[block_idx],
                                 out_channels=out_channels[block_idx], scale="up")
      net = block(net, z=z_per_block[block_idx], y=y_per_block[block_idx], is_training=is_training)
      if name in self._blocks_with_attention:
        net = ops.non_local_block(net, "non_local_block", use_sn=self._spectral_norm)
    net = ops.batch_norm(net, is_training=is_training, name="final_norm")
    net = K.relu(net)
    net = ops.conv2d(net, output_dim=self._image_shape[2], k_h=3, k_w=3, d_h=1, d_w=1, name="final_conv",
                     use_sn=self._spectral_norm)
    return K.tanh01(net)           # (tf.nn.tanh(net) + 1.0) / 2.0


@gin.configurable
class Discriminator(abstract_arch.AbstractDiscriminator):
  """ResNet-based discriminator for 32..512 (reference resnet_biggan.py:305-425)."""

  def __init__(self, ch=96, blocks_with_attention="B1", project_y=True, **kwargs):
    super(Discriminator, self).__init__(**kwargs)
    self._ch = ch
    self._blocks_with_attention = set(blocks_with_attention.split(","))
    self._project_y = project_y

  def _resnet_block(self, name, in_channels, out_channels, scale):
    if scale not in ["down", "none"]:
      raise ValueError("Unknown discriminator ResNet block scaling: {}.".format(scale))
    return BigGanResNetBlock(name=name, in_channels=in_channels, out_channels=out_channels, scale=scale,
                             is_gen_block=False, add_shortcut=in_channels != out_channels,
                             layer_norm=self._layer_norm, spectral_norm=self._spectral_norm,
                             batch_norm=self.batch_norm)

  def _get_in_out_channels(self, colors, resolution):
    if colors not in [1, 3]:
      raise ValueError("Unsupported color channels: {}".format(colors))
    table = {512: [1, 1, 2, 4, 8, 8, 16, 16], 256: [1, 2, 4, 8, 8, 16, 16], 128: [1, 2, 4, 8, 16, 16],
             64: [2, 4, 8, 16, 16], 32: [2, 2, 2, 2]}
    if resolution not in table:
      raise ValueError("Unsupported resolution: {}".format(resolution))
    out_channels = [self._ch * c for c in table[resolution]]
    return [colors] + out_channels[:-1], out_channels

  def apply(self, x, y, is_training):
    resnet_ops.validate_image_inputs(x)
    in_channels, out_channels = self._get_in_out_channels(colors=x.shape[-1], resolution=x.shape[1])
    num_blocks = len(in_channels)
    net = x
    for block_idx in range(num_blocks):
      name = "B{}".format(block_idx + 1)
      is_last_block = block_idx == num_blocks - 1
      block = self._resnet_block(name=name, in_channels=in_channels[block_idx],
                                 out_channels=out_channels[block_idx], scale="none" if is_last_block else "down")
      net = block(net, z=None, y=y, is_training=is_training)
      if name in self._blocks_with_attention:
        net = ops.non_local_block(net, "non_local_block", use_sn=self._spectral_norm)
    net = K.relu(net)
    h = K.globalpool(net, mean=False)
    out_logit = ops.linear(h, 1, scope="final_fc", use_sn=self._spectral_norm)
    if self._project_y:
      if y is None:
        raise ValueError("You must provide class information y to project.")
      with V.variable_scope("embedding_fc"):
        y_embedding_dim = out_channels[-1]
        kernel = V.get_variable("kernel", (y.shape[1], y_embedding_dim), ops.glorot_normal)
        if self._spectral_norm:
          kernel = ops.spectral_norm(kernel)
        embedded_y = K.matmul(y, kernel)
        out_logit = K.add(out_logit, K.rowdot(embedded_y, h))
    out = K.sigmoid(out_logit)
    return out, out_logit, h
