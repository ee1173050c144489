"""ResNet building blocks (reference architectures/resnet_ops.py:35-219)."""
import math

from .. import kernels as K
from .. import variables as V
from . import abstract_arch
from . import arch_ops as ops


def validate_image_inputs(inputs, validate_power2=True):
  """reference resnet_ops.py:59-67."""
  if len(inputs.shape) != 4:
    raise ValueError("Input tensor must have rank 4.")
  if inputs.shape[1] != inputs.shape[2]:
    raise ValueError("Input tensor does not have equal width and height: ", inputs.shape[1:3])
  width = inputs.shape[1]
  if validate_power2 and math.log(width, 2) != int(math.log(width, 2)):
    raise ValueError("Input tensor `width` is not a power of 2: ", width)


class ResNetBlock(object):
  """ResNet block with options for various normalizations (reference resnet_ops.py:70-182)."""

  def __init__(self, name, in_channels, out_channels, scale, is_gen_block, layer_norm=False, spectral_norm=False,
               batch_norm=None):
    assert scale in ["up", "down", "none"]
    self._name = name
    self._in_channels = in_channels
    self._out_channels = out_channels
    self._scale = scale
    self._scale1 = scale if is_gen_block else "none"
    self._scale2 = "none" if is_gen_block else scale
    self._layer_norm = layer_norm
    self._spectral_norm = spectral_norm
    self.batch_norm = batch_norm

  def __call__(self, inputs, z, y, is_training):
    return self.apply(inputs=inputs, z=z, y=y, is_training=is_training)

  def _get_conv(self, inputs, in_channels, out_channels, scale, suffix, kernel_size=(3, 3), strides=(1, 1)):
    """A convolution of the block; "up" fuses the zero-insertion unpool into the conv kernel's gather
    (reference resnet_ops.py:104-134 materialises unpool(inputs) first)."""
    if inputs.shape[-1] != in_channels:
      raise ValueError("Unexpected number of input channels.")
    if scale not in ["up", "down", "none"]:
      raise ValueError("Scale: got {}, expected 'up', 'down', or 'none'.".format(scale))
    outputs = ops.conv2d(inputs, output_dim=out_channels, k_h=kernel_size[0], k_w=kernel_size[1],
                         d_h=strides[0], d_w=strides[1], use_sn=self._spectral_norm,
                         name="{}_{}".format("same" if scale == "none" else scale, suffix),
                         _upsample=(scale == "up"))
    if scale == "down":
      outputs = K.avgpool2(outputs)
    return outputs

  def apply(self, inputs, z, y, is_training):
    if inputs.shape[-1] != self._in_channels:
      raise ValueError("Unexpected number of input channels.")
    with V.variable_scope(self._name):
      output = inputs
      shortcut = self._get_conv(output, self._in_channels, self._out_channels, self._scale,
                                suffix="conv_shortcut")
      output = ops.norm_relu(self.batch_norm, output, z=z, y=y, is_training=is_training, name="bn1")
      output = self._get_conv(output, self._in_channels, self._out_channels, self._scale1, suffix="conv1")
      output = ops.norm_relu(self.batch_norm, output, z=z, y=y, is_training=is_training, name="bn2")
      output = self._get_conv(output, self._out_channels, self._out_channels, self._scale2, suffix="conv2")
      return K.add(output, shortcut)


class ResNetGenerator(abstract_arch.AbstractGenerator):
  """reference resnet_ops.py:185-200."""

  def _resnet_block(self, name, in_channels, out_channels, scale):
    if scale not in ["up", "none"]:
      raise ValueError("Unknown generator ResNet block scaling: {}.".format(scale))
    return ResNetBlock(name=name, in_channels=in_channels, out_channels=out_channels, scale=scale,
                       is_gen_block=True, spectral_norm=self._spectral_norm, batch_norm=self.batch_norm)


class ResNetDiscriminator(abstract_arch.AbstractDiscriminator):
  """reference resnet_ops.py:203-219."""

  def _resnet_block(self, name, in_channels, out_channels, scale):
    if scale not in ["down", "none"]:
      raise ValueError("Unknown discriminator ResNet block scaling: {}.".format(scale))
    return ResNetBlock(name=name, in_channels=in_channels, out_channels=out_channels, scale=scale,
                       is_gen_block=False, layer_norm=self._layer_norm, spectral_norm=self._spectral_norm,
                       batch_norm=self.batch_norm)
