"""SNDCGAN generator/discriminator (reference architectures/sndcgan.py:36-127)."""
from .. import kernels as K
from . import abstract_arch
from .arch_ops import conv2d, deconv2d, linear, lrelu


def conv_out_size_same(size, stride):
  return -(-size // stride)


class Generator(abstract_arch.AbstractGenerator):
  """reference sndcgan.py:36-79."""

  def apply(self, z, y, is_training):
    batch_size = z.shape[0]
    s_h, s_w, colors = self._image_shape
    s_h2, s_w2 = conv_out_size_same(s_h, 2), conv_out_size_same(s_w, 2)
    s_h4, s_w4 = conv_out_size_same(s_h2, 2), conv_out_size_same(s_w2, 2)
    s_h8, s_w8 = conv_out_size_same(s_h4, 2), conv_out_size_same(s_w4, 2)
    net = linear(z, s_h8 * s_w8 * 512, scope="g_fc1")
    net = self.batch_norm(net, z=z, y=y, is_training=is_training, name="g_bn1")
    net = K.relu(net)
    net = K.reshape(net, batch_size, s_h8, s_w8, 512)
    net = deconv2d(net, [batch_size, s_h4, s_w4, 256], 4, 4, 2, 2, name="g_dc2")
    net = self.batch_norm(net, z=z, y=y, is_training=is_training, name="g_bn2")
    net = K.relu(net)
    net = deconv2d(net, [batch_size, s_h2, s_w2, 128], 4, 4, 2, 2, name="g_dc3")
    net = self.batch_norm(net, z=z, y=y, is_training=is_training, name="g_bn3")
    net = K.relu(net)
    net = deconv2d(net, [batch_size, s_h, s_w, 64], 4, 4, 2, 2, name="g_dc4")
    net = self.batch_norm(net, z=z, y=y, is_training=is_training, name="g_bn4")
    net = K.relu(net)
    net = deconv2d(net, [batch_size, s_h, s_w, colors], 3, 3, 1, 1, name="g_dc5")
    return K.tanh01(net)          # tf.div(tf.tanh(net) + 1.0, 2.0)


class Discriminator(abstract_arch.AbstractDiscriminator):
  """reference sndcgan.py:82-127."""

  def apply(self, x, y, is_training):
    del is_training, y
    use_sn = self._spectral_norm
    net = K.affine(x, 2.0, -1.0)   # x * 2.0 - 1.0
    spec = [(64, 3, 1), (128, 4, 2), (128, 3, 1), (256, 4, 2), (256, 3, 1), (512, 4, 2), (512, 3, 1)]
    for i, (c, k, s) in enumerate(spec):
      net = conv2d(net, c, k, k, s, s, name="d_conv%d" % (i + 1), use_sn=use_sn)
      net = lrelu(net, leak=0.1)
    batch_size = x.shape[0]
    net = K.reshape(net, batch_size, -1)
    out_logit = linear(net, 1, scope="d_fc1", use_sn=use_sn)
    out = K.sigmoid(out_logit)
    return out, out_logit, net
