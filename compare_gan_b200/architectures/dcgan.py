"""DCGAN generator/discriminator (reference architectures/dcgan.py:39-129): 5x5 stride-2 (transposed) convolutions with
TF's asymmetric SAME padding (pad_before = 1, pad_after = 2 on even sizes)."""
from .. import kernels as K
from . import abstract_arch
from . import arch_ops as ops
from .arch_ops import conv2d, deconv2d, linear, lrelu
from .sndcgan import conv_out_size_same


class Generator(abstract_arch.AbstractGenerator):
  """reference dcgan.py:39-84."""

  def apply(self, z, y, is_training):
    gf_dim = 64
    bs = z.shape[0]
    s_h, s_w, colors = self._image_shape
    s_h2, s_w2 = conv_out_size_same(s_h, 2), conv_out_size_same(s_w, 2)
    s_h4, s_w4 = conv_out_size_same(s_h2, 2), conv_out_size_same(s_w2, 2)
    s_h8, s_w8 = conv_out_size_same(s_h4, 2), conv_out_size_same(s_w4, 2)
    s_h16, s_w16 = conv_out_size_same(s_h8, 2), conv_out_size_same(s_w8, 2)
    net = linear(z, gf_dim * 8 * s_h16 * s_w16, scope="g_fc1")
    net = K.reshape(net, -1, s_h16, s_w16, gf_dim * 8)
    net = ops.norm_relu(self.batch_norm, net, z=z, y=y, is_training=is_training, name="g_bn1")
    net = deconv2d(net, [bs, s_h8, s_w8, gf_dim * 4], 5, 5, 2, 2, name="g_dc1")
    net = ops.norm_relu(self.batch_norm, net, z=z, y=y, is_training=is_training, name="g_bn2")
    net = deconv2d(net, [bs, s_h4, s_w4, gf_dim * 2], 5, 5, 2, 2, name="g_dc2")
    net = ops.norm_relu(self.batch_norm, net, z=z, y=y, is_training=is_training, name="g_bn3")
    net = deconv2d(net, [bs, s_h2, s_w2, gf_dim * 1], 5, 5, 2, 2, name="g_dc3")
    net = ops.norm_relu(self.batch_norm, net, z=z, y=y, is_training=is_training, name="g_bn4")
    net = deconv2d(net, [bs, s_h, s_w, colors], 5, 5, 2, 2, name="g_dc4")
    return K.tanh01(net)            # 0.5 * tf.nn.tanh(net) + 0.5


class Discriminator(abstract_arch.AbstractDiscriminator):
  """reference dcgan.py:87-129."""

  def apply(self, x, y, is_training):
    bs = x.shape[0]
    df_dim = 64
    net = lrelu(conv2d(x, df_dim, 5, 5, 2, 2, name="d_conv1", use_sn=self._spectral_norm))
    net = conv2d(net, df_dim * 2, 5, 5, 2, 2, name="d_conv2", use_sn=self._spectral_norm)
    net = self.batch_norm(net, y=y, is_training=is_training, name="d_bn1")
    net = lrelu(net)
    net = conv2d(net, df_dim * 4, 5, 5, 2, 2, name="d_conv3", use_sn=self._spectral_norm)
    net = self.batch_norm(net, y=y, is_training=is_training, name="d_bn2")
    net = lrelu(net)
    net = conv2d(net, df_dim * 8, 5, 5, 2, 2, name="d_conv4", use_sn=self._spectral_norm)
    net = self.batch_norm(net, y=y, is_training=is_training, name="d_bn3")
    net = lrelu(net)
    out_logit = linear(K.reshape(net, bs, -1), 1, scope="d_fc4", use_sn=self._spectral_norm)
    out = K.sigmoid(out_logit)
    return out, out_logit, net
