"""SNDCGAN pair (reference architectures/sndcgan.py:36-127; Miyato et al. 2018, table 4).  Generator: dense seed at 1/8
of the image side, three 4x4 stride-2 transposed convolutions (256, 128, 64 channels) and a 3x3 one to the colours, BN +
ReLU in between, (tanh + 1) / 2.  Discriminator: inputs mapped to [-1, 1], seven convolutions alternating 3x3 stride 1
and 4x4 stride 2 with leaky ReLU(0.1), dense logit."""
from .. import kernels as K
from . import abstract_arch
from . import netdef

G_STAGES = ((256, 4, 2), (128, 4, 2), (64, 4, 2))                  # (channels, kernel, stride) of g_dc2..g_dc4
D_STAGES = ((64, 3, 1), (128, 4, 2), (128, 3, 1), (256, 4, 2), (256, 3, 1), (512, 4, 2), (512, 3, 1))


def conv_out_size_same(size, stride):
  return -(-size // stride)


class Generator(abstract_arch.AbstractGenerator):

  def apply(self, z, y, is_training):
    height, width, colors = self._image_shape
    pyramid = netdef.halvings(height, width, len(G_STAGES))       # [full, 1/2, 1/4, 1/8]
    seed_h, seed_w = pyramid[-1]
    flow = netdef.Flow(self, z, z=z, y=y, is_training=is_training)
    flow.linear(seed_h * seed_w * 512, "g_fc1").norm_relu("g_bn1", tf32=True).reshape(z.shape[0], seed_h, seed_w, 512)
    for i, (channels, kernel, stride) in enumerate(G_STAGES):
      flow.deconv(pyramid[len(G_STAGES) - 1 - i], channels, kernel, stride, "g_dc%d" % (i + 2))
      flow.norm_relu("g_bn%d" % (i + 2), tf32=True)
    flow.deconv(pyramid[0], colors, 3, 1, "g_dc5")
    return K.tanh01(flow.x)


class Discriminator(abstract_arch.AbstractDiscriminator):

  def apply(self, x, y, is_training):
    del y, is_training
    sn = self._spectral_norm
    flow = netdef.Flow(self, K.affine(x, 2.0, -1.0))               # [0, 1] -> [-1, 1]
    for i, (channels, kernel, stride) in enumerate(D_STAGES):
      flow.conv(channels, kernel, stride, "d_conv%d" % (i + 1), use_sn=sn).lrelu(leak=0.1, _tf32=i + 1 < len(D_STAGES))
    features = flow.reshape(x.shape[0], -1).x
    logit = flow.linear(1, "d_fc1", use_sn=sn).x
    return K.sigmoid(logit), logit, features
