"""DCGAN pair (reference architectures/dcgan.py:39-129; Radford et al. 2015): 5x5 stride-2 (transposed) convolutions,
which on even sizes exercise TF's asymmetric SAME padding (one pixel before, two after).  Generator: dense seed at 1/16
of the image side, widths 512-256-128-64-colours with BN + ReLU, (tanh + 1) / 2.  Discriminator: widths 64-128-256-512
with leaky ReLU(0.2), BN from the second layer on, dense logit."""
from .. import kernels as K
from . import abstract_arch
from . import netdef

BASE = 64
KERNEL, STRIDE = 5, 2


class Generator(abstract_arch.AbstractGenerator):

  def apply(self, z, y, is_training):
    height, width, colors = self._image_shape
    pyramid = netdef.halvings(height, width, 4)                      # [full, 1/2, 1/4, 1/8, 1/16]
    seed_h, seed_w = pyramid[-1]
    flow = netdef.Flow(self, z, z=z, y=y, is_training=is_training)
    flow.linear(BASE * 8 * seed_h * seed_w, "g_fc1").reshape(-1, seed_h, seed_w, BASE * 8).norm_relu("g_bn1", tf32=True)
    for i, channels in enumerate((BASE * 4, BASE * 2, BASE)):
      flow.deconv(pyramid[3 - i], channels, KERNEL, STRIDE, "g_dc%d" % (i + 1)).norm_relu("g_bn%d" % (i + 2), tf32=True)
    flow.deconv(pyramid[0], colors, KERNEL, STRIDE, "g_dc4")
    return K.tanh01(flow.x)


class Discriminator(abstract_arch.AbstractDiscriminator):

  def apply(self, x, y, is_training):
    sn = self._spectral_norm
    flow = netdef.Flow(self, x, y=y, is_training=is_training)
    flow.conv(BASE, KERNEL, STRIDE, "d_conv1", use_sn=sn).lrelu(_tf32=True)
    for i, channels in enumerate((BASE * 2, BASE * 4, BASE * 8)):
      flow.conv(channels, KERNEL, STRIDE, "d_conv%d" % (i + 2), use_sn=sn).norm("d_bn%d" % (i + 1)).lrelu(_tf32=i < 2)
    features = flow.x
    logit = flow.reshape(x.shape[0], -1).linear(1, "d_fc4", use_sn=sn).x
    return K.sigmoid(logit), logit, features
