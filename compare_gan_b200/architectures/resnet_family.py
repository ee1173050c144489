"""The two "plain" ResNet pairs (CIFAR 32x32 and the five-block 128x128 one) share everything but their tables: a dense
seed, residual blocks, BN-ReLU-conv3x3-sigmoid on the generator side; residual blocks, ReLU, spatial mean, dense logit
(+ optional class projection) on the discriminator side.  The concrete modules (`resnet_cifar`, `resnet5`) only provide
a `GeneratorPlan` / `DiscriminatorPlan`; this module runs them."""
import collections

from .. import kernels as K
from . import arch_ops as ops
from . import netdef
from . import resnet_ops

SEED = 4

GeneratorPlan = collections.namedtuple(
    "GeneratorPlan", "widths scales hierarchical_z embed_z embed_y spectral_norm_outside_blocks")
# widths: seed width followed by each block's output width; scales: one of "up"/"none" per block;
# spectral_norm_outside_blocks: whether the dense seed, the embeddings and the final conv follow G.spectral_norm

DiscriminatorPlan = collections.namedtuple("DiscriminatorPlan", "first_block widths scales project_y")
# first_block: index in the block names ("B0" or "B1"); widths: output width per block; scales: "down"/"none" per block


class PlainResNetGenerator(resnet_ops.ResNetGenerator):

  def _plan(self):
    raise NotImplementedError

  def apply(self, z, y, is_training):
    plan = self._plan()
    sn = self._spectral_norm and plan.spectral_norm_outside_blocks
    if plan.embed_z:
      z = ops.linear(z, z.shape[1], scope="embed_z", use_sn=sn)
    if plan.embed_y:
      y = ops.linear(y, z.shape[1], scope="embed_y", use_sn=sn)
    z_seed, z_blocks, y_blocks = netdef.split_latent(z, y, len(plan.scales), plan.hierarchical_z)
    flow = netdef.Flow(self, z_seed, z=z, y=y, is_training=is_training)
    flow.linear(SEED * SEED * plan.widths[0], "fc_noise", use_sn=sn).reshape(-1, SEED, SEED, plan.widths[0])
    for i, scale in enumerate(plan.scales):
      block = self._resnet_block("B%d" % (i + 1), plan.widths[i], plan.widths[i + 1], scale)
      flow.x = block(flow.x, z=z_blocks[i], y=y_blocks[i], is_training=is_training)
    flow.norm_relu("final_norm", tf32=True).conv(self._image_shape[2], 3, 1, "final_conv", use_sn=sn)
    return K.sigmoid(flow.x)


class PlainResNetDiscriminator(resnet_ops.ResNetDiscriminator):

  def _plan(self, colors):
    raise NotImplementedError

  def apply(self, x, y, is_training):
    resnet_ops.validate_image_inputs(x)
    colors = x.shape[3]
    if colors not in (1, 3):
      raise ValueError("Number of color channels not supported: {}".format(colors))
    plan = self._plan(colors)
    net, width = x, colors
    for i, (out_width, scale) in enumerate(zip(plan.widths, plan.scales)):
      block = self._resnet_block("B%d" % (plan.first_block + i), width, out_width, scale)
      net, width = block(net, z=None, y=y, is_training=is_training), out_width
    features = K.globalpool(K.relu(net), mean=True)
    logit = ops.linear(features, 1, scope="disc_final_fc", use_sn=self._spectral_norm)
    if plan.project_y:
      if y is None:
        raise ValueError("You must provide class information y to project.")
      embedded = ops.linear(y, width, use_bias=False, scope="embedding_fc", use_sn=self._spectral_norm)
      logit = K.add(logit, netdef.projection_term(embedded, features))
    return K.sigmoid(logit), logit, features
