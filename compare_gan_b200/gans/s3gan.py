"""S3GAN (reference gans/s3gan.py:39-321; Lucic et al. 2019, "High-Fidelity Image Generation With Fewer Labels"): ModularGAN
plus auxiliary heads on the discriminator's feature representation —

  (1) a rotation-prediction head trained on the last `batch / rotated_batch_fraction / 4` real and generated images of a
      sub-step, rotated by 90 / 180 / 270 degrees (self_supervision = "rotation"),
  (2) with project_y: a projection discriminator on the class embedding, whose label is the real one where available
      and otherwise the prediction of (3),
  (3) with use_predictor: a linear classifier on the features, trained with cross entropy on the labelled real examples.

Everything runs through the taped C-ABI ops of ModularGAN; `kernels.row_has_label`, `kernels.argmax_one_hot` and
`kernels.softmax_xent` (include/cgan_b200.h "S3GAN heads") are the additions.  As in the reference, create_loss has no
gradient-penalty term."""
from .. import gin_lite as gin
from .. import kernels as K
from .. import tape
from .. import variables as V
from ..architectures import arch_ops as ops
from . import loss_lib
from .modular_gan import ModularGAN
from .ssgan import NUM_ROTATIONS, _tile_rows, rotate_images


@gin.configurable(blacklist=["dataset", "parameters", "model_dir"])
class S3GAN(ModularGAN):
  """S3GAN which enables auxiliary heads for the modular GAN (reference gans/s3gan.py:40-101 constructor contract)."""

  def __init__(self, self_supervision="rotation", rotated_batch_fraction=gin.REQUIRED, weight_rotation_loss_d=1.0,
               weight_rotation_loss_g=0.2, project_y=False, use_predictor=False, use_soft_pred=False, weight_class_loss=1.0,
               use_soft_labels=False, **kwargs):
    super(S3GAN, self).__init__(**kwargs)
    if use_predictor and not project_y:
      raise ValueError("Using predictor requires projection.")
    if self_supervision not in ("none", "rotation"):
      raise ValueError("Unknown self_supervision {}.".format(self_supervision))
    if self_supervision == "rotation" and rotated_batch_fraction is gin.REQUIRED:
      raise ValueError("S3GAN.rotated_batch_fraction is required")
    if project_y and not self.conditional:
      raise ValueError("project_y needs a conditional GAN (labels)")
    self._self_supervision = self_supervision
    self._rotated_batch_fraction = rotated_batch_fraction
    self._weight_rotation_loss_d = weight_rotation_loss_d
    self._weight_rotation_loss_g = weight_rotation_loss_g
    self._project_y = project_y
    self._use_predictor = use_predictor
    self._use_soft_pred = use_soft_pred
    self._weight_class_loss = weight_class_loss
    self._use_soft_labels = use_soft_labels

  def get_class_embedding(self, y, embedding_dim, use_sn):
    """reference :160-173 (tf.initializers.glorot_normal kernel, optionally spectrally normalised)."""
    with V.variable_scope("discriminator_projection"):
      kernel = V.get_variable("kernel", (y.shape[1], embedding_dim), ops.glorot_normal)
      if use_sn:
        kernel = ops.spectral_norm(kernel)
      return K.matmul(y, kernel)

  def discriminator_with_additonal_heads(self, x, y, is_training):
    """reference :103-158 -> (probabilities, logits, rotation logits | None, class logits | None, is_label_available)."""
    d_probs, d_logits, x_rep = self.discriminator(x, y=y, is_training=is_training)
    use_sn = self.discriminator._spectral_norm
    if len(x_rep.shape) != 2:
      raise ValueError("S3GAN needs a rank-2 feature representation, got %s" % (x_rep.shape,))
    is_label_available = K.row_has_label(y) if y is not None else None
    rotation_logits = None
    if "rotation" in self._self_supervision:
      with V.variable_scope("discriminator_rotation"):
        rotation_logits = ops.linear(x_rep, NUM_ROTATIONS, scope="score_classify", use_sn=use_sn)
    if not self._project_y:
      return d_probs, d_logits, rotation_logits, None, is_label_available
    aux_logits = None
    if self._use_predictor:
      with V.variable_scope("discriminator_predictor"):
        aux_logits = ops.linear(x_rep, y.shape[1], use_bias=True, scope="predictor_linear", use_sn=use_sn)
      with tape.no_record():                                      # y = tf.stop_gradient(mix), :151-152
        y_predicted = K.softmax(tape.DT(aux_logits.t)) if self._use_soft_pred else K.argmax_one_hot(aux_logits)
        keep = K.rowscale(tape.DT(y.t), is_label_available)
        unlabelled = K.affine(is_label_available, -1.0, 1.0)     # 1 - is_label_available
        y = K.add(K.rowscale(y_predicted, unlabelled), keep)
    class_embedding = self.get_class_embedding(y=y, embedding_dim=x_rep.shape[-1], use_sn=use_sn)
    d_logits = K.add(d_logits, K.rowdot(class_embedding, x_rep))
    d_probs = K.sigmoid(d_logits)
    return d_probs, d_logits, rotation_logits, aux_logits, is_label_available

  def _build_networks(self, f):
    """Variable creation pass: the heads' variables must exist before the flat packing."""
    gen, all_y = super(S3GAN, self)._build_networks(f)
    self.discriminator_with_additonal_heads(K.concat_rows(f["images"], gen), y=all_y, is_training=True)
    return gen, all_y

  def merge_with_rotation_data(self, real, fake, real_labels, fake_labels, num_rot_examples):
    """reference :175-196: [real, real rotated by 90/180/270, fake, fake rotated]; labels tiled accordingly."""
    bs = real.shape[0]
    real_rotated = rotate_images(K.slice_rows(real, bs - num_rot_examples, bs), rot90_scalars=(1, 2, 3))
    fake_rotated = rotate_images(K.slice_rows(fake, bs - num_rot_examples, bs), rot90_scalars=(1, 2, 3))
    all_features = K.concat_rows(K.concat_rows(real, real_rotated), K.concat_rows(fake, fake_rotated))
    all_labels = None
    if self.conditional:
      real_rotated_labels = _tile_rows(K.slice_rows(real_labels, bs - num_rot_examples, bs), 3)
      fake_rotated_labels = _tile_rows(K.slice_rows(fake_labels, bs - num_rot_examples, bs), 3)
      all_labels = K.concat_rows(K.concat_rows(real_labels, real_rotated_labels),
                                 K.concat_rows(fake_labels, fake_rotated_labels))
    return all_features, all_labels

  def create_loss(self, features, labels, params=None, is_training=True, for_discriminator=True):
    """reference :198-321.  `labels` are class indices, or [B, num_classes] soft labels with use_soft_labels (a row of
    zeros = no label)."""
    real_images, fake_images = features["images"], features["generated"]
    real_labels = fake_labels = None
    if self.conditional:
      if self._use_soft_labels:
        if len(labels.shape) != 2 or labels.shape[1] != self._dataset.num_classes:
          raise ValueError("Need soft labels of dimension {} but got {}".format(self._dataset.num_classes, labels.shape))
        real_labels = labels
      else:
        real_labels = self._get_one_hot_labels(labels)
      fake_labels = self._get_one_hot_labels(features["sampled_labels"])
    bs = real_images.shape[0]
    rotation = self._self_supervision == "rotation"
    if rotation:
      if bs % self._rotated_batch_fraction != 0:
        raise ValueError("Rotated batch fraction is invalid: %d doesn't divide %d" % (self._rotated_batch_fraction, bs))
      rotated_bs = bs // self._rotated_batch_fraction
      num_rot_examples = rotated_bs // NUM_ROTATIONS
      if num_rot_examples <= 0:
        raise ValueError("rotated batch of %d examples holds no rotation group" % rotated_bs)
      all_features, all_labels = self.merge_with_rotation_data(real_images, fake_images, real_labels, fake_labels,
                                                               num_rot_examples)
    else:
      all_features = K.concat_rows(real_images, fake_images)
      all_labels = K.concat_rows(real_labels, fake_labels) if self.conditional else None
    d_predictions, d_logits, rot_logits, aux_logits, is_label_available = self.discriminator_with_additonal_heads(
        x=all_features, y=all_labels, is_training=is_training)
    expected = 2 * bs + (2 * (NUM_ROTATIONS - 1) * num_rot_examples if rotation else 0)
    if d_logits.shape[0] != expected:
      raise ValueError("Batch size unexpected: got %r expected %r" % (d_logits.shape[0], expected))
    half = expected // 2
    prob_real, prob_fake = K.slice_rows(d_predictions, 0, bs), K.slice_rows(d_predictions, half, half + bs)
    logits_real, logits_fake = K.slice_rows(d_logits, 0, bs), K.slice_rows(d_logits, half, half + bs)
    self.d_loss, _, _, self.g_loss = loss_lib.get_losses(
        d_real=prob_real, d_fake=prob_fake, d_real_logits=logits_real, d_fake_logits=logits_fake)
    if rotation:
      # the last 4 * num_rot_examples rows of each half: the upright tail of the batch followed by its three rotations,
      # labels [0..0, 1..1, 2..2, 3..3] (:279-290)
      rows = NUM_ROTATIONS * num_rot_examples
      real_loss = K.rotation_loss(K.slice_rows(rot_logits, half - rows, half), NUM_ROTATIONS)
      fake_loss = K.rotation_loss(K.slice_rows(rot_logits, 2 * half - rows, 2 * half), NUM_ROTATIONS)
      self.d_loss = K.add(self.d_loss, K.affine(real_loss, self._weight_rotation_loss_d))
      self.g_loss = K.add(self.g_loss, K.affine(fake_loss, self._weight_rotation_loss_g))
      self.rot_real_loss, self.rot_fake_loss = real_loss, fake_loss
    if self._use_predictor:
      # the predictor learns from the labelled REAL examples only (:305-318)
      real_aux_logits = K.slice_rows(aux_logits, 0, bs)
      weights = K.slice_rows(is_label_available, 0, bs)
      class_loss_real = K.softmax_xent(real_aux_logits, tape.DT(real_labels.t), weights)
      self.d_loss = K.add(self.d_loss, K.affine(class_loss_real, self._weight_class_loss))
      self.class_loss_real = class_loss_real
