"""Error behaviour of the mirrored Python surface: the same exception types, raised for the same misuse, as
the reference (arch_ops.py:255-279, 427-430, 470-472; resnet_ops.py:115-119; loss_lib.py:39-43;
modular_gan.py:184-187; runner_lib.py:300-301; datasets.py:646-647).  All of them fire before any kernel is touched,
so this runs without a GPU."""
import numpy as np
import pytest

from compare_gan_b200 import datasets, gin_lite as gin, runner_lib
from compare_gan_b200.architectures import arch_ops, resnet_ops
from compare_gan_b200.gans import loss_lib, modular_gan


class Shaped(object):
  """Stands in for a device tensor where only `.shape` is inspected."""

  def __init__(self, *shape):
    self.shape = tuple(shape)


def setup_function(_):
  gin.clear_config()


def test_arch_ops_reject_bad_arguments():
  with pytest.raises(ValueError, match="Invalid data_format"):
    arch_ops.standardize_batch(Shaped(2, 4, 4, 3), is_training=True, data_format="NWHC")
  with pytest.raises(ValueError, match="unsupported rank"):
    arch_ops.standardize_batch(Shaped(2, 4, 3), is_training=True)
  with pytest.raises(ValueError, match="provide y"):
    arch_ops.conditional_batch_norm(Shaped(2, 4, 4, 3), None, is_training=True, use_sn=False)
  with pytest.raises(ValueError, match="rank 2"):
    arch_ops.conditional_batch_norm(Shaped(2, 4, 4, 3), Shaped(2, 1, 10), is_training=True, use_sn=False)
  with pytest.raises(ValueError, match="multi-dimensional"):
    arch_ops.spectral_norm(Shaped(7))
  with pytest.raises(ValueError, match="square strides"):
    arch_ops.conv2d(Shaped(2, 4, 4, 3), 8, 3, 3, 1, 2)
  with pytest.raises(ValueError, match="Unknown weight initializer"):
    arch_ops.weight_initializer(initializer="he")


def test_resnet_ops_reject_bad_arguments():
  with pytest.raises(ValueError, match="rank 4"):
    resnet_ops.validate_image_inputs(Shaped(2, 4, 4))
  with pytest.raises(ValueError, match="equal width and height"):
    resnet_ops.validate_image_inputs(Shaped(2, 4, 8, 3))
  with pytest.raises(ValueError, match="power of 2"):
    resnet_ops.validate_image_inputs(Shaped(2, 6, 6, 3))


def test_losses_reject_mismatched_logits():
  with pytest.raises(ValueError, match="Shape mismatch"):
    loss_lib.check_dimensions(Shaped(4, 1), Shaped(3, 1), Shaped(4, 1), Shaped(4, 1))
  with pytest.raises(ValueError, match="Rank"):
    loss_lib.check_dimensions(Shaped(4), Shaped(4), Shaped(4), Shaped(4))


def test_modular_gan_runner_and_dataset_errors():
  ds = datasets.get_dataset("cifar10")
  gan = modular_gan.ModularGAN(dataset=ds, parameters={"architecture": "no_such_arch", "z_dim": 8, "lambda": 1,
                                                       "disc_iters": 1}, model_dir="/tmp/x")
  with pytest.raises(NotImplementedError, match="not implemented"):
    gan.generator
  with pytest.raises(NotImplementedError, match="not implemented"):
    gan.discriminator
  celeb = datasets.get_dataset("celeb_a")
  with pytest.raises(ValueError, match="does not have labels"):
    modular_gan.ModularGAN(dataset=celeb, parameters={"architecture": "resnet5_arch", "z_dim": 8, "lambda": 1,
                                                      "disc_iters": 1}, model_dir="/tmp/x", conditional=True)
  with pytest.raises(ValueError, match="not supported"):
    runner_lib.run_with_schedule("train_and_dance", options={})
  with pytest.raises(ValueError, match="not available"):
    datasets.get_dataset("mnist_on_mars")
