"""Runner surface (reference runner_lib.py:72-111, 280-354): the gin `options` dict and a plain Python
training loop in place of TPUEstimator.train.  Checkpoint polling / CSV task manager are out of the
accelerated path."""
import time

import numpy as np

from . import datasets
from . import gin_lite as gin


@gin.configurable("options")
def get_options_dict(batch_size=gin.REQUIRED, gan_class=gin.REQUIRED, architecture=gin.REQUIRED,
                     training_steps=gin.REQUIRED, discriminator_normalization=None, lamba=1, disc_iters=1,
                     z_dim=128):
  """Parse legacy options from Gin configurations into a Python dict (reference runner_lib.py:72-111)."""
  del discriminator_normalization
  return {
      "use_tpu": False,
      "batch_size": batch_size,
      "gan_class": gan_class,
      "architecture": architecture,
      "training_steps": training_steps,
      "lambda": lamba,  # Different spelling intended.
      "disc_iters": disc_iters,
      "z_dim": z_dim,
  }


@gin.configurable("z")
def z_generator(shape, distribution_fn=None, minval=-1.0, maxval=1.0, stddev=1.0, rng=None):
  """reference gans/modular_gan.py:365-391 (uniform [-1,1) default; normal when bound to tf.random.normal)."""
  rng = rng or np.random
  if distribution_fn is not None and getattr(distribution_fn, "__name__", "") == "normal":
    return (rng.standard_normal(shape) * stddev).astype(np.float32)
  return rng.uniform(minval, maxval, shape).astype(np.float32)


def _normal(*a, **k):
  raise RuntimeError("placeholder for @tf.random.normal; z_generator samples on the host")


def _uniform(*a, **k):
  raise RuntimeError("placeholder for @tf.random.uniform; z_generator samples on the host")


_normal.__name__ = "normal"
_uniform.__name__ = "uniform"
gin.external_configurable(_normal, "tf.random.normal")
gin.external_configurable(_uniform, "tf.random.uniform")


def sample_cycle_inputs(gan, dataset, batch_size, rng):
  """One unrolled cycle of synthetic inputs: disc_iters+1 sub-steps, fresh images/z each
  (reference gans/modular_gan.py:218-223, 410-426)."""
  k = gan._disc_iters
  images = [dataset.sample_images(batch_size) for _ in range(k + 1)]
  z = [z_generator((batch_size, gan._z_dim), rng=rng) for _ in range(k + 1)]
  labels = sampled = None
  if gan.conditional:
    labels = [dataset.sample_labels(batch_size) for _ in range(k + 1)]
    sampled = [rng.randint(0, dataset.num_classes, batch_size).astype(np.int32) for _ in range(k + 1)]
  alphas = [rng.rand(batch_size, 1, 1, 1).astype(np.float32) for _ in range(k + 1)]
  return images, z, labels, sampled, alphas


def run_with_schedule(schedule, options=None, model_dir="/tmp/compare_gan_b200", num_cycles=None, use_graph=True,
                      seed=0):
  """Run the `train` schedule on synthetic data (reference runner_lib.py:280-354)."""
  if schedule != "train":
    raise ValueError("Schedule {} not supported on the accelerated path.".format(schedule))
  options = options or get_options_dict()
  dataset = datasets.get_dataset()
  gan = options["gan_class"](dataset=dataset, parameters=options, model_dir=model_dir)
  from .tpu import tpu_ops
  per_replica = options["batch_size"] // tpu_ops.num_replicas()
  gan.build(per_replica)
  if use_graph:
    gan.capture()
  rng = np.random.RandomState(seed)
  cycles = num_cycles if num_cycles is not None else options["training_steps"] // max(1, options["disc_iters"])
  t0 = time.time()
  for _ in range(cycles):
    gan.set_inputs(*sample_cycle_inputs(gan, dataset, per_replica, rng))
    gan.run_cycle()
  d_losses, g_loss = gan.read_losses()
  return {"d_loss": d_losses, "g_loss": g_loss, "cycles": cycles, "seconds": time.time() - t0, "gan": gan}


@gin.configurable("run_config")
def get_run_config(tf_random_seed=None, single_core=False, iterations_per_loop=1000, save_checkpoints_steps=5000,
                   keep_checkpoint_max=1000):
  """reference main.py:79-95 (only the seed matters on this path: it seeds the host-side initialisers)."""
  return {"tf_random_seed": tf_random_seed, "single_core": single_core, "iterations_per_loop": iterations_per_loop,
          "save_checkpoints_steps": save_checkpoints_steps, "keep_checkpoint_max": keep_checkpoint_max}


@gin.configurable("eval_z")
def eval_z_generator(shape, distribution_fn=None, minval=-1.0, maxval=1.0, stddev=1.0, rng=None):
  """reference eval_gan_lib.py:43-62."""
  return z_generator(shape, distribution_fn=distribution_fn, minval=minval, maxval=maxval, stddev=stddev, rng=rng)
