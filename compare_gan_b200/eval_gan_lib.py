"""Evaluation of a trained GAN (reference eval_gan_lib.py:65-212): sample the generator in inference mode in batches
of 64, run Inception on the samples, compute FID / IS / KID.  Everything up to the [N,2048] statistics stays on the
GPU; sample generation can be sharded across ranks with a final all-reduce of the statistics."""
import time

import numpy as np
import torch

from . import eval_utils
from . import kernels as K
from . import tape
from . import variables as V
from .runner_lib import eval_z_generator
from .tpu import tpu_ops

NAN_DETECTED = eval_utils.NAN_DETECTED


def _update_bn_accumulators(gan, batch_size, num_accu_examples=204800, rng=None):
  """reference eval_gan_lib.py:65-92: fill accu_mean/accu_variance by running G with update_accus=1."""
  accus = [v for k, v in gan.store.vars.items() if k.endswith("accu/update_accus")]
  if not accus:
    return False
  rng = rng or np.random
  for v in accus:
    K.fill_(v, 1.0)
  for _ in range(int(np.ceil(num_accu_examples / batch_size))):
    generate_batch(gan, batch_size, rng)
  for v in accus:
    K.fill_(v, 0.0)
  return True


def generate_batch(gan, batch_size, rng):
  """One inference-mode G call (z from gin `eval_z`, labels ~ U{0..C-1}, reference :127-146)."""
  z = K.from_numpy(eval_z_generator((batch_size, gan._z_dim), rng=rng))
  y = None
  if gan.conditional:
    labels = rng.randint(0, gan._dataset.num_classes, batch_size).astype(np.int32)
    y = K.one_hot(tape.DT(torch.from_numpy(labels).to(z.t.device)), gan._dataset.num_classes)
  with V.use(gan.store), tape.no_record():
    return gan.generator(z, y=y, is_training=False)


class use_ema_weights(object):
  """Evaluate G with its EMA shadow (reference modular_gan.py:266-285 exports the hub module with the EMA getter)."""

  def __init__(self, gan):
    self.gan = gan

  def __enter__(self):
    g = self.gan
    if g.ema is not None:
      self.saved = g.flat_g["param"].t.clone()
      K.copy_(g.flat_g["param"], g.ema)
    return self

  def __exit__(self, *a):
    g = self.gan
    if g.ema is not None:
      g.flat_g["param"].t.copy_(self.saved)


def evaluate(gan, eval_tasks, num_averaging_runs=1, num_samples=None, batch_size=64, seed=42, num_accu_examples=204800,
             keep_features=True, real_images=None):
  """Mirrors evaluate_tfhub_module (reference eval_gan_lib.py:95-212).  Returns the result dict with
  `<label>_mean/_std/_list` keys plus `eval_samples_per_sec` (generation + Inception + statistics)."""
  dataset = gan._dataset
  n_total = num_samples or dataset.eval_test_samples
  world, rank = tpu_ops.num_replicas(), (torch.distributed.get_rank() if tpu_ops.num_replicas() > 1 else 0)
  n_local = n_total // world + (1 if rank < n_total % world else 0)
  rng = np.random.RandomState(seed + 1000 * rank)
  K.sync_stream()
  fake_dsets, timings = [], []
  with use_ema_weights(gan):
    _update_bn_accumulators(gan, batch_size, num_accu_examples, rng)
    for _ in range(num_averaging_runs):
      acc = eval_utils.FeatureAccumulator(keep_features=keep_features)
      torch.cuda.synchronize()
      t0 = time.time()
      done = 0
      while done < n_local:
        imgs = generate_batch(gan, batch_size, rng)
        pool, logits = eval_utils.inception_transform(imgs)
        valid = min(batch_size, n_local - done)
        acc.add(pool, logits, valid)
        done += valid
      torch.cuda.synchronize()
      timings.append(time.time() - t0)
      sample = acc.finish(eval_utils.EvalDataSample())
      if sample.activations is not None and not np.isfinite(sample.activations).all():
        raise eval_utils.NanFoundError("NaN in generated samples")
      fake_dsets.append(sample)
  if real_images is None:
    real_images = dataset.sample_images(n_local)
  racc = eval_utils.inception_transform_np(real_images * 255.0, batch_size, keep_features=keep_features)
  real_dset = racc.finish(eval_utils.EvalDataSample())
  result = {}
  for task in eval_tasks:
    dicts = [task.run_after_session(f, real_dset) for f in fake_dsets]
    for key in dicts[0]:
      scores = np.array([d[key] for d in dicts])
      result[key + "_mean"] = float(np.mean(scores))
      result[key + "_std"] = float(np.std(scores))
      result[key + "_list"] = "_".join(str(x) for x in scores)
  result["eval_samples_per_sec"] = n_total / float(np.mean(timings))
  return result
