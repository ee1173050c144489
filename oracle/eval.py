"""Oracle (test infrastructure): the evaluation loop of the reference restated on PyTorch-CPU —
compare_gan/eval_gan_lib.py:65-92 (`_update_bn_accumulators`), :95-212 (`evaluate_tfhub_module`: fixed seed, batches of
64, z from `eval_z`, labels ~ U{0..C-1}, `num_averaging_runs` fake data sets, real features, per-task mean/std/list) and
the hub-module export it evaluates: compare_gan/gans/modular_gan.py:266-285 (the generator runs with is_training=False
and, when `g_use_ema`, with the ExponentialMovingAverage shadows in place of the raw weights).

The RNG stream is a numpy RandomState instead of TF's (the reference seeds np.random and tf with 42; TF's stream is
not restated): the engine's eval loop draws z, then labels, batch by batch from the same kind of stream, so feeding
both the same seed makes them see identical latents.
"""
import numpy as np
import torch

from . import inception as oinc
from . import metrics as ometrics
from . import nets


def z_generator(shape, rng, distribution="uniform", minval=-1.0, maxval=1.0, stddev=1.0):
  """eval_gan_lib.py:43-62 (gin `eval_z`): tf.random.uniform(minval, maxval) by default, tf.random.normal(stddev) for BigGAN."""
  if distribution == "uniform":
    return rng.uniform(minval, maxval, shape).astype(np.float32)
  if distribution == "normal":
    return (rng.standard_normal(shape) * stddev).astype(np.float32)
  raise ValueError(distribution)


class _EmaWeights(object):
  """modular_gan.py:266-285: the exported generator reads `<var>/ExponentialMovingAverage` where a shadow exists."""

  def __init__(self, oracle):
    self.o = oracle

  def __enter__(self):
    self.saved = None
    if self.o.g_use_ema and self.o.ema is not None:
      params = self.o.store.trainable_under("generator")
      self.saved = {k: v.detach().clone() for k, v in params.items()}
      with torch.no_grad():
        for k, v in params.items():
          v.copy_(self.o.ema[k])
    return self

  def __exit__(self, *a):
    if self.saved is not None:
      with torch.no_grad():
        for k, v in self.o.store.trainable_under("generator").items():
          v.copy_(self.saved[k])


def sample_batch(oracle, batch_size, rng, z_kw=None):
  """One `generated` fetch of sample_from_generator (eval_gan_lib.py:127-146): z, then labels, inference-mode G."""
  z = z_generator((batch_size, oracle.z_dim), rng, **(z_kw or {}))
  y = None
  if oracle.conditional:
    labels = rng.randint(0, oracle.cfg.num_classes, batch_size)
    y = oracle.one_hot(labels)
  with torch.no_grad():
    return nets.generator(oracle.store, oracle.cfg, torch.from_numpy(z).to(oracle.dtype), y, False)


def update_bn_accumulators(oracle, batch_size, num_accu_examples, rng, z_kw=None):
  """eval_gan_lib.py:65-92: switch every `accu/update_accus` to 1, fetch `generated` num_accu_examples // batch_size
  times, switch back.  The accumulators start from their initial values (a freshly loaded module: 0, 0, 1e-12)."""
  switches = [k for k in oracle.store.vars if k.endswith("accu/update_accus")]
  if not switches:
    return False
  with torch.no_grad():
    for k, v in oracle.store.vars.items():
      if k.endswith("accu/accu_mean") or k.endswith("accu/accu_variance"):
        v.zero_()
      elif k.endswith("accu/accu_counter"):
        v.fill_(1e-12)
    for k in switches:
      oracle.store.vars[k].fill_(1.0)
  for _ in range(num_accu_examples // batch_size):
    sample_batch(oracle, batch_size, rng, z_kw)
  with torch.no_grad():
    for k in switches:
      oracle.store.vars[k].fill_(0.0)
  return True


def evaluate(oracle, inception_weights, real_images, num_samples, batch_size=64, seed=42, num_averaging_runs=1,
             num_accu_examples=204800, z_kw=None):
  """evaluate_tfhub_module (eval_gan_lib.py:95-212) with the default task list's three scores.  Returns the result
  dict (`<label>_mean/_std/_list`) and the generated images of the first run."""
  rng = np.random.RandomState(seed)
  num_batches = int(np.ceil(num_samples / float(batch_size)))
  w = {k: torch.from_numpy(np.asarray(v)) for k, v in inception_weights.items()}

  def features(images):
    acts, logits = [], []
    for i in range(0, len(images), batch_size):
      p, l = oinc.inception_v3(oinc.preprocess(images[i:i + batch_size]), w)
      acts.append(p.numpy()); logits.append(l.numpy())
    return np.concatenate(acts), np.concatenate(logits)
  fakes = []
  with _EmaWeights(oracle):
    update_bn_accumulators(oracle, batch_size, num_accu_examples, rng, z_kw)
    for _ in range(num_averaging_runs):
      imgs = np.concatenate([sample_batch(oracle, batch_size, rng, z_kw).numpy() for _ in range(num_batches)])[:num_samples]
      fakes.append(imgs)
  real_acts, _ = features(np.asarray(real_images, np.float32)[:num_samples])
  per_run = []
  for imgs in fakes:
    if not np.isfinite(imgs).all():
      raise ValueError("NaN in generated samples")
    acts, logits = features(imgs)
    per_run.append({"fid_score": ometrics.compute_fid_from_activations(real_acts, acts),
                    "inception_score": ometrics.inception_score_from_logits(logits),
                    "kid_score": ometrics.kid(acts, real_acts)})
  result = {}
  for key in per_run[0]:
    scores = np.array([d[key] for d in per_run])
    result[key + "_mean"], result[key + "_std"] = float(np.mean(scores)), float(np.std(scores))
    result[key + "_list"] = "_".join(str(x) for x in scores)
  return result, fakes[0]
