"""Runner surface (reference runner_lib.py:72-111, 280-354): the gin `options` dict and a plain Python
training loop in place of TPUEstimator.train.  Checkpoint polling / CSV task manager are out of the
accelerated path."""
import csv
import glob
import collections
import os
import re
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


def sample_cycle_inputs(gan, dataset, batch_size, rng, real=None):
  """One unrolled cycle of inputs: disc_iters+1 sub-steps, fresh images/z each (reference
  gans/modular_gan.py:218-223, 410-426).  Images/labels are synthetic unless `real` = (images, labels) lists from the
  input pipeline are given."""
  k = gan._disc_iters
  images = real[0] if real is not None else [dataset.sample_images(batch_size) for _ in range(k + 1)]
  z = [z_generator((batch_size, gan._z_dim), rng=rng) for _ in range(k + 1)]
  labels = sampled = None
  if gan.conditional:
    labels = real[1] if real is not None else [dataset.sample_labels(batch_size) for _ in range(k + 1)]
    sampled = [rng.randint(0, dataset.num_classes, batch_size).astype(np.int32) for _ in range(k + 1)]
  alphas = [rng.rand(batch_size, 1, 1, 1).astype(np.float32) for _ in range(k + 1)]
  return images, z, labels, sampled, alphas


class PipelineFeeder(object):
  """Feeds training cycles from `dataset.train_input_fn` (reference datasets.py:261-291 + the TPUEstimator infeed):
  each cycle takes disc_iters+1 page-locked batches from the native loader, `set_inputs` copies them asynchronously, and
  the ring slots of a cycle are handed back once a CUDA event recorded behind its copies has completed — checked two
  cycles later, so the host never waits on the copies it has just issued."""

  def __init__(self, gan, dataset, batch_size, rank=0):
    self.k1 = gan._disc_iters + 1
    self.it = dataset.train_input_fn({"batch_size": batch_size}, rank=rank, ring=3 * self.k1 + 1)
    self.pending = collections.deque()

  def feed(self, gan, dataset, batch_size, rng):
    import torch
    while len(self.pending) >= 2:
      self.pending.popleft().synchronize()
      self.it.release(self.k1)
    batches = [next(self.it) for _ in range(self.k1)]
    real = ([b[0] for b in batches], [b[1] for b in batches])
    gan.set_inputs(*sample_cycle_inputs(gan, dataset, batch_size, rng, real=real))
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream())
    self.pending.append(ev)

  def close(self):
    while self.pending:
      self.pending.popleft().synchronize()
    self.it.close()


class TaskManager(object):
  """Interface for managing a task (reference runner_lib.py:114-199); checkpoints are this package's
  `model.ckpt-<step>.npz` or the reference's own TensorFlow checkpoints `model.ckpt-<step>.index` (+ `.data-*`), so a
  model_dir trained by the reference can be evaluated by `continuous_eval` as is."""

  def __init__(self, model_dir):
    self._model_dir = model_dir

  @property
  def model_dir(self):
    return self._model_dir

  def mark_training_done(self):
    os.makedirs(self.model_dir, exist_ok=True)
    open(os.path.join(self.model_dir, "TRAIN_DONE"), "w").close()

  def is_training_done(self):
    return os.path.exists(os.path.join(self.model_dir, "TRAIN_DONE"))

  def add_eval_result(self, checkpoint_path, result_dict, default_value):
    pass

  def get_checkpoints_with_results(self):
    return set()

  def unevaluated_checkpoints(self, timeout=0, eval_every_steps=None):
    """Generator for checkpoints without evaluation results (reference :137-180)."""
    evaluated = self.get_checkpoints_with_results()
    last_eval = time.time()
    while True:
      ckpts = set(glob.glob(os.path.join(self.model_dir, "model.ckpt-*.npz")))
      ckpts |= set(glob.glob(os.path.join(self.model_dir, "model.ckpt-*.index")))
      todo = sorted([(int(re.findall(r"ckpt-(\d+)\.(?:npz|index)$", c)[0]), c) for c in ckpts - evaluated])
      if eval_every_steps:
        todo = [(s, c) for s, c in todo if s > 0 and s % eval_every_steps == 0]
      for _, c in todo:
        yield c
      if todo:
        evaluated |= set(c for _, c in todo)
        last_eval = time.time()
        continue
      if time.time() - last_eval > timeout or self.is_training_done():
        break
      time.sleep(60)

  def report_progress(self, message):
    pass


class TaskManagerWithCsvResults(TaskManager):
  """Task Manager that writes results to a CSV file (reference runner_lib.py:182-232): columns are
  checkpoint_path, step, the sorted result keys, then the sorted operative gin bindings."""

  def __init__(self, model_dir, score_file=None):
    super(TaskManagerWithCsvResults, self).__init__(model_dir)
    self._score_file = score_file or os.path.join(model_dir, "scores.csv")

  def _get_config_for_step(self, step):
    saved = glob.glob(os.path.join(self.model_dir, "operative_config-*.gin"))
    steps = sorted(int(re.findall(r"operative_config-(\d+).gin", fn)[0]) for fn in saved)
    assert steps
    last = [s for s in steps if s <= int(step)][-1]
    config = {}
    for line in open(os.path.join(self.model_dir, "operative_config-%d.gin" % last)):
      if "=" in line:
        k, v = line.split("=", 1)
        config[k.strip()] = v.strip()
    return config

  def add_eval_result(self, checkpoint_path, result_dict, default_value):
    step = re.findall(r"ckpt-(\d+)\.(?:npz|index)$", checkpoint_path)[0]
    config = self._get_config_for_step(step)
    header = ["checkpoint_path", "step"] + sorted(result_dict) + sorted(config)
    write_header = not os.path.exists(self._score_file)
    row = dict(checkpoint_path=checkpoint_path, step=step, **config)
    for k, v in result_dict.items():
      row[k] = "{:.3f}".format(v) if isinstance(v, float) else v
    with open(self._score_file, "a") as f:
      writer = csv.DictWriter(f, fieldnames=header, extrasaction="ignore")
      if write_header:
        writer.writeheader()
      writer.writerow(row)

  def get_checkpoints_with_results(self):
    if not os.path.exists(self._score_file):
      return set()
    with open(self._score_file) as f:
      return {r["checkpoint_path"] for r in csv.DictReader(f)}


def _run_eval(gan, task_manager, eval_tasks=None, num_averaging_runs=1, num_samples=None, eval_every_steps=None,
              timeout=0, write=True):
  """Evaluates all unevaluated checkpoints (reference runner_lib.py:235-277); NaN samples score NAN_DETECTED."""
  from . import eval_gan_lib
  from .metrics import fid_score, inception_score
  eval_tasks = eval_tasks or [inception_score.InceptionScoreTask(), fid_score.FIDScoreTask()]
  results = {}
  for ckpt in task_manager.unevaluated_checkpoints(timeout=timeout, eval_every_steps=eval_every_steps):
    gan.load_checkpoint(ckpt)
    default_value = -1.0
    try:
      result_dict = eval_gan_lib.evaluate(gan, eval_tasks, num_averaging_runs=num_averaging_runs, num_samples=num_samples)
    except ValueError:
      result_dict = {}
    except Exception as e:       # NanFoundError
      if type(e).__name__ != "NanFoundError":
        raise
      result_dict = {}
      default_value = eval_gan_lib.NAN_DETECTED
    if write:          # (sharded evaluation: every rank takes part, rank 0 records the scores)
      task_manager.add_eval_result(ckpt, result_dict, default_value)
    results[ckpt] = result_dict
  return results


def run_with_schedule(schedule, options=None, model_dir="/tmp/compare_gan_b200", num_cycles=None, use_graph=True,
                      seed=0, task_manager=None, save_every_cycles=None, eval_kwargs=None, input_pipeline=False):
  """Run the schedule `train` / `eval_after_train` / `continuous_eval` (reference runner_lib.py:280-354).  Training
  images are synthetic draws by default; with `input_pipeline` they come from `dataset.train_input_fn` (the reference's
  fake data set or on-disk shards) through the native prefetching loader."""
  if schedule not in ("train", "eval_after_train", "continuous_eval"):
    raise ValueError("Schedule {} not supported.".format(schedule))
  options = options or get_options_dict()
  task_manager = task_manager or TaskManagerWithCsvResults(model_dir)
  if schedule == "continuous_eval":
    dataset = datasets.get_dataset()
    gan = options["gan_class"](dataset=dataset, parameters=options, model_dir=model_dir)
    from .tpu import tpu_ops as _t
    gan.build(options["batch_size"] // _t.num_replicas())
    return {"eval": _run_eval(gan, task_manager, timeout=24 * 3600, **(eval_kwargs or {})), "gan": gan}
  dataset = datasets.get_dataset()
  gan = options["gan_class"](dataset=dataset, parameters=options, model_dir=model_dir)
  from .tpu import tpu_ops
  world = tpu_ops.num_replicas()
  rank = 0
  if world > 1:
    import torch.distributed as dist
    rank = dist.get_rank()
  per_replica = options["batch_size"] // world
  gan.build(per_replica)
  if use_graph:
    gan.capture()
  # every replica draws its OWN shard of the global batch (z, sampled labels, alphas, synthetic images): identical streams
  # would make the all-reduced gradient equal one replica's, i.e. an effective batch of per_replica
  rng = np.random.RandomState(seed + rank)
  if rank and hasattr(dataset, "_rng"):
    dataset._rng = np.random.RandomState(getattr(dataset, "_seed", 547) + rank)
  cycles = num_cycles if num_cycles is not None else options["training_steps"] // max(1, options["disc_iters"])
  t0 = time.time()
  feeder = None
  if input_pipeline:
    feeder = PipelineFeeder(gan, dataset, per_replica, rank=rank)
  for _ in range(cycles):
    if feeder is not None:
      feeder.feed(gan, dataset, per_replica, rng)
    else:
      gan.set_inputs(*sample_cycle_inputs(gan, dataset, per_replica, rng))
    gan.run_cycle()
    if rank == 0 and save_every_cycles and (_ + 1) % save_every_cycles == 0:
      gan.save_checkpoint(model_dir)
  d_losses, g_loss = gan.read_losses()
  if feeder is not None:
    feeder.close()
  if rank == 0:      # replicas hold identical weights: one writer for checkpoints, the operative config and the markers
    os.makedirs(model_dir, exist_ok=True)
    with open(os.path.join(model_dir, "operative_config-0.gin"), "w") as f:     # GinConfigSaverHook, runner_lib.py:319
      f.write(gin.operative_config_str())
    gan.save_checkpoint(model_dir)
    task_manager.mark_training_done()
  if world > 1:
    import torch.distributed as dist
    dist.barrier()
  out = {"d_loss": d_losses, "g_loss": g_loss, "cycles": cycles, "seconds": time.time() - t0, "gan": gan}
  if schedule == "eval_after_train":
    out["eval"] = _run_eval(gan, task_manager, write=rank == 0, **(eval_kwargs or {}))
  return out


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
