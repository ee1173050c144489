"""Shared helpers for the -m gpu parity tests: build the engine (CUDA, through the C-ABI) and the CPU oracle
with identical configuration and identical initial weights."""
import numpy as np
import torch

from oracle import gan as ogan
from oracle import nets as onets


def rel_err(a, b):
  a = np.asarray(a, np.float64)
  b = np.asarray(b, np.float64)
  d = np.linalg.norm((a - b).ravel())
  n = np.linalg.norm(b.ravel())
  return d / max(n, 1e-30)


def assert_close(a, b, tol, what=""):
  a, b = np.asarray(a), np.asarray(b)
  assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
  e = rel_err(a, b)
  assert np.isfinite(a).all(), "%s: non-finite values" % what
  assert e <= tol, "%s: rel-L2 error %.3e > %.1e" % (what, e, tol)


ARCH_IMAGE = {"resnet_cifar_arch": (32, 32, 3)}


def make_pair(arch, image_shape, batch, loss="non_saturating", penalty="no_penalty", lamba=1.0, disc_iters=1,
              g_bn="batch_norm", g_sn=False, d_sn=False, sn_singular="left", conditional=False, num_classes=0,
              initializer="normal", use_moving_averages=True, bn_decay=0.9, bn_eps=1e-5, g_lr=2e-4, d_lr=None,
              beta1=0.5, beta2=0.999, z_dim=128, g_use_ema=False, ema_start_step=0, ch=8, extra_bindings=(),
              project_y=False, seed=0):
  """Returns (engine ModularGAN built for `batch`, GanOracle) sharing config and weights."""
  from compare_gan_b200 import gin_lite as gin
  from compare_gan_b200 import datasets
  from compare_gan_b200.gans import modular_gan  # noqa: F401  (registers configurables)
  gin.clear_config()
  bn_ref = {"batch_norm": "@batch_norm", "conditional_batch_norm": "@conditional_batch_norm", None: "None"}[g_bn]
  cfg = [
      "G.batch_norm_fn = %s" % bn_ref,
      "G.spectral_norm = %s" % g_sn,
      "D.spectral_norm = %s" % d_sn,
      "spectral_norm.singular_value = '%s'" % sn_singular,
      "standardize_batch.decay = %r" % bn_decay,
      "standardize_batch.epsilon = %r" % bn_eps,
      "standardize_batch.use_moving_averages = %s" % use_moving_averages,
      "weights.initializer = '%s'" % initializer,
      "loss.fn = @%s" % loss,
      "penalty.fn = @%s" % penalty,
      "ModularGAN.g_lr = %r" % g_lr,
      "ModularGAN.g_optimizer_fn = @tf.train.AdamOptimizer",
      "tf.train.AdamOptimizer.beta1 = %r" % beta1,
      "tf.train.AdamOptimizer.beta2 = %r" % beta2,
      "ModularGAN.conditional = %s" % conditional,
      "ModularGAN.g_use_ema = %s" % g_use_ema,
      "ModularGAN.ema_start_step = %d" % ema_start_step,
      "resnet_biggan.Generator.ch = %d" % ch,
      "resnet_biggan.Discriminator.ch = %d" % ch,
      "resnet_biggan.Discriminator.project_y = %s" % project_y,
      "resnet_cifar.Discriminator.project_y = %s" % project_y,
  ]
  if d_lr is not None:
    cfg.append("ModularGAN.d_lr = %r" % d_lr)
  cfg += list(extra_bindings)
  gin.parse_config("\n".join(cfg))
  ds = datasets.ImageDatasetV2("synthetic", image_shape[0], image_shape[2], num_classes or None, 100)
  params = {"architecture": arch, "z_dim": z_dim, "lambda": lamba, "disc_iters": disc_iters, "seed": seed}
  eng = modular_gan.ModularGAN(dataset=ds, parameters=params, model_dir="/tmp/cgan_test")
  eng.build(batch)

  hier = arch == "resnet_biggan_arch"
  ocfg = onets.Cfg(architecture=arch, image_shape=tuple(image_shape), g_bn=g_bn, g_sn=g_sn, d_sn=d_sn,
                   sn_singular=sn_singular, bn_decay=bn_decay, bn_eps=bn_eps,
                   use_moving_averages=use_moving_averages, initializer=initializer, ch=ch, project_y=project_y,
                   hierarchical_z=hier, embed_y=hier, num_classes=num_classes)
  for b in extra_bindings:
    if "Generator.blocks_with_attention" in b:
      ocfg.g_attention = b.split("=")[1].strip().strip("'\"")
    if "Discriminator.blocks_with_attention" in b:
      ocfg.d_attention = b.split("=")[1].strip().strip("'\"")
  orc = ogan.GanOracle(ocfg, loss=loss, penalty=penalty, lamba=lamba, disc_iters=disc_iters, g_lr=g_lr, d_lr=d_lr,
                       beta1=beta1, beta2=beta2, conditional=conditional, g_use_ema=g_use_ema,
                       ema_start_step=ema_start_step, z_dim=z_dim).build(batch)
  state = eng.state_numpy()
  onames = list(orc.store.vars.keys())
  enames = list(state.keys())
  assert sorted(onames) == sorted(enames), ("variable sets differ", sorted(set(onames) ^ set(enames))[:10])
  assert [k for k in orc.store.trainable] == [k for k in eng.store.trainable], "trainable variable order differs"
  orc.store.load_numpy(state)
  return eng, orc


def make_inputs(rng, k, batch, image_shape, z_dim, num_classes=0, z_normal=False, gp=False):
  imgs = [rng.rand(batch, *image_shape).astype(np.float32) for _ in range(k + 1)]
  if z_normal:
    zs = [rng.standard_normal((batch, z_dim)).astype(np.float32) for _ in range(k + 1)]
  else:
    zs = [rng.uniform(-1, 1, (batch, z_dim)).astype(np.float32) for _ in range(k + 1)]
  labels = sampled = None
  if num_classes:
    labels = [rng.randint(0, num_classes, batch).astype(np.int32) for _ in range(k + 1)]
    sampled = [rng.randint(0, num_classes, batch).astype(np.int32) for _ in range(k + 1)]
  alphas = [rng.rand(batch, 1, 1, 1).astype(np.float32) for _ in range(k + 1)] if gp else None
  return imgs, zs, labels, sampled, alphas


def compare_states(eng, orc, tol, skip=()):
  es = eng.state_numpy()
  worst = (0.0, None)
  for k, v in orc.store.vars.items():
    if any(s in k for s in skip):
      continue
    a, b = es[k], v.detach().numpy()
    e = rel_err(a, b)
    if e > worst[0]:
      worst = (e, k)
    assert np.isfinite(a).all(), k
    assert e <= tol, "%s: rel-L2 error %.3e > %.1e" % (k, e, tol)
  return worst
