"""CPU trace harness for the host-side network definitions (no GPU, no kernels).

The architecture code (`compare_gan_b200/architectures/*`, `arch_ops`, `resnet_ops`) only composes calls into
`compare_gan_b200.kernels` and `variables.get_variable`.  Here every kernel-layer function it uses is replaced by a
shape-propagating stand-in that records (function, argument shapes / scalars / data-flow ids, result shape), and every
variable creation is recorded with its full name, shape, trainability and a hash of its initial value.  Two
implementations that produce the same trace launch the same kernels on the same operands in the same order, so a
refactoring of the definitions can be proven behaviour-preserving without a GPU; the trace also pins the engine's
variable key space (the reference's checkpoint names) on the CPU.
"""
import contextlib
import hashlib
import math

import numpy as np

from compare_gan_b200 import kernels as K
from compare_gan_b200 import gin_lite as gin
from compare_gan_b200 import variables as V


class FakeDT(object):
  """Stands in for tape.DT: shape, numel, an identity for data-flow tracking."""
  _next = [0]

  def __init__(self, shape, req=False):
    self.shape = tuple(int(s) for s in shape)
    self.numel = int(np.prod(self.shape)) if self.shape else 1
    self.req = req
    self.node = None
    self.id = FakeDT._next[0]
    FakeDT._next[0] += 1


class Tracer(object):

  def __init__(self):
    self.ops = []
    self.variables = []

  def describe(self, a):
    if isinstance(a, FakeDT):
      return ["T", a.id, list(a.shape)]
    if isinstance(a, K.BNState):
      return ["BNState"] + [self.describe(getattr(a, f)) for f in K.BNState.__slots__]
    if isinstance(a, np.ndarray):
      return ["ndarray", list(a.shape), str(a.dtype)]
    if isinstance(a, (list, tuple)):
      return [self.describe(v) for v in a]
    if callable(a):
      return ["callable", getattr(a, "__name__", type(a).__name__)]
    if isinstance(a, float):
      return float(np.float32(a)) if math.isfinite(a) else str(a)
    if isinstance(a, (np.integer,)):
      return int(a)
    if isinstance(a, (np.floating,)):
      return float(a)
    return a          # int, bool, str, None

  def record(self, name, args, kwargs, out):
    self.ops.append([name, [self.describe(a) for a in args], {k: self.describe(v) for k, v in sorted(kwargs.items())},
                     self.describe(out)])


def _same(i=0):
  return lambda *a, **k: a[i].shape


def _reshape(x, *shape):
  shape = list(shape[0]) if len(shape) == 1 and isinstance(shape[0], (list, tuple)) else list(shape)
  if -1 in shape:
    known = int(np.prod([s for s in shape if s != -1]))
    shape[shape.index(-1)] = x.numel // known
  assert int(np.prod(shape)) == x.numel, (x.shape, shape)
  return shape


def _conv2d(x, w, bias=None, stride=1, upsample=False, padding="SAME", relu=False, residual=None, round_out=False):
  n, h, ww, cin = x.shape
  kh, kw, wcin, cout = w.shape
  if wcin != cin:
    raise ValueError("conv2d: kernel expects %d input channels, got %d" % (wcin, cin))
  vh, vw = (2 * h, 2 * ww) if upsample else (h, ww)
  if padding == "SAME":
    out = (n, -(-vh // stride), -(-vw // stride), cout)
  else:
    out = (n, (vh - kh) // stride + 1, (vw - kw) // stride + 1, cout)
  if residual is not None and tuple(residual.shape) != out:
    raise ValueError("conv2d: residual shape %s does not match the output %s" % (residual.shape, out))
  return out


def _deconv2d(x, w, bias, out_hw, stride):
  n, h, ww, cin = x.shape
  kh, kw, cout, wcin = w.shape
  if wcin != cin:
    raise ValueError("deconv2d: kernel expects %d input channels, got %d" % (wcin, cin))
  oh, ow = out_hw
  if (-(-oh // stride), -(-ow // stride)) != (h, ww):
    raise ValueError("deconv2d: output shape %s incompatible with input %s" % ((oh, ow), (h, ww)))
  return (n, oh, ow, cout)


def _matmul(a, b, ta=False, tb=False):
  m = a.shape[1] if ta else a.shape[0]
  k = a.shape[0] if ta else a.shape[1]
  kb = b.shape[1] if tb else b.shape[0]
  n = b.shape[0] if tb else b.shape[1]
  if k != kb:
    raise ValueError("matmul: inner dimensions differ: %d vs %d" % (k, kb))
  return (m, n)


def _bmm(a, b, ta=False, tb=False):
  m = a.shape[2] if ta else a.shape[1]
  n = b.shape[1] if tb else b.shape[2]
  return (a.shape[0], m, n)


SHAPE_RULES = {
    "reshape": _reshape, "relu": _same(), "sigmoid": _same(), "tanh01": _same(), "lrelu": _same(), "affine": _same(),
    "add": _same(), "scale_by_param": _same(), "softmax": _same(), "bias_add": _same(), "bn_train": _same(),
    "bn_infer": _same(), "spectral_normalize": _same(),
    "globalpool": lambda x, mean: (x.shape[0], x.shape[3]),
    "slice_cols": lambda x, lo, hi: (x.shape[0], hi - lo),
    "slice_rows": lambda x, lo, hi: (hi - lo,) + tuple(x.shape[1:]),
    "concat_cols": lambda a, b: (a.shape[0], a.shape[1] + b.shape[1]),
    "concat_rows": lambda a, b: (a.shape[0] + b.shape[0],) + tuple(a.shape[1:]),
    "rowdot": lambda a, b: (a.shape[0], 1),
    "maxpool2": lambda x: (x.shape[0], x.shape[1] // 2, x.shape[2] // 2, x.shape[3]),
    "avgpool2": lambda x: (x.shape[0], x.shape[1] // 2, x.shape[2] // 2, x.shape[3]),
    "unpool": lambda x: (x.shape[0], 2 * x.shape[1], 2 * x.shape[2], x.shape[3]),
    "matmul": _matmul, "bmm": _bmm, "conv2d": _conv2d, "deconv2d": _deconv2d,
    "one_hot": lambda labels, classes: (labels.shape[0], classes),
    "attention": lambda theta, phi, g: (theta.shape[0], theta.shape[1], g.shape[2]),
}


@contextlib.contextmanager
def traced_kernels():
  """Swaps the kernel-layer functions for recording stand-ins; any OTHER kernel function raises, so a definition that
  starts using a new op cannot slip through untraced."""
  tracer = Tracer()
  saved = {}
  FakeDT._next[0] = 0

  def fake(name, rule):
    def f(*args, **kwargs):
      out = FakeDT(rule(*args, **kwargs))
      tracer.record(name, args, kwargs, out)
      return out
    return f

  def forbidden(name):
    def f(*a, **k):
      raise AssertionError("kernels.%s is not covered by the trace harness" % name)
    return f

  keep = {"BNState", "same_pad", "conv_desc"}
  saved["attention_shape_ok"] = K.attention_shape_ok
  K.attention_shape_ok = lambda *a: False          # a host-side predicate (math_mode 0 in the traces), not a kernel call
  keep.add("attention_shape_ok")
  for name in dir(K):
    obj = getattr(K, name)
    if name.startswith("_") or not callable(obj) or isinstance(obj, type) or name in keep:
      continue
    saved[name] = obj
    setattr(K, name, fake(name, SHAPE_RULES[name]) if name in SHAPE_RULES else forbidden(name))

  def from_numpy(a, req=False):
    return FakeDT(np.asarray(a).shape, req)
  K.from_numpy = from_numpy

  orig_get = V.VariableStore.get

  def get(self, name, shape, initializer, trainable=True):
    full = self.full_name(name)
    new = full not in self.vars
    v = orig_get(self, name, shape, initializer, trainable)
    if new:
      digest = hashlib.sha1(np.ascontiguousarray(self.init_values[full]).tobytes()).hexdigest()[:12]
      tracer.variables.append([full, list(v.shape), bool(trainable), digest])
    tracer.ops.append(["get_variable", full, list(v.shape), ["T", v.id, list(v.shape)]])
    return v
  V.VariableStore.get = get
  try:
    yield tracer
  finally:
    V.VariableStore.get = orig_get
    for name, obj in saved.items():
      setattr(K, name, obj)


def trace_networks(gin_text, architecture, image_shape, batch=2, z_dim=128, num_classes=0, conditional=False, seed=0):
  """Builds G and D of `architecture` under `gin_text` exactly as ModularGAN does and traces: G(z, y, training),
  D([x; G(z)], y, training), G(z, y, inference).  Returns {"variables": [...], "ops": [...]}."""
  from compare_gan_b200 import datasets
  from compare_gan_b200.gans import modular_gan
  gin.clear_config()
  gin.parse_config(gin_text)
  ds = datasets.ImageDatasetV2("synthetic", image_shape[0], image_shape[2], num_classes or None, 100)
  params = {"architecture": architecture, "z_dim": z_dim, "lambda": 1, "disc_iters": 1, "seed": seed}
  with traced_kernels() as tracer:
    gan = modular_gan.ModularGAN(dataset=ds, parameters=params, model_dir="/tmp/arch_trace", conditional=conditional)
    store = V.VariableStore(seed=seed)
    z = FakeDT((batch, z_dim))
    x = FakeDT((batch,) + tuple(image_shape))
    y = FakeDT((batch, num_classes)) if conditional else None
    with V.use(store):
      fake = gan.generator(z, y=y, is_training=True)
      both = K.concat_rows(x, fake)
      yy = K.concat_rows(y, y) if conditional else None
      gan.discriminator(both, y=yy, is_training=True)
      gan.generator(z, y=y, is_training=False)
  gin.clear_config()
  return {"variables": tracer.variables, "ops": tracer.ops}


# The configurations whose traces are pinned in tests/golden/arch_traces.json: every architecture the engine builds,
# at small widths where the structure allows (the trace is shape-only, but initial values are really drawn).
CASES = {
    "resnet_cifar": dict(gin_text="G.batch_norm_fn = @batch_norm\nD.spectral_norm = True\nstandardize_batch.decay = 0.9\n"
                                  "standardize_batch.epsilon = 1e-5", architecture="resnet_cifar_arch",
                         image_shape=(32, 32, 3)),
    "resnet_cifar_gsn_projection": dict(gin_text="G.batch_norm_fn = @conditional_batch_norm\nG.spectral_norm = True\n"
                                                 "D.spectral_norm = True\nresnet_cifar.Discriminator.project_y = True",
                                        architecture="resnet_cifar_arch", image_shape=(32, 32, 3), num_classes=10,
                                        conditional=True),
    "sndcgan": dict(gin_text="G.batch_norm_fn = @batch_norm\nD.spectral_norm = True", architecture="sndcgan_arch",
                    image_shape=(32, 32, 3)),
    "sndcgan_128": dict(gin_text="G.batch_norm_fn = @batch_norm\nD.spectral_norm = True", architecture="sndcgan_arch",
                        image_shape=(128, 128, 3)),
    "dcgan": dict(gin_text="G.batch_norm_fn = @batch_norm\nD.batch_norm_fn = @batch_norm", architecture="dcgan_arch",
                  image_shape=(64, 64, 3)),
    "resnet5": dict(gin_text="G.batch_norm_fn = @batch_norm", architecture="resnet5_arch", image_shape=(128, 128, 3)),
    "resnet5_64": dict(gin_text="G.batch_norm_fn = @batch_norm", architecture="resnet5_arch", image_shape=(64, 64, 3)),
    "biggan_32": dict(gin_text="G.batch_norm_fn = @conditional_batch_norm\nG.spectral_norm = True\nD.spectral_norm = True\n"
                               "spectral_norm.singular_value = 'auto'\nweights.initializer = 'orthogonal'\n"
                               "standardize_batch.use_moving_averages = False\nresnet_biggan.Generator.ch = 8\n"
                               "resnet_biggan.Discriminator.ch = 8\nresnet_biggan.Discriminator.project_y = True\n"
                               "resnet_biggan.Generator.blocks_with_attention = 'B2'\n"
                               "resnet_biggan.Discriminator.blocks_with_attention = 'B1'",
                      architecture="resnet_biggan_arch", image_shape=(32, 32, 3), z_dim=120, num_classes=10, conditional=True),
    "biggan_128": dict(gin_text="G.batch_norm_fn = @conditional_batch_norm\nG.spectral_norm = True\nD.spectral_norm = True\n"
                                "spectral_norm.singular_value = 'auto'\nweights.initializer = 'orthogonal'\n"
                                "standardize_batch.use_moving_averages = False\nresnet_biggan.Generator.ch = 8\n"
                                "resnet_biggan.Discriminator.ch = 8\nresnet_biggan.Discriminator.project_y = True",
                       architecture="resnet_biggan_arch", image_shape=(128, 128, 3), z_dim=120, num_classes=1000,
                       conditional=True),
    "biggan_deep_64": dict(gin_text="G.batch_norm_fn = @conditional_batch_norm\nG.spectral_norm = True\nD.spectral_norm = True\n"
                                    "spectral_norm.singular_value = 'auto'\nweights.initializer = 'orthogonal'\n"
                                    "standardize_batch.use_moving_averages = False\nresnet_biggan_deep.Generator.ch = 4\n"
                                    "resnet_biggan_deep.Discriminator.ch = 4",
                           architecture="resnet_biggan_deep_arch", image_shape=(64, 64, 3), z_dim=128, num_classes=10,
                           conditional=True),
    "biggan_128_unconditional_plain": dict(gin_text="G.batch_norm_fn = @batch_norm\nresnet_biggan.Generator.ch = 8\n"
                                                    "resnet_biggan.Discriminator.ch = 8\n"
                                                    "resnet_biggan.Generator.hierarchical_z = False\n"
                                                    "resnet_biggan.Generator.embed_y = False\n"
                                                    "resnet_biggan.Discriminator.project_y = False",
                                           architecture="resnet_biggan_arch", image_shape=(64, 64, 3), z_dim=128),
}


def canonical_sha1(ops):
  import json
  return hashlib.sha1(json.dumps(ops, sort_keys=True, separators=(",", ":")).encode()).hexdigest()


def write_golden(path):
  """Regenerates tests/golden/arch_traces.json from the CURRENT definitions (only after an intended change)."""
  import json
  golden = {"_about": "Traces of the host-side network definitions recorded by tests/arch_trace.py (shape-only stand-ins "
                      "of the kernel layer). ops_sha1 = sha1 of the canonical JSON of the recorded op list; regenerate "
                      "with tests/arch_trace.py:write_golden() ONLY when a definition is changed on purpose."}
  for name, kw in CASES.items():
    tr = trace_networks(**kw)
    golden[name] = {"variables": tr["variables"], "n_ops": len(tr["ops"]), "ops_sha1": canonical_sha1(tr["ops"])}
  json.dump(golden, open(path, "w"), indent=0)
