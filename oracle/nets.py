"""Oracle (test infrastructure): variable store, arch_ops and the four BASELINE
architectures restated on PyTorch-CPU.  Variable names / creation order follow
the reference so the name+shape lists of architectures/resnet_norm_test.py are
reproduced (see tests/test_oracle_golden.py).
"""
import contextlib
from collections import OrderedDict

import numpy as np
import torch

from . import tf_ops as T


# test hook: callables fn(scope_name, tensor) receiving the output of every linear / conv2d / deconv2d / non_local_block
ACT_OBSERVERS = []


def _observe(store, y):
  if ACT_OBSERVERS:
    name = "/".join(store._scope)
    for fn in ACT_OBSERVERS:
      fn(name, y)
  return y


class Cfg(object):
  """The gin bindings that matter on the hot path."""

  def __init__(self, **kw):
    self.architecture = "resnet_cifar_arch"
    self.image_shape = (32, 32, 3)
    self.g_bn = "batch_norm"            # G.batch_norm_fn: None|batch_norm|conditional_batch_norm
    self.d_bn = None                    # D.batch_norm_fn
    self.g_sn = False                   # G.spectral_norm
    self.d_sn = False                   # D.spectral_norm
    self.sn_singular = "left"           # spectral_norm.singular_value
    self.sn_eps = 1e-12                 # spectral_norm.epsilon
    self.bn_decay = 0.999               # standardize_batch.decay
    self.bn_eps = 1e-3                  # standardize_batch.epsilon
    self.use_moving_averages = True     # standardize_batch.use_moving_averages
    self.initializer = "normal"         # weights.initializer
    self.stddev = 0.02                  # weights.stddev
    self.ch = 96                        # resnet_biggan.{Generator,Discriminator}.ch
    self.project_y = False              # Discriminator.project_y
    self.hierarchical_z = False
    self.embed_y = False
    self.embed_y_dim = 128
    self.num_classes = 0
    self.g_attention = "B4"
    self.d_attention = "B1"
    self.bn_replicas = 1                # simulate cross-replica moments over R batch shards
    for k, v in kw.items():
      if not hasattr(self, k):
        raise ValueError("unknown cfg key %s" % k)
      setattr(self, k, v)


class VarStore(object):
  """tf.get_variable with AUTO_REUSE over a flat ordered name->tensor dict."""

  def __init__(self, seed=0, dtype=torch.float32):
    self.dtype = dtype               # float64 gives the "exact" gradients used to calibrate fp32 tolerances
    self.vars = OrderedDict()
    self.trainable = OrderedDict()
    self._scope = []
    self.rng = np.random.RandomState(seed)

  @contextlib.contextmanager
  def scope(self, name):
    self._scope.append(name)
    try:
      yield
    finally:
      self._scope.pop()

  def full(self, name):
    return "/".join(self._scope + [name])

  def get(self, name, shape, init, trainable=True):
    full = self.full(name)
    if full in self.vars:
      v = self.vars[full]
      if tuple(v.shape) != tuple(shape):
        raise ValueError("shape mismatch for %s: %s vs %s" % (full, tuple(v.shape), shape))
      return v
    kind = init[0]
    if kind == "normal":
      a = (self.rng.standard_normal(shape) * init[1]).astype(np.float32)
    elif kind == "zeros":
      a = np.zeros(shape, np.float32)
    elif kind == "ones":
      a = np.ones(shape, np.float32)
    elif kind == "const":
      a = np.full(shape, init[1], np.float32)
    elif kind == "orthogonal":
      a = T.orthogonal_init(self.rng, shape)
    elif kind == "glorot_normal":
      a = T.glorot_normal_init(self.rng, shape)
    else:
      raise ValueError(kind)
    v = torch.from_numpy(np.ascontiguousarray(a).reshape(tuple(shape))).to(self.dtype)
    if trainable:
      v.requires_grad_(True)
      self.trainable[full] = v
    self.vars[full] = v
    return v

  def names(self, prefix="", trainable_only=False):
    src = self.trainable if trainable_only else self.vars
    return [(k, list(v.shape)) for k, v in src.items() if k.startswith(prefix)]

  def trainable_under(self, prefix):
    return OrderedDict((k, v) for k, v in self.trainable.items() if k.startswith(prefix + "/"))

  def state_numpy(self):
    return OrderedDict((k, v.detach().numpy().copy()) for k, v in self.vars.items())

  def load_numpy(self, state):
    with torch.no_grad():
      for k, a in state.items():
        if k in self.vars:
          self.vars[k].copy_(torch.from_numpy(np.asarray(a, np.float32).reshape(self.vars[k].shape)).to(self.dtype))


# --------------------------------------------------------------------------- arch_ops

def weight_init(cfg, stddev=None):
  """arch_ops.py:46-63."""
  if cfg.initializer == "normal":
    return ("normal", cfg.stddev if stddev is None else stddev)
  if cfg.initializer == "orthogonal":
    return ("orthogonal",)
  raise ValueError("Unknown weight initializer {}.".format(cfg.initializer))


def spectral_norm(store, cfg, w, var_name="kernel"):
  """arch_ops.py:453-535.  ``u_var`` lives at <kernel>/u_var, advances on every call."""
  if w.dim() < 2:
    raise ValueError("Spectral norm can only be applied to multi-dimensional tensors")
  w2 = w.reshape(-1, w.shape[-1])
  sv = cfg.sn_singular
  if sv == "auto":
    sv = "left" if w2.shape[0] <= w2.shape[1] else "right"
  u_shape = (w2.shape[0], 1) if sv == "left" else (1, w2.shape[1])
  u = store.get(var_name + "/u_var", u_shape, ("normal", 1.0), trainable=False)
  sigma, u_new, _ = T.spectral_sigma(w2, u, sv, cfg.sn_eps)
  with torch.no_grad():
    u.copy_(u_new)
  return (w2 / sigma).reshape(w.shape)


def linear(store, cfg, x, out, scope, use_sn=False, use_bias=True, bias_start=0.0, stddev=0.02):
  """arch_ops.py:538-556."""
  with store.scope(scope or "linear"):
    k = store.get("kernel", (x.shape[1], out), weight_init(cfg, stddev))
    if use_sn:
      k = spectral_norm(store, cfg, k)
    y = x @ k
    if use_bias:
      y = y + store.get("bias", (out,), ("const", bias_start))
    return _observe(store, y)


def conv2d(store, cfg, x, out, kh, kw, stride, name, use_sn=False, use_bias=True):
  """arch_ops.py:559-573."""
  with store.scope(name):
    w = store.get("kernel", (kh, kw, x.shape[-1], out), weight_init(cfg))
    if use_sn:
      w = spectral_norm(store, cfg, w)
    y = T.conv2d_same(x, w, stride)
    if use_bias:
      y = y + store.get("bias", (out,), ("zeros",))
    return _observe(store, y)


def deconv2d(store, cfg, x, out_shape, kh, kw, stride, name, use_sn=False):
  """arch_ops.py:579-592 (bias always)."""
  with store.scope(name):
    w = store.get("kernel", (kh, kw, out_shape[-1], x.shape[-1]), weight_init(cfg))
    if use_sn:
      w = spectral_norm(store, cfg, w)
    y = T.conv2d_transpose_same(x, w, (out_shape[1], out_shape[2]), stride)
    return _observe(store, y + store.get("bias", (out_shape[-1],), ("zeros",)))


def standardize_batch(store, cfg, x, is_training):
  """arch_ops.py:194-319 (+ :66-119 moving averages, :122-191 accumulators)."""
  if x.dim() not in (2, 4):
    raise ValueError("Inputs has unsupported rank. Expected 2 or 4 but got %d" % x.dim())
  shape = x.shape
  c = shape[-1]
  x4 = x.reshape(-1, 1, 1, c) if x.dim() == 2 else x
  if cfg.bn_replicas > 1:
    mean, var = T.cross_replica_moments(list(torch.chunk(x4, cfg.bn_replicas, dim=0)))
  else:
    mean, var = T.batch_moments(x4)
  if cfg.use_moving_averages:
    mm = store.get("moving_mean", (c,), ("zeros",), trainable=False)
    mv = store.get("moving_variance", (c,), ("ones",), trainable=False)
    if is_training:
      with torch.no_grad():   # assign_moving_average, zero_debias=False
        mm.sub_((mm - mean) * (1.0 - cfg.bn_decay))
        mv.sub_((mv - var) * (1.0 - cfg.bn_decay))
    else:
      mean, var = mm, mv
  else:
    with store.scope("accu"):
      am = store.get("accu_mean", (c,), ("zeros",), trainable=False)
      av = store.get("accu_variance", (c,), ("zeros",), trainable=False)
      ac = store.get("accu_counter", (), ("const", 1e-12), trainable=False)
      ua = store.get("update_accus", (), ("zeros",), trainable=False)
    if not is_training:
      if float(ua) == 1.0:
        with torch.no_grad():
          am.add_(mean)
          av.add_(var)
          ac.add_(1.0)
      mean, var = am / ac, av / ac
  y = T.normalize(x4, mean, var, cfg.bn_eps)
  return y.reshape(shape)


def batch_norm(store, cfg, x, is_training, name="batch_norm", center=True, scale=True):
  """arch_ops.py:327-367."""
  with store.scope(name):
    y = standardize_batch(store, cfg, x, is_training)
    c = x.shape[-1]
    if scale:
      y = y * store.get("gamma", (c,), ("ones",))
    if center:
      y = y + store.get("beta", (c,), ("zeros",))
    return y


def conditional_batch_norm(store, cfg, x, y, is_training, use_sn, name="batch_norm"):
  """arch_ops.py:423-445 (gamma multiplies directly, no +1; use_bias False)."""
  if y is None:
    raise ValueError("You must provide y for conditional batch normalization.")
  if y.dim() != 2:
    raise ValueError("Conditioning must have rank 2.")
  with store.scope(name):
    out = standardize_batch(store, cfg, x, is_training)
    c = x.shape[-1]
    with store.scope("condition"):
      gamma = linear(store, cfg, y, c, "gamma", use_sn=use_sn, use_bias=False)
      out = out * gamma.reshape(-1, 1, 1, c)
      beta = linear(store, cfg, y, c, "beta", use_sn=use_sn, use_bias=False)
      out = out + beta.reshape(-1, 1, 1, c)
    return out


def apply_bn(store, cfg, which, x, y, is_training, name, use_sn):
  """AbstractGenerator.batch_norm dispatch — abstract_arch.py:76-83."""
  if which is None:
    return x
  if which == "batch_norm":
    return batch_norm(store, cfg, x, is_training, name=name)
  if which == "conditional_batch_norm":
    return conditional_batch_norm(store, cfg, x, y, is_training, use_sn, name=name)
  raise ValueError(which)


def non_local_block(store, cfg, x, name, use_sn):
  """arch_ops.py:709-758."""
  with store.scope(name):
    n, h, w, c = x.shape
    ca, cg = c // 8, c // 2
    theta = conv2d(store, cfg, x, ca, 1, 1, 1, "conv2d_theta", use_sn, use_bias=False)
    theta = theta.reshape(n, h * w, ca)
    phi = conv2d(store, cfg, x, ca, 1, 1, 1, "conv2d_phi", use_sn, use_bias=False)
    phi = T.max_pool2(phi).reshape(n, h * w // 4, ca)
    attn = torch.softmax(T.bmm(theta, phi, False, True), dim=-1)
    g = conv2d(store, cfg, x, cg, 1, 1, 1, "conv2d_g", use_sn, use_bias=False)
    g = T.max_pool2(g).reshape(n, h * w // 4, cg)
    attn_g = T.bmm(attn, g).reshape(n, h, w, cg)
    sigma = store.get("sigma", (), ("zeros",))
    attn_g = conv2d(store, cfg, attn_g, c, 1, 1, 1, "conv2d_attn_g", use_sn, use_bias=False)
    return _observe(store, x + sigma * attn_g)


# --------------------------------------------------------------------------- resnet_ops

def _get_conv(store, cfg, x, cin, cout, scale, suffix, use_sn, ksize=3):
  """resnet_ops.py:104-134."""
  if x.shape[-1] != cin:
    raise ValueError("Unexpected number of input channels.")
  if scale not in ("up", "down", "none"):
    raise ValueError("Scale: got {}, expected 'up', 'down', or 'none'.".format(scale))
  h = T.unpool(x) if scale == "up" else x
  h = conv2d(store, cfg, h, cout, ksize, ksize, 1,
             "{}_{}".format("same" if scale == "none" else scale, suffix), use_sn)
  if scale == "down":
    h = T.avg_pool2(h)
  return h


def resnet_block(store, cfg, x, name, cin, cout, scale, is_gen, y, is_training, bn, use_sn):
  """resnet_ops.ResNetBlock.apply — resnet_ops.py:136-182."""
  scale1 = scale if is_gen else "none"
  scale2 = "none" if is_gen else scale
  with store.scope(name):
    shortcut = _get_conv(store, cfg, x, cin, cout, scale, "conv_shortcut", use_sn)
    h = apply_bn(store, cfg, bn, x, y, is_training, "bn1", use_sn)
    h = torch.relu(h)
    h = _get_conv(store, cfg, h, cin, cout, scale1, "conv1", use_sn)
    h = apply_bn(store, cfg, bn, h, y, is_training, "bn2", use_sn)
    h = torch.relu(h)
    h = _get_conv(store, cfg, h, cout, cout, scale2, "conv2", use_sn)
    return _observe(store, h + shortcut)


def biggan_block(store, cfg, x, name, cin, cout, scale, is_gen, y, is_training, bn, use_sn,
                 add_shortcut=True):
  """resnet_biggan.BigGanResNetBlock.apply — resnet_biggan.py:99-151."""
  scale1 = scale if is_gen else "none"
  scale2 = "none" if is_gen else scale
  with store.scope(name):
    h = apply_bn(store, cfg, bn, x, y, is_training, "bn1", use_sn)
    h = torch.relu(h)
    h = _get_conv(store, cfg, h, cin, cout, scale1, "conv1", use_sn)
    h = apply_bn(store, cfg, bn, h, y, is_training, "bn2", use_sn)
    h = torch.relu(h)
    h = _get_conv(store, cfg, h, cout, cout, scale2, "conv2", use_sn)
    if add_shortcut:
      h = h + _get_conv(store, cfg, x, cin, cout, scale, "conv_shortcut", use_sn, ksize=1)
    return _observe(store, h)


# --------------------------------------------------------------------------- architectures

def _gen_resnet_cifar(store, cfg, z, y, is_training):
  """resnet_cifar.Generator.apply — resnet_cifar.py:58-112."""
  sn, bn = cfg.g_sn, cfg.g_bn
  assert cfg.image_shape[0] == 32 and cfg.image_shape[1] == 32
  h = linear(store, cfg, z, 4 * 4 * 256, "fc_noise", use_sn=sn)
  h = h.reshape(-1, 4, 4, 256)
  for i in range(3):
    h = resnet_block(store, cfg, h, "B%d" % (i + 1), 256, 256, "up", True, y, is_training, bn, sn)
  h = apply_bn(store, cfg, bn, h, y, is_training, "final_norm", sn)
  h = torch.relu(h)
  h = conv2d(store, cfg, h, cfg.image_shape[2], 3, 3, 1, "final_conv", use_sn=sn)
  return torch.sigmoid(h)


def _disc_resnet_cifar(store, cfg, x, y, is_training):
  """resnet_cifar.Discriminator.apply — resnet_cifar.py:123-167."""
  sn, bn = cfg.d_sn, cfg.d_bn
  colors = x.shape[3]
  if colors not in (1, 3):
    raise ValueError("Number of color channels not supported: {}".format(colors))
  h = x
  for i in range(4):
    h = resnet_block(store, cfg, h, "B%d" % (i + 1), colors if i == 0 else 128, 128,
                     "down" if i <= 1 else "none", False, y, is_training, bn, sn)
  h = torch.relu(h)
  feat = h.mean(dim=(1, 2))
  logit = linear(store, cfg, feat, 1, "disc_final_fc", use_sn=sn)
  if cfg.project_y:
    if y is None:
      raise ValueError("You must provide class information y to project.")
    emb = linear(store, cfg, y, 128, "embedding_fc", use_sn=sn, use_bias=False)
    logit = logit + (emb * feat).sum(dim=1, keepdim=True)
  return torch.sigmoid(logit), logit, feat


def _gen_sndcgan(store, cfg, z, y, is_training):
  """sndcgan.Generator.apply — sndcgan.py:42-79 (no SN in G: plain linear/deconv2d)."""
  bn = cfg.g_bn
  b = z.shape[0]
  sh, sw, colors = cfg.image_shape
  c2 = lambda s: -(-s // 2)
  sh2, sw2 = c2(sh), c2(sw)
  sh4, sw4 = c2(sh2), c2(sw2)
  sh8, sw8 = c2(sh4), c2(sw4)
  h = linear(store, cfg, z, sh8 * sw8 * 512, "g_fc1")
  h = apply_bn(store, cfg, bn, h, y, is_training, "g_bn1", False)
  h = torch.relu(h).reshape(b, sh8, sw8, 512)
  h = deconv2d(store, cfg, h, (b, sh4, sw4, 256), 4, 4, 2, "g_dc2")
  h = torch.relu(apply_bn(store, cfg, bn, h, y, is_training, "g_bn2", False))
  h = deconv2d(store, cfg, h, (b, sh2, sw2, 128), 4, 4, 2, "g_dc3")
  h = torch.relu(apply_bn(store, cfg, bn, h, y, is_training, "g_bn3", False))
  h = deconv2d(store, cfg, h, (b, sh, sw, 64), 4, 4, 2, "g_dc4")
  h = torch.relu(apply_bn(store, cfg, bn, h, y, is_training, "g_bn4", False))
  h = deconv2d(store, cfg, h, (b, sh, sw, colors), 3, 3, 1, "g_dc5")
  return (torch.tanh(h) + 1.0) / 2.0


def _disc_sndcgan(store, cfg, x, y, is_training):
  """sndcgan.Discriminator.apply — sndcgan.py:88-127."""
  sn = cfg.d_sn
  h = x * 2.0 - 1.0
  spec = [(64, 3, 1), (128, 4, 2), (128, 3, 1), (256, 4, 2), (256, 3, 1), (512, 4, 2), (512, 3, 1)]
  for i, (co, k, s) in enumerate(spec):
    h = conv2d(store, cfg, h, co, k, k, s, "d_conv%d" % (i + 1), use_sn=sn)
    h = T.lrelu(h, 0.1)
  feat = h.reshape(x.shape[0], -1)
  logit = linear(store, cfg, feat, 1, "d_fc1", use_sn=sn)
  return torch.sigmoid(logit), logit, feat


def _gen_resnet5(store, cfg, z, y, is_training, ch=64, channels=(8, 8, 4, 4, 2, 1)):
  """resnet5.Generator.apply — resnet5.py:45-93 (fc_noise without SN kwarg)."""
  sn, bn = cfg.g_sn, cfg.g_bn
  seed = 4
  size = cfg.image_shape[0]
  up_layers = np.log2(float(size) / seed)
  if not float(up_layers).is_integer():
    raise ValueError("log2({}/{}) must be an integer.".format(size, seed))
  if up_layers < 0 or up_layers > 5:
    raise ValueError("Invalid image_size {}.".format(size))
  up_layers = int(up_layers)
  h = linear(store, cfg, z, ch * channels[0] * seed * seed, "fc_noise")
  h = h.reshape(-1, seed, seed, ch * channels[0])
  for i in range(5):
    h = resnet_block(store, cfg, h, "B%d" % (i + 1), ch * channels[i], ch * channels[i + 1],
                     "up" if i < up_layers else "none", True, y, is_training, bn, sn)
  h = torch.relu(apply_bn(store, cfg, bn, h, y, is_training, "final_norm", sn))
  h = conv2d(store, cfg, h, cfg.image_shape[2], 3, 3, 1, "final_conv")
  return torch.sigmoid(h)


def _disc_resnet5(store, cfg, x, y, is_training, ch=64, channels=(1, 2, 4, 4, 8, 8)):
  """resnet5.Discriminator.apply — resnet5.py:105-145."""
  sn, bn = cfg.d_sn, cfg.d_bn
  colors = x.shape[3]
  if colors not in (1, 3):
    raise ValueError("Number of color channels not supported: {}".format(colors))
  h = resnet_block(store, cfg, x, "B0", colors, ch, "down", False, y, is_training, bn, sn)
  for i in range(5):
    h = resnet_block(store, cfg, h, "B%d" % (i + 1), ch * channels[i], ch * channels[i + 1],
                     "down", False, y, is_training, bn, sn)
  h = torch.relu(h)
  feat = h.mean(dim=(1, 2))
  logit = linear(store, cfg, feat, 1, "disc_final_fc", use_sn=sn)
  return torch.sigmoid(logit), logit, feat


def _gen_dcgan(store, cfg, z, y, is_training):
  """dcgan.Generator.apply — dcgan.py:39-84 (5x5 stride-2 transposed convs, asymmetric SAME padding)."""
  bn = cfg.g_bn
  b = z.shape[0]
  sh, sw, colors = cfg.image_shape
  c2 = lambda s: -(-s // 2)
  sh2, sw2 = c2(sh), c2(sw); sh4, sw4 = c2(sh2), c2(sw2); sh8, sw8 = c2(sh4), c2(sw4); sh16, sw16 = c2(sh8), c2(sw8)
  h = linear(store, cfg, z, 512 * sh16 * sw16, "g_fc1").reshape(-1, sh16, sw16, 512)
  h = torch.relu(apply_bn(store, cfg, bn, h, y, is_training, "g_bn1", False))
  h = deconv2d(store, cfg, h, (b, sh8, sw8, 256), 5, 5, 2, "g_dc1")
  h = torch.relu(apply_bn(store, cfg, bn, h, y, is_training, "g_bn2", False))
  h = deconv2d(store, cfg, h, (b, sh4, sw4, 128), 5, 5, 2, "g_dc2")
  h = torch.relu(apply_bn(store, cfg, bn, h, y, is_training, "g_bn3", False))
  h = deconv2d(store, cfg, h, (b, sh2, sw2, 64), 5, 5, 2, "g_dc3")
  h = torch.relu(apply_bn(store, cfg, bn, h, y, is_training, "g_bn4", False))
  h = deconv2d(store, cfg, h, (b, sh, sw, colors), 5, 5, 2, "g_dc4")
  return 0.5 * torch.tanh(h) + 0.5


def _disc_dcgan(store, cfg, x, y, is_training):
  """dcgan.Discriminator.apply — dcgan.py:87-129."""
  sn, bn = cfg.d_sn, cfg.d_bn
  h = T.lrelu(conv2d(store, cfg, x, 64, 5, 5, 2, "d_conv1", use_sn=sn))
  for i, co in enumerate((128, 256, 512)):
    h = conv2d(store, cfg, h, co, 5, 5, 2, "d_conv%d" % (i + 2), use_sn=sn)
    h = T.lrelu(apply_bn(store, cfg, bn, h, y, is_training, "d_bn%d" % (i + 1), sn))
  feat = h
  logit = linear(store, cfg, h.reshape(x.shape[0], -1), 1, "d_fc4", use_sn=sn)
  return torch.sigmoid(logit), logit, feat


_BIGGAN_G = {512: [16, 16, 8, 8, 4, 2, 1, 1], 256: [16, 16, 8, 8, 4, 2, 1],
             128: [16, 16, 8, 4, 2, 1], 64: [16, 16, 8, 4, 2], 32: [4, 4, 4, 4]}
_BIGGAN_D = {512: [1, 1, 2, 4, 8, 8, 16, 16], 256: [1, 2, 4, 8, 8, 16, 16],
             128: [1, 2, 4, 8, 16, 16], 64: [2, 4, 8, 16, 16], 32: [2, 2, 2, 2]}


def _gen_biggan(store, cfg, z, y, is_training):
  """resnet_biggan.Generator.apply — resnet_biggan.py:223-302."""
  sn, bn = cfg.g_sn, cfg.g_bn
  res = cfg.image_shape[0]
  if res not in _BIGGAN_G:
    raise ValueError("Unsupported resolution: {}".format(res))
  mult = _BIGGAN_G[res]
  cin = [cfg.ch * c for c in mult[:-1]]
  cout = [cfg.ch * c for c in mult[1:]]
  nb = len(cin)
  if cfg.embed_y:
    y = linear(store, cfg, y, cfg.embed_y_dim, "embed_y", use_sn=False, use_bias=False)
  y_per_block = nb * [y]
  if cfg.hierarchical_z:
    zs = torch.chunk(z, nb + 1, dim=1)
    z0, z_per_block = zs[0], zs[1:]
    if y is not None:
      y_per_block = [torch.cat([zi, y], 1) for zi in z_per_block]
  else:
    z0 = z
  h = linear(store, cfg, z0, cin[0] * 16, "fc_noise", use_sn=sn)
  h = h.reshape(-1, 4, 4, cin[0])
  attn = set(cfg.g_attention.split(","))
  for i in range(nb):
    name = "B%d" % (i + 1)
    h = biggan_block(store, cfg, h, name, cin[i], cout[i], "up", True, y_per_block[i],
                     is_training, bn, sn)
    if name in attn:
      h = non_local_block(store, cfg, h, "non_local_block", sn)
  h = batch_norm(store, cfg, h, is_training, name="final_norm")
  h = torch.relu(h)
  h = conv2d(store, cfg, h, cfg.image_shape[2], 3, 3, 1, "final_conv", use_sn=sn)
  return (torch.tanh(h) + 1.0) / 2.0


def _disc_biggan(store, cfg, x, y, is_training):
  """resnet_biggan.Discriminator.apply — resnet_biggan.py:363-425."""
  sn, bn = cfg.d_sn, cfg.d_bn
  colors, res = x.shape[-1], x.shape[1]
  if colors not in (1, 3):
    raise ValueError("Unsupported color channels: {}".format(colors))
  if res not in _BIGGAN_D:
    raise ValueError("Unsupported resolution: {}".format(res))
  cout = [cfg.ch * c for c in _BIGGAN_D[res]]
  cin = [colors] + cout[:-1]
  attn = set(cfg.d_attention.split(","))
  h = x
  nb = len(cin)
  for i in range(nb):
    name = "B%d" % (i + 1)
    last = i == nb - 1
    h = biggan_block(store, cfg, h, name, cin[i], cout[i], "none" if last else "down", False, y,
                     is_training, bn, sn, add_shortcut=cin[i] != cout[i])
    if name in attn:
      h = non_local_block(store, cfg, h, "non_local_block", sn)
  h = torch.relu(h)
  feat = h.sum(dim=(1, 2))
  logit = linear(store, cfg, feat, 1, "final_fc", use_sn=sn)
  if cfg.project_y:
    if y is None:
      raise ValueError("You must provide class information y to project.")
    with store.scope("embedding_fc"):
      k = store.get("kernel", (y.shape[1], cout[-1]), ("glorot_normal",))
      if sn:
        k = spectral_norm(store, cfg, k)
      emb = y @ k
    logit = logit + (emb * feat).sum(dim=1, keepdim=True)
  return torch.sigmoid(logit), logit, feat


# --------------------------------------------------------------------------- resnet_biggan_deep

def biggan_deep_block(store, cfg, x, name, cin, cout, scale, y, is_training, bn, use_sn):
  """resnet_biggan_deep.BigGanDeepResNetBlock — resnet_biggan_deep.py:61-177: bottleneck (1x1 -> 3x3 -> 3x3 -> 1x1 at
  max(cin, cout) / 4 channels) with an identity-preserving skip: channels are DROPPED on the way up and ADDED by a 1x1
  conv on the way down."""
  if x.shape[-1] != cin:
    raise ValueError("Unexpected number of input channels (expected {}, got {}).".format(cin, x.shape[-1]))
  mid = max(cin, cout) // 4
  with store.scope(name):
    h = x
    with store.scope("conv1"):
      h = torch.relu(apply_bn(store, cfg, bn, h, y, is_training, "bn", use_sn))
      h = conv2d(store, cfg, h, mid, 1, 1, 1, "1x1_conv", use_sn)
    with store.scope("conv2"):
      h = torch.relu(apply_bn(store, cfg, bn, h, y, is_training, "bn", use_sn))
      if scale == "up":
        h = T.unpool(h)
      h = conv2d(store, cfg, h, mid, 3, 3, 1, "3x3_conv", use_sn)
    with store.scope("conv3"):
      h = torch.relu(apply_bn(store, cfg, bn, h, y, is_training, "bn", use_sn))
      h = conv2d(store, cfg, h, mid, 3, 3, 1, "3x3_conv", use_sn)
    with store.scope("conv4"):
      h = torch.relu(apply_bn(store, cfg, bn, h, y, is_training, "bn", use_sn))
      if scale == "down":
        h = T.avg_pool2(h)
      h = conv2d(store, cfg, h, cout, 1, 1, 1, "1x1_conv", use_sn)
    with store.scope("shortcut"):                                   # :94-117
      sc = x
      if cin > cout:
        assert scale == "up"
        sc = sc[..., :cout]
      if scale == "up":
        sc = T.unpool(sc)
      if scale == "down":
        sc = T.avg_pool2(sc)
      if cin < cout:
        assert scale == "down"
        sc = torch.cat([sc, conv2d(store, cfg, sc, cout - cin, 1, 1, 1, "add_channels", use_sn)], dim=-1)
    return h + sc


_DEEP_G = {512: 4 * [16] + 4 * [8] + [4, 4, 2, 2, 1, 1, 1], 256: 4 * [16] + 4 * [8] + [4, 4, 2, 2, 1],
           128: 4 * [16] + 2 * [8] + [4, 4, 2, 2, 1], 64: 4 * [16] + 2 * [8] + [4, 4, 2], 32: 8 * [4]}
_DEEP_D = {512: [1, 1, 1, 2, 2, 4, 4] + 4 * [8] + 4 * [16], 256: [1, 2, 2, 4, 4] + 4 * [8] + 4 * [16],
           128: [1, 2, 2, 4, 4] + 2 * [8] + 4 * [16], 64: [2, 4, 4] + 2 * [8] + 4 * [16], 32: 8 * [2]}


def _gen_biggan_deep(store, cfg, z, y, is_training):
  """resnet_biggan_deep.Generator.apply — resnet_biggan_deep.py:243-311: z is not chunked, every BN sees [z, embed(y)];
  blocks alternate none / up; attention after the up block that reaches 64x64."""
  sn, bn = cfg.g_sn, cfg.g_bn
  res = cfg.image_shape[0]
  if res not in _DEEP_G:
    raise ValueError("Unsupported resolution: {}".format(res))
  mult = _DEEP_G[res]
  cin = [cfg.ch * c for c in mult[:-1]]
  cout = [cfg.ch * c for c in mult[1:]]
  if cfg.embed_y:
    y = linear(store, cfg, y, cfg.embed_y_dim, "embed_y", use_sn=False, use_bias=False)
  if y is not None:
    y = torch.cat([z, y], dim=1)
    z = y
  h = linear(store, cfg, z, cin[0] * 16, "fc_noise", use_sn=sn).reshape(-1, 4, 4, cin[0])
  for i in range(len(cin)):
    scale = "none" if i % 2 == 0 else "up"
    h = biggan_deep_block(store, cfg, h, "B%d" % (i + 1), cin[i], cout[i], scale, y, is_training, bn, sn)
    if scale == "up" and h.shape[1] == 64:
      h = non_local_block(store, cfg, h, "non_local_block", sn)
  h = torch.relu(batch_norm(store, cfg, h, is_training, name="final_norm"))
  h = conv2d(store, cfg, h, cfg.image_shape[2], 3, 3, 1, "final_conv", use_sn=sn)
  return (torch.tanh(h) + 1.0) / 2.0


def _disc_biggan_deep(store, cfg, x, y, is_training):
  """resnet_biggan_deep.Discriminator.apply — resnet_biggan_deep.py:373-434: initial 3x3 conv, blocks alternate
  down / none, attention after the none block at 64x64, relu, SUM over space, final_fc (+ projection)."""
  sn, bn = cfg.d_sn, cfg.d_bn
  colors, res = x.shape[-1], x.shape[1]
  if colors not in (1, 3):
    raise ValueError("Unsupported color channels: {}".format(colors))
  if res not in _DEEP_D:
    raise ValueError("Unsupported resolution: {}".format(res))
  mult = _DEEP_D[res]
  cin = [cfg.ch * c for c in mult[:-1]]
  cout = [cfg.ch * c for c in mult[1:]]
  h = conv2d(store, cfg, x, cin[0], 3, 3, 1, "initial_conv", use_sn=sn)
  for i in range(len(cin)):
    scale = "down" if i % 2 == 0 else "none"
    h = biggan_deep_block(store, cfg, h, "B%d" % (i + 1), cin[i], cout[i], scale, y, is_training, bn, sn)
    if scale == "none" and h.shape[1] == 64:
      h = non_local_block(store, cfg, h, "non_local_block", sn)
  h = torch.relu(h)
  feat = h.sum(dim=(1, 2))
  logit = linear(store, cfg, feat, 1, "final_fc", use_sn=sn)
  if cfg.project_y:
    if y is None:
      raise ValueError("You must provide class information y to project.")
    with store.scope("embedding_fc"):
      k = store.get("kernel", (y.shape[1], cout[-1]), ("glorot_normal",))
      if sn:
        k = spectral_norm(store, cfg, k)
      emb = y @ k
    logit = logit + (emb * feat).sum(dim=1, keepdim=True)
  return torch.sigmoid(logit), logit, feat


_GENS = {"dcgan_arch": _gen_dcgan, "resnet_cifar_arch": _gen_resnet_cifar, "sndcgan_arch": _gen_sndcgan,
         "resnet5_arch": _gen_resnet5, "resnet_biggan_arch": _gen_biggan, "resnet_biggan_deep_arch": _gen_biggan_deep}
_DISCS = {"dcgan_arch": _disc_dcgan, "resnet_cifar_arch": _disc_resnet_cifar, "sndcgan_arch": _disc_sndcgan,
          "resnet5_arch": _disc_resnet5, "resnet_biggan_arch": _disc_biggan, "resnet_biggan_deep_arch": _disc_biggan_deep}


def generator(store, cfg, z, y, is_training):
  """AbstractGenerator.__call__ — abstract_arch.py:71-74 (scope "generator")."""
  if cfg.architecture not in _GENS:
    raise NotImplementedError("Architecture %s not implemented." % cfg.architecture)
  with store.scope("generator"):
    return _GENS[cfg.architecture](store, cfg, z, y, is_training)


def discriminator(store, cfg, x, y, is_training):
  """AbstractDiscriminator.__call__ — abstract_arch.py:116-119 (scope "discriminator")."""
  if cfg.architecture not in _DISCS:
    raise NotImplementedError("Architecture %s not implemented." % cfg.architecture)
  with store.scope("discriminator"):
    return _DISCS[cfg.architecture](store, cfg, x, y, is_training)
