"""Oracle (test infrastructure): TF1 op semantics restated on PyTorch-CPU fp32.

All activations are NHWC ``torch.float32`` tensors; conv kernels are HWIO
``[kh,kw,cin,cout]``; deconv kernels ``[kh,kw,cout,cin]``; linear kernels
``[in,out]`` — exactly the reference layouts (architectures/arch_ops.py:543-546,
563-565, 583-585).  Backward passes come from torch.autograd (an independent
implementation from the hand-written CUDA backward under test).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def _same_pads(n, k, s):
  """TF SAME padding: out=ceil(n/s); before=total//2."""
  out = -(-n // s)
  total = max((out - 1) * s + k - n, 0)
  return out, total // 2, total - total // 2


def conv2d_same(x, w_hwio, stride=1):
  """tf.nn.conv2d(x, w, strides=[1,s,s,1], padding="SAME") — arch_ops.py:568."""
  kh, kw = w_hwio.shape[0], w_hwio.shape[1]
  _, pt, pb = _same_pads(x.shape[1], kh, stride)
  _, pl, pr = _same_pads(x.shape[2], kw, stride)
  xn = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
  y = F.conv2d(xn, w_hwio.permute(3, 2, 0, 1), stride=stride)
  return y.permute(0, 2, 3, 1)


def conv2d_transpose_same(x, w_hwoi, out_hw, stride):
  """tf.nn.conv2d_transpose (default SAME) — arch_ops.py:588-589.

  Adjoint of "SAME-pad then VALID conv": full transposed conv, then crop the
  SAME padding.  ``w_hwoi`` is ``[kh,kw,cout,cin]`` (cin = x channels).
  """
  kh, kw = w_hwoi.shape[0], w_hwoi.shape[1]
  oh, ow = out_hw
  _, pt, _ = _same_pads(oh, kh, stride)
  _, pl, _ = _same_pads(ow, kw, stride)
  # torch conv_transpose2d weight: [in_channels(=cin of x), out_channels, kh, kw]
  wt = w_hwoi.permute(3, 2, 0, 1)
  full = F.conv_transpose2d(x.permute(0, 3, 1, 2), wt, stride=stride)
  need_h, need_w = pt + oh, pl + ow
  if full.shape[2] < need_h or full.shape[3] < need_w:
    full = F.pad(full, (0, max(0, need_w - full.shape[3]),
                        0, max(0, need_h - full.shape[2])))
  y = full[:, :, pt:pt + oh, pl:pl + ow]
  return y.permute(0, 2, 3, 1)


def unpool(x):
  """Zero-insertion 2x upsampling — resnet_ops.py:35-56 (value at even r,c)."""
  n, h, w, c = x.shape
  out = torch.zeros(n, 2 * h, 2 * w, c, dtype=x.dtype)
  out[:, ::2, ::2, :] = x
  return out


def avg_pool2(x):
  """tf.nn.pool(x,[2,2],"AVG","SAME",strides=[2,2]) — resnet_ops.py:131-133."""
  return F.avg_pool2d(x.permute(0, 3, 1, 2), 2, 2, ceil_mode=True,
                      count_include_pad=False).permute(0, 2, 3, 1)


def max_pool2(x):
  """tf.layers.max_pooling2d(pool 2, stride 2, VALID) — arch_ops.py:741,750."""
  return F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)


def lrelu(x, leak=0.2):
  """arch_ops.py:595-597: max(x, leak*x)."""
  return torch.maximum(x, leak * x)


def l2_normalize(x, eps=1e-12):
  """tf.math.l2_normalize over all elements: x*rsqrt(max(sum x^2, eps))."""
  return x * torch.rsqrt(torch.clamp((x * x).sum(), min=eps))


def batch_moments(x4):
  """arch_ops.py:289-297: fp32 mean and BIASED variance, E[x^2]-E[x]^2 form."""
  mean = x4.mean(dim=(0, 1, 2))
  mean_sq = (x4 * x4).mean(dim=(0, 1, 2))
  return mean, mean_sq - mean * mean


def cross_replica_mean(shards):
  """tpu/tpu_ops.py:75-91: sum over replicas / group size (list of tensors)."""
  tot = shards[0].clone()
  for s in shards[1:]:
    tot = tot + s
  return tot / float(len(shards))


def cross_replica_moments(shards4):
  """tpu/tpu_ops.py:94-125 with parallel=True: mean of per-replica means and of
  per-replica mean-of-squares; var = E[x^2] - E[x]^2."""
  means = [s.mean(dim=(0, 1, 2)) for s in shards4]
  msqs = [(s * s).mean(dim=(0, 1, 2)) for s in shards4]
  mean = cross_replica_mean(means)
  msq = cross_replica_mean(msqs)
  return mean, msq - mean * mean


def normalize(x4, mean, var, eps):
  """tf.nn.batch_normalization with scale/offset None — arch_ops.py:306-312."""
  inv = torch.rsqrt(var + eps)
  return x4 * inv + (-mean * inv)


def spectral_sigma(w2d, u, singular_value="left", eps=1e-12):
  """One power iteration — arch_ops.py:503-527.  Returns (sigma, u_new, v).

  u_new and v are detached (stop_gradient, :521-522); sigma keeps grad wrt w.
  """
  with torch.no_grad():
    if singular_value == "left":
      v = l2_normalize(w2d.t() @ u, eps)
      u_new = l2_normalize(w2d @ v, eps)
    else:
      v = l2_normalize(u @ w2d.t(), eps)
      u_new = l2_normalize(v @ w2d, eps)
  if singular_value == "left":
    sigma = (u_new.t() @ w2d) @ v
  else:
    sigma = (v @ w2d) @ u_new.t()
  return sigma.reshape(()), u_new, v


def sigmoid_ce(logits, labels_one):
  """tf.nn.sigmoid_cross_entropy_with_logits: max(x,0)-x*z+log1p(exp(-|x|))."""
  z = 1.0 if labels_one else 0.0
  return torch.clamp(logits, min=0) - logits * z + torch.log1p(torch.exp(-logits.abs()))


def orthogonal_init(rng, shape, gain=1.0):
  """tf.initializers.orthogonal (weights.initializer="orthogonal", arch_ops.py:60-61)."""
  rows = int(np.prod(shape[:-1]))
  cols = int(shape[-1])
  a = rng.standard_normal((max(rows, cols), min(rows, cols)))
  q, r = np.linalg.qr(a)
  q = q * np.sign(np.diag(r))
  if rows < cols:
    q = q.T
  return (gain * q.reshape(shape)).astype(np.float32)


def glorot_normal_init(rng, shape):
  """tf.initializers.glorot_normal: truncated normal, stddev sqrt(2/(fan_in+fan_out))
  (TF divides by .87962566103423978 to correct for truncation)."""
  fan_in, fan_out = shape[0], shape[1]
  std = math.sqrt(2.0 / (fan_in + fan_out)) / .87962566103423978
  a = rng.standard_normal(shape)
  bad = np.abs(a) > 2.0
  while bad.any():
    a[bad] = rng.standard_normal(int(bad.sum()))
    bad = np.abs(a) > 2.0
  return (a * std).astype(np.float32)
