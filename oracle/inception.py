"""Oracle (test infrastructure): Inception-v3 (2015 classify_image graph, as run by tfgan.eval.run_inception —
eval_utils.py:165-175) restated on PyTorch-CPU with TF pooling/padding semantics, plus TF's legacy bilinear resize and
the TF-GAN pre-processing.  Weights are supplied by the caller (HWIO dict `inception/<layer>/{kernel,bias}`).
PARITY UNPINNED: the reference's tests replace the Inception graph by a fake (test_utils.py:37-55), and the real frozen
graph is not available offline; topology follows the published Inception-v3 (Szegedy et al. 2015) layer table.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import tf_ops as T


def resize_bilinear_tf(x, oh, ow):
  """tf.image.resize_bilinear, align_corners=False (TF1 kernel): src = dst * (in/out), no half-pixel offset."""
  n, h, w, c = x.shape
  ys = torch.arange(oh, dtype=torch.float32) * (h / oh)
  xs = torch.arange(ow, dtype=torch.float32) * (w / ow)
  y0 = ys.floor().long(); x0 = xs.floor().long()
  y1 = torch.clamp(y0 + 1, max=h - 1); x1 = torch.clamp(x0 + 1, max=w - 1)
  ly = (ys - y0.float()).view(1, oh, 1, 1); lx = (xs - x0.float()).view(1, 1, ow, 1)
  top = x[:, y0][:, :, x0] + (x[:, y0][:, :, x1] - x[:, y0][:, :, x0]) * lx
  bot = x[:, y1][:, :, x0] + (x[:, y1][:, :, x1] - x[:, y1][:, :, x0]) * lx
  return top + (bot - top) * ly


def preprocess(images01):
  """eval_utils.py:157-175: images*255, bilinear resize to 299x299, (x-128)/128."""
  x = resize_bilinear_tf(torch.as_tensor(images01), 299, 299)
  return (x * 255.0 - 128.0) / 128.0


def _conv(x, w, name, stride=1, padding="SAME"):
  k = w["inception/%s/kernel" % name]; b = w["inception/%s/bias" % name]
  k = torch.as_tensor(k); b = torch.as_tensor(b)
  if padding == "SAME":
    y = T.conv2d_same(x, k, stride)
  else:
    y = F.conv2d(x.permute(0, 3, 1, 2), k.permute(3, 2, 0, 1), stride=stride).permute(0, 2, 3, 1)
  return torch.relu(y + b)


def _pool(x, mode, k, s, padding):
  xn = x.permute(0, 3, 1, 2)
  if padding == "VALID":
    y = F.max_pool2d(xn, k, s) if mode == "max" else F.avg_pool2d(xn, k, s)
  else:
    _, pt, pb = T._same_pads(x.shape[1], k, s)
    _, pl, pr = T._same_pads(x.shape[2], k, s)
    if mode == "max":
      y = F.max_pool2d(F.pad(xn, (pl, pr, pt, pb), value=float("-inf")), k, s)
    else:   # TF avg_pool SAME divides by the number of VALID cells
      num = F.avg_pool2d(F.pad(xn, (pl, pr, pt, pb)), k, s, divisor_override=1)
      den = F.avg_pool2d(F.pad(torch.ones_like(xn[:, :1]), (pl, pr, pt, pb)), k, s, divisor_override=1)
      y = num / den
  return y.permute(0, 2, 3, 1)


def _a(x, w, p):
  b1 = _conv(x, w, p + "/b1x1")
  b5 = _conv(_conv(x, w, p + "/b5x5_1"), w, p + "/b5x5_2")
  b3 = _conv(_conv(_conv(x, w, p + "/b3x3dbl_1"), w, p + "/b3x3dbl_2"), w, p + "/b3x3dbl_3")
  bp = _conv(_pool(x, "avg", 3, 1, "SAME"), w, p + "/bpool")
  return torch.cat([b1, b5, b3, bp], 3)


def _b(x, w, p):
  b3 = _conv(x, w, p + "/b3x3", 2, "VALID")
  bd = _conv(_conv(_conv(x, w, p + "/b3x3dbl_1"), w, p + "/b3x3dbl_2"), w, p + "/b3x3dbl_3", 2, "VALID")
  return torch.cat([b3, bd, _pool(x, "max", 3, 2, "VALID")], 3)


def _c(x, w, p):
  b1 = _conv(x, w, p + "/b1x1")
  b7 = _conv(_conv(_conv(x, w, p + "/b7x7_1"), w, p + "/b7x7_2"), w, p + "/b7x7_3")
  bd = x
  for i in range(1, 6):
    bd = _conv(bd, w, p + "/b7x7dbl_%d" % i)
  bp = _conv(_pool(x, "avg", 3, 1, "SAME"), w, p + "/bpool")
  return torch.cat([b1, b7, bd, bp], 3)


def _d(x, w, p):
  b3 = _conv(_conv(x, w, p + "/b3x3_1"), w, p + "/b3x3_2", 2, "VALID")
  b7 = x
  for i in range(1, 4):
    b7 = _conv(b7, w, p + "/b7x7x3_%d" % i)
  b7 = _conv(b7, w, p + "/b7x7x3_4", 2, "VALID")
  return torch.cat([b3, b7, _pool(x, "max", 3, 2, "VALID")], 3)


def _e(x, w, p, pool_mode):
  b1 = _conv(x, w, p + "/b1x1")
  t = _conv(x, w, p + "/b3x3_1")
  b3 = torch.cat([_conv(t, w, p + "/b3x3_2a"), _conv(t, w, p + "/b3x3_2b")], 3)
  t = _conv(_conv(x, w, p + "/b3x3dbl_1"), w, p + "/b3x3dbl_2")
  bd = torch.cat([_conv(t, w, p + "/b3x3dbl_3a"), _conv(t, w, p + "/b3x3dbl_3b")], 3)
  bp = _conv(_pool(x, pool_mode, 3, 1, "SAME"), w, p + "/bpool")
  return torch.cat([b1, b3, bd, bp], 3)


def inception_v3(x, w):
  """x: [N,299,299,3] in [-1,1].  Returns (pool_3 [N,2048], logits [N,1008])."""
  with torch.no_grad():
    x = torch.as_tensor(x)
    x = _conv(x, w, "conv_1a", 2, "VALID")
    x = _conv(x, w, "conv_2a", 1, "VALID")
    x = _conv(x, w, "conv_2b")
    x = _pool(x, "max", 3, 2, "VALID")
    x = _conv(x, w, "conv_3b")
    x = _conv(x, w, "conv_4a", 1, "VALID")
    x = _pool(x, "max", 3, 2, "VALID")
    x = _a(x, w, "mixed"); x = _a(x, w, "mixed_1"); x = _a(x, w, "mixed_2")
    x = _b(x, w, "mixed_3")
    for p in ("mixed_4", "mixed_5", "mixed_6", "mixed_7"):
      x = _c(x, w, p)
    x = _d(x, w, "mixed_8")
    x = _e(x, w, "mixed_9", "avg")
    x = _e(x, w, "mixed_10", "max")
    pool = x.mean(dim=(1, 2))
    logits = pool @ torch.as_tensor(w["inception/logits/kernel"]) + torch.as_tensor(w["inception/logits/bias"])
    return pool, logits
