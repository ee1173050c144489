"""Pin the CPU oracle against every golden the reference's own tests hold for the
hot path.  Runs on CPU."""
import json
import os

import numpy as np
import torch

from oracle import gan, metrics, nets
from oracle import tf_ops as T

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_goldens.json")))
V = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "resnet_cifar_variables.json")))


def test_batch_norm_golden():
  # architectures/arch_ops_test.py:29-61
  x = torch.tensor(G["bn_input"]["x"], dtype=torch.float32)
  assert list(x.shape) == [4, 2, 1, 3]
  store, cfg = nets.VarStore(), nets.Cfg()
  y = nets.batch_norm(store, cfg, x, True)
  np.testing.assert_allclose(y.detach().numpy(), np.array(G["bn_expected"]["y"], np.float32),
                             rtol=1e-6, atol=1e-6)


def test_cross_replica_batch_norm_equals_single():
  # architectures/arch_ops_tpu_test.py:112-133: 2-replica sync BN == the single-device golden
  x = torch.tensor(G["bn_input"]["x"], dtype=torch.float32)
  store, cfg = nets.VarStore(), nets.Cfg(bn_replicas=2)
  y = nets.batch_norm(store, cfg, x, True)
  np.testing.assert_allclose(y.detach().numpy(), np.array(G["bn_expected"]["y"], np.float32),
                             rtol=1e-6, atol=1e-6)


def test_cross_replica_mean_golden():
  # tpu/tpu_ops_test.py:79-83
  inp = np.array(G["cross_replica_mean"]["inputs"], np.float32)
  out = T.cross_replica_mean([torch.from_numpy(inp[0]), torch.from_numpy(inp[1])])
  np.testing.assert_allclose(out.numpy(), G["cross_replica_mean"]["expected"], atol=1e-6)


def test_accumulated_moments_state_machine():
  # architectures/arch_ops_test.py:63-132 driven through standardize_batch's accumulator path.
  cfg = nets.Cfg(use_moving_averages=False, bn_eps=0.0)
  store = nets.VarStore()

  def feed(mean, var, training):
    # a 2-sample batch with exactly the requested per-channel mean / biased variance
    m, v = np.array(mean, np.float32), np.array(var, np.float32)
    x = torch.from_numpy(np.stack([m - np.sqrt(v), m + np.sqrt(v)]).reshape(2, 1, 1, 2))
    y = nets.standardize_batch(store, cfg, x, training)
    # recover the (mean, var) the layer used from its output: y = (x-mean)/sqrt(var)
    x0, x1 = x[0, 0, 0], x[1, 0, 0]
    y0, y1 = y[0, 0, 0], y[1, 0, 0]
    inv = (y1 - y0) / (x1 - x0)
    used_mean = x0 - y0 / inv
    return used_mean.numpy(), (1.0 / inv ** 2).numpy()

  # training: batch moments pass through, accumulators untouched (:63-89)
  for mean, var in G["accu_moments"]["feeds"][:2]:
    m, v = feed(mean, var, True)
    np.testing.assert_allclose(m, mean, rtol=1e-5)
    np.testing.assert_allclose(v, var, rtol=1e-4)
  assert float(store.vars["accu/accu_mean"].sum()) == 0.0
  assert abs(float(store.vars["accu/accu_counter"])) < 1e-6
  # eval with update_accus=1, then 0 (:91-132)
  store.vars["accu/update_accus"].fill_(1.0)
  feeds, outs = G["accu_moments"]["feeds"], G["accu_moments"]["eval_outputs"]
  for i in range(2):
    m, v = feed(feeds[i][0], feeds[i][1], False)
    np.testing.assert_allclose(m, outs[i][0], rtol=1e-4)
    np.testing.assert_allclose(v, outs[i][1], rtol=1e-4)
  store.vars["accu/update_accus"].fill_(0.0)
  m, v = feed(feeds[2][0], feeds[2][1], False)
  np.testing.assert_allclose(m, outs[2][0], rtol=1e-4)
  np.testing.assert_allclose(v, outs[2][1], rtol=1e-4)
  aa = G["accu_moments"]["accu_after"]
  np.testing.assert_allclose(store.vars["accu/accu_mean"].numpy(), aa["mean"], rtol=1e-5)
  np.testing.assert_allclose(store.vars["accu/accu_variance"].numpy(), aa["variance"], rtol=1e-5)
  np.testing.assert_allclose(float(store.vars["accu/accu_counter"]), aa["counter"], rtol=1e-6)


def test_fid_golden():
  # metrics/fid_score_test.py:31-40
  real = np.ones((100, 2)); real[:50, 0] = 2
  gen = np.ones((100, 2)) * 9; gen[50:, 0] = 2
  assert abs(metrics.compute_fid_from_activations(real, gen) - G["fid"]["value"]) < G["fid"]["tol"]


def _names(cfg, which, trainable):
  store = nets.VarStore()
  z = torch.zeros(8, 128)
  with torch.no_grad():
    if which == "g":
      out = nets.generator(store, cfg, z, None, True)
      assert list(out.shape) == [8, 32, 32, 3]
    else:
      nets.discriminator(store, cfg, torch.zeros(8, 32, 32, 3), None, True)
  src = store.trainable if trainable else store.vars
  return [[k, list(v.shape)] for k, v in src.items()]


def test_resnet_cifar_variable_lists():
  # architectures/resnet_norm_test.py:39-63, 78-106, 124-162, 326-361
  assert _names(nets.Cfg(g_bn=None), "g", True) == V["g_default"]
  assert _names(nets.Cfg(g_bn=None), "d", True) == V["d_default"]
  assert _names(nets.Cfg(g_bn="batch_norm"), "g", True) == V["g_batch_norm"]
  assert _names(nets.Cfg(g_bn=None, g_sn=True), "g", False) == V["g_spectral_norm_global"]


def test_biggan128_parameter_counts():
  # architectures/resnet_biggan_test.py:112-154
  cfg = nets.Cfg(architecture="resnet_biggan_arch", image_shape=(128, 128, 3),
                 g_bn="conditional_batch_norm", g_sn=True, d_sn=True, sn_singular="auto",
                 hierarchical_z=True, embed_y=True, project_y=True, num_classes=1000,
                 use_moving_averages=False)
  store = nets.VarStore()
  with torch.no_grad():
    z = torch.zeros(2, 120)
    y = torch.zeros(2, 1000); y[:, 1] = 1
    x = nets.generator(store, cfg, z, y, True)
    assert list(x.shape) == [2, 128, 128, 3]
    assert float(x.min()) >= 0.0 and float(x.max()) <= 1.0   # architectures_test.py:51-57
    d, _, _ = nets.discriminator(store, cfg, x, y, True)
    assert float(d.min()) >= 0.0 and float(d.max()) <= 1.0
  ng = sum(v.numel() for k, v in store.trainable.items() if k.startswith("generator/"))
  nd = sum(v.numel() for k, v in store.trainable.items() if k.startswith("discriminator/"))
  assert ng == G["biggan128_params"]["generator"]
  assert nd == G["biggan128_params"]["discriminator"]
  # resnet_biggan_test.py: CBN input width 148, embed_y [1000,128], 1x1 shortcuts, no shortcut in D B6
  assert list(store.vars["generator/B1/bn1/condition/gamma/kernel"].shape) == [148, 1536]
  assert list(store.vars["generator/embed_y/kernel"].shape) == [1000, 128]
  assert list(store.vars["generator/B1/up_conv_shortcut/kernel"].shape) == [1, 1, 1536, 1536]
  assert not any(k.startswith("discriminator/B6/") and "shortcut" in k for k in store.vars)


def test_step_counters_and_cycle_runs():
  # gans/modular_gan_test.py:142-177: global_step_disc == global_step * disc_iters
  cfg = nets.Cfg(d_sn=True, bn_decay=0.9, bn_eps=1e-5)
  k = 2
  o = gan.GanOracle(cfg, disc_iters=k).build(2)
  rng = np.random.RandomState(0)
  for _ in range(2):
    imgs = [rng.rand(2, 32, 32, 3).astype(np.float32) for _ in range(k + 1)]
    zs = [rng.uniform(-1, 1, (2, 128)).astype(np.float32) for _ in range(k + 1)]
    dl, gl = o.cycle(imgs, zs)
    assert len(dl) == k and np.isfinite(gl)
  assert o.global_step == 2 and o.global_step_disc == o.global_step * k


def test_wgangp_matches_finite_difference():
  # penalty (gans/penalty_lib.py:59-82) has no golden: check the autograd double-backward
  # against a central finite difference of the penalty wrt one D weight.
  torch.manual_seed(0)
  cfg = nets.Cfg(architecture="resnet5_arch", image_shape=(16, 16, 3), g_bn=None)
  store = nets.VarStore(1)
  x = torch.rand(2, 16, 16, 3).double()
  xf = torch.rand(2, 16, 16, 3).double()
  alpha = torch.rand(2, 1, 1, 1).double()
  with torch.no_grad():
    nets.discriminator(store, cfg, x.float(), None, True)
  for k in list(store.vars):   # run this check in float64
    v = store.vars[k].detach().double() * 5.0
    v.requires_grad_(k in store.trainable)
    store.vars[k] = v
    if k in store.trainable:
      store.trainable[k] = v
  w = store.vars["discriminator/B1/same_conv1/kernel"]
  p = gan.wgangp_penalty(store, cfg, x, xf, None, True, alpha)
  gw = torch.autograd.grad(p, w)[0]
  idx = (1, 1, 3, 5)
  eps = 1e-5
  with torch.no_grad():
    w[idx] += eps
  pp = float(gan.wgangp_penalty(store, cfg, x, xf, None, True, alpha))
  with torch.no_grad():
    w[idx] -= 2 * eps
  pm = float(gan.wgangp_penalty(store, cfg, x, xf, None, True, alpha))
  fd = (pp - pm) / (2 * eps)
  assert abs(fd - float(gw[idx])) <= 1e-5 * max(1.0, abs(fd)), (fd, float(gw[idx]))


def test_conv_transpose_is_adjoint_of_conv():
  # deconv2d semantics (arch_ops.py:588-589): <conv(x,w), y> == <x, conv_transpose(y,w)>
  rng = np.random.RandomState(0)
  for k, s, n in [(4, 2, 8), (3, 1, 6), (5, 2, 8), (5, 2, 7)]:
    x = torch.from_numpy(rng.randn(2, n, n, 3)).double()
    w = torch.from_numpy(rng.randn(k, k, 3, 5)).double()     # conv: 3 -> 5
    y = torch.from_numpy(rng.randn(2, -(-n // s), -(-n // s), 5)).double()
    lhs = (T.conv2d_same(x, w, s) * y).sum()
    rhs = (x * T.conv2d_transpose_same(y, w, (n, n), s)).sum()
    assert abs(float(lhs - rhs)) < 1e-9 * max(1.0, abs(float(lhs)))


def _naive_conv2d_same(x, w, s):
  """tf.nn.conv2d(padding="SAME") written out from its definition (TF docs "convolution" / nn_ops padding notes):
  out = ceil(in / s); pad_total = max((out-1)*s + k - in, 0); pad_before = pad_total // 2 (the extra pixel goes AFTER);
  out[b,i,j,o] = sum_{di,dj,c} x[b, s*i+di-pad_top, s*j+dj-pad_left, c] * w[di,dj,c,o]."""
  b, h, wd, c = x.shape
  kh, kw, _, o = w.shape
  oh, ow = -(-h // s), -(-wd // s)
  pt = max((oh - 1) * s + kh - h, 0) // 2
  pl = max((ow - 1) * s + kw - wd, 0) // 2
  y = np.zeros((b, oh, ow, o), np.float64)
  for i in range(oh):
    for j in range(ow):
      for di in range(kh):
        for dj in range(kw):
          r, q = s * i + di - pt, s * j + dj - pl
          if 0 <= r < h and 0 <= q < wd:
            y[:, i, j, :] += x[:, r, q, :].astype(np.float64) @ w[di, dj].astype(np.float64)
  return y


def test_conv2d_same_matches_the_written_out_definition():
  """The oracle's conv2d_same (explicit asymmetric padding + torch conv) against plain loops over TF's definition:
  odd and even sizes, strides 1/2, kernels 1/3/4/5 (the 4x4 and 5x5 stride-2 cases put the extra padding pixel after,
  SURVEY App. A)."""
  rng = np.random.RandomState(0)
  for (h, w_, k, s) in [(5, 5, 3, 1), (6, 4, 3, 2), (7, 7, 5, 2), (8, 8, 4, 2), (5, 6, 1, 1), (9, 9, 5, 1), (4, 4, 4, 2), (7, 5, 3, 2)]:
    x = rng.randn(2, h, w_, 3).astype(np.float32)
    w = rng.randn(k, k, 3, 4).astype(np.float32)
    got = T.conv2d_same(torch.from_numpy(x), torch.from_numpy(w), s).numpy()
    want = _naive_conv2d_same(x, w, s)
    assert got.shape == want.shape, (h, w_, k, s)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5, err_msg=str((h, w_, k, s)))


def test_conv2d_transpose_matches_scatter_definition():
  """tf.nn.conv2d_transpose(..., "SAME") is the gradient of conv2d_same w.r.t. its input: every input pixel scatters
  its value times the kernel into the output window the forward conv would have read (arch_ops.py:588-589)."""
  rng = np.random.RandomState(1)
  for (oh, k, s) in [(8, 4, 2), (8, 5, 2), (6, 3, 1), (10, 4, 2)]:
    ih = -(-oh // s)
    x = rng.randn(2, ih, ih, 3).astype(np.float32)
    w = rng.randn(k, k, 5, 3).astype(np.float32)              # [kh, kw, out_channels, in_channels]
    got = T.conv2d_transpose_same(torch.from_numpy(x), torch.from_numpy(w), (oh, oh), s).numpy()
    pt = max((ih - 1) * s + k - oh, 0) // 2
    want = np.zeros((2, oh, oh, 5), np.float64)
    for i in range(ih):
      for j in range(ih):
        for di in range(k):
          for dj in range(k):
            r, q = s * i + di - pt, s * j + dj - pt
            if 0 <= r < oh and 0 <= q < oh:
              want[:, r, q, :] += x[:, i, j, :].astype(np.float64) @ w[di, dj].astype(np.float64).T
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5, err_msg=str((oh, k, s)))


def test_losses_match_their_mathematical_definitions():
  """loss_lib.py:53-148 against float64 numpy from the textbook forms: -log sigma(x) / -log(1 - sigma(x)) for the
  non-saturating loss (TF evaluates them through the stable max(x,0) - x z + log1p(exp(-|x|)) form), hinge,
  Wasserstein and least squares.  Includes logits of +-30 where the naive fp32 form would overflow."""
  from oracle import gan as ogan
  rng = np.random.RandomState(0)
  real = np.concatenate([rng.randn(30, 1) * 3, [[30.0], [-30.0]]]).astype(np.float32)
  fake = np.concatenate([rng.randn(30, 1) * 3, [[-30.0], [30.0]]]).astype(np.float32)
  r64, f64 = real.astype(np.float64), fake.astype(np.float64)
  sig = lambda v: 1.0 / (1.0 + np.exp(-v))
  want = {
      # -log sigma(x) = log(1 + e^-x), -log(1 - sigma(x)) = log(1 + e^x): evaluated with logaddexp, exact in float64
      "non_saturating": (np.mean(np.logaddexp(0, -r64)) + np.mean(np.logaddexp(0, f64)), np.mean(np.logaddexp(0, -r64)),
                         np.mean(np.logaddexp(0, f64)), np.mean(np.logaddexp(0, -f64))),
      "hinge": (np.mean(np.maximum(0, 1 - r64)) + np.mean(np.maximum(0, 1 + f64)), np.mean(np.maximum(0, 1 - r64)),
                np.mean(np.maximum(0, 1 + f64)), -np.mean(f64)),
      "wasserstein": (-np.mean(r64) + np.mean(f64), -np.mean(r64), np.mean(f64), -np.mean(f64)),
      "least_squares": (0.5 * (np.mean((sig(r64) - 1) ** 2) + np.mean(sig(f64) ** 2)), np.mean((sig(r64) - 1) ** 2),
                        np.mean(sig(f64) ** 2), 0.5 * np.mean((sig(f64) - 1) ** 2)),
  }
  tr, tf_ = torch.from_numpy(real), torch.from_numpy(fake)
  for fn, expect in want.items():
    got = ogan.get_losses(fn, torch.sigmoid(tr), torch.sigmoid(tf_), tr, tf_)
    np.testing.assert_allclose([float(v) for v in got], expect, rtol=2e-6, atol=1e-6, err_msg=fn)


def test_spectral_norm_is_one_power_iteration_and_converges_to_the_top_singular_value():
  """arch_ops.py:503-531: v = normalize(W^T u), u' = normalize(W v), sigma = u'^T W v, W / sigma; iterating the stored u
  drives sigma to the largest singular value (numpy SVD), for both the left and the right variant."""
  rng = np.random.RandomState(0)
  w = rng.randn(12, 7).astype(np.float32)
  smax = np.linalg.svd(w.astype(np.float64), compute_uv=False)[0]
  for mode, ushape in (("left", (12, 1)), ("right", (1, 7))):
    u = rng.randn(*ushape).astype(np.float32)
    u64 = u.astype(np.float64)
    w64 = w.astype(np.float64)
    # one step, written out in float64
    if mode == "left":
      v = w64.T @ u64; v /= np.sqrt((v ** 2).sum() + 1e-12)
      un = w64 @ v; un /= np.sqrt((un ** 2).sum() + 1e-12)
      s_ref = float((un.T @ w64 @ v).item())
    else:
      v = u64 @ w64.T; v /= np.sqrt((v ** 2).sum() + 1e-12)
      un = v @ w64; un /= np.sqrt((un ** 2).sum() + 1e-12)
      s_ref = float((v @ w64 @ un.T).item())
    sigma, u_new, _ = T.spectral_sigma(torch.from_numpy(w), torch.from_numpy(u), mode)
    np.testing.assert_allclose(float(sigma), s_ref, rtol=1e-5)
    np.testing.assert_allclose(u_new.numpy(), un, rtol=1e-4, atol=1e-6)
    ut = torch.from_numpy(u)
    for _ in range(200):
      sigma, ut, _ = T.spectral_sigma(torch.from_numpy(w), ut, mode)
    np.testing.assert_allclose(float(sigma), smax, rtol=1e-4)


def test_unpool_and_pools_match_definitions():
  """resnet_ops.unpool (resnet_ops.py:35-56): value at the even position of each 2x2 cell, zeros elsewhere;
  tf.nn.pool AVG 2x2 stride 2 (resnet_ops.py:131) and the 2x2 max pool of the attention block (arch_ops.py:741)."""
  rng = np.random.RandomState(0)
  x = rng.randn(2, 4, 6, 3).astype(np.float32)
  up = T.unpool(torch.from_numpy(x)).numpy()
  want = np.zeros((2, 8, 12, 3), np.float32)
  want[:, ::2, ::2, :] = x
  np.testing.assert_array_equal(up, want)
  cells = x.reshape(2, 2, 2, 3, 2, 3)
  np.testing.assert_allclose(T.avg_pool2(torch.from_numpy(x)).numpy(), cells.mean(axis=(2, 4)), rtol=1e-6, atol=1e-7)
  np.testing.assert_array_equal(T.max_pool2(torch.from_numpy(x)).numpy(), cells.max(axis=(2, 4)))


def test_non_local_block_matches_written_out_attention():
  """arch_ops.py:709-758 from its definition in float64 numpy: theta = x W_t [hw, c/8]; phi, g = 2x2-max-pooled
  x W_p, x W_g [hw/4, .]; beta = softmax(theta phi^T) over the hw/4 keys; out = x + sigma * (beta g) W_o."""
  from oracle import nets as onets
  cfg = onets.Cfg(architecture="resnet_biggan_arch", image_shape=(32, 32, 3))
  store = onets.VarStore(seed=3)
  rng = np.random.RandomState(0)
  x = rng.randn(2, 4, 4, 16).astype(np.float32)
  with torch.no_grad():
    onets.non_local_block(store, cfg, torch.from_numpy(x), "nl", False)          # creates the variables
    store.vars["nl/sigma"].fill_(0.7)
    for k in ("conv2d_theta", "conv2d_phi", "conv2d_g", "conv2d_attn_g"):
      store.vars["nl/%s/kernel" % k].mul_(20.0)                                   # away from a uniform softmax
    got = onets.non_local_block(store, cfg, torch.from_numpy(x), "nl", False).numpy()
  w = {k: store.vars["nl/%s/kernel" % k].detach().numpy().astype(np.float64)[0, 0] for k in ("conv2d_theta", "conv2d_phi", "conv2d_g", "conv2d_attn_g")}
  x64 = x.astype(np.float64)

  def pool(t):          # [n,4,4,c] -> [n,4,c]: 2x2 max over each cell, row-major cells
    n, h, w_, c = t.shape
    return t.reshape(n, h // 2, 2, w_ // 2, 2, c).max(axis=(2, 4)).reshape(n, -1, c)
  theta = (x64 @ w["conv2d_theta"]).reshape(2, 16, -1)
  phi = pool(x64 @ w["conv2d_phi"])
  g = pool(x64 @ w["conv2d_g"])
  logits = theta @ phi.transpose(0, 2, 1)
  beta = np.exp(logits - logits.max(-1, keepdims=True))
  beta /= beta.sum(-1, keepdims=True)
  want = x64 + 0.7 * ((beta @ g).reshape(2, 4, 4, -1) @ w["conv2d_attn_g"])
  np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5)


def test_tf_adam_matches_the_update_rule():
  """tf.train.AdamOptimizer (modular_gan.py:606-616 via the optimizer fn): t += 1; lr_t = lr sqrt(1-b2^t)/(1-b1^t);
  m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; theta -= lr_t m / (sqrt(v) + eps) — epsilon OUTSIDE the bias-corrected
  root, unlike the paper's form."""
  from collections import OrderedDict
  from oracle import gan as ogan
  rng = np.random.RandomState(0)
  p0 = rng.randn(5, 3).astype(np.float32)
  params = OrderedDict(w=torch.from_numpy(p0.copy()))
  opt = ogan.TFAdam(params, lr=2e-4, beta1=0.5, beta2=0.999)
  theta, m, v = p0.astype(np.float64), 0.0, 0.0
  for t in range(1, 6):
    g = rng.randn(5, 3).astype(np.float32)
    opt.step({"w": torch.from_numpy(g)})
    m = 0.5 * m + 0.5 * g.astype(np.float64)
    v = 0.999 * v + 0.001 * g.astype(np.float64) ** 2
    theta = theta - 2e-4 * np.sqrt(1 - 0.999 ** t) / (1 - 0.5 ** t) * m / (np.sqrt(v) + 1e-8)
    np.testing.assert_allclose(params["w"].numpy(), theta, rtol=1e-5, atol=1e-7, err_msg="step %d" % t)


def test_inception_score_and_kid_match_their_estimators():
  """IS = exp(E_x KL(p(y|x) || p(y))) (tfgan classifier_score_from_logits, inception_score.py:44) and KID = the unbiased
  MMD^2 with k(a,b) = (a.b/d + 1)^3 (kid_score.py:44-149), both as explicit double loops in float64 for a single block
  with equally many real and generated samples (where the reference's `n = r_e - r_s` quirk is immaterial)."""
  rng = np.random.RandomState(0)
  logits = rng.randn(17, 9) * 2
  p = np.exp(logits) / np.exp(logits).sum(1, keepdims=True)
  marg = p.mean(0)
  kl = [sum(p[i, j] * (np.log(p[i, j]) - np.log(marg[j])) for j in range(9)) for i in range(17)]
  np.testing.assert_allclose(metrics.inception_score_from_logits(logits), np.exp(np.mean(kl)), rtol=1e-10)

  real, fake = rng.randn(12, 6), rng.randn(12, 6) + 0.3
  k = lambda a, b: (a @ b / 6.0 + 1.0) ** 3
  m = 12
  rr = sum(k(real[i], real[j]) for i in range(m) for j in range(m) if i != j) / (m * (m - 1))
  gg = sum(k(fake[i], fake[j]) for i in range(m) for j in range(m) if i != j) / (m * (m - 1))
  rg = sum(k(real[i], fake[j]) for i in range(m) for j in range(m)) / (m * m)
  np.testing.assert_allclose(metrics.kid(fake, real), rr + gg - 2 * rg, rtol=1e-10)


def test_biggan_deep128_parameter_counts():
  # architectures/resnet_biggan_deep_test.py:30-60
  cfg = nets.Cfg(architecture="resnet_biggan_deep_arch", image_shape=(128, 128, 3), g_bn="conditional_batch_norm",
                 embed_y=True, project_y=True, num_classes=1000, ch=128)
  store = nets.VarStore()
  with torch.no_grad():
    z = torch.zeros(2, 128)
    y = torch.zeros(2, 1000); y[:, 1] = 1
    x = nets.generator(store, cfg, z, y, True)
    assert list(x.shape) == [2, 128, 128, 3]
    out = nets.discriminator(store, cfg, x, y, True)
    assert len(out) == 3 and list(out[1].shape) == [2, 1]
  ng = sum(v.numel() for k, v in store.trainable.items() if k.startswith("generator/"))
  nd = sum(v.numel() for k, v in store.trainable.items() if k.startswith("discriminator/"))
  assert ng == G["biggan_deep128_params"]["generator"]
  assert nd == G["biggan_deep128_params"]["discriminator"]
  # structure: every BN is conditioned on [z, embed(y)] = 256 wide; the first G block keeps 2048 channels through a
  # 512-wide bottleneck; D grows channels through `add_channels`; attention sits at 64x64 in both networks
  assert list(store.vars["generator/B1/conv1/bn/condition/gamma/kernel"].shape) == [256, 2048]
  assert list(store.vars["generator/B1/conv2/3x3_conv/kernel"].shape) == [3, 3, 512, 512]
  assert list(store.vars["discriminator/B1/shortcut/add_channels/kernel"].shape) == [1, 1, 128, 128]
  assert "generator/non_local_block/sigma" in store.vars and "discriminator/non_local_block/sigma" in store.vars


def test_single_training_step_over_losses_penalties_and_architectures():
  """modular_gan_test.py:65-97 (testSingleTrainingStepArchitectures / Losses / Penalties): one cycle at batch 2 for
  every loss, for WGAN-GP, and for every restated architecture, with finite losses and the step-counter rule."""
  rng = np.random.RandomState(0)

  def one(arch, loss, penalty, conditional=False, **cfg_kw):
    shape = (32, 32, 3)
    cfg = nets.Cfg(architecture=arch, image_shape=shape, num_classes=10 if conditional else 0, **cfg_kw)
    o = gan.GanOracle(cfg, loss=loss, penalty=penalty, lamba=1.0, disc_iters=1, g_lr=2e-4, beta1=0.5, beta2=0.999,
                      conditional=conditional, z_dim=120 if conditional else 128).build(2)
    imgs = [rng.rand(2, *shape).astype(np.float32) for _ in range(2)]
    zs = [rng.uniform(-1, 1, (2, 120 if conditional else 128)).astype(np.float32) for _ in range(2)]
    labels = [rng.randint(0, 10, 2).astype(np.int32) for _ in range(2)] if conditional else None
    alphas = [rng.rand(2, 1, 1, 1).astype(np.float32) for _ in range(2)] if penalty == "wgangp_penalty" else None
    d_losses, g_loss = o.cycle(imgs, zs, labels, labels, alphas)
    assert np.isfinite(g_loss) and all(np.isfinite(v) for v in d_losses), (arch, loss, penalty)
    assert o.global_step == 1 and o.global_step_disc == 1

  for loss in ("non_saturating", "hinge", "wasserstein", "least_squares"):
    one("resnet_cifar_arch", loss, "no_penalty")
  one("resnet_cifar_arch", "hinge", "wgangp_penalty")
  for arch in ("sndcgan_arch", "dcgan_arch", "resnet5_arch"):
    one(arch, "hinge", "no_penalty")
  one("resnet_biggan_arch", "hinge", "no_penalty", conditional=True, g_bn="conditional_batch_norm", g_sn=True, d_sn=True,
      hierarchical_z=True, embed_y=True, project_y=True, ch=8, g_attention="B2", d_attention="B1")
  one("resnet_biggan_deep_arch", "hinge", "no_penalty", conditional=True, g_bn="conditional_batch_norm", g_sn=True,
      d_sn=True, embed_y=True, project_y=True, ch=8)
