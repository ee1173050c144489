"""Network- and step-level parity: engine (CUDA via the C-ABI) vs the CPU oracle with identical weights
and identical fed tensors (z, images, labels, alpha) — the analogue of the reference's
runner_lib_test.py:108-147 / modular_gan_test.py:83-95 — plus CUDA-graph replay == eager."""
import numpy as np
import pytest
import torch

from oracle import nets as onets
from tests.gpu_util import assert_close, compare_states, make_inputs, make_pair, rel_err

pytestmark = pytest.mark.gpu

# math_mode 0 (exact fp32 contraction): activations 1e-4, updated weights after Adam 1e-3 abs-ish
ACT_TOL = 2e-4


def _forward_both(eng, orc, batch, z_dim, num_classes=0, z_normal=False):
  from compare_gan_b200 import kernels as K, tape, variables as V
  rng = np.random.RandomState(11)
  z = (rng.standard_normal((batch, z_dim)) if z_normal else rng.uniform(-1, 1, (batch, z_dim))).astype(np.float32)
  labels = rng.randint(0, num_classes, batch).astype(np.int32) if num_classes else None
  snap = eng.snapshot()
  with V.use(eng.store), tape.no_record():
    y = K.one_hot(tape.DT(torch.from_numpy(labels).cuda()), num_classes) if num_classes else None
    img = eng.generator(K.from_numpy(z), y=y, is_training=True)
    d, logit, h = eng.discriminator(img, y=y, is_training=True)
  with torch.no_grad():
    oy = orc.one_hot(labels) if num_classes else None
    oimg = onets.generator(orc.store, orc.cfg, torch.from_numpy(z), oy, True)
    od, ologit, oh = onets.discriminator(orc.store, orc.cfg, oimg, oy, True)
  assert_close(img.cpu(), oimg.numpy(), ACT_TOL, "generator output")
  assert float(img.cpu().min()) >= 0.0 and float(img.cpu().max()) <= 1.0      # architectures_test.py:51-57
  assert_close(h.cpu(), oh.numpy(), 5e-4, "discriminator features")
  assert_close(logit.cpu(), ologit.numpy(), 1e-3, "discriminator logits")
  assert float(d.cpu().min()) >= 0.0 and float(d.cpu().max()) <= 1.0
  eng.restore(snap)
  orc.store.load_numpy(eng.state_numpy())


def _cycles_both(eng, orc, batch, image_shape, z_dim, k, n_cycles=2, num_classes=0, gp=False, z_normal=False,
                 tol=2e-3):
  rng = np.random.RandomState(5)
  for c in range(n_cycles):
    imgs, zs, labels, sampled, alphas = make_inputs(rng, k, batch, image_shape, z_dim, num_classes, z_normal, gp)
    eng.set_inputs(imgs, zs, labels, sampled, alphas)
    eng.run_cycle()
    dl, gl = eng.read_losses()
    odl, ogl = orc.cycle(imgs, zs, labels, sampled, alphas)
    for a, b in zip(dl, odl):
      assert abs(a - b) <= 1e-3 * max(1.0, abs(b)), ("d_loss", c, dl, odl)
    assert abs(gl - ogl) <= 1e-3 * max(1.0, abs(ogl)), ("g_loss", c, gl, ogl)
  assert eng.global_step == n_cycles and eng.global_step_disc == n_cycles * k     # modular_gan_test.py:175-177
  # Adam's first steps move every weight by ~lr regardless of gradient scale, so tiny gradient differences are
  # amplified for near-zero-gradient weights; compare with a tolerance relative to each tensor's norm.
  return compare_states(eng, orc, tol)


def test_resnet_cifar_forward():
  eng, orc = make_pair("resnet_cifar_arch", (32, 32, 3), 4, d_sn=True)
  _forward_both(eng, orc, 4, 128)


def test_resnet_cifar_cycle_sn_bn():
  # BASELINE config 1/2 structure: resnet_cifar10.gin (NS loss, SN on D, BN in G, disc_iters=5 -> 2 here for time)
  eng, orc = make_pair("resnet_cifar_arch", (32, 32, 3), 4, d_sn=True, disc_iters=2)
  _cycles_both(eng, orc, 4, (32, 32, 3), 128, 2)


def test_resnet_cifar_cycle_hinge_gsn_ema():
  eng, orc = make_pair("resnet_cifar_arch", (32, 32, 3), 4, d_sn=True, g_sn=True, loss="hinge", disc_iters=1,
                       g_use_ema=True, ema_start_step=1)
  _cycles_both(eng, orc, 4, (32, 32, 3), 128, 1)
  ema = eng.ema.cpu()
  for name, (off, n) in eng.flat_g["views"].items():
    assert rel_err(ema[off:off + n], orc.ema[name].numpy().ravel()) <= 2e-3, name


def test_sndcgan_forward_and_cycle():
  # config 3 structure (sndcgan_celebahq128.gin) at 32x32 to keep the CPU oracle fast
  eng, orc = make_pair("sndcgan_arch", (32, 32, 3), 4, d_sn=True, disc_iters=1)
  _forward_both(eng, orc, 4, 128)
  _cycles_both(eng, orc, 4, (32, 32, 3), 128, 1)


def test_resnet5_wgangp_cycle():
  # config 4 structure (resnet_lsun-bedroom128.gin: WGAN-GP, lambda 10, no SN, Adam(0.5,0.9) lr 1e-4) at 64x64
  eng, orc = make_pair("resnet5_arch", (64, 64, 3), 2, loss="wasserstein", penalty="wgangp_penalty", lamba=10.0,
                       disc_iters=2, g_lr=1e-4, beta1=0.5, beta2=0.9)
  _forward_both(eng, orc, 2, 128)
  _cycles_both(eng, orc, 2, (64, 64, 3), 128, 2, gp=True, tol=5e-3)


def test_biggan_forward_and_cycle():
  # config 5 structure (biggan_imagenet128.gin) at 32x32, ch=8: conditional BN, attention in G and D, hinge,
  # SN auto, orthogonal init, projection D, accumulators instead of moving averages, EMA, N(0,1) z
  eb = ["resnet_biggan.Generator.blocks_with_attention = 'B2'", "resnet_biggan.Discriminator.blocks_with_attention = 'B1'"]
  eng, orc = make_pair("resnet_biggan_arch", (32, 32, 3), 4, loss="hinge", disc_iters=2, g_bn="conditional_batch_norm",
                       g_sn=True, d_sn=True, sn_singular="auto", conditional=True, num_classes=10,
                       initializer="orthogonal", use_moving_averages=False, g_lr=1e-4, d_lr=5e-4, beta1=0.0,
                       beta2=0.999, z_dim=120, g_use_ema=True, ema_start_step=0, ch=8, extra_bindings=eb,
                       project_y=True)
  # make the attention gate non-zero so the block matters
  for k in ("generator/non_local_block/sigma", "discriminator/non_local_block/sigma"):
    eng.store.vars[k].t.fill_(0.5)
  orc.store.load_numpy(eng.state_numpy())
  _forward_both(eng, orc, 4, 120, num_classes=10, z_normal=True)
  _cycles_both(eng, orc, 4, (32, 32, 3), 120, 2, num_classes=10, z_normal=True, tol=5e-3)


def test_cuda_graph_replay_equals_eager():
  eng, _ = make_pair("resnet_cifar_arch", (32, 32, 3), 4, d_sn=True, disc_iters=2)
  rng = np.random.RandomState(9)
  batches = [make_inputs(rng, 2, 4, (32, 32, 3), 128) for _ in range(2)]
  snap = eng.snapshot()
  eager = []
  for b in batches:
    eng.set_inputs(*b)
    eng.run_cycle()
    eager.append(eng.read_losses())
  state_eager = eng.state_numpy()
  eng.restore(snap)
  eng.capture(warmup=2)
  n0 = eng_launches()
  for i, b in enumerate(batches):
    eng.set_inputs(*b)
    eng.run_cycle()
    dl, gl = eng.read_losses()
    assert dl == eager[i][0] and gl == eager[i][1], "graph replay must be bit-identical to eager"
  assert eng_launches() == n0, "replay launches no new host-side kernels"
  for k, v in eng.state_numpy().items():
    np.testing.assert_array_equal(v, state_eager[k], err_msg=k)


def eng_launches():
  from compare_gan_b200 import kernels as K
  return K.lib().launch_count()
