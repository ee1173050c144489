"""SSGAN (reference gans/ssgan.py): rotation self-supervision on top of ModularGAN — the engine (taped C-ABI ops,
`kernels.rot90` / `kernels.rotation_loss`) against the oracle restatement (oracle/gan.py SsganOracle), one full cycle:
losses, every gradient incl. the rotation head's, post-Adam weights.  The same body runs above the CPU emulator of the
ABI (`not gpu`) and on the device (`gpu`), mirroring gans/ssgan_test.py:41-79."""
import numpy as np
import pytest
import torch

from oracle import gan as ogan
from oracle import nets as onets
from tests.gpu_util import compare_grads, make_inputs


def _pair(batch, rotated, loss="hinge", self_supervision="rotation_gan", conditional=False, num_classes=0):
  from compare_gan_b200 import datasets, gin_lite as gin
  from compare_gan_b200.gans import modular_gan, ssgan  # noqa: F401
  gin.clear_config()
  gin.parse_config("\n".join([
      "G.batch_norm_fn = @batch_norm", "D.spectral_norm = True", "standardize_batch.decay = 0.9",
      "standardize_batch.epsilon = 1e-5", "loss.fn = @%s" % loss, "penalty.fn = @no_penalty",
      "tf.train.AdamOptimizer.beta1 = 0.5", "tf.train.AdamOptimizer.beta2 = 0.999"]))
  ds = datasets.ImageDatasetV2("synthetic", 32, 3, num_classes or None, 100)
  params = {"architecture": "resnet_cifar_arch", "z_dim": 128, "lambda": 1.0, "disc_iters": 1, "seed": 0}
  eng = ssgan.SSGAN(dataset=ds, parameters=params, model_dir="/tmp/cgan_ssgan", g_lr=2e-4, d_lr=1e-30,
                    rotated_batch_size=rotated, self_supervision=self_supervision, conditional=conditional)
  eng.build(batch)
  cfg = onets.Cfg(architecture="resnet_cifar_arch", image_shape=(32, 32, 3), g_bn="batch_norm", d_sn=True, bn_decay=0.9,
                  bn_eps=1e-5, num_classes=num_classes)
  orc = ogan.SsganOracle(cfg, rotated_batch_size=rotated, self_supervision=self_supervision, loss=loss, disc_iters=1,
                         g_lr=2e-4, d_lr=1e-30, beta1=0.5, beta2=0.999, conditional=conditional).build(batch)
  state = eng.state_numpy()
  assert sorted(state) == sorted(orc.store.vars), sorted(set(state) ^ set(orc.store.vars))[:6]
  assert "discriminator_rotation/score_classify/kernel" in state and "discriminator_rotation/score_classify/kernel/u_var" in state
  orc.store.load_numpy(state)
  return eng, orc


def _rot_reference(x):
  """The four rotations as the reference composes them (gans/utils.py:38-49), in numpy."""
  tr = lambda a: a.transpose(0, 2, 1, 3)
  return [x, tr(x)[:, ::-1], x[:, ::-1, ::-1], tr(x[:, ::-1])]


def _body():
  from compare_gan_b200 import kernels as K, tape
  # rot90 and its adjoint
  rng = np.random.RandomState(0)
  x = rng.rand(3, 8, 8, 5).astype(np.float32)
  xd = K.from_numpy(x, req=True)
  for k, want in enumerate(_rot_reference(x)):
    y = K.rot90(xd, k)
    np.testing.assert_array_equal(y.cpu(), want)
    if k:
      gy = rng.rand(3, 8, 8, 5).astype(np.float32)
      (gx,) = tape.backward([(y, K.from_numpy(gy))], [xd], K.add_grad)
      xt = torch.from_numpy(x).requires_grad_(True)
      ogan.rotate_images(xt, (k,)).backward(torch.from_numpy(gy))
      np.testing.assert_array_equal(gx.cpu(), xt.grad.numpy())
  # the rotation loss and its gradient
  z = (rng.randn(8, 4) * 2).astype(np.float32)
  zd = K.from_numpy(z, req=True)
  loss = K.rotation_loss(zd)
  zt = torch.from_numpy(z).requires_grad_(True)
  oh = torch.nn.functional.one_hot(torch.arange(4).repeat_interleave(2), 4).float()
  ref = -(oh * torch.log(torch.softmax(zt, -1) + 1e-10)).sum(1).mean()
  ref.backward()
  assert abs(float(loss.cpu()[0]) - float(ref)) < 1e-5
  (gz,) = tape.backward([(loss, K.fill_(K.empty(1), 1.0))], [zd], K.add_grad)
  np.testing.assert_allclose(gz.cpu(), zt.grad.numpy(), rtol=1e-4, atol=1e-6)
  # one cycle of the whole model, both self-supervision modes
  for mode in ("rotation_gan", "rotation_only"):
    eng, orc = _pair(4, 8, self_supervision=mode)
    inputs = make_inputs(np.random.RandomState(3), 1, 4, (32, 32, 3), 128)
    eng.set_inputs(*inputs)
    eng.run_cycle()
    dl, gl = eng.read_losses()
    odl, ogl = orc.cycle(*inputs)
    assert abs(dl[0] - odl[0]) <= 1e-4 * max(1.0, abs(odl[0])) and abs(gl - ogl) <= 1e-4 * max(1.0, abs(ogl)), (mode, dl, odl, gl, ogl)
    assert "discriminator_rotation/score_classify/kernel" in orc.last_d_grads
    compare_grads(eng, orc, 2e-3, g_tol=5e-2)
    assert eng.global_step == 1 and eng.global_step_disc == 1
  with pytest.raises(ValueError):
    _pair(4, 6)[0].run_cycle()          # rotated_batch_size must be a multiple of 4


def test_ssgan_cycle_on_the_emulator():
  from tests.abi_emulator import emulated_library
  with emulated_library():
    _body()


@pytest.mark.gpu
def test_ssgan_cycle_gpu():
  from compare_gan_b200 import kernels as K
  K.init(0)
  _body()
