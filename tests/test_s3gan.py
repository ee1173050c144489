"""S3GAN (reference gans/s3gan.py): rotation head, projection discriminator on real-or-predicted labels and the label
predictor on top of ModularGAN — the engine (`gans/s3gan.py`, kernels `row_has_label` / `argmax_one_hot` / `softmax_xent`)
against the oracle restatement (oracle/gan.py S3ganOracle): the new ops alone, then one full cycle per head configuration
with partly unlabelled real examples (label -1): losses, every gradient incl. the heads', step counters.  The same body runs
above the CPU emulator of the ABI (`not gpu`) and on the device (`gpu`), mirroring gans/s3gan_test.py:43-78."""
import numpy as np
import pytest
import torch

from oracle import gan as ogan
from oracle import nets as onets
from tests.gpu_util import compare_grads, make_inputs

NUM_CLASSES = 10


def _pair(batch, **heads):
  from compare_gan_b200 import datasets, gin_lite as gin
  from compare_gan_b200.gans import modular_gan, s3gan  # noqa: F401
  gin.clear_config()
  gin.parse_config("\n".join([
      "G.batch_norm_fn = @conditional_batch_norm", "G.spectral_norm = True", "D.spectral_norm = True",
      "spectral_norm.singular_value = 'auto'", "standardize_batch.use_moving_averages = False", "weights.initializer = 'orthogonal'",
      "loss.fn = @hinge", "penalty.fn = @no_penalty", "tf.train.AdamOptimizer.beta1 = 0.0", "tf.train.AdamOptimizer.beta2 = 0.999",
      "resnet_biggan.Generator.ch = 8", "resnet_biggan.Discriminator.ch = 8", "resnet_biggan.Discriminator.project_y = False",
      "resnet_biggan.Generator.blocks_with_attention = ''", "resnet_biggan.Discriminator.blocks_with_attention = ''"]))
  ds = datasets.ImageDatasetV2("synthetic", 32, 3, NUM_CLASSES, 100)
  params = {"architecture": "resnet_biggan_arch", "z_dim": 120, "lambda": 1.0, "disc_iters": 1, "seed": 0}
  eng = s3gan.S3GAN(dataset=ds, parameters=params, model_dir="/tmp/cgan_s3gan", g_lr=1e-4, d_lr=1e-30, conditional=True, **heads)
  eng.build(batch)
  cfg = onets.Cfg(architecture="resnet_biggan_arch", image_shape=(32, 32, 3), g_bn="conditional_batch_norm", g_sn=True, d_sn=True,
                  sn_singular="auto", use_moving_averages=False, initializer="orthogonal", ch=8, project_y=False,
                  hierarchical_z=True, embed_y=True, num_classes=NUM_CLASSES)
  cfg.g_attention = cfg.d_attention = ""
  okw = dict(heads)
  okw.pop("use_soft_labels", None)
  orc = ogan.S3ganOracle(cfg, loss="hinge", disc_iters=1, g_lr=1e-4, d_lr=1e-30, beta1=0.0, beta2=0.999, conditional=True,
                         z_dim=120, **okw).build(batch)
  state = eng.state_numpy()
  assert sorted(state) == sorted(orc.store.vars), sorted(set(state) ^ set(orc.store.vars))[:6]
  orc.store.load_numpy(state)
  return eng, orc


def _ops():
  from compare_gan_b200 import kernels as K, tape
  rng = np.random.RandomState(0)
  y = np.zeros((6, 5), np.float32)
  y[0, 2] = y[3, 4] = 1.0
  y[4] = [0.1, 0.2, 0.3, 0.2, 0.2]                     # a soft label counts as available, an all-zero row does not
  np.testing.assert_array_equal(K.row_has_label(K.from_numpy(y)).cpu()[:, 0], [1, 0, 0, 1, 1, 0])
  z = rng.randn(7, 5).astype(np.float32)
  z[2, 1] = z[2, 3] = 9.0                               # a tie: the first maximum wins (tf.argmax)
  np.testing.assert_array_equal(K.argmax_one_hot(K.from_numpy(z)).cpu(), np.eye(5, dtype=np.float32)[z.argmax(1)])
  lab = np.eye(5, dtype=np.float32)[rng.randint(0, 5, 7)]
  lab[5] = [0.5, 0.5, 0, 0, 0]
  for w in (None, np.array([1, 0, 1, 1, 0, 1, 1], np.float32), np.zeros(7, np.float32)):
    zd = K.from_numpy(z, req=True)
    loss = K.softmax_xent(zd, K.from_numpy(lab), None if w is None else K.from_numpy(w))
    zt = torch.from_numpy(z).requires_grad_(True)
    ce = -(torch.from_numpy(lab) * torch.log_softmax(zt, -1)).sum(1)
    wt = torch.ones(7) if w is None else torch.from_numpy(w)
    present = float((wt != 0).sum())
    ref = (wt * ce).sum() / present if present else (wt * ce).sum() * 0.0       # tf.losses SUM_BY_NONZERO_WEIGHTS
    ref.backward()
    assert abs(float(loss.cpu()[0]) - float(ref.detach())) < 1e-5
    (gz,) = tape.backward([(loss, K.fill_(K.empty(1), 1.0))], [zd], K.add_grad)
    np.testing.assert_allclose(gz.cpu(), zt.grad.numpy(), rtol=1e-4, atol=1e-6)


def _body():
  _ops()
  configs = [
      dict(self_supervision="rotation", rotated_batch_fraction=2),
      dict(self_supervision="rotation", rotated_batch_fraction=1, project_y=True, use_predictor=True),
      dict(self_supervision="none", rotated_batch_fraction=1, project_y=True, use_predictor=True, use_soft_pred=True, weight_class_loss=0.7),
      dict(self_supervision="none", rotated_batch_fraction=1, project_y=True),
  ]
  for heads in configs:
    eng, orc = _pair(8, **heads)
    imgs, zs, labels, sampled, _ = make_inputs(np.random.RandomState(3), 1, 8, (32, 32, 3), 120, NUM_CLASSES, z_normal=True)
    for lab in labels:
      lab[[1, 4, 6]] = -1                   # unlabelled real examples: one-hot rows of zeros (s3gan.py:113)
    eng.set_inputs(imgs, zs, labels, sampled)
    eng.run_cycle()
    dl, gl = eng.read_losses()
    odl, ogl = orc.cycle(imgs, zs, labels, sampled)
    assert abs(dl[0] - odl[0]) <= 1e-4 * max(1.0, abs(odl[0])) and abs(gl - ogl) <= 1e-4 * max(1.0, abs(ogl)), (heads, dl, odl, gl, ogl)
    names = set(orc.last_d_grads)
    if heads["self_supervision"] == "rotation":
      assert "discriminator_rotation/score_classify/kernel" in names
    if heads.get("project_y"):
      assert "discriminator_projection/kernel" in names
    if heads.get("use_predictor"):
      assert "discriminator_predictor/predictor_linear/kernel" in names and "discriminator_predictor/predictor_linear/bias" in names
    compare_grads(eng, orc, 2e-3, g_tol=5e-2)
    assert eng.global_step == 1 and eng.global_step_disc == 1
  from compare_gan_b200.gans import s3gan
  with pytest.raises(ValueError):
    _pair(8, self_supervision="rotation", rotated_batch_fraction=3)[0].run_cycle()      # 3 does not divide the batch
  with pytest.raises(ValueError):
    _pair(8, self_supervision="none", rotated_batch_fraction=1, use_predictor=True)                             # predictor requires projection


def test_s3gan_cycle_on_the_emulator():
  from tests.abi_emulator import emulated_library
  with emulated_library():
    _body()


@pytest.mark.gpu
def test_s3gan_cycle_gpu():
  from compare_gan_b200 import kernels as K
  K.init(0)
  _body()
