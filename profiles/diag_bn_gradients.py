"""Diagnostic: gradients of g_loss wrt every BN input/output of G, engine vs fp32/fp64 oracle."""
import numpy as np
import torch
from tests.gpu_util import make_inputs, make_pair
from compare_gan_b200 import kernels as K, tape, variables as V
from oracle import nets as onets

K.init(0)
eng, orc, orc64 = make_pair("resnet_cifar_arch", (32, 32, 3), 4, d_sn=True, disc_iters=1, d_lr=1e-30, with64=True)
rng = np.random.RandomState(0)
imgs, zs, _, _, _ = make_inputs(rng, 1, 4, (32, 32, 3), 128)

# ---- engine: record BN in/out
rec = []
orig_bn = K.bn_train
def bn_rec(x, *a, **k):
  y = orig_bn(x, *a, **k)
  rec.append((x, y))
  return y
K.bn_train = bn_rec
with V.use(eng.store):
  z = K.from_numpy(zs[1]); img = K.from_numpy(imgs[1])
  gen = eng.generator(z, y=None, is_training=True)
  K.bn_train = orig_bn
  eng.create_loss({"images": img, "generated": gen}, None, for_discriminator=False)
  ones = K.fill_(K.empty(1), 1.0)
  wrt = [t for pair in rec for t in pair] + [gen]
  grads = tape.backward([(eng.g_loss, ones)], wrt, K.add)
eg = [g.cpu() for g in grads]
efwd = [t.cpu() for pair in rec for t in pair] + [gen.cpu()]

def run_oracle(o, dtype):
  recs = []
  orig = onets.apply_bn
  def rec_bn(store, cfg, which, x, y, is_training, name, use_sn):
    x.retain_grad()
    out = orig(store, cfg, which, x, y, is_training, name, use_sn)
    out.retain_grad()
    recs.append((x, out))
    return out
  onets.apply_bn = rec_bn
  gen = onets.generator(o.store, o.cfg, torch.as_tensor(zs[1]).to(dtype), None, True)
  onets.apply_bn = orig
  gen.retain_grad()
  _, g_loss = o.create_loss(torch.as_tensor(imgs[1]).to(dtype), gen, None, None, None, for_d=False)
  g_loss.backward()
  return ([t.grad.numpy() for pair in recs for t in pair] + [gen.grad.numpy()],
          [t.detach().numpy() for pair in recs for t in pair] + [gen.detach().numpy()])

o32, f32 = run_oracle(orc, torch.float32)
o64, f64 = run_oracle(orc64, torch.float64)
names = []
for i in range(len(rec)):
  names += ["bn%d.in" % i, "bn%d.out" % i]
names.append("gen")
for nm, a, b32, b64 in zip(names, eg, o32, o64):
  nb = np.linalg.norm(b64)
  print("%-10s shape %-18s |ref| %.3e eng %.2e orc32 %.2e  sum-err eng %.2e orc32 %.2e" % (
      nm, a.shape, nb, np.linalg.norm(a - b64) / nb, np.linalg.norm(b32 - b64) / nb,
      np.linalg.norm(a.reshape(-1, a.shape[-1]).sum(0) - b64.reshape(-1, a.shape[-1]).sum(0)) / (np.linalg.norm(b64.reshape(-1, a.shape[-1]).sum(0)) + 1e-30),
      np.linalg.norm(b32.reshape(-1, a.shape[-1]).sum(0) - b64.reshape(-1, a.shape[-1]).sum(0)) / (np.linalg.norm(b64.reshape(-1, a.shape[-1]).sum(0)) + 1e-30)))

print("forward values")
for nm, a, b32, b64 in zip(names, efwd, f32, f64):
  nb = np.linalg.norm(b64)
  print("%-10s |ref| %.3e eng %.2e orc32 %.2e  maxabs eng %.2e orc32 %.2e  signflips eng %d orc32 %d" % (
      nm, nb, np.linalg.norm(a - b64) / nb, np.linalg.norm(b32 - b64) / nb, np.abs(a - b64).max(), np.abs(b32 - b64).max(),
      int(((a > 0) != (b64 > 0)).sum()), int(((b32 > 0) != (b64 > 0)).sum())))
