"""Diagnostic (not a test; run from the repo root: python -m profiles.diag_gradients): per-tensor gradient errors of the smoke configuration vs fp32 / fp64 oracles."""
import sys
import numpy as np
import torch
from tests.gpu_util import make_inputs, make_pair
from compare_gan_b200 import kernels as K

K.init(0)
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
eng, orc, orc64 = make_pair("resnet_cifar_arch", (32, 32, 3), 4, d_sn=True, disc_iters=1, d_lr=1e-30, with64=True)
rng = np.random.RandomState(seed)
inputs = make_inputs(rng, 1, 4, (32, 32, 3), 128)
eng.set_inputs(*inputs)
eng.run_cycle()
print("losses", eng.read_losses(), orc.cycle(*inputs), orc64.cycle(*inputs))
for prefix, flat, r32, r64 in (("discriminator", eng.flat_d, orc.last_d_grads, orc64.last_d_grads),
                               ("generator", eng.flat_g, orc.last_g_grads, orc64.last_g_grads)):
  g = flat["grad"].cpu().astype(np.float64)
  for name, (off, n) in flat["views"].items():
    a, b64, b32 = g[off:off + n], r64[name].numpy().ravel(), r32[name].numpy().ravel().astype(np.float64)
    nb = np.linalg.norm(b64) + 1e-30
    print("%-55s |ref| %.3e  eng %.2e  orc32 %.2e" % (name, nb, np.linalg.norm(a - b64) / nb, np.linalg.norm(b32 - b64) / nb))
