"""Launch the dominant kernels of the resnet_cifar10 B=256 cycle in isolation (for `ncu --set full`):
3x3 256->256 conv at 32x32 (G B3 / conv2): forward (conv_tc_kernel), input gradient, filter gradient (wgrad_tc_kernel),
and the D-side 128->128 conv at 32x32 with B=512."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from compare_gan_b200 import kernels as K, tape

K.init(0)
K.set_math_mode(1)
rng = np.random.RandomState(0)
for (b, h, cin, cout) in [(256, 32, 256, 256), (512, 32, 128, 128)]:
  x = K.from_numpy(rng.randn(b, h, h, cin).astype(np.float32), req=True)
  w = K.from_numpy((rng.randn(3, 3, cin, cout) * 0.02).astype(np.float32), req=True)
  bias = K.zeros(cout)
  for _ in range(3):
    y = K.conv2d(x, w, bias)
    g = K.from_numpy(rng.randn(*y.shape).astype(np.float32))
    tape.backward([(y, g)], [x, w], K.add)
  torch.cuda.synchronize()
print("done")
