"""Inception-v3 forward of one batch of 64 (math_mode 1) for an ncu launch list."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from compare_gan_b200 import kernels as K, inception
K.init(0)
K.set_math_mode(1)
net = inception.InceptionV3()
x = K.from_numpy(np.random.RandomState(0).rand(64, 32, 32, 3).astype(np.float32))
for _ in range(2):
  y = K.resize_bilinear(x, 299, 299, inception_scale=True)
  net(y)
torch.cuda.synchronize()
print("done")
