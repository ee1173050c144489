"""Two eager evaluation batches of 64 (resnet_cifar10 generator in inference mode -> bilinear 299x299 -> Inception-v3 ->
float64 statistics), math_mode 1, for an ncu launch list of the FID path.  The last ~half of the launches are one batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from compare_gan_b200 import eval_gan_lib, eval_utils, kernels as K

eng, ds, options = bench.build_engine("resnet_cifar10", 64)
rng = np.random.RandomState(0)
acc = eval_utils.FeatureAccumulator(keep_features=False)
for it in range(3):
  imgs = eval_gan_lib.generate_batch(eng, 64, rng)
  torch.cuda.synchronize()
  n0 = K.lib().launch_count()
  if it == 2:
    torch.cuda.profiler.start()
  pool, logits = eval_utils.inception_transform(imgs)
  acc.add(pool, logits, 64)
  torch.cuda.synchronize()
  print("launches per batch (resize + inception + statistics):", K.lib().launch_count() - n0)
torch.cuda.profiler.stop()
print("done")
