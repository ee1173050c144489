"""Eager evaluation batches (resnet_cifar10 generator in inference mode -> bilinear 299x299 -> Inception-v3 -> float64
statistics), math_mode 1, for an ncu launch list of the FID path (--profile-from-start off: the third batch only).
Batch = argv[1] (default 256 = four reference batches of 64 per launch, as eval_gan_lib.evaluate fuses them)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from compare_gan_b200 import eval_gan_lib, eval_utils, kernels as K

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng, ds, options = bench.build_engine("resnet_cifar10", 64)
rng = np.random.RandomState(0)
acc = eval_utils.FeatureAccumulator(keep_features=False)
for it in range(3):
  imgs = eval_gan_lib.generate_batch(eng, B, rng)
  torch.cuda.synchronize()
  n0 = K.lib().launch_count()
  if it == 2:
    torch.cuda.profiler.start()
  pool, logits = eval_utils.inception_transform(imgs)
  acc.add(pool, logits, B)
  torch.cuda.synchronize()
  print("launches per batch (resize + inception + statistics):", K.lib().launch_count() - n0)
torch.cuda.profiler.stop()
print("done")
