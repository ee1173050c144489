"""N-GPU == 1-GPU equivalence (the analogue of the reference's arch_ops_tpu_test.py:112-133 for the whole step):
run under `torchrun --nproc-per-node N`.  Every rank trains one cycle on its shard of a global batch with NCCL gradient
all-reduce + cross-replica BN moments; rank 0 then repeats the cycle alone on the concatenated batch and compares the
averaged gradients and the updated BN statistics."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def build(batch, arch="resnet_cifar_arch"):
  from compare_gan_b200 import datasets, gin_lite as gin
  from compare_gan_b200.gans import modular_gan
  gin.clear_config()
  gin.parse_config("\n".join([
      "G.batch_norm_fn = @batch_norm", "D.spectral_norm = True", "standardize_batch.decay = 0.9",
      "standardize_batch.epsilon = 1e-5", "loss.fn = @non_saturating", "penalty.fn = @no_penalty",
      "ModularGAN.g_lr = 0.0002", "ModularGAN.d_lr = 1e-30", "ModularGAN.g_optimizer_fn = @tf.train.AdamOptimizer",
      "tf.train.AdamOptimizer.beta1 = 0.5", "tf.train.AdamOptimizer.beta2 = 0.999", "ModularGAN.math_mode = 0"]))
  ds = datasets.ImageDatasetV2("synthetic", 32, 3, None, 100)
  params = {"architecture": arch, "z_dim": 128, "lambda": 1.0, "disc_iters": 1, "seed": 0}
  return modular_gan.ModularGAN(dataset=ds, parameters=params, model_dir="/tmp/x").build(batch)


def main():
  rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
  os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
  dist.init_process_group("nccl", device_id=torch.device("cuda", local))
  from compare_gan_b200 import kernels as K
  from compare_gan_b200.tpu import tpu_ops
  K.init(local)
  per = 4
  rng = np.random.RandomState(0)
  imgs = [rng.rand(per * world, 32, 32, 3).astype(np.float32) for _ in range(2)]
  zs = [rng.uniform(-1, 1, (per * world, 128)).astype(np.float32) for _ in range(2)]
  eng = build(per)
  sl = slice(rank * per, (rank + 1) * per)
  eng.set_inputs([a[sl] for a in imgs], [a[sl] for a in zs])
  eng.run_cycle()
  torch.cuda.synchronize()
  gd, gg = eng.flat_d["grad"].cpu() / world, eng.flat_g["grad"].cpu() / world       # all-reduced sums -> means
  state = eng.state_numpy()
  dist.barrier()
  ok = True
  if rank == 0:
    tpu_ops.force_local(True)
    ref = build(per * world)
    ref.set_inputs(imgs, zs)
    ref.run_cycle()
    torch.cuda.synchronize()
    rd, rg = ref.flat_d["grad"].cpu(), ref.flat_g["grad"].cpu()
    rstate = ref.state_numpy()
    def rel(a, b):
      return float(np.linalg.norm(a.astype(np.float64) - b) / (np.linalg.norm(b) + 1e-30))
    e_d, e_g = rel(gd, rd), rel(gg, rg)
    e_bn = max(rel(state[k], rstate[k]) for k in state if "moving_" in k)
    e_w = max(rel(state[k], rstate[k]) for k in state if k.startswith("generator/") and "kernel" in k and "u_var" not in k)
    print("world %d: rel err D-grad %.2e  G-grad %.2e  BN moving stats %.2e  G weights after Adam %.2e" % (world, e_d, e_g, e_bn, e_w))
    ok = e_d < 1e-4 and e_g < 2e-3 and e_bn < 1e-5
    print("DIST_EQUIVALENCE", "PASS" if ok else "FAIL")
  dist.barrier()
  dist.destroy_process_group()
  sys.exit(0 if ok else 1)


if __name__ == "__main__":
  main()
