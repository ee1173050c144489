"""world_size-2 `gloo` tests (CPU) of the data-parallel exchange logic: the only collectives on the path are
all-reduce(sum) of a flat gradient buffer and of BN `[mean, mean-of-squares]`."""
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_goldens.json")))


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, q):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from compare_gan_b200.tape import DT
  from compare_gan_b200.tpu import tpu_ops
  out = {}
  assert tpu_ops.num_replicas() == world
  # (1) reference tpu/tpu_ops_test.py:79-83: each replica feeds one vector, all get the mean
  inp = np.array(G["cross_replica_mean"]["inputs"], np.float32)
  x = DT(torch.from_numpy(inp[rank].copy()))
  tpu_ops.cross_replica_mean(x)
  out["mean"] = x.t.numpy().copy()
  # (2) reference arch_ops_tpu_test.py:112-133: sync-BN over 2 replicas == single-device BN on the whole batch
  xb = np.array(G["bn_input"]["x"], np.float32)          # [4,2,1,3]; replica r holds images 2r, 2r+1
  shard = torch.from_numpy(xb[2 * rank:2 * rank + 2])
  local = torch.cat([shard.mean(dim=(0, 1, 2)), (shard * shard).mean(dim=(0, 1, 2))]).contiguous()
  mean, var = tpu_ops.cross_replica_moments_from_local(DT(local))
  y = (shard - mean) * torch.rsqrt(var + 1e-3)
  out["bn"] = y.numpy().copy()
  # (3) gradient exchange: flat buffer all-reduce(sum) then 1/world inside the optimizer == mean of per-replica grads
  g = DT(torch.full((1000,), float(rank + 1)))
  tpu_ops.cross_replica_sum_(g)
  out["grad"] = float(g.t[0]) / world
  q.put((rank, out))
  dist.barrier()
  dist.destroy_process_group()


def test_two_replica_exchange_logic():
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = dict(q.get(timeout=120) for _ in range(2))
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  exp_bn = np.array(G["bn_expected"]["y"], np.float32)
  for r in range(2):
    np.testing.assert_allclose(res[r]["mean"], G["cross_replica_mean"]["expected"], atol=1e-6)
    np.testing.assert_allclose(res[r]["bn"], exp_bn[2 * r:2 * r + 2], rtol=1e-5, atol=1e-5)
    assert abs(res[r]["grad"] - 1.5) < 1e-6


def _engine_worker(rank, world, port, q):
  """One data-parallel rank of the whole training cycle: the engine's host code above the emulated C-ABI
  (tests/abi_emulator.py), collectives over gloo."""
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  torch.set_num_threads(2)
  import sys
  sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
  from tests.abi_emulator import emulated_library
  from tests.dist_gpu_check import build
  from compare_gan_b200.tpu import tpu_ops
  per = 2
  rng = np.random.RandomState(0)
  imgs = [rng.rand(per * world, 32, 32, 3).astype(np.float32) for _ in range(2)]
  zs = [rng.uniform(-1, 1, (per * world, 128)).astype(np.float32) for _ in range(2)]
  out = {}
  with emulated_library():
    eng = build(per)
    sl = slice(rank * per, (rank + 1) * per)
    eng.set_inputs([a[sl] for a in imgs], [a[sl] for a in zs])
    eng.run_cycle()
    out["d_grad"] = eng.flat_d["grad"].cpu() / world          # all-reduced sums -> means
    out["g_grad"] = eng.flat_g["grad"].cpu() / world
    out["state"] = {k: v for k, v in eng.state_numpy().items() if "moving_" in k or k.endswith("u_var")}
    dist.barrier()
    if rank == 0:                                              # the same cycle alone on the concatenated batch
      tpu_ops.force_local(True)
      ref = build(per * world)
      ref.set_inputs(imgs, zs)
      ref.run_cycle()
      out["ref_d_grad"], out["ref_g_grad"] = ref.flat_d["grad"].cpu(), ref.flat_g["grad"].cpu()
      out["ref_state"] = {k: v for k, v in ref.state_numpy().items() if k in out["state"]}
      tpu_ops.force_local(False)
    # sharded evaluation statistics: every rank streams its shard of the activations into float64
    # (n, sum x, sum x x^T); FeatureAccumulator.finish all-reduces them -> moments of the whole set on every rank
    from compare_gan_b200 import eval_utils, kernels as K
    acts = np.random.RandomState(7).randn(12, 16).astype(np.float32)
    acc = eval_utils.FeatureAccumulator(dim=16, keep_features=False)
    mine = acts[rank::world]
    acc.add(K.from_numpy(mine), K.from_numpy(np.zeros((len(mine), 4), np.float32)), len(mine))
    sample = acc.finish(eval_utils.EvalDataSample())
    out["eval_n"], out["eval_mu"], out["eval_sigma"] = acc.n, sample.moments[0], sample.moments[1]
  q.put((rank, out))
  dist.barrier()
  dist.destroy_process_group()


def test_two_rank_training_cycle_equals_one_rank_on_the_concatenated_batch():
  """SURVEY §8e acceptance, on the CPU: two ranks, each on its shard, with the flat-gradient all-reduce after the D and
  the G update and cross-replica BN moments (forward) / sums (backward), reproduce the single-rank cycle on the whole
  batch — the analogue of arch_ops_tpu_test.py:112-133 for the full step (tests/dist_gpu_check.py is the NCCL twin)."""
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_engine_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = dict(q.get(timeout=600) for _ in range(2))
  for p in procs:
    p.join(60)
    assert p.exitcode == 0

  def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / (np.linalg.norm(b) + 1e-30))
  r0 = res[0]
  np.testing.assert_array_equal(res[0]["d_grad"], res[1]["d_grad"])      # both ranks hold the same reduced gradients
  np.testing.assert_array_equal(res[0]["g_grad"], res[1]["g_grad"])
  assert rel(r0["d_grad"], r0["ref_d_grad"]) < 1e-4 and rel(r0["g_grad"], r0["ref_g_grad"]) < 2e-3
  for k, v in r0["ref_state"].items():
    assert rel(r0["state"][k], v) < 1e-5, k                               # BN moving stats (sync-BN) and SN u vectors
    np.testing.assert_array_equal(res[0]["state"][k], res[1]["state"][k], err_msg=k)
  acts = np.random.RandomState(7).randn(12, 16).astype(np.float32).astype(np.float64)
  for r in (0, 1):
    assert res[r]["eval_n"] == 12
    np.testing.assert_allclose(res[r]["eval_mu"], acts.mean(0), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(res[r]["eval_sigma"], np.cov(acts, rowvar=False, ddof=1), rtol=1e-10, atol=1e-12)
